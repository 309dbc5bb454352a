#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dual or twin_layout or ax_atx" > gpurun_out/r2_g33_tests.txt 2>&1; tail -5 gpurun_out/r2_g33_tests.txt
( timeout 300 python profiles/dual_timing.py; timeout 300 python profiles/dual_timing.py --twin 0 ) > gpurun_out/r2_g33_dual_timing.txt 2>&1; cat gpurun_out/r2_g33_dual_timing.txt
