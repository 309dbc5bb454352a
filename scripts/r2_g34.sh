#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vamp.py -m gpu -x -q > gpurun_out/r2_g34_tests.txt 2>&1; tail -15 gpurun_out/r2_g34_tests.txt
