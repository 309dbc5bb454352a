#!/bin/bash
# round 2, GPU call 46: full suite on the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g46_tests.txt 2>&1; tail -6 gpurun_out/r2_g46_tests.txt
