#!/bin/bash
# round 2, GPU call 48: the two-rank tests on the final code (companion steps in lockstep over two ranks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_g48_tests.txt 2>&1; tail -3 gpurun_out/r2_g48_tests.txt
