#!/bin/bash
# round 2, GPU call 44: memcheck over the final code (dual sweep, prepared solve included) and the kernel tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_small.py > gpurun_out/r2_g44_memcheck.txt 2>&1; tail -3 gpurun_out/r2_g44_memcheck.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r2_g44_tests.txt 2>&1; tail -3 gpurun_out/r2_g44_tests.txt
