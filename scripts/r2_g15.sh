#!/bin/bash
# round 2, GPU call 15: randomized differential run of the sweep kernels against the oracle
mkdir -p gpurun_out
( time timeout 1500 python tests/fuzz_gpu.py --cases 400 --seed 1 ) > gpurun_out/r2_g15_fuzz.txt 2>&1
tail -5 gpurun_out/r2_g15_fuzz.txt
