#!/bin/bash
# round 2, GPU call 14: ncu of the missing-genotype gather after the software pipeline over segments
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:miss_sum -s 1 -c 1 -o gpurun_out/r02_miss_sum_pipelined_c4shard_1pct -f python profiles/run_sweeps.py --reps 2 --miss 0.01 > gpurun_out/r2_g14_ncu.log 2>&1
tail -3 gpurun_out/r2_g14_ncu.log
