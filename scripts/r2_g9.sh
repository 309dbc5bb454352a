#!/bin/bash
# round 2, GPU call 9: ncu --set full of the round-2 kernels (pair-mode walks, gathering X.v, bank-aware missing-genotype gather) + racecheck
# of the pair mode + the default bench on the single-GPU cut with the final kernel defaults
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
( $NCU -k regex:pair_kernel -s 2 -c 2 -o gpurun_out/r02_pair_kernels_c4shard -f python profiles/run_sweeps.py --reps 2 ) > gpurun_out/r2_g9_ncu1.log 2>&1
( GVB_TWIN=0 $NCU -k regex:ax_tile_kernel -s 1 -c 1 -o gpurun_out/r02_ax_gather_c4shard -f python profiles/run_sweeps.py --reps 2 ) > gpurun_out/r2_g9_ncu2.log 2>&1
( $NCU -k regex:miss_sum -s 1 -c 1 -o gpurun_out/r02_miss_sum_c4shard_1pct -f python profiles/run_sweeps.py --reps 2 --miss 0.01 ) > gpurun_out/r2_g9_ncu3.log 2>&1
( time timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python profiles/sanitize_small.py ) > gpurun_out/r2_g9_racecheck_pair.txt 2>&1
( time timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_g9_bench_c4.json 2> gpurun_out/r2_g9_bench_c4.err ) > gpurun_out/r2_g9_time.txt 2>&1
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_g9_bench_reference.json 2> gpurun_out/r2_g9_bench_reference.err ) >> gpurun_out/r2_g9_time.txt 2>&1
tail -3 gpurun_out/r2_g9_ncu*.log gpurun_out/r2_g9_racecheck_pair.txt; ls -la gpurun_out/*.ncu-rep
