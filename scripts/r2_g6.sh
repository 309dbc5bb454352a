#!/bin/bash
# round 2, GPU call 6: pair-mode (TMA-staged tables) vs producer-warp cp.async tables; bank-aware missing-genotype list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_g6_pytest.txt 2>&1
( time GVB_TAB=cpasync timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -5 ) >> gpurun_out/r2_g6_pytest.txt 2>&1
for mode in tma cpasync; do
  ( time GVB_TAB=$mode timeout 400 python bench.py --workload c4shard --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g6_bench_c4shard_$mode.json 2> gpurun_out/r2_g6_bench_c4shard_$mode.err ) 2>> gpurun_out/r2_g6_pytest.txt
done
for mode in tma cpasync; do
  ( time GVB_TAB=$mode timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g6_bench_c4_$mode.json 2> gpurun_out/r2_g6_bench_c4_$mode.err ) 2>> gpurun_out/r2_g6_pytest.txt
done
( time timeout 400 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g6_bench_config5.json 2> gpurun_out/r2_g6_bench_config5.err ) 2>> gpurun_out/r2_g6_pytest.txt
grep -v "^$" gpurun_out/r2_g6_pytest.txt | tail -8
