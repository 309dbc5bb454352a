#!/bin/bash
# round 2, GPU call 2: GPU test suite (CG on device scalars, partial twin, lane-per-marker gather) + the new bench protocol
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2_g2_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 5 --warmup 3 > gpurun_out/r2_g2_bench_c4shard.json 2> gpurun_out/r2_g2_bench_c4shard.err ) 2>> gpurun_out/r2_g2_pytest.txt
( time GVB_VERBOSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g2_bench_c4.json 2> gpurun_out/r2_g2_bench_c4.err ) 2>> gpurun_out/r2_g2_pytest.txt
( time timeout 300 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g2_bench_config1.json 2> gpurun_out/r2_g2_bench_config1.err ) 2>> gpurun_out/r2_g2_pytest.txt
( time timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g2_bench_config2.json 2> gpurun_out/r2_g2_bench_config2.err ) 2>> gpurun_out/r2_g2_pytest.txt
( time GVB_VERBOSE=1 timeout 400 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g2_bench_config5.json 2> gpurun_out/r2_g2_bench_config5.err ) 2>> gpurun_out/r2_g2_pytest.txt
( time GVB_MISS_SUM=warp timeout 400 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g2_bench_config5_warp.json 2> gpurun_out/r2_g2_bench_config5_warp.err ) 2>> gpurun_out/r2_g2_pytest.txt
tail -5 gpurun_out/r2_g2_pytest.txt
