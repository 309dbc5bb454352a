#!/bin/bash
# round 2, GPU call 1: full GPU test suite + the new bench protocol on the weak shard and on the single-GPU cut of config 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/r2_g1_smi.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2_g1_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 5 --warmup 3 > gpurun_out/r2_g1_bench_c4shard.json 2> gpurun_out/r2_g1_bench_c4shard.err ) 2>> gpurun_out/r2_g1_pytest.txt
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g1_bench_c4.json 2> gpurun_out/r2_g1_bench_c4.err ) 2>> gpurun_out/r2_g1_pytest.txt
( time timeout 300 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g1_bench_config1.json 2> gpurun_out/r2_g1_bench_config1.err ) 2>> gpurun_out/r2_g1_pytest.txt
tail -5 gpurun_out/r2_g1_pytest.txt
