#!/bin/bash
# round 2, GPU call 3: scale classes (precision tests) + whole GPU suite
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2_g3_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g3_bench_c4shard.json 2> gpurun_out/r2_g3_bench_c4shard.err ) 2>> gpurun_out/r2_g3_pytest.txt
tail -5 gpurun_out/r2_g3_pytest.txt
