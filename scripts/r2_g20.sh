#!/bin/bash
# round 2, GPU call 20: suite after the reduction-grid change; launch list of the c4shard bench (where does the non-sweep time go)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2_g20_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g20_bench_c4shard.json 2> gpurun_out/r2_g20_bench_c4shard.err ) >> gpurun_out/r2_g20_pytest.txt 2>&1
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g20_ncu_launches_c4shard.csv python bench.py --workload c4shard --steps 6 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g20_ncu.log 2>&1 ) >> gpurun_out/r2_g20_pytest.txt 2>&1
grep -v "^$\|^user\|^sys" gpurun_out/r2_g20_pytest.txt | tail -8
