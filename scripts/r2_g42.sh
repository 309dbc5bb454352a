#!/bin/bash
# round 2, GPU call 42: ncu launch list of the default bench command on the final code (shares, not absolutes)
mkdir -p gpurun_out
( time timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g42_ncu_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g42_ncu_bench.log 2>&1 ) 2> gpurun_out/r2_g42_time.txt
tail -3 gpurun_out/r2_g42_time.txt; wc -l gpurun_out/r2_g42_ncu_launches_bench_c4.csv; tail -c 600 gpurun_out/r2_g42_ncu_bench.log
