#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g35_tests.txt 2>&1; tail -8 gpurun_out/r2_g35_tests.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --workload c4shard --no-cpu-baseline > gpurun_out/r2_g35_bench_c4shard.json 2> gpurun_out/r2_g35_bench_err.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g35_bench_c4_1gpu.json 2>> gpurun_out/r2_g35_bench_err.txt
tail -3 gpurun_out/r2_g35_bench_err.txt
python - <<'P'
import json
for f in ["gpurun_out/r2_g35_bench_c4shard.json","gpurun_out/r2_g35_bench_c4_1gpu.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["clocks"]["sm_mhz"], d["config"]["sweeps_per_step"])
P
