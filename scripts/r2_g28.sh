#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vamp.py -m gpu -x -q > gpurun_out/r2_g28_tests.txt 2>&1; tail -3 gpurun_out/r2_g28_tests.txt
python bench.py --gpus 1 --steps 8 --warmup 3 --workload config3 --no-cpu-baseline > gpurun_out/r2_g28_bench_config3.json 2> gpurun_out/r2_g28_bench_err.txt
python - <<'P'
import json
for f in ["gpurun_out/r2_g28_bench_config3.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["per_kernel_GBps"], d["clocks"], d["config"].get("non_sweep_ms_per_step"))
P
