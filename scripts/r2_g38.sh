#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vamp.py -m gpu -x -q > gpurun_out/r2_g38_tests.txt 2>&1; tail -25 gpurun_out/r2_g38_tests.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --workload c4shard --no-cpu-baseline > gpurun_out/r2_g38_bench_c4shard.json 2> gpurun_out/r2_g38_bench_err.txt
tail -3 gpurun_out/r2_g38_bench_err.txt
python - <<'P'
import json
for f in ["gpurun_out/r2_g38_bench_c4shard.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"], d["clocks"]["sm_mhz"], d["config"]["sweeps_per_step"], d["config"]["cg_iters_per_step"][:3])
P
