#!/bin/bash
# round 2, GPU call 50: the probit loop with the riding Lanczos steps (host library rebuilt): probit driver tests + config 3
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_vamp.py -m gpu -x -q -k probit > gpurun_out/r2_g50_tests.txt 2>&1; tail -2 gpurun_out/r2_g50_tests.txt
timeout 60 python bench.py --gpus 1 --steps 8 --warmup 2 --workload config3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g50_bench_config3.json 2> gpurun_out/r2_g50_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g50_bench_config3.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["per_kernel_GBps"], d["config"]["non_sweep_ms_per_step"], d["config"]["sweeps_per_step"])
P
