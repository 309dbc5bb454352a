#!/bin/bash
# round 2, GPU call 49: the probit loop with the riding Lanczos steps: driver tests + config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vamp.py -m gpu -x -q > gpurun_out/r2_g49_tests.txt 2>&1; tail -3 gpurun_out/r2_g49_tests.txt
python bench.py --gpus 1 --steps 8 --warmup 3 --workload config3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g49_bench_config3.json 2> gpurun_out/r2_g49_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g49_bench_config3.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["config"]["non_sweep_ms_per_step"], d["config"]["sweeps_per_step"])
P
