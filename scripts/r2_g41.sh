#!/bin/bash
# round 2, GPU call 41: final-code suite, smoke, the driver's 1-GPU command and its reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g41_tests.txt 2>&1; tail -3 gpurun_out/r2_g41_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g41_bench_c4_1gpu.json 2> gpurun_out/r2_g41_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g41_bench_c4_1gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"], d["config"]["sweeps_per_step"], d["config"]["non_sweep_ms_per_step"], d["cpu_baseline"]["value"])
P
