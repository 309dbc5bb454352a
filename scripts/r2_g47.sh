#!/bin/bash
# round 2, GPU call 47: the driver's 1-GPU command on the final code
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g47_bench_c4_1gpu.json 2> gpurun_out/r2_g47_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g47_bench_c4_1gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"], d["config"]["sweeps_per_step"], d["config"]["non_sweep_ms_per_step"], d["cpu_baseline"]["value"], d["config"]["ms_per_step_each_host_clock"][:4])
P
