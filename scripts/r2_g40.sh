#!/bin/bash
# round 2, GPU call 40: the 8-GPU bench line of the final code
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_g40_bench_c4_8gpu.json 2> gpurun_out/r2_g40_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g40_bench_c4_8gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"], d["config"]["sweeps_per_step"], d["config"]["non_sweep_ms_per_step"], d["parity_check"]["small"]["cg_rel"])
P
