#!/bin/bash
# round 2, GPU call 43: the 4-GPU bench line of the final code
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g43_bench_c4_4gpu.json 2> gpurun_out/r2_g43_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g43_bench_c4_4gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"], d["config"]["non_sweep_ms_per_step"])
P
