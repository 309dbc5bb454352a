#!/bin/bash
# round 2, GPU call 36: 8- and 4-GPU bench lines with the dual sweep
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_g36_bench_c4_8gpu.json 2> gpurun_out/r2_g36_bench_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g36_bench_c4_4gpu.json 2>> gpurun_out/r2_g36_bench_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g36_bench_c4_2gpu.json 2>> gpurun_out/r2_g36_bench_err.txt
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
python - <<'P'
import json
for n in (8,4,2):
    d=json.loads(open(f"gpurun_out/r2_g36_bench_c4_{n}gpu.json").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"]["sm_mhz"], d["config"]["non_sweep_ms_per_step"], d["parity_check"]["small"]["cg_rel"])
P
