#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_g26_bench_c4_8gpu.json 2> gpurun_out/r2_g26_bench_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g26_bench_c4_4gpu.json 2>> gpurun_out/r2_g26_bench_err.txt
python - <<'P'
import json
for f in ["gpurun_out/r2_g26_bench_c4_8gpu.json","gpurun_out/r2_g26_bench_c4_4gpu.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["e2e"]["files_written_per_step"])
P
