#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_g25_tests.txt 2>&1; tail -3 gpurun_out/r2_g25_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_g25_bench_c4_2gpu.json 2> gpurun_out/r2_g25_bench_err.txt
python - <<'P'
import json
for f in ["gpurun_out/r2_g25_bench_c4_2gpu.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["e2e"]["files_written_per_step"])
P
