#!/bin/bash
# round 2, GPU call 12 (4 GPUs): config 4 strong-scaling point at 4 GPUs (55 GB per GPU + twin)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_g12_bench_c4_4gpu.json 2> gpurun_out/r2_g12_bench_c4_4gpu.err ) > gpurun_out/r2_g12_time.txt 2>&1
grep real gpurun_out/r2_g12_time.txt
