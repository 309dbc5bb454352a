#!/bin/bash
# round 2, GPU call 13 (2 GPUs): config 4 strong-scaling point at 2 GPUs with the driver's flags (110 GB per GPU, partial twin)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_g13_bench_c4_2gpu.json 2> gpurun_out/r2_g13_bench_c4_2gpu.err ) > gpurun_out/r2_g13_time.txt 2>&1
grep real gpurun_out/r2_g13_time.txt
