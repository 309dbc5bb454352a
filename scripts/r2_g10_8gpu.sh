#!/bin/bash
# round 2, GPU call 10 (8 GPUs): config 4 itself, strong scaling point at 8 GPUs with the parity check over 8 ranks; config 5 (8 x 105 GB)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_g10_bench_c4_8gpu.json 2> gpurun_out/r2_g10_bench_c4_8gpu.err ) > gpurun_out/r2_g10_time.txt 2>&1
( time timeout 900 $TR --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_g10_bench_c4_8gpu_5steps.json 2> gpurun_out/r2_g10_bench_c4_8gpu_5steps.err ) >> gpurun_out/r2_g10_time.txt 2>&1
( time timeout 900 $TR --master-port 29543 bench.py --gpus 8 --workload config5 --steps 10 --warmup 3 > gpurun_out/r2_g10_bench_config5_8gpu.json 2> gpurun_out/r2_g10_bench_config5_8gpu.err ) >> gpurun_out/r2_g10_time.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 ) >> gpurun_out/r2_g10_time.txt 2>&1
grep -v "^$" gpurun_out/r2_g10_time.txt | tail -12
