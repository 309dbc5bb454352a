#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the two-rank parity tests and the strong-scaling config-4 point at 2 GPUs (110 GB per GPU, partial twin)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_g4_pytest_multi.txt 2>&1
( time GVB_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_g4_bench_c4_2gpu.json 2> gpurun_out/r2_g4_bench_c4_2gpu.err ) 2>> gpurun_out/r2_g4_pytest_multi.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload config3 --steps 4 --warmup 2 > gpurun_out/r2_g4_bench_config3_2gpu.json 2> gpurun_out/r2_g4_bench_config3_2gpu.err ) 2>> gpurun_out/r2_g4_pytest_multi.txt
tail -12 gpurun_out/r2_g4_pytest_multi.txt
