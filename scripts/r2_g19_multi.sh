#!/bin/bash
# round 2: strong-scaling point of config 4 at $1 GPUs with the driver's flags (final defaults)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_g19_bench_c4_${N}gpu.json 2> gpurun_out/r2_g19_bench_c4_${N}gpu.err ) > gpurun_out/r2_g19_time_$N.txt 2>&1
if [ "$N" = "8" ]; then
 ( time timeout 900 $TR --master-port 29579 bench.py --gpus 8 --workload config5 --steps 10 --warmup 3 > gpurun_out/r2_g19_bench_config5_8gpu.json 2> gpurun_out/r2_g19_bench_config5_8gpu.err ) >> gpurun_out/r2_g19_time_$N.txt 2>&1
 ( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 ) >> gpurun_out/r2_g19_time_$N.txt 2>&1
fi
grep -v "^$\|^user\|^sys" gpurun_out/r2_g19_time_$N.txt | tail -6
