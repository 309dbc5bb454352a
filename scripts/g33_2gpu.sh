set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01_bench_c4shard_2gpu.json 2> gpurun_out/bench_2gpu.err
grep "^{" gpurun_out/r01_bench_c4shard_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks'])"
