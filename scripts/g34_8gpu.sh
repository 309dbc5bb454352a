set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r01_bench_c4_8gpu.json 2> gpurun_out/bench_8gpu.err
grep "^{" gpurun_out/r01_bench_c4_8gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks'])"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -1
timeout 600 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r01_bench_c4shard_4gpu.json 2> gpurun_out/bench_4gpu.err
grep "^{" gpurun_out/r01_bench_c4shard_4gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks'])"
