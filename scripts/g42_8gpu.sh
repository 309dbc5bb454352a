mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r01_bench_c4_8gpu.json 2> gpurun_out/bench_8gpu.err
grep "^{" gpurun_out/r01_bench_c4_8gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks'])"
