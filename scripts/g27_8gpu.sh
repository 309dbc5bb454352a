set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -1
for n in 8 4 2; do
timeout 600 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
grep "^{" gpurun_out/bench_${n}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu_same_box.json 2> gpurun_out/bench_1gpu_same_box.err
grep "^{" gpurun_out/bench_1gpu_same_box.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['roofline']['per_kernel_GBps'], d['e2e']['value'], d['clocks']['sm_mhz'])"
