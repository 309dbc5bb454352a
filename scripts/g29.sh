set -x
mkdir -p gpurun_out
GVB_BENCH_KEEP_LOG=gpurun_out/host_log_r01k.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01k.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['twin_layout'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
