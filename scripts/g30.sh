set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vamp.py -x -q 2>&1 | tail -5
GVB_BENCH_KEEP_LOG=gpurun_out/host_log_r01l.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01l.json 2> gpurun_out/bench_r01l.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01l.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['twin_layout'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
