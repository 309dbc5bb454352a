set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01i.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
python tests/cmp_config1_warm_start.py 2>&1 | grep "^warm"
