set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "twin or smoke or ax" 2>&1 | tail -5
for tw in 0 1; do
  GVB_TWIN=$tw timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -6
done
GVB_TWIN_MAD=1 timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -3
GVB_TILE_VARIANT=1 timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01j.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['twin_layout'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
