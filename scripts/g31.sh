mkdir -p gpurun_out
for rep in 1 2; do for ao in 0 1; do
GVB_ASYNC_OUT=$ao timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('async=$ao', d['ms_per_step'], d['config']['sweeps_per_step'], d['roofline']['sweep_ms'], d['roofline']['sweep_share_of_step'], d['e2e']['ms_per_step'], d['clocks'])"
done; done
