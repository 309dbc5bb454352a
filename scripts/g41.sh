mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | tail -3
timeout 300 python profiles/run_sweeps.py --reps 3 --miss 0.01 2>&1 | tail -3
for rep in 1 2; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01m.json 2> gpurun_out/bench_r01m.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01m.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['ms_per_step'], d['e2e']['sweeps_per_step'], d['clocks'])"
done
