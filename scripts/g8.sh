set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err
tail -c 2500 gpurun_out/bench_r01d.json
timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | tail -4
timeout 300 python profiles/run_sweeps.py --reps 3 --miss 0.01 2>&1 | tail -4
