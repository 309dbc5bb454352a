set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err
tail -c 1800 gpurun_out/bench_r01g.json
