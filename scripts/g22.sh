set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err
tail -c 2300 gpurun_out/bench_r01h.json
