set -x
mkdir -p gpurun_out
for k in 0 1 2 3 4; do echo "== MADK $k"; GVB_TILE_MAD=$k timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -2; done
GVB_BENCH_KEEP_LOG=gpurun_out/bench_host_rank0.log timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
tail -c 1500 gpurun_out/bench_r01e.json
