set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv,noheader
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1700 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_c4shard_1gpu.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/r01_bench_c4shard_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2>> gpurun_out/bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01_ncu_launches_bench_c4shard.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -c 300 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 2 -c 2 -o gpurun_out/r01_tile_kernels_twin -f python profiles/run_sweeps.py --reps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
