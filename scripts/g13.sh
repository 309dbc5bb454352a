set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
tail -c 1200 gpurun_out/bench_r01f.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01f.json 2> gpurun_out/bench_ref_r01f.err
tail -c 1200 gpurun_out/bench_ref_r01f.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 2 -c 2 -o gpurun_out/r01b_tile_full python profiles/run_sweeps.py --reps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01b_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
