set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01c_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 900 python profiles/run_config3_probit.py --iterations 5 > gpurun_out/config3_1gpu.json 2> gpurun_out/config3_1gpu.err; echo rc=$?
tail -c 300 gpurun_out/config3_1gpu.json
