set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python profiles/run_sweeps.py --reps 3 --miss 0.01 2>&1 | tail -4
GVB_MISS=twopass timeout 300 python profiles/run_sweeps.py --reps 2 --miss 0.01 2>&1 | tail -3
timeout 600 python profiles/run_config5.py > gpurun_out/config5_1gpu.json 2> gpurun_out/config5_1gpu.err
tail -c 2000 gpurun_out/config5_1gpu.json; tail -5 gpurun_out/config5_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:miss_sum_kernel -s 1 -c 1 -o gpurun_out/r01_miss_sum_full python profiles/run_sweeps.py --reps 2 --miss 0.01 > gpurun_out/ncu_miss.log 2>&1
tail -3 gpurun_out/ncu_miss.log
