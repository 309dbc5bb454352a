#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_g27_host_mem.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"ax_pair_kernel|ax_tile_kernel|atx_pair_kernel" --csv --log-file gpurun_out/r2_g27_ncu_traffic_160GB.csv python profiles/run_sweeps.py --M 1600000 --reps 2 > gpurun_out/r2_g27_run.txt 2>&1
echo "rc=$?" >> gpurun_out/r2_g27_run.txt
cat gpurun_out/r2_g27_host_mem.txt; tail -5 gpurun_out/r2_g27_run.txt; tail -20 gpurun_out/r2_g27_ncu_traffic_160GB.csv | cut -c1-300
