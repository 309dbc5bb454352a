#!/bin/bash
# round 2, GPU call 37: final-code suite, the driver's 1-GPU command, c4shard, ncu --set full of the dual kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g37_tests.txt 2>&1; tail -3 gpurun_out/r2_g37_tests.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g37_bench_c4_1gpu.json 2> gpurun_out/r2_g37_bench_err.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --workload c4shard --no-cpu-baseline > gpurun_out/r2_g37_bench_c4shard.json 2>> gpurun_out/r2_g37_bench_err.txt
GVB_TWIN_STRIPES=1560 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ax_dual_kernel -s 2 -c 2 -o gpurun_out/r2_g37_dual_full python profiles/dual_timing.py --pairs 2 > gpurun_out/r2_g37_ncu.log 2>&1
tail -2 gpurun_out/r2_g37_ncu.log
python - <<'P'
import json
for f in ["gpurun_out/r2_g37_bench_c4_1gpu.json","gpurun_out/r2_g37_bench_c4shard.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"]["sm_mhz"], d["config"]["non_sweep_ms_per_step"])
P
