#!/bin/bash
# round 2, GPU call 45: the dual X^T.u: kernel tests, driver tests (file identity with GVB_DUAL_SWEEP=0), its timing, c4shard bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vamp.py -m gpu -x -q > gpurun_out/r2_g45_tests.txt 2>&1; tail -12 gpurun_out/r2_g45_tests.txt
timeout 200 python profiles/dual_timing.py > gpurun_out/r2_g45_dual_timing.txt 2>&1; cat gpurun_out/r2_g45_dual_timing.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --workload c4shard --no-cpu-baseline > gpurun_out/r2_g45_bench_c4shard.json 2> gpurun_out/r2_g45_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g45_bench_c4shard.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"]["sm_mhz"], d["config"]["sweeps_per_step"][:3], d["config"]["ms_per_step_each_host_clock"][:3])
P
