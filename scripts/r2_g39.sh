#!/bin/bash
# round 2, GPU call 39: two-rank tests and the 2-GPU bench line with the riding Lanczos steps
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_g39_tests.txt 2>&1; tail -3 gpurun_out/r2_g39_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g39_bench_c4_2gpu.json 2> gpurun_out/r2_g39_bench_err.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_g39_bench_c4_2gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_GBps"], d["roofline"]["dual_sweeps"]["n"], d["roofline"]["dual_sweeps"]["ms_each"], d["clocks"]["sm_mhz"], d["config"]["sweeps_per_step"], d["config"]["non_sweep_ms_per_step"])
P
