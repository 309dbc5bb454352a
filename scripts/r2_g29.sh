#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --gpus 1 --steps 8 --warmup 3 --workload config3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g29_bench_config3_$i.json 2> gpurun_out/r2_g29_bench_err.txt
cp /tmp/gvamp_bench_rank0.log gpurun_out/r2_g29_log_$i.txt
done
python - <<'P'
import json
for i in (1,2):
    d=json.loads(open(f"gpurun_out/r2_g29_bench_config3_{i}.json").read().strip().splitlines()[-1])
    print(i, d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("non_sweep_ms_per_step"))
P
grep -n "total time so far\|time for covariates\|time needed to save" gpurun_out/r2_g29_log_1.txt | head -80
