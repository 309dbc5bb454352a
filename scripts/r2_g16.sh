#!/bin/bash
# round 2, GPU call 16: cached A^T A probe (2 sweeps less per iteration): whole suite + bench on c4shard / c4 cut
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_g16_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g16_bench_c4shard.json 2> gpurun_out/r2_g16_bench_c4shard.err ) >> gpurun_out/r2_g16_pytest.txt 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g16_bench_c4.json 2> gpurun_out/r2_g16_bench_c4.err ) >> gpurun_out/r2_g16_pytest.txt 2>&1
( time timeout 400 python bench.py --workload config3 --steps 8 --warmup 2 --no-cpu-baseline > gpurun_out/r2_g16_bench_config3.json 2> gpurun_out/r2_g16_bench_config3.err ) >> gpurun_out/r2_g16_pytest.txt 2>&1
grep -v "^$\|^user\|^sys" gpurun_out/r2_g16_pytest.txt | tail -12
