#!/bin/bash
# round 2, GPU call 17: Onsager solve on the Lanczos projection: whole suite + benches
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2_g17_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g17_bench_c4shard.json 2> gpurun_out/r2_g17_bench_c4shard.err ) >> gpurun_out/r2_g17_pytest.txt 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g17_bench_c4.json 2> gpurun_out/r2_g17_bench_c4.err ) >> gpurun_out/r2_g17_pytest.txt 2>&1
( time timeout 400 python bench.py --workload config3 --steps 8 --warmup 2 --no-cpu-baseline > gpurun_out/r2_g17_bench_config3.json 2> gpurun_out/r2_g17_bench_config3.err ) >> gpurun_out/r2_g17_pytest.txt 2>&1
( time timeout 300 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g17_bench_config1.json 2> gpurun_out/r2_g17_bench_config1.err ) >> gpurun_out/r2_g17_pytest.txt 2>&1
( time timeout 300 python bench.py --workload config2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g17_bench_config2.json 2> gpurun_out/r2_g17_bench_config2.err ) >> gpurun_out/r2_g17_pytest.txt 2>&1
grep -v "^$\|^user\|^sys" gpurun_out/r2_g17_pytest.txt | tail -30
