#!/bin/bash
# round 2, GPU call 11: whole GPU suite on the final kernels (incl. the 105 / 160 GB full-size cases), smoke, default bench (driver's flags), config 5
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_g11_pytest.txt 2>&1
( time python __graft_entry__.py smoke 2>&1 | tail -3 ) >> gpurun_out/r2_g11_pytest.txt 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g11_bench_c4.json 2> gpurun_out/r2_g11_bench_c4.err ) >> gpurun_out/r2_g11_pytest.txt 2>&1
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g11_bench_reference.json 2> gpurun_out/r2_g11_bench_reference.err ) >> gpurun_out/r2_g11_pytest.txt 2>&1
( time timeout 400 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g11_bench_config5.json 2> gpurun_out/r2_g11_bench_config5.err ) >> gpurun_out/r2_g11_pytest.txt 2>&1
( time timeout 400 python bench.py --workload c4shard --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_g11_bench_c4shard.json 2> gpurun_out/r2_g11_bench_c4shard.err ) >> gpurun_out/r2_g11_pytest.txt 2>&1
grep -v "^$\|^user\|^sys" gpurun_out/r2_g11_pytest.txt | tail -14
