#!/bin/bash
# round 2, GPU call 21: sanitizers on the final kernels; the driver's three commands on the final commit
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_small.py ) > gpurun_out/r2_g21_memcheck.txt 2>&1
( time timeout 1500 compute-sanitizer --tool racecheck --print-limit 40 python profiles/sanitize_small.py ) > gpurun_out/r2_g21_racecheck.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2_g21_pytest.txt 2>&1
( time python __graft_entry__.py smoke 2>&1 | tail -2 ) >> gpurun_out/r2_g21_pytest.txt 2>&1
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g21_bench_reference.json 2> gpurun_out/r2_g21_bench_reference.err ) >> gpurun_out/r2_g21_pytest.txt 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_g21_bench_c4.json 2> gpurun_out/r2_g21_bench_c4.err ) >> gpurun_out/r2_g21_pytest.txt 2>&1
tail -3 gpurun_out/r2_g21_memcheck.txt gpurun_out/r2_g21_racecheck.txt; grep -v "^$\|^user\|^sys" gpurun_out/r2_g21_pytest.txt | tail -10
