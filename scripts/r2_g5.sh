#!/bin/bash
# round 2, GPU call 5: compute-sanitizer (memcheck, racecheck) on small shapes; launch list + full ncu capture of the big-shard X.v kernels;
# config 3 on one GPU after the simulate() fix; whole GPU suite with the test-only library split
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_g5_pytest.txt 2>&1
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_small.py ) > gpurun_out/r2_g5_memcheck.txt 2>&1
( time timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python profiles/sanitize_small.py ) > gpurun_out/r2_g5_racecheck.txt 2>&1
( time timeout 600 python bench.py --workload config3 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_g5_bench_config3.json 2> gpurun_out/r2_g5_bench_config3.err ) 2>> gpurun_out/r2_g5_pytest.txt
# launch list of the default bench command (ncu serialises and cold-starts every launch: shares, not absolutes)
( time timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g5_ncu_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r2_g5_ncu_bench.log 2>&1 ) 2>> gpurun_out/r2_g5_pytest.txt
tail -3 gpurun_out/r2_g5_memcheck.txt gpurun_out/r2_g5_racecheck.txt gpurun_out/r2_g5_pytest.txt
