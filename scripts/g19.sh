set -x
timeout 300 python profiles/run_sweeps.py --reps 3 --N 100000 --M 500000 2>&1 | tail -3
timeout 300 python profiles/run_sweeps.py --reps 3 --N 200000 --M 1000000 2>&1 | tail -3
timeout 300 python profiles/run_sweeps.py --reps 3 --N 10000 --M 20000 2>&1 | tail -3
timeout 300 python profiles/run_sweeps.py --reps 3 --N 400000 --M 1600000 2>&1 | tail -3
timeout 600 python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -c 1500
