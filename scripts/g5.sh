set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | tail -4
timeout 300 python profiles/run_sweeps.py --reps 3 --miss 0.01 2>&1 | tail -4
GVB_KERNELS=lut1 timeout 300 python profiles/run_sweeps.py --reps 3 --miss 0.01 2>&1 | tail -4
timeout 300 python profiles/run_sweeps.py --reps 3 --N 100000 --M 500000 2>&1 | tail -4
timeout 300 python profiles/run_sweeps.py --reps 3 --N 10000 --M 20000 2>&1 | tail -4
