mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for tw in 1; do GVB_TWIN=$tw timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | tail -3; done
