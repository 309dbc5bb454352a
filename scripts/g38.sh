mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5
for tw in 1 0; do GVB_TWIN=$tw timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -4; done
