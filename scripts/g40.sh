for tt in 0 1; do echo "GVB_TAB_TMA=$tt"; GVB_TAB_TMA=$tt timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | grep "rep 2"; done
GVB_TAB_TMA=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
