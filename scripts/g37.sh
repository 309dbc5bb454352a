for d in 0 1 2 3; do echo "GVB_DBG=$d"; GVB_DBG=$d timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | grep "rep 2"; done
