for pf in 0 1 2 4 8; do echo "GVB_BED_PF=$pf"; GVB_BED_PF=$pf timeout 300 python profiles/run_sweeps.py --reps 3 2>&1 | grep "rep 2"; done
