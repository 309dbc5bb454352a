#!/bin/bash
mkdir -p gpurun_out
( python profiles/sweep_tuning.py --pairs 40 --twin 0 --configs "gather_cpasync15x2=GVB_GATHER_TAB:cpasync" "gather_tma12x2=GVB_GATHER_TAB:tma,GVB_PAIR_SHAPE:1" "gather_cpasync15x2_b=GVB_GATHER_TAB:cpasync" "gather_tma12x2_b=GVB_GATHER_TAB:tma,GVB_PAIR_SHAPE:1" "gather_mad0=GVB_TILE_MAD:0" "gather_mad2=GVB_TILE_MAD:2" ) > gpurun_out/r2_g22_tuning_gather.txt 2>&1
cat gpurun_out/r2_g22_tuning_gather.txt
