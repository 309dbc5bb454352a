#!/bin/bash
# round 2, GPU call 7: sustained-rate tuning of the tile kernel variants (pair mode shapes vs cp.async), twin and gather
mkdir -p gpurun_out
( python profiles/sweep_tuning.py --pairs 40 ) > gpurun_out/r2_g7_tuning_c4shard_twin.txt 2>&1
( python profiles/sweep_tuning.py --pairs 40 --twin 0 --configs "gather_cpasync=GVB_GATHER_TAB:cpasync" "gather_tma11x2=GVB_GATHER_TAB:tma,GVB_PAIR_SHAPE:0" "gather_tma7x3=GVB_GATHER_TAB:tma,GVB_PAIR_SHAPE:1" "gather_tma5x4=GVB_GATHER_TAB:tma,GVB_PAIR_SHAPE:2" ) > gpurun_out/r2_g7_tuning_c4shard_gather.txt 2>&1
( python profiles/sweep_tuning.py --M 1050000 --pairs 12 ) > gpurun_out/r2_g7_tuning_105GB.txt 2>&1
cat gpurun_out/r2_g7_tuning_c4shard_twin.txt gpurun_out/r2_g7_tuning_c4shard_gather.txt gpurun_out/r2_g7_tuning_105GB.txt
