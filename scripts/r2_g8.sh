#!/bin/bash
# round 2, GPU call 8: 12 x 2 pair shape vs 11 x 2 vs cp.async (sustained), kernel tests in every shape
mkdir -p gpurun_out
( python profiles/sweep_tuning.py --pairs 40 ) > gpurun_out/r2_g8_tuning_c4shard_twin.txt 2>&1
( python profiles/sweep_tuning.py --pairs 40 --configs "cpasync15x2=GVB_TAB:cpasync" "tma11x2=GVB_TAB:tma,GVB_PAIR_SHAPE:0" "tma12x2=GVB_TAB:tma,GVB_PAIR_SHAPE:1" ) >> gpurun_out/r2_g8_tuning_c4shard_twin.txt 2>&1
for shape in 0 1 2; do
 ( GVB_PAIR_SHAPE=$shape timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3 ) >> gpurun_out/r2_g8_pytest.txt 2>&1
done
cat gpurun_out/r2_g8_tuning_c4shard_twin.txt gpurun_out/r2_g8_pytest.txt
