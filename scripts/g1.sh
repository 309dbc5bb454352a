set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "GVB_KERNELS=lut1" "GVB_TILE_VARIANT=0 GVB_TILE_MAD=1" "GVB_TILE_VARIANT=0 GVB_TILE_MAD=0" "GVB_TILE_VARIANT=1 GVB_TILE_MAD=1" "GVB_TILE_VARIANT=1 GVB_TILE_MAD=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python profiles/run_sweeps.py --reps 4 2>&1 | tail -4
done
