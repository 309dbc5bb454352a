set -x
mkdir -p gpurun_out
GVB_ONSAGER_WARM=1 timeout 900 python -m pytest tests/test_gpu_vamp.py -x -q -k "linear_vamp or config1" 2>&1 | tail -8
GVB_ONSAGER_WARM=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['cg_iters_per_step'], d['config']['final_gamw'])"
