set -x
mkdir -p gpurun_out
for v in "--prior sparse3" "--prior default23 --cg-max-iter 60"; do
timeout 900 python profiles/run_config3_probit.py --iterations 5 $v > gpurun_out/config3_try.json 2> gpurun_out/config3_try.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/config3_try.json').read().strip().splitlines()[-1]); print(d['s_per_iteration'], d['sweeps'], d['corr_x1_truth_local_shard'], d['ax_GBps'], d['atx_GBps'])"
grep -n "^alpha2 =\|^beta2 =\|^tau1 =\|^gam1 =" /tmp/gvamp_c3_rank0.log | tr '\n' ' '; echo
done
