set -x
mkdir -p gpurun_out
timeout 900 python profiles/run_config3_probit.py --iterations 5 > gpurun_out/r01_config3_probit_1gpu.json 2> gpurun_out/config3_1gpu.err; echo rc=$?
tail -c 400 gpurun_out/config3_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_config3_probit_1gpu.json').read().strip().splitlines()[-1]); print(d['s_per_iteration'], d['sweeps'], d['ax_GBps'], d['atx_GBps'], d['corr_x1_truth_local_shard'])"
timeout 900 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_config2_1gpu.json 2> gpurun_out/bench_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench_config2_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['sweeps_per_step'], d['config']['twin_layout'], d['roofline']['per_kernel_GBps'], d['roofline']['sweep_share_of_step'], d['e2e']['value'], d['clocks'])"
