set -x
mkdir -p gpurun_out
GVB_KERNELS=simple timeout 900 python profiles/run_config3_probit.py --iterations 3 > gpurun_out/config3_simple.json 2> gpurun_out/config3_simple.err; echo rc=$?
cp /tmp/gvamp_c3_rank0.log gpurun_out/config3_host_simple.log
tail -c 600 gpurun_out/config3_simple.json
