set -x
mkdir -p gpurun_out
timeout 600 python profiles/run_config3_probit.py --N 20000 --Mt 40000 --iterations 3 > gpurun_out/config3_small.json 2> gpurun_out/config3_small.err; echo rc=$?
tail -c 1500 gpurun_out/config3_small.json; tail -5 gpurun_out/config3_small.err
timeout 900 python profiles/run_config3_probit.py --iterations 4 > gpurun_out/config3_1gpu.json 2> gpurun_out/config3_1gpu.err; echo rc=$?
tail -c 2000 gpurun_out/config3_1gpu.json; tail -5 gpurun_out/config3_1gpu.err
cp /tmp/gvamp_c3_rank0.log gpurun_out/config3_host.log
