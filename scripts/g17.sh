set -x
mkdir -p gpurun_out
timeout 900 python profiles/run_config3_probit.py --iterations 5 > gpurun_out/config3_1gpu.json 2> gpurun_out/config3_1gpu.err; echo rc=$?
tail -c 400 gpurun_out/config3_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 profiles/run_config3_probit.py --iterations 5 > gpurun_out/config3_2gpu.json 2> gpurun_out/config3_2gpu.err; echo rc=$?
tail -c 1500 gpurun_out/config3_2gpu.json; tail -3 gpurun_out/config3_2gpu.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
