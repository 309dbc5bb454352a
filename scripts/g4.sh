set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 2500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
