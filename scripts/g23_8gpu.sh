set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -2
timeout 600 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 2400 gpurun_out/bench_8gpu.json; tail -2 gpurun_out/bench_8gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 700 gpurun_out/bench_2gpu.json
timeout 600 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 700 gpurun_out/bench_4gpu.json
