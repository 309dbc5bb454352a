set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -3
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_8gpu.%h.%p.log timeout 600 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 2600 gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.err
grep -h -i "nvls\|Connected all\|via P2P" gpurun_out/nccl_8gpu.*.log | sort | uniq -c | sort -rn | head -8 > gpurun_out/nccl_8gpu_summary.txt; rm -f gpurun_out/nccl_8gpu.*.log; cat gpurun_out/nccl_8gpu_summary.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29513 profiles/run_config5.py > gpurun_out/config5_8gpu.json 2> gpurun_out/config5_8gpu.err
tail -c 1800 gpurun_out/config5_8gpu.json; tail -3 gpurun_out/config5_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 600 gpurun_out/bench_4gpu.json; tail -3 gpurun_out/bench_4gpu.err
