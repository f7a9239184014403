#!/bin/bash
# 2-GPU pass: 1-GPU bench with the blocked defaults, then the NCCL-routed partitioned table on 2 GPUs.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt; nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== bench native 1 GPU"; timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; echo "rc=$?"
cat gpurun_out/bench_native.json | head -c 3500; tail -3 gpurun_out/bench_native.err
echo "== bench native 2 GPUs (NCCL routing)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_2gpu.json 2> gpurun_out/bench_native_2gpu.err; echo "rc=$?"
cat gpurun_out/bench_native_2gpu.json | head -c 3000; tail -5 gpurun_out/bench_native_2gpu.err
echo "== bench reference 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reference_2gpu.json 2> gpurun_out/bench_reference_2gpu.err; echo "rc=$?"
cat gpurun_out/bench_reference_2gpu.json | head -c 3000; tail -5 gpurun_out/bench_reference_2gpu.err
