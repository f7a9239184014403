#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (exchange)"; timeout 1200 python -m pytest tests -m gpu -x -q -k "exchange or blocked or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
echo "== bench native 2 GPUs (fused routing)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_2gpu.json 2> gpurun_out/bench_native_2gpu.err; echo "rc=$?"
cat gpurun_out/bench_native_2gpu.json | head -c 3000; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_native_2gpu.err | tail -25
