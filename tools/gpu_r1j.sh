#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
echo "== multi gpu check ($N GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 2000000 > gpurun_out/multi_gpu_check.log 2>&1; echo "rc=$?"
grep -E "PASS|FAIL|Error|error" gpurun_out/multi_gpu_check.log | head -40; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multi_gpu_check.log | grep -A25 "Traceback" | head -60
echo "== bench native $N GPUs (fused routing)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_${N}gpu.json 2> gpurun_out/bench_native_${N}gpu.err; echo "rc=$?"
cat gpurun_out/bench_native_${N}gpu.json | head -c 3000; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_native_${N}gpu.err | grep -B2 -A25 "Traceback" | head -60
