#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
echo "== bench native $N GPUs (fused routing, traced)"
CUCO_B200_EXCHANGE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_${N}gpu_traced.json 2> gpurun_out/bench_native_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_native_${N}gpu_traced.json"))
print({k:d[k] for k in ("value","insert_gops","find_gops","insert_ms","find_ms")})
print(json.dumps(d.get("exchange_trace_ms"), indent=1))
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_native_${N}gpu.err | grep -B2 -A25 "Traceback" | head -60
