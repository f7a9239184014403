#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
CUCO_B200_ROUTING=fused CUCO_B200_EXCHANGE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_fused_${N}gpu.json 2> gpurun_out/bench_native_fused_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
line=[l for l in open("gpurun_out/bench_native_fused_${N}gpu.json").read().splitlines() if l.startswith("{")][-1]
d=json.loads(line)
print({k:d[k] for k in ("value","insert_gops","find_gops","insert_ms","find_ms")}, d["e2e"]["value"])
for r,t in enumerate(d.get("exchange_trace_ms") or []):
    print(r, json.dumps(t))
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_native_fused_${N}gpu.err | grep -B2 -A20 "Traceback" | head -40
