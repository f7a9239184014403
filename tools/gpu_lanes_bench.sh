#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
for lanes in ${2:-2}; do
  echo "== bench native $N GPUs, lanes=$lanes"
  CUCO_B200_EXCHANGE_LANES=$lanes timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lanes${lanes}_${N}gpu.json 2> gpurun_out/bench_lanes${lanes}_${N}gpu.err; echo "rc=$?"
  python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/bench_lanes${lanes}_${N}gpu.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    print({k:round(d[k],2) for k in ("value","insert_gops","find_gops","insert_ms","find_ms")}, round(d["e2e"]["value"],2))
except Exception as e:
    print("no result", e)
PY
  grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_lanes${lanes}_${N}gpu.err | grep -B2 -A20 "Traceback" | head -40
done
