#!/bin/bash
# Pipelined (multi-lane) fused exchange: parity in both granularities, then bench lanes 1 vs 2.
set -u
mkdir -p gpurun_out
N=${1:-2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== check, fine-grained exchange"; run 29521 tests/multi_gpu_check.py 2000000 > gpurun_out/check_fine.log 2>&1; echo "rc=$?"
grep -E "FAIL|MULTI_GPU_CHECK" gpurun_out/check_fine.log | head; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/check_fine.log | grep -A22 "Traceback" | head -45
echo "== check, owner-only exchange (1 MiB regions, 20 M keys per rank)"; CUCO_B200_REGION_MIB=1 run 29522 tests/multi_gpu_check.py 20000000 > gpurun_out/check_coarse.log 2>&1; echo "rc=$?"
grep -E "FAIL|MULTI_GPU_CHECK" gpurun_out/check_coarse.log | head; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/check_coarse.log | grep -A22 "Traceback" | head -45
for lanes in 1 2; do
  echo "== bench native $N GPUs, lanes=$lanes"
  CUCO_B200_EXCHANGE_LANES=$lanes run 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lanes${lanes}_${N}gpu.json 2> gpurun_out/bench_lanes${lanes}_${N}gpu.err; echo "rc=$?"
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
