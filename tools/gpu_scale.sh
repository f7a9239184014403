#!/bin/bash
# Scaling pass on N GPUs of one box: parity check, then bench (fused, all_to_all, cuco reference).
set -u
mkdir -p gpurun_out
N=${1:-4}
echo "== multi gpu check ($N GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 2000000 > gpurun_out/multi_gpu_check_${N}.log 2>&1; echo "rc=$?"
grep -E "FAIL|MULTI_GPU_CHECK" gpurun_out/multi_gpu_check_${N}.log | head -20; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multi_gpu_check_${N}.log | grep -A25 "Traceback" | head -50
for arm in "native fused" "native nccl" "reference nccl"; do
  set -- $arm
  echo "== bench $1 $N GPUs ($2 routing)"
  CUCO_B200_ROUTING=$2 CUCO_B200_EXCHANGE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl $1 --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2_${N}gpu.json 2> gpurun_out/bench_$1_$2_${N}gpu.err; echo "rc=$?"
  python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/bench_$1_$2_${N}gpu.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    print({k:d[k] for k in ("value","insert_gops","find_gops","insert_ms","find_ms")}, d["e2e"]["value"])
    print(json.dumps(d.get("exchange_trace_ms")))
except Exception as e:
    print("no result", e)
PY
  grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_$1_$2_${N}gpu.err | grep -B2 -A20 "Traceback" | head -40
done
