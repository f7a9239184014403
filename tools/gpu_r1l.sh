#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
echo "== pytest gpu (exchange)"; timeout 1200 python -m pytest tests -m gpu -x -q -k "exchange" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== multi gpu check ($N GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 2000000 > gpurun_out/multi_gpu_check.log 2>&1; echo "rc=$?"
grep -E "PASS|FAIL" gpurun_out/multi_gpu_check.log | head -40; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multi_gpu_check.log | grep -A25 "Traceback" | head -60
echo "== bench native $N GPUs (fused routing, traced)"
CUCO_B200_EXCHANGE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_${N}gpu_traced.json 2> gpurun_out/bench_native_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
line=[l for l in open("gpurun_out/bench_native_${N}gpu_traced.json").read().splitlines() if l.startswith("{")][-1]
d=json.loads(line)
print({k:d[k] for k in ("value","insert_gops","find_gops","insert_ms","find_ms")}, d["e2e"])
print(json.dumps(d.get("exchange_trace_ms"), indent=1))
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_native_${N}gpu.err | grep -B2 -A25 "Traceback" | head -60
