#!/bin/bash
# round 2, multi-GPU call: parity of the partitioned table under torchrun, then the N-GPU bench line
# (weak-scaling headline + C4 at size + C5 both ways) with the exchange pipeline traced.
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_${N}.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tests/multi_gpu_check.py 2000000 > gpurun_out/${TAG}_multi_gpu_check_${N}.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "FAIL|MULTI_GPU_CHECK|Error|error" gpurun_out/${TAG}_multi_gpu_check_${N}.log | head -20
CUCO_B200_EXCHANGE_TRACE=1 timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_${N}gpu_traced.json 2> gpurun_out/${TAG}_bench_${N}gpu_traced.err
echo "bench (traced, weak only) rc=$?"; tail -n 3 gpurun_out/${TAG}_bench_${N}gpu_traced.err | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_${N}gpu_traced.json'))
    print({k: d[k] for k in ('value', 'insert_ms', 'find_ms', 'insert_ms_best', 'find_ms_best')}, d['e2e'])
    print(json.dumps(d.get('exchange_trace_ms', [None])[0]))
except Exception as e:
    print('no bench line:', e)
PY
timeout 1500 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "bench (full line) rc=$?"; tail -n 3 gpurun_out/${TAG}_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_${N}gpu.json'))
    print({k: d[k] for k in ('value', 'insert_ms', 'find_ms')})
    print('c4', {k: v for k, v in d.get('c4', {}).items() if k not in ('workload',)})
    print('c5', d.get('c5'))
except Exception as e:
    print('no bench line:', e)
PY
CUCO_B200_ROUTING=fused timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_${N}gpu_fused_r01path.json 2> gpurun_out/${TAG}_bench_${N}gpu_fused_r01path.err
echo "bench (round-1 fused path) rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_${N}gpu_fused_r01path.json'))
    print({k: d[k] for k in ('value', 'insert_ms', 'find_ms')})
except Exception as e:
    print('no bench line:', e)
PY
