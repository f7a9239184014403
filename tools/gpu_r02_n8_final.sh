#!/bin/bash
# final N-GPU pass: parity under torchrun, two traced weak-scaling runs (push kernel on / off for the small
# lookup blocks), then the driver-style full bench line (weak headline + C4 at 4 B pairs + C5 both ways)
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tests/multi_gpu_check.py 2000000 > gpurun_out/r02_multi_gpu_check_${N}.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "FAIL|MULTI_GPU_CHECK|Error" gpurun_out/r02_multi_gpu_check_${N}.log | head -5
run() {
  tag=$1; shift
  env "$@" CUCO_B200_EXCHANGE_TRACE=1 timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
    > gpurun_out/r02f_${N}gpu_${tag}.json 2> gpurun_out/r02f_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    txt = open('gpurun_out/r02f_${N}gpu_${tag}.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('${tag}', {k: round(d[k], 3) for k in ('value', 'insert_ms', 'find_ms', 'insert_ms_best', 'find_ms_best')})
    t = d.get('exchange_trace_ms', [None])[0]
    print({k: v for k, v in t.items() if 'chunk 0' in k or 'chunk 3' in k or k.endswith('staged') or 'slice 0' in k or 'slice 7' in k})
except Exception as e:
    print('${tag} no bench line:', e)
PY
}
run default
run push_off_apply1 CUCO_B200_PUSH_MIB=0 CUCO_B200_APPLY_STREAMS=1
timeout 1200 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err
echo "bench full rc=$?"; tail -n 2 gpurun_out/r02_bench_${N}gpu_final.err | cut -c1-300
python - <<PY
import json
try:
    txt = open('gpurun_out/r02_bench_${N}gpu_final.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('full', {k: round(d[k], 3) for k in ('value', 'insert_ms', 'find_ms', 'value_median', 'value_best')}, 'e2e', round(d['e2e']['value'], 2), d['parity'])
    print('c4', {k: v for k, v in d.get('c4', {}).items() if k != 'workload'})
    print('c5', d.get('c5'))
except Exception as e:
    print('no full line:', e)
PY
