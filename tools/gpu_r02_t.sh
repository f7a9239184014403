#!/bin/bash
# traced weak-scaling bench only, for a few settings of the exchange (no rebuild needed between them)
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() {
  tag=$1; shift
  env "$@" CUCO_B200_EXCHANGE_TRACE=1 timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
    > gpurun_out/r02t_${N}gpu_${tag}.json 2> gpurun_out/r02t_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    txt = open('gpurun_out/r02t_${N}gpu_${tag}.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('${tag}', {k: round(d[k], 3) for k in ('value', 'insert_ms', 'find_ms', 'insert_ms_best', 'find_ms_best')})
    t = d.get('exchange_trace_ms', [None])[0]
    print({k: v for k, v in t.items() if 'landed' in k or 'returned' in k or k.endswith('staged') or 'applied' in k})
except Exception as e:
    print('${tag} no bench line:', e)
PY
}
run fan4 CUCO_B200_COPY_STREAMS=4
run fan8 CUCO_B200_COPY_STREAMS=8
run fan1 CUCO_B200_COPY_STREAMS=1
run fan4_slices4 CUCO_B200_COPY_STREAMS=4 CUCO_B200_EXCHANGE_SLICES=4
run fan4_lanes2 CUCO_B200_COPY_STREAMS=4 CUCO_B200_EXCHANGE_LANES=2
run fan4_lanes8 CUCO_B200_COPY_STREAMS=4 CUCO_B200_EXCHANGE_LANES=8
