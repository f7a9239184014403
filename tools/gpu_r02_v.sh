#!/bin/bash
# N-GPU validation of the final exchange code: staged-exchange tests on one GPU (both transports), parity under
# torchrun, traced weak-scaling bench for a few knob settings
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_staged_exchange_gpu.py -x -q > gpurun_out/r02v_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r02v_pytest.log | cut -c1-250
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tests/multi_gpu_check.py 2000000 > gpurun_out/r02v_multi_gpu_check_${N}.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "FAIL|MULTI_GPU_CHECK|Error" gpurun_out/r02v_multi_gpu_check_${N}.log | head
run() {
  tag=$1; shift
  env "$@" CUCO_B200_EXCHANGE_TRACE=1 timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
    > gpurun_out/r02v_${N}gpu_${tag}.json 2> gpurun_out/r02v_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    txt = open('gpurun_out/r02v_${N}gpu_${tag}.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('${tag}', {k: round(d[k], 3) for k in ('value', 'insert_ms', 'find_ms', 'insert_ms_best', 'find_ms_best')})
    t = d.get('exchange_trace_ms', [None])[0]
    print({k: v for k, v in t.items() if 'chunk 0' in k or 'chunk 3' in k or k.endswith('staged') or 'slice 0' in k or 'slice 7' in k})
except Exception as e:
    print('${tag} no bench line:', e)
PY
}
run default CUCO_B200_PUSH_MIB=8
run push_off CUCO_B200_PUSH_MIB=0
run apply2 CUCO_B200_APPLY_STREAMS=2
run push_all CUCO_B200_PUSH_MIB=64
