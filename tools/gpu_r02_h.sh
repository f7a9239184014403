#!/bin/bash
# round 2, call H (2 GPUs): staged-exchange tests on one GPU (coarse + fine), multi-GPU parity under torchrun,
# traced weak-scaling bench for fine and coarse mode
N=${1:-2}
mkdir -p gpurun_out
for t in static_multiset__retrieve_test static_map__erase_test static_set__retrieve_all_test static_map__rehash_test; do
  timeout 200 tests/_build/reftests/${t}_native > gpurun_out/r02h_${t}.log 2>&1; echo "$t rc=$?"; tail -n 2 gpurun_out/r02h_${t}.log | cut -c1-200
done
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err; cat gpurun_out/r02_next_rows.jsonl
timeout 900 python -m pytest tests/test_staged_exchange_gpu.py tests/test_baseline_configs_gpu.py tests/test_matches_gpu.py tests/test_parity_gpu.py -x -q > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r02h_pytest.log | cut -c1-250
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tests/multi_gpu_check.py 2000000 > gpurun_out/r02h_multi_gpu_check_${N}.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "FAIL|MULTI_GPU_CHECK|Error" gpurun_out/r02h_multi_gpu_check_${N}.log | head
for mode in fine coarse; do
  CUCO_B200_EXCHANGE_MODE=$mode CUCO_B200_EXCHANGE_TRACE=1 timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline \
    > gpurun_out/r02h_bench_${N}gpu_${mode}.json 2> gpurun_out/r02h_bench_${N}gpu_${mode}.err
  echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    txt = open('gpurun_out/r02h_bench_${N}gpu_${mode}.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print({k: round(d[k], 3) for k in ('value', 'insert_ms', 'find_ms', 'insert_ms_best', 'find_ms_best')}, round(d['e2e']['value'], 2))
    print(json.dumps(d.get('exchange_trace_ms', [None])[0]))
except Exception as e:
    print('no bench line:', e)
PY
done
