#!/bin/bash
# prefetch distance of pass 2 of the blocked insert (dev build: kind 1 only), LF 0.5 headline
mkdir -p gpurun_out
for d in 1 2 3 4; do
  CUCO_B200_LIB=$PWD/cucollections_b200/libcuco_b200_dev.so CUCO_B200_PREFETCH_DISTANCE=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-points --no-cpu-baseline \
    > gpurun_out/r02_prefetch_distance_$d.json 2> gpurun_out/r02_prefetch_distance_$d.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02_prefetch_distance_$d.json'))
    print('distance $d', {k: round(d[k], 3) for k in ('value', 'insert_gops', 'insert_ms', 'insert_ms_best', 'find_ms')})
except Exception as e:
    print('distance $d failed', e)
PY
done
for r in 8 32; do
  CUCO_B200_LIB=$PWD/cucollections_b200/libcuco_b200_dev.so CUCO_B200_PREFETCH_DISTANCE=2 CUCO_B200_REGION_MIB=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-points --no-cpu-baseline \
    > gpurun_out/r02_prefetch_distance_2_region$r.json 2> /dev/null
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02_prefetch_distance_2_region$r.json'))
    print('distance 2 region $r MiB', {k: round(d[k], 3) for k in ('value', 'insert_gops', 'insert_ms', 'insert_ms_best')})
except Exception as e:
    print('failed', e)
PY
done
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err; cat gpurun_out/r02_next_rows.jsonl
