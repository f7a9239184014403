#!/bin/bash
# round 2, call B: the new blocked-insert kernels (tile_route_kernel, stream_mutate_kernel) against the
# round-1 ones, through tools/insert_lab (public C++ API), then the regression suite and the bench
mkdir -p gpurun_out
L=tools/_build/insert_lab
out=gpurun_out/r02_insert_lab.jsonl; : > $out
run() { env "$@" timeout 120 $L 100000000 0.5 5 0 >> $out 2>> gpurun_out/r02_insert_lab.err || echo "{\"failed\": \"$*\"}" >> $out; }
run CUCO_B200_TILE_ROUTE=0 CUCO_B200_STREAM_PROBE=0
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=0
run CUCO_B200_TILE_ROUTE=0 CUCO_B200_STREAM_PROBE=1
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_STREAM_SLOTS=2
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_STREAM_SLOTS=4
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_BLOCKED_PREFETCH=0
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_REGION_MIB=8
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_REGION_MIB=32
run CUCO_B200_BLOCKED=0
# load factor 0.8 and unique keys
for lf in 0.8; do
  env CUCO_B200_TILE_ROUTE=0 CUCO_B200_STREAM_PROBE=0 timeout 120 $L 100000000 $lf 5 0 >> $out 2>> gpurun_out/r02_insert_lab.err
  env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 $lf 5 0 >> $out 2>> gpurun_out/r02_insert_lab.err
done
env CUCO_B200_TILE_ROUTE=0 CUCO_B200_STREAM_PROBE=0 timeout 120 $L 100000000 0.5 5 1 >> $out 2>> gpurun_out/r02_insert_lab.err
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 0.5 5 1 >> $out 2>> gpurun_out/r02_insert_lab.err
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 0.8 5 1 >> $out 2>> gpurun_out/r02_insert_lab.err
echo "insert_lab done"; cat $out
# per-kernel times of the default configuration
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
  --log-file gpurun_out/r02_insert_lab_ncu.csv $L 100000000 0.5 2 0 > /dev/null 2>&1
grep -o '"[a-z_]*kernel[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*"$' gpurun_out/r02_insert_lab_ncu.csv | tail -24 | cut -c1-60,200-
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r02b_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench.json')); print({k:d[k] for k in ('value','insert_gops','find_gops','insert_ms','find_ms')}, d['e2e'])"
