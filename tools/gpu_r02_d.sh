#!/bin/bash
# round 2, call D: stream_mutate_kernel v2 (converged rows + parked keys)
mkdir -p gpurun_out
L=tools/_build/insert_lab
out=gpurun_out/r02_insert_lab_d.jsonl; : > $out
run() { env "$@" timeout 120 $L 100000000 0.5 5 0 >> $out 2>> gpurun_out/r02_insert_lab_d.err || echo "{\"failed\": \"$*\"}" >> $out; }
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=0
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_STREAM_SLOTS=1
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_STREAM_SLOTS=4
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_BLOCKED_PREFETCH=0
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_REGION_MIB=8
run CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 CUCO_B200_REGION_MIB=32
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 0.8 5 0 >> $out
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 0.5 5 1 >> $out
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 100000000 0.8 5 1 >> $out
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=1 timeout 120 $L 10000000 0.5 5 0 >> $out
env CUCO_B200_TILE_ROUTE=1 CUCO_B200_STREAM_PROBE=0 timeout 120 $L 10000000 0.5 5 0 >> $out
cat $out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv \
  --log-file gpurun_out/r02_insert_lab_d_ncu.csv $L 100000000 0.5 1 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stream_mutate -c 1 -o gpurun_out/r02_stream_mutate_v2 $L 100000000 0.5 1 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tile_route -c 1 -o gpurun_out/r02_tile_route_v1 $L 100000000 0.5 1 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
