#!/bin/bash
# Look-ahead sweep of the all-matches walks (count / retrieve on a multiset) + ncu captures.
mkdir -p gpurun_out
for a in 2 4; do
  CUCO_B200_MATCH_AHEAD=$a timeout 120 python -u -m pytest tests/test_matches_gpu.py -x -q --timeout 100 \
    --timeout-method thread -k "multiset or multimap or golden" > gpurun_out/pytest_matches_ahead$a.log 2>&1
  echo "pytest ahead=$a rc=$?"; tail -1 gpurun_out/pytest_matches_ahead$a.log
done
: > gpurun_out/matches_ahead.jsonl
for a in 1 2 4; do
  MATCHES_NATIVE_ONLY=1 CUCO_B200_MATCH_AHEAD=$a timeout 100 python tools/matches_bench.py 50000000 \
    >> gpurun_out/matches_ahead.jsonl 2>> gpurun_out/matches_ahead.err
done
cat gpurun_out/matches_ahead.jsonl
for k in count_kernel retrieve_kernel; do
  MATCHES_NATIVE_ONLY=1 timeout 150 ncu --set full --clock-control none --import-source on \
    -k regex:$k -c 1 -f -o gpurun_out/r01_ncu_$k python tools/matches_bench.py 20000000 \
    > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
