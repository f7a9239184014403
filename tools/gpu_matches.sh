#!/bin/bash
# GPU-box script for the rows next to the hot path (static_set::retrieve, static_multiset):
# cuco fixtures, parity tests, drop-in device checks against both header trees, bench vs cuco.
# Every step is bounded: an overfilled open-addressing table never terminates.
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 120 python tools/make_golden_matches.py gpurun_out/golden/cuco_golden_matches.npz
echo "make_golden_matches rc=$?"
cp -f gpurun_out/golden/cuco_golden_matches.npz tests/golden/ 2>/dev/null
timeout 300 python -u -m pytest tests/test_matches_gpu.py -x -v --timeout 100 --timeout-method thread \
  > gpurun_out/pytest_matches.log 2>&1
echo "pytest_matches rc=$?"
grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed" gpurun_out/pytest_matches.log | tail -25
timeout 60 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1
echo "device_checks_ref rc=$?"
timeout 60 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1
echo "device_checks_native rc=$?"
grep -c PASS gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log
grep FAIL gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log | head -20
timeout 150 python tools/matches_bench.py ${MATCHES_N:-50000000} > gpurun_out/matches_bench.jsonl 2> gpurun_out/matches_bench.err
echo "matches_bench rc=$?"
cat gpurun_out/matches_bench.jsonl
tail -3 gpurun_out/matches_bench.err
