#!/bin/bash
# GPU-box script for the rows next to the hot path (static_set::retrieve, static_multiset):
# parity tests, drop-in device checks against both header trees, cuco fixtures, bench vs cuco.
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_matches_gpu.py -x -q > gpurun_out/pytest_matches.log 2>&1
echo "pytest_matches rc=$?"
tail -5 gpurun_out/pytest_matches.log
timeout 300 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1
echo "device_checks_native rc=$?"
timeout 300 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1
echo "device_checks_ref rc=$?"
grep -c PASS gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log
grep FAIL gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log | head -20
timeout 300 python tools/make_golden_matches.py gpurun_out/golden/cuco_golden_matches.npz
timeout 600 python tools/matches_bench.py > gpurun_out/matches_bench.jsonl 2> gpurun_out/matches_bench.err
echo "matches_bench rc=$?"
cat gpurun_out/matches_bench.jsonl
tail -3 gpurun_out/matches_bench.err
