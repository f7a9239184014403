#!/bin/bash
# round 2, call I (1 GPU): full -m gpu suite (blocked count / retrieve, erase with two keys per thread, staged
# exchange, cuco's Catch2 suites), matches bench direct vs blocked, next-rows, staging-kernel probe + ncu
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 18 gpurun_out/r02i_pytest.log | cut -c1-250
CUCO_B200_BLOCKED=0 timeout 200 python tools/matches_bench.py 50000000 > gpurun_out/r02_matches_bench_direct.jsonl 2> gpurun_out/r02_matches_bench_direct.err
echo "matches_bench direct rc=$?"; cat gpurun_out/r02_matches_bench_direct.jsonl | cut -c1-700
timeout 200 python tools/matches_bench.py 50000000 > gpurun_out/r02_matches_bench_blocked.jsonl 2> gpurun_out/r02_matches_bench_blocked.err
echo "matches_bench blocked(auto) rc=$?"; cat gpurun_out/r02_matches_bench_blocked.jsonl | cut -c1-700; tail -n 3 gpurun_out/r02_matches_bench_blocked.err
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err
echo "next_rows rc=$?"; cat gpurun_out/r02_next_rows.jsonl
timeout 300 python tools/stage_probe.py 100000000 5 > gpurun_out/r02_stage_probe.jsonl 2> gpurun_out/r02_stage_probe.err
echo "stage_probe rc=$?"; cat gpurun_out/r02_stage_probe.jsonl; tail -n 3 gpurun_out/r02_stage_probe.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:exchange_route_kernel -s 3 -c 1 \
  -o gpurun_out/r02_key_stage_v1 python tools/stage_probe.py 100000000 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
