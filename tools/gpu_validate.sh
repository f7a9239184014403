#!/bin/bash
# Round-end validation on the GPU box, every step bounded: full `-m gpu` suite, smoke(), default bench.
mkdir -p gpurun_out
timeout 100 tests/_build/dynamic_map_checks_native > gpurun_out/dynamic_map_checks_native.log 2>&1
echo "dynamic_map_checks_native rc=$?"; tail -3 gpurun_out/dynamic_map_checks_native.log
timeout 100 oracle/_ref/dynamic_map_checks_ref > gpurun_out/dynamic_map_checks_ref.log 2>&1
echo "dynamic_map_checks_ref rc=$?"; tail -3 gpurun_out/dynamic_map_checks_ref.log
timeout 380 python -u -m pytest tests -m gpu -x -q --timeout 150 --timeout-method thread --durations 8 \
  > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest_gpu rc=$?"
tail -15 gpurun_out/pytest_gpu_full.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 140 python bench.py > gpurun_out/bench_native_r01b.json 2> gpurun_out/bench_native_r01b.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_native_r01b.json; tail -2 gpurun_out/bench_native_r01b.err
timeout 120 python tools/matches_bench.py 50000000 > gpurun_out/matches_bench.jsonl 2> gpurun_out/matches_bench.err
echo "matches_bench rc=$?"; cat gpurun_out/matches_bench.jsonl
