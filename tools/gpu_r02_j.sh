#!/bin/bash
# round 2, call J (1 GPU): the reference suite that failed in call I on its own, then the full -m gpu suite,
# staging probe, next-rows and the bench line
mkdir -p gpurun_out
timeout 120 tests/_build/reftests/static_map__shared_memory_test_native > gpurun_out/r02j_shared_memory_native.log 2>&1
echo "shared_memory native rc=$?"; tail -n 12 gpurun_out/r02j_shared_memory_native.log | cut -c1-300
timeout 120 oracle/_ref/reftests/static_map__shared_memory_test_ref > gpurun_out/r02j_shared_memory_ref.log 2>&1
echo "shared_memory ref rc=$?"; tail -n 6 gpurun_out/r02j_shared_memory_ref.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 16 gpurun_out/r02j_pytest.log | cut -c1-250
timeout 300 python tools/stage_probe.py 100000000 5 > gpurun_out/r02_stage_probe_v2.jsonl 2> gpurun_out/r02_stage_probe_v2.err
echo "stage_probe rc=$?"; cat gpurun_out/r02_stage_probe_v2.jsonl | cut -c1-200; tail -n 3 gpurun_out/r02_stage_probe_v2.err
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err
echo "next_rows rc=$?"; cat gpurun_out/r02_next_rows.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02j_bench.json'))
print({k: d[k] for k in ('value', 'insert_gops', 'find_gops', 'insert_ms', 'find_ms')}, d['e2e']['value'])
PY
