#!/bin/bash
# round 2, call G: whole -m gpu suite (staged exchange incl. fine mode, cuco's Catch2 suites, BASELINE configs),
# next-rows timings, bench line with sweep
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 22 gpurun_out/r02g_pytest.log | cut -c1-250
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err
echo "next_rows rc=$?"; cat gpurun_out/r02_next_rows.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02g_bench.json'))
print({k: d[k] for k in ('value', 'insert_gops', 'find_gops', 'insert_ms', 'find_ms')})
print(d['roofline']); print(d['roofline_other_pass'])
for r in d.get('sweep', []): print(r)
PY
