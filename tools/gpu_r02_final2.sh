#!/bin/bash
# round 2, last 1-GPU pass on the final binaries: whole -m gpu suite, smoke(), next-rows, both bench arms
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest_gpu_full.log 2>&1
echo "pytest rc=$?"; tail -n 10 gpurun_out/r02_pytest_gpu_full.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?"
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err; cat gpurun_out/r02_next_rows.jsonl
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_1gpu_reference_final.json 2> gpurun_out/r02_bench_1gpu_reference_final.err
echo "bench reference rc=$?"
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err
echo "bench rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r02_bench_1gpu_reference_final.json'))
d = json.load(open('gpurun_out/r02_bench_1gpu_final.json'))
for x in (r, d):
    print(x['impl'], {k: round(x[k], 3) for k in ('value', 'insert_gops', 'find_gops', 'insert_ms', 'find_ms')}, 'e2e', round(x['e2e']['value'], 3), x['clocks'])
for row in d['c2_points']: print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()})
for row in d['sweep']: print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()})
PY
