#!/bin/bash
# round 2, call F: the whole -m gpu suite (incl. cuco's own Catch2 suites and the BASELINE-config tests),
# both bench arms, and the ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02f_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/r02f_pytest.log
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
echo "bench reference rc=$?"; cut -c1-400 gpurun_out/r02f_bench_reference.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02f_bench.json'))
print({k: d[k] for k in ('value', 'insert_gops', 'find_gops', 'insert_ms', 'find_ms', 'value_median', 'value_best')})
print(d['e2e']); print(d['roofline']); print(d.get('c2_points'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r02_launches_bench_native.csv python bench.py --steps 2 --warmup 3 --no-points --no-cpu-baseline > gpurun_out/r02f_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
