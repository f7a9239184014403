#!/bin/bash
# round 2, final 1-GPU pass: whole -m gpu suite, smoke(), both bench arms, ncu launch list of the bench command,
# ncu --set full of the three hot kernels (exported to text on the box), next-rows and matches benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > gpurun_out/r02_final_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu_full.log 2>&1
echo "pytest rc=$?"; tail -n 14 gpurun_out/r02_pytest_gpu_full.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_1gpu_reference_final.json 2> gpurun_out/r02_bench_1gpu_reference_final.err
echo "bench reference rc=$?"
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err
echo "bench rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r02_bench_1gpu_reference_final.json'))
d = json.load(open('gpurun_out/r02_bench_1gpu_final.json'))
for x in (r, d):
    print(x['impl'], {k: round(x[k], 3) for k in ('value', 'insert_gops', 'find_gops', 'insert_ms', 'find_ms')}, 'e2e', round(x['e2e']['value'], 3), x['clocks'])
print(d['roofline']); print(d['roofline_other_pass'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r02_launches_bench_native.csv python bench.py --steps 2 --warmup 3 --no-points --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu launch list rc=$?"
for k in tile_route_kernel blocked_mutate_kernel lookup_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/r02_final_$k \
    python bench.py --steps 1 --warmup 3 --no-points --no-cpu-baseline > /dev/null 2>&1
  ncu -i gpurun_out/r02_final_$k.ncu-rep --page details > gpurun_out/r02_final_${k}_details.txt 2>/dev/null
  ncu -i gpurun_out/r02_final_$k.ncu-rep --page raw --csv > gpurun_out/r02_final_${k}_raw.csv 2>/dev/null
  ls -la gpurun_out/r02_final_$k.ncu-rep | awk '{print $5, $NF}'
  rm -f gpurun_out/r02_final_$k.ncu-rep   # the text exports are what is kept (pull limit 64 MiB)
done
timeout 300 python tools/next_rows_bench.py > gpurun_out/r02_next_rows.jsonl 2> gpurun_out/r02_next_rows.err; cat gpurun_out/r02_next_rows.jsonl
timeout 200 python tools/matches_bench.py 50000000 > gpurun_out/r02_matches_bench_blocked.jsonl 2> gpurun_out/r02_matches_bench_blocked.err; cat gpurun_out/r02_matches_bench_blocked.jsonl | cut -c1-400
