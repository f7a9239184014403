#!/bin/bash
# Round-1 evidence pass on one GPU: launch list of the bench command, full ncu capture of the three
# hot kernels (exported to CSV/text on the box: the .ncu-rep files are too large to travel), bench of
# both arms.
set -u
mkdir -p gpurun_out
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "rc=$?"
python tools/ncu_table.py gpurun_out/launches_bench.csv | grep -E "lookup_kernel|route_kernel|blocked_mutate|fill_slots" | tail -12
echo "== ncu full: lookup / route / blocked_mutate"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lookup_kernel|route_kernel|blocked_mutate_kernel' -s 9 -c 3 -f -o /tmp/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
ncu -i /tmp/prof_final.ncu-rep --page details > gpurun_out/prof_final_details.txt 2>/dev/null
ncu -i /tmp/prof_final.ncu-rep --page source --csv > gpurun_out/prof_final_source.csv 2>/dev/null
ls -la /tmp/prof_final.ncu-rep gpurun_out/prof_final*
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 --detail --no-cpu-baseline > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
echo "== bench native"; timeout 900 python bench.py --steps 10 --warmup 3 --detail > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; echo "rc=$?"
cat gpurun_out/bench_native.json | head -c 4000; tail -3 gpurun_out/bench_native.err
