#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== microbench3"; timeout 300 tools/_build/microbench3 3 > gpurun_out/microbench3.jsonl 2> gpurun_out/microbench3.err; echo "rc=$?"
cat gpurun_out/microbench3.jsonl; tail -3 gpurun_out/microbench3.err
echo "== microbench3 under ncu (sector counts)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/microbench3_ncu.csv tools/_build/microbench3 3 > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/microbench3_ncu.csv
echo "== insert probe (timings)"; timeout 900 python tools/insert_probe.py 100000000 3 > gpurun_out/insert_probe.jsonl 2> gpurun_out/insert_probe.err; echo "rc=$?"; cat gpurun_out/insert_probe.jsonl; tail -3 gpurun_out/insert_probe.err
echo "== insert probe under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'mutate|route' --csv --log-file gpurun_out/insert_probe_ncu.csv python tools/insert_probe.py 100000000 1 > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/insert_probe_ncu.csv
