#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (blocked + golden)"; timeout 1800 python -m pytest tests -m gpu -x -q -k "blocked or golden or smoke or round_trip" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== insert probe (timings)"; timeout 900 python tools/insert_probe.py 100000000 3 > gpurun_out/insert_probe.jsonl 2> gpurun_out/insert_probe.err; echo "rc=$?"; cat gpurun_out/insert_probe.jsonl; tail -3 gpurun_out/insert_probe.err
echo "== insert probe under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:'mutate|route' --csv --log-file gpurun_out/insert_probe_ncu.csv python tools/insert_probe.py 100000000 1 "blocked r16 kpt1,blocked r16 kpt2,blocked r16 kpt4,blocked r32 kpt2" > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/insert_probe_ncu.csv
