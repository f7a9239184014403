#!/bin/bash
# Blocked-path v2: parity, variant timings, per-launch DRAM sectors, full ncu capture of both passes.
set -u
mkdir -p gpurun_out
echo "== device checks"; timeout 600 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1; echo "rc=$?"
timeout 600 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1; echo "rc=$?"
grep -c PASS gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log; grep -B3 FAIL gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log | head -30
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== insert probe (timings)"; timeout 900 python tools/insert_probe.py 100000000 3 > gpurun_out/insert_probe.jsonl 2> gpurun_out/insert_probe.err; echo "rc=$?"; cat gpurun_out/insert_probe.jsonl; tail -3 gpurun_out/insert_probe.err
echo "== insert probe under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'mutate|route' --csv --log-file gpurun_out/insert_probe_ncu.csv python tools/insert_probe.py 100000000 1 "direct kpt1,blocked r16 kpt2,blocked r16 kpt4 casfirst" > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/insert_probe_ncu.csv
echo "== ncu full blocked"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'route_kernel|blocked_mutate' -c 2 -f -o /tmp/prof_blocked python tools/insert_probe.py 100000000 1 "blocked r16 kpt2" > gpurun_out/ncu_full_blocked.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_blocked.ncu-rep --page raw --csv > gpurun_out/prof_blocked_raw.csv 2>/dev/null
ncu -i /tmp/prof_blocked.ncu-rep --page source --csv > gpurun_out/prof_blocked_source.csv 2>/dev/null
ncu -i /tmp/prof_blocked.ncu-rep --page details > gpurun_out/prof_blocked_details.txt 2>/dev/null
ls -la gpurun_out/ /tmp/prof_blocked.ncu-rep
