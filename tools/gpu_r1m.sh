#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== exchange probe (timings)"; timeout 600 python tools/exchange_probe.py 50000000 2 > gpurun_out/exchange_probe.json 2> gpurun_out/exchange_probe.err; echo "rc=$?"; cat gpurun_out/exchange_probe.json; tail -5 gpurun_out/exchange_probe.err
echo "== exchange probe under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'exchange|blocked_mutate' -c 14 --csv --log-file gpurun_out/exchange_probe_ncu.csv python tools/exchange_probe.py 50000000 2 > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/exchange_probe_ncu.csv
