#!/bin/bash
set -u
mkdir -p gpurun_out
VARIANT="${1:-blocked r16 kpt2}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'route_kernel|blocked_mutate' -c 2 -f -o /tmp/prof_blocked python tools/insert_probe.py 100000000 1 "$VARIANT" > gpurun_out/ncu_full_blocked.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_blocked.ncu-rep --page raw --csv > gpurun_out/prof_blocked_raw.csv 2>/dev/null
ncu -i /tmp/prof_blocked.ncu-rep --page source --csv > gpurun_out/prof_blocked_source.csv 2>/dev/null
ncu -i /tmp/prof_blocked.ncu-rep --page details > gpurun_out/prof_blocked_details.txt 2>/dev/null
ls -la gpurun_out/prof_blocked*
