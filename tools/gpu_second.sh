#!/bin/bash
# Hardware ceilings, tuning sweep and first ncu captures.
set -u
mkdir -p gpurun_out
echo "== microbench"; timeout 600 tools/_build/microbench 4 > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "rc=$?"
cat gpurun_out/microbench.jsonl
echo "== sweep"; timeout 1500 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "rc=$?"; tail -3 gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_native.csv python tools/profile_target.py native > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
grep -E "lookup_kernel|mutate_kernel|fill_slots" gpurun_out/launches_native.csv | awk -F'","' '{print $5, $NF}' | cut -c1-200 | tail -12
echo "== ncu full native"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'lookup_kernel|mutate_kernel' -s 2 -c 3 -f -o gpurun_out/prof_native python tools/profile_target.py native > gpurun_out/ncu_full_native.log 2>&1; echo "rc=$?"
echo "== ncu full reference"; timeout 1200 ncu --set full --clock-control none -k regex:'insert_if_n|find|contains_if_n' -s 3 -c 3 -f -o gpurun_out/prof_reference python tools/profile_target.py reference > gpurun_out/ncu_full_reference.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
