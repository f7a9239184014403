#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== device checks"; timeout 600 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1; echo "rc=$?"
timeout 600 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1; echo "rc=$?"
grep -c PASS gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log; grep -B3 FAIL gpurun_out/device_checks_native.log gpurun_out/device_checks_ref.log | head -30
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== insert probe (timings)"; timeout 900 python tools/insert_probe.py 100000000 3 > gpurun_out/insert_probe.jsonl 2> gpurun_out/insert_probe.err; echo "rc=$?"; cat gpurun_out/insert_probe.jsonl; tail -3 gpurun_out/insert_probe.err
echo "== insert probe under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'mutate|route' --csv --log-file gpurun_out/insert_probe_ncu.csv python tools/insert_probe.py 100000000 1 "direct kpt1,direct kpt2,blocked r16 kpt2,blocked r16 kpt4,blocked r16 kpt4 casfirst" > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_table.py gpurun_out/insert_probe_ncu.csv
echo "== bench native"; timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; echo "rc=$?"
cat gpurun_out/bench_native.json | head -c 3500; tail -3 gpurun_out/bench_native.err
echo "== bench native blocked"; CUCO_B200_BLOCKED=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native_blocked.json 2> gpurun_out/bench_native_blocked.err; echo "rc=$?"
cat gpurun_out/bench_native_blocked.json | head -c 2500; tail -3 gpurun_out/bench_native_blocked.err
ls -la gpurun_out/
