#!/bin/bash
# Round-1 second GPU pass: hardware probes (window blocking, L2 fetch granularity, partition),
# drop-in device checks against both header trees, golden fixtures from cuco's own build,
# parity tests, bench (native + cuco), ncu launch list and full captures.
set -u
mkdir -p gpurun_out/golden
nproc > gpurun_out/nproc.txt
echo "== microbench2"; timeout 600 tools/_build/microbench2 3 > gpurun_out/microbench2.jsonl 2> gpurun_out/microbench2.err; echo "rc=$?"
cat gpurun_out/microbench2.jsonl; tail -3 gpurun_out/microbench2.err
echo "== device checks (native headers)"; timeout 600 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1; echo "rc=$?"
grep -c PASS gpurun_out/device_checks_native.log; grep FAIL gpurun_out/device_checks_native.log | head -20; tail -2 gpurun_out/device_checks_native.log
echo "== device checks (reference headers)"; timeout 600 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1; echo "rc=$?"
grep -c PASS gpurun_out/device_checks_ref.log; grep FAIL gpurun_out/device_checks_ref.log | head -20; tail -2 gpurun_out/device_checks_ref.log
echo "== golden"; timeout 600 python tools/make_golden.py gpurun_out/golden/cuco_golden.npz > gpurun_out/make_golden.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/make_golden.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 --detail --no-cpu-baseline > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"
cat gpurun_out/bench_reference.json | head -c 3000; tail -3 gpurun_out/bench_reference.err
echo "== bench native"; timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; echo "rc=$?"
cat gpurun_out/bench_native.json | head -c 3500; tail -3 gpurun_out/bench_native.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_native.csv python tools/profile_target.py native > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full native"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'lookup_kernel|mutate_kernel' -s 3 -c 3 -f -o gpurun_out/prof_native python tools/profile_target.py native > gpurun_out/ncu_full_native.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
