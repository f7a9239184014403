#!/bin/bash
# First on-GPU pass: smoke, parity tests, bench (native + cuco reference), ncu launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== bench native"; timeout 900 python bench.py --steps 5 --warmup 3 --detail > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; echo "rc=$?"
cat gpurun_out/bench_native.json | head -c 3000; tail -3 gpurun_out/bench_native.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 --detail --no-cpu-baseline > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"
cat gpurun_out/bench_reference.json | head -c 3000; tail -3 gpurun_out/bench_reference.err
