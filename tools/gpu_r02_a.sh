#!/bin/bash
# round 2, call A: hardware probes for pass 1 / pass 2 of the blocked insert + regression check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 300 tools/_build/microbench4 3 > gpurun_out/r02_microbench4.jsonl 2> gpurun_out/r02_microbench4.err
echo "microbench4 rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r02a_bench.json
