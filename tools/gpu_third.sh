#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 2400 python tools/sweep.py > gpurun_out/sweep2.jsonl 2> gpurun_out/sweep2.err; echo "rc=$?"; tail -3 gpurun_out/sweep2.err
