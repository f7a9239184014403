#!/bin/bash
# ncu --set full of the multiset count kernel and the multiset retrieve kernel (final versions).
mkdir -p gpurun_out
MATCHES_NATIVE_ONLY=1 timeout 70 ncu --set full --clock-control none --import-source on \
  -k regex:retrieve_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r01_ncu_multiset_retrieve_v2 \
  python tools/matches_bench.py 20000000 > gpurun_out/ncu_multiset_retrieve.log 2>&1
echo "ncu retrieve rc=$?"
MATCHES_NATIVE_ONLY=1 timeout 70 ncu --set full --clock-control none --import-source on \
  -k regex:count_kernel -c 1 -f -o gpurun_out/r01_ncu_multiset_count_v2 \
  python tools/matches_bench.py 20000000 > gpurun_out/ncu_multiset_count.log 2>&1
echo "ncu count rc=$?"
ls -la gpurun_out/*v2.ncu-rep
