#!/bin/bash
set -u
mkdir -p gpurun_out
for k in 1 2 4; do echo "== exchange probe lookup kpt=$k"; CUCO_B200_EXCHANGE_LOOKUP_KPT=$k timeout 600 python tools/exchange_probe.py 50000000 2 2>&1 | tail -1; done
