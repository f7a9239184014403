#!/bin/bash
# Diagnosis pass: load-then-CAS variants with DRAM sector counts, blocked insert launch breakdown,
# device checks logs, golden fixtures.
set -u
mkdir -p gpurun_out/golden
echo "== microbench3"; timeout 300 tools/_build/microbench3 3 > gpurun_out/microbench3.jsonl 2> gpurun_out/microbench3.err; echo "rc=$?"
cat gpurun_out/microbench3.jsonl; tail -3 gpurun_out/microbench3.err
echo "== microbench3 under ncu (sector counts)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_atom.sum,lts__t_sectors_op_read.sum --clock-control none --csv --log-file gpurun_out/microbench3_ncu.csv tools/_build/microbench3 3 > /dev/null 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/microbench3_ncu.csv')) if len(r)>10]
hdr=rows[0]
ki,mi,vi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
idi=hdr.index('ID')
out={}
for r in rows[1:]:
    out.setdefault((r[idi],r[ki][:60]),{})[r[mi]]=r[vi]
for k,v in out.items():
    print(k, v)
PY
echo "== device checks"; timeout 600 tests/_build/device_checks_native > gpurun_out/device_checks_native.log 2>&1; echo "rc=$?"
timeout 600 oracle/_ref/device_checks_ref > gpurun_out/device_checks_ref.log 2>&1; echo "rc=$?"
grep -B3 FAIL gpurun_out/device_checks_native.log | head -40; echo ----; grep -B3 FAIL gpurun_out/device_checks_ref.log | head -40
echo "== golden"; timeout 600 python tools/make_golden.py gpurun_out/golden/cuco_golden.npz > gpurun_out/make_golden.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/make_golden.log
echo "== ncu launches blocked"; CUCO_B200_BLOCKED=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum --clock-control none --csv --log-file gpurun_out/launches_blocked.csv python tools/profile_target.py native > gpurun_out/ncu_launch_blocked.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_blocked.csv')) if len(r)>10]
hdr=rows[0]
ki,mi,vi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
idi=hdr.index('ID')
out={}
for r in rows[1:]:
    out.setdefault((r[idi],r[ki][:50]),{})[r[mi]]=r[vi]
for k,v in out.items():
    print(k, v)
PY
ls -la gpurun_out/
