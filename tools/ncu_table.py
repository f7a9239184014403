#!/usr/bin/env python3
"""Prints an `ncu --csv --metrics ...` log as one line per launch: id, kernel, metric=value ..."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((int(r[idi]), r[ki][:70]), {})[r[mi].split("__")[-1]] = r[vi]
for (i, k), v in sorted(out.items()):
    print(i, k, " ".join(f"{a}={b}" for a, b in v.items()))
