#!/usr/bin/env python3
"""Tuning sweep on the GPU box: times insert and find of the C2 configurations for every launch
variant of the native library and for cuco's build, prints one JSON line per point."""
import itertools
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi, key_generator as kg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
keys = kg.uniform(n, 1, torch.int64, dev, seed=42)
pairs = torch.stack([keys, keys], dim=1).contiguous()
ukeys = kg.unique(n, torch.int64, dev, seed=7)
upairs = torch.stack([ukeys, ukeys], dim=1).contiguous()
miss = keys + 2 * n
out = torch.empty(n, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream(dev)


def timeit(fn, reps=3):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def point(lib, label, probing, cg, lf, inp, q, tag):
    t = cb.static_map(n=n, load_factor=lf, probing=probing, cg_size=cg, device=dev, _library=lib)
    ins = []
    for _ in range(3):
        t.clear_async()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); t.insert_async(inp); b.record(stream)
        torch.cuda.synchronize()
        ins.append(a.elapsed_time(b))
    f = timeit(lambda: t.find(q, out))
    fm = timeit(lambda: t.find(miss, out))
    c = timeit(lambda: t.contains(q))
    print(json.dumps({"impl": label, "probing": f"{probing}<{cg}>", "lf": lf, "keys": tag,
                      "insert_gops": round(n / statistics.median(ins) / 1e6, 2),
                      "find_hit_gops": round(n / f / 1e6, 2), "find_miss_gops": round(n / fm / 1e6, 2),
                      "contains_gops": round(n / c / 1e6, 2)}), flush=True)
    t.close()


native = _cabi.native()
try:
    ref = _cabi.reference()
except Exception:
    ref = None

configs = [("linear_probing", 1, 0.5), ("linear_probing", 1, 0.8), ("double_hashing", 8, 0.5), ("double_hashing", 8, 0.8)]
if ref is not None:
    for probing, cg, lf in configs:
        point(ref, "cuco", probing, cg, lf, pairs, keys, "uniform")
    point(ref, "cuco", "linear_probing", 1, 0.5, upairs, ukeys, "unique")
    point(ref, "cuco", "linear_probing", 1, 0.8, upairs, ukeys, "unique")

for kpt, cas_first, sector, waves in itertools.product((1, 2, 4), (0, 1), (0, 1), (1, 2)):
    native.set_tuning(kpt, cas_first, sector, waves, 0, 1)
    label = f"ours kpt={kpt} cas_first={cas_first} sector={sector} waves={waves}"
    for probing, cg, lf in configs[:1] + configs[2:3]:
        point(native, label, probing, cg, lf, pairs, keys, "uniform")
# best-guess config across all points incl. unique keys (true fill = LF)
for kpt, cas_first in ((2, 1), (4, 1), (4, 0), (2, 0)):
    native.set_tuning(kpt, cas_first, 1, 1, 0, 1)
    label = f"ours kpt={kpt} cas_first={cas_first} sector=1 waves=1"
    for probing, cg, lf in configs:
        point(native, label, probing, cg, lf, pairs, keys, "uniform")
    point(native, label, "linear_probing", 1, 0.5, upairs, ukeys, "unique")
    point(native, label, "linear_probing", 1, 0.8, upairs, ukeys, "unique")
native.set_tuning(2, 1, 1, 1, 1, 1)
point(native, "ours generic", "linear_probing", 1, 0.5, pairs, keys, "uniform")
