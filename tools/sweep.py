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

def tune(kpt=2, cas_first=0, sector=1, waves=1, generic=0, coherent=0, blocked=0, region=32):
    native.set_tuning(kpt, cas_first, sector, waves, generic, 1, coherent)
    native.set_blocking(blocked, region)
    return f"ours kpt={kpt} casf={cas_first} sec={sector} waves={waves} gen={generic} coh={coherent} blk={blocked} reg={region}"


lp, dh = configs[0], configs[2]
# 1. unblocked insert: load policy, cas-first, keys per thread, waves
for kpt, cas_first, coherent, waves in itertools.product((1, 2, 4), (0, 1), (0, 1), (1, 4)):
    label = tune(kpt=kpt, cas_first=cas_first, coherent=coherent, waves=waves)
    point(native, label, *lp, pairs, keys, "uniform")
# 2. lookups: waves and kpt
for kpt, waves in itertools.product((1, 2, 4), (1, 2, 4, 8, 64)):
    label = tune(kpt=kpt, waves=waves)
    point(native, label, *lp, pairs, keys, "uniform")
# 3. blocked insert: region size and kpt
for kpt, region in itertools.product((1, 2, 4), (8, 16, 32, 48, 64)):
    label = tune(kpt=kpt, blocked=1, region=region)
    point(native, label, *lp, pairs, keys, "uniform")
# 4. full table of the C2 points with the blocked path on/off
for blocked in (0, 1):
    label = tune(blocked=blocked)
    for probing, cg, lf in configs:
        point(native, label, probing, cg, lf, pairs, keys, "uniform")
    point(native, label, "linear_probing", 1, 0.5, upairs, ukeys, "unique")
    point(native, label, "linear_probing", 1, 0.8, upairs, ukeys, "unique")
label = tune(generic=1)
point(native, label, *lp, pairs, keys, "uniform")
