#!/usr/bin/env python3
"""SURVEY.md §8(f) rank 2-3 rows - join probe and multiset - timed for this implementation and for
cuco's own build through the same shim, outputs cross-checked:
  static_set<int64> double_hashing<4>: insert of n unique keys, retrieve with 2n probes (50 % hits)
  static_multiset<int64> linear_probing<1> storage<2>: insert of n elements (multiplicity 4), count and
  retrieve with n/2 probes (50 % hits).
The synchronous calls (count, retrieve) are timed by wall clock around the call; insert by events."""
import json
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
mult = 4
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)


def ms(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def wall_ms(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out


def best(fn, reps=3):
    times, out = [], None
    for _ in range(reps):
        t, out = wall_ms(fn)
        times.append(t)
    return min(times), out


libs = [("native", _cabi.native())]
if not os.environ.get("MATCHES_NATIVE_ONLY"):
    try:
        libs.append(("reference", _cabi.reference()))
    except (FileNotFoundError, OSError):
        pass

set_keys = torch.randperm(n, device=dev, dtype=torch.int64)
set_probes = torch.randperm(2 * n, device=dev, dtype=torch.int64)
multi_keys = (torch.randperm(n, device=dev) // mult).to(torch.int64)
multi_probes = torch.randperm(n // 2, device=dev, dtype=torch.int64)  # keys < n/4 are present

rows, checks = [], []
for name, lib in libs:
    s = cb.static_set(n=n, load_factor=0.5, key_dtype=torch.int64, device=dev, _library=lib)
    s.insert_async(set_keys); s.clear(); torch.cuda.synchronize()
    set_insert_ms = ms(lambda: s.insert_async(set_keys))
    set_retrieve_ms, (p, m) = best(lambda: s.retrieve(set_probes))
    set_rows = p.numel()
    set_ok = bool(torch.equal(p, m)) and bool(torch.equal(p.sort().values, torch.arange(n, device=dev)))
    s.close(); del s, p, m
    torch.cuda.empty_cache()

    t = cb.static_multiset(n=n, load_factor=0.5, key_dtype=torch.int64, probing="linear_probing",
                           cg_size=1, window_size=2, device=dev, _library=lib)
    t.insert_async(multi_keys); t.clear(); torch.cuda.synchronize()
    multi_insert_ms = ms(lambda: t.insert_async(multi_keys))
    size = t.size()
    count_ms, count = best(lambda: t.count(multi_probes))
    retrieve_ms, (p, m) = best(lambda: t.retrieve(multi_probes))   # includes the sizing count
    multi_ok = bool(torch.equal(p, m))
    hist = torch.bincount(p, minlength=n // 2)
    rows.append({
        "impl": name, "n": n, "match_ahead": int(os.environ.get("CUCO_B200_MATCH_AHEAD", "1")),
        "set_insert_gops": round(n / set_insert_ms / 1e6, 2),
        "set_retrieve_gprobes": round(2 * n / set_retrieve_ms / 1e6, 2), "set_retrieve_rows": set_rows,
        "multiset_insert_gops": round(n / multi_insert_ms / 1e6, 2),
        "multiset_count_gprobes": round((n // 2) / count_ms / 1e6, 2), "multiset_count": count,
        "multiset_retrieve_grows": round(p.numel() / retrieve_ms / 1e6, 2), "multiset_rows": p.numel(),
        "checks_ok": set_ok and multi_ok and size == n})
    checks.append((set_rows, size, count, hist.clone()))
    t.close(); del t, p, m
    torch.cuda.empty_cache()
for c in checks[1:]:
    assert c[:3] == checks[0][:3], (c[:3], checks[0][:3])
    assert torch.equal(c[3], checks[0][3]), "retrieve rows differ between the implementations"
for r in rows:
    print(json.dumps(r))
