#!/usr/bin/env python3
"""Times the other single-GPU BASELINE.json configurations for this implementation and for cuco's own
build through the same C shim (one JSON line per configuration and arm; results are cross-checked):
  C1  static_set<int32>, 1 M uniform keys, LF 0.5: insert + contains (8 MB table, L2-resident)
  C3  static_map<int32,int32>, 100 M keys, LF 0.5, linear_probing<4>: find / contains with a 50 % miss
      rate, uniform and Gaussian (skew 0.5 and 0.1: heavy duplication) build streams
  C5  (one GPU's worth) static_map<int64,int64> insert_or_apply(plus): 250 M rows over 10 M distinct
      keys into a 20 M-slot table
usage: config_bench.py [scale]   scale divides the C3 / C5 sizes (default 1)."""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi, key_generator as kg  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)


def ms(fn, reps, setup=None):
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def libs():
    out = [("native", _cabi.native())]
    try:
        out.append(("reference", _cabi.reference()))
    except (FileNotFoundError, OSError):
        pass
    return out


def c1():
    n = 1_000_000
    keys = kg.uniform(n, 1, torch.int32, dev, seed=42)
    rows, outputs = [], []
    for name, lib in libs():
        t = cb.static_set(n=n, load_factor=0.5, key_dtype=torch.int32, probing="double_hashing", cg_size=4,
                          device=dev, _library=lib)
        out = torch.empty(n, dtype=torch.bool, device=dev)
        ins = ms(lambda: t.insert_async(keys), 20, setup=t.clear_async)
        con = ms(lambda: t.contains(keys, out), 20)
        outputs.append((t.size(), out.clone()))
        rows.append({"config": "C1 static_set<int32> 1M uniform LF0.5 double_hashing<4>", "impl": name,
                     "insert_gops": round(n / ins / 1e6, 2), "contains_gops": round(n / con / 1e6, 2),
                     "insert_us": round(ins * 1e3, 1), "contains_us": round(con * 1e3, 1), "size": outputs[-1][0]})
        t.close()
    check(outputs)
    return rows


def check(outputs):
    for o in outputs[1:]:
        assert o[0] == outputs[0][0], "sizes differ between the two implementations"
        for x, y in zip(o[1:], outputs[0][1:]):
            assert torch.equal(x, y), "per-key outputs differ between the two implementations"


def c3():
    n = 100_000_000 // scale
    rows = []
    for dist, build in (("uniform", kg.uniform(n, 1, torch.int32, dev, seed=42)),
                        ("gaussian skew 0.5", kg.gaussian(n, 0.5, torch.int32, dev, seed=42)),
                        ("gaussian skew 0.1", kg.gaussian(n, 0.1, torch.int32, dev, seed=42))):
        probe = kg.dropout(build, 0.5, seed=43)
        outputs = []
        for name, lib in libs():
            t = cb.static_map(n=n, load_factor=0.5, key_dtype=torch.int32, value_dtype=torch.int32,
                              probing="linear_probing", cg_size=4, device=dev, _library=lib)
            found = torch.empty(n, dtype=torch.int32, device=dev)
            present = torch.empty(n, dtype=torch.bool, device=dev)
            ins = ms(lambda: t.insert_async(build, build), 3, setup=t.clear_async)
            fnd = ms(lambda: t.find(probe, found), 5)
            con = ms(lambda: t.contains(probe, present), 5)
            outputs.append((t.size(), found.clone(), present.clone()))
            rows.append({"config": f"C3 static_map<int32,int32> {n} keys LF0.5 linear_probing<4>, {dist} build, "
                                   "50% miss probes", "impl": name, "insert_gops": round(n / ins / 1e6, 2),
                         "find_gops": round(n / fnd / 1e6, 2), "contains_gops": round(n / con / 1e6, 2),
                         "distinct": outputs[-1][0], "hit_rate": round(float(present.float().mean().item()), 3)})
            t.close()
            del t
        check(outputs)
    return rows


def c5():
    rows_n, distinct = 250_000_000 // scale, 10_000_000 // scale
    keys = torch.randint(1, distinct + 1, (rows_n,), device=dev, dtype=torch.int64,
                         generator=torch.Generator(device=dev).manual_seed(42))
    pairs = torch.stack([keys, torch.ones_like(keys)], dim=1).contiguous()
    del keys
    q = torch.arange(1, distinct + 1, device=dev, dtype=torch.int64)
    rows, outputs = [], []
    for name, lib in libs():
        t = cb.static_map(n=distinct, load_factor=0.5, empty_value=0, probing="linear_probing", cg_size=1,
                          device=dev, _library=lib)
        agg = ms(lambda: t.insert_or_apply(pairs, op="plus"), 3, setup=t.clear_async)
        sums = t.find(q)
        outputs.append((t.size(), sums.clone()))
        rows.append({"config": f"C5 (one shard) insert_or_apply(plus) {rows_n} rows / {distinct} distinct, "
                               "static_map<int64,int64> LF0.5 linear_probing<1>", "impl": name,
                     "rows_gops": round(rows_n / agg / 1e6, 2), "ms": round(agg, 3),
                     "sum_check": int(sums.sum().item()) == rows_n})
        t.close()
    check(outputs)
    return rows


for fn in (c1, c3, c5):
    for row in fn():
        print(json.dumps(row), flush=True)
    torch.cuda.empty_cache()
