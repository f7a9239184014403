#!/usr/bin/env python3
"""Every bulk operation of the hot path on the C2 table - 100 M uniform
int64 pairs, linear_probing<1>, LF 0.5 - for this implementation and for cuco's own build through the
same shim; outputs cross-checked between the two. One JSON line per arm."""
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
stream = torch.cuda.current_stream(dev)
keys = kg.uniform(n, 1, torch.int64, dev, seed=42)
pairs = torch.stack([keys, keys * 3 + 1], dim=1).contiguous()
probe = kg.dropout(keys, 0.5, seed=43)
stencil = (torch.arange(n, device=dev) % 3 != 0)


def ms(fn, reps=3, setup=None):
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


libs = [("native", _cabi.native())]
try:
    libs.append(("reference", _cabi.reference()))
except (FileNotFoundError, OSError):
    pass
rows, checks = [], []
for name, lib in libs:
    t = cb.static_map(n=n, load_factor=0.5, probing="linear_probing", cg_size=1, device=dev, _library=lib)
    out64 = torch.empty(n, dtype=torch.int64, device=dev)
    outb = torch.empty(n, dtype=torch.bool, device=dev)
    r = {"impl": name, "n": n}
    r["insert"] = ms(lambda: t.insert_async(pairs), setup=t.clear_async)
    r["insert_if (2/3 pass)"] = ms(lambda: t.insert_if_async(pairs, stencil), setup=t.clear_async)
    holder = {}
    r["insert_and_find"] = ms(lambda: holder.update(r=t.insert_and_find(pairs)), setup=t.clear_async)
    found_iaf, inserted = holder["r"]
    r["insert_or_assign"] = ms(lambda: t.insert_or_assign(pairs), setup=t.clear_async)
    t.clear_async()
    t.insert_async(pairs)
    r["insert_or_assign (all present)"] = ms(lambda: t.insert_or_assign(pairs))
    r["find (50% miss)"] = ms(lambda: t.find(probe, out64))
    r["contains (50% miss)"] = ms(lambda: t.contains(probe, outb))
    r["contains_if (2/3 pass)"] = ms(lambda: t.contains_if(probe, stencil, outb))
    size = t.size()
    checks.append((size, int(inserted.sum().item()), found_iaf.clone(), out64.clone(), outb.clone()))
    t.close()
    del t
    z = cb.static_map(n=n, load_factor=0.5, empty_value=0, probing="linear_probing", cg_size=1, device=dev,
                      _library=lib)
    ones = torch.stack([keys, torch.ones_like(keys)], dim=1).contiguous()
    r["insert_or_apply(plus)"] = ms(lambda: z.insert_or_apply(ones, op="plus"), setup=z.clear_async)
    sums = z.find(keys)
    checks[-1] = checks[-1] + (sums.clone(),)
    z.close()
    del z, ones
    torch.cuda.empty_cache()
    rows.append({k: (round(n / v / 1e6, 2) if isinstance(v, float) else v) for k, v in r.items()})
for c in checks[1:]:
    assert c[0] == checks[0][0] and c[1] == checks[0][1] == checks[0][0], (c[:2], checks[0][:2])
    assert all(torch.equal(x, y) for x, y in zip(c[2:], checks[0][2:])), "outputs differ between the implementations"
for row in rows:
    row["unit"] = "Gops/s"
    print(json.dumps(row))
