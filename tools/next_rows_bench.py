#!/usr/bin/env python3
"""SURVEY.md §8(f) rank 1-2 rows around the hot path - erase, size, clear, rehash, retrieve_all - timed
for this implementation and for cuco's own build through the same shim, outputs cross-checked.
static_map<int64,int64>, linear_probing<1>, 50 M unique pairs in a capacity-100 M table."""
import json
import statistics
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
keys = torch.randperm(n, device=dev, dtype=torch.int64)
pairs = torch.stack([keys, keys + 1], dim=1).contiguous()


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


rows, checks = [], []
libs = [("native", _cabi.native())]
try:
    libs.append(("reference", _cabi.reference()))
except (FileNotFoundError, OSError):
    pass
for name, lib in libs:
    t = cb.static_map(capacity=2 * n, erased_key=-2, probing="linear_probing", cg_size=1, device=dev, _library=lib)
    t.insert_async(pairs)
    t.size()  # first call allocates the device counter
    size_runs = [wall_ms(t.size) for _ in range(5)]  # host call to host result, as a user sees it
    size_ms, size0 = statistics.median(r[0] for r in size_runs), size_runs[0][1]
    t.erase(keys[:1024] + 3 * n)  # absent keys: loads the kernel (CUDA loads modules lazily) outside the timed call
    torch.cuda.synchronize()
    erase_ms = ms(lambda: t.erase(keys[: n // 2]))
    size1 = t.size()
    present = t.contains(keys)
    retrieve_runs = [wall_ms(t.retrieve_all) for _ in range(4)]  # the first run pays the cold allocations
    retrieve_first, got = retrieve_runs[0]
    retrieve_ms = statistics.median(r[0] for r in retrieve_runs[1:])
    rk = got[0].sort().values
    del retrieve_runs
    rehash_runs = [wall_ms(lambda: t.rehash())[0] for _ in range(4)]
    rehash_first, rehash_ms = rehash_runs[0], statistics.median(rehash_runs[1:])
    found = t.find(keys)
    clear_ms = statistics.median(ms(t.clear_async) for _ in range(3))
    rows.append({"impl": name, "n": n, "capacity": t.capacity(), "size_ms": round(size_ms, 3),
                 "erase_gops": round(n / 2 / erase_ms / 1e6, 2), "retrieve_all_ms": round(retrieve_ms, 3),
                 "rehash_ms": round(rehash_ms, 3), "retrieve_all_first_ms": round(retrieve_first, 3),
                 "rehash_first_ms": round(rehash_first, 3), "clear_ms": round(clear_ms, 3),
                 "clear_GBps": round(t.capacity() * 16 / clear_ms / 1e6, 1)})
    checks.append((size0, size1, present.clone(), rk.clone(), found.clone()))
    t.close()
    del t
    torch.cuda.empty_cache()
for c in checks[1:]:
    assert c[0] == checks[0][0] and c[1] == checks[0][1]
    assert all(torch.equal(x, y) for x, y in zip(c[2:], checks[0][2:])), "outputs differ between the implementations"
for r in rows:
    print(json.dumps(r))
