#!/usr/bin/env python3
"""Times the source-side staging kernels of the staged exchange on ONE GPU (they never touch a peer):
pairs grouped by (owner, bucket) for several rank counts / bucket counts, keys grouped by owner.

    python tools/stage_probe.py [n_pairs] [reps]      one JSON line per configuration
"""
import ctypes as C
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi  # noqa: E402
from cucollections_b200 import key_generator as kg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = _cabi.native()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
keys = kg.uniform(n, 1, torch.int64, dev, seed=42)
pairs = torch.stack([keys, keys], dim=1).contiguous()
table = cb.static_map(n=n, load_factor=0.5, probing="linear_probing", cg_size=1, device=dev, _library=lib)


def vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def ms(fn):
    out = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            out.append(a.elapsed_time(b))
    return statistics.median(out)


def stage(elems, count, keys_only, P, buckets):
    cap, spill = C.c_uint32(), C.c_uint32()
    lib.check(lib.exchange_stage_plan(table._handle, count, P, buckets, C.byref(cap), C.byref(spill)))
    eb = 8 if keys_only else 16
    buf = torch.empty(P * buckets * cap.value * eb, dtype=torch.uint8, device=dev)
    counts = torch.zeros(P * buckets, dtype=torch.int32, device=dev)
    pos = torch.empty(count, dtype=torch.int32, device=dev)
    sp = torch.empty(spill.value * eb, dtype=torch.uint8, device=dev)
    spi = torch.empty(spill.value, dtype=torch.int32, device=dev)
    spc = torch.zeros(1, dtype=torch.int32, device=dev)

    def run():
        lib.check(lib.exchange_stage(table._handle, vp(elems), None, count, int(keys_only), buckets, cap.value,
                                     spill.value, P, 0, 0x9E3779B97F4A7C15, vp(buf), vp(counts), vp(pos), vp(sp),
                                     vp(spi), vp(spc), C.c_void_p(stream.cuda_stream)))
    t = ms(run)
    assert int(counts.sum().item()) + int(spc.item()) == count
    return t


for P, buckets in ((1, 197), (2, 1), (2, 8), (2, 99), (2, 197), (8, 1), (8, 8), (8, 16), (8, 50), (8, 99), (8, 128)):
    t = stage(pairs, n, False, P, buckets)
    print(json.dumps({"what": "pairs", "n": n, "ranks": P, "buckets_per_owner": buckets, "total_buckets": P * buckets,
                      "ms": round(t, 3), "GBps_in_plus_out": round(32 * n / t / 1e6, 1)}), flush=True)
for P in (2, 4, 8):
    for m in (n // 4, n):
        t = stage(keys[:m], m, True, P, 1)
        print(json.dumps({"what": "keys", "n": m, "ranks": P, "ms": round(t, 3),
                          "GBps_in_plus_out": round(20 * m / t / 1e6, 1)}), flush=True)
