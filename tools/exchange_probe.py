#!/usr/bin/env python3
"""Times the kernels of the fused exchange path with P simulated ranks living on ONE GPU (peer
pointers are plain device pointers), so each stage can be profiled with ncu (which cannot wrap a
multi-rank run). usage: exchange_probe.py [n_per_rank] [P]"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi, key_generator as kg, partitioned as cbp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = _cabi.native()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)


class Rank:
    def __init__(self, me):
        self.me = me
        self.table = cb.static_map(n=int(n * 1.03) + 1, load_factor=0.5, probing="linear_probing", cg_size=1,
                                   device=dev, _library=lib)
        r, cap, sp = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib.check(lib.exchange_plan(self.table._handle, n, P, C.byref(r), C.byref(cap), C.byref(sp)))
        self.R, self.cap, self.spill_cap = r.value, cap.value, sp.value
        seg = P * self.R * self.cap
        z = dict(device=dev)
        self.segments = torch.empty(seg * 16, dtype=torch.uint8, **z)
        self.counts = torch.zeros(self.R * P, dtype=torch.int32, **z)
        self.flags = torch.zeros(P, dtype=torch.int32, **z)
        self.results = torch.empty(seg * 8, dtype=torch.uint8, **z)
        self.counts_local = torch.zeros(P * self.R, dtype=torch.int32, **z)
        self.position_local = torch.empty(n, dtype=torch.int32, **z)
        self.spill = torch.empty(self.spill_cap * 16, dtype=torch.uint8, **z)
        self.spill_index = torch.empty(self.spill_cap, dtype=torch.int32, **z)
        self.spill_count = torch.zeros(1, dtype=torch.int32, **z)
        self.keys = kg.uniform(n, 1, torch.int64, dev, seed=42 + me) + me * n
        self.pairs = torch.stack([self.keys, self.keys], dim=1).contiguous()
        self.out = torch.empty(n, dtype=torch.int64, **z)


ranks = [Rank(me) for me in range(P)]
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
peers = lambda attr: (C.c_void_p * P)(*[getattr(r, attr).data_ptr() for r in ranks])  # noqa: E731
times = {}


def timed(label, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    times.setdefault(label, []).append(a.elapsed_time(b))


def route(r, keys_only):
    lib.check(lib.exchange_route(r.table._handle, vp(r.keys if keys_only else r.pairs), None, n, int(keys_only),
                                 r.R, r.cap, r.spill_cap, P, r.me, cbp.DEFAULT_SALT, peers("segments"),
                                 peers("counts"), peers("flags"), vp(r.counts_local), vp(r.position_local),
                                 vp(r.spill), vp(r.spill_index), vp(r.spill_count), None))


for rep in range(3):
    for r in ranks:
        r.table.clear_async()
    for r in ranks:
        timed("route pairs", lambda: route(r, False))
    for r in ranks:
        timed("probe received segments",
              lambda: lib.check(lib.exchange_mutate(r.table._handle, vp(r.segments), vp(r.counts), r.R, r.cap, P, -1, None)))
    for r in ranks:
        timed("route keys", lambda: route(r, True))
    for r in ranks:
        timed("lookup received segments",
              lambda: lib.check(lib.exchange_lookup(r.table._handle, vp(r.segments), vp(r.counts), peers("results"),
                                                    r.R, r.cap, P, r.me, 0, None)))
    for r in ranks:
        timed("unpermute", lambda: lib.check(lib.exchange_unpermute(r.table._handle, vp(r.results),
                                                                     vp(r.position_local), n, vp(r.out), 0, None)))
for r in ranks:
    assert bool((r.out == r.keys).all().item()), "wrong payloads"
    assert int(r.flags.sum().item()) == 0
print(json.dumps({"n_per_rank": n, "simulated_ranks": P, "R": ranks[0].R, "segment_capacity": ranks[0].cap,
                  "ms": {k: round(min(v), 4) for k, v in times.items()}}))
