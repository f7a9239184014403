#!/usr/bin/env python3
"""Minimal launch sequence for ncu: clear, insert, find, contains on the headline configuration.
usage: profile_target.py [native|reference] [n] [probing cg]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi, key_generator as kg  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "native"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
probing = sys.argv[3] if len(sys.argv) > 3 else "linear_probing"
cg = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lib = _cabi.native() if impl == "native" else _cabi.reference()
dev = torch.device("cuda", 0)
keys = kg.uniform(n, 1, torch.int64, dev, seed=42)
pairs = torch.stack([keys, keys], dim=1).contiguous()
out = torch.empty(n, dtype=torch.int64, device=dev)
t = cb.static_map(n=n, load_factor=0.5, probing=probing, cg_size=cg, device=dev, _library=lib)
torch.cuda.synchronize()
for _ in range(2):
    t.clear_async()
    t.insert_async(pairs)
    t.find(keys, out)
    t.contains(keys)
torch.cuda.synchronize()
assert bool((out == keys).all().item())
print("ok", impl, n, probing, cg, t.size())
