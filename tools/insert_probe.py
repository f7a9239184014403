#!/usr/bin/env python3
"""Times one bulk insert (and find) of the headline configuration for a list of launch variants of
the native library; run plain for event timings, or under
  ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct -k regex:mutate
for DRAM sector counts per variant (launch order = print order)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi, key_generator as kg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
lib = _cabi.native()
stream = torch.cuda.current_stream(dev)
keys = kg.uniform(n, 1, torch.int64, dev, seed=42)
pairs = torch.stack([keys, keys], dim=1).contiguous()
ukeys = kg.unique(n, torch.int64, dev, seed=7)
upairs = torch.stack([ukeys, ukeys], dim=1).contiguous()

# label -> (keys_per_thread code, cas_first, waves, generic, blocked mode, region MiB, blocked kpt,
#           blocked cas_first, blocked prefetch)
VARIANTS = [
    ("direct kpt1", 11, 0, 0, 0, 0, 16, 2, 0, 1),
    ("direct kpt2", 22, 0, 0, 0, 0, 16, 2, 0, 1),
    ("generic", 11, 0, 0, 1, 0, 16, 2, 0, 1),
    ("blocked r16 kpt1", 11, 0, 0, 0, 1, 16, 1, 0, 1),
    ("blocked r16 kpt2", 11, 0, 0, 0, 1, 16, 2, 0, 1),
    ("blocked r16 kpt4", 11, 0, 0, 0, 1, 16, 4, 0, 1),
    ("blocked r16 kpt2 noprefetch", 11, 0, 0, 0, 1, 16, 2, 0, 0),
    ("blocked r16 kpt1 casfirst", 11, 0, 0, 0, 1, 16, 1, 1, 1),
    ("blocked r16 kpt2 casfirst", 11, 0, 0, 0, 1, 16, 2, 1, 1),
    ("blocked r16 kpt4 casfirst", 11, 0, 0, 0, 1, 16, 4, 1, 1),
    ("blocked r16 kpt2 casfirst noprefetch", 11, 0, 0, 0, 1, 16, 2, 1, 0),
    ("blocked r8 kpt2", 11, 0, 0, 0, 1, 8, 2, 0, 1),
    ("blocked r32 kpt2", 11, 0, 0, 0, 1, 32, 2, 0, 1),
    ("blocked r32 kpt4 casfirst", 11, 0, 0, 0, 1, 32, 4, 1, 1),
]
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for lf in (0.5, 0.8):
    for label, kpt, casf, waves, gen, blk, region, bkpt, bcas, bpf in VARIANTS:
        if only and not any(o in label for o in only):
            continue
        lib.set_tuning(kpt, casf, 1, waves, gen, 1, 0)
        lib.set_blocking(blk, region)
        lib.set_blocking_variant(bkpt, bcas, bpf)
        t = cb.static_map(n=n, load_factor=lf, probing="linear_probing", cg_size=1, device=dev, _library=lib)
        row = {"variant": label, "lf": lf}
        for tag, inp, distinct in (("uniform", pairs, None), ("unique", upairs, n)):
            ts = []
            for _ in range(reps):
                t.clear_async()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); t.insert_async(inp); b.record(stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            row[f"insert_{tag}_gops"] = round(n / min(ts) / 1e6, 2)
            row[f"size_{tag}"] = t.size()
        print(json.dumps(row), flush=True)
        t.close()
