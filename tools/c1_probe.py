import sys, json, statistics
sys.path.insert(0, "/root/repo")
import torch
import cucollections_b200 as cb
from cucollections_b200 import _cabi, key_generator as kg
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev)
n = 1_000_000
keys = kg.uniform(n, 1, torch.int32, dev, seed=42)
for name, lib in (("native", _cabi.native()), ("reference", _cabi.reference())):
    for l2 in ((1, 0) if name == "native" else (1,)):
        if name == "native":
            lib.set_tuning(-1, -1, -1, -1, -1, l2, -1)
        t = cb.static_set(n=n, load_factor=0.5, key_dtype=torch.int32, probing="double_hashing", cg_size=4, device=dev, _library=lib)
        out = torch.empty(n, dtype=torch.bool, device=dev)
        def ms(fn, setup=None):
            ts = []
            for _ in range(30):
                if setup: setup()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            return statistics.median(ts) * 1e3
        print(name, "l2_window", l2, "insert_us", round(ms(lambda: t.insert_async(keys), setup=t.clear_async), 1), "contains_us", round(ms(lambda: t.contains(keys, out)), 1), flush=True)
        t.close()
