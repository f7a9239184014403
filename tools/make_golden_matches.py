#!/usr/bin/env python3
"""Generates tests/golden/cuco_golden_matches.npz: outputs of cuCollections' OWN implementation
(oracle/_ref/libcuco_ref.so) for the rows next to the hot path - `static_set::retrieve`,
`static_multiset` insert / insert_if / contains / find / count / count_outer / retrieve /
retrieve_outer and `experimental::static_multimap` insert / insert_if / contains / contains_if /
count - on seeded inputs, run on a B200.

Run on the GPU box:  python tools/make_golden_matches.py gpurun_out/golden/cuco_golden_matches.npz
then copy the file to tests/golden/. `run_kind` is the recorded scenario; tests/
test_golden_matches.py replays it on the CPU oracle (no GPU) and tests/test_matches_gpu.py on the
native kernels. Row order of retrieve is unspecified by the reference, so rows are recorded sorted.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

N = 3000
SEED = 20241017
KINDS = (0, 5, 10, 11, 12)  # static_set<int32>, <int64>, static_multiset<int32>, <int64>, multimap<int64,int64>
MULTISETS = (10, 11)
MULTIMAPS = (12,)


def inputs(kind: int):
    """Build side with multiplicities 1..~12 plus a few heavy keys; probes half present, half absent."""
    rng = np.random.default_rng(SEED + kind)
    a = rng.integers(1, N // 3, size=N, dtype=np.int64)
    a[: N // 20] = rng.integers(1, 6, size=N // 20, dtype=np.int64)
    rng.shuffle(a)
    q = np.concatenate([np.arange(1, N // 3, dtype=np.int64),
                        rng.integers(N, 2 * N, size=N // 3, dtype=np.int64)])
    rng.shuffle(q)
    stencil = rng.integers(0, 2, size=N, dtype=np.uint8)
    return a, q, stencil


def sorted_rows(probe, match):
    probe, match = np.asarray(probe, np.int64), np.asarray(match, np.int64)
    order = np.lexsort((match, probe))
    return np.stack([probe[order], match[order]], axis=1)


class GpuBackend:
    """Tables of one C-ABI library (native or reference build) on cuda:0."""

    def __init__(self, kind, lib):
        import torch

        import cucollections_b200 as cb
        self.torch, self.cb, self.lib = torch, cb, lib
        self.k = cb.KINDS[kind]

    def make(self, **kw):
        k = self.k
        common = dict(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size, window_size=k.window_size,
                      hash=k.hash, device=self.torch.device("cuda", 0), _library=self.lib, **kw)
        if k.value is not None:
            return self.cb.static_multimap(value_dtype=k.value, **common)
        return self.cb.static_multiset(**common) if k.multi else self.cb.static_set(**common)

    def keys(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(self.k.key)

    def values(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(self.k.value)

    def stencil(self, s):
        return self.torch.from_numpy(s).to("cuda")

    @staticmethod
    def host(x):
        return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


class OracleBackend:
    """The CPU oracle behind the same call shapes."""

    def __init__(self, kind):
        from oracle import oracle
        self.oracle, self.kind = oracle, kind

    def make(self, n=None, load_factor=0.0, capacity=None):
        size = n if n is not None else capacity
        return self.oracle.Table.for_kind(self.kind, size, load_factor if n is not None else 0.0)

    @staticmethod
    def keys(a):
        return a

    @staticmethod
    def values(a):
        return a

    @staticmethod
    def stencil(s):
        return s

    @staticmethod
    def host(x):
        return np.asarray(x)


def run_scenario(kind, b):
    a, q, stencil = inputs(kind)
    tag = f"k{kind}_"
    out = {tag + "a": a, tag + "q": q, tag + "stencil": stencil}

    caps = []
    for lf in (0.5, 0.8, 1.0):
        caps.append(b.make(n=N, load_factor=lf).capacity())
    for c in (0, 1, 400, 1234, 2 * N):
        caps.append(b.make(capacity=c).capacity())
    out[tag + "capacities"] = np.asarray(caps, dtype=np.int64)

    t = b.make(n=2 * N, load_factor=0.7)  # insert_if (~N/2) + insert (N) elements stay below capacity
    if kind in MULTIMAPS:
        vals = a * 3 + 1
        out[tag + "insert_if_new"] = np.int64(t.insert_if(b.keys(a), b.stencil(stencil), b.values(vals)))
        out[tag + "count_after_insert_if"] = np.int64(t.count(b.keys(q)))
        out[tag + "insert_new"] = np.int64(t.insert(b.keys(a), b.values(vals)))
        out[tag + "contains_q"] = b.host(t.contains(b.keys(q))).astype(bool)
        out[tag + "contains_if_q"] = b.host(t.contains_if(b.keys(q), b.stencil(stencil[: q.size]))).astype(bool)
        out[tag + "count_inner"] = np.int64(t.count(b.keys(q)))
        out[tag + "count_self"] = np.int64(t.count(b.keys(a)))
        return out
    out[tag + "insert_if_new"] = np.int64(t.insert_if(b.keys(a), b.stencil(stencil)))
    out[tag + "size_after_insert_if"] = np.int64(t.size())
    out[tag + "insert_new"] = np.int64(t.insert(b.keys(a)))
    out[tag + "size"] = np.int64(t.size())
    out[tag + "contains_q"] = b.host(t.contains(b.keys(q))).astype(bool)
    out[tag + "find_q"] = b.host(t.find(b.keys(q))).astype(np.int64)
    if kind in MULTISETS:
        for outer in (False, True):
            name = "outer" if outer else "inner"
            out[tag + f"count_{name}"] = np.int64(t.count(b.keys(q), outer))
            out[tag + f"rows_{name}"] = sorted_rows(*(b.host(x) for x in t.retrieve(b.keys(q), outer)))
        out[tag + "count_self"] = np.int64(t.count(b.keys(a)))
    else:
        out[tag + "rows_inner"] = sorted_rows(*(b.host(x) for x in t.retrieve(b.keys(q))))
    return out


def run_kind(kind, lib):
    return run_scenario(kind, GpuBackend(kind, lib))


def run_kind_oracle(kind):
    return run_scenario(kind, OracleBackend(kind))


def main(out_path: str):
    import torch

    from cucollections_b200 import _cabi
    lib = _cabi.reference()
    assert lib.flavour.startswith("reference"), lib.flavour
    out = {"n": np.int64(N), "seed": np.int64(SEED)}
    for kind in KINDS:
        out.update(run_kind(kind, lib))
    torch.cuda.synchronize()
    Path(out_path).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **out)
    print(f"wrote {out_path}: {len(out)} arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden/cuco_golden_matches.npz")
