#!/usr/bin/env python3
"""Generates tests/golden/cuco_golden.npz: outputs of cuCollections' OWN implementation
(oracle/_ref/libcuco_ref.so = the reference headers compiled behind the C-ABI shim) on seeded inputs,
run on a B200. The fixtures pin both the CPU oracle (tests/test_golden_fixtures.py, no GPU needed)
and the native kernels (-m gpu) to what the reference really returns.

Run on the GPU box:  python tools/make_golden.py gpurun_out/golden/cuco_golden.npz
then copy the file to tests/golden/. Only layout-independent results are recorded (SURVEY.md §8a'):
per-key find/contains outputs, insert counts, sizes, capacities, sorted retrieve_all, and for
insert_and_find the found values plus the *set of keys* that reported `inserted` (which duplicate
reports it is unspecified)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import _cabi  # noqa: E402

N = 2000
SEED = 20240917


def inputs(kind: int):
    """Seeded inputs of one kind: batch A (duplicates), batch B (half present, half new), queries."""
    rng = np.random.default_rng(SEED + kind)
    a = rng.integers(1, N, size=N, dtype=np.int64)                    # ~63 % distinct
    b = np.concatenate([a[: N // 2], rng.integers(N, 2 * N, size=N // 2, dtype=np.int64)])
    rng.shuffle(b)
    q = np.concatenate([a[: N // 2], rng.integers(2 * N, 4 * N, size=N // 2, dtype=np.int64)])
    rng.shuffle(q)
    stencil = rng.integers(0, 2, size=N, dtype=np.uint8)
    return a, b, q, stencil


def value_of(keys):
    return keys * 3 + 1


def main(out_path: str):
    lib = _cabi.reference()
    assert lib.flavour.startswith("reference"), lib.flavour
    dev = torch.device("cuda", 0)
    out = {"n": np.int64(N), "seed": np.int64(SEED)}

    def dev_t(a, dtype):
        return torch.from_numpy(a.astype(np.int64)).to(dev).to(dtype)

    for kind, k in cb.KINDS.items():
        if k.multi:
            continue  # static_multiset kinds are recorded by tools/make_golden_matches.py
        a, b, q, stencil = inputs(kind)
        is_map = k.value is not None
        tag = f"k{kind}_"
        out[tag + "a"], out[tag + "b"], out[tag + "q"], out[tag + "stencil"] = a, b, q, stencil

        def make(**kw):
            if is_map:
                return cb.static_map(key_dtype=k.key, value_dtype=k.value, probing=k.probing,
                                     cg_size=k.cg_size, window_size=k.window_size, hash=k.hash,
                                     device=dev, _library=lib, **kw)
            return cb.static_set(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size,
                                 window_size=k.window_size, hash=k.hash, device=dev, _library=lib, **kw)

        def args(keys):
            kt = dev_t(keys, k.key)
            return (kt, dev_t(value_of(keys), k.value)) if is_map else (kt,)

        # capacities: (n, load factor) and plain capacity constructors
        caps = []
        for lf in (0.5, 0.8, 1.0):
            t = make(n=N, load_factor=lf)
            caps.append(t.capacity())
            t.close()
        for c in (0, 1, 400, 1234, 2 * N):
            t = make(capacity=c)
            caps.append(t.capacity())
            t.close()
        out[tag + "capacities"] = np.asarray(caps, dtype=np.int64)

        # insert / size / find / contains / contains_if
        t = make(n=N, load_factor=0.5)
        out[tag + "insert_new"] = np.int64(t.insert(*args(a)))
        out[tag + "size_after_a"] = np.int64(t.size())
        out[tag + "find_q"] = t.find(dev_t(q, k.key)).cpu().numpy().astype(np.int64)
        out[tag + "contains_q"] = t.contains(dev_t(q, k.key)).cpu().numpy()
        out[tag + "contains_if_q"] = t.contains_if(dev_t(q, k.key), torch.from_numpy(stencil).to(dev)).cpu().numpy()
        # insert_and_find of batch B on top
        try:
            found, inserted = t.insert_and_find(*args(b))
            found, inserted = found.cpu().numpy().astype(np.int64), inserted.cpu().numpy()
            out[tag + "iaf_found"] = found
            out[tag + "iaf_inserted_count"] = np.int64(inserted.sum())
            out[tag + "iaf_inserted_keys"] = np.unique(b[inserted])
        except cb.CucoError:
            # the reference build of kind 6 lacks insert_and_find (nvcc aborts on it, see
            # cucollections_b200/csrc/cabi_kind.cu); plain insert keeps the later results defined
            t.insert(*args(b))
        out[tag + "size_after_b"] = np.int64(t.size())
        ra = t.retrieve_all()
        if is_map:
            rk, rv = (x.cpu().numpy().astype(np.int64) for x in ra)
            order = np.argsort(rk, kind="stable")
            out[tag + "retrieve_keys"], out[tag + "retrieve_values"] = rk[order], rv[order]
        else:
            out[tag + "retrieve_keys"] = np.sort(ra.cpu().numpy().astype(np.int64))
        t.close()

        # insert_if
        t = make(n=N, load_factor=0.8)
        out[tag + "insert_if_new"] = np.int64(t.insert_if(args(a)[0], torch.from_numpy(stencil).to(dev), *args(a)[1:]))
        out[tag + "insert_if_contains"] = t.contains(dev_t(a, k.key)).cpu().numpy()
        t.close()

        # erase with a tombstone sentinel (capacity constructor, like the reference's erase_test)
        t = make(capacity=2 * N, erased_key=-2)
        t.insert(*args(a))
        t.erase(dev_t(a[: N // 2], k.key))
        out[tag + "erase_contains"] = t.contains(dev_t(a, k.key)).cpu().numpy()
        out[tag + "erase_size"] = np.int64(t.size())
        t.close()

        if is_map:
            # upserts: assign (value pure function of key), apply plus/min/max with and without init
            t = make(n=N, load_factor=0.5)
            t.insert_or_assign(dev_t(a, k.key), dev_t(value_of(a) + 7, k.value))
            rk, rv = (x.cpu().numpy().astype(np.int64) for x in t.retrieve_all())
            order = np.argsort(rk, kind="stable")
            out[tag + "assign_keys"], out[tag + "assign_values"] = rk[order], rv[order]
            t.close()
            ones = np.ones(N, dtype=np.int64)
            for name, op, empty_value, init, vals in (
                    ("plus", "plus", 0, None, ones), ("plus_init", "plus", 0, 0, a),
                    ("min", "min", np.iinfo(np.int32).max, None, a % 17 + np.arange(N) % 5),
                    ("max", "max", np.iinfo(np.int32).min, None, a % 17 + np.arange(N) % 5)):
                t = make(n=N, load_factor=0.5, empty_value=int(empty_value))
                t.insert_or_apply(dev_t(a, k.key), dev_t(vals, k.value), op=op, init=init)
                rk, rv = (x.cpu().numpy().astype(np.int64) for x in t.retrieve_all())
                order = np.argsort(rk, kind="stable")
                out[tag + f"apply_{name}_in"] = vals.astype(np.int64)
                out[tag + f"apply_{name}_keys"], out[tag + f"apply_{name}_values"] = rk[order], rv[order]
                t.close()
    torch.cuda.synchronize()
    Path(out_path).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **out)
    print(f"wrote {out_path}: {len(out)} arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden/cuco_golden.npz")
