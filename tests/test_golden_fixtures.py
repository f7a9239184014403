"""Golden fixtures recorded from cuCollections' OWN implementation on a B200
(tests/golden/cuco_golden.npz, written by tools/make_golden.py through oracle/_ref/libcuco_ref.so).

  * CPU (`-m "not gpu"`): the C oracle must reproduce every recorded result - this is what pins the
    oracle to the real reference rather than to our reading of it.
  * GPU (`-m gpu`): the native sm_100a kernels, through the C ABI, must reproduce them too.

Only layout-independent results are recorded (SURVEY.md §8a'); everything is compared bit-exactly.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

GOLDEN = Path(__file__).resolve().parent / "golden" / "cuco_golden.npz"
KINDS = list(range(10))


@pytest.fixture(scope="module")
def gold():
    assert GOLDEN.exists(), f"{GOLDEN} is missing (regenerate with tools/make_golden.py on a GPU box)"
    return np.load(GOLDEN)


def value_of(keys):
    return keys * 3 + 1


class OracleFactory:
    """Builds oracle tables with the call shapes of the recorded scenario."""

    def __init__(self, kind):
        self.kind = kind
        self.is_map = oracle.KIND_GEOMETRY[kind][1] != 0

    def make(self, n=None, load_factor=0.0, capacity=None, **kw):
        size = n if n is not None else capacity
        return oracle.Table.for_kind(self.kind, size, load_factor if n is not None else 0.0, **kw)

    def args(self, keys, values=None):
        if not self.is_map:
            return (keys,)
        return (keys, value_of(keys) if values is None else values)

    @staticmethod
    def keys(a):
        return a

    @staticmethod
    def np(x):
        return np.asarray(x)

    @staticmethod
    def sorted_all(t, is_map):
        if is_map:
            k, v = t.retrieve_all()
            order = np.argsort(k, kind="stable")
            return k[order], v[order]
        return np.sort(t.retrieve_all()), None


def run_scenario(f, g, kind):
    """Replays tools/make_golden.py's scenario on factory `f`, asserting against fixture `g`."""
    tag = f"k{kind}_"
    a, b, q, stencil = (g[tag + x] for x in ("a", "b", "q", "stencil"))
    n = int(g["n"])

    caps = []
    for lf in (0.5, 0.8, 1.0):
        caps.append(f.make(n=n, load_factor=lf).capacity())
    for c in (0, 1, 400, 1234, 2 * n):
        caps.append(f.make(capacity=c).capacity())
    assert caps == g[tag + "capacities"].tolist()

    t = f.make(n=n, load_factor=0.5)
    assert t.insert(*f.args(a)) == int(g[tag + "insert_new"])
    assert t.size() == int(g[tag + "size_after_a"])
    assert np.array_equal(f.np(t.find(f.keys(q))).astype(np.int64), g[tag + "find_q"])
    assert np.array_equal(f.np(t.contains(f.keys(q))), g[tag + "contains_q"])
    assert np.array_equal(f.contains_if(t, q, stencil), g[tag + "contains_if_q"])
    if tag + "iaf_found" in g.files:
        found, inserted = t.insert_and_find(*f.args(b))
        found, inserted = f.np(found).astype(np.int64), f.np(inserted).astype(bool)
        assert np.array_equal(found, g[tag + "iaf_found"])
        assert int(inserted.sum()) == int(g[tag + "iaf_inserted_count"])
        assert np.array_equal(np.unique(b[inserted]), g[tag + "iaf_inserted_keys"])
    else:
        t.insert(*f.args(b))
    assert t.size() == int(g[tag + "size_after_b"])
    rk, rv = f.sorted_all(t, f.is_map)
    assert np.array_equal(rk, g[tag + "retrieve_keys"])
    if f.is_map:
        assert np.array_equal(rv, g[tag + "retrieve_values"])

    t = f.make(n=n, load_factor=0.8)
    assert f.insert_if(t, a, stencil) == int(g[tag + "insert_if_new"])
    assert np.array_equal(f.np(t.contains(f.keys(a))), g[tag + "insert_if_contains"])

    t = f.make(capacity=2 * n, erased_key=-2)
    t.insert(*f.args(a))
    t.erase(f.keys(a[: n // 2]))
    assert np.array_equal(f.np(t.contains(f.keys(a))), g[tag + "erase_contains"])
    assert t.size() == int(g[tag + "erase_size"])

    if f.is_map:
        t = f.make(n=n, load_factor=0.5)
        t.insert_or_assign(*f.args(a, value_of(a) + 7))
        rk, rv = f.sorted_all(t, True)
        assert np.array_equal(rk, g[tag + "assign_keys"])
        assert np.array_equal(rv, g[tag + "assign_values"])
        for name, op, empty_value, init in (("plus", "plus", 0, None), ("plus_init", "plus", 0, 0),
                                            ("min", "min", np.iinfo(np.int32).max, None),
                                            ("max", "max", np.iinfo(np.int32).min, None)):
            vals = g[tag + f"apply_{name}_in"]
            t = f.make(n=n, load_factor=0.5, empty_value=int(empty_value))
            f.apply(t, a, vals, op, init)
            rk, rv = f.sorted_all(t, True)
            assert np.array_equal(rk, g[tag + f"apply_{name}_keys"]), name
            assert np.array_equal(rv, g[tag + f"apply_{name}_values"]), name


class OracleRunner(OracleFactory):
    def contains_if(self, t, q, stencil):
        return t.contains(q, stencil)

    def insert_if(self, t, a, stencil):
        return t.insert_if(a, stencil, value_of(a) if self.is_map else None)

    def apply(self, t, keys, vals, op, init):
        code = {"plus": oracle.PLUS, "min": oracle.MIN, "max": oracle.MAX}[op]
        t.insert_or_apply(keys, vals, code, init)


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reproduces_cuco_outputs(kind, gold):
    run_scenario(OracleRunner(kind), gold, kind)


class NativeRunner:
    """The product path: sm_100a kernels through the C ABI, torch tensors as device buffers."""

    def __init__(self, kind, lib):
        import torch

        import cucollections_b200 as cb
        self.torch, self.cb, self.lib = torch, cb, lib
        self.k = cb.KINDS[kind]
        self.is_map = self.k.value is not None
        self.dev = torch.device("cuda", 0)

    def dev_t(self, a, dtype):
        return self.torch.from_numpy(np.ascontiguousarray(a).astype(np.int64)).to(self.dev).to(dtype)

    def keys(self, a):
        return self.dev_t(a, self.k.key)

    def make(self, **kw):
        k = self.k
        common = dict(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size, window_size=k.window_size,
                      hash=k.hash, device=self.dev, _library=self.lib, **kw)
        if self.is_map:
            return self.cb.static_map(value_dtype=k.value, **common)
        common.pop("empty_value", None)
        return self.cb.static_set(**common)

    def args(self, keys, values=None):
        if not self.is_map:
            return (self.keys(keys),)
        return (self.keys(keys), self.dev_t(value_of(keys) if values is None else values, self.k.value))

    @staticmethod
    def np(x):
        return x.cpu().numpy()

    def contains_if(self, t, q, stencil):
        return t.contains_if(self.keys(q), self.torch.from_numpy(stencil).to(self.dev)).cpu().numpy()

    def insert_if(self, t, a, stencil):
        st = self.torch.from_numpy(stencil).to(self.dev)
        args = self.args(a)
        return t.insert_if(args[0], st, *args[1:])

    def apply(self, t, keys, vals, op, init):
        t.insert_or_apply(self.keys(keys), self.dev_t(vals, self.k.value), op=op, init=init)

    def sorted_all(self, t, is_map):
        if is_map:
            k, v = (x.cpu().numpy().astype(np.int64) for x in t.retrieve_all())
            order = np.argsort(k, kind="stable")
            return k[order], v[order]
        return np.sort(t.retrieve_all().cpu().numpy().astype(np.int64)), None


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_native_kernels_reproduce_cuco_outputs(kind, gold, native_lib):
    run_scenario(NativeRunner(kind, native_lib), gold, kind)
