"""GPU parity tests of the rows next to the hot path: `static_set::retrieve`
`static_multiset` (insert, insert_if, contains, find, count, count_outer, retrieve, retrieve_outer)
and `experimental::static_multimap` (insert, insert_if, contains, contains_if, count), through the C
ABI, against
  * the CPU oracle (oracle/cuco_oracle.c: oracle_count / oracle_retrieve, allows_duplicates), and
  * cuco itself (oracle/_ref/libcuco_ref.so) on the same seeded inputs,
plus the fixtures cuco produced on a B200 (tests/golden/cuco_golden_matches.npz).
Row order of retrieve is unspecified by the reference: rows are compared sorted. Bit-exact.

Mirrors tests/static_set/retrieve_test.cu and tests/static_multiset/{insert,contains,find,count,
custom_count,retrieve,large_input}_test.cu of the reference.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

import cucollections_b200 as cb
from cucollections_b200 import _cabi
from oracle import oracle

pytestmark = pytest.mark.gpu

MULTISET_KINDS = [_cabi.MULTISET_I32_DH4_W2, _cabi.MULTISET_I64_LP1_W2]
MULTIMAP_KINDS = [_cabi.MULTIMAP_I64_LP4]
SET_KINDS = [_cabi.SET_I32_DH4, _cabi.SET_I64_DH4]
GOLDEN = Path(__file__).resolve().parent / "golden" / "cuco_golden_matches.npz"


def make(kind, lib, **kw):
    k = cb.KINDS[kind]
    common = dict(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size, window_size=k.window_size,
                  hash=k.hash, _library=lib, **kw)
    if k.value is not None:
        return cb.static_multimap(value_dtype=k.value, **common)
    return cb.static_multiset(**common) if k.multi else cb.static_set(**common)


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype)


def rows(probe, match):
    """Row multiset of a retrieve result as a lexicographically sorted [n, 2] int64 array."""
    p = probe.cpu().numpy().astype(np.int64) if isinstance(probe, torch.Tensor) else np.asarray(probe, np.int64)
    m = match.cpu().numpy().astype(np.int64) if isinstance(match, torch.Tensor) else np.asarray(match, np.int64)
    order = np.lexsort((m, p))
    return np.stack([p[order], m[order]], axis=1)


def skewed_keys(n, distinct, seed):
    """Keys with multiplicities from 1 to ~20 (a few heavy hitters), like a join build side."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, distinct, size=n, dtype=np.int64)
    heavy = rng.integers(0, 16, size=n // 8, dtype=np.int64)
    out = np.concatenate([base[: n - heavy.size], heavy])
    rng.shuffle(out)
    return out


@pytest.fixture(autouse=True)
def reset_tuning(native_lib):
    yield
    native_lib.set_tuning(12, 0, 1, 0, 0, 1, 0)
    native_lib.set_blocking(-1, 16)


# blocked: force the L2-blocked count / retrieve (include/cuco/b200/blocked_match.cuh) with tiny 16 KiB
# regions, so that even these small tables are staged by region and probed region by region
BLOCKING = [(-1, 16), (1, -16)]


@pytest.mark.parametrize("blocking", BLOCKING)
@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("kind", MULTISET_KINDS)
def test_multiset_matches_oracle(kind, generic, blocking, native_lib):
    native_lib.set_tuning(12, 0, 1, 0, generic, 1, 0)
    native_lib.set_blocking(*blocking)
    k = cb.KINDS[kind]
    n = 20_000
    keys = skewed_keys(n, n // 4, 11)
    queries = np.concatenate([np.arange(0, n // 4, dtype=np.int64), np.arange(n, n + 3000, dtype=np.int64)])
    for lf in (0.5, 0.9):
        # sized for 2n elements: the stream is inserted twice below (overfilling never terminates)
        t = make(kind, native_lib, n=2 * n, load_factor=lf)
        ref = oracle.Table.for_kind(kind, 2 * n, lf)
        assert t.capacity() == ref.capacity()
        assert t.count(dev(queries, k.key)) == 0
        assert t.insert(dev(keys, k.key)) == ref.insert(keys) == n   # every element is stored
        assert t.size() == ref.size() == n
        dq = dev(queries, k.key)
        assert np.array_equal(t.contains(dq).cpu().numpy(), ref.contains(queries))
        assert np.array_equal(t.find(dq).cpu().numpy(), ref.find(queries))
        for outer in (False, True):
            assert t.count(dq, outer) == ref.count(queries, outer)
            assert np.array_equal(rows(*t.retrieve(dq, outer)), rows(*ref.retrieve(queries, outer)))
        # probing with the stored stream: sum of multiplicity squared
        _, counts = np.unique(keys, return_counts=True)
        assert t.count(dev(keys, k.key)) == int((counts.astype(np.int64) ** 2).sum())
        # inserting the same stream again doubles every multiplicity
        t.insert_async(dev(keys, k.key))
        assert t.size() == 2 * n
        assert t.count(dq) == 2 * ref.count(queries)
        t.close()


@pytest.mark.parametrize("kind", MULTISET_KINDS)
def test_multiset_insert_if_and_blocked_insert(kind, native_lib):
    """insert_if returns the number of stored elements; the L2-blocked insert path (route +
    region-ordered probe), forced on with small regions, stores duplicates too."""
    k = cb.KINDS[kind]
    n = 60_000
    keys = skewed_keys(n, n // 3, 12)
    stencil = (np.arange(n) % 3 != 0)
    queries = np.arange(0, n // 3, dtype=np.int64)
    ref = oracle.Table.for_kind(kind, n, 0.6)
    want_new = ref.insert_if(keys, stencil)
    want_count = ref.count(queries)
    want_rows = rows(*ref.retrieve(queries))
    for blocking in ((0, 16), (1, -64)):
        native_lib.set_blocking(*blocking)
        t = make(kind, native_lib, n=n, load_factor=0.6)
        assert t.insert_if(dev(keys, k.key), torch.from_numpy(stencil).to("cuda")) == want_new == int(stencil.sum())
        assert t.size() == want_new
        assert t.count(dev(queries, k.key)) == want_count
        assert np.array_equal(rows(*t.retrieve(dev(queries, k.key))), want_rows)
        t.close()


@pytest.mark.parametrize("blocking", BLOCKING)
@pytest.mark.parametrize("kind", SET_KINDS)
def test_set_retrieve_matches_oracle(kind, blocking, native_lib):
    native_lib.set_blocking(*blocking)
    k = cb.KINDS[kind]
    n = 30_000
    keys = np.random.default_rng(13).integers(0, n, size=n, dtype=np.int64)
    queries = np.random.default_rng(14).integers(0, 2 * n, size=2 * n, dtype=np.int64)  # duplicates + misses
    t = make(kind, native_lib, n=n, load_factor=0.5)
    ref = oracle.Table.for_kind(kind, n, 0.5)
    p, m = t.retrieve(dev(queries, k.key))
    assert p.numel() == 0 and m.numel() == 0
    t.insert(dev(keys, k.key))
    ref.insert(keys)
    got = rows(*t.retrieve(dev(queries, k.key)))
    want = rows(*ref.retrieve(queries))
    assert np.array_equal(got, want)
    assert np.array_equal(got[:, 0], got[:, 1])
    assert got.shape[0] == int(np.isin(queries, keys).sum())
    t.close()


@pytest.mark.parametrize("blocking", BLOCKING)
@pytest.mark.parametrize("kind", MULTISET_KINDS + SET_KINDS)
def test_matches_equal_cuco_itself(kind, blocking, native_lib, reference_lib):
    """Same calls into our build and into cuco's own build of the same shim."""
    native_lib.set_blocking(*blocking)
    k = cb.KINDS[kind]
    n = 40_000
    keys = skewed_keys(n, n // 5, 15)
    queries = np.concatenate([keys[::3], np.arange(n, n + 5000, dtype=np.int64)])
    stencil = (np.arange(n) % 4 != 1)
    results = {}
    for name, lib in (("ours", native_lib), ("cuco", reference_lib)):
        t = make(kind, lib, n=n, load_factor=0.7)
        dq = dev(queries, k.key)
        r = {"capacity": t.capacity()}
        r["insert_if"] = t.insert_if(dev(keys, k.key), torch.from_numpy(stencil).to("cuda"))
        r["size1"] = t.size()
        t.insert_async(dev(keys[~stencil], k.key))
        r["size2"] = t.size()
        r["contains"] = t.contains(dq).cpu().numpy()
        r["find"] = t.find(dq).cpu().numpy()
        if k.multi:
            for outer in (False, True):
                r[f"count{int(outer)}"] = t.count(dq, outer)
                r[f"rows{int(outer)}"] = rows(*t.retrieve(dq, outer))
        else:
            r["rows"] = rows(*t.retrieve(dq))
        results[name] = r
        t.close()
    for key, ours in results["ours"].items():
        theirs = results["cuco"][key]
        assert np.array_equal(ours, theirs), key


@pytest.mark.parametrize("kind", MULTIMAP_KINDS)
def test_multimap_matches_oracle_and_cuco(kind, native_lib, reference_lib):
    """experimental::static_multimap (tests/static_multimap/{insert_contains,insert_if,count}_test.cu):
    our build, cuco's build of the same shim and the CPU oracle on the same inputs; the direct and the
    L2-blocked insert path."""
    k = cb.KINDS[kind]
    n = 50_000
    keys = skewed_keys(n, n // 5, 16)
    vals = keys * 5 + 2
    queries = np.concatenate([np.arange(0, n // 5, dtype=np.int64), np.arange(n, n + 4000, dtype=np.int64)])
    stencil = (np.arange(n) % 3 != 1)
    qst = (np.arange(queries.size) % 2 == 0)
    ref = oracle.Table.for_kind(kind, 2 * n, 0.7)
    want = {"capacity": ref.capacity(), "insert_if": ref.insert_if(keys, stencil, vals)}
    want["count1"] = ref.count(queries)
    ref.insert(keys, vals)
    want["contains"] = ref.contains(queries)
    want["contains_if"] = ref.contains(queries, qst)
    want["count2"] = ref.count(queries)
    want["count_self"] = ref.count(keys)
    assert want["insert_if"] == int(stencil.sum()) and want["count2"] == want["count1"] + n
    for name, lib, blocking in (("ours", native_lib, (0, 16)), ("ours-blocked", native_lib, (1, -64)),
                                ("cuco", reference_lib, None)):
        if blocking:
            native_lib.set_blocking(*blocking)
        t = make(kind, lib, n=2 * n, load_factor=0.7)
        dk, dv, dq = dev(keys, k.key), dev(vals, k.value), dev(queries, k.key)
        got = {"capacity": t.capacity()}
        got["insert_if"] = t.insert_if(dk, torch.from_numpy(stencil).to("cuda"), dv)
        got["count1"] = t.count(dq)
        assert t.insert(dk, dv) == n
        got["contains"] = t.contains(dq).cpu().numpy()
        got["contains_if"] = t.contains_if(dq, torch.from_numpy(qst).to("cuda")).cpu().numpy()
        got["count2"] = t.count(dq)
        got["count_self"] = t.count(dk)
        for key, value in want.items():
            assert np.array_equal(got[key], value), (name, key)
        t.close()


@pytest.mark.skipif(not GOLDEN.exists(), reason="fixtures not recorded yet (tools/make_golden_matches.py)")
@pytest.mark.parametrize("kind", MULTISET_KINDS + SET_KINDS + MULTIMAP_KINDS)
def test_native_matches_golden_fixtures(kind, native_lib):
    from tools.make_golden_matches import run_kind
    g = np.load(GOLDEN)
    got = run_kind(kind, native_lib)
    mine = [name for name in g.files if name.startswith(f"k{kind}_")]
    if not mine:
        pytest.skip(f"kind {kind} not in the recorded fixtures yet")
    for name in mine:
        assert np.array_equal(got[name], g[name]), name


@pytest.mark.parametrize("blocking", BLOCKING)
def test_empty_inputs_and_high_multiplicity(blocking, native_lib):
    """n == 0 returns without launching (impl.cuh:335 convention); a key stored more often than the
    retrieve kernel keeps matches in registers takes the second-walk path."""
    native_lib.set_blocking(*blocking)
    kind = _cabi.MULTISET_I64_LP1_W2
    k = cb.KINDS[kind]
    t = make(kind, native_lib, capacity=4096)
    empty = torch.empty(0, dtype=k.key, device="cuda")
    assert t.count(empty) == 0 and t.count(empty, outer=True) == 0
    p, m = t.retrieve(empty)
    assert p.numel() == 0 and m.numel() == 0
    assert t.insert(empty) == 0 and t.size() == 0
    mult = np.array([1, 2, 3, 4, 5, 9, 33, 100], dtype=np.int64)          # key i stored mult[i] times
    keys = np.repeat(np.arange(mult.size, dtype=np.int64), mult)
    np.random.default_rng(3).shuffle(keys)
    t.insert(dev(keys, k.key))
    probes = np.arange(0, mult.size + 3, dtype=np.int64)
    assert t.count(dev(probes, k.key)) == int(mult.sum())
    p, m = t.retrieve(dev(probes, k.key))
    assert torch.equal(p, m)
    assert np.array_equal(np.bincount(p.cpu().numpy(), minlength=probes.size)[: mult.size], mult)
    p, m = t.retrieve(dev(probes, k.key), outer=True)
    assert p.numel() == int(mult.sum()) + 3 and int((m == -1).sum().item()) == 3
    # one hot probe key: with the blocked path its region's segment overflows - first into the spill
    # list (100 k probes), then past it (300 k probes: the whole batch is redone by the direct kernel)
    for hot in (100_000, 300_000):
        batch = torch.cat([torch.full((hot,), 6, dtype=k.key, device="cuda"), dev(probes, k.key)])
        assert t.count(batch) == hot * 33 + int(mult.sum())
        p, m = t.retrieve(batch[: hot // 10 + 11])
        assert torch.equal(p, m) and p.numel() == (hot // 10 + 11) * 33
    t.close()
    s = make(_cabi.SET_I64_DH4, native_lib, capacity=1000)
    p, m = s.retrieve(empty)
    assert p.numel() == 0
    s.close()


def test_multiset_large_input_properties(native_lib):
    """Size-independent properties at a size the oracle does not reach (large_input_test.cu style):
    40 M elements, each key stored 4 times."""
    kind = _cabi.MULTISET_I64_LP1_W2
    n, mult = 40_000_000, 4
    keys = (torch.randperm(n, device="cuda") // mult).to(torch.int64)
    t = make(kind, native_lib, n=n, load_factor=0.5)
    t.insert_async(keys)
    assert t.size() == n
    probes = torch.arange(0, 2 * (n // mult), device="cuda", dtype=torch.int64)
    assert t.count(probes) == n
    assert t.count(probes, outer=True) == n + n // mult
    assert bool(t.contains(probes[: n // mult]).all()) and not bool(t.contains(probes[n // mult:]).any())
    sample = probes[: 1_000_000]
    p, m = t.retrieve(sample)
    assert p.numel() == mult * sample.numel()
    assert torch.equal(p, m)
    assert torch.equal(torch.sort(p).values, torch.repeat_interleave(sample, mult))
    t.close()
