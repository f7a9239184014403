"""CPU side of the fixtures cuCollections itself produced on a B200 for `static_set::retrieve` and
`static_multiset` (tests/golden/cuco_golden_matches.npz, written by tools/make_golden_matches.py
through oracle/_ref/libcuco_ref.so): the C oracle must reproduce every recorded result, which pins
its count / retrieve / allows_duplicates restatement to the real reference. The GPU side (native
kernels against the same file) is tests/test_matches_gpu.py::test_native_matches_golden_fixtures.

Independently of the fixtures, the oracle's multiset semantics are checked against a plain Python
model (collections.Counter), which is what the reference's own static_multiset tests assert.
"""
from collections import Counter
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tools.make_golden_matches import KINDS, run_kind_oracle

GOLDEN = Path(__file__).resolve().parent / "golden" / "cuco_golden_matches.npz"


@pytest.mark.skipif(not GOLDEN.exists(), reason="fixtures not recorded yet (tools/make_golden_matches.py)")
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reproduces_cuco_fixtures(kind):
    g = np.load(GOLDEN)
    got = run_kind_oracle(kind)
    mine = [name for name in g.files if name.startswith(f"k{kind}_")]
    if not mine:
        pytest.skip(f"kind {kind} not in the recorded fixtures yet")
    assert sorted(mine) == sorted(got)
    for name in mine:
        assert np.array_equal(got[name], g[name]), name


@pytest.mark.parametrize("kind", [10, 11])
@pytest.mark.parametrize("load_factor", [0.5, 0.95])
def test_oracle_multiset_against_counter_model(kind, load_factor):
    rng = np.random.default_rng(kind)
    n = 5000
    keys = rng.integers(0, n // 6, size=n, dtype=np.int64)
    model = Counter(keys.tolist())
    t = oracle.Table.for_kind(kind, n, load_factor)
    assert t.insert(keys) == n            # static_multiset/insert_test.cu: every element is stored
    assert t.size() == n
    queries = np.arange(0, n // 3, dtype=np.int64)   # upper half absent
    want = np.array([model.get(int(q), 0) for q in queries])
    assert t.count(queries) == int(want.sum())                      # count_test.cu
    assert t.count(queries, outer=True) == int(np.maximum(want, 1).sum())
    assert np.array_equal(t.contains(queries), want > 0)            # contains_test.cu
    found = t.find(queries)                                         # find_test.cu
    assert np.array_equal(found, np.where(want > 0, queries, -1))
    probe, match = t.retrieve(queries)                              # retrieve_test.cu
    assert np.array_equal(probe, match)
    assert Counter(probe.tolist()) == Counter({k: c for k, c in model.items() if k < n // 3})
    probe, match = t.retrieve(queries, outer=True)
    assert int((match == -1).sum()) == int((want == 0).sum())
    assert np.array_equal(np.sort(probe[match == -1]), queries[want == 0])


@pytest.mark.parametrize("kind", [0, 5])
def test_oracle_set_retrieve_is_a_semi_join(kind):
    rng = np.random.default_rng(kind + 100)
    keys = rng.integers(0, 3000, size=3000, dtype=np.int64)
    queries = rng.integers(0, 6000, size=5000, dtype=np.int64)
    t = oracle.Table.for_kind(kind, 3000, 0.5)
    t.insert(keys)
    probe, match = t.retrieve(queries)
    assert np.array_equal(probe, match)
    assert np.array_equal(probe, queries[np.isin(queries, keys)])   # input order, one row per hit


def test_oracle_multimap_against_counter_model():
    """experimental::static_multimap semantics (tests/static_multimap/{insert_contains,insert_if,
    count}_test.cu): every pair is stored, count sums multiplicities, contains is key presence."""
    rng = np.random.default_rng(12)
    n = 4000
    keys = rng.integers(0, n // 5, size=n, dtype=np.int64)
    model = Counter(keys.tolist())
    t = oracle.Table.for_kind(12, n, 0.8)
    assert t.insert(keys, keys * 2) == n
    queries = np.arange(0, n // 2, dtype=np.int64)
    want = np.array([model.get(int(q), 0) for q in queries])
    assert t.count(queries) == int(want.sum()) == n
    assert np.array_equal(t.contains(queries), want > 0)
    probe, mk, mv = t.retrieve(queries)
    assert np.array_equal(probe, mk) and np.array_equal(mv, mk * 2)
    # the same pair over and over (insert_if_test.cu, "same element n / 2 times")
    t = oracle.Table.for_kind(12, n, 0.8)
    assert t.insert_if(np.ones(n, dtype=np.int64), np.arange(n) % 2 == 0, np.ones(n, dtype=np.int64)) == n // 2
    assert t.count(queries) == n // 2


def test_multiset_insert_if_counts_every_selected_element():
    t = oracle.Table.for_kind(10, 1000, 0.5)
    keys = np.zeros(1000, dtype=np.int64) + 7
    stencil = np.arange(1000) % 2
    assert t.insert_if(keys, stencil) == 500
    assert t.count(np.array([7])) == 500
