"""Staged exchange of the hash-partitioned table (cucollections_b200/partitioned.py::StagedExchange,
C ABI cuco_b200_exchange_stage / _publish / _copy_async / _apply / _lookup_local / _unpermute) with
every 'rank' living on ONE GPU: the peer pointers are ordinary device pointers and the cross-rank
barrier is stream order, so exactly the kernels and copies of the multi-process path run, step by
step. The union of the shards must behave like ONE table - the CPU oracle - holding every rank's
batch: per-key find / contains, total size, insert_or_apply sums. (The multi-process version runs
under torchrun: tests/multi_gpu_check.py and bench.py's parity gate.)"""
import numpy as np
import pytest
import torch

import cucollections_b200 as cb
from cucollections_b200 import _cabi
from cucollections_b200 import partitioned as cbp
from oracle import oracle

pytestmark = pytest.mark.gpu


class SimTransport:
    """All ranks in one process: a shared registry of buffer addresses, barriers are no-ops (the test
    drives the ranks in lock-step on one stream)."""

    def __init__(self, registry, P, me):
        self.registry, self.P, self.me = registry, P, me

    def allocate(self, nbytes):
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        self.registry[self.me] = self.buf.data_ptr()
        return self.buf

    def peer_ptrs(self):
        assert all(p is not None for p in self.registry)
        return list(self.registry)

    def barrier(self, channel):
        pass

    def ready(self):
        torch.cuda.synchronize()


def make(kind, lib, **kw):
    k = cb.KINDS[kind]
    common = dict(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size, window_size=k.window_size,
                  hash=k.hash, _library=lib, **kw)
    if k.value is None:
        return cb.static_set(**common)
    return cb.static_map(value_dtype=k.value, **common)


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype)


def keyset(kind, n, seed, hi):
    return np.random.default_rng(seed).integers(0, hi, size=n, dtype=np.int64)


def build_ranks(lib, kind, P, n, slices, lanes, mode="auto", **table_kw):
    registry = [None] * P
    ranks = []
    for me in range(P):
        table = make(kind, lib, n=n + n // 8 + 64, load_factor=0.5, **table_kw)
        ranks.append(cbp.StagedExchange(table, n, None, "cuda", cbp.DEFAULT_SALT, slices=slices, lanes=lanes,
                                        transport=SimTransport(registry, P, me), mode=mode))
    return ranks


def mutate_all(ranks, batches, op=-1):
    for r, pairs in zip(ranks, batches):
        r.stage_pairs(pairs)
    for r in ranks:
        r.publish_pairs()
        for g in range(r.G):
            r.send_slice(g)
    for r in ranks:
        for g in range(r.G):
            r.apply_slice(g, op)
    torch.cuda.synchronize()
    return [r._spilled(mutation=True) for r in ranks]


def lookup_all(ranks, queries, outs, what):
    for c in range(ranks[0].L):
        chunks = [r._chunks(q.shape[0]) + [(q.shape[0], q.shape[0])] * r.L for r, q in zip(ranks, queries)]
        for r, q, ch in zip(ranks, queries, chunks):
            lo, hi = ch[c]
            r.stage_keys(r.lanes[c], q[lo:hi], lo)
        for r in ranks:
            r.send_keys(r.lanes[c])
        for r in ranks:
            r.answer_keys(r.lanes[c], what)
        for r in ranks:
            r.return_results(r.lanes[c], what)
        for r, o, ch in zip(ranks, outs, chunks):
            r.unpermute(r.lanes[c], what, o, *ch[c])
    torch.cuda.synchronize()
    return [r._spilled(mutation=False) for r in ranks]


@pytest.mark.parametrize("mode", ["coarse", "fine"])
@pytest.mark.parametrize("kind,P,slices,lanes,region_kib", [
    (_cabi.MAP_I64_LP1, 3, 4, 2, 16), (_cabi.MAP_I64_DH8, 2, 2, 1, 64), (_cabi.MAP_I32_LP4, 4, 3, 3, 16),
    (_cabi.SET_I64_DH4, 2, 2, 2, 16), (_cabi.MAP_I64_LP1_X64, 8, 5, 4, 4), (_cabi.MAP_I64_LP1, 1, 2, 2, 16)])
def test_staged_exchange_with_simulated_ranks(kind, P, slices, lanes, region_kib, mode, native_lib, monkeypatch):
    # blocks travel by cudaMemcpyAsync in the coarse runs and by the push kernel in the fine ones
    monkeypatch.setenv("CUCO_B200_PUSH_MIB", "0" if mode == "coarse" else "8")
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 60_000  # per rank
    try:
        native_lib.set_blocking(1, -region_kib)  # small regions: every slice is regrouped and probed blocked
        ranks = build_ranks(native_lib, kind, P, n, slices, lanes, mode)
        assert all(r.fine == (mode == "fine") for r in ranks)
        ref = oracle.Table.for_kind(kind, 2 * n * P, 0.0)
        batches, raw = [], []
        for me in range(P):
            keys = keyset(kind, n - 17 * me, 40 + me, hi=2 * n)  # ragged batch sizes, overlaps between ranks
            vals = keys * 7 + 3
            raw.append(keys)
            ref.insert(keys, vals if is_map else None)
            batches.append(dev(np.stack([keys, vals], axis=1), k.key) if is_map else dev(keys, k.key))
        spilled = mutate_all(ranks, batches)
        assert all(total == 0 for total, _ in spilled)
        assert sum(r.table.size() for r in ranks) == ref.size()
        sizes = [r.table.size() for r in ranks]
        assert min(sizes) > 0.7 * ref.size() / P  # the owner hash balances the shards
        # a second pass through the same buffers changes nothing
        mutate_all(ranks, batches)
        assert sum(r.table.size() for r in ranks) == ref.size()
        queries = [np.concatenate([raw[me][: n // 2], keyset(kind, n // 3 + me, 60 + me, hi=2 * n) + 4 * n])
                   for me in range(P)]
        dq = [dev(q, k.key) for q in queries]
        for what in (0, 1):
            if what == 0:
                outs = [torch.full((q.shape[0],), -7, dtype=k.value if is_map else k.key, device="cuda") for q in queries]
            else:
                outs = [torch.full((q.shape[0],), 9, dtype=torch.uint8, device="cuda") for q in queries]
            spilled = lookup_all(ranks, dq, outs, what)
            assert all(total == 0 for total, _ in spilled)
            for me in range(P):
                got = outs[me].cpu().numpy()
                if what == 0:
                    assert np.array_equal(got.astype(np.int64), ref.find(queries[me])), (what, me)
                else:
                    assert np.array_equal(got.astype(bool), ref.contains(queries[me])), (what, me)
        if is_map:  # aggregate variant: occurrences of every key over all ranks' batches
            for r in ranks:
                r.table.close()
            agg = oracle.Table.for_kind(kind, 2 * n * P, 0.0, empty_value=0)
            ranks = build_ranks(native_lib, kind, P, n, slices, lanes, mode, empty_value=0)
            ones = []
            for me in range(P):
                one = np.ones(raw[me].shape[0], dtype=np.int64)
                agg.insert_or_apply(raw[me], one, oracle.PLUS)
                ones.append(dev(np.stack([raw[me], one], axis=1), k.key))
            mutate_all(ranks, ones, _cabi.PLUS)
            got = {}
            for r in ranks:
                ks, vs = r.table.retrieve_all()
                got.update(zip(ks.cpu().tolist(), vs.cpu().tolist()))
            wk, wv = agg.retrieve_all()
            assert got == dict(zip(wk.tolist(), wv.tolist()))
        for r in ranks:
            r.table.close()
    finally:
        native_lib.set_blocking(-1, 16)


def test_staged_exchange_spills_a_hot_key(native_lib):
    """One hot key overflows its (owner, slice) segment: the surplus lands in the spill list, the
    flags tell every rank, and the routed part still counts once."""
    P, n = 2, 40_000
    ranks = build_ranks(native_lib, _cabi.MAP_I64_LP1, P, n, 2, 1, "coarse")
    hot = np.full(n, 12345, dtype=np.int64)
    cold = np.arange(n, dtype=np.int64) + 100_000
    batches = [dev(np.stack([hot, hot], axis=1), torch.int64), dev(np.stack([cold, cold], axis=1), torch.int64)]
    spilled = mutate_all(ranks, batches)
    total, mine = spilled[0]
    assert total == n - ranks[0].cap and mine == [total]
    assert spilled[1][0] == total and spilled[1][1] == [0]
    assert sum(r.table.size() for r in ranks) == n + 1
    # lookups of a hot batch: the surplus keys are reported with their source positions
    q = [dev(hot, torch.int64), dev(cold, torch.int64)]
    outs = [torch.full((n,), -7, dtype=torch.int64, device="cuda") for _ in range(P)]
    spilled = lookup_all(ranks, q, outs, 0)
    assert spilled[0][0] == n - ranks[0].cap_l
    idx = ranks[0].lanes[0].spill_index[: spilled[0][1][0]].cpu().numpy()
    answered = np.ones(n, dtype=bool)
    answered[idx] = False
    assert np.array_equal(outs[0].cpu().numpy()[answered], hot[answered])
    assert bool((outs[0].cpu().numpy()[~answered] == -7).all())
    assert np.array_equal(outs[1].cpu().numpy(), cold)
    for r in ranks:
        r.table.close()


def test_staged_lookup_with_short_and_empty_batches(native_lib):
    """Ranks whose batch is shorter than a lane - or empty - still run every lane (empty chunks), so the
    barriers of the multi-process path match; answers stay correct for the ranks that do have keys."""
    P, n, lanes = 3, 30_000, 3
    kind = _cabi.MAP_I64_LP1
    ranks = build_ranks(native_lib, kind, P, n, 2, lanes, "coarse")
    ref = oracle.Table.for_kind(kind, 4 * n * P, 0.0)
    batches = []
    for me in range(P):
        keys = keyset(kind, [n, 11, 0][me], 70 + me, hi=2 * n)   # full, shorter than one lane, empty
        ref.insert(keys, keys + 5)
        batches.append(dev(np.stack([keys, keys + 5], axis=1).reshape(-1, 2), torch.int64))
    assert all(total == 0 for total, _ in mutate_all(ranks, batches))
    assert sum(r.table.size() for r in ranks) == ref.size()
    queries = [keyset(kind, m, 80 + me, hi=3 * n) for me, m in enumerate([0, n, 7])]
    outs = [torch.full((q.shape[0],), -7, dtype=torch.int64, device="cuda") for q in queries]
    assert all(total == 0 for total, _ in lookup_all(ranks, [dev(q, torch.int64) for q in queries], outs, 0))
    for me in range(P):
        assert np.array_equal(outs[me].cpu().numpy(), ref.find(queries[me])), me
    for r in ranks:
        r.table.close()
