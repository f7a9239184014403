"""CPU, world_size 2, gloo: the exchange plumbing of the hash-partitioned table
(cucollections_b200/partitioned.py) with a stand-in backend - numpy routing with the same owner
function as the CUDA kernels and the C oracle as each rank's local table. Checks that results are
partition invariant: identical to one table holding the union of both ranks' keys."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

MASK = (1 << 64) - 1


def mix64(x: int) -> int:
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & MASK
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & MASK
    x ^= x >> 33
    return x


def owner_of(key: int, salt: int, parts: int) -> int:
    """Same arithmetic as owner_of() in cucollections_b200/csrc/cabi_core.cu."""
    return (mix64((key & MASK) ^ salt) * parts) >> 64


class _OracleTable:
    """Adapts oracle.Table to the slice of the static_map API partitioned.py uses."""

    def __init__(self, n, load_factor):
        from oracle import oracle
        self.t = oracle.Table.for_kind(1, n, load_factor)

    def insert(self, pairs):
        a = pairs.numpy()
        return self.t.insert(a[:, 0], a[:, 1])

    def insert_async(self, pairs):
        self.insert(pairs)

    def insert_or_apply(self, pairs, op="plus", init=None):
        from oracle import oracle
        a = pairs.numpy()
        self.t.insert_or_apply(a[:, 0], a[:, 1], {"plus": oracle.PLUS, "min": oracle.MIN, "max": oracle.MAX}[op], init)

    def find(self, keys):
        return torch.from_numpy(self.t.find(keys.numpy()))

    def contains(self, keys):
        return torch.from_numpy(self.t.contains(keys.numpy()))

    def size(self):
        return self.t.size()

    def _payload_dtype(self):
        return torch.int64

    def clear_async(self):
        self.t.clear()

    def close(self):
        pass


class CpuBackend:
    device = torch.device("cpu")

    def make_table(self, n_local, load_factor, **kw):
        return _OracleTable(n_local, load_factor)

    def partition(self, elems, num_parts, salt, pair_aos, want_index):
        keys = (elems[:, 0] if pair_aos else elems).numpy()
        owners = np.array([owner_of(int(k), salt, num_parts) for k in keys], dtype=np.int64)
        order = np.argsort(owners, kind="stable")
        counts = np.bincount(owners, minlength=num_parts).tolist()
        routed = elems[torch.from_numpy(order)]
        return routed.contiguous(), counts, (torch.from_numpy(order) if want_index else None)

    def unpermute(self, values, index, out):
        out[index] = values
        return out

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cucollections_b200.partitioned import partitioned_static_map
        rng = np.random.default_rng(100 + rank)
        keys = rng.integers(1, 3 * n, size=n, dtype=np.int64)          # overlaps between ranks
        pairs = torch.from_numpy(np.stack([keys, keys * 2 + 1], axis=1).copy())
        table = partitioned_static_map(n * world, 0.5, backend=CpuBackend())
        new_keys = table.insert(pairs)
        total = table.size()
        queries = torch.from_numpy(np.concatenate([keys[: n // 2], rng.integers(3 * n, 4 * n, size=n // 2)]))
        found = table.find(queries)
        present = table.contains(queries)
        # aggregate variant on a second table: sum of ones per key over both ranks
        agg = partitioned_static_map(n * world, 0.5, backend=CpuBackend())
        agg.insert_or_apply(torch.from_numpy(np.stack([keys % 50, np.ones(n, dtype=np.int64)], axis=1).copy()), op="plus")
        sums = agg.find(torch.arange(50, dtype=torch.int64))
        queue.put((rank, keys, new_keys, total, queries.numpy(), found.numpy(), present.numpy(), sums.numpy(),
                   table.local_size()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_partitioned_table_is_partition_invariant():
    world, n = 2, 4000
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((queue.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    union = {}
    for _, keys, *_ in results:
        for k in keys:
            union.setdefault(int(k), int(k) * 2 + 1)
    counts = {}
    for _, keys, *_ in results:
        for k in keys:
            counts[int(k) % 50] = counts.get(int(k) % 50, 0) + 1
    local_sizes = []
    for rank, keys, new_keys, total, queries, found, present, sums, local_size in results:
        assert new_keys == len(union)            # all-reduced number of new keys
        assert total == len(union)               # global size()
        want_found = np.array([union.get(int(q), -1) for q in queries])
        assert np.array_equal(found, want_found), rank
        assert np.array_equal(present, want_found != -1), rank
        assert np.array_equal(sums, np.array([counts.get(i, 0) if counts.get(i, 0) else -1 for i in range(50)])), rank
        local_sizes.append(local_size)
    assert sum(local_sizes) == len(union)
    assert min(local_sizes) > 0.4 * len(union)   # the owner hash balances the shards


def test_owner_function_is_balanced_and_independent_of_table_hash():
    parts = 8
    keys = np.arange(1, 80_001)
    owners = np.array([owner_of(int(k), 0x9E3779B97F4A7C15, parts) for k in keys])
    hist = np.bincount(owners, minlength=parts)
    assert hist.min() > 0.9 * len(keys) / parts and hist.max() < 1.1 * len(keys) / parts
