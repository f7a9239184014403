#!/usr/bin/env python3
"""Multi-GPU parity check of the hash-partitioned table (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/multi_gpu_check.py [n_per_rank]

Every rank inserts its own batch (overlapping key ranges, duplicates) through the staged exchange, the
fused P2P exchange and the all_to_all path into separate partitioned tables, and looks up a mixed hit/miss
batch in both. Rank 0 also builds ONE single-GPU table over the union of all batches; results are
partition invariant, so per-key find / contains outputs and the total size must be identical.
Prints one line per check and "MULTI_GPU_CHECK PASS" / "FAIL"; exit code 0 iff everything passed."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cucollections_b200 as cb  # noqa: E402
from cucollections_b200 import partitioned  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def check(cond, what):
        nonlocal ok
        cond = bool(cond)
        ok = ok and cond
        if rank == 0 or not cond:
            print(f"[rank {rank}] {'PASS' if cond else 'FAIL'} {what}", flush=True)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    keys = torch.randint(1, 3 * n, (n,), generator=gen, device=dev, dtype=torch.int64)  # overlaps across ranks
    pairs = torch.stack([keys, keys * 3 + 1], dim=1).contiguous()
    queries = torch.cat([keys[: n // 2], torch.randint(3 * n, 6 * n, (n - n // 2,), generator=gen, device=dev)])

    tables = {
        "staged": partitioned.partitioned_static_map(n * world, 0.5, backend=partitioned.GpuBackend(dev),
                                                     fused_batch=n, routing="staged", probing="linear_probing",
                                                     cg_size=1),
        "fused": partitioned.partitioned_static_map(n * world, 0.5, backend=partitioned.GpuBackend(dev),
                                                    fused_batch=n, routing="fused", probing="linear_probing",
                                                    cg_size=1),
        "fused, 2 pipelined lanes": partitioned.partitioned_static_map(
            n * world, 0.5, backend=partitioned.GpuBackend(dev), fused_batch=n, fused_lanes=2, routing="fused",
            probing="linear_probing", cg_size=1),
        "nccl": partitioned.partitioned_static_map(n * world, 0.5, backend=partitioned.GpuBackend(dev),
                                                   probing="linear_probing", cg_size=1),
    }
    results = {}
    for name, t in tables.items():
        t.insert_async(pairs)
        found = t.find(queries)
        present = t.contains(queries)
        torch.cuda.synchronize(dev)
        results[name] = (found.clone(), present.clone(), t.size())
    for name in ("staged", "fused, 2 pipelined lanes"):
        check(torch.equal(results[name][0], results["nccl"][0]), f"{name}: find == all_to_all find")
        check(torch.equal(results[name][1], results["nccl"][1]), f"{name}: contains == all_to_all contains")
        check(results[name][2] == results["nccl"][2], f"{name}: total size")
    check(torch.equal(results["fused"][0], results["nccl"][0]), "fused find == all_to_all find")
    check(torch.equal(results["fused"][1], results["nccl"][1]), "fused contains == all_to_all contains")
    check(results["fused"][2] == results["nccl"][2], f"total size {results['fused'][2]} == {results['nccl'][2]}")

    # a second round through the same buffers (reuse across calls: barriers / lane hand-over)
    more = torch.stack([keys + 7 * n, keys], dim=1).contiguous()
    sizes = []
    for name, t in tables.items():
        t.insert_async(more)
        sizes.append(t.size())
        again = t.find(queries)
        check(torch.equal(again, results[name][0]), f"{name}: second round leaves earlier keys intact")
    check(len(set(sizes)) == 1, f"sizes after the second round agree: {sizes}")

    # the union on ONE GPU (rank 0), then every rank checks its own queries against it
    all_pairs = [torch.empty_like(pairs) for _ in range(world)]
    dist.all_gather(all_pairs, pairs)
    if rank == 0:
        # the single table is cuco's OWN build when oracle/_ref/libcuco_ref.so travelled with the snapshot
        from cucollections_b200 import _cabi
        try:
            judge, judge_name = _cabi.reference(), "cuco (libcuco_ref.so)"
        except (FileNotFoundError, OSError):
            judge, judge_name = None, "this repository's single-GPU table"
        print(f"[rank 0] single-table judge: {judge_name}", flush=True)
        single = cb.static_map(n=n * world, load_factor=0.5, probing="linear_probing", cg_size=1, device=dev,
                               _library=judge)
        single.insert_async(torch.cat(all_pairs))
        size = torch.tensor([single.size()], device=dev)
    else:
        single, size = None, torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(size, 0)
    check(results["staged"][2] == int(size.item()), f"partitioned size == single-table size {int(size.item())}")
    all_queries = [torch.empty_like(queries) for _ in range(world)]
    dist.all_gather(all_queries, queries)
    all_found = [torch.empty_like(results["staged"][0]) for _ in range(world)]
    dist.all_gather(all_found, results["staged"][0])
    all_present = [torch.empty_like(results["staged"][1]) for _ in range(world)]
    dist.all_gather(all_present, results["staged"][1])
    if rank == 0:
        for r in range(world):
            check(torch.equal(single.find(all_queries[r]), all_found[r]), f"rank {r} find == single table")
            check(torch.equal(single.contains(all_queries[r]), all_present[r]), f"rank {r} contains == single table")

    # aggregate variant: count occurrences of key % 1000 over all ranks
    agg = partitioned.partitioned_static_map(4000, 0.5, backend=partitioned.GpuBackend(dev), fused_batch=n,
                                             routing="staged", empty_value=0, probing="linear_probing", cg_size=1)
    ones = torch.stack([keys % 1000, torch.ones_like(keys)], dim=1).contiguous()
    agg.insert_or_apply(ones, op="plus")
    sums = agg.find(torch.arange(1000, device=dev, dtype=torch.int64))
    hist = torch.bincount(keys % 1000, minlength=1000)
    dist.all_reduce(hist)
    check(torch.equal(sums, hist), "insert_or_apply(plus) over all ranks == global histogram (skewed batch, spill path)")
    check(int(sums.sum().item()) == n * world, "every row counted exactly once")

    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) == 0 else "FAIL", flush=True)
    for t in list(tables.values()) + [agg]:
        t.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
