#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's headline configuration.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--no-points] [--no-c4] [--no-c5]

Workload (config[1] of BASELINE.json, the one the metric is quoted on): cuco::static_map<int64,int64>,
100 M uniform key/value pairs (uniform_int[1, n] -> ~63 % distinct, value = key, fixed seed) bulk
`insert_async` into an empty table sized n / 0.5 and then bulk `find_async` of the same 100 M keys,
linear_probing<1>, xxhash_32, one slot per window, sentinels -1/-1.  One *step* = one insert pass +
one find pass over that batch.  `value` = (inserts + finds) / (insert + find kernel time), in Gops/s,
inputs already resident in HBM; the table is cleared between steps outside the timed events (the
reference's own static_set/insert_or_apply benchmarks do the same).  Both working sets
(1.6 GB of pairs, 3.2 GB of slots) are far larger than the 126 MB L2, so nothing carries over
between timed kernels.

N > 1 (torchrun, one rank per GPU): the hash-partitioned table (cucollections_b200/partitioned.py). First a
checker leg, never timed: 1 M pairs per rank through the routing under test, every rank's mixed hit / miss
find / contains and the global size() compared bit-exactly with the CPU oracle over the union of all batches
(non-zero exit on mismatch). Then the headline: every rank brings its own 100 M pairs (weak scaling), keys are
grouped by owner rank locally, delivered over NVLink by the copy engines while the owners apply what has
already landed, and lookups return the same way (staged exchange; `CUCO_B200_ROUTING=fused|nccl` select the
round-1 paths). Timing is CUDA events per rank, max over ranks. Extra legs in the same line: `c4` = BASELINE
configs[3] at its stated size (4 B pairs over the N GPUs), `c5` = configs[4] (insert_or_apply sum over 2 B
rows with 10 M distinct keys, with and without per-GPU pre-aggregation).

`--impl reference` times cuco's own headers (oracle/_ref/libcuco_ref.so: the unmodified reference
compiled for sm_100a behind the same C shim) on the same inputs with the same events - that is the
"cuco on the same B200" bar of the north star.  cuco has no CPU path; the `cpu_baseline` object is
the host std::unordered_map baseline the north star asks for, on a bounded 10 M-pair sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

N_KEYS = 100_000_000
LOAD_FACTOR = 0.5
INSERT_BYTES_PER_OP = 80.0  # 16 B pair in + 32 B sector read + 32 B sector write-back
FIND_BYTES_PER_OP = 48.0    # 8 B key in + 8 B value out + 32 B sector read


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def make_inputs(n, device, seed):
    from cucollections_b200 import key_generator as kg
    keys = kg.uniform(n, 1, torch.int64, device, seed=seed)
    pairs = torch.stack([keys, keys], dim=1).contiguous()  # value = key: result independent of winner
    return keys, pairs


def timed(fn, stream):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    fn()
    stop.record(stream)
    return start, stop


def cpu_baseline(sample_n=10_000_000):
    """Host std::unordered_map<int64,int64> baseline on a bounded sample of the same key stream."""
    import numpy as np
    from oracle import oracle
    rng = np.random.default_rng(42)
    keys = rng.integers(1, sample_n + 1, size=sample_n, dtype=np.int64)
    threads = oracle.hardware_threads()
    ti, tf, _ = oracle.baseline_map_i64(keys, keys, keys, threads, LOAD_FACTOR)
    gops = (2 * sample_n) / (ti + tf) / 1e9
    return {"value": gops, "unit": "Gops/s", "cores": threads, "kind": "port",
            "sample": f"std::unordered_map<int64,int64> sharded over {threads} host threads, "
                      f"{sample_n} uniform pairs insert + {sample_n} finds at max_load_factor {LOAD_FACTOR}",
            "insert_gops": sample_n / ti / 1e9, "find_gops": sample_n / tf / 1e9}


def run_single(args, lib, impl):
    """One GPU, table resident on it. Returns the JSON dict."""
    import cucollections_b200 as cb

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream(dev)
    n = args.n
    keys, pairs = make_inputs(n, dev, seed=42)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    table = cb.static_map(n=n, load_factor=LOAD_FACTOR, probing="linear_probing", cg_size=1,
                          device=dev, _library=lib)

    def step():
        table.clear_async()
        ei = timed(lambda: table.insert_async(pairs), stream)
        ef = timed(lambda: table.find(keys, out), stream)
        return ei, ef

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    # correctness gate on the warm-up result: every inserted key is found with value == key
    assert bool((out == keys).all().item()), "find returned a wrong payload"

    with ClockSampler(dev.index) as clocks:
        torch.cuda.synchronize(dev)
        events = [step() for _ in range(args.steps)]
        torch.cuda.synchronize(dev)
    t_ins = [a.elapsed_time(b) for (a, b), _ in events]
    t_find = [a.elapsed_time(b) for _, (a, b) in events]
    ms_ins, ms_find = sum(t_ins) / len(t_ins), sum(t_find) / len(t_find)
    ms_step = ms_ins + ms_find
    value = 2 * n / (ms_step * 1e-3) / 1e9
    spread = {"insert_ms_median": statistics.median(t_ins), "insert_ms_best": min(t_ins),
              "find_ms_median": statistics.median(t_find), "find_ms_best": min(t_find),
              "value_median": 2 * n / ((statistics.median(t_ins) + statistics.median(t_find)) * 1e-3) / 1e9,
              "value_best": 2 * n / ((min(t_ins) + min(t_find)) * 1e-3) / 1e9}

    # ---- end to end: host pinned buffers in, host results out, copies inside the timed region ----
    h_pairs = torch.empty((n, 2), dtype=torch.int64, pin_memory=True)
    h_pairs.copy_(pairs)
    h_keys = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_keys.copy_(keys)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    d_pairs, d_keys = torch.empty_like(pairs), torch.empty_like(keys)

    def e2e_step():
        table.clear_async()
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(stream)
        if impl == "native":
            # the product's host-buffer calls: chunked, uploads / kernels / downloads overlapped
            table.insert_host(h_pairs)
            table.find_host(h_keys, h_out)
        else:
            # what a cuco user writes: copy in, bulk call, copy out (cuco takes device ranges only)
            d_pairs.copy_(h_pairs, non_blocking=True)
            table.insert_async(d_pairs)
            d_keys.copy_(h_keys, non_blocking=True)
            table.find(d_keys, out)
            h_out.copy_(out, non_blocking=True)
        stop.record(stream)
        return start, stop

    h_out.zero_()
    e2e_step()
    torch.cuda.synchronize(dev)
    assert bool((h_out == h_keys).all().item()), "end-to-end find returned a wrong payload"
    e2e_events = []
    for _ in range(max(1, min(args.steps, 3))):
        e2e_events.append(e2e_step())
        torch.cuda.synchronize(dev)
    e2e_ms = statistics.mean(a.elapsed_time(b) for a, b in e2e_events)
    assert bool((h_out == h_keys).all().item())
    e2e = {"value": 2 * n / (e2e_ms * 1e-3) / 1e9, "unit": "Gops/s",
           "h2d_bytes_per_step": int(h_pairs.numel() * 8 + h_keys.numel() * 8),
           "d2h_bytes_per_step": int(h_out.numel() * 8), "ms_per_step": e2e_ms}

    # our kernels per step: blocked insert = route + probe, direct insert = 1; find = 1
    table_bytes = table.capacity() * 16
    blocked = (impl == "native" and os.environ.get("CUCO_B200_BLOCKED", "-1") != "0"
               and table_bytes >= (256 << 20) and n >= (1 << 22) and n * 64 >= table_bytes)
    if impl == "native" and os.environ.get("CUCO_B200_BLOCKED") == "1":
        blocked = True
    launches_per_step = 3 if blocked else 2
    peak, peak_src = measured_hbm_peak()
    ins_gbs = INSERT_BYTES_PER_OP * n / (ms_ins * 1e-3) / 1e9
    find_gbs = FIND_BYTES_PER_OP * n / (ms_find * 1e-3) / 1e9
    # measured DRAM traffic per launch of the same kernels (ncu --set full capture under profiles/)
    traffic = {}
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(impl, {})
        except Exception:
            traffic = {}
    if impl == "native":
        find_traffic = traffic.get("lookup_kernel")
        router = "tile_route_kernel" if "tile_route_kernel" in traffic else "route_kernel"
        ins_traffic = (traffic.get(router, 0) + traffic.get("blocked_mutate_kernel", 0)) if blocked else None
        find_kernel = "find: lookup_kernel (one launch = the whole find pass)"
        ins_kernel = (f"insert: {router} + blocked_mutate_kernel (two launches)" if blocked
                      else "insert: mutate_kernel")
    else:
        find_traffic, ins_traffic = traffic.get("find"), traffic.get("insert_if_n")
        find_kernel, ins_kernel = "find: cuco::detail::find", "insert: cuco::detail::insert_if_n"
    find_roof = {"bound": "hbm", "kernel": find_kernel, "achieved": find_gbs, "peak": peak,
                 "peak_source": peak_src, "unit": "GB/s", "frac": find_gbs / peak, "traffic": find_traffic,
                 "algorithmic_bytes_per_op": FIND_BYTES_PER_OP, "launch_ms": ms_find}
    ins_roof = {"bound": "hbm", "kernel": ins_kernel, "achieved": ins_gbs, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": ins_gbs / peak, "traffic": ins_traffic,
                "algorithmic_bytes_per_op": INSERT_BYTES_PER_OP, "launch_ms": ms_ins}
    # Dominant single kernel of the step. The find pass is one launch; the blocked insert is two, whose
    # shares of the pass come from the committed ncu launch list of this command (profiles/launch_shares.json,
    # written by tools/ncu_table.py from the latest capture), never from a constant in this file.
    probe_share = launch_share("insert", "probe") if (impl == "native" and blocked) else 1.0
    longest_insert_launch = probe_share * ms_ins
    dominant, other = (find_roof, ins_roof) if ms_find >= longest_insert_launch else (ins_roof, find_roof)
    ins_roof["longest_launch_share"] = probe_share
    result = {
        "metric": "Gops/s insert & find (int64 pairs, LF 0.5)",
        "value": value,
        "unit": "Gops/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64",
        "data": "synthetic",
        "impl": impl,
        "config": {"workload": "static_map<int64,int64> 100M uniform pairs insert + find, LF 0.5, "
                               "linear_probing<1>, 1 GPU",
                   "n": n, "load_factor": LOAD_FACTOR, "capacity": table.capacity(),
                   "timing": "CUDA events around insert_async and find_async; clear outside; "
                             "inputs (1.6 GB) and table (3.2 GB) larger than L2"},
        "insert_gops": n / (ms_ins * 1e-3) / 1e9,
        "find_gops": n / (ms_find * 1e-3) / 1e9,
        "insert_ms": ms_ins,
        "find_ms": ms_find,
        **spread,
        "roofline": dominant,
        "roofline_other_pass": other,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
    }
    table.close()
    del table, d_pairs, d_keys, h_pairs, h_keys, h_out
    torch.cuda.empty_cache()
    # the other points of BASELINE configs[1] (LF 0.8, double_hashing<8>), same protocol, outside the
    # timed step of the headline configuration
    if not args.no_points:
        result["c2_points"] = run_points(args, lib, dev, keys, pairs, out)
        result["sweep"] = run_sweep(lib, dev, keys, pairs, out)
    return result


def run_sweep(lib, dev, keys, pairs, out, reps=5):
    """The reference benchmarks' sweep axes (benchmarks/benchmark_defaults.hpp:28-52) on the headline
    instantiation (static_map<int64,int64>, linear_probing<1>), one axis at a time around the defaults:
    occupancy 0.1 - 0.9 (insert_bench.cu, find_bench.cu), matching rate of the find queries 0 / 0.5
    (find_bench.cu:57-59: dropout), key multiplicity 8 and Gaussian skew 0.5 (key_generator.cuh).
    Median of `reps` after 2 warm-ups, Gops/s; every find output is checked."""
    import cucollections_b200 as cb
    from cucollections_b200 import key_generator as kg
    stream = torch.cuda.current_stream(dev)
    n = keys.numel()

    def measure(label, lf, k, p, q, expect):
        t = cb.static_map(n=n, load_factor=lf, probing="linear_probing", cg_size=1, device=dev, _library=lib)
        ins, fnd = [], []
        for i in range(2 + reps):
            t.clear_async()
            ei = timed(lambda: t.insert_async(p), stream)
            ef = timed(lambda: t.find(q, out), stream)
            torch.cuda.synchronize(dev)
            if i >= 2:
                ins.append(ei[0].elapsed_time(ei[1]))
                fnd.append(ef[0].elapsed_time(ef[1]))
        ok = bool((out == expect).all().item())
        t.close()
        del t
        torch.cuda.empty_cache()
        return {**label, "load_factor": lf, "insert_gops": n / (statistics.median(ins) * 1e-3) / 1e9,
                "find_gops": n / (statistics.median(fnd) * 1e-3) / 1e9, "find_output_checked": ok}

    rows = []
    for lf in (0.1, 0.3, 0.7, 0.9):
        rows.append(measure({"axis": "occupancy"}, lf, keys, pairs, keys, keys))
    for rate in (0.0, 0.5):
        q = kg.dropout(keys, rate, seed=43) if rate > 0 else keys + (n + 1)  # uniform keys live in [1, n]
        q = torch.where(q == n, q + 1, q)  # n itself may or may not have been drawn: keep the expectation exact
        present = q < n
        expect = torch.where(present, q, torch.full_like(q, -1))
        rows.append(measure({"axis": "matching_rate", "matching_rate": rate}, 0.5, keys, pairs, q, expect))
        del q, present, expect
    k8 = kg.uniform(n, 8, torch.int64, dev, seed=42)
    rows.append(measure({"axis": "multiplicity", "multiplicity": 8}, 0.5, k8, torch.stack([k8, k8], dim=1).contiguous(),
                        k8, k8))
    del k8
    kgauss = kg.gaussian(n, 0.5, torch.int64, dev, seed=42)
    rows.append(measure({"axis": "skew", "distribution": "gaussian(0.5)"}, 0.5, kgauss,
                        torch.stack([kgauss, kgauss], dim=1).contiguous(), kgauss, kgauss))
    del kgauss
    torch.cuda.empty_cache()
    return rows


def launch_share(op: str, launch: str) -> float:
    """Share of one launch in a multi-launch pass, from the committed ncu launch list."""
    f = ROOT / "profiles" / "launch_shares.json"
    try:
        return float(json.loads(f.read_text())[op][launch])
    except Exception:
        return 1.0  # unknown: treat the pass as one launch (never favours the reported fraction)


def run_points(args, lib, dev, keys, pairs, out, reps=5):
    """Insert / find rates (median and best of `reps`, 3 warm-ups) of the C2 points next to the headline
    one: linear_probing<1> at LF 0.8, double_hashing<8> at LF 0.5 and 0.8 (reference sweep axes:
    benchmarks/static_map/insert_bench.cu, find_bench.cu)."""
    import cucollections_b200 as cb
    stream = torch.cuda.current_stream(dev)
    n = keys.numel()
    rows = []
    for probing, cg, lf in (("linear_probing", 1, 0.8), ("double_hashing", 8, 0.5), ("double_hashing", 8, 0.8)):
        t = cb.static_map(n=n, load_factor=lf, probing=probing, cg_size=cg, device=dev, _library=lib)
        ins, fnd = [], []
        for i in range(3 + reps):
            t.clear_async()
            ei = timed(lambda: t.insert_async(pairs), stream)
            ef = timed(lambda: t.find(keys, out), stream)
            torch.cuda.synchronize(dev)
            if i >= 3:
                ins.append(ei[0].elapsed_time(ei[1]))
                fnd.append(ef[0].elapsed_time(ef[1]))
        ok = bool((out == keys).all().item())
        rows.append({"probing": f"{probing}<{cg}>", "load_factor": lf, "capacity": t.capacity(),
                     "insert_gops_median": n / (statistics.median(ins) * 1e-3) / 1e9,
                     "insert_gops_best": n / (min(ins) * 1e-3) / 1e9,
                     "find_gops_median": n / (statistics.median(fnd) * 1e-3) / 1e9,
                     "find_gops_best": n / (min(fnd) * 1e-3) / 1e9,
                     "insert_frac_of_hbm_roofline": INSERT_BYTES_PER_OP * n / (statistics.median(ins) * 1e-3) / 1e9
                     / measured_hbm_peak()[0],
                     "find_frac_of_hbm_roofline": FIND_BYTES_PER_OP * n / (statistics.median(fnd) * 1e-3) / 1e9
                     / measured_hbm_peak()[0],
                     "all_found": ok})
        t.close()
        del t
        torch.cuda.empty_cache()
    return rows


def multi_gpu_parity_gate(world, rank, dev, lib, routing, n=1_000_000):
    """Checker leg of the multi-GPU bench (never timed): a partitioned table of n pairs per rank is
    filled and queried through the SAME routing path as the timed run and compared, bit-exact, with the
    CPU oracle (oracle/cuco_oracle.c) holding the union of every rank's batch: per-key find / contains
    on a mixed hit/miss batch of every rank, and the global size(). Raises on any mismatch."""
    import numpy as np
    import torch.distributed as dist

    from cucollections_b200 import _cabi
    from cucollections_b200.partitioned import GpuBackend, partitioned_static_map
    from oracle import oracle

    gen = torch.Generator(device=dev).manual_seed(977 + rank)
    keys = torch.randint(1, 3 * n, (n,), generator=gen, device=dev, dtype=torch.int64)  # overlaps across ranks
    pairs = torch.stack([keys, keys * 5 + 2], dim=1).contiguous()
    queries = torch.cat([keys[: n // 2], torch.randint(3 * n, 6 * n, (n - n // 2,), generator=gen, device=dev)])
    table = partitioned_static_map(n * world, 0.5, backend=GpuBackend(dev, lib),
                                   fused_batch=n if routing != "nccl" else None, routing=routing,
                                   probing="linear_probing", cg_size=1)
    table.insert_async(pairs)
    found, present, size = table.find(queries), table.contains(queries), table.size()
    table.insert_async(pairs)  # idempotent second pass through the same buffers
    size_again = table.size()
    gathered = [torch.empty_like(pairs) for _ in range(world)]
    dist.all_gather(gathered, pairs)
    outs = [torch.empty_like(found) for _ in range(world)] if rank == 0 else None
    dist.gather(found, outs, dst=0)
    pres = [torch.empty_like(present) for _ in range(world)] if rank == 0 else None
    dist.gather(present, pres, dst=0)
    qs = [torch.empty_like(queries) for _ in range(world)] if rank == 0 else None
    dist.gather(queries, qs, dst=0)
    verdict = torch.zeros(1, dtype=torch.int64, device=dev)
    detail = ""
    if rank == 0:
        ref = oracle.Table.for_kind(_cabi.MAP_I64_LP1, 2 * n * world, 0.0)
        for g in gathered:
            a = g.cpu().numpy()
            ref.insert(a[:, 0], a[:, 1])
        bad = []
        if ref.size() != size or size_again != size:
            bad.append(f"size {size} / {size_again} != oracle {ref.size()}")
        for r in range(world):
            q = qs[r].cpu().numpy()
            if not np.array_equal(outs[r].cpu().numpy(), ref.find(q)):
                bad.append(f"find of rank {r}")
            if not np.array_equal(pres[r].cpu().numpy().astype(bool), ref.contains(q)):
                bad.append(f"contains of rank {r}")
        verdict[0] = len(bad)
        detail = "; ".join(bad)
    dist.broadcast(verdict, 0)
    table.close()
    if int(verdict.item()) != 0:
        raise AssertionError(f"multi-GPU parity gate failed against the oracle: {detail}")
    return {"checked_against": "oracle/cuco_oracle.c (CPU restatement) over the union of all ranks' batches",
            "pairs_per_rank": n, "queries_per_rank": int(queries.numel()), "ops": ["find", "contains", "size"],
            "size": size, "passed": True}


def run_cpu_reference(args):
    """Fallback reference arm when cuco's GPU build is not available: the host baseline."""
    base = cpu_baseline()
    n = 10_000_000
    ms = 2 * n / (base["value"] * 1e9) * 1e3
    return {"metric": "Gops/s insert & find (int64 pairs, LF 0.5)", "value": base["value"],
            "unit": "Gops/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "static_map<int64,int64> 100M uniform pairs insert + find, LF 0.5, "
                                   "linear_probing<1>, 1 GPU"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "Gops/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--n", type=int, default=N_KEYS, help="pairs on one GPU (N = 1)")
    ap.add_argument("--no-points", action="store_true",
                    help="skip the LF 0.8 / double_hashing<8> points (c2_points)")
    ap.add_argument("--total", type=int, default=0,
                    help="N > 1: pairs over all GPUs (default: BASELINE configs[3], 4 B)")
    ap.add_argument("--batch", type=int, default=0, help="N > 1: pairs per rank and bulk call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true",
                    help="N > 1: skip the BASELINE configs[4] leg (insert_or_apply sum, 2 B rows / 10 M keys)")
    ap.add_argument("--c5-rows", type=int, default=0, help="N > 1: rows of the C5 leg (default 2 B)")
    ap.add_argument("--c5-distinct", type=int, default=0, help="N > 1: distinct keys of the C5 leg (default 10 M)")
    ap.add_argument("--no-c4", action="store_true",
                    help="N > 1: skip the BASELINE configs[3] leg at its stated size (4 B pairs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    from cucollections_b200 import _cabi
    if args.impl == "reference":
        try:
            lib = _cabi.reference()
        except (FileNotFoundError, OSError):
            lib = None
        if lib is None:
            # cuco's GPU build is not in this snapshot: report the host baseline instead (rank 0 only)
            if rank == 0:
                print(json.dumps(run_cpu_reference(args)))
            return
    else:
        lib = _cabi.native()  # raises if the CUDA library is missing: no fallback

    if world > 1 or args.gpus > 1:
        from cucollections_b200 import partitioned
        result = partitioned.bench(args, lib, args.impl, clock_sampler=ClockSampler,
                                   parity_gate=multi_gpu_parity_gate)
    else:
        result = run_single(args, lib, args.impl)

    if rank == 0:
        if not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline()
        print(json.dumps(result))


if __name__ == "__main__":
    main()
