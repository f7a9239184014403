"""Hash-partitioned static_map across the GPUs of one node (one process per GPU, torch.distributed).

There is no reference counterpart (cuCollections is single-GPU); this is the multi-GPU row of the
hot path (BASELINE configs 4-5). Every key has exactly one owner rank,

    owner(key) = mulhi64(fmix64(key ^ salt), world_size)

(the high bits of a 64-bit mix that shares nothing with the in-table xxhash, so each shard still
sees uniformly spread hashes). A bulk operation on a rank's local batch is

    1. partition  - our kernels: per-owner histogram, then a block-aggregated scatter of the keys /
                    pairs (and, for lookups, their source index) into contiguous per-owner segments
    2. exchange   - counts, then payload, with all_to_all_single over NCCL (NVLink 5 / NVSwitch:
                    uniform bandwidth to every peer, so a flat all-to-all is the right shape)
    3. local op   - the single-GPU sm_100a kernels on the received segment
    4. lookups only: results travel back through the mirrored all-to-all and are un-permuted.

Results (per-key found/contains, total size) do not depend on the partitioning, so they are
bit-identical to one big table holding the union of all ranks' keys.

The exchange logic is independent of where the table lives: `backend` supplies partition + local
table operations. `GpuBackend` is the product path (C ABI, no fallback); the gloo/CPU tests plug in a
stand-in from tests/ to exercise the plumbing without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import statistics

import torch
import torch.distributed as dist

from . import _cabi
from .containers import static_map

DEFAULT_SALT = 0x9E3779B97F4A7C15
DEFAULT_ROUTING = "staged"  # "staged" | "fused" | "nccl"; CUCO_B200_ROUTING overrides


def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class GpuBackend:
    """Partition kernels + local table of the native library on this rank's GPU."""

    def __init__(self, device, library=None):
        self.device = torch.device(device)
        self.lib = library or _cabi.native()

    def make_table(self, n_local, load_factor, **kw):
        return static_map(n=n_local, load_factor=load_factor, device=self.device,
                          _library=self.lib, **kw)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def partition(self, keys, num_parts, salt, pair_aos, want_index):
        """keys: [n] int64 keys or [n,2] AoS pairs. Returns (routed, counts[num_parts] cpu list,
        src_index or None); `routed` holds the elements grouped by owner, owners ascending."""
        native = _cabi.native()  # the routing kernels are ours regardless of the table flavour
        n = keys.shape[0]
        counts = torch.zeros(num_parts, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            native.check(native.partition_count(_vp(keys), 8, int(pair_aos), n, num_parts, salt,
                                                _vp(counts), self._stream()))
            cursors = torch.cumsum(counts, 0) - counts
            routed = torch.empty_like(keys)
            index = torch.empty(n, dtype=torch.int64, device=self.device) if want_index else None
            native.check(native.partition_scatter(_vp(keys), None, 8, 8, int(pair_aos), n, num_parts,
                                                  salt, _vp(cursors), _vp(routed), None, _vp(index),
                                                  self._stream()))
        return routed, counts.tolist(), index

    def unpermute(self, values, index, out):
        native = _cabi.native()
        with torch.cuda.device(self.device):
            native.check(native.scatter_by_index(_vp(values), _vp(index), _vp(out),
                                                 values.element_size(), values.numel(), self._stream()))
        return out

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)


class _Lane:
    """One set of exchange buffers (symmetric + local scratch) sized for `n_lane` elements per rank."""


class FusedExchange:
    """Fused routing over peer memory (cuco_b200_exchange_*, include/cuco_b200.h): one kernel groups
    the batch by (owner, L2 region of the owner's shard) and stores it into the owners' buffers over
    NVLink; the owner probes region by region; lookup results are stored back the same way. Buffers
    live in torch symmetric memory so every rank can address every peer's copy. Native build only.

    `lanes` > 1 cuts every batch into that many chunks, each with its own buffer set, and software-
    pipelines them: chunk c+1 is routed (NVLink-bound, on a side stream) while chunk c is probed on the
    caller's stream (SM / L2-bound)."""

    def __init__(self, table, n_max, group, device, salt, lanes=1):
        import torch.distributed._symmetric_memory as symm_mem

        self.table, self.lib, self.device, self.salt = table, table._lib, torch.device(device), salt
        self.group = group
        self.P, self.me = dist.get_world_size(group), dist.get_rank(group)
        self.n_max, self.K = int(n_max), max(1, int(lanes))
        self.n_lane = -(-self.n_max // self.K)
        r, cap, spill = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.lib.check(self.lib.exchange_plan(table._handle, self.n_lane, self.P, C.byref(r), C.byref(cap),
                                              C.byref(spill)))
        self.R, self.cap, self.spill_cap = r.value, cap.value, spill.value
        k = table.kind
        self.key_bytes = k.key.itemsize
        self.slot_bytes = self.key_bytes + (k.value.itemsize if k.value is not None else 0)
        seg_elems = self.P * self.R * self.cap

        def pad(x):
            return (x + 255) // 256 * 256

        off_segments = 0
        off_counts = pad(seg_elems * self.slot_bytes)
        off_flags = off_counts + pad(self.R * self.P * 4)
        off_results = off_flags + pad(self.P * 4)
        lane_bytes = off_results + pad(seg_elems * 8)
        self.buf = symm_mem.empty(lane_bytes * self.K, dtype=torch.uint8, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        vp = C.c_void_p * self.P
        self.lanes = []
        for i in range(self.K):
            lane, shift = _Lane(), i * lane_bytes
            lane.id = i
            lane.peer_segments = vp(*[p + shift + off_segments for p in ptrs])
            lane.peer_counts = vp(*[p + shift + off_counts for p in ptrs])
            lane.peer_flags = vp(*[p + shift + off_flags for p in ptrs])
            lane.peer_results = vp(*[p + shift + off_results for p in ptrs])
            base = self.buf.data_ptr() + shift
            lane.my_segments = C.c_void_p(base + off_segments)
            lane.my_counts = C.c_void_p(base + off_counts)
            lane.my_results = C.c_void_p(base + off_results)
            lane.flags = self.buf[shift + off_flags: shift + off_flags + self.P * 4].view(torch.int32)
            lane.flags.zero_()
            lane.counts_local = torch.zeros(self.P * self.R, dtype=torch.int32, device=self.device)
            lane.position_local = torch.empty(max(1, self.n_lane), dtype=torch.int32, device=self.device)
            lane.spill = torch.empty(self.spill_cap * self.slot_bytes, dtype=torch.uint8, device=self.device)
            lane.spill_index = torch.empty(self.spill_cap, dtype=torch.int32, device=self.device)
            lane.spill_count = torch.zeros(1, dtype=torch.int32, device=self.device)
            lane.landed = torch.cuda.Event()
            lane.consumed = torch.cuda.Event()
            self.lanes.append(lane)
        self.flags_host = torch.zeros(self.K * self.P, dtype=torch.int32).pin_memory()
        self.side = torch.cuda.Stream(self.device) if self.K > 1 else None
        # CUCO_B200_EXCHANGE_TRACE=1: CUDA events around every stage, summarised by trace_summary()
        self.trace = [] if os.environ.get("CUCO_B200_EXCHANGE_TRACE") else None
        torch.cuda.synchronize(self.device)
        dist.barrier(group)

    def launches_per_step(self) -> int:
        # insert = route + publish + probe; find = route + publish + lookup + unpermute (per lane)
        return 7 * self.K

    def _mark(self, label):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))
            self.trace.append((label, e))

    def trace_summary(self):
        """Median milliseconds per stage label (time since the previous mark ON THE SAME STREAM is only
        meaningful for lanes == 1) over all traced calls; mutations and lookups share the routing labels."""
        if not self.trace:
            return {}
        torch.cuda.synchronize(self.device)
        samples = {}
        for (_, a), (label, b) in zip(self.trace, self.trace[1:]):
            if label != "begin":
                samples.setdefault(label, []).append(a.elapsed_time(b))
        return {k: round(statistics.median(v), 4) for k, v in samples.items()}

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _route(self, lane, elems, n, keys_only):
        """Runs on the current stream: wait until every owner has consumed this lane, route, wait until
        every source's segments, counts and flags have landed."""
        self._mark("begin")
        self.hdl.barrier(channel=lane.id)
        self._mark("barrier (previous exchange consumed)")
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_route(
                self.table._handle, _vp(elems), None, n, int(keys_only), self.R, self.cap, self.spill_cap,
                self.P, self.me, self.salt, lane.peer_segments, lane.peer_counts, lane.peer_flags,
                _vp(lane.counts_local), _vp(lane.position_local), _vp(lane.spill), _vp(lane.spill_index),
                _vp(lane.spill_count), self._stream()))
        self._mark("route + publish")
        self.hdl.barrier(channel=lane.id)
        self._mark("barrier (segments landed)")

    def _chunks(self, n):
        return [(lo, min(n, lo + self.n_lane)) for lo in range(0, max(n, 1), self.n_lane)] if n else []

    def _run(self, elems, keys_only, consume):
        """Routes `elems` chunk by chunk (lane = chunk index mod lanes) and calls `consume(lane, lo, hi)`
        on the caller's stream once the chunk has landed everywhere."""
        n = elems.shape[0]
        if n > self.n_max:
            raise ValueError(f"batch of {n} exceeds the exchange buffers sized for {self.n_max}")
        main = torch.cuda.current_stream(self.device)
        # every rank runs all K lanes in every call (empty chunks included): ranks with different batch
        # sizes would otherwise disagree on the number of barriers, and a lane that is skipped would keep
        # the spill flags of an earlier call
        chunks = self._chunks(n)
        chunks += [(n, n)] * (self.K - len(chunks))
        if self.side is None:
            for i, (lo, hi) in enumerate(chunks):
                lane = self.lanes[i % self.K]
                self._route(lane, elems[lo:hi], hi - lo, keys_only)
                consume(lane, lo, hi)
            return
        elems.record_stream(self.side)
        ready = torch.cuda.Event()
        ready.record(main)
        self.side.wait_event(ready)  # the batch is complete before the side stream reads it
        for i, (lo, hi) in enumerate(chunks):
            lane = self.lanes[i % self.K]
            with torch.cuda.stream(self.side):
                self.side.wait_event(lane.consumed)  # our own owner-side work on this lane is done
                self._route(lane, elems[lo:hi], hi - lo, keys_only)
                lane.landed.record(self.side)
            main.wait_event(lane.landed)
            consume(lane, lo, hi)
            lane.consumed.record(main)

    def _spilled(self):
        """(total spilled over all ranks and lanes, [spilled here per lane]): one small read-back per
        bulk call."""
        for i, lane in enumerate(self.lanes):
            self.flags_host[i * self.P:(i + 1) * self.P].copy_(lane.flags, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        mine = [int(self.flags_host[i * self.P + self.me].item()) for i in range(self.K)]
        if max(mine) > self.spill_cap:
            raise RuntimeError("batch too skewed for the exchange buffers: spill list overflowed")
        return int(self.flags_host.sum().item()), mine

    def mutate(self, pairs, reduce_op=-1):
        """Routes and applies a batch of [n, 2] pairs; returns this rank's spilled pairs (or None)."""
        if len(self._chunks(pairs.shape[0])) > self.K:
            raise ValueError("batch needs more chunks than there are lanes")

        def consume(lane, lo, hi):
            with torch.cuda.device(self.device):
                self.lib.check(self.lib.exchange_mutate(self.table._handle, lane.my_segments, lane.my_counts,
                                                        self.R, self.cap, self.P, reduce_op, self._stream()))
            self._mark("probe received segments")

        self._run(pairs, False, consume)
        total, mine = self._spilled()
        if total == 0:
            return None
        parts = [lane.spill[: m * self.slot_bytes].view(pairs.dtype).view(m, 2)
                 for lane, m in zip(self.lanes, mine) if m]
        return torch.cat(parts) if parts else pairs[:0]

    def lookup(self, keys, out, what):
        """what: 0 find, 1 contains. Returns (spilled keys, their source indices) or None."""
        if len(self._chunks(keys.shape[0])) > self.K:
            raise ValueError("batch needs more chunks than there are lanes")
        offsets = {}

        def consume(lane, lo, hi):
            offsets[lane.id] = lo
            with torch.cuda.device(self.device):
                self.lib.check(self.lib.exchange_lookup(self.table._handle, lane.my_segments, lane.my_counts,
                                                        lane.peer_results, self.R, self.cap, self.P, self.me,
                                                        what, self._stream()))
            self._mark("lookup received segments")
            self.hdl.barrier(channel=self.K + lane.id)  # every owner has stored this rank's results
            self._mark("barrier (results landed)")
            with torch.cuda.device(self.device):
                self.lib.check(self.lib.exchange_unpermute(self.table._handle, lane.my_results,
                                                           _vp(lane.position_local), hi - lo, _vp(out[lo:hi]),
                                                           what, self._stream()))
            self._mark("unpermute")

        self._run(keys, True, consume)
        total, mine = self._spilled()
        if total == 0:
            return None
        spilled_keys = [lane.spill[: m * self.key_bytes].view(keys.dtype) for lane, m in zip(self.lanes, mine) if m]
        spilled_at = [lane.spill_index[:m].to(torch.int64) + offsets.get(lane.id, 0)
                      for lane, m in zip(self.lanes, mine) if m]
        if not spilled_keys:
            return keys[:0], torch.empty(0, dtype=torch.int64, device=self.device)
        return torch.cat(spilled_keys), torch.cat(spilled_at)


class SymmTransport:
    """Symmetric-memory buffer of the staged exchange: every rank can address every peer's copy
    (NVLink peer mappings), and a device-side barrier over the group's signal pads."""

    def __init__(self, group, device):
        self.group, self.device = group, torch.device(device)
        self.P, self.me = dist.get_world_size(group), dist.get_rank(group)
        self.hdl = None

    def allocate(self, nbytes):
        import torch.distributed._symmetric_memory as symm_mem
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group if self.group is not None else dist.group.WORLD)
        return self.buf

    def peer_ptrs(self):
        return [int(p) for p in self.hdl.buffer_ptrs]

    def barrier(self, channel):
        self.hdl.barrier(channel=channel)

    def ready(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)


class StagedExchange:
    """Exchange of the hash-partitioned table with the transfer on the copy engines (C ABI
    cuco_b200_exchange_stage / _publish / _copy_async / _apply / _lookup_local / _unpermute).

    Mutations: ONE kernel groups the local batch by (owner, table slice) into a local staging buffer;
    the copy engines then deliver slice after slice to the owners over NVLink while the owners apply
    the slices that have already landed - each slice is a dense batch confined to 1/G of the shard, so
    the owner's L2-blocked path (regroup by 16 MB region, probe region by region) runs on it at full
    density. Transfer and probe overlap without competing for SMs, and nothing depends on a fixed
    number of (owner, region) buckets, so shards of any size (C4: 16 - 64 GB) take the same path.

    Lookups: the batch is cut into `lanes` chunks; per chunk the keys are grouped by owner (recording
    where each answer will come back), copied to the owners, answered there into a local buffer in
    arrival order, copied back and un-permuted; the chunks are software-pipelined so the copy engines
    move chunk c + 1 in and chunk c - 1 back while the SMs answer chunk c.

    `transport` supplies the peer-addressable buffer and the cross-rank barrier (SymmTransport; the
    one-GPU tests plug in simulated ranks)."""

    def __init__(self, table, n_max, group, device, salt, slices=None, lanes=None, transport=None, mode=None):
        self.table, self.lib, self.device, self.salt = table, table._lib, torch.device(device), salt
        self.t = transport or SymmTransport(group, device)
        self.P, self.me = self.t.P, self.t.me
        self.n_max = int(n_max)
        k = table.kind
        self.key_bytes = k.key.itemsize
        self.slot_bytes = self.key_bytes + (k.value.itemsize if k.value is not None else 0)
        self.result_bytes = 8
        shard_bytes = table.capacity() * self.slot_bytes
        explicit_slices = slices or int(os.environ.get("CUCO_B200_EXCHANGE_SLICES", "0"))
        if slices is None:
            slices = explicit_slices or max(1, min(16, shard_bytes // (384 << 20)))
        if lanes is None:
            lanes = int(os.environ.get("CUCO_B200_EXCHANGE_LANES", "0")) or max(1, min(4, self.n_max // (1 << 22)))
        self.G, self.L = int(slices), int(lanes)
        self.n_lane = -(-self.n_max // self.L)
        # Mutations, two modes. fine: the SOURCE groups by (owner, L2 region of the owner's shard), the
        # owner probes the received segments region by region and never regroups - possible while
        # ranks x regions fits the router's bucket budget (shards up to a few GB). coarse: the source groups
        # by (owner, one of G table slices) and the owner runs its L2-blocked path (regroup + probe) on
        # each slice as it lands - any shard size. B = buckets per owner of the staging buffer.
        mode = mode or os.environ.get("CUCO_B200_EXCHANGE_MODE", "auto")
        fine = C.c_uint32(0)
        if mode != "coarse":
            self.lib.check(self.lib.exchange_fine_regions(table._handle, self.P, C.byref(fine)))
            if mode == "fine" and fine.value == 0:
                raise ValueError("shard too large for the fine mode of the staged exchange")
        self.fine = fine.value != 0
        self.B = fine.value if self.fine else self.G
        if self.fine:
            self.G = max(1, min(explicit_slices or 8, self.B))  # transfer phases (groups of consecutive regions)
        cap, spill = C.c_uint32(), C.c_uint32()
        self.lib.check(self.lib.exchange_stage_plan(table._handle, self.n_max, self.P, self.B, C.byref(cap), C.byref(spill)))
        self.cap, self.spill_cap = cap.value, spill.value
        self.lib.check(self.lib.exchange_stage_plan(table._handle, self.n_lane, self.P, 1, C.byref(cap), C.byref(spill)))
        self.cap_l, self.spill_cap_l = cap.value, spill.value
        P, G, L = self.P, self.B, self.L  # G: buckets per owner in the buffer layouts below

        def pad(x):
            return (x + 255) // 256 * 256

        # ---- peer-addressable part (identical layout on every rank) ----
        self.off_recv = 0             # coarse: [slice][source][cap] slot images; fine: [source][region][cap]
        self.off_counts = self.off_recv + pad(G * P * self.cap * self.slot_bytes)      # [G][P] uint32
        self.off_flags = self.off_counts + pad(G * P * 4)                   # [P] uint32: spill counts of mutations
        self.off_keys = self.off_flags + pad(P * 4)                         # L x [P][cap_l] keys
        self.keys_lane = pad(P * self.cap_l * self.key_bytes)
        self.off_kcounts = self.off_keys + L * self.keys_lane               # L x [P] uint32
        self.off_kflags = self.off_kcounts + L * pad(P * 4)                 # L x [P] uint32
        self.off_back = self.off_kflags + L * pad(P * 4)                    # L x [P][cap_l] results (8 bytes each)
        self.back_lane = pad(P * self.cap_l * self.result_bytes)
        total = self.off_back + L * self.back_lane
        self.buf = self.t.allocate(total)
        self.base = self.buf.data_ptr()
        self.buf[self.off_counts:self.off_keys].zero_()
        self.buf[self.off_kcounts:self.off_back].zero_()
        # ---- local part ----
        dev = self.device
        self.stage = torch.empty(P * G * self.cap * self.slot_bytes, dtype=torch.uint8, device=dev)
        self.counts_local = torch.zeros(P * G, dtype=torch.int32, device=dev)
        self.spill = torch.empty(self.spill_cap * self.slot_bytes, dtype=torch.uint8, device=dev)
        self.spill_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.position_local = torch.empty(max(1, self.n_max), dtype=torch.int32, device=dev)
        self.lanes = []
        for i in range(L):
            lane = _Lane()
            lane.id = i
            lane.stage = torch.empty(P * self.cap_l * self.key_bytes, dtype=torch.uint8, device=dev)
            lane.counts_local = torch.zeros(P, dtype=torch.int32, device=dev)
            lane.results = torch.empty(P * self.cap_l * self.result_bytes, dtype=torch.uint8, device=dev)
            lane.spill = torch.empty(self.spill_cap_l * self.key_bytes, dtype=torch.uint8, device=dev)
            lane.spill_index = torch.empty(self.spill_cap_l, dtype=torch.int32, device=dev)
            lane.spill_count = torch.zeros(1, dtype=torch.int32, device=dev)
            lane.staged, lane.landed, lane.answered, lane.returned = (torch.cuda.Event() for _ in range(4))
            self.lanes.append(lane)
        self.flags_host = torch.zeros((1 + L) * P, dtype=torch.int32).pin_memory()
        # high priority: the tiny barrier / publish kernels on these streams must not queue behind the
        # probe kernels' CTAs, or a landed slice would be announced late
        self.copy_in = torch.cuda.Stream(dev, priority=-1)    # source -> owner transfers
        self.copy_back = torch.cuda.Stream(dev, priority=-1)  # owner -> source transfers (lookup results)
        # measured at 8 GPUs (profiles/r02_exchange_sweep_8gpu.txt): 1 / 4 / 8 side streams -> 178 / 176 / 158
        # Gops/s: the copies of one phase do not gain from running side by side, 8 at once lose to ingress
        # contention. One stream is the default; the fan-out stays for other topologies.
        fan = int(os.environ.get("CUCO_B200_COPY_STREAMS", "1"))
        # blocks up to this size travel by push kernel (0 = always the copy engines)
        self.push_bytes = int(os.environ.get("CUCO_B200_PUSH_MIB", "8")) << 20
        self.push_ctas = int(os.environ.get("CUCO_B200_PUSH_CTAS", "8"))
        self.fan = [torch.cuda.Stream(dev, priority=-1) for _ in range(max(0, fan))]       # source -> owner
        self.fan_back = [torch.cuda.Stream(dev, priority=-1) for _ in range(max(0, fan))]  # owner -> source
        self.fork_events = [torch.cuda.Event() for _ in range(64)]
        self.join_events = [torch.cuda.Event() for _ in range(256)]
        self.fork_next = self.join_next = 0
        # streams the slices are applied on, round-robin. Measured: 2 streams gain 4 % in fine mode (probe-only
        # applies: the next probe fills the tail of the previous one, 2 GPUs 3.60 -> 3.45 ms) but cost 20 % on the
        # 16.5 GB shards of C4 in coarse mode (two regroup + probe pairs at once evict each other's L2 regions:
        # 21.0 -> 26.1 ms per 500 M pairs at 8 GPUs), so: 2 in fine mode, 1 in coarse mode
        default_apply = "2" if self.fine else "1"
        self.apply_streams = [torch.cuda.Stream(dev)
                              for _ in range(int(os.environ.get("CUCO_B200_APPLY_STREAMS", default_apply)) - 1)]
        self.apply_done = [torch.cuda.Event() for _ in self.apply_streams]
        self.staged = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.slice_landed = [torch.cuda.Event() for _ in range(self.G)]
        self.consumed.record(torch.cuda.current_stream(dev))
        self._peers = None
        self.trace = [] if os.environ.get("CUCO_B200_EXCHANGE_TRACE") else None
        self.t.ready()

    # ---- plumbing --------------------------------------------------------------------------------
    def launches_per_step(self) -> int:
        # insert = stage + publish + G x (probe [+ regroup in coarse mode]); find = lanes x (stage + publish +
        # lookup + unpermute)
        return 2 + (1 if self.fine else 2) * self.G + 4 * self.L

    def _tick(self, label, stream=None):
        """CUCO_B200_EXCHANGE_TRACE=1: a timing event on `stream`; trace_summary() reports every label as
        milliseconds since the call's "begin" (median over the traced calls), i.e. the pipeline as it ran."""
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream or torch.cuda.current_stream(self.device))
            if label == "begin":
                self.trace.append([])
            self.trace[-1].append((label, e))

    def trace_summary(self):
        if not self.trace:
            return {}
        torch.cuda.synchronize(self.device)
        samples = {}
        for call in self.trace[1:] or self.trace:  # the first call is a warm-up when there are more
            for label, e in call[1:]:
                samples.setdefault(label, []).append(call[0][1].elapsed_time(e))
        return {k: round(statistics.median(v), 3) for k, v in samples.items()}

    def peers(self):
        if self._peers is None:
            self._peers = self.t.peer_ptrs()
        return self._peers

    def _ptr_array(self, offset):
        return (C.c_void_p * self.P)(*[p + offset for p in self.peers()])

    def _s(self, stream=None):
        return C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)

    def _copy(self, dst, src, nbytes, stream):
        self.lib.check(self.lib.copy_async(C.c_void_p(dst), C.c_void_p(src), nbytes, self._s(stream)))

    def _fan_out(self, copies, stream, fan=None):
        """Issues the (dst, src, bytes) copies of one phase. On `stream` alone they would run one after
        the other, each paying the fixed cost of a peer copy (measured at 8 GPUs: 8 x 3 MB took 0.5 ms);
        spread over `fan` side streams they overlap and keep several copy engines busy. Fork / join with
        events, so that for `stream` the phase still looks like one operation."""
        if copies and max(c[2] for c in copies) <= self.push_bytes and len(copies) <= 16:
            # small blocks (lookup chunks): one push kernel instead of a peer copy per destination
            n = len(copies)
            dst = (C.c_void_p * n)(*[c[0] for c in copies])
            src = (C.c_void_p * n)(*[c[1] for c in copies])
            size = (C.c_int64 * n)(*[c[2] for c in copies])
            self.lib.check(self.lib.push_async(dst, src, size, n, self.push_ctas, self._s(stream)))
            return
        fan = self.fan if fan is None else fan
        if stream is None or len(fan) <= 1 or len(copies) <= 1:
            for dst, src, nbytes in copies:
                self._copy(dst, src, nbytes, stream)
            return
        fork = self.fork_events[self.fork_next % len(self.fork_events)]
        self.fork_next += 1
        fork.record(stream)
        used = min(len(fan), len(copies))
        for j in range(used):
            fan[j].wait_event(fork)
        for i, (dst, src, nbytes) in enumerate(copies):
            self._copy(dst, src, nbytes, fan[i % used])
        for j in range(used):
            join = self.join_events[(self.join_next + j) % len(self.join_events)]
            join.record(fan[j])
            stream.wait_event(join)
        self.join_next += used

    # ---- mutations: the steps (driven in lock-step by the simulated-rank tests) --------------------
    def stage_pairs(self, pairs):
        n = pairs.shape[0]
        if n > self.n_max:
            raise ValueError(f"batch of {n} exceeds the exchange buffers sized for {self.n_max}")
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_stage(
                self.table._handle, _vp(pairs), None, n, 0, self.B, self.cap, self.spill_cap, self.P, self.me,
                self.salt, _vp(self.stage), _vp(self.counts_local), None, _vp(self.spill), None,
                _vp(self.spill_count), self._s()))

    def publish_pairs(self, stream=None):
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_publish(
                _vp(self.counts_local), _vp(self.spill_count), self._ptr_array(self.off_counts),
                self._ptr_array(self.off_flags), self.B, self.cap, self.P, self.me, int(self.fine),
                self._s(stream)))

    def _buckets(self, g):
        """Buckets (table slices, or L2 regions in fine mode) that travel in phase g."""
        return (g * self.B // self.G, (g + 1) * self.B // self.G)

    def send_slice(self, g, stream=None):
        """Phase g: this rank's blocks for every owner -> the owner's receive buffer, owners in
        rank-rotated order (at any moment every rank writes to a different peer)."""
        block = self.cap * self.slot_bytes
        b0, b1 = self._buckets(g)
        copies = []
        for i in range(self.P):
            o = (self.me + i) % self.P
            src = self.stage.data_ptr() + (o * self.B + b0) * block
            if self.fine:   # recv[source][region][cap]
                dst = self.peers()[o] + self.off_recv + (self.me * self.B + b0) * block
            else:           # recv[slice][source][cap], one slice per phase
                dst = self.peers()[o] + self.off_recv + (b0 * self.P + self.me) * block
            copies.append((dst, src, (b1 - b0) * block))
        with torch.cuda.device(self.device):
            self._fan_out(copies, stream)

    def apply_slice(self, g, reduce_op=-1):
        block = self.cap * self.slot_bytes
        b0, b1 = self._buckets(g)
        with torch.cuda.device(self.device):
            if self.fine:
                self.lib.check(self.lib.exchange_probe(
                    self.table._handle, C.c_void_p(self.base + self.off_recv), C.c_void_p(self.base + self.off_counts),
                    self.B, self.cap, self.P, b0, b1 - b0, reduce_op, self._s()))
            else:
                self.lib.check(self.lib.exchange_apply(
                    self.table._handle, C.c_void_p(self.base + self.off_recv + g * self.P * block),
                    C.c_void_p(self.base + self.off_counts + g * self.P * 4), self.cap, self.P, g, self.G,
                    reduce_op, self._s()))

    def mutate(self, pairs, reduce_op=-1):
        """Routes and applies a batch of [n, 2] pairs; returns this rank's spilled pairs (or None)."""
        main = torch.cuda.current_stream(self.device)
        self._tick("begin")
        self.stage_pairs(pairs)
        self.staged.record(main)
        self._tick("mutate: staged")
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.staged)
            self.copy_in.wait_event(self.consumed)   # this rank's previous exchange has been applied
            self.t.barrier(0)                        # ... and so has everybody else's: the buffers are free
            self.publish_pairs(self.copy_in)
            for g in range(self.G):
                self.send_slice(g, self.copy_in)
                self.t.barrier(0)                    # slice g has landed on every owner
                self.slice_landed[g].record(self.copy_in)
                self._tick(f"mutate: slice {g} landed", self.copy_in)
        if not self.apply_streams:
            for g in range(self.G):
                main.wait_event(self.slice_landed[g])
                self.apply_slice(g, reduce_op)
                self._tick(f"mutate: slice {g} applied")
        else:
            # slices alternate between the caller's stream and side streams: the regrouping pass of slice
            # g + 1 fills the tail of the probe of slice g (slices touch disjoint parts of the table)
            lanes = [main] + self.apply_streams
            self.staged.record(main)
            for st in self.apply_streams:
                st.wait_event(self.staged)  # everything the caller queued before this call
            for g in range(self.G):
                st = lanes[g % len(lanes)]
                with torch.cuda.stream(st):
                    st.wait_event(self.slice_landed[g])
                    self.apply_slice(g, reduce_op)
                    self._tick(f"mutate: slice {g} applied", st)
            for i, st in enumerate(self.apply_streams):
                self.apply_done[i].record(st)
                main.wait_event(self.apply_done[i])
        self.consumed.record(main)
        total, mine = self._spilled(mutation=True)
        if total == 0:
            return None
        m = mine[0]
        return self.spill[: m * self.slot_bytes].view(pairs.dtype).view(m, 2)

    # ---- lookups ---------------------------------------------------------------------------------
    def _chunks(self, n):
        return [(lo, min(n, lo + self.n_lane)) for lo in range(0, n, self.n_lane)] or [(0, 0)]

    def stage_keys(self, lane, keys, lo):
        n = keys.shape[0]
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_stage(
                self.table._handle, _vp(keys), None, n, 1, 1, self.cap_l, self.spill_cap_l, self.P, self.me,
                self.salt, _vp(lane.stage), _vp(lane.counts_local), _vp(self.position_local[lo:lo + max(n, 1)]),
                _vp(lane.spill), _vp(lane.spill_index), _vp(lane.spill_count), self._s()))

    def send_keys(self, lane, stream=None):
        pad4 = (self.P * 4 + 255) // 256 * 256
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_publish(
                _vp(lane.counts_local), _vp(lane.spill_count), self._ptr_array(self.off_kcounts + lane.id * pad4),
                self._ptr_array(self.off_kflags + lane.id * pad4), 1, self.cap_l, self.P, self.me, 0,
                self._s(stream)))
            block = self.cap_l * self.key_bytes
            self._fan_out([(self.peers()[o] + self.off_keys + lane.id * self.keys_lane + self.me * block,
                            lane.stage.data_ptr() + o * block, block)
                           for o in ((self.me + i) % self.P for i in range(self.P))], stream)

    def answer_keys(self, lane, what):
        pad4 = (self.P * 4 + 255) // 256 * 256
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_lookup_local(
                self.table._handle, C.c_void_p(self.base + self.off_keys + lane.id * self.keys_lane),
                C.c_void_p(self.base + self.off_kcounts + lane.id * pad4), _vp(lane.results), self.cap_l, self.P,
                what, self._s()))

    def return_results(self, lane, what, stream=None):
        rbytes = self._result_bytes(what)
        block = self.cap_l * rbytes
        with torch.cuda.device(self.device):
            self._fan_out([(self.peers()[s] + self.off_back + lane.id * self.back_lane + self.me * block,
                            lane.results.data_ptr() + s * block, block)
                           for s in ((self.me + i) % self.P for i in range(self.P))], stream, self.fan_back)

    def _result_bytes(self, what):
        k = self.table.kind
        return 1 if what == 1 else (k.value.itemsize if k.value is not None else k.key.itemsize)

    def unpermute(self, lane, what, out, lo, hi):
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.exchange_unpermute(
                self.table._handle, C.c_void_p(self.base + self.off_back + lane.id * self.back_lane),
                _vp(self.position_local[lo:lo + max(hi - lo, 1)]), hi - lo, _vp(out[lo:hi]), what, self._s()))

    def lookup(self, keys, out, what):
        """what: 0 find, 1 contains. Returns (spilled keys, their source indices) or None."""
        n = keys.shape[0]
        if n > self.n_max:
            raise ValueError(f"batch of {n} exceeds the exchange buffers sized for {self.n_max}")
        main = torch.cuda.current_stream(self.device)
        # every rank runs all L lanes in every call (empty chunks included): the barriers must match
        chunks = self._chunks(n)
        chunks += [(n, n)] * (self.L - len(chunks))

        def stage(c):
            lo, hi = chunks[c]
            self.stage_keys(self.lanes[c], keys[lo:hi], lo)
            self.lanes[c].staged.record(main)
            self._tick(f"lookup: chunk {c} staged")

        def transfer_in(c):
            lane = self.lanes[c]
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(lane.staged)
                if c == 0:
                    self.copy_in.wait_event(self.consumed)
                    self.t.barrier(1)                 # every rank is done with the previous lookup's buffers
                self.send_keys(lane, self.copy_in)
                self.t.barrier(1)                     # chunk c has landed on every owner
                lane.landed.record(self.copy_in)
                self._tick(f"lookup: chunk {c} landed", self.copy_in)

        def transfer_back(c):
            lane = self.lanes[c]
            with torch.cuda.stream(self.copy_back):
                self.copy_back.wait_event(lane.answered)
                self.return_results(lane, what, self.copy_back)
                self.t.barrier(2)                     # every owner has returned this rank's answers of chunk c
                lane.returned.record(self.copy_back)
                self._tick(f"lookup: chunk {c} returned", self.copy_back)

        self._tick("begin")
        stage(0)
        transfer_in(0)
        for c in range(self.L):
            if c + 1 < self.L:
                stage(c + 1)
                transfer_in(c + 1)
            lane = self.lanes[c]
            main.wait_event(lane.landed)
            self.answer_keys(lane, what)
            lane.answered.record(main)
            self._tick(f"lookup: chunk {c} answered")
            transfer_back(c)
            if c >= 1:
                main.wait_event(self.lanes[c - 1].returned)
                self.unpermute(self.lanes[c - 1], what, out, *chunks[c - 1])
                self._tick(f"lookup: chunk {c - 1} unpermuted")
        last = self.L - 1
        main.wait_event(self.lanes[last].returned)
        self.unpermute(self.lanes[last], what, out, *chunks[last])
        self._tick(f"lookup: chunk {last} unpermuted")
        self.consumed.record(main)
        total, mine = self._spilled(mutation=False)
        if total == 0:
            return None
        parts = [(c, m) for c, m in enumerate(mine) if m]
        if not parts:
            return keys[:0], torch.empty(0, dtype=torch.int64, device=self.device)
        spilled_keys = [self.lanes[c].spill[: m * self.key_bytes].view(keys.dtype) for c, m in parts]
        spilled_at = [self.lanes[c].spill_index[:m].to(torch.int64) + chunks[c][0] for c, m in parts]
        return torch.cat(spilled_keys), torch.cat(spilled_at)

    def _spilled(self, mutation):
        """(total spilled over all ranks, [spilled on this rank per buffer set]); one small read-back."""
        P = self.P
        pad4 = (P * 4 + 255) // 256 * 256
        if mutation:
            views = [self.buf[self.off_flags: self.off_flags + P * 4].view(torch.int32)]
            caps = [self.spill_cap]
        else:
            views = [self.buf[self.off_kflags + i * pad4: self.off_kflags + i * pad4 + P * 4].view(torch.int32)
                     for i in range(self.L)]
            caps = [self.spill_cap_l] * self.L
        host = self.flags_host[: len(views) * P]
        for i, v in enumerate(views):
            host[i * P:(i + 1) * P].copy_(v, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        mine = [int(host[i * P + self.me].item()) for i in range(len(views))]
        if any(m > c for m, c in zip(mine, caps)):
            raise RuntimeError("batch too skewed for the exchange buffers: spill list overflowed")
        return int(host.sum().item()), mine


class partitioned_static_map:
    """static_map<int64,int64> sharded over the ranks of `group` by owner(key)."""

    def __init__(self, n_total, load_factor=0.5, *, backend, group=None, salt=DEFAULT_SALT,
                 headroom=1.03, fused_batch=None, fused_lanes=None, routing=None, **table_kw):
        """`fused_batch`: largest batch (elements per rank and call) the exchange buffers are sized for;
        None keeps the all_to_all routing (also the fallback for spilled elements). It is a contract of
        the collective: every rank may pass batches of any size up to it (ragged and empty batches are
        fine - all ranks always run the same number of exchange rounds), but a rank that exceeds it
        raises ValueError before it takes part in the call and leaves its peers waiting, like a size
        mismatch in any collective. `routing`: "staged" (default; copy-engine transfers overlapped with the
        owners' probes), "fused" (round 1: peer stores from the routing kernel) or "nccl".
        `fused_lanes`: chunks per batch that are software-pipelined (routing of chunk c+1 overlaps the
        owner-side probe of chunk c); default 1 or CUCO_B200_EXCHANGE_LANES."""
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.backend = backend
        self.salt = salt
        # the owner hash spreads keys evenly up to binomial noise; a few percent of head-room keeps
        # every shard at or below the requested load factor
        n_local = int(-(-n_total // self.world) * headroom) + 1
        self.table = backend.make_table(n_local, load_factor, **table_kw)
        self._table_kw, self._scratch, self._scratch_capacity = dict(table_kw), None, 0
        self.fused = None
        self.routing = "nccl"
        if fused_batch:
            self.routing = routing or os.environ.get("CUCO_B200_ROUTING", DEFAULT_ROUTING)
            if self.routing == "nccl":
                pass
            elif self.routing == "staged":
                self.fused = StagedExchange(self.table, fused_batch, group, backend.device, salt)
            else:
                lanes = fused_lanes or int(os.environ.get("CUCO_B200_EXCHANGE_LANES", "1"))
                self.fused = FusedExchange(self.table, fused_batch, group, backend.device, salt, lanes=lanes)

    def launches_per_step(self) -> int:
        """Kernels of this repository per (insert + find) step and rank, for bench.py's `gpu_launches`."""
        if self.fused is not None:
            return self.fused.launches_per_step()
        # all_to_all routing: 2 x (count + scatter) + insert (route + probe when blocked) + find + scatter_by_index
        return 8

    # ---- exchange helpers ------------------------------------------------------------------------
    def _exchange_counts(self, send_counts):
        send = torch.tensor(send_counts, dtype=torch.int64)
        recv = torch.empty(self.world, dtype=torch.int64)
        if dist.get_backend(self.group) == "nccl":
            dev = self.backend.device
            send, recv = send.to(dev), recv.to(dev)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.tolist()

    def _route(self, elems, pair_aos, want_index):
        routed, send_counts, index = self.backend.partition(elems, self.world, self.salt, pair_aos,
                                                            want_index)
        recv_counts = self._exchange_counts(send_counts)
        shape = (sum(recv_counts), 2) if pair_aos else (sum(recv_counts),)
        received = self.backend.empty(shape, routed.dtype)
        dist.all_to_all_single(received, routed, output_split_sizes=recv_counts,
                               input_split_sizes=send_counts, group=self.group)
        return received, send_counts, recv_counts, index

    def _return(self, results, send_counts, recv_counts, index, out):
        back = self.backend.empty((sum(send_counts),), results.dtype)
        dist.all_to_all_single(back, results, output_split_sizes=send_counts,
                               input_split_sizes=recv_counts, group=self.group)
        return self.backend.unpermute(back, index, out)

    # ---- bulk API --------------------------------------------------------------------------------
    def insert_async(self, pairs):
        """pairs: [n, 2] int64 (key, value) on this rank. Stream-ordered on every rank."""
        if self.fused is not None:
            pairs = self.fused.mutate(pairs)
            if pairs is None:
                return
        received, *_ = self._route(pairs, True, False)
        self.table.insert_async(received)

    def insert(self, pairs) -> int:
        """Returns the number of new keys over all ranks (collective routing: it needs the count)."""
        received, *_ = self._route(pairs, True, False)
        new = torch.tensor([self.table.insert(received)], dtype=torch.int64)
        return self._allreduce_sum(new)

    def insert_or_apply(self, pairs, op="plus", init=None, pre_aggregate=None):
        """payload[key] = fold of `op` over every rank's rows carrying key.

        `pre_aggregate` = capacity hint (distinct keys this rank may see): the rows are first folded into
        a rank-local scratch table (reference idea: insert_or_apply_shmem, static_map/kernels.cuh:171-264,
        one level up - the whole GPU instead of one block), and only its <= `pre_aggregate` partial
        results cross NVLink. op must be associative and commutative (plus / min / max are)."""
        if pre_aggregate:
            if self._scratch is None or self._scratch_capacity < pre_aggregate:
                if self._scratch is not None:
                    self._scratch.close()
                self._scratch = self.backend.make_table(int(pre_aggregate), 0.5, **self._table_kw)
                self._scratch_capacity = int(pre_aggregate)
            self._scratch.clear_async()
            self._scratch.insert_or_apply(pairs, op=op, init=init)
            keys, vals = self._scratch.retrieve_all()
            pairs = torch.stack([keys, vals], dim=1).contiguous()
        if self.fused is not None and init is None:
            pairs = self.fused.mutate(pairs, {"plus": _cabi.PLUS, "min": _cabi.MIN, "max": _cabi.MAX}[op])
            if pairs is None:
                return
        received, *_ = self._route(pairs, True, False)
        self.table.insert_or_apply(received, op=op, init=init)

    def _lookup(self, keys, out, what):
        """what: 0 find, 1 contains (uint8 results)."""
        if self.fused is not None:
            spilled = self.fused.lookup(keys, out, what)
            if spilled is None:
                return out
            keys, where = spilled  # finish the spilled keys through the collective path
            part = self.backend.empty((keys.shape[0],), out.dtype)
            self._lookup_collective(keys, part, what)
            out[where] = part
            return out
        return self._lookup_collective(keys, out, what)

    def _lookup_collective(self, keys, out, what):
        received, send_counts, recv_counts, index = self._route(keys, False, True)
        results = self.table.find(received) if what == 0 else self.table.contains(received).view(torch.uint8)
        return self._return(results, send_counts, recv_counts, index, out)

    def find(self, keys, out=None):
        if out is None:
            out = self.backend.empty((keys.shape[0],), self.table._payload_dtype())
        return self._lookup(keys, out, 0)

    def contains(self, keys, out=None):
        if out is None:
            out = self.backend.empty((keys.shape[0],), torch.uint8)
        return self._lookup(keys, out.view(torch.uint8), 1).view(torch.bool)

    def clear_async(self):
        self.table.clear_async()

    def _allreduce_sum(self, t):
        if dist.get_backend(self.group) == "nccl":
            t = t.to(self.backend.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def size(self) -> int:
        return self._allreduce_sum(torch.tensor([self.table.size()], dtype=torch.int64))

    def local_size(self) -> int:
        return self.table.size()

    def close(self):
        self.table.close()
        if self._scratch is not None:
            self._scratch.close()
            self._scratch = None


# ==================================================================================================
# benchmark entry (bench.py --gpus N under torchrun)
# ==================================================================================================
def _hbm_peak():
    import json
    from pathlib import Path
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"])
    except Exception:
        return 6650.0


def _timed_pass(table, dev, stream, steps, warmup, pairs, keys, out):
    """warmup + steps of (clear, insert, find); returns per-step (insert ms, find ms) lists."""
    def step():
        table.clear_async()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(stream)
        table.insert_async(pairs)
        b.record(stream)
        table.find(keys, out)
        c.record(stream)
        return a, b, c

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    assert bool((out == keys).all().item()), "partitioned find returned a wrong payload"
    dist.barrier()
    torch.cuda.synchronize(dev)
    events = [step() for _ in range(steps)]
    torch.cuda.synchronize(dev)
    dist.barrier()
    return [a.elapsed_time(b) for a, b, _ in events], [b.elapsed_time(c) for _, b, c in events]


def _max_over_ranks(values, dev):
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def _c4_leg(args, lib, dev, rank, world, routing):
    """BASELINE configs[3] at its stated size: `total` pairs (default 4 B) over the `world` GPUs, LF 0.5,
    linear_probing<1, xxhash_64> (a 32-bit hash folds unevenly onto shards of 1 - 4 G slots). Every rank
    feeds its share in batches of `batch` pairs (keys of batch b are regenerated from its seed for the
    find pass instead of being kept: at 2 GPUs a rank's 2 G pairs are 32 GB next to a 64 GB shard).
    Events around the bulk calls only, max over ranks per batch, summed. Properties checked at this size:
    every key found with its own payload, absent keys miss, global size == number of distinct keys."""
    from . import key_generator as kg

    stream = torch.cuda.current_stream(dev)
    total = args.total or 4_000_000_000
    share = total // world
    batch = min(args.batch or 500_000_000, share)
    batches = -(-share // batch)
    table = partitioned_static_map(total, 0.5, backend=GpuBackend(dev, lib),
                                   fused_batch=batch if routing != "nccl" else None, routing=routing,
                                   probing="linear_probing", cg_size=1, hash="xxhash_64")
    out = torch.empty(batch, dtype=torch.int64, device=dev)

    def make(b):
        m = min(batch, share - b * batch)
        # rank r draws uniformly from its own value range [r * share + 1, (r + 1) * share] (~63 % distinct)
        g = torch.Generator(device=dev).manual_seed(1000 + 64 * b + rank)
        k = torch.randint(1 + rank * share, 1 + (rank + 1) * share, (m,), generator=g, device=dev, dtype=torch.int64)
        return k, torch.stack([k, k], dim=1).contiguous()

    ins_ms, find_ms = [], []
    seen = torch.zeros(share, dtype=torch.bool, device=dev)  # exact distinct count without a sort
    for b in range(batches):
        keys, pairs = make(b)
        for _ in range(1 if b else 2):  # one untimed warm-up of the path on the first batch
            if not b:
                table.clear_async()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            table.insert_async(pairs)
            e1.record(stream)
            torch.cuda.synchronize(dev)
        ins_ms.append(e0.elapsed_time(e1))
        del pairs
        seen[keys - (1 + rank * share)] = True
        del keys
    distinct = int(seen.sum().item())
    del seen
    torch.cuda.empty_cache()
    size = table.size()
    ok = True
    for b in range(batches):
        keys, _ = make(b)
        del _
        o = out[: keys.numel()]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        table.find(keys, o)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        find_ms.append(e0.elapsed_time(e1))
        ok = ok and bool((o == keys).all().item())
        if b == 0:
            absent = keys[: 1 << 20] + total + 7
            ok = ok and bool((table.find(absent) == -1).all().item())
        del keys
    ins = _max_over_ranks(ins_ms, dev)
    fnd = _max_over_ranks(find_ms, dev)
    counts = torch.tensor([distinct, 0 if ok else 1], dtype=torch.int64, device=dev)
    dist.all_reduce(counts)
    capacity = table.table.capacity()
    table.close()
    torch.cuda.empty_cache()
    t_ins, t_find = sum(ins) * 1e-3, sum(fnd) * 1e-3
    pairs_done = share * world
    result = {
        "workload": f"hash-partitioned static_map<int64,int64>, {pairs_done} uniform pairs over {world} GPUs "
                    f"({share} per GPU in {batches} bulk calls of {batch}), LF 0.5, linear_probing<1, xxhash_64>",
        "pairs_total": pairs_done, "pairs_per_gpu": share, "batch": batch, "shard_capacity": capacity,
        "shard_bytes": capacity * 16,
        "insert_gops": pairs_done / t_ins / 1e9, "find_gops": pairs_done / t_find / 1e9,
        "value": 2 * pairs_done / (t_ins + t_find) / 1e9, "unit": "Gops/s",
        "insert_ms_per_batch": ins, "find_ms_per_batch": fnd,
        "size": size, "distinct_keys": int(counts[0].item()),
        "properties": {"all_found_with_own_payload_and_absent_miss": int(counts[1].item()) == 0,
                       "size_equals_distinct": size == int(counts[0].item())},
    }
    if not all(result["properties"].values()):
        raise AssertionError(f"C4 leg failed its properties: {result}")
    return result


def _c5_leg(args, lib, dev, rank, world, routing):
    """BASELINE configs[4]: group-by aggregate, insert_or_apply(plus) of 2 B rows with 10 M distinct keys
    (uniform_int[1, 1e7], value 1, empty value 0: benchmarks/static_map/insert_or_apply_bench.cu:59) into the
    partitioned static_map<int64,int64>, measured both ways: every row routed to its owner, and rows folded
    into a rank-local scratch table first so that only <= 10 M partial sums per rank cross NVLink. Both
    results are compared bit-exactly with each other and with the all-reduced histogram of the rows."""
    stream = torch.cuda.current_stream(dev)
    rows_total = args.c5_rows or 2_000_000_000
    distinct = args.c5_distinct or 10_000_000
    share = rows_total // world
    batch = min(share, 250_000_000)
    batches = -(-share // batch)

    def make(b):
        m = min(batch, share - b * batch)
        g = torch.Generator(device=dev).manual_seed(5000 + 64 * b + rank)
        k = torch.randint(1, distinct + 1, (m,), generator=g, device=dev, dtype=torch.int64)
        return torch.stack([k, torch.ones_like(k)], dim=1).contiguous()

    probe = torch.arange(0, distinct + 8, device=dev, dtype=torch.int64)
    hist = torch.zeros(distinct + 8, dtype=torch.int64, device=dev)
    for b in range(batches):
        p = make(b)
        hist += torch.bincount(p[:, 0], minlength=distinct + 8)
        del p
    dist.all_reduce(hist)
    out = {}
    for name, hint in (("every_row_routed", None), ("pre_aggregated_per_gpu", distinct)):
        n_max = batch if hint is None else min(batch, distinct + distinct // 8)
        table = partitioned_static_map(distinct, 0.5, backend=GpuBackend(dev, lib),
                                       fused_batch=n_max if routing != "nccl" else None, routing=routing,
                                       empty_value=0, probing="linear_probing", cg_size=1)
        times = []
        for rep in range(3):  # the first repetition is the warm-up
            table.clear_async()
            total = 0.0
            for b in range(batches):
                p = make(b)
                torch.cuda.synchronize(dev)
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                table.insert_or_apply(p, op="plus", pre_aggregate=hint)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                total += e0.elapsed_time(e1)
                del p
            times.append(total)
        ms = _max_over_ranks([min(times[1:]), statistics.median(times[1:])], dev)
        sums = table.find(probe)
        size = table.size()
        exact = torch.tensor([int(torch.equal(sums, hist))], dtype=torch.int64, device=dev)
        dist.all_reduce(exact, op=dist.ReduceOp.MIN)
        out[name] = {"ms_best": ms[0], "ms_median": ms[1], "grows_per_s": share * world / (ms[1] * 1e-3) / 1e9,
                     "size": size, "sums_equal_histogram_of_rows": bool(exact.item()),
                     "rows_counted": int(sums.sum().item())}
        table.close()
        torch.cuda.empty_cache()
        if not (out[name]["sums_equal_histogram_of_rows"] and out[name]["rows_counted"] == share * world):
            raise AssertionError(f"C5 leg ({name}) lost or double-counted rows: {out[name]}")
    out["workload"] = (f"hash-partitioned static_map<int64,int64> insert_or_apply(plus), {share * world} rows "
                       f"({share} per GPU in {batches} bulk calls), {distinct} distinct keys, {world} GPUs")
    out["unit"] = "G rows/s"
    return out


def bench(args, lib, impl, clock_sampler=None, parity_gate=None):
    import json  # noqa: F401
    import os

    from . import key_generator as kg

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream(dev)
    routing = os.environ.get("CUCO_B200_ROUTING", DEFAULT_ROUTING if impl == "native" else "nccl")

    # ---- checker leg: the same routing path against the CPU oracle, before anything is timed ----
    parity = parity_gate(world, rank, dev, lib, routing) if parity_gate is not None else None

    n = args.n  # pairs per rank: weak scaling
    # rank r owns the value range [r*n, (r+1)*n) of the global uniform stream (fixed seeds)
    keys = kg.uniform(n, 1, torch.int64, dev, seed=42 + rank) + rank * n
    pairs = torch.stack([keys, keys], dim=1).contiguous()
    out = torch.empty(n, dtype=torch.int64, device=dev)
    table = partitioned_static_map(n * world, 0.5, backend=GpuBackend(dev, lib),
                                   fused_batch=n if routing != "nccl" else None, routing=routing,
                                   probing="linear_probing", cg_size=1)

    sampler = clock_sampler(dev.index) if (clock_sampler is not None and rank == 0) else None
    if sampler is not None:
        sampler.__enter__()
    t_ins, t_find = _timed_pass(table, dev, stream, args.steps, args.warmup, pairs, keys, out)
    if sampler is not None:
        sampler.__exit__(None, None, None)
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    ins, fnd, total = _max_over_ranks([mean(t_ins), mean(t_find), mean(t_ins) + mean(t_find)], dev)
    ins_med, fnd_med, ins_best, fnd_best = _max_over_ranks(
        [statistics.median(t_ins), statistics.median(t_find), min(t_ins), min(t_find)], dev)
    total_size = table.size()
    distinct = torch.tensor([int(torch.unique(keys).numel())], dtype=torch.int64, device=dev)
    dist.all_reduce(distinct)  # the ranks' value ranges are disjoint
    assert total_size == int(distinct.item()), "global size() != number of distinct keys"

    # end to end: pinned host buffers in, host results out
    h_pairs = torch.empty((n, 2), dtype=torch.int64, pin_memory=True).copy_(pairs)
    h_keys = torch.empty(n, dtype=torch.int64, pin_memory=True).copy_(keys)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    d_pairs, d_keys = torch.empty_like(pairs), torch.empty_like(keys)

    def e2e_step():
        table.clear_async()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        d_pairs.copy_(h_pairs, non_blocking=True)
        table.insert_async(d_pairs)
        d_keys.copy_(h_keys, non_blocking=True)
        table.find(d_keys, out)
        h_out.copy_(out, non_blocking=True)
        b.record(stream)
        return a, b

    e2e_step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    ev = [e2e_step() for _ in range(max(1, min(args.steps, 3)))]
    torch.cuda.synchronize(dev)
    assert bool((h_out == h_keys).all().item()), "end-to-end partitioned find returned a wrong payload"
    e2e_ms = _max_over_ranks([statistics.mean(a.elapsed_time(b) for a, b in ev)], dev)[0]
    ops = 2 * n * world
    described = {"staged": "staged exchange: local grouping by (owner, table slice), copy-engine transfers over "
                           "NVLink overlapped slice by slice with the owners' L2-blocked probe",
                 "fused": "fused P2P routing over NVLink (one partition kernel stores into the owners' memory)",
                 "nccl": "NCCL all-to-all routing"}[routing]
    result = {
        "metric": "Gops/s insert & find (int64 pairs, LF 0.5)",
        "value": ops / (total * 1e-3) / 1e9,
        "unit": "Gops/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64",
        "data": "synthetic",
        "impl": impl,
        "config": {"workload": f"hash-partitioned static_map<int64,int64>, {n} uniform pairs per GPU "
                               f"insert + find, LF 0.5, linear_probing<1>, {world} GPUs, " + described,
                   "n_per_gpu": n, "total_size": total_size,
                   "timing": "CUDA events per rank around routing + exchange + local kernels, "
                             "max over ranks; clear outside; working sets larger than L2"},
        "insert_gops": n * world / (ins * 1e-3) / 1e9,
        "find_gops": n * world / (fnd * 1e-3) / 1e9,
        "insert_ms": ins,
        "find_ms": fnd,
        "insert_ms_median": ins_med, "find_ms_median": fnd_med,
        "insert_ms_best": ins_best, "find_ms_best": fnd_best,
        "value_median": ops / ((ins_med + fnd_med) * 1e-3) / 1e9,
        "value_best": ops / ((ins_best + fnd_best) * 1e-3) / 1e9,
        "parity": parity,
        # per GPU the insert moves the single-GPU algorithmic bytes (80 B/op) through HBM and
        # 16 B * (P-1)/P per pair through NVLink; report the HBM view (same definition as N=1) and
        # the NVLink egress next to it
        "roofline": {"bound": "hbm", "kernel": {"staged": "insert (exchange_route_kernel -> copy engines -> "
                                                          "tile_route_kernel + blocked_mutate_kernel per slice)",
                                                "fused": "insert (exchange_route_kernel + blocked_mutate_kernel)",
                                                "nccl": "insert (partition + all_to_all + local insert)"}[routing],
                     "achieved": 80.0 * n / (ins * 1e-3) / 1e9, "peak": _hbm_peak(), "unit": "GB/s",
                     "frac": 80.0 * n / (ins * 1e-3) / 1e9 / _hbm_peak(), "traffic": None,
                     "per_gpu": True,
                     "nvlink_egress_gbs": 16.0 * n * (world - 1) / world / (ins * 1e-3) / 1e9,
                     "nvlink_peak_gbs": 770.0},
        "e2e": {"value": ops / (e2e_ms * 1e-3) / 1e9, "unit": "Gops/s",
                "h2d_bytes_per_step": int(n * 24 * world), "d2h_bytes_per_step": int(n * 8 * world),
                "ms_per_step": e2e_ms},
        "gpu_launches": table.launches_per_step() * args.steps * world,
        "clocks": (sampler.summary() if sampler is not None else
                   {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampled on rank 0 only"]}),
    }
    if impl != "native":
        # cuco has no multi-GPU path: this arm is cuco's own local table behind THIS repository's partition
        # kernels and an NCCL all-to-all. It is a constructed baseline, not a reference measurement.
        result["reference_class"] = "constructed"
        result["config"]["reference_note"] = ("cuCollections is single-GPU; this arm = repo partition kernels + "
                                              "NCCL all_to_all_single + cuco's local static_map per rank")
    if table.fused is not None and getattr(table.fused, "trace", None) is not None:
        traces = [None] * world
        dist.all_gather_object(traces, table.fused.trace_summary())
        result["exchange_trace_ms"] = traces  # one dict per rank
    table.close()
    del table, pairs, keys, out, h_pairs, h_keys, h_out, d_pairs, d_keys
    torch.cuda.empty_cache()
    dist.barrier()
    if impl == "native" and not getattr(args, "no_c4", False):
        result["c4"] = _c4_leg(args, lib, dev, rank, world, routing)
        result["config"]["c4_workload"] = result["c4"]["workload"]
    if impl == "native" and not getattr(args, "no_c5", False):
        result["c5"] = _c5_leg(args, lib, dev, rank, world, routing)
        result["config"]["c5_workload"] = result["c5"]["workload"]
    dist.barrier()
    return result
