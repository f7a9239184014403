// Stream kernels of the L2-blocked mutation path (round 2): both passes fed by the bulk-copy
// engine (cp.async.bulk + mbarrier, "TMA" without a tensor map), so that no thread ever waits
// on the DRAM latency of the input stream it is about to consume.
//
// Why (measured, profiles/r02_microbench4.jsonl and profiles/r01_prof_final_details.txt):
//   * the claims of pass 2, confined to an L2-resident table slice, run at 86 G CAS/s when nothing
//     but the CAS is on the dependency chain - that is the DRAM bound of the slice traffic
//     (5.5 TB/s). The round-1 `blocked_mutate_kernel` reached 54 G keys/s: every CTA paid, one after
//     the other, the DRAM round trip of its input tile, the CAS round trip, and then a tail of
//     probe rounds during which most lanes idle (17 of 32 threads active on average).
//   * shared-memory atomics rank 100 M elements in 0.22 ms: the round-1 `route_kernel` (0.98 ms) was
//     not bound by its ranking but by four block-wide phases per tile that could not overlap the
//     input loads of the next tile.
//
// Pass 1, `tile_route_kernel`: persistent CTAs; the tile after the current one is already on its
// way into the second shared-memory buffer (one cp.async.bulk per tile, completion on an
// mbarrier) while the current tile is ranked, grouped in place and copied out.
//
// Pass 2, `stream_mutate_kernel`: persistent WARPS. Each warp owns a private double buffer in
// shared memory that the bulk-copy engine fills with 128-key chunks of the staged segments, in
// region order (global ticket). Phase 1 is fully converged - rows of 32 keys, one claim CAS per key
// on its home slot, several rows in flight - and finishes ~82 % of the keys; the others are parked
// in shared memory and worked off 32 at a time by the general driver, so no lane ever holds its
// warp in a drain loop. The only latency left on the critical path is one L2 round trip per group
// of rows.
#pragma once

#include <cuco/b200/bulk_kernels.cuh>

#include <cstdint>

namespace cuco::b200 {

// -------------------------------------------------------------------------------------------------
// PTX: mbarrier + bulk copy (global -> shared), generic enough for both kernels
// -------------------------------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ std::uint32_t shared_address(void const* p) noexcept
{
  return static_cast<std::uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbarrier_init(std::uint64_t* barrier, unsigned arrivals) noexcept
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(shared_address(barrier)), "r"(arrivals)
               : "memory");
}

/// Makes freshly initialised barriers visible to the async proxy (the bulk-copy engine).
__device__ __forceinline__ void fence_barrier_init() noexcept
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

/// Orders this thread's generic-proxy accesses to shared memory before later async-proxy accesses.
__device__ __forceinline__ void fence_proxy_async() noexcept
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/// One arrival that also announces `bytes` of bulk-copy traffic to wait for.
__device__ __forceinline__ void mbarrier_arrive_expect(std::uint64_t* barrier, unsigned bytes) noexcept
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(shared_address(barrier)),
               "r"(bytes)
               : "memory");
}

/// Non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbarrier_test(std::uint64_t* barrier, unsigned parity) noexcept
{
  unsigned done;
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
    "selp.u32 %0, 1, 0, p;\n\t}"
    : "=r"(done)
    : "r"(shared_address(barrier)), "r"(parity)
    : "memory");
  return done != 0;
}

/// Blocking wait (the hardware suspends the thread for a bounded time per attempt).
__device__ __forceinline__ void mbarrier_wait(std::uint64_t* barrier, unsigned parity) noexcept
{
  unsigned done;
  do {
    asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(shared_address(barrier)), "r"(parity)
      : "memory");
  } while (done == 0);
}

/// cp.async.bulk global -> shared (UBLKCP.S.G): `bytes` a multiple of 16, both addresses 16-byte
/// aligned; completion is signalled as `bytes` of transaction count on `barrier`.
__device__ __forceinline__ void bulk_load(void* shared_destination,
                                          void const* global_source,
                                          unsigned bytes,
                                          std::uint64_t* barrier) noexcept
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
      shared_address(shared_destination)),
    "l"(global_source),
    "r"(bytes),
    "r"(shared_address(barrier))
    : "memory");
}

/// The same copy with an L2 eviction-priority policy (createpolicy) attached to the source lines.
__device__ __forceinline__ void bulk_load_hinted(void* shared_destination,
                                                 void const* global_source,
                                                 unsigned bytes,
                                                 std::uint64_t* barrier,
                                                 std::uint64_t policy) noexcept
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
      shared_address(shared_destination)),
    "l"(global_source),
    "r"(bytes),
    "r"(shared_address(barrier)),
    "l"(policy)
    : "memory");
}

/// Policy "evict first": lines of a read-once stream must not displace the table slices in L2.
__device__ __forceinline__ std::uint64_t policy_evict_first() noexcept
{
  std::uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

/// Fire-and-forget: bring the line holding `address` into L2.
__device__ __forceinline__ void prefetch_l2(void const* address) noexcept
{
  asm volatile("prefetch.global.L2 [%0];" ::"l"(address));
}

}  // namespace ptx

// =================================================================================================
// pass 1: tile router with bulk-copy input
// =================================================================================================

constexpr int tile_route_block_size = 256;  ///< tile = 2048 elements: three CTAs per SM for 16-byte slots

/// Dynamic shared memory of `tile_route_kernel`: two element tiles, the bucket of every staged
/// position, three bucket arrays.
template <int BlockSize, typename Slot, bool WithIndex = false>
constexpr std::size_t tile_route_smem_bytes(std::uint32_t num_buckets) noexcept
{
  std::size_t const tile = std::size_t{BlockSize} * route_items_per_thread;
  return 2 * tile * sizeof(Slot) + tile * sizeof(std::uint16_t) + (WithIndex ? tile * sizeof(std::uint32_t) : 0) +
         3 * std::size_t{num_buckets} * sizeof(unsigned int);
}

/// Bucket of a key for the single-GPU blocked path: the L2 region of its home slot.
struct region_router {
  region_map regions;
  static constexpr bool spills = false;  ///< a full segment is finished on the spot (local table)

  [[nodiscard]] __host__ __device__ std::uint32_t num_buckets() const noexcept { return regions.num_regions; }

  template <typename Engine, typename Key>
  [[nodiscard]] __device__ std::uint32_t operator()(Engine const& engine, Key const& key) const noexcept
  {
    return regions(engine.make_cursor(key).slot);
  }
};

/// Bucket of a key for the staged multi-GPU exchange: (owner rank, region of the OWNER's shard) -
/// every shard has the same geometry, so the source can compute both.
struct exchange_router {
  region_map regions;
  std::uint32_t num_ranks;
  std::uint64_t salt;
  static constexpr bool spills = true;  ///< a full segment overflows into the spill list (foreign table)

  [[nodiscard]] __host__ __device__ std::uint32_t num_buckets() const noexcept
  {
    return num_ranks * regions.num_regions;
  }

  template <typename Engine, typename Key>
  [[nodiscard]] __device__ std::uint32_t operator()(Engine const& engine, Key const& key) const noexcept
  {
    auto const owner = exchange_owner(exchange_key_bits(key), salt, num_ranks);
    if (regions.num_regions == 1) { return owner; }  // routing by owner only: no need for the table hash
    return owner * regions.num_regions + regions(engine.make_cursor(key).slot);
  }
};

/// `region_router` for batches whose overflow cannot be finished on the spot (probe keys of the
/// blocked count / retrieve): elements of a full segment go to the spill list.
struct region_spill_router : region_router {
  static constexpr bool spills = true;
};

/// Key of a routed element: the element itself when probe keys are routed, else the slot image's key.
template <typename Engine, typename Elem>
[[nodiscard]] __device__ __forceinline__ auto const& routed_key(Elem const& element) noexcept
{
  if constexpr (cuda::std::is_same_v<Elem, typename Engine::value_type>) {
    return Engine::key_of(element);
  } else {
    return element;
  }
}

/// Where elements go whose segment is full (routers with `spills`).
struct route_spill {
  void* list                = nullptr;
  unsigned int* count       = nullptr;
  std::uint32_t capacity    = 0;
  /// routers `WithIndex` (lookups: the answers must find their way back): position[i] = where input
  /// element i was staged (bucket * segment_capacity + offset; 0xffffffff = spilled), and the input
  /// index of every spilled element
  std::uint32_t* position   = nullptr;
  std::uint32_t* index      = nullptr;
};

/// `route_kernel` for a contiguous input range of slot images (`Slot const*`, 16-byte aligned):
/// same ranking, staging and copy-out, but persistent, with the next tile loaded by the bulk-copy
/// engine while the current one is processed, and with the staged tile grouped IN PLACE in the
/// buffer the tile arrived in (every thread holds its elements in registers by then).
template <int BlockSize,
          int ChunkSlots,
          bool Counted,
          typename StencilIt,
          typename Predicate,
          typename Counter,
          typename Engine,
          typename Action,
          typename Router = region_router,
          typename Elem   = typename Engine::value_type,
          bool WithIndex  = false>
CUCO_KERNEL __launch_bounds__(BlockSize) void tile_route_kernel(
  Elem const* first,
  index_type n,
  StencilIt stencil,
  Predicate pred,
  Elem* segments,
  unsigned int* region_counts,
  Router router,
  std::uint32_t segment_capacity,
  Counter* num_new,
  Engine engine,
  Action action,
  route_spill spill = {})
{
  using slot_type           = Elem;  // routed element: a slot image, or a bare probe key
  constexpr int items       = route_items_per_thread;
  constexpr index_type tile = index_type{BlockSize} * items;

  extern __shared__ __align__(128) unsigned char route_dynamic_smem[];
  std::uint32_t const num_regions = router.num_buckets();
  auto* const buffer0    = reinterpret_cast<slot_type*>(route_dynamic_smem);
  auto* const buffer1    = buffer0 + tile;
  auto* const origin     = reinterpret_cast<std::uint32_t*>(buffer1 + tile);                  // [tile] (WithIndex)
  auto* const owner      = reinterpret_cast<std::uint16_t*>(origin + (WithIndex ? tile : 0));  // [tile]
  auto* const tile_hist  = reinterpret_cast<unsigned int*>(owner + tile);                     // [R]
  auto* const tile_start = tile_hist + num_regions;                                           // [R]
  auto* const run_start  = tile_start + num_regions;                                          // [R]
  __shared__ unsigned int warp_sums[BlockSize / 32];
  __shared__ __align__(8) std::uint64_t arrived[2];

  index_type const num_tiles = (n + tile - 1) / tile;
  // Elements of tile `which` that arrive through the bulk copy: whole 16-byte units only, so the
  // copy never reads past the end of the caller's range; the (at most 15-byte) tail of the LAST
  // tile is read by its threads directly.
  auto bulk_elements = [&](index_type which) -> unsigned {
    index_type const base = which * tile;
    auto const count      = static_cast<unsigned>((n - base) < tile ? (n - base) : tile);
    unsigned const bytes  = (count * static_cast<unsigned>(sizeof(slot_type))) & ~15u;
    return bytes / static_cast<unsigned>(sizeof(slot_type));
  };
  auto issue = [&](index_type which, int slot) {
    // thread 0 only: start the bulk copy of tile `which` into buffer `slot`
    unsigned const bytes = bulk_elements(which) * static_cast<unsigned>(sizeof(slot_type));
    if (bytes != 0) {
      ptx::mbarrier_arrive_expect(&arrived[slot], bytes);
      ptx::bulk_load(slot ? buffer1 : buffer0, first + which * tile, bytes, &arrived[slot]);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::shared_address(&arrived[slot])) : "memory");
    }
  };

  if (threadIdx.x == 0) {
    ptx::mbarrier_init(&arrived[0], 1);
    ptx::mbarrier_init(&arrived[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    index_type const t0 = blockIdx.x;
    if (t0 < num_tiles) { issue(t0, 0); }
    if (t0 + gridDim.x < num_tiles) { issue(t0 + gridDim.x, 1); }
  }

  unsigned long long mine = 0;
  int round               = 0;
  for (index_type t = blockIdx.x; t < num_tiles; t += gridDim.x, ++round) {
    int const slot          = round & 1;
    unsigned const parity   = (round >> 1) & 1;
    slot_type* const stage  = slot ? buffer1 : buffer0;
    index_type const base   = t * tile;

    for (std::uint32_t r = threadIdx.x; r < num_regions; r += BlockSize) {
      tile_hist[r] = 0;
    }
    ptx::mbarrier_wait(&arrived[slot], parity);
    __syncthreads();  // histogram zeroed everywhere

    uninitialized<slot_type> val[items];
    std::uint32_t region[items];
    std::uint32_t rank[items];
    unsigned const in_buffer = bulk_elements(t);
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * BlockSize + threadIdx.x;
      region[j]            = 0xffffffffu;
      if (idx < n) {
        unsigned const at = j * BlockSize + threadIdx.x;
        val[j].value      = at < in_buffer ? stage[at] : first[idx];
        bool keep;
        if constexpr (cuda::std::is_same_v<StencilIt, slot_type const*>) {
          // the batch is its own stencil (rehash: "insert the slots that are filled"): the element is
          // already here, no second - dependent - load of the same 16 bytes
          keep = (stencil == first) ? pred(val[j].value) : pred(*(stencil + idx));
        } else {
          keep = pred(*(stencil + idx));
        }
        if (keep) { region[j] = router(engine, routed_key<Engine>(val[j].value)); }
      }
    }
    if (num_regions <= 8) {
      // few buckets (measured, profiles/r02_stage_probe.jsonl: the election wins up to ~8 buckets, plain
      // shared-memory atomics from 16 up): the lanes of a warp that share a bucket reserve their ranks with ONE atomic
      // (one atomic per element would serialise the tile on a handful of counters)
      unsigned const lane = threadIdx.x & 31u;
#pragma unroll
      for (int j = 0; j < items; ++j) {
        unsigned const same   = __match_any_sync(0xffffffffu, region[j]);
        unsigned const leader = static_cast<unsigned>(__ffs(same)) - 1u;
        unsigned int first_rank = 0;
        if (lane == leader && region[j] != 0xffffffffu) {
          first_rank = atomicAdd(&tile_hist[region[j]], static_cast<unsigned int>(__popc(same)));
        }
        first_rank = __shfl_sync(0xffffffffu, first_rank, static_cast<int>(leader));
        rank[j]    = first_rank + static_cast<unsigned int>(__popc(same & ((1u << lane) - 1u)));
      }
    } else {
#pragma unroll
      for (int j = 0; j < items; ++j) {
        if (region[j] != 0xffffffffu) { rank[j] = atomicAdd(&tile_hist[region[j]], 1u); }
      }
    }
    __syncthreads();  // every element of the tile is in registers: the buffer may be overwritten

    // exclusive scan of the tile histogram: thread t owns regions [t * rpt, (t + 1) * rpt)
    {
      std::uint32_t const rpt = (num_regions + BlockSize - 1) / BlockSize;
      std::uint32_t const r0  = threadIdx.x * rpt;
      unsigned int sum        = 0;
      for (std::uint32_t i = 0; i < rpt; ++i) {
        if (r0 + i < num_regions) { sum += tile_hist[r0 + i]; }
      }
      unsigned int inclusive = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned int const up = __shfl_up_sync(0xffffffffu, inclusive, d);
        if ((threadIdx.x & 31) >= d) { inclusive += up; }
      }
      if ((threadIdx.x & 31) == 31) { warp_sums[threadIdx.x >> 5] = inclusive; }
      __syncthreads();
      unsigned int running = inclusive - sum;
      for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) {
        running += warp_sums[w];
      }
      for (std::uint32_t i = 0; i < rpt; ++i) {
        std::uint32_t const r = r0 + i;
        if (r < num_regions) {
          unsigned int const held = tile_hist[r];
          tile_start[r]           = running;
          run_start[r]            = held ? atomicAdd(&region_counts[r], held) : 0u;
          running += held;
        }
      }
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (region[j] != 0xffffffffu) {
        unsigned int const pos = tile_start[region[j]] + rank[j];
        stage[pos]             = val[j].value;
        owner[pos]             = static_cast<std::uint16_t>(region[j]);
        if constexpr (WithIndex) {
          origin[pos] = static_cast<std::uint32_t>(base + index_type{j} * BlockSize + threadIdx.x);
        }
      }
    }
    __syncthreads();

    unsigned int const count = tile_start[num_regions - 1] + tile_hist[num_regions - 1];
    for (unsigned int pos = threadIdx.x; pos < count; pos += BlockSize) {
      std::uint32_t const r     = owner[pos];
      std::uint64_t const where = std::uint64_t{run_start[r]} + (pos - tile_start[r]);
      if (where < segment_capacity) {
        segments[std::uint64_t{r} * segment_capacity + where] = stage[pos];
        // the tile's positions are written together, so these scattered 4-byte stores merge in L2
        if constexpr (WithIndex) {
          spill.position[origin[pos]] = static_cast<std::uint32_t>(std::uint64_t{r} * segment_capacity + where);
        }
      } else if constexpr (Router::spills) {
        // segment full (heavily skewed input): the caller finishes the spilled elements another way
        unsigned int const at = atomicAdd(spill.count, 1u);
        if (at < spill.capacity) {
          static_cast<slot_type*>(spill.list)[at] = stage[pos];
          if constexpr (WithIndex) { spill.index[at] = origin[pos]; }
        }
        if constexpr (WithIndex) { spill.position[origin[pos]] = 0xffffffffu; }
      } else if constexpr (cuda::std::is_same_v<Elem, typename Engine::value_type>) {
        // segment full (heavily skewed input): finish this element now, unblocked
        mine += mutate_slow_path<ChunkSlots, load_policy::streaming>(engine, stage[pos], base + pos, action);
      }
    }
    __syncthreads();  // the buffer and the bucket arrays are free again

    if (threadIdx.x == 0) {
      index_type const ahead = t + 2 * index_type{gridDim.x};
      if (ahead < num_tiles) {
        ptx::fence_proxy_async();  // our generic-proxy stores into this buffer precede the engine's
        issue(ahead, slot);
      }
    }
  }
  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

// =================================================================================================
// pass 2: warp-persistent probe stream over the staged segments
// =================================================================================================

constexpr int stream_chunk_keys = 64;  ///< keys per bulk copy (one warp's refill unit)

/// Geometry of pass 2, shared by host and device (segments as laid out by the routers:
/// segment = region * sources + source, `segment_capacity` elements each).
struct stream_layout {
  unsigned int const* counts;        ///< elements routed to each segment (may exceed the capacity)
  std::uint32_t segment_capacity;    ///< elements a segment can hold (multiple of 16)
  std::uint32_t chunks_per_segment;  ///< ceil(segment_capacity / stream_chunk_keys)
  std::uint32_t num_segments;        ///< regions * sources
  std::uint32_t sources;             ///< segments per region (1; number of ranks for exchanged batches)
  std::uint64_t region_slots;        ///< ceil(capacity / num_regions): slots per region
  std::uint64_t table_bytes;         ///< end of the slot array (prefetch clamp)
  std::uint32_t prefetch;            ///< bit 0: bulk-prefetch the next region's slots into L2;
                                     ///< bit 1: prefetch every key's home line one chunk ahead of its CAS
};

/// Takes this warp's next non-empty work item (warp-uniform), starts its bulk copy into
/// `destination` and, if asked to, the bulk L2 prefetch of its share of the NEXT region's slots.
/// `next_ticket` (meaningful in lane 0) holds a ticket requested EARLIER - the atomic's round trip
/// is off the critical path - and is replaced by a fresh request. Returns the number of keys on
/// their way (0: no work left).
template <typename Slot, int SlotBytes>
__device__ __forceinline__ unsigned stream_start_load(unsigned long long* ticket,
                                                      unsigned long long& next_ticket,
                                                      std::uint64_t total_items,
                                                      stream_layout const& layout,
                                                      Slot const* segments,
                                                      Slot* destination,
                                                      std::uint64_t* barrier,
                                                      Slot const* table,
                                                      unsigned lane) noexcept
{
  constexpr unsigned chunk_keys = stream_chunk_keys;
  while (true) {
    // Work items are handed out in order by one global ticket counter, so the chunks in flight
    // anywhere on the chip are always neighbours in region order: that is what keeps the table
    // slices they probe resident in L2 (a static round-robin lets warps drift regions apart -
    // measured: 14 GB of DRAM traffic per 100 M keys instead of 6.4).
    unsigned long long const item = __shfl_sync(0xffffffffu, next_ticket, 0);
    if (item >= total_items) { return 0u; }
    if (lane == 0) { next_ticket = atomicAdd(ticket, 1ull); }
    auto const segment = static_cast<std::uint32_t>(item / layout.chunks_per_segment);
    auto const chunk   = static_cast<std::uint32_t>(item - std::uint64_t{segment} * layout.chunks_per_segment);
    unsigned int const stored = layout.counts[segment];
    unsigned int const count  = stored < layout.segment_capacity ? stored : layout.segment_capacity;
    std::uint64_t const begin = std::uint64_t{chunk} * chunk_keys;
    if (begin < count) {
      unsigned const keys = static_cast<unsigned>(count - begin < chunk_keys ? count - begin : chunk_keys);
      if (lane == 0) {
        unsigned const bytes = (keys * static_cast<unsigned>(sizeof(Slot)) + 15u) & ~15u;
        ptx::fence_proxy_async();  // the lanes' reads of this buffer (ordered by __syncwarp) come first
        ptx::mbarrier_arrive_expect(barrier, bytes);
        // the staged batch is read exactly once: its lines leave L2 first, the table slices stay
        ptx::bulk_load_hinted(destination,
                              segments + std::uint64_t{segment} * layout.segment_capacity + begin,
                              bytes,
                              barrier,
                              ptx::policy_evict_first());
        if (layout.prefetch & 1u) {
          std::uint32_t const region      = segment / layout.sources;
          std::uint32_t const num_regions = layout.num_segments / layout.sources;
          if (region + 1 < num_regions) {
            // the non-empty chunks of this region's segments share the next region between them
            std::uint32_t const used   = (count + chunk_keys - 1) / chunk_keys;
            std::uint64_t const slice  = layout.region_slots * SlotBytes / layout.sources;
            std::uint64_t const share  = ((slice + used - 1) / used + 127) & ~std::uint64_t{127};
            std::uint64_t const origin = (std::uint64_t{region} + 1) * layout.region_slots * SlotBytes +
                                         std::uint64_t{segment - region * layout.sources} * slice;
            std::uint64_t const begin_byte = (origin + std::uint64_t{chunk} * share) & ~std::uint64_t{15};
            std::uint64_t end_byte         = begin_byte + share;
            std::uint64_t const stop       = (origin + slice + 127) & ~std::uint64_t{15};
            if (end_byte > stop) { end_byte = stop; }
            if (end_byte > (layout.table_bytes & ~std::uint64_t{15})) { end_byte = layout.table_bytes & ~std::uint64_t{15}; }
            if (begin_byte < end_byte) {
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                             reinterpret_cast<char const*>(table) + begin_byte),
                           "r"(static_cast<std::uint32_t>(end_byte - begin_byte))
                           : "memory");
            }
          }
        }
      }
      return keys;
    }
    // a chunk beyond the segment's fill count (the segments have slack): take the next ticket
  }
}

/// Keys a warp may have parked for the second phase before it must work some off: one batch that
/// is waiting for company plus everything one chunk can add.
constexpr int stream_parked_capacity = 32 + stream_chunk_keys;

/// Dynamic shared memory of `stream_mutate_kernel`, per warp: three chunk buffers, the home slots
/// of two chunks, the parked keys.
template <int BlockSize, typename Slot>
constexpr std::size_t stream_mutate_smem_bytes() noexcept
{
  return std::size_t{BlockSize / 32} *
         ((3 * stream_chunk_keys + stream_parked_capacity) * sizeof(Slot) +
          2 * stream_chunk_keys * sizeof(std::uint64_t));
}

/// Pass 2. Every warp streams 64-key chunks of the staged segments (region order, handed out by a
/// global ticket) through a private ring of three shared-memory buffers filled by the bulk-copy
/// engine. A chunk goes through three stages, one per loop iteration of its warp:
///   landing   the bulk copy is in flight (issued two iterations before the chunk is needed);
///   scouting  fully converged: every lane hashes its keys, stores the home slots in shared memory
///             and sends an L2 prefetch for each home line. Measured on B200: a CAS that MISSES L2
///             is served at ~22 G/s chip-wide whatever the DRAM load (profiles/r01_hardware_probes.md),
///             a load or prefetch miss at DRAM speed - so the claim must never be the first touch;
///   claiming  one iteration later, fully converged again: ONE claim CAS per key on its home slot,
///             `Rows` rows of 32 keys in flight per warp. A key is finished if the CAS won or found
///             the key already there (~82 % at load factor 0.5); the others are PARKED in shared
///             memory and worked off 32 at a time through the general insertion driver, so the
///             claiming stage never diverges and the second phase starts with full warps.
template <int BlockSize,
          int Rows,
          int ChunkSlots,
          int MinBlocks,
          bool Counted,
          typename Counter,
          typename Engine,
          typename Action>
CUCO_KERNEL __launch_bounds__(BlockSize, MinBlocks) void stream_mutate_kernel(
  typename Engine::value_type const* segments,
  stream_layout layout,
  unsigned long long* ticket,  ///< zeroed by the host before the launch
  Counter* num_new,
  Engine engine,
  Action action)
{
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  static_assert(Engine::single_cas, "the stream path needs one-shot claimable slots");
  static_assert(Action::blockable, "per-element outputs cannot follow a regrouped batch");
  static_assert(stream_chunk_keys % (32 * Rows) == 0 || Rows == 1);
  constexpr int warps           = BlockSize / 32;
  constexpr int chunk_keys      = stream_chunk_keys;
  constexpr bool key_only_claim = Action::key_then_apply && sizeof(slot_type) > 8;
  constexpr auto policy         = load_policy::streaming;
  constexpr std::size_t warp_bytes =
    (3 * chunk_keys + stream_parked_capacity) * sizeof(slot_type) + 2 * chunk_keys * sizeof(std::uint64_t);

  extern __shared__ __align__(128) unsigned char stream_dynamic_smem[];
  __shared__ __align__(8) std::uint64_t arrived_all[warps][3];

  unsigned const lane = threadIdx.x & 31u;
  unsigned const warp = threadIdx.x >> 5;
  auto* const chunk0  = reinterpret_cast<slot_type*>(stream_dynamic_smem + std::size_t{warp} * warp_bytes);
  auto* const parked  = chunk0 + 3 * chunk_keys;
  auto* const home0   = reinterpret_cast<std::uint64_t*>(parked + stream_parked_capacity);  // [2][chunk_keys]
  std::uint64_t* const arrived = arrived_all[warp];

  if (lane == 0) {
    ptx::mbarrier_init(&arrived[0], 1);
    ptx::mbarrier_init(&arrived[1], 1);
    ptx::mbarrier_init(&arrived[2], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();

  std::uint64_t const total_items = std::uint64_t{layout.num_segments} * layout.chunks_per_segment;
  auto* const table               = engine.slots();
  auto const empty_slot           = engine.empty_slot_sentinel();
  bool const scout_lines          = (layout.prefetch & 2u) != 0;

  unsigned long long next_ticket = 0;
  if (lane == 0) { next_ticket = atomicAdd(ticket, 1ull); }
  auto start_next_load = [&](unsigned b) -> unsigned {
    return stream_start_load<slot_type, Engine::slot_bytes>(
      ticket, next_ticket, total_items, layout, segments, chunk0 + b * chunk_keys, &arrived[b], table, lane);
  };

  // scouting stage of the chunk in buffer `b`: home slots into shared memory, home lines into L2
  auto scout = [&](unsigned b, unsigned count, unsigned stash) {
    slot_type const* const keys = chunk0 + b * chunk_keys;
    std::uint64_t* const home   = home0 + stash * chunk_keys;
#pragma unroll
    for (int r = 0; r < chunk_keys / 32; ++r) {
      unsigned const at = r * 32 + lane;
      if (at < count) {
        auto const slot = static_cast<std::uint64_t>(engine.make_cursor(Engine::key_of(keys[at])).slot);
        home[at]        = slot;
        if (scout_lines) { ptx::prefetch_l2(table + slot); }
      }
    }
  };

  // warp-uniform pipeline state in scalars (indexed arrays would live in local memory).
  // Ring position p = iteration % 3 holds the chunk being claimed, p+1 the one being scouted,
  // p+2 the one landing.
  unsigned phases     = 0;  // bit b: parity of the mbarrier phase buffer b completes next
  unsigned keys_claim = start_next_load(0);
  unsigned keys_scout = start_next_load(1);
  unsigned keys_land  = start_next_load(2);
  unsigned ring       = 0;  // buffer of the chunk being claimed
  unsigned stash      = 0;  // which half of `home0` holds its home slots
  unsigned num_parked = 0;
  unsigned long long mine = 0;

  auto wait_landed = [&](unsigned b) {
    ptx::mbarrier_wait(&arrived[b], (phases >> b) & 1u);
    phases ^= 1u << b;
  };

  // phase 2 for up to 32 parked keys: each lane finishes one through the general driver
  auto work_off = [&](unsigned count) {
    if (lane < count) {
      slot_type const val = parked[num_parked - count + lane];
      mine += mutate_slow_path<ChunkSlots, policy>(engine, val, index_type{0}, action);
    }
    __syncwarp();
    num_parked -= count;
  };

  if (keys_claim != 0) {
    wait_landed(0);
    scout(0, keys_claim, 0);
  }

  while (keys_claim != 0) {
    unsigned const next_buffer = ring == 2 ? 0u : ring + 1;
    if (keys_scout != 0) {
      wait_landed(next_buffer);
      scout(next_buffer, keys_scout, stash ^ 1u);
    }
    __syncwarp();

    // ---- claiming stage of the chunk in buffer `ring` ----
    slot_type const* const keys     = chunk0 + ring * chunk_keys;
    std::uint64_t const* const home = home0 + stash * chunk_keys;
    for (unsigned base = 0; base < keys_claim; base += 32 * Rows) {
      uninitialized<slot_type> val[Rows];
      slot_type seen[Rows];
      std::uint64_t slot[Rows];
      unsigned live = 0;
#pragma unroll
      for (int r = 0; r < Rows; ++r) {
        unsigned const at = base + r * 32 + lane;
        if (at < keys_claim) {
          val[r].value = keys[at];
          slot[r]      = home[at];
          live |= 1u << r;
        }
      }
#pragma unroll
      for (int r = 0; r < Rows; ++r) {
        if (live & (1u << r)) {
          if constexpr (key_only_claim) {
            key_type expected_key = Engine::key_of(empty_slot);
            cuda::atomic_ref<key_type, Engine::thread_scope> key_ref{(table + slot[r])->first};
            key_ref.compare_exchange_strong(expected_key,
                                            static_cast<key_type>(Engine::key_of(val[r].value)),
                                            cuda::memory_order_relaxed);
            seen[r]       = empty_slot;
            seen[r].first = expected_key;
          } else {
            seen[r] = cas_slot<Engine::thread_scope>(table + slot[r], empty_slot, val[r].value);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < Rows; ++r) {
        bool park = false;
        if (live & (1u << r)) {
          auto* const address = table + slot[r];
          bool const won      = key_only_claim
                                  ? same_bits(Engine::key_of(seen[r]), Engine::key_of(empty_slot))
                                  : same_bits(seen[r], empty_slot);
          if (won) {
            action.on_new(engine, index_type{0}, address, val[r].value);
            ++mine;
          } else if (engine.classify_insert(Engine::key_of(val[r].value), Engine::key_of(seen[r])) ==
                     equal_result::EQUAL) {
            action.on_present(engine, index_type{0}, address, seen[r], val[r].value);
          } else {
            park = true;
          }
        }
        unsigned const parking = __ballot_sync(0xffffffffu, park);
        if (park) { parked[num_parked + __popc(parking & ((1u << lane) - 1u))] = val[r].value; }
        num_parked += __popc(parking);
      }
    }
    __syncwarp();
    while (num_parked >= 32) { work_off(32); }

    // the claimed chunk's buffer takes the chunk after the two that are already under way
    unsigned const refilled = start_next_load(ring);
    keys_claim = keys_scout;
    keys_scout = keys_land;
    keys_land  = refilled;
    ring       = next_buffer;
    stash ^= 1u;
  }
  if (num_parked != 0) { work_off(num_parked); }
  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

}  // namespace cuco::b200
