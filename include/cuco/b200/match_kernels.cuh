// Bulk "all matches" operations: count / count_outer and retrieve / retrieve_outer.
//
// the join-probe step that follows the insert/find path (the reference's
// `static_set::retrieve`, static_set/kernels.cuh:33-283, and the shared `count` / `retrieve`
// of multi-containers, open_addressing/kernels.cuh:437-471, 587-627 with the block-buffered device
// code in open_addressing_ref_impl.cuh:834-892, 1009-1282). Design here:
//
//  * thread-per-key on the resumable probe cursor of probe_engine, sector-wide read-only loads like
//    `lookup_kernel`; a key's probe walk ends at the first EMPTY slot, every EQUAL slot on the way is a
//    match (tables that forbid duplicates stop at the first match);
//  * walks over tables with duplicates are long (all copies of the key, then the rest of the cluster
//    up to the first empty slot), so they can run with look-ahead (`probe_engine::walk_ahead`,
//    tuning().match_ahead): the following chunks of the probe sequence that lie in the same 128-byte
//    line - which DRAM delivers whole anyway - are loaded together instead of one dependent round
//    trip per 32-byte chunk;
//  * retrieve reserves output space once per CTA round: per-thread match counts -> block exclusive
//    scan -> ONE global atomic per round -> every thread writes its matches at its own offset, so the
//    output of a round is one contiguous run. The first matches of a key (up to 32 bytes worth) stay
//    in registers during the counting walk; only keys with more matches walk their (now L1/L2-
//    resident) probe sequence a second time. No shared-memory flush buffers, no per-match atomics;
//  * `outer` variants account a key without matches as one output row {key, empty slot sentinel}.
//
// Tried and measured on B200, not kept (profiles/r01_matches_bench_v2_flattened.jsonl,
// profiles/r01_ncu_count_kernel_v1_details.txt): the per-key count loop runs with 7.7 of 32 lanes
// active (walk lengths differ widely inside a warp), but a flattened loop in which a lane that
// finishes a key fetches its next one at once was SLOWER (11.6 vs 13.6 G probes/s: the key loads
// stop being coalesced and nearly every iteration pays the refill path for a few lanes); a
// lookup-kernel-style probe phase (two keys per thread in flight) for tables without duplicates
// changed nothing (32.1 vs 31.7 G probes/s), so the one `block_retrieve` serves both.
//
// The order of the output rows is unspecified by contract (reference: "copies ... to unspecified
// locations"); parity is defined on the multiset of rows.
#pragma once

#include <cuco/b200/bulk_kernels.cuh>
#include <cuco/b200/probe_engine.cuh>
#include <cuco/detail/utility/cuda.cuh>

#include <cuda/atomic>

#include <cstdint>

namespace cuco::b200 {

/// Number of slots matching `key` (walk ends at the first empty slot); the first `Cached` matches
/// are kept in `cache` (registers: the index is resolved by an unrolled compare, never dynamically).
template <int ChunkSlots, load_policy Policy, int Ahead, int Cached, typename Engine, typename ProbeKey>
__device__ __forceinline__ unsigned int count_matches(Engine const& engine,
                                                      ProbeKey const& key,
                                                      typename Engine::value_type (&cache)[Cached])
{
  using size_type  = typename Engine::size_type;
  using slot_type  = typename Engine::value_type;
  unsigned int hits = 0;
  engine.template walk_ahead<ChunkSlots, Policy, Ahead>(engine.make_cursor(key), [&](size_type, slot_type slot) {
    auto const state = engine.classify_lookup(key, Engine::key_of(slot));
    if (state == equal_result::EMPTY) { return true; }
    if (state == equal_result::EQUAL) {
#pragma unroll
      for (int k = 0; k < Cached; ++k) {
        if (hits == static_cast<unsigned>(k)) { cache[k] = slot; }
      }
      ++hits;
      if constexpr (!Engine::allows_duplicates) { return true; }
    }
    return false;
  });
  return hits;
}

/// Matches of one key kept in registers by `block_retrieve`: one for tables without duplicates,
/// otherwise as many as fit 32 bytes. Keys with more matches walk their probe sequence again.
template <typename Engine>
constexpr int cached_matches() noexcept
{
  if constexpr (!Engine::allows_duplicates) {
    return 1;
  } else {
    return Engine::slot_bytes <= 8 ? 4 : 2;
  }
}

/// Exclusive scan of one value per thread over the CTA; returns {exclusive prefix, CTA total}.
/// `warp_totals` is shared scratch of BlockSize / 32 entries; ends with the CTA synchronised.
template <int BlockSize>
__device__ __forceinline__ void block_exclusive_scan(unsigned int mine,
                                                     unsigned int* warp_totals,
                                                     unsigned int& prefix,
                                                     unsigned int& total)
{
  static_assert(BlockSize % 32 == 0 && BlockSize <= 1024);
  constexpr int warps    = BlockSize / 32;
  unsigned int const lane = threadIdx.x & 31;
  unsigned int inclusive  = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned int const up = __shfl_up_sync(0xffffffffu, inclusive, d);
    if (lane >= static_cast<unsigned>(d)) { inclusive += up; }
  }
  if (lane == 31) { warp_totals[threadIdx.x >> 5] = inclusive; }
  __syncthreads();
  unsigned int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < warps; ++w) {
    unsigned int const t = warp_totals[w];
    if (w < static_cast<int>(threadIdx.x >> 5)) { before += t; }
    all += t;
  }
  prefix = before + inclusive - mine;
  total  = all;
  __syncthreads();
}

/// The whole CTA retrieves the matches of keys [first, first + n): writes {probe key, matched slot}
/// rows at positions reserved from `counter` (anything with `fetch_add(n, memory_order)`).
template <bool IsOuter,
          int BlockSize,
          int ChunkSlots,
          load_policy Policy,
          int Ahead,
          typename Engine,
          typename InputIt,
          typename OutputProbeIt,
          typename OutputMatchIt,
          typename AtomicCounter>
__device__ void block_retrieve(Engine const& engine,
                               InputIt first,
                               index_type n,
                               OutputProbeIt output_probe,
                               OutputMatchIt output_match,
                               AtomicCounter& counter)
{
  using size_type  = typename Engine::size_type;
  using slot_type  = typename Engine::value_type;
  using probe_type = decltype(read_input(first, index_type{0}));
  __shared__ unsigned int warp_totals[BlockSize / 32];
  __shared__ unsigned long long round_base;

  for (index_type base = 0; base < n; base += BlockSize) {
    index_type const idx = base + threadIdx.x;
    uninitialized<probe_type> key;
    constexpr int cached = cached_matches<Engine>();
    slot_type match[cached];
    match[0]          = engine.empty_slot_sentinel();  // the row of an outer retrieve without matches
    unsigned int hits = 0, rows = 0;
    if (idx < n) {
      key.value = read_input(first, idx);
      hits      = count_matches<ChunkSlots, Policy, Ahead>(engine, key.value, match);
      rows      = (IsOuter && hits == 0) ? 1u : hits;
    }
    unsigned int offset, total;
    block_exclusive_scan<BlockSize>(rows, warp_totals, offset, total);
    if (total == 0) { continue; }  // uniform across the CTA
    if (threadIdx.x == 0) {
      round_base = static_cast<unsigned long long>(
        counter.fetch_add(static_cast<size_type>(total), cuda::memory_order_relaxed));
    }
    __syncthreads();
    auto const where = static_cast<index_type>(round_base) + offset;
    if (rows != 0 && rows <= static_cast<unsigned>(cached)) {
#pragma unroll
      for (int k = 0; k < cached; ++k) {
        if (static_cast<unsigned>(k) < rows) {
          *(output_probe + (where + k)) = key.value;
          *(output_match + (where + k)) = match[k];
        }
      }
    } else if (rows != 0) {
      unsigned int written = 0;
      engine.template walk_ahead<ChunkSlots, Policy, Ahead>(
        engine.make_cursor(key.value), [&](size_type, slot_type slot) {
          auto const state = engine.classify_lookup(key.value, Engine::key_of(slot));
          if (state == equal_result::EMPTY) { return true; }
          if (state == equal_result::EQUAL) {
            *(output_probe + (where + written)) = key.value;
            *(output_match + (where + written)) = slot;
            // the table may have grown since the counting walk: never write past the reservation
            if (++written == rows) { return true; }
          }
          return false;
        });
    }
    __syncthreads();  // round_base is rewritten next round
  }
}

/// Counter view handed to block_retrieve by the bulk kernel.
template <typename T>
struct global_counter {
  T* address;
  __device__ T fetch_add(T n, cuda::memory_order order) const noexcept
  {
    cuda::atomic_ref<T, cuda::thread_scope_device> ref{*address};
    return ref.fetch_add(n, order);
  }
};

template <bool IsOuter,
          int BlockSize,
          int ChunkSlots,
          int Ahead,
          typename InputIt,
          typename OutputProbeIt,
          typename OutputMatchIt,
          typename Counter,
          typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize) void retrieve_kernel(InputIt first,
                                                              index_type n,
                                                              OutputProbeIt output_probe,
                                                              OutputMatchIt output_match,
                                                              Counter* num_rows,
                                                              Engine engine)
{
  // CTA b owns the keys [b * span, min(n, (b + 1) * span)), span a multiple of BlockSize
  index_type const rounds = cuco::detail::int_div_ceil(
    cuco::detail::int_div_ceil(n, index_type{BlockSize}), index_type{gridDim.x});
  index_type const span  = rounds * BlockSize;
  index_type const begin = index_type{blockIdx.x} * span;
  if (begin >= n) { return; }
  index_type const count = (n - begin) < span ? (n - begin) : span;
  global_counter<Counter> counter{num_rows};
  block_retrieve<IsOuter, BlockSize, ChunkSlots, load_policy::readonly, Ahead>(
    engine, first + begin, count, output_probe, output_match, counter);
}

template <bool IsOuter,
          int BlockSize,
          int ChunkSlots,
          int Ahead,
          typename InputIt,
          typename Counter,
          typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize) void count_kernel(InputIt first,
                                                           index_type n,
                                                           Counter* total,
                                                           Engine engine)
{
  using slot_type = typename Engine::value_type;
  unsigned long long mine = 0;
  for (index_type idx = cuco::detail::global_thread_id(); idx < n;
       idx += cuco::detail::grid_stride()) {
    auto const key = read_input(first, idx);
    slot_type unused[1];
    unsigned int const hits =
      count_matches<ChunkSlots, load_policy::readonly, Ahead>(engine, key, unused);
    mine += (IsOuter && hits == 0) ? 1u : hits;
  }
  accumulate_count(total, mine);
}

}  // namespace cuco::b200
