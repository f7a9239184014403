// L2-blocked "all matches" operations: count / count_outer and retrieve / retrieve_outer on tables
// far larger than L2, for callers whose result does not depend on the order of the probe keys - a
// total (count) or rows at unspecified positions (retrieve; reference: "copies ... to unspecified
// locations", static_set.cuh:620, open_addressing_impl.cuh:604-660).
//
// Why: a direct probe of a table beyond L2 moves a whole 128-byte DRAM line per key
// (profiles/r01_hardware_probes.md; 116-143 B of DRAM traffic per probe in
// profiles/r01_ncu_*count*_details.txt). Lookups that must answer in input order cannot avoid that
// on one GPU (a second partition plus the un-permute cost more than they save); these
// operations can: the probe keys are grouped by the L2 region their home slot lies in (pass 1, the
// same bulk-copy fed router as the blocked insert, on 4/8-byte keys), and the regions are probed one
// after the other with the region's slots resident in L2 (pass 2), so every table line is fetched
// once per batch and shared by all keys that land in it. DRAM traffic per probe: key in + key
// staged out and in again + table bytes / batch size + output rows.
//
// Pass 2 reuses the per-key device code of match_kernels.cuh unchanged (`count_matches`,
// `block_retrieve`), so the row semantics - including `outer` - are those of the direct kernels.
#pragma once

#include <cuco/b200/bulk_kernels.cuh>
#include <cuco/b200/match_kernels.cuh>
#include <cuco/b200/probe_engine.cuh>

#include <cstdint>

namespace cuco::b200 {

/// One thread of the CTA streams its share of the NEXT region's slots into L2 (same scheme as
/// `blocked_mutate_kernel`): ctas = CTAs per region, cta = this CTA's index inside the region.
template <typename Engine>
__device__ __forceinline__ void prefetch_next_region(Engine const& engine,
                                                     blocked_layout const& layout,
                                                     std::uint32_t region,
                                                     std::uint32_t num_regions,
                                                     std::uint32_t cta) noexcept
{
  if (layout.prefetch_bytes == 0 || region + 1 >= num_regions) { return; }
  std::uint64_t const begin =
    (layout.first_slot + (std::uint64_t{region} + 1) * layout.region_slots) * Engine::slot_bytes +
    std::uint64_t{cta} * layout.prefetch_bytes;
  std::uint64_t const limit =
    (layout.first_slot + (std::uint64_t{region} + 2) * layout.region_slots) * Engine::slot_bytes;
  std::uint64_t end = begin + layout.prefetch_bytes;
  if (end > limit) { end = limit; }
  if (end > (layout.table_bytes & ~std::uint64_t{15})) { end = layout.table_bytes & ~std::uint64_t{15}; }
  if (begin < end) {
    auto const* address = reinterpret_cast<char const*>(engine.slots()) + (begin & ~std::uint64_t{15});
    auto const bytes    = static_cast<std::uint32_t>((end - (begin & ~std::uint64_t{15})) & ~std::uint64_t{15});
    if (bytes != 0) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(address), "r"(bytes) : "memory");
    }
  }
}

/// Pass 2 of a blocked count: grid (CTAs per region, regions); CTA x of region y strides over the
/// region's staged keys.
template <bool IsOuter, int BlockSize, int ChunkSlots, int Ahead, typename Counter, typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize) void blocked_count_kernel(
  typename Engine::key_type const* segments, blocked_layout layout, Counter* total, Engine engine)
{
  using slot_type            = typename Engine::value_type;
  std::uint32_t const region = blockIdx.y;
  if (threadIdx.x == 0) { prefetch_next_region(engine, layout, region, gridDim.y, blockIdx.x); }
  auto const stored = layout.counts[region];
  index_type const count =
    stored < layout.segment_capacity ? index_type{stored} : index_type{layout.segment_capacity};
  auto const* keys        = segments + std::uint64_t{region} * layout.segment_capacity;
  unsigned long long mine = 0;
  for (index_type idx = index_type{blockIdx.x} * BlockSize + threadIdx.x; idx < count;
       idx += index_type{gridDim.x} * BlockSize) {
    auto const key = read_input(keys, idx);
    slot_type unused[1];
    unsigned int const hits = count_matches<ChunkSlots, load_policy::readonly, Ahead>(engine, key, unused);
    mine += (IsOuter && hits == 0) ? 1u : hits;
  }
  accumulate_count(total, mine);
}

/// Pass 2 of a blocked retrieve: CTA x of region y owns a contiguous span of the region's staged
/// keys and runs the ordinary CTA-wide retrieve on it.
template <bool IsOuter,
          int BlockSize,
          int ChunkSlots,
          int Ahead,
          typename OutputProbeIt,
          typename OutputMatchIt,
          typename Counter,
          typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize) void blocked_retrieve_kernel(
  typename Engine::key_type const* segments,
  blocked_layout layout,
  OutputProbeIt output_probe,
  OutputMatchIt output_match,
  Counter* num_rows,
  Engine engine)
{
  std::uint32_t const region = blockIdx.y;
  if (threadIdx.x == 0) { prefetch_next_region(engine, layout, region, gridDim.y, blockIdx.x); }
  auto const stored = layout.counts[region];
  index_type const n =
    stored < layout.segment_capacity ? index_type{stored} : index_type{layout.segment_capacity};
  index_type const rounds =
    cuco::detail::int_div_ceil(cuco::detail::int_div_ceil(n, index_type{BlockSize}), index_type{gridDim.x});
  index_type const span  = rounds * BlockSize;
  index_type const begin = index_type{blockIdx.x} * span;
  if (begin >= n) { return; }
  index_type const count = (n - begin) < span ? (n - begin) : span;
  auto const* keys       = segments + std::uint64_t{region} * layout.segment_capacity + begin;
  global_counter<Counter> counter{num_rows};
  block_retrieve<IsOuter, BlockSize, ChunkSlots, load_policy::readonly, Ahead>(
    engine, keys, count, output_probe, output_match, counter);
}

}  // namespace cuco::b200
