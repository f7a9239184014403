// Whole-table passes: retrieve_all, for_each, rehash.
//
// These are the rows SURVEY.md §8(f) ranks "next" after the probe path; the reference implements
// them with cub::DeviceSelect::If in INT32_MAX chunks (open_addressing_impl.cuh:726-785),
// cub::DeviceFor (:796-812) and a shared-memory staged rehash kernel (open_addressing/kernels.cuh:
// 670-715). Here each is one streaming kernel over the slot array: sector-wide loads, a
// warp-aggregated atomic to reserve output space (order of the output is unspecified by contract),
// 64-bit indexing throughout so there is no chunking.
#pragma once

#include <cuco/b200/bulk_kernels.cuh>
#include <cuco/b200/match_kernels.cuh>
#include <cuco/b200/probe_engine.cuh>
#include <cuco/detail/error.hpp>
#include <cuco/detail/utility/cuda.cuh>

#include <cuda/atomic>
#include <cuda/stream_ref>

#include <cooperative_groups.h>

namespace cuco::b200 {

template <typename Engine>
struct filled_slot {
  typename Engine::key_type empty;
  typename Engine::key_type erased;

  template <typename Slot>
  __device__ bool operator()(Slot const& slot) const noexcept
  {
    auto const& k = Engine::key_of(slot);
    return !(same_bits(k, empty) || same_bits(k, erased));
  }
};

/// Appends every filled slot to the output through `write(position, slot)`.
/// One tile = BlockSize x 4 slots: four independent slot loads per thread, per-thread fill count,
/// CTA-wide exclusive scan, ONE atomic on the output cursor per tile (a warp-level ballot + atomic per
/// 32 slots serialises 3 M atomics on one address for a 100 M-slot table: 4.7 ms against cuco's 3.7),
/// then every thread writes its survivors at its own offset - the rows of a tile form one run.
template <int BlockSize, typename Engine, typename Write>
CUCO_KERNEL __launch_bounds__(BlockSize) void compact_kernel(Engine engine,
                                                             unsigned long long* cursor,
                                                             Write write)
{
  using slot_type      = typename Engine::value_type;
  constexpr int items  = 4;
  constexpr index_type tile = index_type{BlockSize} * items;
  auto const* table    = engine.slots();
  auto const n         = static_cast<index_type>(engine.capacity());
  auto const filled    = filled_slot<Engine>{engine.empty_key_sentinel(), engine.erased_key_sentinel()};
  __shared__ unsigned int warp_totals[BlockSize / 32];
  __shared__ unsigned long long tile_base;

  for (index_type base = index_type{blockIdx.x} * tile; base < n; base += index_type{gridDim.x} * tile) {
    slot_type slot[items];
    unsigned keep = 0;
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const i = base + index_type{j} * BlockSize + threadIdx.x;
      if (i < n) { slot[j] = load_streaming(table + i); }
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const i = base + index_type{j} * BlockSize + threadIdx.x;
      if (i < n && filled(slot[j])) { keep |= 1u << j; }
    }
    unsigned int offset, total;
    block_exclusive_scan<BlockSize>(static_cast<unsigned int>(__popc(keep)), warp_totals, offset, total);
    if (total == 0) { continue; }  // uniform across the CTA
    if (threadIdx.x == 0) {
      cuda::atomic_ref<unsigned long long, cuda::thread_scope_device> ref{*cursor};
      tile_base = ref.fetch_add(total, cuda::memory_order_relaxed);
    }
    __syncthreads();
    unsigned long long where = tile_base + offset;
#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (keep & (1u << j)) { write(where++, slot[j]); }
    }
    __syncthreads();  // tile_base is rewritten next round
  }
}

template <typename Engine, typename Write>
inline unsigned long long compact_filled(Engine const& engine, Write write, cuda::stream_ref stream)
{
  constexpr int block = 256;
  unsigned long long* cursor{};
  // from the library's own pool, which keeps freed blocks cached: the device's DEFAULT pool gives memory
  // back to the driver at every synchronisation (release threshold 0), and this function synchronises -
  // the next call then pays a physical allocation (measured: retrieve_all between 2 and 120 ms)
  auto const pool = device_scratch_pool();
  CUCO_EXPECTS(pool != nullptr, "no stream-ordered memory pool on this device");
  CUCO_CUDA_TRY(cudaMallocFromPoolAsync(
    reinterpret_cast<void**>(&cursor), sizeof(unsigned long long), pool, stream.get()));
  CUCO_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), stream.get()));
  auto const n = static_cast<index_type>(engine.capacity());
  if (n > 0) {
    auto const kernel = compact_kernel<block, Engine, Write>;
    auto const grid =
      persistent_grid(kernel, block, cuco::detail::int_div_ceil(n, index_type{block} * 4));
    kernel<<<grid, block, 0, stream.get()>>>(engine, cursor, write);
  }
  unsigned long long count = 0;
  CUCO_CUDA_TRY(
    cudaMemcpyAsync(&count, cursor, sizeof(count), cudaMemcpyDeviceToHost, stream.get()));
  CUCO_CUDA_TRY(cudaFreeAsync(cursor, stream.get()));
  stream.wait();
  return count;
}

template <typename KeyOut, typename ValueOut>
struct write_pair_columns {
  KeyOut keys;
  ValueOut values;
  template <typename Slot>
  __device__ void operator()(unsigned long long pos, Slot const& slot) const
  {
    *(keys + pos)   = slot.first;
    *(values + pos) = slot.second;
  }
};

template <typename Out>
struct write_elements {
  Out out;
  template <typename Slot>
  __device__ void operator()(unsigned long long pos, Slot const& slot) const
  {
    *(out + pos) = slot;
  }
};

/// static_map::retrieve_all — returns the number of pairs written.
template <typename Engine, typename KeyOut, typename ValueOut>
inline unsigned long long retrieve_all_pairs(Engine const& engine,
                                             KeyOut keys_out,
                                             ValueOut values_out,
                                             cuda::stream_ref stream)
{
  auto k = unwrap(keys_out);
  auto v = unwrap(values_out);
  return compact_filled(engine, write_pair_columns<decltype(k), decltype(v)>{k, v}, stream);
}

/// static_set::retrieve_all — returns the number of keys written.
template <typename Engine, typename Out>
inline unsigned long long retrieve_all_elements(Engine const& engine, Out out, cuda::stream_ref stream)
{
  auto o = unwrap(out);
  return compact_filled(engine, write_elements<decltype(o)>{o}, stream);
}

template <int BlockSize, typename Engine, typename Callback>
CUCO_KERNEL __launch_bounds__(BlockSize) void for_each_filled_kernel(Engine engine, Callback callback)
{
  auto const* table = engine.slots();
  auto const n      = static_cast<index_type>(engine.capacity());
  auto const filled = filled_slot<Engine>{engine.empty_key_sentinel(), engine.erased_key_sentinel()};
  for (index_type i = cuco::detail::global_thread_id(); i < n; i += cuco::detail::grid_stride()) {
    auto const slot = table[i];
    if (filled(slot)) { callback(slot); }
  }
}

template <typename Engine, typename Callback>
inline void for_each_filled_async(Engine const& engine, Callback callback, cuda::stream_ref stream)
{
  constexpr int block = 256;
  auto const n        = static_cast<index_type>(engine.capacity());
  if (n == 0) { return; }
  auto const kernel = for_each_filled_kernel<block, Engine, Callback>;
  auto const grid   = persistent_grid(kernel, block, cuco::detail::int_div_ceil(n, index_type{block}));
  kernel<<<grid, block, 0, stream.get()>>>(engine, callback);
}

template <int BlockSize, typename Engine, typename InputIt, typename Callback>
CUCO_KERNEL __launch_bounds__(BlockSize) void for_each_key_kernel(Engine engine,
                                                                  InputIt first,
                                                                  index_type n,
                                                                  Callback callback)
{
  for (index_type i = cuco::detail::global_thread_id(); i < n; i += cuco::detail::grid_stride()) {
    engine.scalar_for_each(read_input(first, i), callback);
  }
}

template <typename Engine, typename InputIt, typename Callback>
inline void for_each_key_async(
  Engine const& engine, InputIt first, InputIt last, Callback callback, cuda::stream_ref stream)
{
  constexpr int block = 256;
  auto const n        = cuco::detail::distance(first, last);
  if (n == 0) { return; }
  auto in           = unwrap(first);
  auto const kernel = for_each_key_kernel<block, Engine, decltype(in), Callback>;
  auto const grid   = persistent_grid(kernel, block, cuco::detail::int_div_ceil(n, index_type{block}));
  kernel<<<grid, block, 0, stream.get()>>>(engine, in, n, callback);
}

/// Moves every entry of `table` into fresh storage of `extent` windows (drops tombstones). The old
/// slot array doubles as input range and stencil of an ordinary bulk insert_if.
template <typename TableEngine, typename Extent, typename InsertRef>
inline void rehash_into(TableEngine& table, Extent extent, InsertRef, cuda::stream_ref stream)
{
  using engine_t   = typename TableEngine::engine_type;
  using slot_t     = typename TableEngine::value_type;
  auto const old_engine = table.make_engine();
  auto const old_slots  = static_cast<index_type>(old_engine.capacity());
  auto old_storage      = table.exchange_storage(extent, stream);
  if (old_slots == 0) { return; }
  auto const* first = reinterpret_cast<slot_t const*>(old_storage.data());
  auto const filled =
    filled_slot<engine_t>{old_engine.empty_key_sentinel(), old_engine.erased_key_sentinel()};
  table.insert_slots_if(first, first + old_slots, first, filled, stream);
  stream.wait();  // old storage is released when this function returns
}

}  // namespace cuco::b200
