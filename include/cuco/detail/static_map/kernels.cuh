// Block-local pre-aggregation for insert_or_apply (reference:
// include/cuco/detail/static_map/kernels.cuh:171-264, `insert_or_apply_shmem`, dispatched by
// detail/static_map/helpers.cuh:50-114 when cg_size == 1).
//
// The bulk `insert_or_apply_async` of this tree does not need it - batches that are dense in a
// table beyond L2 are regrouped by table region instead (b200/stream_kernels.cuh), and tables that
// fit in L2 take the direct kernel - but the kernel is part of what the reference's users and its own
// tests (tests/static_map/insert_or_apply_test.cu:118) launch by name, so it exists here with the
// same template parameters and arguments. It is also the building block of the per-GPU
// pre-aggregation of partitioned group-by batches (cucollections_b200/partitioned.py).
//
// Every CTA owns a table of `window_extent` windows in shared memory (block scope, same key
// predicate and probing scheme as the global table). Elements are folded into it round by round
// - one element per thread and round, the number of keys that were new to the CTA counted by the
// round's barrier itself (`__syncthreads_count`) - until the CTA has seen more than `BlockSize`
// distinct keys; then the shared table is folded into the global one, and whatever input is left
// goes to the global table directly.
#pragma once

#include <cuco/b200/probe_engine.cuh>
#include <cuco/detail/utility/cuda.cuh>
#include <cuco/pair.cuh>
#include <cuco/types.cuh>

#include <cooperative_groups.h>

#include <cstdint>
#include <iterator>

namespace cuco::static_map_ns::detail {

template <bool HasInit,
          std::int32_t CGSize,
          std::int32_t BlockSize,
          class SharedMapRefType,
          class InputIt,
          class Init,
          class Op,
          class Ref>
CUCO_KERNEL __launch_bounds__(BlockSize) void insert_or_apply_shmem(
  InputIt first,
  cuco::detail::index_type n,
  [[maybe_unused]] Init init,
  Op op,
  Ref ref,
  typename SharedMapRefType::extent_type window_extent)
{
  static_assert(CGSize == 1, "the shared-memory pre-aggregation is a one-thread-per-key kernel");
  namespace cg      = cooperative_groups;
  using key_type    = typename Ref::key_type;
  using mapped_type = typename Ref::mapped_type;
  using input_type  = typename std::iterator_traits<InputIt>::value_type;
  using index_type  = cuco::detail::index_type;

  auto const block = cg::this_thread_block();

  __shared__ typename SharedMapRefType::window_type block_windows[window_extent.value()];
  auto const block_storage = typename SharedMapRefType::storage_ref_type(window_extent, block_windows);
  auto block_table = SharedMapRefType{cuco::empty_key<key_type>{ref.empty_key_sentinel()},
                                      cuco::empty_value<mapped_type>{ref.empty_value_sentinel()},
                                      ref.key_eq(),
                                      ref.probing_scheme(),
                                      {},
                                      block_storage}
                       .rebind_operators(cuco::op::insert_or_apply);
  block_table.initialize(block);  // fills with the sentinels and synchronises the CTA

  auto fold = [&](auto& table, auto const& element) {
    if constexpr (HasInit) {
      return table.insert_or_apply(element, init, op);
    } else {
      return table.insert_or_apply(element, op);
    }
  };

  index_type const stride = cuco::detail::grid_stride();
  index_type idx          = cuco::detail::global_thread_id();
  index_type const lead   = idx - threadIdx.x;  // the CTA's first element of the current round
  int distinct_in_block   = 0;                   // same value in every thread of the CTA
  index_type round        = 0;

  // aggregate while the CTA-local table has room: at most BlockSize new keys per round
  while (lead + round * stride < n && distinct_in_block <= BlockSize) {
    index_type const mine = idx + round * stride;
    bool was_new          = false;
    if (mine < n) {
      input_type const element = *(first + mine);
      was_new                  = fold(block_table, element);
    }
    distinct_in_block += __syncthreads_count(was_new);
    ++round;
  }

  // fold the CTA's partial results into the global table
  auto const windows = block_storage.num_windows();
  for (auto w = static_cast<decltype(windows)>(threadIdx.x); w < windows; w += BlockSize) {
    auto const slot = block_storage[w][0];
    if (!cuco::detail::bitwise_compare(slot.first, ref.empty_key_sentinel())) { fold(ref, slot); }
  }

  // the CTA saw too many distinct keys: the rest of its share bypasses shared memory
  for (index_type mine = idx + round * stride; mine < n; mine += stride) {
    input_type const element = *(first + mine);
    fold(ref, element);
  }
}

}  // namespace cuco::static_map_ns::detail
