// Device-side index helpers (reference: include/cuco/detail/utility/cuda.cuh:24-86).
#pragma once

#include <cuco/detail/utility/cuda.hpp>

#include <cooperative_groups.h>

#ifndef CUCO_KERNEL
// Kernels are templates living in headers; hidden visibility keeps two shared objects built from
// different versions of these headers from resolving to each other's kernels.
#define CUCO_KERNEL __attribute__((visibility("hidden"))) __global__
#endif

namespace cuco::detail {

__device__ constexpr std::int32_t warp_size() noexcept { return 32; }

__device__ inline index_type global_thread_id() noexcept
{
  return index_type{blockIdx.x} * blockDim.x + threadIdx.x;
}

__device__ inline index_type grid_stride() noexcept { return index_type{gridDim.x} * blockDim.x; }

template <typename Tile>
struct tile_size;

template <unsigned N, typename Parent>
struct tile_size<cooperative_groups::thread_block_tile<N, Parent>> {
  static constexpr unsigned value = N;
};

template <typename Tile>
inline constexpr unsigned tile_size_v = tile_size<Tile>::value;

}  // namespace cuco::detail
