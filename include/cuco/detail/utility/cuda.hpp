// Host-side launch arithmetic shared by the bulk launchers
// (reference: include/cuco/detail/utility/cuda.hpp:24-77). The reference launches one key per
// cooperative group in blocks of 128; the sm_100a kernels use their own tiling
// (cuco/b200/bulk_launch.cuh) but `grid_size` keeps its meaning for generic callers and tests.
#pragma once

#include <cuco/detail/error.hpp>
#include <cuco/detail/utility/math.cuh>

#include <cuda_runtime_api.h>

#include <cstdint>

namespace cuco::detail {

using index_type = std::int64_t;  ///< 64-bit element index used by every kernel

constexpr std::int32_t default_block_size() noexcept { return 128; }
constexpr std::int32_t default_stride() noexcept { return 1; }

/// Blocks needed so that `num` items, `cg_size` threads each, `stride` items per thread, are covered.
constexpr auto grid_size(index_type num,
                         std::int32_t cg_size    = 1,
                         std::int32_t stride     = default_stride(),
                         std::int32_t block_size = default_block_size()) noexcept
{
  return int_div_ceil(cg_size * num, static_cast<index_type>(stride) * block_size);
}

/// Number of SMs of the current device (148 on B200); cached per device after the first query.
inline std::int32_t multiprocessor_count()
{
  constexpr int max_devices = 64;
  static std::int32_t cache[max_devices] = {};
  int dev = 0;
  CUCO_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < max_devices && cache[dev] > 0) { return cache[dev]; }
  std::int32_t sms = 0;
  CUCO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dev >= 0 && dev < max_devices) { cache[dev] = sms; }
  return sms;
}

/// Largest grid of `kernel` that is fully resident at once (blocks/SM x SM count).
template <typename Kernel>
auto max_occupancy_grid_size(std::int32_t block_size, Kernel kernel, std::size_t dynamic_smem = 0)
{
  std::int32_t per_sm = 0;
  CUCO_CUDA_TRY(
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_size, dynamic_smem));
  return per_sm * multiprocessor_count();
}

}  // namespace cuco::detail
