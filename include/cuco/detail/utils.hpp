// Iterator helpers used by the host bulk API (reference: include/cuco/detail/utils.hpp:27-71).
#pragma once

#include <cuco/detail/utility/cuda.hpp>

#include <cuda/std/iterator>
#include <cuda/std/type_traits>

#include <iterator>

namespace cuco::detail {

/// `last - first` as a 64-bit count; bulk inputs must be random access.
template <typename It>
__host__ __device__ constexpr index_type distance(It first, It last)
{
  using category = typename cuda::std::iterator_traits<It>::iterator_category;
  static_assert(cuda::std::is_base_of_v<cuda::std::random_access_iterator_tag, category>,
                "Input iterator should be a random access iterator.");
  return static_cast<index_type>(cuda::std::distance(first, last));
}

/// constexpr binary search: first position in [first, last) whose element is not less than value.
template <typename It, typename T>
constexpr It lower_bound(It first, It last, T const& value)
{
  auto len = std::distance(first, last);
  while (len > 0) {
    auto const half = len / 2;
    It mid          = first;
    std::advance(mid, half);
    if (static_cast<T>(*mid) < value) {
      first = ++mid;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return first;
}

}  // namespace cuco::detail
