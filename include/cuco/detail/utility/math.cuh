#pragma once

#include <type_traits>

namespace cuco::detail {

/// ceil(a / b) for non-negative integers.
template <typename A, typename B>
__host__ __device__ constexpr A int_div_ceil(A a, B b) noexcept
{
  static_assert(std::is_integral_v<A> && std::is_integral_v<B>);
  return (a + b - 1) / b;
}

}  // namespace cuco::detail
