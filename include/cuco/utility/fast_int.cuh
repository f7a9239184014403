// cuco::utility::fast_int — an integer whose value is fixed at construction and that other integers
// can be divided / reduced by without a hardware divide. It is the type behind the dynamic
// `window_extent`, i.e. the `hash % num_windows` of every probe step.
//
// Same public surface as the reference's include/cuco/utility/fast_int.cuh:34-171 (explicit ctor,
// value(), explicit conversion, `lhs / fast`, `lhs % fast`, `fast - x`, `fast / x`), but the
// arithmetic is the branch-free round-up multiply of Granlund & Montgomery ("Division by invariant
// integers using multiplication", 1994, fig. 4.1):
//
//   L  = ceil(log2 d)                     m = floor(2^N (2^L - d) / d) + 1
//   t  = mulhi(m, n)                      q = (t + ((n - t) >> min(L,1))) >> max(L-1, 0)
//
// which is exact for every N-bit unsigned n and every d >= 1 (d == 1 and powers of two included), so
// there are no special cases on the device: one mulhi, one subtract, one add, two shifts.
#pragma once

#include <cuco/detail/__config>

#include <cuda/std/bit>
#include <cuda/std/type_traits>

#include <climits>
#include <cstdint>

namespace cuco::utility {

template <typename T>
struct fast_int {
  static_assert(cuda::std::is_same_v<T, std::int32_t> || cuda::std::is_same_v<T, std::uint32_t> ||
                  cuda::std::is_same_v<T, std::int64_t> || cuda::std::is_same_v<T, std::uint64_t>,
                "fast_int supports 32- and 64-bit integers");

  using value_type = T;

  /// `value` must be positive.
  __host__ __device__ explicit constexpr fast_int(T value) noexcept : value_{value}
  {
    auto const d        = static_cast<unsigned_type>(value);
    int const ceil_log2 = (d <= 1) ? 0 : static_cast<int>(cuda::std::bit_width(d - 1));
    // 2^L - d, computed without overflowing N bits when L == N
    wide_type const pow = wide_type{1} << ceil_log2;
    multiplier_         = static_cast<unsigned_type>(((pow - d) << bits) / d) + 1;
    pre_shift_          = static_cast<std::uint8_t>(ceil_log2 < 1 ? ceil_log2 : 1);
    post_shift_         = static_cast<std::uint8_t>(ceil_log2 > 1 ? ceil_log2 - 1 : 0);
  }

  __host__ __device__ constexpr value_type value() const noexcept { return value_; }
  __host__ __device__ explicit constexpr operator value_type() const noexcept { return value_; }

  /// floor(n / value) for non-negative n.
  __host__ __device__ constexpr value_type quotient(value_type n) const noexcept
  {
    auto const un = static_cast<unsigned_type>(n);
    auto const t  = mulhi(multiplier_, un);
    return static_cast<value_type>((t + ((un - t) >> pre_shift_)) >> post_shift_);
  }

  /// n mod value for non-negative n.
  __host__ __device__ constexpr value_type remainder(value_type n) const noexcept
  {
    return static_cast<value_type>(static_cast<unsigned_type>(n) -
                                   static_cast<unsigned_type>(quotient(n)) *
                                     static_cast<unsigned_type>(value_));
  }

 private:
  using unsigned_type = cuda::std::make_unsigned_t<T>;
  using wide_type = cuda::std::conditional_t<sizeof(T) == 4, std::uint64_t, unsigned __int128>;
  static constexpr int bits = CHAR_BIT * sizeof(T);

  __host__ __device__ static constexpr unsigned_type mulhi(unsigned_type a,
                                                           unsigned_type b) noexcept
  {
#if defined(__CUDA_ARCH__)
    if (!cuda::std::is_constant_evaluated()) {
      if constexpr (sizeof(T) == 4) {
        return __umulhi(a, b);
      } else {
        return __umul64hi(a, b);
      }
    }
#endif
    return static_cast<unsigned_type>((wide_type{a} * wide_type{b}) >> bits);
  }

  value_type value_;
  unsigned_type multiplier_;
  std::uint8_t pre_shift_;
  std::uint8_t post_shift_;

  template <typename Lhs>
  friend __host__ __device__ constexpr value_type operator/(Lhs lhs, fast_int const& rhs) noexcept
  {
    static_assert(cuda::std::is_same_v<Lhs, value_type>,
                  "Left-hand side operand must be of type value_type.");
    return rhs.quotient(lhs);
  }

  template <typename Lhs>
  friend __host__ __device__ constexpr value_type operator%(Lhs lhs, fast_int const& rhs) noexcept
  {
    static_assert(cuda::std::is_same_v<Lhs, value_type>,
                  "Left-hand side operand must be of type value_type.");
    return rhs.remainder(lhs);
  }

  template <typename Rhs>
  friend __host__ __device__ constexpr auto operator-(fast_int const& lhs, Rhs rhs) noexcept
  {
    return lhs.value() - rhs;
  }

  template <typename Rhs>
  friend __host__ __device__ constexpr auto operator/(fast_int const& lhs, Rhs rhs) noexcept
  {
    return lhs.value() / rhs;
  }
};

template <typename T>
fast_int(T) -> fast_int<T>;

}  // namespace cuco::utility
