// cuco::pair — the slot type of static_map (reference: include/cuco/pair.cuh:40-111 and
// detail/pair/*). A dedicated pair type exists because the slot must be *over-aligned*: a
// pair<int64,int64> is 16-byte aligned so that one slot can be read with a single 128-bit load and
// claimed with a single 128-bit compare-and-swap (`atom.cas.b128` on sm_100a), and a
// pair<int32,int32> is 8-byte aligned for a packed 64-bit CAS. thrust/std pairs only guarantee the
// alignment of their widest member.
#pragma once

#include <cuda/std/bit>
#include <cuda/std/tuple>
#include <cuda/std/type_traits>
#include <cuda/std/utility>
#include <thrust/device_reference.h>
#include <thrust/pair.h>
#include <thrust/tuple.h>

#include <cstddef>
#include <cstdint>
#include <tuple>
#include <type_traits>
#include <utility>

namespace cuco {
namespace detail {

/// min(16, next power of two >= sizeof(First) + sizeof(Second))
template <typename First, typename Second>
__host__ __device__ constexpr std::size_t pair_alignment()
{
  constexpr std::size_t pow2 = cuda::std::bit_ceil(sizeof(First) + sizeof(Second));
  return pow2 < 16 ? pow2 : 16;
}

/// Unsigned integer with the same size as an object, or void when there is none (<= 8 bytes only).
template <std::size_t Bytes>
struct packed {
  using type = void;
};
template <>
struct packed<4> {
  using type = std::uint32_t;
};
template <>
struct packed<8> {
  using type = std::uint64_t;
};
template <typename Slot>
using packed_t = typename packed<sizeof(Slot)>::type;

/// A slot is "packable" when a single 32/64-bit CAS covers all of its bits and all bits matter.
template <typename Slot>
__host__ __device__ constexpr bool is_packable()
{
  return !std::is_void_v<packed_t<Slot>> && std::has_unique_object_representations_v<Slot>;
}

/// View a packable slot as its integer image and back.
template <typename Slot>
union pair_converter {
  using packed_type = packed_t<Slot>;
  packed_type packed;
  Slot pair;

  template <typename T>
  __device__ pair_converter(T&& p) : pair{p}
  {
  }
  __device__ pair_converter(packed_type p) : packed{p} {}
};

// --- "looks like a 2-tuple" detection, for std:: and cuda::std:: get<> families -------------------
template <typename T, typename = void>
struct is_std_pair_like : cuda::std::false_type {};

template <typename T>
struct is_std_pair_like<T,
                        cuda::std::void_t<decltype(std::get<0>(cuda::std::declval<T>())),
                                          decltype(std::get<1>(cuda::std::declval<T>()))>>
  : cuda::std::bool_constant<std::tuple_size<T>::value == 2> {};

template <typename T, typename = void>
struct is_cuda_std_pair_like_impl : cuda::std::false_type {};

template <typename T>
struct is_cuda_std_pair_like_impl<
  T,
  cuda::std::void_t<decltype(cuda::std::get<0>(cuda::std::declval<T>())),
                    decltype(cuda::std::get<1>(cuda::std::declval<T>())),
                    decltype(cuda::std::tuple_size<T>::value)>>
  : cuda::std::bool_constant<cuda::std::tuple_size<T>::value == 2> {};

/// thrust::device_reference<X> is looked through so that `*device_vector_iterator` qualifies.
template <typename T>
struct is_cuda_std_pair_like
  : is_cuda_std_pair_like_impl<cuda::std::remove_reference_t<decltype(thrust::raw_reference_cast(
      cuda::std::declval<T>()))>> {};

}  // namespace detail

template <typename First, typename Second>
struct alignas(detail::pair_alignment<First, Second>()) pair {
  using first_type  = First;
  using second_type = Second;

  pair()                       = default;
  ~pair()                      = default;
  pair(pair const&)            = default;
  pair(pair&&)                 = default;
  pair& operator=(pair const&) = default;
  pair& operator=(pair&&)      = default;

  __host__ __device__ constexpr pair(First const& f, Second const& s) : first{f}, second{s} {}

  /// Converting copy from another cuco::pair.
  template <typename F, typename S>
  __host__ __device__ constexpr pair(pair<F, S> const& p) : first{p.first}, second{p.second}
  {
  }

  /// From anything std::get<0/1> works on (std::pair, std::tuple of two).
  template <typename T, std::enable_if_t<detail::is_std_pair_like<T>::value>* = nullptr>
  __host__ __device__ constexpr pair(T const& p)
    : pair{std::get<0>(thrust::raw_reference_cast(p)), std::get<1>(thrust::raw_reference_cast(p))}
  {
  }

  /// From anything cuda::std::get<0/1> works on (thrust::pair/tuple, cuda::std::pair/tuple).
  template <typename T, std::enable_if_t<detail::is_cuda_std_pair_like<T>::value>* = nullptr>
  __host__ __device__ constexpr pair(T const& p)
    : pair{cuda::std::get<0>(thrust::raw_reference_cast(p)),
           cuda::std::get<1>(thrust::raw_reference_cast(p))}
  {
  }

  First first;
  Second second;
};

template <typename F, typename S>
pair(F, S) -> pair<F, S>;

template <typename F, typename S>
__host__ __device__ constexpr pair<std::decay_t<F>, std::decay_t<S>> make_pair(F&& f,
                                                                               S&& s) noexcept
{
  return pair<std::decay_t<F>, std::decay_t<S>>(std::forward<F>(f), std::forward<S>(s));
}

template <typename T1, typename T2, typename U1, typename U2>
__host__ __device__ constexpr bool operator==(pair<T1, T2> const& lhs,
                                              pair<U1, U2> const& rhs) noexcept
{
  return lhs.first == rhs.first && lhs.second == rhs.second;
}

template <typename T>
struct is_cuco_pair : cuda::std::false_type {};
template <typename F, typename S>
struct is_cuco_pair<pair<F, S>> : cuda::std::true_type {};

/// cuco::pair, or anything the std / cuda::std 2-tuple protocols accept.
template <typename T>
struct is_tuple_like : cuda::std::disjunction<is_cuco_pair<T>,
                                              detail::is_std_pair_like<T>,
                                              detail::is_cuda_std_pair_like<T>> {};

}  // namespace cuco

// Tuple protocol so that cuda::std::get<I>(slot) and thrust zip machinery work on cuco::pair.
// (std::tuple_size is deliberately NOT specialised: host structured bindings then bind the public
// members first/second directly.)
namespace cuda::std {

template <typename F, typename S>
struct tuple_size<cuco::pair<F, S>> : integral_constant<size_t, 2> {};
template <typename F, typename S>
struct tuple_size<const cuco::pair<F, S>> : tuple_size<cuco::pair<F, S>> {};
template <typename F, typename S>
struct tuple_size<volatile cuco::pair<F, S>> : tuple_size<cuco::pair<F, S>> {};
template <typename F, typename S>
struct tuple_size<const volatile cuco::pair<F, S>> : tuple_size<cuco::pair<F, S>> {};

template <size_t I, typename F, typename S>
struct tuple_element<I, cuco::pair<F, S>> {
  static_assert(I < 2, "cuco::pair has two elements");
  using type = conditional_t<I == 0, F, S>;
};
template <size_t I, typename F, typename S>
struct tuple_element<I, const cuco::pair<F, S>> : tuple_element<I, cuco::pair<F, S>> {};
template <size_t I, typename F, typename S>
struct tuple_element<I, volatile cuco::pair<F, S>> : tuple_element<I, cuco::pair<F, S>> {};
template <size_t I, typename F, typename S>
struct tuple_element<I, const volatile cuco::pair<F, S>> : tuple_element<I, cuco::pair<F, S>> {};

template <size_t I, typename F, typename S>
__host__ __device__ constexpr auto& get(cuco::pair<F, S>& p) noexcept
{
  if constexpr (I == 0) {
    return p.first;
  } else {
    return p.second;
  }
}

template <size_t I, typename F, typename S>
__host__ __device__ constexpr auto const& get(cuco::pair<F, S> const& p) noexcept
{
  if constexpr (I == 0) {
    return p.first;
  } else {
    return p.second;
  }
}

template <size_t I, typename F, typename S>
__host__ __device__ constexpr auto&& get(cuco::pair<F, S>&& p) noexcept
{
  if constexpr (I == 0) {
    return ::cuda::std::move(p.first);
  } else {
    return ::cuda::std::move(p.second);
  }
}

template <size_t I, typename F, typename S>
__host__ __device__ constexpr auto const&& get(cuco::pair<F, S> const&& p) noexcept
{
  if constexpr (I == 0) {
    return ::cuda::std::move(p.first);
  } else {
    return ::cuda::std::move(p.second);
  }
}

}  // namespace cuda::std
