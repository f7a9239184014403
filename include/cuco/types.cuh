// Strong types for sentinel arguments (reference: include/cuco/types.cuh:35-45), so that
// `static_map{n, empty_key{-1}, empty_value{-1}}` cannot silently swap the two and CTAD can deduce
// Key and T from them.
#pragma once

namespace cuco {
namespace detail {

/// One value of type T, tagged by `Tag` so that differently-purposed values are distinct types.
template <typename T, typename Tag>
struct tagged_value {
  using value_type = T;

  __host__ __device__ explicit constexpr tagged_value(T v) : value{v} {}
  __host__ __device__ constexpr operator T() const noexcept { return value; }

  T value;
};

}  // namespace detail

/// Key that marks a never-used slot.
template <typename T>
struct empty_key : detail::tagged_value<T, struct empty_key_role> {
  __host__ __device__ explicit constexpr empty_key(T v)
    : detail::tagged_value<T, struct empty_key_role>{v}
  {
  }
};

/// Payload stored next to the empty key.
template <typename T>
struct empty_value : detail::tagged_value<T, struct empty_value_role> {
  __host__ __device__ explicit constexpr empty_value(T v)
    : detail::tagged_value<T, struct empty_value_role>{v}
  {
  }
};

/// Key that marks a slot whose entry was erased (tombstone).
template <typename T>
struct erased_key : detail::tagged_value<T, struct erased_key_role> {
  __host__ __device__ explicit constexpr erased_key(T v)
    : detail::tagged_value<T, struct erased_key_role>{v}
  {
  }
};

// deduction guides: empty_key{-1} -> empty_key<int>
template <typename T>
empty_key(T) -> empty_key<T>;
template <typename T>
empty_value(T) -> empty_value<T>;
template <typename T>
erased_key(T) -> erased_key<T>;

}  // namespace cuco
