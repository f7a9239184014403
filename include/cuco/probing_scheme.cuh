// Probing schemes: the rule that turns a key into the sequence of windows an operation visits.
//
// Public surface and arithmetic follow the reference (include/cuco/probing_scheme.cuh:39-204,
// detail/probing_scheme/probing_scheme_impl.inl:30-200, detail/utils.cuh:89-142), because the
// sequence is observable through `scheme(key, extent)` / `scheme(tile, key, extent)` and must be the
// same no matter which code path (bulk kernel, per-thread ref op, per-tile ref op) touches a table:
//
//   linear_probing<CG, H>      start = H(key) mod N (+ tile rank)        step = CG
//   double_hashing<CG, H1, H2> start = H1(key) mod N (+ tile rank)       step = (H2(key) mod (N/CG - 1) + 1) * CG
//
// with N the number of windows (a prime times CG). One probe step therefore covers the CG
// consecutive windows starting at `start + k*step (mod N)`: a *bucket* of CG * window_size slots.
// The sm_100a bulk kernels scan that bucket from a single thread with vector loads and early exit
// (cuco/b200/probe_engine.cuh); tiles of CG threads scan it cooperatively with one ballot. Both get
// their arithmetic from `home_hash` / `step_hash` below so they cannot drift apart.
#pragma once

#include <cuco/detail/__config>
#include <cuco/hash_functions.cuh>
#include <cuco/pair.cuh>

#include <cuda/std/array>
#include <cuda/std/limits>
#include <cuda/std/tuple>
#include <cuda/std/type_traits>

#include <cooperative_groups.h>

#include <cstdint>
#include <cstring>

namespace cuco {
namespace detail {

/// Tag base so containers can check "is a cuco probing scheme" and read cg_size.
template <std::int32_t CGSize>
class probing_scheme_base {
 public:
  static constexpr std::int32_t cg_size = CGSize;
};

/// Maps any hash result onto a non-negative value of the container's size type.
/// 128-bit hashes (arrays of two 64-bit words) are truncated to their low bits; signed size types
/// take the absolute value.
template <typename SizeType, typename HashType>
__host__ __device__ constexpr SizeType sanitize_hash(HashType hash) noexcept
{
  SizeType narrowed{};
  if constexpr (cuda::std::is_same_v<HashType, cuda::std::array<std::uint64_t, 2>>) {
    unsigned __int128 wide{};
    memcpy(&wide, &hash, sizeof(wide));
    narrowed = static_cast<SizeType>(wide);
  } else {
    narrowed = static_cast<SizeType>(hash);
  }
  if constexpr (cuda::std::is_signed_v<SizeType>) {
    return narrowed < 0 ? static_cast<SizeType>(-narrowed) : narrowed;
  } else {
    return narrowed;
  }
}

/// `hash + rank` without overflowing SizeType (wraps past max the way the reference does).
template <typename SizeType>
__host__ __device__ constexpr SizeType add_rank(SizeType base, SizeType rank) noexcept
{
  constexpr auto top = cuda::std::numeric_limits<SizeType>::max();
  return (base > top - rank) ? static_cast<SizeType>(rank - (top - base))
                             : static_cast<SizeType>(base + rank);
}

template <typename SizeType, typename CG, typename HashType>
__device__ constexpr SizeType sanitize_hash(CG const& group, HashType hash) noexcept
{
  return add_rank<SizeType>(sanitize_hash<SizeType>(hash),
                            static_cast<SizeType>(group.thread_rank()));
}

/// Forward iterator over window indices: i, i + step, i + 2 step, ... (mod upper_bound).
template <typename Extent>
class probing_iterator {
 public:
  using extent_type = Extent;
  using size_type   = typename extent_type::value_type;

  __host__ __device__ constexpr probing_iterator(size_type start,
                                                 size_type step_size,
                                                 extent_type upper_bound) noexcept
    : index_{start}, step_{step_size}, bound_{upper_bound}
  {
  }

  __host__ __device__ constexpr size_type operator*() const noexcept { return index_; }

  __host__ __device__ constexpr probing_iterator& operator++() noexcept
  {
    index_ = (index_ + step_) % bound_;
    return *this;
  }

  __host__ __device__ constexpr probing_iterator operator++(std::int32_t) noexcept
  {
    auto const before = *this;
    ++(*this);
    return before;
  }

 private:
  size_type index_;
  size_type step_;
  extent_type bound_;
};

}  // namespace detail

template <std::int32_t CGSize, typename Hash>
class linear_probing : private detail::probing_scheme_base<CGSize> {
 public:
  using detail::probing_scheme_base<CGSize>::cg_size;
  using hasher = Hash;

  static constexpr bool is_double_hashing_scheme = false;  // engine hook

  __host__ __device__ constexpr linear_probing(Hash const& hash = {}) : hash_{hash} {}

  template <typename NewHash>
  [[nodiscard]] __host__ __device__ constexpr auto rebind_hash_function(
    NewHash const& hash) const noexcept
  {
    return linear_probing<cg_size, NewHash>{hash};
  }

  /// Per-thread sequence (cg_size == 1 semantics: consecutive windows).
  template <typename ProbeKey, typename Extent>
  __host__ __device__ constexpr auto operator()(ProbeKey const& probe_key,
                                                Extent upper_bound) const noexcept
  {
    using size_type = typename Extent::value_type;
    return detail::probing_iterator<Extent>{
      detail::sanitize_hash<size_type>(hash_(probe_key)) % upper_bound, 1, upper_bound};
  }

  /// Per-tile sequence: rank r starts r windows further and every rank strides by cg_size.
  template <typename ProbeKey, typename Extent>
  __host__ __device__ constexpr auto operator()(
    cooperative_groups::thread_block_tile<cg_size> const& g,
    ProbeKey const& probe_key,
    Extent upper_bound) const noexcept
  {
    using size_type = typename Extent::value_type;
    return detail::probing_iterator<Extent>{
      detail::sanitize_hash<size_type>(g, hash_(probe_key)) % upper_bound, cg_size, upper_bound};
  }

  __host__ __device__ constexpr hasher hash_function() const noexcept { return hash_; }

  // engine hooks -------------------------------------------------------------------------------
  template <typename SizeType, typename ProbeKey>
  __host__ __device__ constexpr SizeType home_hash(ProbeKey const& key) const noexcept
  {
    return detail::sanitize_hash<SizeType>(hash_(key));
  }

 private:
  Hash hash_;
};

template <std::int32_t CGSize, typename Hash1, typename Hash2 = Hash1>
class double_hashing : private detail::probing_scheme_base<CGSize> {
 public:
  using detail::probing_scheme_base<CGSize>::cg_size;
  using hasher = cuda::std::tuple<Hash1, Hash2>;

  static constexpr bool is_double_hashing_scheme = true;  // engine hook

  /// The second hash defaults to the same family seeded with 1 so the two are independent.
  __host__ __device__ constexpr double_hashing(Hash1 const& hash1 = {}, Hash2 const& hash2 = {1})
    : hash1_{hash1}, hash2_{hash2}
  {
  }

  __host__ __device__ constexpr double_hashing(cuda::std::tuple<Hash1, Hash2> const& hash)
    : hash1_{cuda::std::get<0>(hash)}, hash2_{cuda::std::get<1>(hash)}
  {
  }

  template <typename NewHash,
            typename Enable = cuda::std::enable_if_t<cuco::is_tuple_like<NewHash>::value>>
  [[nodiscard]] __host__ __device__ constexpr auto rebind_hash_function(NewHash const& hash) const
  {
    auto const first  = cuda::std::get<0>(hash);
    auto const second = cuda::std::get<1>(hash);
    return double_hashing<cg_size,
                          cuda::std::decay_t<decltype(first)>,
                          cuda::std::decay_t<decltype(second)>>{first, second};
  }

  template <typename ProbeKey, typename Extent>
  __host__ __device__ constexpr auto operator()(ProbeKey const& probe_key,
                                                Extent upper_bound) const noexcept
  {
    using size_type = typename Extent::value_type;
    return detail::probing_iterator<Extent>{
      detail::sanitize_hash<size_type>(hash1_(probe_key)) % upper_bound,
      static_cast<size_type>(
        detail::sanitize_hash<size_type>(hash2_(probe_key)) % (upper_bound - 1) + 1),
      upper_bound};
  }

  template <typename ProbeKey, typename Extent>
  __host__ __device__ constexpr auto operator()(
    cooperative_groups::thread_block_tile<cg_size> const& g,
    ProbeKey const& probe_key,
    Extent upper_bound) const noexcept
  {
    using size_type = typename Extent::value_type;
    return detail::probing_iterator<Extent>{
      detail::sanitize_hash<size_type>(g, hash1_(probe_key)) % upper_bound,
      static_cast<size_type>(
        (detail::sanitize_hash<size_type>(hash2_(probe_key)) % (upper_bound / cg_size - 1) + 1) *
        cg_size),
      upper_bound};
  }

  __host__ __device__ constexpr hasher hash_function() const noexcept { return {hash1_, hash2_}; }

  // engine hooks -------------------------------------------------------------------------------
  template <typename SizeType, typename ProbeKey>
  __host__ __device__ constexpr SizeType home_hash(ProbeKey const& key) const noexcept
  {
    return detail::sanitize_hash<SizeType>(hash1_(key));
  }

  template <typename SizeType, typename ProbeKey>
  __host__ __device__ constexpr SizeType step_hash(ProbeKey const& key) const noexcept
  {
    return detail::sanitize_hash<SizeType>(hash2_(key));
  }

 private:
  Hash1 hash1_;
  Hash2 hash2_;
};

template <typename T>
struct is_double_hashing : cuda::std::false_type {};

template <std::int32_t CGSize, typename Hash1, typename Hash2>
struct is_double_hashing<double_hashing<CGSize, Hash1, Hash2>> : cuda::std::true_type {};

}  // namespace cuco
