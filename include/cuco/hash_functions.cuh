// Hash functors of the drop-in surface (reference: include/cuco/hash_functions.cuh:34-105 and
// detail/hash_functions/*). All are stateless apart from a seed, callable on host and device, and
// expose `argument_type`, `result_type`, `operator()(Key const&)` and
// `compute_hash(bytes, size)` like the reference so that probing schemes and user code can swap
// them freely. The algorithms are the published ones, written here from their specifications:
//   * XXH32 / XXH64: https://github.com/Cyan4973/xxHash/blob/dev/doc/xxhash_spec.md
//   * MurmurHash3 x86_32 / x86_128 / x64_128 and the fmix finalizers: Austin Appleby, public domain
// Known-answer vectors from the reference's tests/utility/hash_test.cu are replayed in
// tests/test_hash_vectors.py against both this header (host build) and the C oracle.
#pragma once

#include <cuco/extent.cuh>

#include <cuda/std/array>
#include <cuda/std/cstddef>
#include <cuda/std/type_traits>

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace cuco {
namespace detail {

/// Reads the `index`-th T-sized little-endian word from an unaligned byte stream.
template <typename T, typename Byte, typename Index>
__host__ __device__ constexpr T load_chunk(Byte const* data, Index index) noexcept
{
  T word;
  memcpy(&word, reinterpret_cast<cuda::std::byte const*>(data) + index * sizeof(T), sizeof(T));
  return word;
}

__host__ __device__ constexpr std::uint32_t rotl32(std::uint32_t x, int r) noexcept
{
  return (x << r) | (x >> (32 - r));
}

__host__ __device__ constexpr std::uint64_t rotl64(std::uint64_t x, int r) noexcept
{
  return (x << r) | (x >> (64 - r));
}

__host__ __device__ constexpr std::uint32_t murmur_fmix32(std::uint32_t h) noexcept
{
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ constexpr std::uint64_t murmur_fmix64(std::uint64_t h) noexcept
{
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}

/// Hashes the object representation of `key` with `Derived::compute_hash`; small keys are copied to
/// a local first so the byte reads fold into register moves.
template <typename Derived, typename Key>
struct bytewise_hasher {
  using argument_type = Key;

  __host__ __device__ constexpr auto operator()(Key const& key) const noexcept
  {
    auto const& self = static_cast<Derived const&>(*this);
    if constexpr (sizeof(Key) <= 16) {
      Key const local = key;
      return self.compute_hash(reinterpret_cast<cuda::std::byte const*>(&local),
                               cuco::extent<std::size_t, sizeof(Key)>{});
    } else {
      return self.compute_hash(reinterpret_cast<cuda::std::byte const*>(&key),
                               cuco::extent<std::size_t, sizeof(Key)>{});
    }
  }
};

// ------------------------------------------------------------------------------------------------
// XXH32
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct XXHash_32 : bytewise_hasher<XXHash_32<Key>, Key> {
  using result_type = std::uint32_t;

  __host__ __device__ constexpr XXHash_32(std::uint32_t seed = 0) : seed_{seed} {}

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(cuda::std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    constexpr std::uint32_t P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu,
                            P5      = 0x165667B1u;
    std::size_t const len           = static_cast<std::size_t>(size);
    std::size_t const num_stripes   = len / 16;
    std::uint32_t h                 = 0;

    if (num_stripes > 0) {
      std::uint32_t acc[4] = {seed_ + P1 + P2, seed_ + P2, seed_, seed_ - P1};
      for (std::size_t s = 0; s != num_stripes; ++s) {
        for (int lane = 0; lane < 4; ++lane) {
          acc[lane] += load_chunk<std::uint32_t>(bytes, 4 * s + lane) * P2;
          acc[lane] = rotl32(acc[lane], 13) * P1;
        }
      }
      h = rotl32(acc[0], 1) + rotl32(acc[1], 7) + rotl32(acc[2], 12) + rotl32(acc[3], 18);
    } else {
      h = seed_ + P5;
    }
    h += static_cast<std::uint32_t>(len);

    std::size_t pos = num_stripes * 16;
    for (; pos + 4 <= len; pos += 4) {
      h += load_chunk<std::uint32_t>(bytes, pos / 4) * P3;
      h = rotl32(h, 17) * P4;
    }
    for (; pos < len; ++pos) {
      h += (cuda::std::to_integer<std::uint32_t>(bytes[pos]) & 0xffu) * P5;
      h = rotl32(h, 11) * P1;
    }

    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    h ^= h >> 16;
    return h;
  }

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    return compute_hash(reinterpret_cast<cuda::std::byte const*>(bytes), size);
  }

 private:
  std::uint32_t seed_;
};

// ------------------------------------------------------------------------------------------------
// XXH64
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct XXHash_64 : bytewise_hasher<XXHash_64<Key>, Key> {
  using result_type = std::uint64_t;

  __host__ __device__ constexpr XXHash_64(std::uint64_t seed = 0) : seed_{seed} {}

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(cuda::std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    std::size_t const len         = static_cast<std::size_t>(size);
    std::size_t const num_stripes = len / 32;
    std::uint64_t h               = 0;

    if (num_stripes > 0) {
      std::uint64_t acc[4] = {seed_ + P1 + P2, seed_ + P2, seed_, seed_ - P1};
      for (std::size_t s = 0; s != num_stripes; ++s) {
        for (int lane = 0; lane < 4; ++lane) {
          acc[lane] = round(acc[lane], load_chunk<std::uint64_t>(bytes, 4 * s + lane));
        }
      }
      h = rotl64(acc[0], 1) + rotl64(acc[1], 7) + rotl64(acc[2], 12) + rotl64(acc[3], 18);
      for (int lane = 0; lane < 4; ++lane) {
        h = (h ^ round(0, acc[lane])) * P1 + P4;
      }
    } else {
      h = seed_ + P5;
    }
    h += static_cast<std::uint64_t>(len);

    std::size_t pos = num_stripes * 32;
    for (; pos + 8 <= len; pos += 8) {
      h ^= round(0, load_chunk<std::uint64_t>(bytes, pos / 8));
      h = rotl64(h, 27) * P1 + P4;
    }
    if (pos + 4 <= len) {
      h ^= static_cast<std::uint64_t>(load_chunk<std::uint32_t>(bytes, pos / 4)) * P1;
      h = rotl64(h, 23) * P2 + P3;
      pos += 4;
    }
    for (; pos < len; ++pos) {
      h ^= (cuda::std::to_integer<std::uint64_t>(bytes[pos]) & 0xffu) * P5;
      h = rotl64(h, 11) * P1;
    }

    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
  }

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    return compute_hash(reinterpret_cast<cuda::std::byte const*>(bytes), size);
  }

 private:
  static constexpr std::uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full,
                                 P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull,
                                 P5 = 0x27D4EB2F165667C5ull;

  __host__ __device__ static constexpr std::uint64_t round(std::uint64_t acc,
                                                           std::uint64_t input) noexcept
  {
    return rotl64(acc + input * P2, 31) * P1;
  }

  std::uint64_t seed_;
};

// ------------------------------------------------------------------------------------------------
// MurmurHash3 finalizers used directly as integer hashes: fmix(key ^ seed)
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct MurmurHash3_fmix32 {
  static_assert(sizeof(Key) == 4, "Key type must be 4 bytes in size.");
  using argument_type = Key;
  using result_type   = std::uint32_t;

  __host__ __device__ constexpr MurmurHash3_fmix32(std::uint32_t seed = 0) : seed_{seed} {}
  __host__ __device__ constexpr result_type operator()(Key const& key) const noexcept
  {
    return murmur_fmix32(static_cast<std::uint32_t>(key) ^ seed_);
  }

 private:
  std::uint32_t seed_;
};

template <typename Key>
struct MurmurHash3_fmix64 {
  static_assert(sizeof(Key) == 8, "Key type must be 8 bytes in size.");
  using argument_type = Key;
  using result_type   = std::uint64_t;

  __host__ __device__ constexpr MurmurHash3_fmix64(std::uint64_t seed = 0) : seed_{seed} {}
  __host__ __device__ constexpr result_type operator()(Key const& key) const noexcept
  {
    return murmur_fmix64(static_cast<std::uint64_t>(key) ^ seed_);
  }

 private:
  std::uint64_t seed_;
};

// ------------------------------------------------------------------------------------------------
// MurmurHash3_x86_32
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct MurmurHash3_32 : bytewise_hasher<MurmurHash3_32<Key>, Key> {
  using result_type = std::uint32_t;

  __host__ __device__ constexpr MurmurHash3_32(std::uint32_t seed = 0) : seed_{seed} {}

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(cuda::std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    constexpr std::uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    std::size_t const len      = static_cast<std::size_t>(size);
    std::size_t const blocks   = len / 4;
    std::uint32_t h            = seed_;

    for (std::size_t i = 0; i != blocks; ++i) {
      std::uint32_t k = load_chunk<std::uint32_t>(bytes, i);
      k *= c1;
      k = rotl32(k, 15);
      k *= c2;
      h ^= k;
      h = rotl32(h, 13);
      h = h * 5 + 0xe6546b64u;
    }

    std::uint32_t tail = 0;
    for (std::size_t i = len & 3; i > 0; --i) {
      tail = (tail << 8) | (cuda::std::to_integer<std::uint32_t>(bytes[blocks * 4 + i - 1]) & 0xffu);
    }
    if (len & 3) {
      tail *= c1;
      tail = rotl32(tail, 15);
      tail *= c2;
      h ^= tail;
    }

    h ^= static_cast<std::uint32_t>(len);
    return murmur_fmix32(h);
  }

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    return compute_hash(reinterpret_cast<cuda::std::byte const*>(bytes), size);
  }

 private:
  std::uint32_t seed_;
};

// ------------------------------------------------------------------------------------------------
// MurmurHash3_x64_128
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct MurmurHash3_x64_128 : bytewise_hasher<MurmurHash3_x64_128<Key>, Key> {
  using result_type = cuda::std::array<std::uint64_t, 2>;

  __host__ __device__ constexpr MurmurHash3_x64_128(std::uint64_t seed = 0) : seed_{seed} {}

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(cuda::std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    constexpr std::uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    std::size_t const len      = static_cast<std::size_t>(size);
    std::size_t const blocks   = len / 16;
    std::uint64_t h1 = seed_, h2 = seed_;

    for (std::size_t i = 0; i != blocks; ++i) {
      std::uint64_t k1 = load_chunk<std::uint64_t>(bytes, 2 * i);
      std::uint64_t k2 = load_chunk<std::uint64_t>(bytes, 2 * i + 1);
      k1 *= c1;
      k1 = rotl64(k1, 31);
      k1 *= c2;
      h1 ^= k1;
      h1 = rotl64(h1, 27);
      h1 += h2;
      h1 = h1 * 5 + 0x52dce729u;
      k2 *= c2;
      k2 = rotl64(k2, 33);
      k2 *= c1;
      h2 ^= k2;
      h2 = rotl64(h2, 31);
      h2 += h1;
      h2 = h2 * 5 + 0x38495ab5u;
    }

    // tail: up to 15 bytes, low 8 feed k1 and the rest k2
    std::size_t const rem = len & 15;
    std::uint64_t k1 = 0, k2 = 0;
    for (std::size_t i = rem; i > 8; --i) {
      k2 = (k2 << 8) | (cuda::std::to_integer<std::uint64_t>(bytes[blocks * 16 + i - 1]) & 0xffu);
    }
    for (std::size_t i = (rem < 8 ? rem : 8); i > 0; --i) {
      k1 = (k1 << 8) | (cuda::std::to_integer<std::uint64_t>(bytes[blocks * 16 + i - 1]) & 0xffu);
    }
    if (rem > 8) {
      k2 *= c2;
      k2 = rotl64(k2, 33);
      k2 *= c1;
      h2 ^= k2;
    }
    if (rem > 0) {
      k1 *= c1;
      k1 = rotl64(k1, 31);
      k1 *= c2;
      h1 ^= k1;
    }

    h1 ^= static_cast<std::uint64_t>(len);
    h2 ^= static_cast<std::uint64_t>(len);
    h1 += h2;
    h2 += h1;
    h1 = murmur_fmix64(h1);
    h2 = murmur_fmix64(h2);
    h1 += h2;
    h2 += h1;
    return {h1, h2};
  }

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    return compute_hash(reinterpret_cast<cuda::std::byte const*>(bytes), size);
  }

 private:
  std::uint64_t seed_;
};

// ------------------------------------------------------------------------------------------------
// MurmurHash3_x86_128
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct MurmurHash3_x86_128 : bytewise_hasher<MurmurHash3_x86_128<Key>, Key> {
  using result_type = cuda::std::array<std::uint32_t, 4>;

  __host__ __device__ constexpr MurmurHash3_x86_128(std::uint32_t seed = 0) : seed_{seed} {}

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(cuda::std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    constexpr std::uint32_t c[4]    = {0x239b961bu, 0xab0e9789u, 0x38b34ae5u, 0xa1e38b93u};
    constexpr int krot[4]           = {15, 16, 17, 18};
    constexpr int hrot[4]           = {19, 17, 15, 13};
    constexpr std::uint32_t hadd[4] = {0x561ccd1bu, 0x0bcaa747u, 0x96cd1c35u, 0x32ac3b17u};
    std::size_t const len           = static_cast<std::size_t>(size);
    std::size_t const blocks        = len / 16;
    std::uint32_t h[4]              = {seed_, seed_, seed_, seed_};

    for (std::size_t i = 0; i != blocks; ++i) {
      for (int j = 0; j < 4; ++j) {
        std::uint32_t k = load_chunk<std::uint32_t>(bytes, 4 * i + j);
        k *= c[j];
        k = rotl32(k, krot[j]);
        k *= c[(j + 1) & 3];
        h[j] ^= k;
        h[j] = rotl32(h[j], hrot[j]);
        h[j] += h[(j + 1) & 3];
        h[j] = h[j] * 5 + hadd[j];
      }
    }

    // tail: byte t (0..14) feeds lane t / 4
    std::size_t const rem = len & 15;
    for (int j = 3; j >= 0; --j) {
      std::size_t const lo = 4 * static_cast<std::size_t>(j);
      if (rem > lo) {
        std::size_t const hi = (rem < lo + 4) ? rem : lo + 4;
        std::uint32_t k      = 0;
        for (std::size_t t = hi; t > lo; --t) {
          k = (k << 8) | (cuda::std::to_integer<std::uint32_t>(bytes[blocks * 16 + t - 1]) & 0xffu);
        }
        k *= c[j];
        k = rotl32(k, krot[j]);
        k *= c[(j + 1) & 3];
        h[j] ^= k;
      }
    }

    for (int j = 0; j < 4; ++j) {
      h[j] ^= static_cast<std::uint32_t>(len);
    }
    h[0] += h[1];
    h[0] += h[2];
    h[0] += h[3];
    h[1] += h[0];
    h[2] += h[0];
    h[3] += h[0];
    for (int j = 0; j < 4; ++j) {
      h[j] = murmur_fmix32(h[j]);
    }
    h[0] += h[1];
    h[0] += h[2];
    h[0] += h[3];
    h[1] += h[0];
    h[2] += h[0];
    h[3] += h[0];
    return {h[0], h[1], h[2], h[3]};
  }

  template <typename Extent>
  __host__ __device__ constexpr result_type compute_hash(std::byte const* bytes,
                                                         Extent size) const noexcept
  {
    return compute_hash(reinterpret_cast<cuda::std::byte const*>(bytes), size);
  }

 private:
  std::uint32_t seed_;
};

// ------------------------------------------------------------------------------------------------
// identity: the key is its own hash (perfect hashing when keys < capacity)
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct identity_hash {
  using argument_type = Key;
  using result_type   = cuda::std::conditional_t<sizeof(Key) <= 4, std::uint32_t, std::uint64_t>;

  static_assert(cuda::std::is_convertible_v<Key, result_type>,
                "Key type must be convertible to result_type");

  __host__ __device__ constexpr result_type operator()(Key const& key) const noexcept
  {
    return static_cast<result_type>(key);
  }
};

}  // namespace detail

template <typename Key>
using identity_hash = detail::identity_hash<Key>;
template <typename Key>
using murmurhash3_fmix_32 = detail::MurmurHash3_fmix32<Key>;
template <typename Key>
using murmurhash3_fmix_64 = detail::MurmurHash3_fmix64<Key>;
template <typename Key>
using murmurhash3_32 = detail::MurmurHash3_32<Key>;
template <typename Key>
using murmurhash3_x64_128 = detail::MurmurHash3_x64_128<Key>;
template <typename Key>
using murmurhash3_x86_128 = detail::MurmurHash3_x86_128<Key>;
template <typename Key>
using xxhash_32 = detail::XXHash_32<Key>;
template <typename Key>
using xxhash_64 = detail::XXHash_64<Key>;

/// Hash used by containers when none is named.
template <typename Key>
using default_hash_function = xxhash_32<Key>;

}  // namespace cuco
