// sm_100a slot primitives: whole-chunk vector loads with explicit cache policy and single-shot
// compare-and-swap on 4-, 8- and 16-byte slots.
//
// These replace the reference's per-field `cuda::atomic_ref` sequences
// (include/cuco/detail/open_addressing/open_addressing_ref_impl.cuh:1549-1664: packed_cas,
// back_to_back_cas, cas_dependent_write) wherever the slot is "one-shot claimable":
//   *  4/8-byte slots  -> one 32/64-bit CAS on the packed slot image (same as the reference)
//   * 16-byte slots     -> one native `atom.cas.b128` (ATOMG.E.CAS.128 in SASS). Key and payload
//                          become visible together, so writers need no second atomic and readers
//                          never observe a key whose payload is still the sentinel.
// Table reads go through `ld.global` with the width of a whole probing chunk (up to 256 bits =
// one 32-byte DRAM sector, LDG.E.256) and a cache policy chosen by the caller.
#pragma once

#include <cuco/detail/__config>
#include <cuco/pair.cuh>

#include <cuda/atomic>
#include <cuda/std/type_traits>

#include <cstdint>
#include <cstring>

namespace cuco::b200 {

/// How a table read interacts with the caches.
enum class load_policy : int {
  plain,     ///< ordinary generic load (works on shared memory too)
  readonly,  ///< table is not written during the kernel: non-coherent path, no L1 allocation
  streaming, ///< weak global load that does not allocate in L1 (mutating kernels: staleness is
             ///< harmless because every claim is validated by the CAS result)
  coherent   ///< relaxed.gpu load served by L2 (only needed when a stale value would be acted on)
};

/// Raw bytes of one probing chunk.
template <int Bytes>
struct alignas(Bytes < 16 ? Bytes : 16) raw_chunk {
  static_assert(Bytes == 4 || Bytes == 8 || Bytes == 16 || Bytes == 32);
  std::uint32_t w[Bytes / 4];
};

#define CUCO_B200_LD(PREFIX, BYTES_TAG)                                                          \
  if constexpr (BYTES_TAG == 4) {                                                                \
    asm volatile(PREFIX ".u32 %0, [%1];" : "=r"(r.w[0]) : "l"(p) : "memory");                    \
  } else if constexpr (BYTES_TAG == 8) {                                                         \
    asm volatile(PREFIX ".v2.u32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p) : "memory"); \
  } else if constexpr (BYTES_TAG == 16) {                                                        \
    asm volatile(PREFIX ".v4.u32 {%0,%1,%2,%3}, [%4];"                                           \
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])                        \
                 : "l"(p)                                                                        \
                 : "memory");                                                                    \
  } else {                                                                                       \
    asm volatile(PREFIX ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                               \
                 : "=r"(r.w[0]),                                                                 \
                   "=r"(r.w[1]),                                                                 \
                   "=r"(r.w[2]),                                                                 \
                   "=r"(r.w[3]),                                                                 \
                   "=r"(r.w[4]),                                                                 \
                   "=r"(r.w[5]),                                                                 \
                   "=r"(r.w[6]),                                                                 \
                   "=r"(r.w[7])                                                                  \
                 : "l"(p)                                                                        \
                 : "memory");                                                                    \
  }

/// Loads `Bytes` (4/8/16/32, naturally aligned) as one instruction: LDG.E.{32,64,128,256}.
template <int Bytes, load_policy Policy>
__device__ __forceinline__ raw_chunk<Bytes> load_chunk_bytes(void const* p) noexcept
{
  raw_chunk<Bytes> r;
  if constexpr (Policy == load_policy::plain) {
    if constexpr (Bytes == 4) {
      r.w[0] = *static_cast<std::uint32_t const*>(p);
    } else if constexpr (Bytes == 8) {
      auto const v = *static_cast<uint2 const*>(p);
      r.w[0]       = v.x;
      r.w[1]       = v.y;
    } else {
      auto const* q = static_cast<uint4 const*>(p);
#pragma unroll
      for (int i = 0; i < Bytes / 16; ++i) {
        auto const v   = q[i];
        r.w[4 * i]     = v.x;
        r.w[4 * i + 1] = v.y;
        r.w[4 * i + 2] = v.z;
        r.w[4 * i + 3] = v.w;
      }
    }
  } else if constexpr (Policy == load_policy::readonly) {
    CUCO_B200_LD("ld.global.nc.L1::no_allocate", Bytes)
  } else if constexpr (Policy == load_policy::streaming) {
    CUCO_B200_LD("ld.global.L1::no_allocate", Bytes)
  } else {
    CUCO_B200_LD("ld.relaxed.gpu.global.L1::no_allocate", Bytes)
  }
  return r;
}
#undef CUCO_B200_LD

/// Extracts slot `i` of a chunk as a typed value (folds to register moves).
template <typename Slot, int Bytes>
__device__ __forceinline__ Slot chunk_slot(raw_chunk<Bytes> const& c, int i) noexcept
{
  Slot s;
  memcpy(&s, reinterpret_cast<char const*>(c.w) + i * sizeof(Slot), sizeof(Slot));
  return s;
}

/// Streaming (read-once) global load of an input element; does not displace table lines in L1.
template <typename T>
__device__ __forceinline__ T load_streaming(T const* p) noexcept
{
  if constexpr (sizeof(T) == 16 && alignof(T) >= 16) {
    raw_chunk<16> r = load_chunk_bytes<16, load_policy::readonly>(p);
    return chunk_slot<T>(r, 0);
  } else if constexpr (sizeof(T) == 8 && alignof(T) >= 8) {
    raw_chunk<8> r = load_chunk_bytes<8, load_policy::readonly>(p);
    return chunk_slot<T>(r, 0);
  } else if constexpr (sizeof(T) == 4 && alignof(T) >= 4) {
    raw_chunk<4> r = load_chunk_bytes<4, load_policy::readonly>(p);
    return chunk_slot<T>(r, 0);
  } else {
    return *p;
  }
}

// ------------------------------------------------------------------------------------------------
// compare-and-swap on a whole slot
// ------------------------------------------------------------------------------------------------

/// True when a slot can be claimed with one hardware CAS covering all of its bits.
template <typename Slot>
__host__ __device__ constexpr bool is_single_cas_slot() noexcept
{
  if constexpr (sizeof(Slot) == 4 || sizeof(Slot) == 8) {
    return std::has_unique_object_representations_v<Slot> && alignof(Slot) >= sizeof(Slot);
  } else if constexpr (sizeof(Slot) == 16) {
    return std::has_unique_object_representations_v<Slot> && alignof(Slot) >= 16;
  } else {
    return false;
  }
}

template <cuda::thread_scope Scope>
struct scope_tag {};

/// 128-bit relaxed CAS; returns the value found at `address` (== expected on success).
/// Generic addressing, so it also serves block-scoped tables living in shared memory.
template <cuda::thread_scope Scope>
__device__ __forceinline__ void cas_b128(void* address,
                                         std::uint64_t exp_lo,
                                         std::uint64_t exp_hi,
                                         std::uint64_t des_lo,
                                         std::uint64_t des_hi,
                                         std::uint64_t& old_lo,
                                         std::uint64_t& old_hi) noexcept
{
#define CUCO_B200_CAS128(SCOPE)                                        \
  asm volatile(                                                        \
    "{\n\t"                                                            \
    ".reg .b128 e, d, o;\n\t"                                          \
    "mov.b128 e, {%2, %3};\n\t"                                        \
    "mov.b128 d, {%4, %5};\n\t"                                        \
    "atom.relaxed." SCOPE ".cas.b128 o, [%6], e, d;\n\t"               \
    "mov.b128 {%0, %1}, o;\n\t"                                        \
    "}"                                                                \
    : "=l"(old_lo), "=l"(old_hi)                                       \
    : "l"(exp_lo), "l"(exp_hi), "l"(des_lo), "l"(des_hi), "l"(address) \
    : "memory")
  if constexpr (Scope == cuda::thread_scope_system) {
    CUCO_B200_CAS128("sys");
  } else if constexpr (Scope == cuda::thread_scope_device) {
    CUCO_B200_CAS128("gpu");
  } else {
    // block and thread scope: CTA is the narrowest scope PTX atomics have
    CUCO_B200_CAS128("cta");
  }
#undef CUCO_B200_CAS128
}

/// One-shot CAS of a whole slot. Returns the slot value observed at `address`; the swap happened
/// iff that equals `expected` bit for bit.
template <cuda::thread_scope Scope, typename Slot>
__device__ __forceinline__ Slot cas_slot(Slot* address, Slot const& expected, Slot const& desired) noexcept
{
  static_assert(is_single_cas_slot<Slot>());
  Slot observed;
  if constexpr (sizeof(Slot) == 16) {
    std::uint64_t e[2], d[2], o[2];
    memcpy(e, &expected, 16);
    memcpy(d, &desired, 16);
    cas_b128<Scope>(address, e[0], e[1], d[0], d[1], o[0], o[1]);
    memcpy(&observed, o, 16);
  } else {
    using word = cuda::std::conditional_t<sizeof(Slot) == 4, std::uint32_t, std::uint64_t>;
    word e, d;
    memcpy(&e, &expected, sizeof(word));
    memcpy(&d, &desired, sizeof(word));
    cuda::atomic_ref<word, Scope> ref{*reinterpret_cast<word*>(address)};
    ref.compare_exchange_strong(e, d, cuda::memory_order_relaxed);
    memcpy(&observed, &e, sizeof(word));  // compare_exchange leaves the observed value in `e`
  }
  return observed;
}

/// Bitwise equality of two slots / keys (what "the CAS succeeded" and "is sentinel" mean).
template <typename T>
__host__ __device__ __forceinline__ bool same_bits(T const& a, T const& b) noexcept
{
  if constexpr (sizeof(T) == 4) {
    std::uint32_t x{}, y{};
    memcpy(&x, &a, 4);
    memcpy(&y, &b, 4);
    return x == y;
  } else if constexpr (sizeof(T) == 8) {
    std::uint64_t x{}, y{};
    memcpy(&x, &a, 8);
    memcpy(&y, &b, 8);
    return x == y;
  } else if constexpr (sizeof(T) == 16) {
    std::uint64_t x[2]{}, y[2]{};
    memcpy(x, &a, 16);
    memcpy(y, &b, 16);
    return x[0] == y[0] && x[1] == y[1];
  } else {
    auto const* pa = reinterpret_cast<unsigned char const*>(&a);
    auto const* pb = reinterpret_cast<unsigned char const*>(&b);
    for (std::size_t i = 0; i < sizeof(T); ++i) {
      if (pa[i] != pb[i]) { return false; }
    }
    return true;
  }
}

}  // namespace cuco::b200
