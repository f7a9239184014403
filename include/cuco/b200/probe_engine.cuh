// probe_engine — the device-side core every operation of static_map / static_set runs on.
//
// It plays the role of the reference's `open_addressing_ref_impl`
// (include/cuco/detail/open_addressing/open_addressing_ref_impl.cuh:88-1752) and of the map-only
// upsert logic (detail/static_map/static_map_ref.inl:464-1053), re-designed for sm_100a:
//
//  * One definition of the probe sequence. A probe step of the scheme covers a *bucket*: the
//    cg_size consecutive windows starting at window `base`, i.e. cg_size*window_size contiguous slots
//    (wrapping at the end of the table). Linear probing moves to the adjacent bucket, double hashing
//    jumps `step` windows. Tiles scan a bucket cooperatively (one window per lane, one ballot);
//    single threads scan it slot by slot in memory order with early exit. Both pick the lowest
//    available slot in bucket order, so they build and read the same tables.
//
//  * Thread-per-key scanning is expressed through a resumable `cursor`, so bulk kernels can keep
//    several keys in flight per thread (issue all chunk loads / CAS, then consume them) instead of
//    one dependent probe chain per thread.
//
//  * Table reads are chunk loads (up to 256 bit = one 32-byte sector) with an explicit cache
//    policy; slot claims are single 32/64/128-bit CAS when the slot allows it (slot_ops.cuh) and
//    fall back to key-CAS + payload store otherwise.
//
//  * Insertion takes the FIRST AVAILABLE slot (empty or erased) of the probe sequence, the
//    reference's rule (ref_impl.cuh:385-409): re-inserting a key that still sits further down its
//    cluster after a neighbour was erased stores it a second time, exactly as cuco does
//    (tests/test_baseline_configs_gpu.py::test_insert_after_erase_follows_cuco pins size(), the insert
//    count and lookups against cuco's own build). Compile with -DCUCO_B200_TOMBSTONE_AWARE_INSERT=1
//    for the stricter opt-in rule: an erased slot is only remembered as candidate while the scan
//    continues to the first truly empty slot (at most one full cycle), so a key can never be stored
//    twice. Without tombstones the two rules coincide.
//
// Result semantics (what parity is defined over) are those of SURVEY.md §8(a').
#pragma once

#include <cuco/b200/slot_ops.cuh>
#include <cuco/detail/__config>
#include <cuco/detail/utility/cuda.cuh>
#include <cuco/extent.cuh>
#include <cuco/pair.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/utility/traits.hpp>

#include <cuda/atomic>
#include <cuda/std/type_traits>
#include <thrust/device_reference.h>
#include <thrust/pair.h>

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include <cstdint>

namespace cuco {
namespace detail {

/// Outcome of comparing a probe key with a slot key. The numeric values are relied upon
/// (`count` sums them; EQUAL must be 1, UNEQUAL 0).
enum class equal_result : std::int32_t { UNEQUAL = 0, EQUAL = 1, EMPTY = 2, AVAILABLE = 3 };

/// Outcome of trying to claim a slot.
enum class insert_result : std::int32_t { CONTINUE = 0, SUCCESS = 1, DUPLICATE = 2 };

/// Public-ish helper kept under its reference name.
template <typename T>
__host__ __device__ inline bool bitwise_compare(T const& lhs, T const& rhs)
{
  static_assert(cuco::is_bitwise_comparable_v<T>,
                "Bitwise compared objects must have unique object representations or be explicitly "
                "declared as safe for bitwise comparison via cuco::is_bitwise_comparable_v.");
  return cuco::b200::same_bits(lhs, rhs);
}

}  // namespace detail

namespace b200 {

namespace cg = cooperative_groups;

using detail::equal_result;
using detail::insert_result;

/// Stand-in divisor for schemes without a second hash: `x % no_modulus{}` is x.
struct no_modulus {
  template <typename T>
  friend __host__ __device__ constexpr T operator%(T lhs, no_modulus) noexcept
  {
    return lhs;
  }
};

/// Divisor of the double-hashing step, N / cg_size - 1: a fast_int when N is a run-time value so the
/// per-key `hash2 % divisor` is a multiply-shift (the reference divides, probing_scheme_impl.inl:188).
template <bool IsDoubleHashing, typename Extent, typename SizeType, int CGSize>
struct step_modulus {
  using type = no_modulus;
  __host__ __device__ static constexpr type make(Extent const&) noexcept { return {}; }
};

template <typename Extent, typename SizeType, int CGSize>
struct step_modulus<true, Extent, SizeType, CGSize> {
  static constexpr bool dynamic =
    cuda::std::is_base_of_v<cuco::utility::fast_int<SizeType>, Extent>;
  using type = cuda::std::conditional_t<dynamic, cuco::utility::fast_int<SizeType>, SizeType>;
  __host__ __device__ static constexpr type make(Extent const& extent) noexcept
  {
    auto const n = static_cast<SizeType>(static_cast<SizeType>(extent) / CGSize - 1);
    return type{n < 1 ? SizeType{1} : n};
  }
};

template <typename Key,
          cuda::thread_scope Scope,
          typename KeyEqual,
          typename ProbingScheme,
          typename StorageRef,
          bool AllowsDuplicates>
class probe_engine {
  static_assert(sizeof(Key) <= 8, "Container does not support key types larger than 8 bytes.");
  static_assert(
    cuco::is_bitwise_comparable_v<Key>,
    "Key type must have unique object representations or have been explicitly declared as safe for "
    "bitwise comparison via specialization of cuco::is_bitwise_comparable_v<Key>.");
  static_assert(
    std::is_base_of_v<cuco::detail::probing_scheme_base<ProbingScheme::cg_size>, ProbingScheme>,
    "ProbingScheme must inherit from cuco::detail::probing_scheme_base");

 public:
  using key_type            = Key;
  using probing_scheme_type = ProbingScheme;
  using hasher              = typename probing_scheme_type::hasher;
  using storage_ref_type    = StorageRef;
  using window_type         = typename storage_ref_type::window_type;
  using value_type          = typename storage_ref_type::value_type;  ///< slot type
  using extent_type         = typename storage_ref_type::extent_type;
  using size_type           = typename storage_ref_type::size_type;
  using key_equal           = KeyEqual;
  using iterator            = typename storage_ref_type::iterator;
  using const_iterator      = typename storage_ref_type::const_iterator;

  static constexpr auto cg_size           = probing_scheme_type::cg_size;
  static constexpr auto window_size       = storage_ref_type::window_size;
  static constexpr auto thread_scope      = Scope;
  static constexpr bool has_payload       = !std::is_same_v<Key, value_type>;
  static constexpr bool allows_duplicates = AllowsDuplicates;
  static constexpr bool is_double_hashing = probing_scheme_type::is_double_hashing_scheme;

  static constexpr int slot_bytes   = sizeof(value_type);
  static constexpr int bucket_slots = cg_size * window_size;  ///< slots covered by one probe step
  static constexpr bool pow2_slot =
    (slot_bytes & (slot_bytes - 1)) == 0 && alignof(value_type) >= sizeof(value_type);
  /// Slots claimable with one hardware CAS (32/64/128 bit).
  static constexpr bool single_cas = is_single_cas_slot<value_type>();
  /// Insert rule on tables with an erased-key sentinel: false = the reference's first-AVAILABLE rule
  /// (default, bit-exact with cuco), true = opt-in stricter rule (see the header comment).
#if defined(CUCO_B200_TOMBSTONE_AWARE_INSERT) && CUCO_B200_TOMBSTONE_AWARE_INSERT
  static constexpr bool tombstone_aware_insert = true;
#else
  static constexpr bool tombstone_aware_insert = false;
#endif

  /// Chunk (in slots) that is always safe to load: one window, or one slot for odd geometries.
  static constexpr int window_chunk_slots =
    (pow2_slot && (window_size & (window_size - 1)) == 0 && window_size * slot_bytes <= 32)
      ? window_size
      : 1;
  /// Chunk (in slots) used by bulk kernels on container-owned (padded, 32 B aligned) storage.
  static constexpr int sector_chunk_slots = (pow2_slot && slot_bytes <= 16) ? 32 / slot_bytes : 1;

  // ----------------------------------------------------------------------------------------------
  // construction / accessors
  // ----------------------------------------------------------------------------------------------
  __host__ __device__ explicit constexpr probe_engine(value_type empty_slot_sentinel,
                                                      key_equal const& predicate,
                                                      probing_scheme_type const& probing_scheme,
                                                      storage_ref_type storage_ref) noexcept
    : empty_slot_{empty_slot_sentinel},
      erased_key_{key_of(empty_slot_sentinel)},
      eq_{predicate},
      scheme_{probing_scheme},
      storage_{storage_ref},
      step_modulus_{step_modulus_helper::make(storage_ref.window_extent())}
  {
  }

  __host__ __device__ explicit constexpr probe_engine(value_type empty_slot_sentinel,
                                                      key_type erased_key_sentinel,
                                                      key_equal const& predicate,
                                                      probing_scheme_type const& probing_scheme,
                                                      storage_ref_type storage_ref) noexcept
    : empty_slot_{empty_slot_sentinel},
      erased_key_{erased_key_sentinel},
      eq_{predicate},
      scheme_{probing_scheme},
      storage_{storage_ref},
      step_modulus_{step_modulus_helper::make(storage_ref.window_extent())}
  {
  }

  [[nodiscard]] __host__ __device__ constexpr key_type empty_key_sentinel() const noexcept
  {
    return key_of(empty_slot_);
  }
  template <bool Dummy = true, typename = cuda::std::enable_if_t<has_payload and Dummy>>
  [[nodiscard]] __host__ __device__ constexpr auto empty_value_sentinel() const noexcept
  {
    return empty_slot_.second;
  }
  [[nodiscard]] __host__ __device__ constexpr key_type erased_key_sentinel() const noexcept
  {
    return erased_key_;
  }
  [[nodiscard]] __host__ __device__ constexpr value_type empty_slot_sentinel() const noexcept
  {
    return empty_slot_;
  }
  [[nodiscard]] __host__ __device__ constexpr key_equal key_eq() const noexcept { return eq_; }
  [[nodiscard]] __host__ __device__ constexpr probing_scheme_type const& probing_scheme()
    const noexcept
  {
    return scheme_;
  }
  [[nodiscard]] __host__ __device__ constexpr hasher hash_function() const noexcept
  {
    return scheme_.hash_function();
  }
  [[nodiscard]] __host__ __device__ constexpr storage_ref_type storage_ref() const noexcept
  {
    return storage_;
  }
  [[nodiscard]] __host__ __device__ constexpr auto capacity() const noexcept
  {
    return storage_.capacity();
  }
  [[nodiscard]] __host__ __device__ constexpr extent_type window_extent() const noexcept
  {
    return storage_.window_extent();
  }
  [[nodiscard]] __device__ constexpr const_iterator end() const noexcept { return storage_.end(); }
  [[nodiscard]] __device__ constexpr iterator end() noexcept { return storage_.end(); }

  /// True when tombstones can exist in this table (an erased sentinel distinct from empty).
  [[nodiscard]] __host__ __device__ bool has_tombstones() const noexcept
  {
    return !same_bits(erased_key_, key_of(empty_slot_));
  }

  [[nodiscard]] __host__ __device__ value_type* slots() const noexcept
  {
    return reinterpret_cast<value_type*>(storage_.data());
  }

  // ----------------------------------------------------------------------------------------------
  // whole-table helpers callable by a group
  // ----------------------------------------------------------------------------------------------
  /// Copies all windows into `memory_to_use` (typically shared memory) with the whole group.
  template <typename CG>
  __device__ void make_copy(CG const& g, window_type* const memory_to_use) const noexcept
  {
    auto const n        = static_cast<size_type>(this->window_extent());
    auto const* src     = storage_.data();
    constexpr int bytes = sizeof(window_type);
    if constexpr (bytes % 16 == 0) {
      // 128-bit copies; windows are at least 16-byte aligned in both spaces
      auto const* s = reinterpret_cast<uint4 const*>(src);
      auto* d       = reinterpret_cast<uint4*>(memory_to_use);
      auto const m  = static_cast<size_type>(n) * (bytes / 16);
      for (size_type i = g.thread_rank(); i < m; i += g.size()) {
        d[i] = s[i];
      }
    } else {
      for (size_type i = g.thread_rank(); i < n; i += g.size()) {
        memory_to_use[i] = src[i];
      }
    }
    g.sync();
  }

  /// Sets every slot to the empty sentinel with the whole group, then synchronises it.
  template <typename CG>
  __device__ constexpr void initialize(CG const& g) noexcept
  {
    auto* const s = slots();
    auto const n  = static_cast<size_type>(this->capacity());
    for (size_type i = g.thread_rank(); i < n; i += g.size()) {
      s[i] = empty_slot_;
    }
    g.sync();
  }

  // ----------------------------------------------------------------------------------------------
  // value plumbing
  // ----------------------------------------------------------------------------------------------
  template <typename Value>
  [[nodiscard]] __host__ __device__ static constexpr auto const& key_of(Value const& value) noexcept
  {
    if constexpr (has_payload) {
      return thrust::raw_reference_cast(value).first;
    } else {
      return thrust::raw_reference_cast(value);
    }
  }

  template <typename Value>
  [[nodiscard]] __host__ __device__ constexpr auto const& extract_key(
    Value const& value) const noexcept
  {
    return key_of(value);
  }

  template <typename Value, typename Enable = std::enable_if_t<has_payload and sizeof(Value)>>
  [[nodiscard]] __device__ constexpr auto const& extract_payload(Value const& value) const noexcept
  {
    return thrust::raw_reference_cast(value).second;
  }

  /// Accepts anything pair-like for maps (converted payload, key left in its own type for
  /// heterogeneous lookup) and passes set elements through.
  template <typename T>
  [[nodiscard]] __device__ constexpr auto heterogeneous_value(T const& value) const noexcept
  {
    if constexpr (has_payload and not cuda::std::is_same_v<T, value_type>) {
      using mapped_type = decltype(empty_slot_.second);
      if constexpr (cuco::detail::is_cuda_std_pair_like<T>::value) {
        return cuco::pair{cuda::std::get<0>(value),
                          static_cast<mapped_type>(cuda::std::get<1>(value))};
      } else {
        return cuco::pair{thrust::raw_reference_cast(value.first),
                          static_cast<mapped_type>(value.second)};
      }
    } else {
      return thrust::raw_reference_cast(value);
    }
  }

  /// The slot image that gets stored for an input value.
  template <typename T>
  [[nodiscard]] __device__ constexpr value_type native_value(T const& value) const noexcept
  {
    if constexpr (has_payload) {
      return value_type{static_cast<key_type>(key_of(value)), value.second};
    } else {
      return static_cast<value_type>(value);
    }
  }

  [[nodiscard]] __device__ constexpr value_type erased_slot_sentinel() const noexcept
  {
    if constexpr (has_payload) {
      return value_type{erased_key_, empty_slot_.second};
    } else {
      return erased_key_;
    }
  }

  // ----------------------------------------------------------------------------------------------
  // slot classification (sentinels are tested bitwise BEFORE the user predicate ever sees them)
  // ----------------------------------------------------------------------------------------------
  template <typename ProbeKey>
  [[nodiscard]] __device__ constexpr equal_result compare_keys(ProbeKey const& probe,
                                                               key_type const& slot_key) const noexcept
  {
    return eq_(probe, slot_key) ? equal_result::EQUAL : equal_result::UNEQUAL;
  }

  /// Lookup view: EMPTY ends the search, tombstones are ordinary mismatches.
  template <typename ProbeKey>
  [[nodiscard]] __device__ constexpr equal_result classify_lookup(
    ProbeKey const& probe, key_type const& slot_key) const noexcept
  {
    return same_bits(slot_key, key_of(empty_slot_)) ? equal_result::EMPTY
                                                    : compare_keys(probe, slot_key);
  }

  /// Insert view: empty and erased slots are both AVAILABLE. Tables that allow duplicates never
  /// report EQUAL here: an occupied slot is just occupied, whatever key it holds.
  template <typename ProbeKey>
  [[nodiscard]] __device__ constexpr equal_result classify_insert(
    ProbeKey const& probe, key_type const& slot_key) const noexcept
  {
    if (same_bits(slot_key, key_of(empty_slot_)) || same_bits(slot_key, erased_key_)) {
      return equal_result::AVAILABLE;
    }
    if constexpr (allows_duplicates) {
      return equal_result::UNEQUAL;
    } else {
      return compare_keys(probe, slot_key);
    }
  }

  [[nodiscard]] __device__ constexpr bool is_empty_key(key_type const& slot_key) const noexcept
  {
    return same_bits(slot_key, key_of(empty_slot_));
  }

  // ----------------------------------------------------------------------------------------------
  // cursor: resumable position in the probe sequence of one key (thread-per-key scanning)
  // ----------------------------------------------------------------------------------------------
  struct cursor {
    size_type slot;     ///< absolute index of the next slot to examine
    size_type base;     ///< base window of the current bucket      (double hashing only)
    size_type step;     ///< windows between consecutive buckets    (double hashing only)
    std::int32_t used;  ///< slots of the current bucket already examined (double hashing only)
  };

  template <typename ProbeKey>
  [[nodiscard]] __device__ constexpr cursor make_cursor(ProbeKey const& key) const noexcept
  {
    auto const extent = storage_.window_extent();
    auto const home   = static_cast<size_type>(scheme_.template home_hash<size_type>(key) % extent);
    cursor c{};
    c.slot = home * window_size;
    if constexpr (is_double_hashing) {
      c.base = home;
      c.step = static_cast<size_type>(
        (scheme_.template step_hash<size_type>(key) % step_modulus_ + 1) * cg_size);
      c.used = 0;
    }
    return c;
  }

  /// First slot of `c`'s current chunk when chunks hold `ChunkSlots` slots.
  template <int ChunkSlots>
  [[nodiscard]] __device__ static constexpr size_type chunk_begin(cursor const& c) noexcept
  {
    if constexpr (ChunkSlots == 1) {
      return c.slot;
    } else {
      return c.slot - (c.slot % ChunkSlots);
    }
  }

  /// Number of slots of the current chunk that belong to the probe sequence, starting at c.slot.
  template <int ChunkSlots>
  [[nodiscard]] __device__ constexpr int chunk_valid(cursor const& c) const noexcept
  {
    auto const cap  = static_cast<size_type>(storage_.capacity());
    int const first = static_cast<int>(c.slot - chunk_begin<ChunkSlots>(c));
    int valid       = ChunkSlots - first;
    auto const to_end = cap - c.slot;  // slots before the table wraps
    if (to_end < static_cast<size_type>(valid)) { valid = static_cast<int>(to_end); }
    if constexpr (is_double_hashing) {
      int const left = bucket_slots - c.used;
      if (left < valid) { valid = left; }
    }
    return valid;
  }

  /// Moves the cursor past `n` examined slots (n <= chunk_valid).
  __device__ constexpr void advance(cursor& c, int n) const noexcept
  {
    auto const cap = static_cast<size_type>(storage_.capacity());
    c.slot += n;
    if (c.slot >= cap) { c.slot -= cap; }
    if constexpr (is_double_hashing) {
      c.used += n;
      if (c.used >= bucket_slots) {
        auto const windows = static_cast<size_type>(storage_.window_extent());
        c.base += c.step;
        if (c.base >= windows) { c.base -= windows; }
        c.slot = c.base * window_size;
        c.used = 0;
      }
    }
  }

  template <int ChunkSlots, load_policy Policy>
  [[nodiscard]] __device__ auto load_chunk(cursor const& c) const noexcept
  {
    return load_chunk_bytes<ChunkSlots * slot_bytes, chunk_policy<ChunkSlots, Policy>()>(
      slots() + chunk_begin<ChunkSlots>(c));
  }

  /// Visits the slots of `key`'s probe sequence in order until `visit(index, slot)` returns true.
  template <int ChunkSlots, load_policy Policy, typename Visit>
  __device__ void walk(cursor c, Visit&& visit) const noexcept
  {
    while (true) {
      auto const begin = chunk_begin<ChunkSlots>(c);
      int const first  = static_cast<int>(c.slot - begin);
      int const valid  = chunk_valid<ChunkSlots>(c);
      auto const raw   = load_chunk<ChunkSlots, Policy>(c);
#pragma unroll
      for (int i = 0; i < ChunkSlots; ++i) {
        if (i >= first && i < first + valid) {
          if (visit(static_cast<size_type>(begin + i), chunk_slot<value_type>(raw, i))) { return; }
        }
      }
      advance(c, valid);
    }
  }

  /// `walk` with look-ahead inside the 128-byte line: together with the chunk the cursor points at,
  /// up to `Ahead - 1` following chunks of the probe sequence are loaded before the first is
  /// examined, as long as they lie in the SAME 128-byte line of the slot array. Where a chunk lies
  /// depends only on the cursor, never on table contents, so the loads are independent and their
  /// latencies overlap; DRAM delivers the whole line on the first miss anyway, so the extra loads
  /// are L2 hits and never pull in a line the walk might not need (measured: unbounded look-ahead
  /// costs 15-25 % on count / retrieve because a third of the walks then touch one more line).
  /// Meant for walks that are expected to be long (enumerating duplicates up to the first empty slot)
  /// on container-owned storage (256-byte aligned base).
  template <int ChunkSlots, load_policy Policy, int Ahead, typename Visit>
  __device__ void walk_ahead(cursor c, Visit&& visit) const noexcept
  {
    static_assert(Ahead >= 1);
    if constexpr (Ahead == 1) {
      walk<ChunkSlots, Policy>(c, visit);
    } else {
      constexpr size_type line_slots = 128 / slot_bytes;
      while (true) {
        cursor at[Ahead];
        int valid[Ahead];
        bool live[Ahead];
        at[0]   = c;
        live[0] = true;
#pragma unroll
        for (int a = 0; a < Ahead; ++a) {
          valid[a] = chunk_valid<ChunkSlots>(at[a]);
          if (a + 1 < Ahead) {
            at[a + 1] = at[a];
            advance(at[a + 1], valid[a]);
            live[a + 1] = live[a] && (chunk_begin<ChunkSlots>(at[a + 1]) / line_slots ==
                                      chunk_begin<ChunkSlots>(at[a]) / line_slots);
          }
        }
        decltype(load_chunk<ChunkSlots, Policy>(c)) raw[Ahead];
#pragma unroll
        for (int a = 0; a < Ahead; ++a) {
          if (live[a]) { raw[a] = load_chunk<ChunkSlots, Policy>(at[a]); }
        }
        bool resumed = false;
#pragma unroll
        for (int a = 0; a < Ahead; ++a) {
          if (!live[a]) {
            if (!resumed) {
              c       = at[a];  // first chunk outside the line: next round starts here
              resumed = true;
            }
            continue;
          }
          auto const begin = chunk_begin<ChunkSlots>(at[a]);
          int const first  = static_cast<int>(at[a].slot - begin);
#pragma unroll
          for (int i = 0; i < ChunkSlots; ++i) {
            if (i >= first && i < first + valid[a]) {
              if (visit(static_cast<size_type>(begin + i), chunk_slot<value_type>(raw[a], i))) { return; }
            }
          }
        }
        if (!resumed) {
          c = at[Ahead - 1];
          advance(c, valid[Ahead - 1]);
        }
      }
    }
  }

  // ----------------------------------------------------------------------------------------------
  // claiming a slot
  // ----------------------------------------------------------------------------------------------
  /// Tries to replace `expected` (an available slot image we observed) by `desired` at `address`.
  /// On CONTINUE/DUPLICATE, `expected` is updated to what is there now (key always; payload too for
  /// single-CAS slots).
  [[nodiscard]] __device__ insert_result try_claim(value_type* address,
                                                   value_type& expected,
                                                   value_type const& desired) const noexcept
  {
    return this->try_claim(address, expected, desired, key_of(desired));
  }

  /// Same, for a key that arrived in its own (heterogeneous) type: the duplicate test applies the
  /// user's predicate to THAT key, as the reference does (ref_impl.cuh:385-409), never to the
  /// converted one.
  template <typename ProbeKey>
  [[nodiscard]] __device__ insert_result try_claim(value_type* address,
                                                   value_type& expected,
                                                   value_type const& desired,
                                                   ProbeKey const& probe_key) const noexcept
  {
    if constexpr (single_cas) {
      auto const observed = cas_slot<Scope>(address, expected, desired);
      if (same_bits(observed, expected)) { return insert_result::SUCCESS; }
      expected = observed;
    } else {
      // padded or otherwise non-packable slot: claim the key, then publish the payload
      static_assert(has_payload, "key-only slots of size 4/8 are always single-CAS");
      auto expected_key = expected.first;
      cuda::atomic_ref<key_type, Scope> key_ref{address->first};
      if (key_ref.compare_exchange_strong(
            expected_key, static_cast<key_type>(desired.first), cuda::memory_order_relaxed)) {
        cuda::atomic_ref<decltype(address->second), Scope> payload_ref{address->second};
        payload_ref.store(desired.second, cuda::memory_order_relaxed);
        return insert_result::SUCCESS;
      }
      expected.first = expected_key;
    }
    if constexpr (!allows_duplicates) {
      if (!same_bits(key_of(expected), key_of(empty_slot_)) && !same_bits(key_of(expected), erased_key_) &&
          eq_(probe_key, key_of(expected))) {
        return insert_result::DUPLICATE;
      }
    }
    return insert_result::CONTINUE;
  }

  /// Spins until the payload next to a freshly claimed key has been published (two-step writers).
  template <typename T>
  __device__ void wait_for_payload(T& payload, T const& sentinel) const noexcept
  {
    cuda::atomic_ref<T, Scope> ref{payload};
    T current;
    do {
      current = ref.load(cuda::memory_order_relaxed);
    } while (same_bits(current, sentinel));
  }

  // ----------------------------------------------------------------------------------------------
  // thread-per-key operations (any bucket geometry). `ChunkSlots` / `Policy` let bulk kernels on
  // container-owned storage use sector chunks and cache hints; the defaults are safe on any memory.
  // ----------------------------------------------------------------------------------------------

  /// Shared insertion driver. Walks to the end of the key's cluster, claims the first available
  /// slot. `on_equal(slot_ptr)` runs when the key is already present (returns nothing);
  /// `Claim(slot_ptr, expected&, desired)` performs the claim.
  /// Returns {slot pointer, true} on a new entry, {slot pointer, false} if present.
  template <int ChunkSlots, load_policy Policy, typename Value, typename Claim>
  __device__ thrust::pair<value_type*, bool> insert_driver(Value const& val, Claim&& claim) noexcept
  {
    auto const& key      = key_of(val);
    auto const start     = make_cursor(key);
    auto* const table    = slots();
    bool const tombstone = tombstone_aware_insert && has_tombstones();
    size_type seen       = 0;  // slots visited (strict rule only): bounds the walk to one full cycle

    while (true) {  // restarted only after losing a tombstone candidate to another key
      value_type* found   = nullptr;
      bool inserted       = false;
      bool restart        = false;
      value_type* cand    = nullptr;
      value_type cand_img = empty_slot_;

      size_type ordinal = 0;  // slots of the probe sequence examined before the current one
      walk<ChunkSlots, Policy>(start, [&](size_type index, value_type slot) {
        auto const state = classify_insert(key, key_of(slot));
        struct count_on_exit {
          size_type& n;
          __device__ ~count_on_exit() { ++n; }
        } const counted{ordinal};
        if (tombstone) { ++seen; }
        if constexpr (!allows_duplicates) {
          if (state == equal_result::EQUAL) {
            found = table + index;
            return true;
          }
        }
        if (state != equal_result::AVAILABLE) { return false; }

        if constexpr (!allows_duplicates && cg_size > 1) {
          // Reference rule for a cg_size-wide probe step (ref_impl.cuh:433-458): every lane reports the
          // first slot of ITS window that is available or equal, and an EQUAL report from any lane of
          // the step outranks an AVAILABLE one from a lower lane. Only an erased slot can precede the
          // key inside one step (an empty one would have taken the key when it was inserted), so the
          // rest of the step is examined only then.
          if (has_tombstones() && !is_empty_key(key_of(slot))) {
            constexpr int group_slots = cg_size * window_size;
            int const pos             = static_cast<int>(ordinal % group_slots);
            auto const cap            = static_cast<size_type>(storage_.capacity());
            for (int p = (pos / window_size + 1) * window_size; p < group_slots;) {
              size_type at = index + static_cast<size_type>(p - pos);
              if (at >= cap) { at -= cap; }
              key_type later_key;  // the key leads the slot image (pair.first, or the slot itself)
              memcpy(&later_key, table + at, sizeof(key_type));
              auto const later = classify_insert(key, later_key);
              if (later == equal_result::EQUAL) {
                found = table + at;
                return true;
              }
              p = (later == equal_result::AVAILABLE) ? (p / window_size + 1) * window_size : p + 1;
            }
          }
        }
        bool truly_empty = !tombstone || is_empty_key(key_of(slot));
        if (tombstone && !truly_empty && cand != nullptr && seen > static_cast<size_type>(capacity())) {
          truly_empty = true;  // one full cycle without an empty slot: settle for the remembered tombstone
        }
        if (!truly_empty) {
          // tombstone: remember the first one, keep looking for the key further down the cluster
          if (cand == nullptr) {
            cand     = table + index;
            cand_img = slot;
          }
          if constexpr (allows_duplicates) {
            // duplicates allowed: no need to look further
          } else {
            return false;
          }
        }
        value_type* target  = (cand != nullptr) ? cand : table + index;
        value_type expected = (cand != nullptr) ? cand_img : slot;
        while (true) {
          auto const res = claim(target, expected, val);
          if (res == insert_result::SUCCESS) {
            found    = target;
            inserted = true;
            return true;
          }
          if (res == insert_result::DUPLICATE) {
            found = target;
            return true;
          }
          // CONTINUE: `expected` now holds the current occupant
          if (classify_insert(key, key_of(expected)) == equal_result::AVAILABLE) { continue; }
          break;
        }
        if (target != table + index) {
          restart = true;  // lost a tombstone candidate: rescan from the start
          return true;
        }
        return false;  // lost this empty slot to another key: keep walking
      });

      if (!restart) { return {found, inserted}; }
    }
  }

  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename Value>
  __device__ bool scalar_insert(Value const& value) noexcept
  {
    auto const val = this->heterogeneous_value(value);
    auto const res = insert_driver<ChunkSlots, Policy>(
      val, [&](value_type* target, value_type& expected, auto const& v) {
        return try_claim(target, expected, native_value(v), key_of(v));
      });
    return res.second;
  }

  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename Value>
  __device__ thrust::pair<iterator, bool> scalar_insert_and_find(Value const& value) noexcept
  {
    auto const val = this->heterogeneous_value(value);
    auto const res = insert_driver<ChunkSlots, Policy>(
      val, [&](value_type* target, value_type& expected, auto const& v) {
        return try_claim(target, expected, native_value(v), key_of(v));
      });
    if constexpr (has_payload) {
      // a two-step writer (padded slots, insert_or_assign/apply) may not have published yet. Never
      // after our OWN successful insert: try_claim has stored the payload by then, and a payload equal
      // to the sentinel would spin forever
      if (!res.second) { wait_for_payload(res.first->second, empty_slot_.second); }
    }
    return {iterator{res.first}, res.second};
  }

  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename ProbeKey>
  [[nodiscard]] __device__ bool scalar_contains(ProbeKey const& key) const noexcept
  {
    bool present = false;
    walk<ChunkSlots, Policy>(make_cursor(key), [&](size_type, value_type slot) {
      auto const state = classify_lookup(key, key_of(slot));
      if (state == equal_result::EQUAL) {
        present = true;
        return true;
      }
      return state == equal_result::EMPTY;
    });
    return present;
  }

  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename ProbeKey>
  [[nodiscard]] __device__ const_iterator scalar_find(ProbeKey const& key) const noexcept
  {
    value_type* hit = nullptr;
    walk<ChunkSlots, Policy>(make_cursor(key), [&](size_type index, value_type slot) {
      auto const state = classify_lookup(key, key_of(slot));
      if (state == equal_result::EQUAL) {
        hit = slots() + index;
        return true;
      }
      return state == equal_result::EMPTY;
    });
    return hit ? const_iterator{hit} : this->end();
  }

  /// Number of entries matching `key` (0/1 unless duplicates are allowed).
  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename ProbeKey>
  [[nodiscard]] __device__ size_type scalar_count(ProbeKey const& key) const noexcept
  {
    if constexpr (!allows_duplicates) {
      return static_cast<size_type>(scalar_contains<ChunkSlots, Policy>(key));
    } else {
      size_type n = 0;
      walk<ChunkSlots, Policy>(make_cursor(key), [&](size_type, value_type slot) {
        auto const state = classify_lookup(key, key_of(slot));
        if (state == equal_result::EMPTY) { return true; }
        n += static_cast<size_type>(state == equal_result::EQUAL);
        return false;
      });
      return n;
    }
  }

  /// Tombstones the entry of `key`; true iff this call removed it.
  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename ProbeKey>
  __device__ bool scalar_erase(ProbeKey const& key) noexcept
  {
    return this->template erase_from<ChunkSlots, Policy>(make_cursor(key), key);
  }

  /// `scalar_erase` resumed at `start` (a cursor further down the key's probe sequence).
  template <int ChunkSlots, load_policy Policy, typename ProbeKey>
  __device__ bool erase_from(cursor start, ProbeKey const& key) noexcept
  {
    bool erased = false;
    walk<ChunkSlots, Policy>(start, [&](size_type index, value_type slot) {
      auto const state = classify_lookup(key, key_of(slot));
      if (state == equal_result::EMPTY) { return true; }
      if (state != equal_result::EQUAL) { return false; }
      erased = retire_observed(slots() + index, slot);
      return true;
    });
    return erased;
  }

  /// Calls `callback(slot content)` for every entry matching `key` (the reference hands the
  /// callback a copy of the slot, ref_impl.cuh:1290-1316).
  template <int ChunkSlots  = window_chunk_slots,
            load_policy Policy = load_policy::plain,
            typename ProbeKey,
            typename Callback>
  __device__ void scalar_for_each(ProbeKey const& key, Callback&& callback) const noexcept
  {
    walk<ChunkSlots, Policy>(make_cursor(key), [&](size_type, value_type slot) {
      auto const state = classify_lookup(key, key_of(slot));
      if (state == equal_result::EMPTY) { return true; }
      if (state == equal_result::EQUAL) {
        callback(slot);
        if constexpr (!allows_duplicates) { return true; }
      }
      return false;
    });
  }

  // ----------------------------------------------------------------------------------------------
  // tile-per-key operations: lane r owns window (base + r) mod N of the current bucket
  // ----------------------------------------------------------------------------------------------
  struct lane_view {
    equal_result state;   ///< first non-UNEQUAL state in this lane's window (scan order), or UNEQUAL
    std::int32_t index;   ///< its intra-window index, -1 if none
    value_type image;     ///< slot image at that index
    value_type* window;   ///< first slot of the lane's window
  };

  /// Loads this lane's window of the bucket at `base` and classifies it.
  template <bool ForInsert, typename Tile, typename ProbeKey>
  [[nodiscard]] __device__ lane_view tile_look(Tile const& tile,
                                               ProbeKey const& key,
                                               size_type base) const noexcept
  {
    auto const windows = static_cast<size_type>(storage_.window_extent());
    size_type w        = base + tile.thread_rank();
    if (w >= windows) { w -= windows; }
    auto const content = storage_[w];
    lane_view v{equal_result::UNEQUAL, -1, empty_slot_, slots() + w * window_size};
#pragma unroll
    for (int i = 0; i < window_size; ++i) {
      auto const s = ForInsert ? classify_insert(key, key_of(content[i]))
                               : classify_lookup(key, key_of(content[i]));
      bool const skip = (s == equal_result::UNEQUAL) ||
                        (ForInsert && allows_duplicates && s == equal_result::EQUAL);
      if (!skip && v.index < 0) {
        v.state = s;
        v.index = i;
        v.image = content[i];
      }
    }
    return v;
  }

  /// Bucket sequence of a key as seen by a tile: {base window, step in windows}.
  template <typename ProbeKey>
  [[nodiscard]] __device__ constexpr thrust::pair<size_type, size_type> tile_plan(
    ProbeKey const& key) const noexcept
  {
    auto const c = make_cursor(key);
    if constexpr (is_double_hashing) {
      return {c.base, c.step};
    } else {
      return {static_cast<size_type>(c.slot / window_size), static_cast<size_type>(cg_size)};
    }
  }

  __device__ constexpr size_type next_base(size_type base, size_type step) const noexcept
  {
    auto const windows = static_cast<size_type>(storage_.window_extent());
    base += step;
    if (base >= windows) { base -= windows; }
    return base;
  }

  /// Tile insertion driver (same contract as insert_driver, result replicated on every lane).
  template <typename Tile, typename Value, typename Claim>
  __device__ thrust::pair<value_type*, bool> tile_insert_driver(Tile const& tile,
                                                                Value const& val,
                                                                Claim&& claim) noexcept
  {
    auto const& key   = key_of(val);
    auto [base, step] = tile_plan(key);

    while (true) {
      auto const view = tile_look<true>(tile, key, base);
      auto* const mine = view.window + (view.index < 0 ? 0 : view.index);

      if constexpr (!allows_duplicates) {
        auto const equal_lanes = tile.ballot(view.state == equal_result::EQUAL);
        if (equal_lanes) {
          auto const src = __ffs(equal_lanes) - 1;
          auto const p   = tile.shfl(reinterpret_cast<std::intptr_t>(mine), src);
          return {reinterpret_cast<value_type*>(p), false};
        }
      }

      auto const open_lanes = tile.ballot(view.state == equal_result::AVAILABLE);
      if (open_lanes) {
        auto const src = __ffs(open_lanes) - 1;
        auto const p   = tile.shfl(reinterpret_cast<std::intptr_t>(mine), src);
        auto status    = insert_result::CONTINUE;
        if (tile.thread_rank() == static_cast<unsigned>(src)) {
          auto expected = view.image;
          status        = claim(mine, expected, val);
        }
        status = static_cast<insert_result>(tile.shfl(static_cast<std::int32_t>(status), src));
        if (status == insert_result::SUCCESS) { return {reinterpret_cast<value_type*>(p), true}; }
        if (status == insert_result::DUPLICATE) { return {reinterpret_cast<value_type*>(p), false}; }
        continue;  // bucket changed under us: look at it again
      }
      base = next_base(base, step);
    }
  }

  template <typename Tile, typename Value>
  __device__ bool tile_insert(Tile const& tile, Value const& value) noexcept
  {
    auto const val = this->heterogeneous_value(value);
    return tile_insert_driver(tile, val, [&](value_type* t, value_type& e, auto const& v) {
             return try_claim(t, e, native_value(v), key_of(v));
           }).second;
  }

  template <typename Tile, typename Value>
  __device__ thrust::pair<iterator, bool> tile_insert_and_find(Tile const& tile,
                                                               Value const& value) noexcept
  {
    auto const val = this->heterogeneous_value(value);
    auto const res = tile_insert_driver(tile, val, [&](value_type* t, value_type& e, auto const& v) {
      return try_claim(t, e, native_value(v), key_of(v));
    });
    if constexpr (has_payload) {
      if (!res.second) {  // see scalar_insert_and_find: never after the tile's own successful insert
        if (tile.thread_rank() == 0) { wait_for_payload(res.first->second, empty_slot_.second); }
        tile.sync();
      }
    }
    return {iterator{res.first}, res.second};
  }

  template <typename Tile, typename ProbeKey>
  [[nodiscard]] __device__ const_iterator tile_find(Tile const& tile,
                                                    ProbeKey const& key) const noexcept
  {
    auto [base, step] = tile_plan(key);
    while (true) {
      auto const view = tile_look<false>(tile, key, base);
      auto const hits = tile.ballot(view.state == equal_result::EQUAL);
      if (hits) {
        auto const src = __ffs(hits) - 1;
        auto const p   = tile.shfl(reinterpret_cast<std::intptr_t>(view.window + view.index), src);
        return const_iterator{reinterpret_cast<value_type*>(p)};
      }
      if (tile.any(view.state == equal_result::EMPTY)) { return this->end(); }
      base = next_base(base, step);
    }
  }

  template <typename Tile, typename ProbeKey>
  [[nodiscard]] __device__ bool tile_contains(Tile const& tile, ProbeKey const& key) const noexcept
  {
    auto [base, step] = tile_plan(key);
    while (true) {
      auto const view = tile_look<false>(tile, key, base);
      if (tile.any(view.state == equal_result::EQUAL)) { return true; }
      if (tile.any(view.state == equal_result::EMPTY)) { return false; }
      base = next_base(base, step);
    }
  }

  template <typename Tile, typename ProbeKey>
  __device__ bool tile_erase(Tile const& tile, ProbeKey const& key) noexcept
  {
    auto [base, step] = tile_plan(key);
    while (true) {
      auto const view = tile_look<false>(tile, key, base);
      auto const hits = tile.ballot(view.state == equal_result::EQUAL);
      if (hits) {
        auto const src = __ffs(hits) - 1;
        bool done      = false;
        if (tile.thread_rank() == static_cast<unsigned>(src)) {
          done = retire_slot(view.window + view.index, key_of(view.image));
        }
        return tile.shfl(done, src);
      }
      if (tile.any(view.state == equal_result::EMPTY)) { return false; }
      base = next_base(base, step);
    }
  }

  /// Per-tile count: every lane counts the matches in its own windows and returns THAT number, like
  /// the reference ("occurrences found by the current thread", ref_impl.cuh:860-892); callers sum
  /// over the tile.
  template <typename Tile, typename ProbeKey>
  [[nodiscard]] __device__ size_type tile_count(Tile const& tile,
                                                ProbeKey const& key) const noexcept
  {
    auto [base, step]  = tile_plan(key);
    auto const windows = static_cast<size_type>(storage_.window_extent());
    size_type mine     = 0;
    while (true) {
      size_type w = base + tile.thread_rank();
      if (w >= windows) { w -= windows; }
      auto const content = storage_[w];
      bool saw_empty     = false;
#pragma unroll
      for (int i = 0; i < window_size; ++i) {
        auto const s = classify_lookup(key, key_of(content[i]));
        if (s == equal_result::EMPTY) { saw_empty = true; }
        if (!saw_empty && s == equal_result::EQUAL) { ++mine; }
      }
      if (tile.any(saw_empty)) { break; }
      base = next_base(base, step);
    }
    return mine;
  }

  /// Per-tile for_each: every lane runs `callback(slot content)` on the matches of its own window
  /// (up to the first empty slot of that window); `sync(tile)` runs after every probe step
  /// (reference ref_impl.cuh:1334-1440).
  template <typename Tile, typename ProbeKey, typename Callback, typename Sync>
  __device__ void tile_for_each(Tile const& tile,
                                ProbeKey const& key,
                                Callback&& callback,
                                Sync&& sync) const noexcept
  {
    auto [base, step]  = tile_plan(key);
    auto const windows = static_cast<size_type>(storage_.window_extent());
    while (true) {
      size_type w = base + tile.thread_rank();
      if (w >= windows) { w -= windows; }
      auto const content = storage_[w];
      bool saw_empty     = false;
      bool saw_equal     = false;
#pragma unroll
      for (int i = 0; i < window_size; ++i) {
        if (!saw_empty) {
          auto const s = classify_lookup(key, key_of(content[i]));
          if (s == equal_result::EMPTY) {
            saw_empty = true;
          } else if (s == equal_result::EQUAL) {
            callback(content[i]);
            saw_equal = true;
          }
        }
      }
      sync(tile);
      if (tile.any(saw_empty)) { return; }
      if constexpr (!allows_duplicates) {
        if (tile.any(saw_equal)) { return; }
      }
      base = next_base(base, step);
    }
  }

  // ----------------------------------------------------------------------------------------------
  // payload-side helpers used by the map upserts
  // ----------------------------------------------------------------------------------------------
  /// Claims only the key half of a 16-byte slot (used when the payload must be combined in place).
  [[nodiscard]] __device__ insert_result try_claim_key(value_type* address,
                                                       key_type& expected_key,
                                                       key_type const& desired_key) const noexcept
  {
    return this->try_claim_key(address, expected_key, desired_key, desired_key);
  }

  template <typename ProbeKey>
  [[nodiscard]] __device__ insert_result try_claim_key(value_type* address,
                                                       key_type& expected_key,
                                                       key_type const& desired_key,
                                                       ProbeKey const& probe_key) const noexcept
  {
    static_assert(has_payload);
    cuda::atomic_ref<key_type, Scope> key_ref{address->first};
    if (key_ref.compare_exchange_strong(expected_key, desired_key, cuda::memory_order_relaxed)) {
      return insert_result::SUCCESS;
    }
    if (!same_bits(expected_key, key_of(empty_slot_)) && !same_bits(expected_key, erased_key_) &&
        eq_(probe_key, expected_key)) {
      return insert_result::DUPLICATE;
    }
    return insert_result::CONTINUE;
  }

 public:
  /// key -> erased sentinel (payload reset to the empty payload); true iff we made the transition.
  /// Same, for a caller that already holds the slot image it observed: packed slots go straight to
  /// the one CAS (no second load of the slot; the claim and the payload reset are one 32/64/128-bit
  /// atomic), the rest is `retire_slot`.
  __device__ bool retire_observed(value_type* address, value_type const& observed) noexcept
  {
    if constexpr (single_cas) {
      value_type expected = observed;
      auto const key      = key_of(observed);
      while (eq_(key, key_of(expected)) && !same_bits(key_of(expected), erased_key_)) {
        auto const seen = cas_slot<Scope>(address, expected, erased_slot_sentinel());
        if (same_bits(seen, expected)) { return true; }
        expected = seen;
      }
      return false;
    } else {
      return this->retire_slot(address, key_of(observed));
    }
  }

  __device__ bool retire_slot(value_type* address, key_type observed_key) noexcept
  {
    if constexpr (has_payload && sizeof(value_type) > 8 && !single_cas) {
      cuda::atomic_ref<key_type, Scope> key_ref{address->first};
      if (key_ref.compare_exchange_strong(observed_key, erased_key_, cuda::memory_order_relaxed)) {
        cuda::atomic_ref<decltype(address->second), Scope> payload_ref{address->second};
        payload_ref.store(empty_slot_.second, cuda::memory_order_relaxed);
        return true;
      }
      return false;
    } else if constexpr (single_cas) {
      // packed slot: swap the whole image; the payload may be changing, so retry while the key stays
      value_type expected = *address;
      while (eq_(observed_key, key_of(expected)) && !same_bits(key_of(expected), erased_key_)) {
        auto const seen = cas_slot<Scope>(address, expected, erased_slot_sentinel());
        if (same_bits(seen, expected)) { return true; }
        expected = seen;
      }
      return false;
    } else {
      cuda::atomic_ref<key_type, Scope> key_ref{*reinterpret_cast<key_type*>(address)};
      return key_ref.compare_exchange_strong(
        observed_key, erased_key_, cuda::memory_order_relaxed);
    }
  }

 private:

  /// Chunks wider than one slot need the vector path; single odd-sized slots use a plain copy.
  template <int ChunkSlots, load_policy Policy>
  __host__ __device__ static constexpr load_policy chunk_policy() noexcept
  {
    return Policy;
  }

  using step_modulus_helper = step_modulus<is_double_hashing, extent_type, size_type, cg_size>;
  using step_modulus_type   = typename step_modulus_helper::type;

  value_type empty_slot_;
  key_type erased_key_;
  key_equal eq_;
  probing_scheme_type scheme_;
  storage_ref_type storage_;
  step_modulus_type step_modulus_;

  template <typename K, cuda::thread_scope S, typename E, typename P, typename R, bool D>
  friend class probe_engine;
};

}  // namespace b200
}  // namespace cuco
