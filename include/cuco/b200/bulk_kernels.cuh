// Bulk kernels of the open-addressing hot path, written for sm_100a.
//
// They replace the reference's one-key-per-cooperative-group kernels
// (include/cuco/detail/open_addressing/kernels.cuh:64-667 and detail/static_map/kernels.cuh:52-264:
// insert_if_n, insert_and_find, find, contains_if_n, erase, size, insert_or_assign,
// insert_or_apply). Design:
//
//  * thread-per-key, several keys per thread. A CTA takes a tile of BlockSize*KeysPerThread inputs;
//    thread t owns inputs t, t+BlockSize, ... so every input load and output store of a warp is one
//    contiguous segment. Each thread keeps KeysPerThread independent probe cursors and runs them in
//    rounds: issue every pending chunk load (or slot CAS), then consume them all. The random-sector
//    latency of one key is overlapped with the others' instead of being exposed per probe step.
//  * the table is read in sector chunks: one 256-bit load brings the whole 32-byte DRAM sector the
//    probe lands in (two 16-byte slots, four 8-byte, eight 4-byte) and is scanned from registers in
//    probe order with early exit. Lookups use the non-coherent path and do not allocate in L1;
//    mutating kernels read through L2 (relaxed.gpu) so a retry always observes the competing write.
//  * a slot is claimed with one CAS on the whole slot image (32/64/128 bit). With `CasFirst` the
//    first probe of an insert is the CAS itself - no load - which is the common case at load
//    factors <= 0.5 where the home slot is usually free.
//  * one CTA per tile by default (a persistent grid-stride launch is a tuning option);
//  * all input loads of a thread are posted before the first is consumed;
//  * mutations of tables far beyond L2 can be routed by table region first ("L2-blocked mutation"
//    below), and hash-partitioned multi-GPU tables route by owner rank as well ("Exchange path").
//
// Kernels with `generic_` prefix are the one-key-per-thread fallbacks used when the fast path's
// preconditions do not hold (tombstones configured, padded slots, storage not 32-byte aligned).
#pragma once

#include <cuco/b200/probe_engine.cuh>
#include <cuco/detail/utility/cuda.cuh>

#include <cuda/atomic>
#include <cuda/std/iterator>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/iterator_traits.h>

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace cuco::b200 {

using cuco::detail::index_type;

/// Identity predicate for the un-stencilled entry points.
/// The stream-ordered pool all containers (and the whole-table passes of table_scan.cuh) on the current
/// device draw their per-call scratch from:
/// one per device for the life of the process (a pool per container costs a 2 MiB granule and a
/// driver round trip each - the reference's shared_memory_test builds 1000 maps). Freed blocks stay
/// cached up to CUCO_B200_SCRATCH_KEEP_MIB (default 4096) so that back-to-back bulk calls do not
/// pay cudaMalloc; anything above goes back to the driver at the next synchronisation.
[[nodiscard]] inline cudaMemPool_t device_scratch_pool() noexcept
{
  constexpr int max_devices = 64;
  static std::mutex guard;
  static cudaMemPool_t pools[max_devices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= max_devices) { return nullptr; }
  std::lock_guard<std::mutex> lock{guard};
  if (pools[dev] == nullptr) {
    cudaMemPoolProps props{};
    props.allocType     = cudaMemAllocationTypePinned;
    props.handleTypes   = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id   = dev;
    if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) {
      cudaGetLastError();
      pools[dev] = nullptr;
      return nullptr;
    }
    std::uint64_t keep = std::uint64_t{4096} << 20;
    if (char const* s = std::getenv("CUCO_B200_SCRATCH_KEEP_MIB")) {
      keep = static_cast<std::uint64_t>(std::max(0, std::atoi(s))) << 20;
    }
    cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return pools[dev];
}


struct always_true {
  template <typename T>
  __host__ __device__ constexpr bool operator()(T const&) const noexcept
  {
    return true;
  }
};

/// Reads element `i` of an input range; raw pointers take the streaming (read-once) path.
template <typename It>
__device__ __forceinline__ auto read_input(It first, index_type i)
{
  using value_t = typename cuda::std::iterator_traits<It>::value_type;
  if constexpr (cuda::std::is_pointer_v<It>) {
    return load_streaming<cuda::std::remove_cv_t<value_t>>(first + i);
  } else {
    return static_cast<value_t>(*(first + i));
  }
}

/// Register storage for a value of a type that may lack a default constructor (probe keys and
/// input elements are user types; they only need to be trivially copyable, like kernel arguments).
template <typename T>
union uninitialized {
  T value;
  __device__ uninitialized() {}
};

/// Adds a per-thread count into a global counter: warp shuffle reduce, one atomic per warp.
template <typename Counter>
__device__ __forceinline__ void accumulate_count(Counter* counter, unsigned long long mine)
{
  auto const warp = cg::tiled_partition<32>(cg::this_thread_block());
  auto const sum  = cg::reduce(warp, mine, cg::plus<unsigned long long>());
  if (warp.thread_rank() == 0 && sum != 0) {
    cuda::atomic_ref<Counter, cuda::thread_scope_device> ref{*counter};
    ref.fetch_add(static_cast<Counter>(sum), cuda::memory_order_relaxed);
  }
}

// =================================================================================================
// lookups: find / contains / contains_if
// =================================================================================================

/// Output policy of `find`: payload (map) or stored key (set), sentinel on a miss
/// (reference: open_addressing/kernels.cuh:358-372).
template <typename Engine>
struct emit_found {
  using slot_type = typename Engine::value_type;
  slot_type empty_slot;

  __device__ auto hit(slot_type const& slot) const noexcept
  {
    if constexpr (Engine::has_payload) {
      return slot.second;
    } else {
      return slot;
    }
  }
  __device__ auto miss() const noexcept { return hit(empty_slot); }
};

/// Output policy of `contains`.
struct emit_present {
  template <typename Slot>
  __device__ bool hit(Slot const&) const noexcept
  {
    return true;
  }
  __device__ bool miss() const noexcept { return false; }
};

template <int BlockSize,
          int KeysPerThread,
          int ChunkSlots,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename OutputIt,
          typename Engine,
          typename Emit>
CUCO_KERNEL __launch_bounds__(BlockSize) void lookup_kernel(InputIt first,
                                                            index_type n,
                                                            StencilIt stencil,
                                                            Predicate pred,
                                                            OutputIt out,
                                                            Engine engine,
                                                            Emit emit)
{
  using slot_type  = typename Engine::value_type;
  using cursor     = typename Engine::cursor;
  using probe_type = decltype(read_input(first, index_type{0}));
  constexpr index_type tile = index_type{BlockSize} * KeysPerThread;
  constexpr auto policy     = load_policy::readonly;

  for (index_type tile_base = index_type{blockIdx.x} * tile; tile_base < n;
       tile_base += index_type{gridDim.x} * tile) {
    uninitialized<probe_type> key[KeysPerThread];
    cursor cur[KeysPerThread];
    unsigned pending = 0;

    // all input loads are posted before the first one is consumed (the loads are volatile asm:
    // hashing key j inside the same loop would serialise the DRAM round trips)
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
      if (idx < n) {
        if (pred(*(stencil + idx))) {
          key[j].value = read_input(first, idx);
          pending |= 1u << j;
        } else {
          *(out + idx) = emit.miss();
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (pending & (1u << j)) { cur[j] = engine.make_cursor(key[j].value); }
    }

    while (pending) {
      raw_chunk<ChunkSlots * Engine::slot_bytes> raw[KeysPerThread];
#pragma unroll
      for (int j = 0; j < KeysPerThread; ++j) {
        if (pending & (1u << j)) { raw[j] = engine.template load_chunk<ChunkSlots, policy>(cur[j]); }
      }
#pragma unroll
      for (int j = 0; j < KeysPerThread; ++j) {
        if (pending & (1u << j)) {
          index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
          int const begin_off =
            static_cast<int>(cur[j].slot - Engine::template chunk_begin<ChunkSlots>(cur[j]));
          int const valid = engine.template chunk_valid<ChunkSlots>(cur[j]);
          bool done       = false;
#pragma unroll
          for (int i = 0; i < ChunkSlots; ++i) {
            if (!done && i >= begin_off && i < begin_off + valid) {
              auto const slot  = chunk_slot<slot_type>(raw[j], i);
              auto const state = engine.classify_lookup(key[j].value, Engine::key_of(slot));
              if (state == equal_result::EQUAL) {
                *(out + idx) = emit.hit(slot);
                done         = true;
              } else if (state == equal_result::EMPTY) {
                *(out + idx) = emit.miss();
                done         = true;
              }
            }
          }
          if (done) {
            pending &= ~(1u << j);
          } else {
            engine.advance(cur[j], valid);
          }
        }
      }
    }
  }
}

/// Fallback lookup: one key per thread through the engine's safe (window-chunk, plain load) path.
template <int BlockSize,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename OutputIt,
          typename Engine,
          typename Emit>
CUCO_KERNEL __launch_bounds__(BlockSize) void generic_lookup_kernel(InputIt first,
                                                                    index_type n,
                                                                    StencilIt stencil,
                                                                    Predicate pred,
                                                                    OutputIt out,
                                                                    Engine engine,
                                                                    Emit emit)
{
  for (index_type idx = cuco::detail::global_thread_id(); idx < n;
       idx += cuco::detail::grid_stride()) {
    if (pred(*(stencil + idx))) {
      auto const key = read_input(first, idx);
      auto const it  = engine.scalar_find(key);
      if (it == engine.end()) {
        *(out + idx) = emit.miss();
      } else {
        *(out + idx) = emit.hit(*it);
      }
    } else {
      *(out + idx) = emit.miss();
    }
  }
}

// =================================================================================================
// mutations: insert / insert_if / insert_and_find / insert_or_assign / insert_or_apply
// =================================================================================================

/// What a mutation does once it knows whether the key was new. `Engine` gives the slot types.
/// Every action sees: the input index, the slot address, the slot image now stored there.
struct action_insert {
  static constexpr bool key_then_apply = false;
  static constexpr bool blockable      = true;  ///< no per-element output: input order is free
  template <typename Engine, typename Slot>
  __device__ void on_new(Engine const&, index_type, Slot*, Slot const&) const noexcept
  {
  }
  template <typename Engine, typename Slot>
  __device__ void on_present(Engine const&, index_type, Slot*, Slot const&, Slot const&) const noexcept
  {
  }
};

/// insert_and_find: report the resident payload/key and whether this element created the entry
/// (reference: open_addressing/kernels.cuh:504-564).
template <typename FoundIt, typename InsertedIt>
struct action_insert_and_find {
  static constexpr bool key_then_apply = false;
  static constexpr bool blockable      = false;  ///< writes outputs at the element's input index
  FoundIt found;
  InsertedIt inserted;

  template <typename Engine, typename Slot>
  __device__ void emit(Engine const& engine, index_type idx, Slot* address, Slot image, bool is_new) const
  {
    if constexpr (Engine::has_payload) {
      if (!is_new && same_bits(image.second, engine.empty_slot_sentinel().second)) {
        // written by a two-step writer that has not published the payload yet
        engine.wait_for_payload(address->second, engine.empty_slot_sentinel().second);
        image.second = address->second;
      }
      *(found + idx) = image.second;
    } else {
      *(found + idx) = image;
    }
    *(inserted + idx) = is_new;
  }
  template <typename Engine, typename Slot>
  __device__ void on_new(Engine const& e, index_type idx, Slot* address, Slot const& desired) const
  {
    emit(e, idx, address, desired, true);
  }
  template <typename Engine, typename Slot>
  __device__ void on_present(
    Engine const& e, index_type idx, Slot* address, Slot const& resident, Slot const&) const
  {
    emit(e, idx, address, resident, false);
  }
};

/// insert_or_assign: last writer wins (reference: static_map_ref.inl:486-620).
struct action_assign {
  static constexpr bool key_then_apply = false;
  static constexpr bool blockable      = true;
  template <typename Engine, typename Slot>
  __device__ void on_new(Engine const&, index_type, Slot*, Slot const&) const noexcept
  {
  }
  template <typename Engine, typename Slot>
  __device__ void on_present(
    Engine const&, index_type, Slot* address, Slot const&, Slot const& desired) const noexcept
  {
    using mapped = decltype(address->second);
    cuda::atomic_ref<mapped, Engine::thread_scope> ref{address->second};
    ref.store(desired.second, cuda::memory_order_relaxed);
  }
};

/// insert_or_apply: first arrival stores its value, later arrivals combine with `op`
/// (reference: static_map_ref.inl:850-1052). With `DirectApply` on 16-byte slots the first arrival
/// claims the key only and combines onto the sentinel payload, exactly like the reference does when
/// `init == empty_value_sentinel`.
template <typename Op, bool DirectApply>
struct action_apply {
  static constexpr bool key_then_apply = DirectApply;
  static constexpr bool blockable      = true;
  Op op;

  template <typename Engine, typename Slot>
  __device__ void on_new(Engine const&, index_type, Slot* address, Slot const& desired) const
  {
    if constexpr (DirectApply && sizeof(Slot) > 8) {
      using mapped = decltype(address->second);
      op(cuda::atomic_ref<mapped, Engine::thread_scope>{address->second}, desired.second);
    }
  }
  template <typename Engine, typename Slot>
  __device__ void on_present(
    Engine const&, index_type, Slot* address, Slot const&, Slot const& desired) const
  {
    using mapped = decltype(address->second);
    op(cuda::atomic_ref<mapped, Engine::thread_scope>{address->second}, desired.second);
  }
};

/// Rare path of the fast mutate kernel (not marked noinline: ptxas 12.9 crashes on that): the
/// slot is claimable but its image differs from the canonical empty slot (foreign payload bits), so
/// the general driver - which compares against what it actually observes - finishes the key.
template <int ChunkSlots, load_policy Policy, typename Engine, typename Input, typename Action>
__device__ bool mutate_slow_path(
  Engine& engine, Input const& val, index_type idx, Action& action)
{
  using slot_type    = typename Engine::value_type;
  auto const desired = engine.native_value(val);
  auto const res     = engine.template insert_driver<ChunkSlots, Policy>(
    val, [&](slot_type* t, slot_type& e, auto const& v) { return engine.try_claim(t, e, desired, Engine::key_of(v)); });
  if (res.second) {
    action.on_new(engine, idx, res.first, desired);
  } else {
    action.on_present(engine, idx, res.first, *res.first, desired);
  }
  return res.second;
}

/// One tile of the fast mutation path: element j of thread t is `tile_base + j * BlockSize + t`, live
/// while that index is below `limit` and its stencil passes. Returns the number of new keys.
template <int BlockSize,
          int KeysPerThread,
          int ChunkSlots,
          bool CasFirst,
          load_policy Policy,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename Engine,
          typename Action>
__device__ __forceinline__ unsigned long long mutate_tile(InputIt first,
                                                          index_type tile_base,
                                                          index_type limit,
                                                          StencilIt stencil,
                                                          Predicate& pred,
                                                          Engine& engine,
                                                          Action& action)
{
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  using cursor    = typename Engine::cursor;
  // inputs stay in their own (possibly heterogeneous) type: hashing and the key predicate are
  // invoked on it, exactly where the reference invokes them; the slot image is derived on demand
  using input_type = decltype(engine.heterogeneous_value(read_input(first, index_type{0})));
  static_assert(Engine::single_cas, "fast mutate path needs one-shot claimable slots");
  constexpr auto policy = Policy;
  // claim only the key half, then combine the payload in place (insert_or_apply direct mode)
  constexpr bool key_only_claim = Action::key_then_apply && sizeof(slot_type) > 8;

  auto* const table       = engine.slots();
  auto const empty_slot   = engine.empty_slot_sentinel();
  index_type const n      = limit;
  unsigned long long mine = 0;
  {
    uninitialized<input_type> val[KeysPerThread];
    cursor cur[KeysPerThread];
    unsigned pending = 0;  // bit j: key j still in flight
    unsigned claim   = 0;  // bit j: next action of key j is a CAS at cur[j].slot (else a chunk load)

    // post every input load before consuming the first (see lookup_kernel)
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
      if (idx < n && pred(*(stencil + idx))) {
        val[j].value = engine.heterogeneous_value(read_input(first, idx));
        pending |= 1u << j;
        if constexpr (CasFirst) { claim |= 1u << j; }
      }
    }
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (pending & (1u << j)) { cur[j] = engine.make_cursor(Engine::key_of(val[j].value)); }
    }

    while (pending) {
      raw_chunk<ChunkSlots * Engine::slot_bytes> raw[KeysPerThread];
      slot_type seen[KeysPerThread];

      // ---- issue: every pending key posts its next memory operation ----
#pragma unroll
      for (int j = 0; j < KeysPerThread; ++j) {
        if (pending & (1u << j)) {
          if (claim & (1u << j)) {
            if constexpr (key_only_claim) {
              key_type expected_key = Engine::key_of(empty_slot);
              cuda::atomic_ref<key_type, Engine::thread_scope> key_ref{(table + cur[j].slot)->first};
              key_ref.compare_exchange_strong(expected_key,
                                              static_cast<key_type>(Engine::key_of(val[j].value)),
                                              cuda::memory_order_relaxed);
              seen[j]       = empty_slot;
              seen[j].first = expected_key;
            } else {
              seen[j] = cas_slot<Engine::thread_scope>(
                table + cur[j].slot, empty_slot, engine.native_value(val[j].value));
            }
          } else {
            raw[j] = engine.template load_chunk<ChunkSlots, policy>(cur[j]);
          }
        }
      }

      // ---- consume ----
#pragma unroll
      for (int j = 0; j < KeysPerThread; ++j) {
        if (!(pending & (1u << j))) { continue; }
        index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
        auto const& key      = Engine::key_of(val[j].value);
        auto const desired   = engine.native_value(val[j].value);

        if (claim & (1u << j)) {
          auto* const address = table + cur[j].slot;
          bool const won      = key_only_claim
                                  ? same_bits(Engine::key_of(seen[j]), Engine::key_of(empty_slot))
                                  : same_bits(seen[j], empty_slot);
          if (won) {
            action.on_new(engine, idx, address, desired);
            ++mine;
            pending &= ~(1u << j);
          } else {
            auto const state = engine.classify_insert(key, Engine::key_of(seen[j]));
            if (state == equal_result::EQUAL) {
              action.on_present(engine, idx, address, seen[j], desired);
              pending &= ~(1u << j);
            } else if (state == equal_result::AVAILABLE) {
              // still claimable but not bit-identical to the empty image (foreign payload)
              mine += mutate_slow_path<ChunkSlots, policy>(engine, val[j].value, idx, action);
              pending &= ~(1u << j);
            } else {
              // somebody else's key lives here now: move on, next round loads
              engine.advance(cur[j], 1);
              claim &= ~(1u << j);
            }
          }
        } else {
          int const begin_off =
            static_cast<int>(cur[j].slot - Engine::template chunk_begin<ChunkSlots>(cur[j]));
          int const valid = engine.template chunk_valid<ChunkSlots>(cur[j]);
          int consumed    = valid;
          bool resolved   = false;
#pragma unroll
          for (int i = 0; i < ChunkSlots; ++i) {
            if (!resolved && consumed == valid && i >= begin_off && i < begin_off + valid) {
              auto const slot  = chunk_slot<slot_type>(raw[j], i);
              auto const state = engine.classify_insert(key, Engine::key_of(slot));
              if (state == equal_result::EQUAL) {
                action.on_present(engine, idx, table + (cur[j].slot + (i - begin_off)), slot, desired);
                resolved = true;
              } else if (state == equal_result::AVAILABLE) {
                consumed = i - begin_off;  // stop in front of this slot and claim it next round
              }
            }
          }
          if (resolved) {
            pending &= ~(1u << j);
          } else {
            if (consumed < valid) { claim |= 1u << j; }
            if (consumed > 0) { engine.advance(cur[j], consumed); }
          }
        }
      }
    }
  }
  return mine;
}

template <int BlockSize,
          int KeysPerThread,
          int ChunkSlots,
          bool CasFirst,
          bool Counted,
          load_policy Policy,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename Counter,
          typename Engine,
          typename Action>
CUCO_KERNEL __launch_bounds__(BlockSize) void mutate_kernel(InputIt first,
                                                            index_type n,
                                                            StencilIt stencil,
                                                            Predicate pred,
                                                            Counter* num_new,
                                                            Engine engine,
                                                            Action action)
{
  constexpr index_type tile = index_type{BlockSize} * KeysPerThread;
  unsigned long long mine   = 0;
  for (index_type tile_base = index_type{blockIdx.x} * tile; tile_base < n;
       tile_base += index_type{gridDim.x} * tile) {
    mine += mutate_tile<BlockSize, KeysPerThread, ChunkSlots, CasFirst, Policy>(
      first, tile_base, n, stencil, pred, engine, action);
  }
  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

/// Fallback mutation: one key per thread through the engine's general drivers (tombstone aware,
/// two-step claims for padded slots).
template <int BlockSize,
          bool Counted,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename Counter,
          typename Engine,
          typename Action>
CUCO_KERNEL __launch_bounds__(BlockSize) void generic_mutate_kernel(InputIt first,
                                                                    index_type n,
                                                                    StencilIt stencil,
                                                                    Predicate pred,
                                                                    Counter* num_new,
                                                                    Engine engine,
                                                                    Action action)
{
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  unsigned long long mine = 0;

  for (index_type idx = cuco::detail::global_thread_id(); idx < n;
       idx += cuco::detail::grid_stride()) {
    if (!pred(*(stencil + idx))) { continue; }
    auto const val     = engine.heterogeneous_value(read_input(first, idx));
    auto const desired = engine.native_value(val);
    constexpr bool key_only_claim = Action::key_then_apply && sizeof(slot_type) > 8;

    auto const res = engine.template insert_driver<Engine::window_chunk_slots, load_policy::plain>(
      val, [&](slot_type* target, slot_type& expected, auto const& v) {
        if constexpr (key_only_claim) {
          key_type expected_key = Engine::key_of(expected);
          auto const r = engine.try_claim_key(target, expected_key, desired.first, Engine::key_of(v));
          expected.first = expected_key;
          return r;
        } else {
          return engine.try_claim(target, expected, desired, Engine::key_of(v));
        }
      });
    if (res.second) {
      action.on_new(engine, idx, res.first, desired);
      ++mine;
    } else {
      if constexpr (Engine::has_payload && !cuda::std::is_same_v<Action, action_insert>) {
        // two-step writers publish the payload after the key: wait before reading or combining,
        // unless the first arrival itself combines onto the sentinel (direct apply)
        if constexpr (!key_only_claim && !Engine::single_cas) {
          if constexpr (!cuda::std::is_same_v<Action, action_assign>) {
            engine.wait_for_payload(res.first->second, engine.empty_slot_sentinel().second);
          }
        }
      }
      action.on_present(engine, idx, res.first, *res.first, desired);
    }
  }

  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

// =================================================================================================
// L2-blocked mutation: route the batch by table region, then probe region by region
// =================================================================================================
//
// Measured on B200 (tools/microbench*.cu, profiles/r01_hardware_probes.md):
//   * an L2 miss to a random table sector moves a whole 128-byte line from DRAM (3.9 sectors per
//     32-byte read), so random probing tops out near 49 G reads/s and ~22 G read+CAS/s chip-wide;
//   * the same number of accesses confined to one 16-32 MB slice of the table at a time run at
//     174 G reads/s and 55-60 G read+CAS/s: the slice is fetched line by line exactly once, the
//     other three sectors of every line are hits, and dirty sectors leave L2 once.
// A mutation batch that touches a table far larger than L2 about once per sector is therefore ~3x
// cheaper when it arrives grouped by table region, and a shared-memory staged partition of the
// batch costs about one streaming read + write of it.
//
// Pass 1 (`route_kernel`): every CTA takes a tile of the batch, ranks its elements per region with
// shared-memory counters, reserves a run in each region's segment with ONE global atomic per
// (tile, region), stages the tile in shared memory grouped by region and copies it out in runs that
// are contiguous in both shared and global memory (full 128-byte lines instead of scattered 16-byte
// stores). Segments have a fixed capacity (expected load + slack); an element whose segment is
// full is finished right there through the general driver, which keeps the path correct for
// arbitrarily skewed inputs without a counting pre-pass.
//
// Pass 2 (`blocked_mutate_kernel`): grid (tiles per segment, regions), so CTAs are dispatched
// region by region; each CTA runs the ordinary tile body (`mutate_tile`) on its slice of one
// segment, after posting one bulk L2 prefetch of its share of the NEXT region's table slice
// (`cp.async.bulk.prefetch.L2`), which turns that region's first touches into L2 hits.

/// Maps a home slot to its region: floor(slot * num_regions / capacity) by multiply-high.
struct region_map {
  std::uint64_t scale;  ///< ceil(2^64 * num_regions / slots) clamped so the result < num_regions
  std::uint32_t num_regions;
  std::uint64_t first_slot = 0;  ///< the regions divide the slice [first_slot, first_slot + slots) of the
                                 ///< table (the whole table unless the batch is known to be confined)

  __host__ __device__ std::uint32_t operator()(std::uint64_t slot) const noexcept
  {
    slot -= first_slot;  // a slot in front of the slice wraps to a huge value and lands in the last region
#if defined(__CUDA_ARCH__)
    auto const r = static_cast<std::uint32_t>(__umul64hi(slot, scale));
#else
    auto const r = static_cast<std::uint32_t>((static_cast<unsigned __int128>(slot) * scale) >> 64);
#endif
    return r < num_regions ? r : num_regions - 1;
  }

  /// First slot of region `r` (host side: exact inverse of operator()).
  [[nodiscard]] std::uint64_t region_begin(std::uint32_t r) const noexcept
  {
    unsigned __int128 const numerator = static_cast<unsigned __int128>(r) << 64;
    return first_slot + static_cast<std::uint64_t>((numerator + scale - 1) / scale);
  }

  /// Map of `num_regions` regions over `slots` slots starting at `first_slot`.
  [[nodiscard]] static region_map over(std::uint64_t slots, std::uint32_t num_regions, std::uint64_t first_slot = 0) noexcept
  {
    unsigned __int128 const scaled = (static_cast<unsigned __int128>(num_regions) << 64) / (slots ? slots : 1);
    return region_map{static_cast<std::uint64_t>(scaled) + 1, num_regions, first_slot};
  }
};

constexpr int route_block_size       = 512;  ///< more warps per CTA hide the input-load latency
constexpr int route_items_per_thread = 8;    ///< tile = 4096 elements
constexpr int route_max_regions      = 1024;

/// Dynamic shared memory of `route_kernel` for a slot type.
template <int BlockSize, typename Slot>
constexpr std::size_t route_smem_bytes() noexcept
{
  return std::size_t{BlockSize} * route_items_per_thread * (sizeof(Slot) + sizeof(std::uint16_t));
}

/// Predicate over a virtual index space made of fixed-capacity segments: element i exists iff its
/// offset inside segment i / capacity is below that segment's fill count (exchange buffers).
struct segment_live {
  unsigned int const* counts;
  std::uint32_t segment_capacity;

  __device__ bool operator()(index_type i) const noexcept
  {
    auto const segment = static_cast<std::uint32_t>(i / segment_capacity);
    auto const local   = static_cast<std::uint32_t>(i - index_type{segment} * segment_capacity);
    return local < counts[segment];
  }
};

template <int BlockSize,
          int ChunkSlots,
          bool Counted,
          typename InputIt,
          typename StencilIt,
          typename Predicate,
          typename Counter,
          typename Engine,
          typename Action>
CUCO_KERNEL __launch_bounds__(BlockSize) void route_kernel(InputIt first,
                                                           index_type n,
                                                           StencilIt stencil,
                                                           Predicate pred,
                                                           typename Engine::value_type* segments,
                                                           unsigned int* region_counts,
                                                           region_map regions,
                                                           std::uint32_t segment_capacity,
                                                           Counter* num_new,
                                                           Engine engine,
                                                           Action action)
{
  using slot_type = typename Engine::value_type;
  constexpr int items          = route_items_per_thread;
  constexpr index_type tile    = index_type{BlockSize} * items;
  constexpr int regions_per_thread = (route_max_regions + BlockSize - 1) / BlockSize;

  extern __shared__ __align__(128) unsigned char route_dynamic_smem[];
  auto* const stage = reinterpret_cast<slot_type*>(route_dynamic_smem);            // [tile]
  auto* const owner = reinterpret_cast<std::uint16_t*>(stage + tile);              // [tile]
  __shared__ unsigned int tile_hist[route_max_regions];    // elements of this tile per region
  __shared__ unsigned int tile_start[route_max_regions];   // first staged position of the region
  __shared__ unsigned int run_start[route_max_regions];    // reserved position in the segment
  __shared__ unsigned int warp_sums[BlockSize / 32];

  std::uint32_t const num_regions = regions.num_regions;
  unsigned long long mine         = 0;

  for (index_type base = index_type{blockIdx.x} * tile; base < n;
       base += index_type{gridDim.x} * tile) {
    for (std::uint32_t r = threadIdx.x; r < num_regions; r += BlockSize) {
      tile_hist[r] = 0;
    }
    __syncthreads();

    using input_type = decltype(engine.heterogeneous_value(read_input(first, index_type{0})));
    uninitialized<input_type> val[items];
    std::uint32_t region[items];
    std::uint32_t rank[items];
    // all loads of the tile in flight first; hashing starts when the first one lands
    unsigned live = 0;
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * BlockSize + threadIdx.x;
      if (idx < n && pred(*(stencil + idx))) {
        val[j].value = engine.heterogeneous_value(read_input(first, idx));
        live |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      region[j] = 0xffffffffu;
      if (live & (1u << j)) { region[j] = regions(engine.make_cursor(Engine::key_of(val[j].value)).slot); }
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (region[j] != 0xffffffffu) { rank[j] = atomicAdd(&tile_hist[region[j]], 1u); }
    }
    __syncthreads();

    // exclusive scan of the tile histogram; thread t owns regions [t * rpt, (t + 1) * rpt)
    {
      unsigned int held[regions_per_thread];
      unsigned int sum = 0;
#pragma unroll
      for (int i = 0; i < regions_per_thread; ++i) {
        std::uint32_t const r = threadIdx.x * regions_per_thread + i;
        held[i]               = r < num_regions ? tile_hist[r] : 0u;
        sum += held[i];
      }
      unsigned int inclusive = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned int const up = __shfl_up_sync(0xffffffffu, inclusive, d);
        if ((threadIdx.x & 31) >= d) { inclusive += up; }
      }
      if ((threadIdx.x & 31) == 31) { warp_sums[threadIdx.x >> 5] = inclusive; }
      __syncthreads();
      unsigned int running = inclusive - sum;
      for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) {
        running += warp_sums[w];
      }
#pragma unroll
      for (int i = 0; i < regions_per_thread; ++i) {
        std::uint32_t const r = threadIdx.x * regions_per_thread + i;
        if (r < num_regions) {
          tile_start[r] = running;
          run_start[r]  = held[i] ? atomicAdd(&region_counts[r], held[i]) : 0u;
          running += held[i];
        }
      }
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (region[j] != 0xffffffffu) {
        unsigned int const pos = tile_start[region[j]] + rank[j];
        stage[pos]             = engine.native_value(val[j].value);
        owner[pos]             = static_cast<std::uint16_t>(region[j]);
      }
    }
    __syncthreads();

    // staged elements of this tile = end of the last region's run
    unsigned int const count = tile_start[num_regions - 1] + tile_hist[num_regions - 1];
    for (unsigned int pos = threadIdx.x; pos < count; pos += BlockSize) {
      std::uint32_t const r     = owner[pos];
      std::uint64_t const where = std::uint64_t{run_start[r]} + (pos - tile_start[r]);
      if (where < segment_capacity) {
        segments[std::uint64_t{r} * segment_capacity + where] = stage[pos];
      } else {
        // segment full (heavily skewed input): finish this element now, unblocked
        mine += mutate_slow_path<ChunkSlots, load_policy::streaming>(
          engine, stage[pos], base + pos, action);
      }
    }
    __syncthreads();
  }
  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

/// Geometry of pass 2, shared by host and device.
struct blocked_layout {
  unsigned int const* counts;       ///< elements routed to each region (may exceed the capacity)
  std::uint32_t segment_capacity;   ///< elements a segment can hold
  std::uint64_t region_slots;       ///< ceil(capacity / num_regions): slots per region
  std::uint64_t table_bytes;        ///< end of the slot array (prefetch clamp)
  std::uint32_t prefetch_bytes;     ///< per-CTA share of the next region (multiple of 128), 0 = off
  std::uint32_t sources;            ///< segments per region (1; number of ranks for exchanged batches)
  std::uint64_t first_slot = 0;     ///< region 0 starts at this slot (batches confined to a table slice)
  std::uint32_t region_begin  = 0;  ///< this launch covers regions [region_begin, region_begin + gridDim.y / sources)
  std::uint32_t total_regions = 0;  ///< regions the table (slice) is divided into; 0 = those of this launch
  std::uint32_t source_major  = 0;  ///< 0: segment (region, source) is number region * sources + source;
                                    ///< 1: source * total_regions + region (exchange buffers filled per source)
};

template <int BlockSize,
          int KeysPerThread,
          int ChunkSlots,
          bool CasFirst,
          bool Counted,
          typename Counter,
          typename Engine,
          typename Action>
CUCO_KERNEL __launch_bounds__(BlockSize) void blocked_mutate_kernel(
  typename Engine::value_type const* segments,
  blocked_layout layout,
  Counter* num_new,
  Engine engine,
  Action action)
{
  constexpr index_type tile = index_type{BlockSize} * KeysPerThread;
  // blockIdx.y = region * sources + source: CTAs are dispatched region by region
  std::uint32_t const local_region = blockIdx.y / layout.sources;
  std::uint32_t const source       = blockIdx.y - local_region * layout.sources;
  std::uint32_t const region       = layout.region_begin + local_region;
  std::uint32_t const num_regions  = layout.total_regions != 0 ? layout.total_regions : gridDim.y / layout.sources;
  std::uint32_t const segment =
    layout.source_major != 0 ? source * num_regions + region : region * layout.sources + source;

  if (layout.prefetch_bytes != 0 && threadIdx.x == 0 && region + 1 < num_regions) {
    // stream this CTA's share of the next region's slots into L2 while this region is probed
    std::uint64_t const share_index = std::uint64_t{source} * gridDim.x + blockIdx.x;
    std::uint64_t const begin =
      (layout.first_slot + (std::uint64_t{region} + 1) * layout.region_slots) * Engine::slot_bytes +
      share_index * layout.prefetch_bytes;
    std::uint64_t const limit =
      (layout.first_slot + (std::uint64_t{region} + 2) * layout.region_slots) * Engine::slot_bytes;
    std::uint64_t end = begin + layout.prefetch_bytes;
    if (end > limit) { end = limit; }
    if (end > (layout.table_bytes & ~std::uint64_t{15})) { end = layout.table_bytes & ~std::uint64_t{15}; }
    if (begin < end) {
      auto const* address = reinterpret_cast<char const*>(engine.slots()) + (begin & ~std::uint64_t{15});
      auto const bytes    = static_cast<std::uint32_t>((end - (begin & ~std::uint64_t{15})) & ~std::uint64_t{15});
      if (bytes != 0) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(address), "r"(bytes) : "memory");
      }
    }
  }

  auto const stored = layout.counts[segment];
  index_type const count =
    stored < layout.segment_capacity ? index_type{stored} : index_type{layout.segment_capacity};
  index_type const tile_base = index_type{blockIdx.x} * tile;
  if (tile_base >= count) { return; }

  // Straight-line tile body. With the region's slots resident in L2 the kernel is bound by issue
  // slots and dependent round trips, not DRAM, so the general round loop of `mutate_tile` (bitmask
  // bookkeeping, divergent re-entry) is replaced by three phases over all keys of the thread -
  // post the input loads, post the home-sector loads, claim - and whatever is left (sector full of
  // other keys, lost race, foreign payload image) goes through the general driver afterwards.
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  using cursor    = typename Engine::cursor;
  constexpr bool key_only_claim = Action::key_then_apply && sizeof(slot_type) > 8;
  constexpr auto policy         = load_policy::streaming;

  index_type const first_idx = index_type{segment} * layout.segment_capacity + tile_base + threadIdx.x;
  index_type const limit     = index_type{segment} * layout.segment_capacity + count;
  auto* const table          = engine.slots();
  auto const empty_slot      = engine.empty_slot_sentinel();
  unsigned long long mine    = 0;

  uninitialized<slot_type> val[KeysPerThread];
  cursor cur[KeysPerThread];
  unsigned live = 0, claiming = 0, todo = 0, rare = 0;

#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    index_type const idx = first_idx + index_type{j} * BlockSize;
    if (idx < limit) {
      val[j].value = read_input(segments, idx);
      live |= 1u << j;
    }
  }
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    if (live & (1u << j)) { cur[j] = engine.make_cursor(Engine::key_of(val[j].value)); }
  }

  slot_type seen[KeysPerThread];
  if constexpr (CasFirst) {
    claiming = live;
  } else {
    raw_chunk<ChunkSlots * Engine::slot_bytes> raw[KeysPerThread];
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (live & (1u << j)) { raw[j] = engine.template load_chunk<ChunkSlots, policy>(cur[j]); }
    }
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (!(live & (1u << j))) { continue; }
      auto const& key = Engine::key_of(val[j].value);
      int const begin_off =
        static_cast<int>(cur[j].slot - Engine::template chunk_begin<ChunkSlots>(cur[j]));
      int const valid = engine.template chunk_valid<ChunkSlots>(cur[j]);
      int state_of    = 0;  // 0: sector exhausted, 1: present, 2: claim at cur[j].slot
#pragma unroll
      for (int i = 0; i < ChunkSlots; ++i) {
        if (state_of == 0 && i >= begin_off && i < begin_off + valid) {
          auto const slot  = chunk_slot<slot_type>(raw[j], i);
          auto const state = engine.classify_insert(key, Engine::key_of(slot));
          if (state == equal_result::EQUAL) {
            action.on_present(engine,
                              first_idx + index_type{j} * BlockSize,
                              table + (cur[j].slot + (i - begin_off)),
                              slot,
                              val[j].value);
            state_of = 1;
          } else if (state == equal_result::AVAILABLE) {
            if (i > begin_off) { engine.advance(cur[j], i - begin_off); }  // now at the free slot
            state_of = 2;
          }
        }
      }
      if (state_of == 0) {
        engine.advance(cur[j], valid);
        todo |= 1u << j;
      }
      if (state_of == 2) { claiming |= 1u << j; }
    }
  }

#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    if (claiming & (1u << j)) {
      if constexpr (key_only_claim) {
        key_type expected_key = Engine::key_of(empty_slot);
        cuda::atomic_ref<key_type, Engine::thread_scope> key_ref{(table + cur[j].slot)->first};
        key_ref.compare_exchange_strong(expected_key,
                                        static_cast<key_type>(Engine::key_of(val[j].value)),
                                        cuda::memory_order_relaxed);
        seen[j]       = empty_slot;
        seen[j].first = expected_key;
      } else {
        seen[j] = cas_slot<Engine::thread_scope>(table + cur[j].slot, empty_slot, val[j].value);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    if (!(claiming & (1u << j))) { continue; }
    index_type const idx = first_idx + index_type{j} * BlockSize;
    auto* const address  = table + cur[j].slot;
    bool const won       = key_only_claim
                             ? same_bits(Engine::key_of(seen[j]), Engine::key_of(empty_slot))
                             : same_bits(seen[j], empty_slot);
    if (won) {
      action.on_new(engine, idx, address, val[j].value);
      ++mine;
    } else if (engine.classify_insert(Engine::key_of(val[j].value), Engine::key_of(seen[j])) ==
               equal_result::EQUAL) {
      action.on_present(engine, idx, address, seen[j], val[j].value);
    } else if (engine.classify_insert(Engine::key_of(val[j].value), Engine::key_of(seen[j])) ==
               equal_result::AVAILABLE) {
      rare |= 1u << j;  // claimable, but not the canonical empty image (foreign payload bits)
    } else {
      engine.advance(cur[j], 1);  // another key won this slot: carry on behind it
      todo |= 1u << j;
    }
  }

  // Leftover keys (home sector full of other keys, or lost a race): warp-converged rounds. Every
  // lane takes its lowest pending key and runs one probe step on it (sector load, scan, claim), so
  // a round costs one pass through this body for the whole warp instead of one divergent walk
  // per key.
  while (__any_sync(0xffffffffu, todo != 0)) {
    if (todo != 0) {
      int const jj = __ffs(todo) - 1;
      slot_type v  = val[0].value;
      cursor c     = cur[0];
#pragma unroll
      for (int j = 1; j < KeysPerThread; ++j) {
        if (j == jj) {
          v = val[j].value;
          c = cur[j];
        }
      }
      index_type const idx = first_idx + index_type{jj} * BlockSize;
      auto const& key      = Engine::key_of(v);
      auto const raw       = engine.template load_chunk<ChunkSlots, policy>(c);
      int const begin_off  = static_cast<int>(c.slot - Engine::template chunk_begin<ChunkSlots>(c));
      int const valid      = engine.template chunk_valid<ChunkSlots>(c);
      int outcome          = 0;  // 0: keep probing, 1: finished, 2: rare path
      int consumed         = valid;
#pragma unroll
      for (int i = 0; i < ChunkSlots; ++i) {
        if (outcome == 0 && consumed == valid && i >= begin_off && i < begin_off + valid) {
          auto const slot  = chunk_slot<slot_type>(raw, i);
          auto const state = engine.classify_insert(key, Engine::key_of(slot));
          if (state == equal_result::EQUAL) {
            action.on_present(engine, idx, table + (c.slot + (i - begin_off)), slot, v);
            outcome = 1;
          } else if (state == equal_result::AVAILABLE) {
            consumed = i - begin_off;
          }
        }
      }
      if (outcome == 0) {
        if (consumed > 0) { engine.advance(c, consumed); }
        if (consumed < valid) {
          // c.slot is the available slot: claim it
          auto* const address = table + c.slot;
          slot_type observed;
          if constexpr (key_only_claim) {
            key_type expected_key = Engine::key_of(empty_slot);
            cuda::atomic_ref<key_type, Engine::thread_scope> key_ref{address->first};
            key_ref.compare_exchange_strong(
              expected_key, static_cast<key_type>(key), cuda::memory_order_relaxed);
            observed       = empty_slot;
            observed.first = expected_key;
          } else {
            observed = cas_slot<Engine::thread_scope>(address, empty_slot, v);
          }
          bool const won = key_only_claim
                             ? same_bits(Engine::key_of(observed), Engine::key_of(empty_slot))
                             : same_bits(observed, empty_slot);
          if (won) {
            action.on_new(engine, idx, address, v);
            ++mine;
            outcome = 1;
          } else {
            auto const state = engine.classify_insert(key, Engine::key_of(observed));
            if (state == equal_result::EQUAL) {
              action.on_present(engine, idx, address, observed, v);
              outcome = 1;
            } else if (state == equal_result::AVAILABLE) {
              outcome = 2;
            } else {
              engine.advance(c, 1);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < KeysPerThread; ++j) {
        if (j == jj) { cur[j] = c; }
      }
      if (outcome != 0) { todo &= ~(1u << jj); }
      if (outcome == 2) { rare |= 1u << jj; }
    }
  }
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    if (rare & (1u << j)) {
      mine += mutate_slow_path<ChunkSlots, policy>(
        engine, val[j].value, first_idx + index_type{j} * BlockSize, action);
    }
  }
  if constexpr (Counted) { accumulate_count(num_new, mine); }
}

// =================================================================================================
// Exchange path of hash-partitioned tables: ONE routing kernel groups a rank's batch by owner rank
// AND by L2 region of the owner's table and stores it straight into the owners' memory
// =================================================================================================
//
// No reference counterpart (cuCollections is single-GPU). Every rank holds one shard
// with identical geometry, so a sender can compute the region an element will land in on its owner.
// `exchange_route_kernel` is `route_kernel` with the combined bucket (owner, region): the staged runs
// are written with plain stores through peer pointers (NVLink 5 / NVSwitch P2P; the run granularity
// is what keeps those stores line-sized), so routing, the all-to-all and the L2 grouping of the
// owner's pass 2 are the same kernel. Layouts (cap = segment capacity in elements):
//   on the owner `o`:   segment g = region * P + source      at element offset g * cap
//   on the source `s`:  segment l = owner * R + region        (counters, source indices, results)
// A source is the only writer of its segments, so their fill counts are local atomics and are
// published to the owners by `exchange_publish_kernel` before the cross-rank barrier. Elements
// that do not fit their segment (heavily skewed batches) go to a local spill list that the host
// side finishes through the collective fallback.

constexpr int exchange_max_ranks = 16;

/// Peer base pointers of one exchange buffer (index = rank).
struct exchange_peers {
  void* base[exchange_max_ranks];
};

/// Owner of a key: high bits of a mix that shares nothing with the in-table hash.
__host__ __device__ inline std::uint32_t exchange_owner(std::uint64_t key_bits,
                                                        std::uint64_t salt,
                                                        std::uint32_t num_ranks) noexcept
{
  std::uint64_t x = key_bits ^ salt;
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
#if defined(__CUDA_ARCH__)
  return static_cast<std::uint32_t>(__umul64hi(x, std::uint64_t{num_ranks}));
#else
  return static_cast<std::uint32_t>((static_cast<unsigned __int128>(x) * num_ranks) >> 64);
#endif
}

struct exchange_geometry {
  std::uint32_t num_ranks;         ///< P
  std::uint32_t my_rank;
  std::uint32_t num_regions;       ///< R (P * R <= route_max_regions)
  std::uint32_t segment_capacity;  ///< cap
  std::uint64_t salt;
  /// Where a routed element lands behind `base[owner]`: segment (region * dest_stride + dest_slot).
  /// Peer stores into the owners' buffers use (P, my_rank): the owner sees its segments region-major,
  /// source-minor. Staging into a LOCAL buffer (copied to the owners later) uses (1, 0).
  std::uint32_t dest_stride;
  std::uint32_t dest_slot;
};

template <typename Key>
__device__ __forceinline__ std::uint64_t exchange_key_bits(Key const& key) noexcept
{
  if constexpr (sizeof(Key) == 8) {
    std::uint64_t bits;
    memcpy(&bits, &key, 8);
    return bits;
  } else {
    static_assert(sizeof(Key) == 4);
    std::int32_t bits;
    memcpy(&bits, &key, 4);
    return static_cast<std::uint64_t>(static_cast<std::int64_t>(bits));  // sign-extended like int64(key)
  }
}

/// Dynamic shared memory of `exchange_route_kernel`.
template <int BlockSize, typename Elem, bool WithIndex>
constexpr std::size_t exchange_smem_bytes() noexcept
{
  return std::size_t{BlockSize} * route_items_per_thread *
         (sizeof(Elem) + sizeof(std::uint16_t) + (WithIndex ? sizeof(std::uint32_t) : 0));
}

/// KeysOnly = false: elements are slot images (mutations); true: elements are keys (lookups), and
/// the position each key's result will come back to is recorded in `position_local`.
template <int BlockSize, bool KeysOnly, typename InputIt, typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize, 2) void exchange_route_kernel(
  InputIt first,
  index_type n,
  exchange_peers peers,            ///< owners' segment buffers
  unsigned int* counts_local,      ///< [P * R] fill counts of this rank's segments
  std::uint32_t* position_local,   ///< [n] source-side position of every routed key (KeysOnly);
                                   ///< 0xffffffff for keys that went to the spill list
  void* spill,                     ///< [spill_capacity] elements that did not fit
  std::uint32_t* spill_index,      ///< their source indices (KeysOnly)
  unsigned int* spill_count,
  std::uint32_t spill_capacity,
  region_map regions,
  exchange_geometry geometry,
  Engine engine)
{
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  using elem_type = cuda::std::conditional_t<KeysOnly, key_type, slot_type>;
  constexpr int items       = route_items_per_thread;
  constexpr index_type tile = index_type{BlockSize} * items;
  constexpr int buckets_per_thread = (route_max_regions + BlockSize - 1) / BlockSize;

  extern __shared__ __align__(128) unsigned char route_dynamic_smem[];
  auto* const stage  = reinterpret_cast<elem_type*>(route_dynamic_smem);   // [tile]
  auto* const origin = reinterpret_cast<std::uint32_t*>(stage + tile);     // [tile] (KeysOnly)
  auto* const bucket_of =
    reinterpret_cast<std::uint16_t*>(origin + (KeysOnly ? tile : 0));       // [tile]
  __shared__ unsigned int tile_hist[route_max_regions];
  __shared__ unsigned int tile_start[route_max_regions];
  __shared__ unsigned int run_start[route_max_regions];
  __shared__ unsigned int warp_sums[BlockSize / 32];

  std::uint32_t const P           = geometry.num_ranks;
  std::uint32_t const R           = geometry.num_regions;
  std::uint32_t const cap         = geometry.segment_capacity;
  std::uint32_t const num_buckets = P * R;

  for (index_type base = index_type{blockIdx.x} * tile; base < n;
       base += index_type{gridDim.x} * tile) {
    for (std::uint32_t b = threadIdx.x; b < num_buckets; b += BlockSize) {
      tile_hist[b] = 0;
    }
    __syncthreads();

    using input_type = decltype(read_input(first, index_type{0}));
    uninitialized<input_type> val[items];
    std::uint32_t bucket[items];
    std::uint32_t rank[items];
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * BlockSize + threadIdx.x;
      if (idx < n) { val[j].value = read_input(first, idx); }
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * BlockSize + threadIdx.x;
      bucket[j]            = 0xffffffffu;
      if (idx < n) {
        key_type key;
        if constexpr (KeysOnly) {
          key = static_cast<key_type>(val[j].value);
        } else {
          key = static_cast<key_type>(Engine::key_of(engine.heterogeneous_value(val[j].value)));
        }
        auto const owner  = exchange_owner(exchange_key_bits(key), geometry.salt, P);
        auto const region = regions(engine.make_cursor(key).slot);
        bucket[j]         = owner * R + region;
      }
    }
    if (num_buckets <= 8) {
      // Few buckets (routing by owner, or by owner and a handful of table slices): one shared-memory
      // atomic per element would serialise the whole tile on a few counters. The lanes of a warp that
      // share a bucket elect a leader, which reserves their ranks with ONE atomic.
      unsigned const lane = threadIdx.x & 31u;
#pragma unroll
      for (int j = 0; j < items; ++j) {
        unsigned const same   = __match_any_sync(0xffffffffu, bucket[j]);
        unsigned const leader = static_cast<unsigned>(__ffs(same)) - 1u;
        unsigned int base     = 0;
        if (lane == leader && bucket[j] != 0xffffffffu) {
          base = atomicAdd(&tile_hist[bucket[j]], static_cast<unsigned int>(__popc(same)));
        }
        base    = __shfl_sync(0xffffffffu, base, static_cast<int>(leader));
        rank[j] = base + static_cast<unsigned int>(__popc(same & ((1u << lane) - 1u)));
      }
    } else {
#pragma unroll
      for (int j = 0; j < items; ++j) {
        if (bucket[j] != 0xffffffffu) { rank[j] = atomicAdd(&tile_hist[bucket[j]], 1u); }
      }
    }
    __syncthreads();
    {
      unsigned int held[buckets_per_thread];
      unsigned int sum = 0;
#pragma unroll
      for (int i = 0; i < buckets_per_thread; ++i) {
        std::uint32_t const b = threadIdx.x * buckets_per_thread + i;
        held[i]               = b < num_buckets ? tile_hist[b] : 0u;
        sum += held[i];
      }
      unsigned int inclusive = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned int const up = __shfl_up_sync(0xffffffffu, inclusive, d);
        if ((threadIdx.x & 31) >= d) { inclusive += up; }
      }
      if ((threadIdx.x & 31) == 31) { warp_sums[threadIdx.x >> 5] = inclusive; }
      __syncthreads();
      unsigned int running = inclusive - sum;
      for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) {
        running += warp_sums[w];
      }
#pragma unroll
      for (int i = 0; i < buckets_per_thread; ++i) {
        std::uint32_t const b = threadIdx.x * buckets_per_thread + i;
        if (b < num_buckets) {
          tile_start[b] = running;
          run_start[b]  = held[i] ? atomicAdd(&counts_local[b], held[i]) : 0u;
          running += held[i];
        }
      }
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (bucket[j] != 0xffffffffu) {
        unsigned int const pos = tile_start[bucket[j]] + rank[j];
        if constexpr (KeysOnly) {
          stage[pos]  = static_cast<key_type>(val[j].value);
          origin[pos] = static_cast<std::uint32_t>(base + index_type{j} * BlockSize + threadIdx.x);
        } else {
          stage[pos] = engine.native_value(engine.heterogeneous_value(val[j].value));
        }
        bucket_of[pos] = static_cast<std::uint16_t>(bucket[j]);
      }
    }
    __syncthreads();

    auto const count = static_cast<unsigned int>((n - base) < tile ? (n - base) : tile);
    for (unsigned int pos = threadIdx.x; pos < count; pos += BlockSize) {
      std::uint32_t const b     = bucket_of[pos];
      std::uint64_t const where = std::uint64_t{run_start[b]} + (pos - tile_start[b]);
      if (where < cap) {
        std::uint32_t const owner  = b / R;
        std::uint32_t const region = b - owner * R;
        auto* const segment =
          static_cast<elem_type*>(peers.base[owner]) +
          (std::uint64_t{region} * geometry.dest_stride + geometry.dest_slot) * cap;
        segment[where] = stage[pos];
        // the tile's 4096 positions are written together, so these scattered 4-byte stores merge
        // into full lines in L2; the return trip then GATHERS by position with coalesced output
        if constexpr (KeysOnly) {
          position_local[origin[pos]] = static_cast<std::uint32_t>(std::uint64_t{b} * cap + where);
        }
      } else {
        unsigned int const at = atomicAdd(spill_count, 1u);
        if (at < spill_capacity) {
          static_cast<elem_type*>(spill)[at] = stage[pos];
          if constexpr (KeysOnly) { spill_index[at] = origin[pos]; }
        }
        if constexpr (KeysOnly) { position_local[origin[pos]] = 0xffffffffu; }
      }
    }
    __syncthreads();
  }
}

/// Tells every owner how many elements this rank put into each of its segments there, and every
/// peer how many elements this rank spilled (peers need to agree on running the fallback).
template <typename Geometry>  // a template only so that this header-defined kernel has vague linkage
CUCO_KERNEL void exchange_publish_kernel(unsigned int const* counts_local,
                                         unsigned int const* spill_count,
                                         exchange_peers counts_recv,  ///< owners' [R * P] arrays
                                         exchange_peers spill_flags,  ///< peers' [P] arrays
                                         Geometry geometry)
{
  std::uint32_t const P = geometry.num_ranks, R = geometry.num_regions;
  for (std::uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < P * R; b += gridDim.x * blockDim.x) {
    std::uint32_t const owner = b / R, region = b - owner * R;
    unsigned int const filled = counts_local[b];
    static_cast<unsigned int*>(counts_recv.base[owner])[region * geometry.dest_stride + geometry.dest_slot] =
      filled < geometry.segment_capacity ? filled : geometry.segment_capacity;
  }
  if (blockIdx.x == 0 && threadIdx.x < P) {
    static_cast<unsigned int*>(spill_flags.base[threadIdx.x])[geometry.my_rank] = *spill_count;
  }
}

/// Owner side of a routed lookup: probes the keys of segment g = region * P + source and stores the
/// results at the key's position behind `results.base[source]`, segment
/// (geometry.dest_slot * R + region): the SOURCE's result buffer (peer stores, dest_slot = this rank)
/// or a local buffer that the copy engines return afterwards (dest_slot = 0, one base per source).
/// Every result is kept in a register until its key is resolved and stored by ONE instruction per
/// key slot, so a warp writes whole 256-byte runs (8-byte results) instead of the hit / miss /
/// leftover fragments of the probe loop - that matters for stores that cross NVLink.
template <int BlockSize, int KeysPerThread, int ChunkSlots, typename Result, typename Engine, typename Emit>
CUCO_KERNEL __launch_bounds__(BlockSize) void exchange_lookup_kernel(
  typename Engine::key_type const* segments,
  unsigned int const* counts_recv,
  exchange_peers results,
  exchange_geometry geometry,
  Engine engine,
  Emit emit)
{
  using slot_type = typename Engine::value_type;
  using key_type  = typename Engine::key_type;
  using cursor    = typename Engine::cursor;
  constexpr index_type tile = index_type{BlockSize} * KeysPerThread;
  constexpr auto policy     = load_policy::readonly;

  // Sources are visited in an order rotated by this rank, so that at any moment the owners store
  // results to DIFFERENT sources: with the same order everywhere all ranks would hit one source's
  // NVLink ingress at a time (measured at 4 GPUs: 5.7 ms vs 3.2 ms for this kernel).
  std::uint32_t const region = blockIdx.y / geometry.num_ranks;
  std::uint32_t const source =
    (blockIdx.y - region * geometry.num_ranks + geometry.my_rank + 1) % geometry.num_ranks;
  std::uint32_t const segment = region * geometry.num_ranks + source;
  auto const stored           = counts_recv[segment];
  index_type const count =
    stored < geometry.segment_capacity ? index_type{stored} : index_type{geometry.segment_capacity};
  index_type const tile_base = index_type{blockIdx.x} * tile;
  if (tile_base >= count) { return; }

  key_type const* const in = segments + std::uint64_t{segment} * geometry.segment_capacity;
  Result* const out        = static_cast<Result*>(results.base[source]) +
                      (std::uint64_t{geometry.dest_slot} * geometry.num_regions + region) *
                        geometry.segment_capacity;

  uninitialized<key_type> key[KeysPerThread];
  cursor cur[KeysPerThread];
  Result answer[KeysPerThread];
  unsigned live = 0;
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
    if (idx < count) {
      key[j].value = read_input(in, idx);
      live |= 1u << j;
    }
  }
  unsigned pending = live;
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    if (pending & (1u << j)) { cur[j] = engine.make_cursor(key[j].value); }
  }
  // first probe of every key, straight line: all sector loads in flight, then all scans
  {
    raw_chunk<ChunkSlots * Engine::slot_bytes> raw[KeysPerThread];
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (pending & (1u << j)) { raw[j] = engine.template load_chunk<ChunkSlots, policy>(cur[j]); }
    }
#pragma unroll
    for (int j = 0; j < KeysPerThread; ++j) {
      if (pending & (1u << j)) {
        int const begin_off =
          static_cast<int>(cur[j].slot - Engine::template chunk_begin<ChunkSlots>(cur[j]));
        int const valid = engine.template chunk_valid<ChunkSlots>(cur[j]);
        bool done       = false;
#pragma unroll
        for (int i = 0; i < ChunkSlots; ++i) {
          if (!done && i >= begin_off && i < begin_off + valid) {
            auto const slot  = chunk_slot<slot_type>(raw[j], i);
            auto const state = engine.classify_lookup(key[j].value, Engine::key_of(slot));
            if (state == equal_result::EQUAL) {
              answer[j] = static_cast<Result>(emit.hit(slot));
              done      = true;
            } else if (state == equal_result::EMPTY) {
              answer[j] = static_cast<Result>(emit.miss());
              done      = true;
            }
          }
        }
        if (done) {
          pending &= ~(1u << j);
        } else {
          engine.advance(cur[j], valid);
        }
      }
    }
  }
  // leftovers (home sector full of other keys): warp-converged rounds, lowest pending key per lane
  while (__any_sync(0xffffffffu, pending != 0)) {
    if (pending != 0) {
      int const jj = __ffs(pending) - 1;
      key_type k   = key[0].value;
      cursor c     = cur[0];
#pragma unroll
      for (int j = 1; j < KeysPerThread; ++j) {
        if (j == jj) {
          k = key[j].value;
          c = cur[j];
        }
      }
      auto const raw      = engine.template load_chunk<ChunkSlots, policy>(c);
      int const begin_off = static_cast<int>(c.slot - Engine::template chunk_begin<ChunkSlots>(c));
      int const valid     = engine.template chunk_valid<ChunkSlots>(c);
      bool done           = false;
      Result found{};
#pragma unroll
      for (int i = 0; i < ChunkSlots; ++i) {
        if (!done && i >= begin_off && i < begin_off + valid) {
          auto const slot  = chunk_slot<slot_type>(raw, i);
          auto const state = engine.classify_lookup(k, Engine::key_of(slot));
          if (state == equal_result::EQUAL) {
            found = static_cast<Result>(emit.hit(slot));
            done  = true;
          } else if (state == equal_result::EMPTY) {
            found = static_cast<Result>(emit.miss());
            done  = true;
          }
        }
      }
      if (done) {
        pending &= ~(1u << jj);
#pragma unroll
        for (int j = 0; j < KeysPerThread; ++j) {
          if (j == jj) { answer[j] = found; }
        }
      } else {
        engine.advance(c, valid);
#pragma unroll
        for (int j = 0; j < KeysPerThread; ++j) {
          if (j == jj) { cur[j] = c; }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KeysPerThread; ++j) {
    index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
    if (live & (1u << j)) { out[idx] = answer[j]; }
  }
}

/// Source side of a routed lookup: results came back in source-side segment order; every input
/// position fetches its own (coalesced stores, gathers that stay within the tile's runs).
template <typename Result, typename OutputIt>
CUCO_KERNEL __launch_bounds__(256) void exchange_unpermute_kernel(Result const* results,
                                                                  std::uint32_t const* position_local,
                                                                  index_type n,
                                                                  OutputIt out)
{
  constexpr int items = 4;  // independent gathers in flight per thread
  for (index_type base = index_type{blockIdx.x} * (256 * items); base < n;
       base += index_type{gridDim.x} * (256 * items)) {
    std::uint32_t position[items];
    Result value[items];
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * 256 + threadIdx.x;
      position[j]          = idx < n ? __ldcs(position_local + idx) : 0xffffffffu;
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      if (position[j] != 0xffffffffu) { value[j] = results[position[j]]; }
    }
#pragma unroll
    for (int j = 0; j < items; ++j) {
      index_type const idx = base + index_type{j} * 256 + threadIdx.x;
      if (position[j] != 0xffffffffu) { *(out + idx) = value[j]; }
    }
  }
}

// =================================================================================================
// erase (tombstoning) - two keys per thread
// =================================================================================================
/// `ChunkSlots` > window size only on container-owned (padded, 32-byte aligned) storage: the walk
/// then reads the whole sector a probe lands in, like the lookup kernels.
template <int BlockSize, int ChunkSlots, typename InputIt, typename Engine>
CUCO_KERNEL __launch_bounds__(BlockSize) void erase_kernel(InputIt first, index_type n, Engine engine)
{
  // Two keys per thread, both input loads and both home-chunk loads posted before either is consumed
  // (one key per thread leaves the DRAM round trips of a thread back to back: 15.6 G keys/s, the same as
  // cuco's one-key-per-thread kernel). A key whose home chunk does not settle it continues with the
  // general walk.
  using slot_type  = typename Engine::value_type;
  using cursor     = typename Engine::cursor;
  using probe_type = decltype(read_input(first, index_type{0}));
  constexpr int keys_per_thread = 2;
  constexpr index_type tile     = index_type{BlockSize} * keys_per_thread;
  constexpr auto policy         = load_policy::streaming;

  for (index_type tile_base = index_type{blockIdx.x} * tile; tile_base < n;
       tile_base += index_type{gridDim.x} * tile) {
    uninitialized<probe_type> key[keys_per_thread];
    cursor cur[keys_per_thread];
    unsigned pending = 0;
#pragma unroll
    for (int j = 0; j < keys_per_thread; ++j) {
      index_type const idx = tile_base + index_type{j} * BlockSize + threadIdx.x;
      if (idx < n) {
        key[j].value = read_input(first, idx);
        pending |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < keys_per_thread; ++j) {
      if (pending & (1u << j)) { cur[j] = engine.make_cursor(key[j].value); }
    }
    raw_chunk<ChunkSlots * Engine::slot_bytes> raw[keys_per_thread];
#pragma unroll
    for (int j = 0; j < keys_per_thread; ++j) {
      if (pending & (1u << j)) { raw[j] = engine.template load_chunk<ChunkSlots, policy>(cur[j]); }
    }
#pragma unroll
    for (int j = 0; j < keys_per_thread; ++j) {
      if (!(pending & (1u << j))) { continue; }
      int const begin_off = static_cast<int>(cur[j].slot - Engine::template chunk_begin<ChunkSlots>(cur[j]));
      int const valid     = engine.template chunk_valid<ChunkSlots>(cur[j]);
      bool done           = false;
#pragma unroll
      for (int i = 0; i < ChunkSlots; ++i) {
        if (!done && i >= begin_off && i < begin_off + valid) {
          auto const slot  = chunk_slot<slot_type>(raw[j], i);
          auto const state = engine.classify_lookup(key[j].value, Engine::key_of(slot));
          if (state == equal_result::EQUAL) {
            engine.retire_observed(engine.slots() + (cur[j].slot + (i - begin_off)), slot);
            done = true;
          } else if (state == equal_result::EMPTY) {
            done = true;
          }
        }
      }
      if (!done) {
        engine.advance(cur[j], valid);
        engine.template erase_from<ChunkSlots, policy>(cur[j], key[j].value);
      }
    }
  }
}

// =================================================================================================
// size: streaming count of filled slots (neither empty nor erased key)
// =================================================================================================
template <int BlockSize, typename Engine, typename Counter>
CUCO_KERNEL __launch_bounds__(BlockSize) void size_kernel(Engine engine, Counter* count)
{
  using slot_type = typename Engine::value_type;
  auto const* table = engine.slots();
  auto const n      = static_cast<index_type>(engine.capacity());
  auto const empty  = engine.empty_key_sentinel();
  auto const erased = engine.erased_key_sentinel();
  unsigned long long mine = 0;

  constexpr int chunk = Engine::sector_chunk_slots;
  bool const vector_ok = (reinterpret_cast<std::uintptr_t>(table) % 32) == 0;
  if (chunk > 1 && vector_ok) {
    // whole sectors; the allocation is padded so the last chunk may be read in full. Four
    // independent sector loads per thread and iteration; the launch is a persistent grid so that the
    // final count costs a few thousand atomics on the one counter, not one per warp of the table.
    index_type const chunks = (n + chunk - 1) / chunk;
    constexpr int unroll    = 4;
    for (index_type base = index_type{blockIdx.x} * (BlockSize * unroll); base < chunks;
         base += index_type{gridDim.x} * (BlockSize * unroll)) {
      raw_chunk<chunk * Engine::slot_bytes> raw[unroll];
#pragma unroll
      for (int u = 0; u < unroll; ++u) {
        index_type const c = base + index_type{u} * BlockSize + threadIdx.x;
        if (c < chunks) {
          raw[u] = load_chunk_bytes<chunk * Engine::slot_bytes, load_policy::readonly>(table + c * chunk);
        }
      }
#pragma unroll
      for (int u = 0; u < unroll; ++u) {
        index_type const c = base + index_type{u} * BlockSize + threadIdx.x;
        if (c < chunks) {
#pragma unroll
          for (int i = 0; i < chunk; ++i) {
            if (c * chunk + i < n) {
              auto const& k = Engine::key_of(chunk_slot<slot_type>(raw[u], i));
              mine += !(same_bits(k, empty) || same_bits(k, erased));
            }
          }
        }
      }
    }
  } else {
    for (index_type i = cuco::detail::global_thread_id(); i < n;
         i += cuco::detail::grid_stride()) {
      auto const& k = Engine::key_of(table[i]);
      mine += !(same_bits(k, empty) || same_bits(k, erased));
    }
  }
  accumulate_count(count, mine);
}

}  // namespace cuco::b200
