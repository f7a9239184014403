// Device-ref operator mixins shared by static_map_ref and static_set_ref.
//
// The reference defines one `operator_impl<op::X_tag, Ref>` specialisation per container and per
// operator (detail/static_map/static_map_ref.inl:400-1370, detail/static_set/static_set_ref.inl:
// 351-633), all forwarding into the shared ref impl. Here the forwarding bodies are written once,
// parameterised on the concrete ref type; the container headers specialise `operator_impl` by
// inheriting from these. Every mixin reaches the probe engine through `Ref::engine()`.
//
// Per-thread overloads require cg_size == 1 like the reference (ref_impl.cuh:376); per-tile
// overloads take a `cooperative_groups::thread_block_tile<cg_size>` and return the same value on
// every lane.
#pragma once

#include <cuco/b200/match_kernels.cuh>
#include <cuco/b200/probe_engine.cuh>
#include <cuco/operator.hpp>

#include <cuda/atomic>
#include <cuda/std/type_traits>
#include <thrust/pair.h>

#include <cooperative_groups.h>

namespace cuco::b200 {

#define CUCO_B200_SCALAR_ONLY()   \
  static_assert(Ref::cg_size == 1, \
                "Non-CG operation is incompatible with the current probing scheme")

/// `Tag` only makes the base unique per mixin: a ref inherits several mixins, and a shared base type
/// would make the downcast to `Ref` ambiguous.
template <typename Ref, typename Tag>
struct mixin_base {
 protected:
  __device__ auto& self() noexcept { return static_cast<Ref&>(*this); }
  __device__ auto const& self() const noexcept { return static_cast<Ref const&>(*this); }
};

// ---- insert --------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_insert : mixin_base<Ref, op::insert_tag> {
  /// Inserts `value`; true iff the key was not present before.
  template <typename Value>
  __device__ bool insert(Value const& value) noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_insert(value);
  }

  template <typename Value>
  __device__ bool insert(cooperative_groups::thread_block_tile<CGSize> const& group,
                         Value const& value) noexcept
  {
    return this->self().engine().tile_insert(group, value);
  }
};

// ---- insert_and_find -----------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_insert_and_find : mixin_base<Ref, op::insert_and_find_tag> {
  /// Inserts `value` if absent; returns {handle to the resident slot, whether this call created it}.
  /// The handle's payload is always a published value, never the empty sentinel.
  template <typename Value>
  __device__ auto insert_and_find(Value const& value) noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_insert_and_find(value);
  }

  template <typename Value>
  __device__ auto insert_and_find(cooperative_groups::thread_block_tile<CGSize> const& group,
                                  Value const& value) noexcept
  {
    return this->self().engine().tile_insert_and_find(group, value);
  }
};

// ---- erase ---------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_erase : mixin_base<Ref, op::erase_tag> {
  /// Replaces the entry of `key` by a tombstone; true iff this call removed it.
  template <typename ProbeKey>
  __device__ bool erase(ProbeKey const& key) noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_erase(key);
  }

  template <typename ProbeKey>
  __device__ bool erase(cooperative_groups::thread_block_tile<CGSize> const& group,
                        ProbeKey const& key) noexcept
  {
    return this->self().engine().tile_erase(group, key);
  }
};

// ---- contains ------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_contains : mixin_base<Ref, op::contains_tag> {
  template <typename ProbeKey>
  [[nodiscard]] __device__ bool contains(ProbeKey const& key) const noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_contains(key);
  }

  template <typename ProbeKey>
  [[nodiscard]] __device__ bool contains(
    cooperative_groups::thread_block_tile<CGSize> const& group, ProbeKey const& key) const noexcept
  {
    return this->self().engine().tile_contains(group, key);
  }
};

// ---- count ---------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_count : mixin_base<Ref, op::count_tag> {
  template <typename ProbeKey>
  [[nodiscard]] __device__ auto count(ProbeKey const& key) const noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_count(key);
  }

  /// Matches found by the calling lane in its own windows; sum over the tile for the key's count.
  template <typename ProbeKey>
  [[nodiscard]] __device__ auto count(cooperative_groups::thread_block_tile<CGSize> const& group,
                                      ProbeKey const& key) const noexcept
  {
    return this->self().engine().tile_count(group, key);
  }
};

// ---- find ----------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_find : mixin_base<Ref, op::find_tag> {
  /// Handle to the slot holding `key`, or `end()`.
  template <typename ProbeKey>
  [[nodiscard]] __device__ auto find(ProbeKey const& key) const noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    return this->self().engine().scalar_find(key);
  }

  template <typename ProbeKey>
  [[nodiscard]] __device__ auto find(cooperative_groups::thread_block_tile<CGSize> const& group,
                                     ProbeKey const& key) const noexcept
  {
    return this->self().engine().tile_find(group, key);
  }
};

// ---- for_each ------------------------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_for_each : mixin_base<Ref, op::for_each_tag> {
  /// Invokes `callback(slot content)` for every entry whose key equals `key`.
  template <typename ProbeKey, typename Callback>
  __device__ void for_each(ProbeKey const& key, Callback&& callback) const noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    this->self().engine().scalar_for_each(key, callback);
  }

  /// Tile flavour: the lane whose window holds a match runs the callback on it.
  template <typename ProbeKey, typename Callback>
  __device__ void for_each(cooperative_groups::thread_block_tile<CGSize> const& group,
                           ProbeKey const& key,
                           Callback&& callback) const noexcept
  {
    this->self().engine().tile_for_each(group, key, callback, [](auto const&) {});
  }

  /// As above; `sync_op(group)` runs after every probe step (e.g. to flush a staging buffer).
  template <typename ProbeKey, typename Callback, typename SyncOp>
  __device__ void for_each(cooperative_groups::thread_block_tile<CGSize> const& group,
                           ProbeKey const& key,
                           Callback&& callback,
                           SyncOp&& sync_op) const noexcept
  {
    this->self().engine().tile_for_each(group, key, callback, sync_op);
  }
};

// ---- retrieve (multi-containers) -----------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_retrieve : mixin_base<Ref, op::retrieve_tag> {
  /// The whole block retrieves all matches of [first, last): rows {probe key, matched element} are
  /// written at positions reserved from `atomic_counter` (reference ref_impl.cuh:1009-1282).
  template <std::int32_t BlockSize,
            class InputProbeIt,
            class OutputProbeIt,
            class OutputMatchIt,
            class AtomicCounter>
  __device__ void retrieve(cooperative_groups::thread_block const&,
                           InputProbeIt input_probe_begin,
                           InputProbeIt input_probe_end,
                           OutputProbeIt output_probe,
                           OutputMatchIt output_match,
                           AtomicCounter& atomic_counter) const
  {
    run<false, BlockSize>(input_probe_begin, input_probe_end, output_probe, output_match, atomic_counter);
  }

  /// As `retrieve`; a key without matches yields {key, empty sentinel}.
  template <std::int32_t BlockSize,
            class InputProbeIt,
            class OutputProbeIt,
            class OutputMatchIt,
            class AtomicCounter>
  __device__ void retrieve_outer(cooperative_groups::thread_block const&,
                                 InputProbeIt input_probe_begin,
                                 InputProbeIt input_probe_end,
                                 OutputProbeIt output_probe,
                                 OutputMatchIt output_match,
                                 AtomicCounter& atomic_counter) const
  {
    run<true, BlockSize>(input_probe_begin, input_probe_end, output_probe, output_match, atomic_counter);
  }

 private:
  template <bool IsOuter, std::int32_t BlockSize, class In, class OutP, class OutM, class Counter>
  __device__ void run(In first, In last, OutP out_probe, OutM out_match, Counter& counter) const
  {
    auto const& e  = this->self().engine();
    using engine_t = cuda::std::remove_cv_t<cuda::std::remove_reference_t<decltype(e)>>;
    block_retrieve<IsOuter, BlockSize, engine_t::window_chunk_slots, load_policy::plain, 1>(
      e, first, static_cast<cuco::detail::index_type>(last - first), out_probe, out_match, counter);
  }
};

// ---- insert_or_assign (maps) ---------------------------------------------------------------------
template <typename Ref, int CGSize>
struct mixin_insert_or_assign : mixin_base<Ref, op::insert_or_assign_tag> {
  /// Upsert: afterwards the key maps to `value.second` (last writer wins among concurrent calls).
  template <typename Value>
  __device__ void insert_or_assign(Value const& value) noexcept
  {
    CUCO_B200_SCALAR_ONLY();
    auto& e          = this->self().engine();
    using engine_t   = cuda::std::remove_reference_t<decltype(e)>;
    using slot_t     = typename engine_t::value_type;
    auto const val   = e.heterogeneous_value(value);
    auto const image = e.native_value(val);
    auto const res   = e.template insert_driver<engine_t::window_chunk_slots, load_policy::plain>(
      val, [&](slot_t* t, slot_t& expected, auto const& v) { return e.try_claim(t, expected, image, engine_t::key_of(v)); });
    if (!res.second) { store_payload<engine_t>(res.first, image); }
  }

  template <typename Value>
  __device__ void insert_or_assign(cooperative_groups::thread_block_tile<CGSize> const& group,
                                   Value const& value) noexcept
  {
    auto& e          = this->self().engine();
    using engine_t   = cuda::std::remove_reference_t<decltype(e)>;
    using slot_t     = typename engine_t::value_type;
    auto const val   = e.heterogeneous_value(value);
    auto const image = e.native_value(val);
    auto const res   = e.tile_insert_driver(
      group, val, [&](slot_t* t, slot_t& expected, auto const& v) { return e.try_claim(t, expected, image, engine_t::key_of(v)); });
    if (!res.second && group.thread_rank() == 0) { store_payload<engine_t>(res.first, image); }
  }

 private:
  template <typename Engine, typename Slot>
  __device__ static void store_payload(Slot* slot, Slot const& image) noexcept
  {
    using mapped = decltype(slot->second);
    cuda::atomic_ref<mapped, Engine::thread_scope> ref{slot->second};
    ref.store(image.second, cuda::memory_order_relaxed);
  }
};

// ---- insert_or_apply (maps) ----------------------------------------------------------------------
template <typename Ref, typename Mapped, int CGSize>
struct mixin_insert_or_apply : mixin_base<Ref, op::insert_or_apply_tag> {
  /// Aggregating upsert: the first arrival stores its payload, every other arrival runs
  /// `op(cuda::atomic_ref<Mapped, Scope>{slot.second}, value.second)`. True iff newly inserted.
  template <typename Value, typename Op>
  __device__ bool insert_or_apply(Value const& value, Op op)
  {
    CUCO_B200_SCALAR_ONLY();
    check_op<Op>();
    return scalar_impl<false>(value, op);
  }

  /// `init` is a hint: when it equals the empty payload, first arrivals may combine onto the
  /// sentinel instead of storing (valid because init is then the identity of `op`).
  template <typename Value,
            typename Init,
            typename Op,
            typename R = Ref,
            typename   = cuda::std::enable_if_t<std::is_convertible_v<Value, typename R::value_type>>>
  __device__ bool insert_or_apply(Value const& value, Init init, Op op)
  {
    CUCO_B200_SCALAR_ONLY();
    check_op<Op>();
    if (same_bits(static_cast<Mapped>(init), this->self().empty_value_sentinel())) {
      return scalar_impl<true>(value, op);
    }
    return scalar_impl<false>(value, op);
  }

  template <typename Value, typename Op>
  __device__ bool insert_or_apply(cooperative_groups::thread_block_tile<CGSize> const& group,
                                  Value const& value,
                                  Op op)
  {
    check_op<Op>();
    return tile_impl<false>(group, value, op);
  }

  template <typename Value, typename Init, typename Op>
  __device__ bool insert_or_apply(cooperative_groups::thread_block_tile<CGSize> const& group,
                                  Value const& value,
                                  Init init,
                                  Op op)
  {
    check_op<Op>();
    if (same_bits(static_cast<Mapped>(init), this->self().empty_value_sentinel())) {
      return tile_impl<true>(group, value, op);
    }
    return tile_impl<false>(group, value, op);
  }

 private:
  template <typename Op>
  __device__ static constexpr void check_op()
  {
    static_assert(
      cuda::std::is_invocable_v<Op, cuda::atomic_ref<Mapped, Ref::thread_scope>, Mapped>,
      "insert_or_apply expects `Op` to be a callable as `Op(cuda::atomic_ref<T, Scope>, T)`");
  }

  /// Claim used by both flavours. Direct mode on 16-byte slots claims the key half only.
  template <bool Direct, typename Engine, typename Slot, typename ProbeKey>
  __device__ static auto claim(
    Engine& e, Slot* target, Slot& expected, Slot const& image, ProbeKey const& probe_key)
  {
    if constexpr (Direct && sizeof(Slot) > 8) {
      auto expected_key = expected.first;
      auto const r      = e.try_claim_key(target, expected_key, image.first, probe_key);
      expected.first    = expected_key;
      return r;
    } else {
      return e.try_claim(target, expected, image, probe_key);
    }
  }

  template <bool Direct, typename Engine, typename Slot, typename Op>
  __device__ static void settle(Engine& e, Slot* slot, Slot const& image, bool is_new, Op& op)
  {
    constexpr bool combines_on_new = Direct && sizeof(Slot) > 8;
    if (is_new) {
      if constexpr (combines_on_new) {
        op(cuda::atomic_ref<Mapped, Engine::thread_scope>{slot->second}, image.second);
      }
      return;
    }
    if constexpr (!combines_on_new && !Engine::single_cas) {
      // the creator publishes the payload after the key: do not combine with the sentinel
      e.wait_for_payload(slot->second, e.empty_slot_sentinel().second);
    }
    op(cuda::atomic_ref<Mapped, Engine::thread_scope>{slot->second}, image.second);
  }

  template <bool Direct, typename Value, typename Op>
  __device__ bool scalar_impl(Value const& value, Op& op)
  {
    auto& e          = this->self().engine();
    using engine_t   = cuda::std::remove_reference_t<decltype(e)>;
    using slot_t     = typename engine_t::value_type;
    auto const val   = e.heterogeneous_value(value);
    auto const image = e.native_value(val);
    auto const res   = e.template insert_driver<engine_t::window_chunk_slots, load_policy::plain>(
      val, [&](slot_t* t, slot_t& expected, auto const& v) {
        return claim<Direct>(e, t, expected, image, engine_t::key_of(v));
      });
    settle<Direct>(e, res.first, image, res.second, op);
    return res.second;
  }

  template <bool Direct, typename Tile, typename Value, typename Op>
  __device__ bool tile_impl(Tile const& group, Value const& value, Op& op)
  {
    auto& e          = this->self().engine();
    using engine_t   = cuda::std::remove_reference_t<decltype(e)>;
    using slot_t     = typename engine_t::value_type;
    auto const val   = e.heterogeneous_value(value);
    auto const image = e.native_value(val);
    auto const res   = e.tile_insert_driver(
      group, val, [&](slot_t* t, slot_t& expected, auto const& v) {
        return claim<Direct>(e, t, expected, image, engine_t::key_of(v));
      });
    if (group.thread_rank() == 0) { settle<Direct>(e, res.first, image, res.second, op); }
    return res.second;
  }
};

#undef CUCO_B200_SCALAR_ONLY

}  // namespace cuco::b200
