// cuco::static_map — fixed-capacity GPU hash map with unique keys, open addressing.
//
// Drop-in for the reference class template (include/cuco/static_map.cuh:88-986,
// detail/static_map/static_map.inl:40-770): same template parameters and defaults, the same three
// constructors, the same stream-ordered bulk API and the same `ref(ops...)` device handle. What is
// behind it is new: the bulk calls run the sm_100a kernels of cuco/b200/bulk_kernels.cuh through
// cuco::b200::table_engine (persistent grids, several keys in flight per thread, sector-wide table
// loads, single 64/128-bit CAS slot claims).
//
// Semantics that parity is checked on (SURVEY.md §8a'): insert keeps one unspecified element per
// distinct key and returns the number of new keys; find writes the payload or the empty value
// sentinel; contains writes bool; insert_and_find reports the resident payload plus whether this
// element created the entry; insert_or_assign overwrites; insert_or_apply folds `op` over all
// elements of a key; size counts filled slots. Overfilling the table or inserting a sentinel key is
// undefined, as in the reference.
#pragma once

#include <cuco/b200/bulk_engine.cuh>
#include <cuco/b200/table_scan.cuh>
#include <cuco/detail/__config>
#include <cuco/detail/static_map/kernels.cuh>
#include <cuco/extent.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/pair.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/static_map_ref.cuh>
#include <cuco/storage.cuh>
#include <cuco/types.cuh>
#include <cuco/utility/allocator.hpp>
#include <cuco/utility/cuda_thread_scope.cuh>
#include <cuco/utility/reduction_functors.cuh>
#include <cuco/utility/traits.hpp>

#include <cuda/atomic>
#include <cuda/stream_ref>
#include <thrust/functional.h>

#include <cstddef>
#include <memory>
#include <utility>

namespace cuco {

template <class Key,
          class T,
          class Extent             = cuco::extent<std::size_t>,
          cuda::thread_scope Scope = cuda::thread_scope_device,
          class KeyEqual           = thrust::equal_to<Key>,
          class ProbingScheme      = cuco::linear_probing<4, cuco::default_hash_function<Key>>,
          class Allocator          = cuco::cuda_allocator<cuco::pair<Key, T>>,
          class Storage            = cuco::storage<1>>
class static_map {
  static_assert(sizeof(Key) <= 8, "Container does not support key types larger than 8 bytes.");
  static_assert(sizeof(T) <= 8, "Container does not support payload types larger than 8 bytes.");
  static_assert(cuco::is_bitwise_comparable_v<T>,
                "Mapped type must have unique object representations or have been explicitly "
                "declared as safe for bitwise comparison via specialization of "
                "cuco::is_bitwise_comparable_v<T>.");

  using impl_type =
    b200::table_engine<Key, cuco::pair<Key, T>, Extent, Scope, KeyEqual, ProbingScheme, Allocator, Storage>;

 public:
  static constexpr auto cg_size      = impl_type::cg_size;
  static constexpr auto window_size  = impl_type::window_size;
  static constexpr auto thread_scope = impl_type::thread_scope;

  using key_type            = typename impl_type::key_type;
  using value_type          = typename impl_type::value_type;  ///< cuco::pair<Key, T>
  using extent_type         = typename impl_type::extent_type;
  using size_type           = typename impl_type::size_type;
  using key_equal           = typename impl_type::key_equal;
  using allocator_type      = typename impl_type::allocator_type;
  using storage_ref_type    = typename impl_type::storage_ref_type;
  using probing_scheme_type = typename impl_type::probing_scheme_type;
  using hasher              = typename probing_scheme_type::hasher;
  using mapped_type         = T;

  template <typename... Operators>
  using ref_type = cuco::static_map_ref<key_type,
                                        mapped_type,
                                        thread_scope,
                                        key_equal,
                                        probing_scheme_type,
                                        storage_ref_type,
                                        Operators...>;

  static_map(static_map const&)            = delete;
  static_map& operator=(static_map const&) = delete;
  static_map(static_map&&)                 = default;
  static_map& operator=(static_map&&)      = default;
  ~static_map()                            = default;

  /// Table with at least `capacity` slots (rounded up to a valid extent), all empty.
  constexpr static_map(Extent capacity,
                       empty_key<Key> empty_key_sentinel,
                       empty_value<T> empty_value_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(capacity,
                                        cuco::pair<Key, T>{empty_key_sentinel, empty_value_sentinel},
                                        pred,
                                        probing_scheme,
                                        alloc,
                                        stream)},
      empty_value_sentinel_{empty_value_sentinel}
  {
  }

  /// Table sized for `n` keys at `desired_load_factor` in (0, 1].
  constexpr static_map(Extent n,
                       double desired_load_factor,
                       empty_key<Key> empty_key_sentinel,
                       empty_value<T> empty_value_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(n,
                                        desired_load_factor,
                                        cuco::pair<Key, T>{empty_key_sentinel, empty_value_sentinel},
                                        pred,
                                        probing_scheme,
                                        alloc,
                                        stream)},
      empty_value_sentinel_{empty_value_sentinel}
  {
  }

  /// Table that supports erase: `erased_key_sentinel` marks tombstones and must differ from empty.
  constexpr static_map(Extent capacity,
                       empty_key<Key> empty_key_sentinel,
                       empty_value<T> empty_value_sentinel,
                       erased_key<Key> erased_key_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(capacity,
                                        cuco::pair<Key, T>{empty_key_sentinel, empty_value_sentinel},
                                        erased_key_sentinel,
                                        pred,
                                        probing_scheme,
                                        alloc,
                                        stream)},
      empty_value_sentinel_{empty_value_sentinel}
  {
  }

  void clear(cuda::stream_ref stream = {}) { impl_->clear(stream); }
  void clear_async(cuda::stream_ref stream = {}) noexcept { impl_->clear_async(stream); }

  // ---- insert ----------------------------------------------------------------------------------
  /// Inserts [first, last); returns how many keys were new. Synchronises `stream`.
  template <typename InputIt>
  size_type insert(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    return impl_->insert(first, last, ref(op::insert), stream);
  }

  template <typename InputIt>
  void insert_async(InputIt first, InputIt last, cuda::stream_ref stream = {}) noexcept
  {
    impl_->insert_async(first, last, ref(op::insert), stream);
  }

  /// Inserts element i only if `pred(stencil[i])`.
  template <typename InputIt, typename StencilIt, typename Predicate>
  size_type insert_if(
    InputIt first, InputIt last, StencilIt stencil, Predicate pred, cuda::stream_ref stream = {})
  {
    return impl_->insert_if(first, last, stencil, pred, ref(op::insert), stream);
  }

  template <typename InputIt, typename StencilIt, typename Predicate>
  void insert_if_async(InputIt first,
                       InputIt last,
                       StencilIt stencil,
                       Predicate pred,
                       cuda::stream_ref stream = {}) noexcept
  {
    impl_->insert_if_async(first, last, stencil, pred, ref(op::insert), stream);
  }

  /// For each element writes the payload now stored under its key and whether it created the entry.
  template <typename InputIt, typename FoundIt, typename InsertedIt>
  void insert_and_find_async(InputIt first,
                             InputIt last,
                             FoundIt found_begin,
                             InsertedIt inserted_begin,
                             cuda::stream_ref stream = {}) noexcept
  {
    impl_->insert_and_find_async(
      first, last, found_begin, inserted_begin, ref(op::insert_and_find), stream);
  }

  template <typename InputIt, typename FoundIt, typename InsertedIt>
  void insert_and_find(InputIt first,
                       InputIt last,
                       FoundIt found_begin,
                       InsertedIt inserted_begin,
                       cuda::stream_ref stream = {})
  {
    insert_and_find_async(first, last, found_begin, inserted_begin, stream);
    stream.wait();
  }

  // ---- upserts ---------------------------------------------------------------------------------
  template <typename InputIt>
  void insert_or_assign(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    insert_or_assign_async(first, last, stream);
    stream.wait();
  }

  template <typename InputIt>
  void insert_or_assign_async(InputIt first, InputIt last, cuda::stream_ref stream = {}) noexcept
  {
    impl_->insert_or_assign_async(first, last, ref(op::insert_or_assign), stream);
  }

  template <typename InputIt, typename Op>
  void insert_or_apply(InputIt first, InputIt last, Op op, cuda::stream_ref stream = {})
  {
    insert_or_apply_async(first, last, op, stream);
    stream.wait();
  }

  template <typename InputIt, typename Init, typename Op>
  void insert_or_apply(InputIt first, InputIt last, Init init, Op op, cuda::stream_ref stream = {})
  {
    insert_or_apply_async(first, last, init, op, stream);
    stream.wait();
  }

  /// payload[key] = fold of `op` over the payloads of all elements carrying `key`.
  template <typename InputIt, typename Op>
  void insert_or_apply_async(InputIt first,
                             InputIt last,
                             Op op,
                             cuda::stream_ref stream = {}) noexcept
  {
    impl_->insert_or_apply_async(first, last, false, op, ref(op::insert_or_apply), stream);
  }

  /// `init` is the identity of `op`; when it equals the empty payload the table combines in place.
  template <typename InputIt,
            typename Init,
            typename Op,
            typename = std::enable_if_t<std::is_convertible_v<Init, T>>>
  void insert_or_apply_async(
    InputIt first, InputIt last, Init init, Op op, cuda::stream_ref stream = {}) noexcept
  {
    bool const direct =
      b200::same_bits(static_cast<T>(init), static_cast<T>(empty_value_sentinel_));
    impl_->insert_or_apply_async(first, last, direct, op, ref(op::insert_or_apply), stream);
  }

  // ---- erase -----------------------------------------------------------------------------------
  template <typename InputIt>
  void erase(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    erase_async(first, last, stream);
    stream.wait();
  }

  template <typename InputIt>
  void erase_async(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    impl_->erase_async(first, last, ref(op::erase), stream);
  }

  // ---- lookups ---------------------------------------------------------------------------------
  template <typename InputIt, typename OutputIt>
  void contains(InputIt first,
                InputIt last,
                OutputIt output_begin,
                cuda::stream_ref stream = {}) const
  {
    contains_async(first, last, output_begin, stream);
    stream.wait();
  }

  template <typename InputIt, typename OutputIt>
  void contains_async(InputIt first,
                      InputIt last,
                      OutputIt output_begin,
                      cuda::stream_ref stream = {}) const noexcept
  {
    impl_->contains_async(first, last, output_begin, ref(op::contains), stream);
  }

  template <typename InputIt, typename StencilIt, typename Predicate, typename OutputIt>
  void contains_if(InputIt first,
                   InputIt last,
                   StencilIt stencil,
                   Predicate pred,
                   OutputIt output_begin,
                   cuda::stream_ref stream = {}) const
  {
    contains_if_async(first, last, stencil, pred, output_begin, stream);
    stream.wait();
  }

  template <typename InputIt, typename StencilIt, typename Predicate, typename OutputIt>
  void contains_if_async(InputIt first,
                         InputIt last,
                         StencilIt stencil,
                         Predicate pred,
                         OutputIt output_begin,
                         cuda::stream_ref stream = {}) const noexcept
  {
    impl_->contains_if_async(first, last, stencil, pred, output_begin, ref(op::contains), stream);
  }

  /// Writes the payload of each key, or the empty value sentinel when absent.
  template <typename InputIt, typename OutputIt>
  void find(InputIt first, InputIt last, OutputIt output_begin, cuda::stream_ref stream = {}) const
  {
    find_async(first, last, output_begin, stream);
    stream.wait();
  }

  template <typename InputIt, typename OutputIt>
  void find_async(InputIt first,
                  InputIt last,
                  OutputIt output_begin,
                  cuda::stream_ref stream = {}) const
  {
    impl_->find_async(first, last, output_begin, ref(op::find), stream);
  }

  // ---- whole-table operations ------------------------------------------------------------------
  /// Applies `callback_op(slot)` to every filled slot.
  template <typename CallbackOp>
  void for_each(CallbackOp&& callback_op, cuda::stream_ref stream = {}) const
  {
    for_each_async(std::forward<CallbackOp>(callback_op), stream);
    stream.wait();
  }

  template <typename CallbackOp>
  void for_each_async(CallbackOp&& callback_op, cuda::stream_ref stream = {}) const
  {
    b200::for_each_filled_async(impl_->make_engine(), callback_op, stream);
  }

  /// Applies `callback_op(slot)` to the entry of each key in [first, last) that is present.
  template <typename InputIt, typename CallbackOp>
  void for_each(InputIt first,
                InputIt last,
                CallbackOp&& callback_op,
                cuda::stream_ref stream = {}) const
  {
    for_each_async(first, last, std::forward<CallbackOp>(callback_op), stream);
    stream.wait();
  }

  template <typename InputIt, typename CallbackOp>
  void for_each_async(InputIt first,
                      InputIt last,
                      CallbackOp&& callback_op,
                      cuda::stream_ref stream = {}) const noexcept
  {
    b200::for_each_key_async(impl_->make_engine(), first, last, callback_op, stream);
  }

  /// Copies all (key, payload) pairs out, in unspecified order; returns the output ends.
  template <typename KeyOut, typename ValueOut>
  std::pair<KeyOut, ValueOut> retrieve_all(KeyOut keys_out,
                                           ValueOut values_out,
                                           cuda::stream_ref stream = {}) const
  {
    auto const n = b200::retrieve_all_pairs(impl_->make_engine(), keys_out, values_out, stream);
    return {keys_out + n, values_out + n};
  }

  /// Rebuilds the table in place (drops tombstones).
  void rehash(cuda::stream_ref stream = {})
  {
    rehash_async(stream);
    stream.wait();
  }

  /// Rebuilds the table with at least `capacity` slots.
  void rehash(size_type capacity, cuda::stream_ref stream = {})
  {
    rehash_async(capacity, stream);
    stream.wait();
  }

  void rehash_async(cuda::stream_ref stream = {})
  {
    b200::rehash_into(*impl_, impl_->storage_ref().window_extent(), ref(op::insert), stream);
  }

  void rehash_async(size_type capacity, cuda::stream_ref stream = {})
  {
    auto const extent = make_window_extent<static_map>(capacity);
    b200::rehash_into(*impl_, extent, ref(op::insert), stream);
  }

  /// Number of entries (full scan). Synchronises `stream`.
  [[nodiscard]] size_type size(cuda::stream_ref stream = {}) const { return impl_->size(stream); }

  [[nodiscard]] constexpr auto capacity() const noexcept { return impl_->capacity(); }
  [[nodiscard]] constexpr key_type empty_key_sentinel() const noexcept
  {
    return impl_->empty_key_sentinel();
  }
  [[nodiscard]] constexpr mapped_type empty_value_sentinel() const noexcept
  {
    return empty_value_sentinel_;
  }
  [[nodiscard]] constexpr key_type erased_key_sentinel() const noexcept
  {
    return impl_->erased_key_sentinel();
  }
  [[nodiscard]] constexpr key_equal key_eq() const noexcept { return impl_->key_eq(); }
  [[nodiscard]] constexpr hasher hash_function() const noexcept { return impl_->hash_function(); }

  /// Device handle exposing the requested operators, e.g. `map.ref(cuco::insert, cuco::find)`.
  template <typename... Operators>
  [[nodiscard]] auto ref(Operators...) const noexcept
  {
    static_assert(sizeof...(Operators), "No operators specified");
    return ref_type<Operators...>{cuco::empty_key<key_type>(this->empty_key_sentinel()),
                                  cuco::empty_value<mapped_type>(this->empty_value_sentinel()),
                                  cuco::erased_key<key_type>(this->erased_key_sentinel()),
                                  impl_->key_eq(),
                                  impl_->probing_scheme(),
                                  cuda_thread_scope<Scope>{},
                                  impl_->storage_ref()};
  }

  /// b200 extension (no reference counterpart): the engine behind this container, used by the
  /// exchange path of hash-partitioned multi-GPU tables (include/cuco/b200/bulk_engine.cuh).
  [[nodiscard]] impl_type& b200_engine() noexcept { return *impl_; }
  [[nodiscard]] impl_type const& b200_engine() const noexcept { return *impl_; }

 private:
  std::unique_ptr<impl_type> impl_;
  mapped_type empty_value_sentinel_;
};

}  // namespace cuco
