// cuco::static_set — fixed-capacity GPU hash set with unique keys, open addressing.
//
// Drop-in for the reference class template (include/cuco/static_set.cuh:82-798,
// detail/static_set/static_set.inl:35-569): same template parameters and defaults (double hashing
// with a tile of 4, xxhash_32, one slot per window), constructors, stream-ordered bulk API and
// `ref(ops...)`. A slot is the key itself, so claiming is a single 32/64-bit CAS; the bulk calls run
// the sm_100a kernels of cuco/b200/bulk_kernels.cuh via cuco::b200::table_engine.
//
// Result semantics: insert returns the number of new keys; find writes the stored key or the empty
// key sentinel; contains writes bool; insert_and_find writes the resident key and whether this
// element created the entry; size counts filled slots.
#pragma once

#include <cuco/b200/bulk_engine.cuh>
#include <cuco/b200/table_scan.cuh>
#include <cuco/detail/__config>
#include <cuco/extent.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/static_set_ref.cuh>
#include <cuco/storage.cuh>
#include <cuco/types.cuh>
#include <cuco/utility/allocator.hpp>
#include <cuco/utility/cuda_thread_scope.cuh>
#include <cuco/utility/traits.hpp>

#include <cuda/atomic>
#include <cuda/stream_ref>
#include <thrust/functional.h>

#include <cstddef>
#include <memory>
#include <utility>

namespace cuco {

template <class Key,
          class Extent             = cuco::extent<std::size_t>,
          cuda::thread_scope Scope = cuda::thread_scope_device,
          class KeyEqual           = thrust::equal_to<Key>,
          class ProbingScheme      = cuco::double_hashing<4, cuco::default_hash_function<Key>>,
          class Allocator          = cuco::cuda_allocator<Key>,
          class Storage            = cuco::storage<1>>
class static_set {
  using impl_type =
    b200::table_engine<Key, Key, Extent, Scope, KeyEqual, ProbingScheme, Allocator, Storage>;

 public:
  static constexpr auto cg_size      = impl_type::cg_size;
  static constexpr auto window_size  = impl_type::window_size;
  static constexpr auto thread_scope = impl_type::thread_scope;

  using key_type            = typename impl_type::key_type;
  using value_type          = typename impl_type::value_type;  ///< == key_type
  using extent_type         = typename impl_type::extent_type;
  using size_type           = typename impl_type::size_type;
  using key_equal           = typename impl_type::key_equal;
  using allocator_type      = typename impl_type::allocator_type;
  using storage_ref_type    = typename impl_type::storage_ref_type;
  using probing_scheme_type = typename impl_type::probing_scheme_type;
  using hasher              = typename probing_scheme_type::hasher;

  template <typename... Operators>
  using ref_type = cuco::static_set_ref<key_type,
                                        thread_scope,
                                        key_equal,
                                        probing_scheme_type,
                                        storage_ref_type,
                                        Operators...>;

  static_set(static_set const&)            = delete;
  static_set& operator=(static_set const&) = delete;
  static_set(static_set&&)                 = default;
  static_set& operator=(static_set&&)      = default;
  ~static_set()                            = default;

  /// Set with at least `capacity` slots (rounded up to a valid extent), all empty.
  constexpr static_set(Extent capacity,
                       empty_key<Key> empty_key_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(
        capacity, static_cast<Key>(empty_key_sentinel), pred, probing_scheme, alloc, stream)}
  {
  }

  /// Set sized for `n` keys at `desired_load_factor` in (0, 1].
  constexpr static_set(Extent n,
                       double desired_load_factor,
                       empty_key<Key> empty_key_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(n,
                                        desired_load_factor,
                                        static_cast<Key>(empty_key_sentinel),
                                        pred,
                                        probing_scheme,
                                        alloc,
                                        stream)}
  {
  }

  /// Set that supports erase: `erased_key_sentinel` marks tombstones and must differ from empty.
  constexpr static_set(Extent capacity,
                       empty_key<Key> empty_key_sentinel,
                       erased_key<Key> erased_key_sentinel,
                       KeyEqual const& pred                = {},
                       ProbingScheme const& probing_scheme = {},
                       cuda_thread_scope<Scope>            = {},
                       Storage                             = {},
                       Allocator const& alloc              = {},
                       cuda::stream_ref stream             = {})
    : impl_{std::make_unique<impl_type>(capacity,
                                        static_cast<Key>(empty_key_sentinel),
                                        static_cast<Key>(erased_key_sentinel),
                                        pred,
                                        probing_scheme,
                                        alloc,
                                        stream)}
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

  /// For each element writes the key now stored for it and whether it created the entry.
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

  /// Writes the stored key equal to each query, or the empty key sentinel when absent.
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

  /// Copies all keys out, in unspecified order; returns the output end.
  template <typename OutputIt>
  OutputIt retrieve_all(OutputIt output_begin, cuda::stream_ref stream = {}) const
  {
    return output_begin + b200::retrieve_all_elements(impl_->make_engine(), output_begin, stream);
  }

  void rehash(cuda::stream_ref stream = {})
  {
    rehash_async(stream);
    stream.wait();
  }

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
    auto const extent = make_window_extent<static_set>(capacity);
    b200::rehash_into(*impl_, extent, ref(op::insert), stream);
  }

  /// Number of keys (full scan). Synchronises `stream`.
  [[nodiscard]] size_type size(cuda::stream_ref stream = {}) const { return impl_->size(stream); }

  [[nodiscard]] constexpr auto capacity() const noexcept { return impl_->capacity(); }
  [[nodiscard]] constexpr key_type empty_key_sentinel() const noexcept
  {
    return impl_->empty_key_sentinel();
  }
  [[nodiscard]] constexpr key_type erased_key_sentinel() const noexcept
  {
    return impl_->erased_key_sentinel();
  }
  [[nodiscard]] constexpr key_equal key_eq() const noexcept { return impl_->key_eq(); }
  [[nodiscard]] constexpr hasher hash_function() const noexcept { return impl_->hash_function(); }

  /// Device handle exposing the requested operators, e.g. `set.ref(cuco::insert, cuco::contains)`.
  template <typename... Operators>
  [[nodiscard]] auto ref(Operators...) const noexcept
  {
    static_assert(sizeof...(Operators), "No operators specified");
    return ref_type<Operators...>{cuco::empty_key<key_type>(this->empty_key_sentinel()),
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
};

}  // namespace cuco
