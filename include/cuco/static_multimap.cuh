// cuco::experimental::static_multimap — fixed-capacity GPU hash multimap (equal keys may map to any
// number of payloads).
//
// Drop-in for the reference's new-style class template (include/cuco/static_multimap.cuh:45-549,
// detail/static_multimap/static_multimap.inl): same template parameters and defaults (linear probing
// with a tile of 4, xxhash_32, one slot per window), constructors, stream-ordered bulk API (insert,
// insert_if, contains, contains_if, count) and `ref(ops...)` with insert / contains / count /
// for_each. It runs on cuco::b200::table_engine with AllowsDuplicates = true, like static_multiset.
// The legacy `cuco::static_multimap` (device_view API, pair_retrieve ...) is out of scope.
#pragma once

#include <cuco/b200/bulk_engine.cuh>
#include <cuco/detail/__config>
#include <cuco/extent.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/pair.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/static_multimap_ref.cuh>
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
namespace experimental {

template <class Key,
          class T,
          class Extent             = cuco::extent<std::size_t>,
          cuda::thread_scope Scope = cuda::thread_scope_device,
          class KeyEqual           = thrust::equal_to<Key>,
          class ProbingScheme      = cuco::linear_probing<4, cuco::default_hash_function<Key>>,
          class Allocator          = cuco::cuda_allocator<cuco::pair<Key, T>>,
          class Storage            = cuco::storage<1>>
class static_multimap {
  static_assert(sizeof(Key) <= 8, "Container does not support key types larger than 8 bytes.");
  static_assert(sizeof(T) <= 8, "Container does not support payload types larger than 8 bytes.");
  static_assert(cuco::is_bitwise_comparable_v<T>,
                "Mapped type must have unique object representations or have been explicitly "
                "declared as safe for bitwise comparison via specialization of "
                "cuco::is_bitwise_comparable_v<T>.");

  using impl_type =
    b200::table_engine<Key, cuco::pair<Key, T>, Extent, Scope, KeyEqual, ProbingScheme, Allocator, Storage, true>;

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
  using ref_type = cuco::static_multimap_ref<key_type,
                                        mapped_type,
                                        thread_scope,
                                        key_equal,
                                        probing_scheme_type,
                                        storage_ref_type,
                                        Operators...>;

  static_multimap(static_multimap const&)            = delete;
  static_multimap& operator=(static_multimap const&) = delete;
  static_multimap(static_multimap&&)                 = default;
  static_multimap& operator=(static_multimap&&)      = default;
  ~static_multimap()                            = default;

  /// Table with at least `capacity` slots (rounded up to a valid extent), all empty.
  constexpr static_multimap(Extent capacity,
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
  constexpr static_multimap(Extent n,
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
  constexpr static_multimap(Extent capacity,
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
  /// Inserts every pair of [first, last) (equal keys are kept side by side). Synchronises `stream`.
  template <typename InputIt>
  void insert(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    impl_->insert_async(first, last, ref(op::insert), stream);
    stream.wait();
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

  /// Total number of stored pairs whose key matches a key of [first, last). Synchronises `stream`.
  template <typename InputIt>
  size_type count(InputIt first, InputIt last, cuda::stream_ref stream = {}) const
  {
    return impl_->template count<false>(first, last, ref(op::count), stream);
  }

  /// b200 extension (the reference class has no `size`): stored pairs, duplicates included.
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

}  // namespace experimental
}  // namespace cuco
