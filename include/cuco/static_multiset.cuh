// cuco::static_multiset — fixed-capacity GPU hash multiset (equal keys may be stored repeatedly).
//
// Drop-in for the reference class template (include/cuco/static_multiset.cuh:81-729,
// detail/static_multiset/static_multiset.inl:25-525): same template parameters and defaults (double
// hashing with a tile of 4, xxhash_32, two slots per window), constructors, stream-ordered bulk API
// (insert, insert_if, contains, contains_if, find, retrieve, retrieve_outer, count, count_outer,
// size) and `ref(ops...)`. It runs on the same cuco::b200::table_engine as static_set with
// AllowsDuplicates = true: the insert kernels claim the first free slot of the probe sequence
// without comparing keys, `count` / `retrieve` enumerate matches up to the first empty slot
// (cuco/b200/match_kernels.cuh).
#pragma once

#include <cuco/b200/bulk_engine.cuh>
#include <cuco/detail/__config>
#include <cuco/extent.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/static_multiset_ref.cuh>
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
          class Storage            = cuco::storage<2>>
class static_multiset {
  using impl_type =
    b200::table_engine<Key, Key, Extent, Scope, KeyEqual, ProbingScheme, Allocator, Storage, true>;

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
  using ref_type = cuco::static_multiset_ref<key_type,
                                        thread_scope,
                                        key_equal,
                                        probing_scheme_type,
                                        storage_ref_type,
                                        Operators...>;

  static_multiset(static_multiset const&)            = delete;
  static_multiset& operator=(static_multiset const&) = delete;
  static_multiset(static_multiset&&)                 = default;
  static_multiset& operator=(static_multiset&&)      = default;
  ~static_multiset()                            = default;

  /// Multiset with at least `capacity` slots (rounded up to a valid extent), all empty.
  constexpr static_multiset(Extent capacity,
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

  /// Multiset sized for `n` keys at `desired_load_factor` in (0, 1].
  constexpr static_multiset(Extent n,
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

  /// With an erased-key sentinel (kept for interface parity; the multiset has no erase).
  constexpr static_multiset(Extent capacity,
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
  /// Inserts every element of [first, last) (duplicates included). Synchronises `stream`.
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

  // ---- all matches: retrieve / count -----------------------------------------------------------
  /// For every key k of [first, last) and every stored element m equal to it, writes k to
  /// `output_probe` and m to `output_match` (same position, unspecified order). Returns the ends of
  /// both outputs; size them with `count`. Synchronises `stream`.
  template <class InputProbeIt, class OutputProbeIt, class OutputMatchIt>
  std::pair<OutputProbeIt, OutputMatchIt> retrieve(InputProbeIt first,
                                                   InputProbeIt last,
                                                   OutputProbeIt output_probe,
                                                   OutputMatchIt output_match,
                                                   cuda::stream_ref stream = {}) const
  {
    auto const rows = impl_->template retrieve<false>(
      first, last, output_probe, output_match, ref(op::retrieve), stream);
    return {output_probe + rows, output_match + rows};
  }

  /// `retrieve` with a custom probe-key equality and hasher (heterogeneous probes).
  template <class InputProbeIt,
            class ProbeEqual,
            class ProbeHash,
            class OutputProbeIt,
            class OutputMatchIt>
  std::pair<OutputProbeIt, OutputMatchIt> retrieve(InputProbeIt first,
                                                   InputProbeIt last,
                                                   ProbeEqual const& probe_equal,
                                                   ProbeHash const& probe_hash,
                                                   OutputProbeIt output_probe,
                                                   OutputMatchIt output_match,
                                                   cuda::stream_ref stream = {}) const
  {
    auto const probe_ref =
      ref(op::retrieve).rebind_key_eq(probe_equal).rebind_hash_function(probe_hash);
    auto const rows =
      impl_->template retrieve<false>(first, last, output_probe, output_match, probe_ref, stream);
    return {output_probe + rows, output_match + rows};
  }

  /// As `retrieve`, but a key without matches yields the row {k, empty key sentinel}; size the
  /// outputs with `count_outer`.
  template <class InputProbeIt,
            class ProbeEqual,
            class ProbeHash,
            class OutputProbeIt,
            class OutputMatchIt>
  std::pair<OutputProbeIt, OutputMatchIt> retrieve_outer(InputProbeIt first,
                                                         InputProbeIt last,
                                                         ProbeEqual const& probe_equal,
                                                         ProbeHash const& probe_hash,
                                                         OutputProbeIt output_probe,
                                                         OutputMatchIt output_match,
                                                         cuda::stream_ref stream = {}) const
  {
    auto const probe_ref =
      ref(op::retrieve).rebind_key_eq(probe_equal).rebind_hash_function(probe_hash);
    auto const rows =
      impl_->template retrieve<true>(first, last, output_probe, output_match, probe_ref, stream);
    return {output_probe + rows, output_match + rows};
  }

  /// Total number of stored elements matching the keys of [first, last). Synchronises `stream`.
  template <typename InputIt>
  size_type count(InputIt first, InputIt last, cuda::stream_ref stream = {}) const
  {
    return impl_->template count<false>(first, last, ref(op::count), stream);
  }

  template <typename InputIt, typename ProbeKeyEqual, typename ProbeHash>
  size_type count(InputIt first,
                  InputIt last,
                  ProbeKeyEqual const& probe_key_equal,
                  ProbeHash const& probe_hash,
                  cuda::stream_ref stream = {}) const
  {
    return impl_->template count<false>(
      first,
      last,
      ref(op::count).rebind_key_eq(probe_key_equal).rebind_hash_function(probe_hash),
      stream);
  }

  /// As `count`, but a key without matches counts as one.
  template <typename InputIt, typename ProbeKeyEqual, typename ProbeHash>
  size_type count_outer(InputIt first,
                        InputIt last,
                        ProbeKeyEqual const& probe_key_equal,
                        ProbeHash const& probe_hash,
                        cuda::stream_ref stream = {}) const
  {
    return impl_->template count<true>(
      first,
      last,
      ref(op::count).rebind_key_eq(probe_key_equal).rebind_hash_function(probe_hash),
      stream);
  }

  /// Number of stored elements, duplicates included (full scan). Synchronises `stream`.
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
