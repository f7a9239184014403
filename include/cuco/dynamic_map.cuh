// cuco::experimental::dynamic_map — a map that grows by chaining static_maps.
//
// Counterpart of the reference's new-style class (include/cuco/dynamic_map.cuh:51-170,
// detail/dynamic_map/dynamic_map.inl:41-170): same template parameters and defaults, constructor,
// `reserve`, `insert`, `contains`. Like there it is host logic over the bulk API of
// cuco::static_map - the sm_100a kernels are the ones static_map launches:
//   * submap i + 1 has twice the capacity of submap i; a submap is filled up to 60 % and only takes
//     a batch slice of at least 10 000 elements (same constants as the reference);
//   * `insert` cuts the batch into slices, one per submap with room, in submap order, and adds the
//     number of new keys each slice reports to `size()`; a key is not looked up in earlier submaps
//     first (neither does the reference), so the class is meant for streams of distinct keys.
// Two deliberate differences from the reference, both where its behaviour is undefined or wrong:
//   * `reserve` counts the remaining elements in signed arithmetic (the reference subtracts a float
//     from an unsigned count and relies on the wrap-around of a negative float -> unsigned cast);
//   * `contains` asks every submap about every key (a key is present if any submap holds it). The
//     reference checks the i-th slice of the QUERY range against submap i only, which is right only
//     when the queries arrive in insertion order; on such inputs both give the same answers.
// The legacy `cuco::dynamic_map` (device views of the legacy static_map) is out of scope.
#pragma once

#include <cuco/detail/error.hpp>
#include <cuco/detail/utils.hpp>
#include <cuco/hash_functions.cuh>
#include <cuco/static_map.cuh>
#include <cuco/types.cuh>

#include <cuda/stream_ref>
#include <thrust/functional.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace cuco {
namespace experimental {
namespace detail {

/// out[i] = out[i] || more[i]
template <typename OutputIt>
CUCO_KERNEL void merge_presence(OutputIt out, bool const* more, cuco::detail::index_type n)
{
  for (cuco::detail::index_type i = cuco::detail::global_thread_id(); i < n;
       i += cuco::detail::grid_stride()) {
    if (more[i]) { *(out + i) = true; }
  }
}

}  // namespace detail

template <class Key,
          class T,
          class Extent             = cuco::extent<std::size_t>,
          cuda::thread_scope Scope = cuda::thread_scope_device,
          class KeyEqual           = thrust::equal_to<Key>,
          class ProbingScheme      = cuco::linear_probing<4, cuco::default_hash_function<Key>>,
          class Allocator          = cuco::cuda_allocator<cuco::pair<Key, T>>,
          class Storage            = cuco::storage<1>>
class dynamic_map {
  using map_type = static_map<Key, T, Extent, Scope, KeyEqual, ProbingScheme, Allocator, Storage>;

 public:
  static constexpr auto thread_scope = map_type::thread_scope;

  using key_type    = typename map_type::key_type;
  using value_type  = typename map_type::value_type;
  using size_type   = typename map_type::size_type;
  using key_equal   = typename map_type::key_equal;
  using mapped_type = T;

  dynamic_map(dynamic_map const&)            = delete;
  dynamic_map& operator=(dynamic_map const&) = delete;
  dynamic_map(dynamic_map&&)                 = default;
  dynamic_map& operator=(dynamic_map&&)      = default;
  ~dynamic_map()                             = default;

  /// One submap of `initial_capacity` slots; later submaps double in size.
  constexpr dynamic_map(Extent initial_capacity,
                        empty_key<Key> empty_key_sentinel,
                        empty_value<T> empty_value_sentinel,
                        KeyEqual const& pred                = {},
                        ProbingScheme const& probing_scheme = {},
                        cuda_thread_scope<Scope> scope      = {},
                        Storage storage                     = {},
                        Allocator const& alloc              = {},
                        cuda::stream_ref stream             = {})
    : size_{0},
      capacity_{static_cast<size_type>(initial_capacity)},
      min_insert_size_{10'000},
      max_load_factor_{0.60f},
      alloc_{alloc}
  {
    submaps_.push_back(std::make_unique<map_type>(initial_capacity,
                                                  empty_key_sentinel,
                                                  empty_value_sentinel,
                                                  pred,
                                                  probing_scheme,
                                                  scope,
                                                  storage,
                                                  alloc,
                                                  stream));
  }

  /// Makes sure the chain of submaps can take `n` elements in total.
  void reserve(size_type n, cuda::stream_ref stream)
  {
    auto remaining = static_cast<std::int64_t>(n);
    for (std::size_t i = 0; remaining > 0; ++i) {
      if (i == submaps_.size()) { this->add_submap(stream); }
      remaining -= this->usable(submaps_[i]->capacity()) - static_cast<std::int64_t>(min_insert_size_);
    }
  }

  /// Inserts [first, last): slice by slice into the submaps that still have room.
  template <typename InputIt>
  void insert(InputIt first, InputIt last, cuda::stream_ref stream = {})
  {
    auto to_insert = cuco::detail::distance(first, last);
    this->reserve(size_ + static_cast<size_type>(to_insert), stream);
    for (std::size_t i = 0; to_insert > 0; ++i) {
      if (i == submaps_.size()) { this->add_submap(stream); }
      auto& submap    = *submaps_[i];
      auto const room = this->usable(submap.capacity()) - static_cast<std::int64_t>(submap.size(stream));
      if (room < static_cast<std::int64_t>(min_insert_size_)) { continue; }
      auto const n = std::min<cuco::detail::index_type>(room, to_insert);
      size_ += submap.insert(first, first + n, stream);
      first += n;
      to_insert -= n;
    }
  }

  /// out[i] = key i is stored in any submap. Synchronises `stream`.
  template <typename InputIt, typename OutputIt>
  void contains(InputIt first, InputIt last, OutputIt output_begin, cuda::stream_ref stream = {}) const
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    submaps_.front()->contains_async(first, last, output_begin, stream);
    if (submaps_.size() > 1) {
      bool* more = nullptr;
      auto const pool = cuco::b200::device_scratch_pool();  // cached blocks; the default pool releases at every sync
      CUCO_EXPECTS(pool != nullptr, "no stream-ordered memory pool on this device");
      CUCO_CUDA_TRY(cudaMallocFromPoolAsync(
        reinterpret_cast<void**>(&more), static_cast<std::size_t>(n), pool, stream.get()));
      auto const grid = static_cast<unsigned>(
        std::min<cuco::detail::index_type>(cuco::detail::int_div_ceil(n, cuco::detail::index_type{256}), 1 << 20));
      for (std::size_t i = 1; i < submaps_.size(); ++i) {
        submaps_[i]->contains_async(first, last, more, stream);
        detail::merge_presence<<<grid, 256, 0, stream.get()>>>(cuco::b200::unwrap(output_begin), more, n);
      }
      CUCO_CUDA_TRY(cudaFreeAsync(more, stream.get()));
    }
    stream.wait();
  }

  /// b200 extensions (the reference class keeps these private): keys counted so far, slots of all
  /// submaps, number of submaps.
  [[nodiscard]] size_type size() const noexcept { return size_; }
  [[nodiscard]] size_type capacity() const noexcept
  {
    size_type total = 0;
    for (auto const& m : submaps_) {
      total += m->capacity();
    }
    return total;
  }
  [[nodiscard]] std::size_t num_submaps() const noexcept { return submaps_.size(); }

 private:
  [[nodiscard]] std::int64_t usable(std::size_t submap_capacity) const noexcept
  {
    return static_cast<std::int64_t>(max_load_factor_ * static_cast<float>(submap_capacity));
  }

  void add_submap(cuda::stream_ref stream)
  {
    auto const& head = *submaps_.front();
    submaps_.push_back(std::make_unique<map_type>(static_cast<Extent>(capacity_),
                                                  empty_key<Key>{head.empty_key_sentinel()},
                                                  empty_value<T>{head.empty_value_sentinel()},
                                                  KeyEqual{},
                                                  ProbingScheme{},
                                                  cuda_thread_scope<Scope>{},
                                                  Storage{},
                                                  alloc_,
                                                  stream));
    capacity_ *= 2;
  }

  size_type size_{};            ///< keys counted so far
  size_type capacity_{};        ///< capacity of the next submap to create
  std::vector<std::unique_ptr<map_type>> submaps_;
  size_type min_insert_size_{}; ///< a submap only takes slices of at least this many elements
  float max_load_factor_{};     ///< fill limit of a submap
  Allocator alloc_{};
};

}  // namespace experimental
}  // namespace cuco
