// Array-of-windows slot storage (reference: include/cuco/aow_storage.cuh:49-224,
// detail/storage/aow_storage.inl:34-203, aow_storage_base.cuh:34-101, storage_base.cuh:30-98,
// storage/kernels.cuh:37-52).
//
// Layout in HBM: one contiguous allocation of `num_windows` windows, each `window_size` slots of
// type T (T = Key for sets, cuco::pair<Key,Payload> for maps), i.e. a plain AoS array of slots.
// Device refs hand out pointers into it (`find` returns a pointer to a real pair), so the layout is
// part of the surface. What is new here is how it is driven: the fill kernel streams 128-bit stores
// from a grid sized to the SM count instead of one scalar slot store per thread, and the owning
// class allocates num_windows windows plus the few it takes to end the allocation on a 32-byte
// sector boundary (`allocated_windows()`; the reference over-allocates by a factor of window_size,
// aow_storage.inl:40). The bulk kernels read the table in whole 32-byte sectors, so the sector the
// last slot lives in must be readable in full whatever allocator the user plugged in; the padding
// slots are filled with the empty sentinel and are never part of a probe sequence.
#pragma once

#include <cuco/detail/error.hpp>
#include <cuco/detail/utility/cuda.cuh>
#include <cuco/extent.cuh>
#include <cuco/utility/allocator.hpp>

#include <cuda/std/array>
#include <cuda/stream_ref>

#include <cstdint>
#include <cstring>
#include <iterator>
#include <memory>

namespace cuco {
namespace detail {

/// `window_size` slots handled as one unit by a probing thread.
template <typename T, std::int32_t WindowSize>
struct window : public cuda::std::array<T, WindowSize> {
  static constexpr std::int32_t window_size = WindowSize;
};

/// Frees through the allocator that produced the pointer (unique_ptr deleter).
template <typename SizeType, typename Allocator>
struct custom_deleter {
  using pointer = typename std::allocator_traits<Allocator>::pointer;

  explicit constexpr custom_deleter(SizeType size, Allocator& allocator)
    : size_{size}, allocator_{allocator}
  {
  }

  void operator()(pointer ptr) { allocator_.deallocate(ptr, size_); }

  SizeType size_;
  Allocator& allocator_;
};

/// Extent holder common to every storage flavour.
template <typename Extent>
class storage_base {
 public:
  using extent_type = Extent;
  using size_type   = typename extent_type::value_type;

  __host__ __device__ explicit constexpr storage_base(Extent size) : extent_{size} {}

  [[nodiscard]] __host__ __device__ constexpr size_type capacity() const noexcept
  {
    return static_cast<size_type>(extent_);
  }

  [[nodiscard]] __host__ __device__ constexpr extent_type extent() const noexcept
  {
    return extent_;
  }

 protected:
  extent_type extent_;
};

/// Geometry shared by the owning storage and its ref.
template <typename T, std::int32_t WindowSize, typename Extent>
class aow_storage_base : public storage_base<Extent> {
 public:
  static constexpr std::int32_t window_size = WindowSize;

  using extent_type = typename storage_base<Extent>::extent_type;
  using size_type   = typename storage_base<Extent>::size_type;
  using value_type  = T;
  using window_type = window<value_type, window_size>;

  __host__ __device__ explicit constexpr aow_storage_base(Extent size) : storage_base<Extent>{size}
  {
  }

  [[nodiscard]] __host__ __device__ constexpr size_type num_windows() const noexcept
  {
    return storage_base<Extent>::capacity();
  }

  /// Total number of slots.
  [[nodiscard]] __host__ __device__ constexpr size_type capacity() const noexcept
  {
    return storage_base<Extent>::capacity() * window_size;
  }

  [[nodiscard]] __host__ __device__ constexpr extent_type window_extent() const noexcept
  {
    return storage_base<Extent>::extent();
  }
};

/// Fills `num_slots` slots with `value`. When the slot size divides 16 the bulk of the range is
/// written as 128-bit stores of a replicated pattern; the (at most 15-byte) tail and odd slot sizes
/// fall back to slot-wise stores.
template <typename Slot>
CUCO_KERNEL void fill_slots(Slot* slots, index_type num_slots, Slot value)
{
  auto const tid    = global_thread_id();
  auto const stride = grid_stride();

  if constexpr (16 % sizeof(Slot) == 0 && alignof(Slot) == sizeof(Slot)) {
    constexpr int per_vec = 16 / sizeof(Slot);
    uint4 pattern;
    for (int i = 0; i < per_vec; ++i) {
      memcpy(reinterpret_cast<char*>(&pattern) + i * sizeof(Slot), &value, sizeof(Slot));
    }
    // cudaMalloc'ed storage is 256-byte aligned; refs over user memory may only be slot aligned,
    // so peel until the pointer is 16-byte aligned.
    auto const misaligned = reinterpret_cast<std::uintptr_t>(slots) & 15u;
    index_type head       = misaligned ? (16 - misaligned) / sizeof(Slot) : 0;
    if (head > num_slots) { head = num_slots; }
    index_type const num_vecs = (num_slots - head) / per_vec;
    auto* vecs                = reinterpret_cast<uint4*>(slots + head);
    for (index_type i = tid; i < num_vecs; i += stride) {
      vecs[i] = pattern;
    }
    index_type const done = head + num_vecs * per_vec;
    if (tid < head) { slots[tid] = value; }
    if (tid < num_slots - done) { slots[done + tid] = value; }
  } else {
    for (index_type i = tid; i < num_slots; i += stride) {
      slots[i] = value;
    }
  }
}

}  // namespace detail

template <typename T, std::int32_t WindowSize>
using window = detail::window<T, WindowSize>;

template <typename T, std::int32_t WindowSize, typename Extent>
class aow_storage_ref;

/// Owning storage: allocates through `Allocator` (rebound to the window type), move-only.
template <typename T,
          std::int32_t WindowSize,
          typename Extent    = cuco::extent<std::size_t>,
          typename Allocator = cuco::cuda_allocator<cuco::window<T, WindowSize>>>
class aow_storage : public detail::aow_storage_base<T, WindowSize, Extent> {
 public:
  using base_type = detail::aow_storage_base<T, WindowSize, Extent>;
  using base_type::window_size;

  using extent_type = typename base_type::extent_type;
  using size_type   = typename base_type::size_type;
  using value_type  = typename base_type::value_type;
  using window_type = typename base_type::window_type;

  using base_type::capacity;
  using base_type::num_windows;

  using allocator_type =
    typename std::allocator_traits<Allocator>::template rebind_alloc<window_type>;
  using window_deleter_type = detail::custom_deleter<size_type, allocator_type>;
  using ref_type            = aow_storage_ref<value_type, window_size, extent_type>;

  explicit constexpr aow_storage(Extent size, Allocator const& allocator = {})
    : base_type{size},
      allocator_{allocator},
      window_deleter_{allocated_windows(), allocator_},
      windows_{allocator_.allocate(allocated_windows()), window_deleter_}
  {
  }

  /// Windows actually allocated: `num_windows()` rounded up so that the allocation covers the whole
  /// 32-byte sector holding the last slot (sector-wide table loads never leave the allocation, even
  /// with a user allocator that sub-allocates tightly).
  [[nodiscard]] constexpr size_type allocated_windows() const noexcept
  {
    constexpr std::size_t sector = 32;
    auto const bytes  = static_cast<std::size_t>(num_windows()) * sizeof(window_type);
    auto const padded = (bytes + sector - 1) / sector * sector;
    return static_cast<size_type>((padded + sizeof(window_type) - 1) / sizeof(window_type));
  }

  aow_storage(aow_storage&&)                 = default;
  aow_storage& operator=(aow_storage&&)      = default;
  ~aow_storage()                             = default;
  aow_storage(aow_storage const&)            = delete;
  aow_storage& operator=(aow_storage const&) = delete;

  [[nodiscard]] constexpr window_type* data() const noexcept { return windows_.get(); }
  [[nodiscard]] constexpr allocator_type allocator() const noexcept { return allocator_; }
  [[nodiscard]] constexpr ref_type ref() const noexcept
  {
    return ref_type{this->window_extent(), this->data()};
  }

  /// Sets every slot to `key` and waits for completion.
  void initialize(value_type key, cuda::stream_ref stream = {})
  {
    this->initialize_async(key, stream);
    stream.wait();
  }

  /// Sets every slot to `key`, stream-ordered. Pure store bandwidth.
  void initialize_async(value_type key, cuda::stream_ref stream = {}) noexcept
  {
    // the padding behind the last window is filled too (see allocated_windows())
    auto const num_slots =
      static_cast<detail::index_type>(this->allocated_windows()) * detail::index_type{window_size};
    if (num_slots == 0) { return; }
    constexpr int block    = 256;
    constexpr int vec_elems = (16 % sizeof(value_type) == 0) ? 16 / sizeof(value_type) : 1;
    auto const work_items  = detail::int_div_ceil(num_slots, detail::index_type{vec_elems});
    int sms                = 148;
    int dev                = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    auto const full_grid = detail::int_div_ceil(work_items, detail::index_type{block});
    // Four grid-strided 16-byte stores per thread. Measured on B200 for a 1.6 GB table
    // (profiles/r01_next_rows.jsonl): a persistent grid of (SM count x 8) CTAs 5.7 TB/s, one store
    // per thread 4.2 TB/s; four stores per thread is the shape that reaches 7 TB/s.
    (void)sms;
    auto const grid = static_cast<unsigned>(std::max<detail::index_type>(
      1, std::min<detail::index_type>(detail::int_div_ceil(full_grid, detail::index_type{4}),
                                      detail::index_type{0x7fffffff})));
    detail::fill_slots<value_type><<<grid, block, 0, stream.get()>>>(
      reinterpret_cast<value_type*>(this->data()), num_slots, key);
  }

 private:
  allocator_type allocator_;
  window_deleter_type window_deleter_;
  std::unique_ptr<window_type, window_deleter_type> windows_;
};

/// Non-owning view handed to device code; trivially copyable.
template <typename T, std::int32_t WindowSize, typename Extent = cuco::extent<std::size_t>>
class aow_storage_ref : public detail::aow_storage_base<T, WindowSize, Extent> {
 public:
  using base_type = detail::aow_storage_base<T, WindowSize, Extent>;
  using base_type::window_size;

  using extent_type = typename base_type::extent_type;
  using size_type   = typename base_type::size_type;
  using value_type  = typename base_type::value_type;
  using window_type = typename base_type::window_type;

  using base_type::capacity;
  using base_type::num_windows;

  __host__ __device__ explicit constexpr aow_storage_ref(Extent size,
                                                         window_type* windows) noexcept
    : base_type{size}, windows_{windows}
  {
  }

  /// Pointer-like handle to one slot. Not incrementable: slots are reached by probing, not walking.
  struct iterator {
    using iterator_category = std::input_iterator_tag;
    using value_type        = T;
    using difference_type   = std::ptrdiff_t;
    using pointer           = T*;
    using reference         = T&;

    __device__ constexpr explicit iterator(T* slot) noexcept : slot_{slot} {}

    __device__ constexpr reference operator*() const { return *slot_; }
    __device__ constexpr pointer operator->() const { return slot_; }

    friend __device__ constexpr bool operator==(iterator const& a, iterator const& b) noexcept
    {
      return a.slot_ == b.slot_;
    }
    friend __device__ constexpr bool operator!=(iterator const& a, iterator const& b) noexcept
    {
      return a.slot_ != b.slot_;
    }

   private:
    T* slot_{};
  };
  using const_iterator = iterator const;

  /// One past the last slot; what `find` returns on a miss.
  [[nodiscard]] __device__ constexpr iterator end() noexcept
  {
    return iterator{reinterpret_cast<value_type*>(windows_) + this->capacity()};
  }
  [[nodiscard]] __device__ constexpr const_iterator end() const noexcept
  {
    return const_iterator{reinterpret_cast<value_type*>(windows_) + this->capacity()};
  }

  [[nodiscard]] __host__ __device__ constexpr window_type* data() noexcept { return windows_; }
  [[nodiscard]] __host__ __device__ constexpr window_type* data() const noexcept
  {
    return windows_;
  }

  /// Loads window `index` by value (one aligned vector load when the window is 4/8/16/32 bytes).
  [[nodiscard]] __device__ constexpr window_type operator[](size_type index) const noexcept
  {
    return *reinterpret_cast<window_type*>(
      __builtin_assume_aligned(windows_ + index, sizeof(value_type) * window_size));
  }

 private:
  window_type* windows_;
};

}  // namespace cuco
