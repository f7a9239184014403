// `cuco::storage<N>`: the public knob selecting array-of-windows slot storage with N slots per
// window (reference: include/cuco/storage.cuh:24-46). A window is the unit a probing thread loads
// at once; with 16-byte slots `storage<2>` makes a window exactly one 32-byte DRAM sector, which the
// sm_100a path fetches with a single 256-bit load.
#pragma once

#include <cuco/aow_storage.cuh>

#include <cstdint>

namespace cuco {

template <std::int32_t WindowSize>
class storage {
 public:
  static constexpr std::int32_t window_size = WindowSize;

  /// Concrete owning storage for slot type T.
  template <typename T, typename Extent, typename Allocator>
  using impl = aow_storage<T, window_size, Extent, Allocator>;
};

namespace detail {

/// Resolves the user-facing `Storage` tag to its implementation and re-exports its interface.
template <typename Storage, typename T, typename Extent, typename Allocator>
class storage : public Storage::template impl<T, Extent, Allocator> {
 public:
  using impl_type      = typename Storage::template impl<T, Extent, Allocator>;
  using ref_type       = typename impl_type::ref_type;
  using value_type     = typename impl_type::value_type;
  using allocator_type = typename impl_type::allocator_type;

  static constexpr int window_size = impl_type::window_size;

  using impl_type::allocator;
  using impl_type::capacity;
  using impl_type::data;
  using impl_type::initialize;
  using impl_type::initialize_async;
  using impl_type::num_windows;
  using impl_type::ref;
  using impl_type::window_extent;

  explicit constexpr storage(Extent size, Allocator const& allocator) : impl_type{size, allocator}
  {
  }
};

}  // namespace detail
}  // namespace cuco
