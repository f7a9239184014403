// One device-resident atomic counter with host-side reset / read-back (reference:
// include/cuco/detail/storage/counter_storage.cuh:37-108). The bulk launchers of this tree take their
// per-call counters from a stream-ordered pool instead (b200/bulk_engine.cuh), but the class is part
// of the surface the reference's own tests and user code include, so it is provided with the same
// interface: construct from an allocator, `reset(stream)`, `data()`, `load_to_host(stream)`.
#pragma once

#include <cuco/aow_storage.cuh>
#include <cuco/detail/error.hpp>
#include <cuco/extent.cuh>

#include <cuda/atomic>
#include <cuda/stream_ref>

#include <memory>

namespace cuco {
namespace detail {

template <typename SizeType, cuda::thread_scope Scope, typename Allocator>
class counter_storage : public storage_base<cuco::extent<SizeType, 1>> {
 public:
  using base_type = storage_base<cuco::extent<SizeType, 1>>;
  using base_type::capacity;

  using size_type      = SizeType;
  using value_type     = cuda::atomic<size_type, Scope>;
  using allocator_type = typename std::allocator_traits<Allocator>::template rebind_alloc<value_type>;
  using counter_deleter_type = custom_deleter<size_type, allocator_type>;

  explicit constexpr counter_storage(Allocator const& allocator)
    : base_type{cuco::extent<size_type, 1>{}},
      allocator_{allocator},
      deleter_{this->capacity(), allocator_},
      counter_{allocator_.allocate(this->capacity()), deleter_}
  {
  }

  /// Zeroes the counter, stream-ordered.
  void reset(cuda::stream_ref stream)
  {
    static_assert(sizeof(size_type) == sizeof(value_type), "the atomic wrapper must not add state");
    CUCO_CUDA_TRY(cudaMemsetAsync(this->data(), 0, sizeof(value_type), stream.get()));
  }

  [[nodiscard]] constexpr value_type* data() noexcept { return counter_.get(); }
  [[nodiscard]] constexpr value_type* data() const noexcept { return counter_.get(); }

  /// Copies the current value to the host; synchronises `stream`.
  [[nodiscard]] size_type load_to_host(cuda::stream_ref stream) const
  {
    size_type value{};
    CUCO_CUDA_TRY(
      cudaMemcpyAsync(&value, this->data(), sizeof(size_type), cudaMemcpyDeviceToHost, stream.get()));
    stream.wait();
    return value;
  }

 private:
  allocator_type allocator_;
  counter_deleter_type deleter_;
  std::unique_ptr<value_type, counter_deleter_type> counter_;
};

}  // namespace detail
}  // namespace cuco
