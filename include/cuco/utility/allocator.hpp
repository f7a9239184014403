// Default device allocator of the drop-in surface (reference: include/cuco/utility/allocator.hpp:28-92):
// stateless, rebindable, cudaMalloc/cudaFree. Containers only require allocate/deallocate/value_type
// plus rebinding through std::allocator_traits, so user allocators (RMM-style) plug in unchanged.
#pragma once

#include <cuco/detail/error.hpp>

#include <cuda_runtime_api.h>

#include <cstddef>

namespace cuco {

template <typename T>
class cuda_allocator {
 public:
  using value_type = T;

  cuda_allocator() = default;

  template <typename U>
  cuda_allocator(cuda_allocator<U> const&) noexcept
  {
  }

  /// Allocates room for `n` objects of T in device memory (throws cuco::cuda_error on failure).
  value_type* allocate(std::size_t n)
  {
    void* raw = nullptr;
    CUCO_CUDA_TRY(cudaMalloc(&raw, n * sizeof(value_type)));
    return static_cast<value_type*>(raw);
  }

  void deallocate(value_type* p, std::size_t /*n*/) { CUCO_CUDA_TRY(cudaFree(p)); }
};

// All cuda_allocators are interchangeable.
template <typename T, typename U>
bool operator==(cuda_allocator<T> const&, cuda_allocator<U> const&) noexcept
{
  return true;
}

template <typename T, typename U>
bool operator!=(cuda_allocator<T> const&, cuda_allocator<U> const&) noexcept
{
  return false;
}

}  // namespace cuco
