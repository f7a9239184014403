// Strong type carrying a cuda::thread_scope as a constructor argument so CTAD can pick it up
// (reference: include/cuco/utility/cuda_thread_scope.cuh:29-46).
#pragma once

#include <cuda/atomic>

namespace cuco {

template <cuda::thread_scope Scope>
struct cuda_thread_scope {
  /// Lets `cuda_thread_scope<S>{}` decay to the enum where a plain scope value is wanted.
  __host__ __device__ constexpr operator cuda::thread_scope() const noexcept { return Scope; }
};

inline constexpr cuda_thread_scope<cuda::thread_scope_system> thread_scope_system{};
inline constexpr cuda_thread_scope<cuda::thread_scope_device> thread_scope_device{};
inline constexpr cuda_thread_scope<cuda::thread_scope_block> thread_scope_block{};
inline constexpr cuda_thread_scope<cuda::thread_scope_thread> thread_scope_thread{};

}  // namespace cuco
