// Aggregation functors for static_map::insert_or_apply
// (reference: include/cuco/utility/reduction_functors.cuh:25-82). Each is invoked as
// `op(cuda::atomic_ref<T, Scope>{slot.second}, value)` and must be commutative + associative for the
// result to be independent of arrival order; all use relaxed ordering like the rest of the path.
#pragma once

#include <cuda/atomic>

namespace cuco::reduce {

/// payload += value
struct plus {
  template <typename T, cuda::thread_scope Scope>
  __device__ void operator()(cuda::atomic_ref<T, Scope> payload, T const& value) const
  {
    payload.fetch_add(value, cuda::memory_order_relaxed);
  }
};

/// payload = max(payload, value)
struct max {
  template <typename T, cuda::thread_scope Scope>
  __device__ void operator()(cuda::atomic_ref<T, Scope> payload, T const& value) const
  {
    payload.fetch_max(value, cuda::memory_order_relaxed);
  }
};

/// payload = min(payload, value)
struct min {
  template <typename T, cuda::thread_scope Scope>
  __device__ void operator()(cuda::atomic_ref<T, Scope> payload, T const& value) const
  {
    payload.fetch_min(value, cuda::memory_order_relaxed);
  }
};

}  // namespace cuco::reduce
