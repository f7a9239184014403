// CUCO_CUDA_TRY / CUCO_EXPECTS / CUCO_FAIL / CUCO_ASSERT_CUDA_SUCCESS with the same spelling,
// arities and default exception types as the reference macros (include/cuco/detail/error.hpp:47-142)
// so callers and tests written against them compile unchanged. The work is done by two small
// inline functions instead of macro bodies.
#pragma once

#include <cuco/utility/error.hpp>

#include <cuda_runtime_api.h>

#include <cassert>
#include <string>
#include <type_traits>

namespace cuco::detail {

template <typename Exception>
[[noreturn]] inline void raise(char const* kind, char const* file, int line, std::string const& msg)
{
  static_assert(std::is_base_of_v<std::exception, Exception>);
  throw Exception{std::string{kind} + file + ":" + std::to_string(line) + ": " + msg};
}

template <typename Exception>
inline void check_cuda(cudaError_t status, char const* file, int line)
{
  if (status == cudaSuccess) { return; }
  cudaGetLastError();  // clear the sticky-less error so the next call starts clean
  raise<Exception>("CUDA error at: ",
                   file,
                   line,
                   std::string{cudaGetErrorName(status)} + " " + cudaGetErrorString(status));
}

}  // namespace cuco::detail

#define CUCO_STRINGIFY_DETAIL(x) #x
#define CUCO_STRINGIFY(x)        CUCO_STRINGIFY_DETAIL(x)

#define CUCO_B200_PICK3(_1, _2, _3, NAME, ...) NAME
#define CUCO_B200_PICK2(_1, _2, NAME, ...)     NAME

#define CUCO_CUDA_TRY_2(call, exception_type) \
  ::cuco::detail::check_cuda<exception_type>((call), __FILE__, __LINE__)
#define CUCO_CUDA_TRY_1(call) CUCO_CUDA_TRY_2(call, ::cuco::cuda_error)
#define CUCO_CUDA_TRY(...) \
  CUCO_B200_PICK2(__VA_ARGS__, CUCO_CUDA_TRY_2, CUCO_CUDA_TRY_1)(__VA_ARGS__)

#define CUCO_EXPECTS_3(cond, reason, exception_type)                                        \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::cuco::detail::raise<exception_type>("CUCO failure at: ", __FILE__, __LINE__, reason); \
    }                                                                                       \
  } while (0)
#define CUCO_EXPECTS_2(cond, reason) CUCO_EXPECTS_3(cond, reason, ::cuco::logic_error)
#define CUCO_EXPECTS(...) \
  CUCO_B200_PICK3(__VA_ARGS__, CUCO_EXPECTS_3, CUCO_EXPECTS_2)(__VA_ARGS__)

#define CUCO_FAIL_2(what, exception_type) \
  ::cuco::detail::raise<exception_type>("CUCO failure at: ", __FILE__, __LINE__, what)
#define CUCO_FAIL_1(what) CUCO_FAIL_2(what, ::cuco::logic_error)
#define CUCO_FAIL(...)    CUCO_B200_PICK2(__VA_ARGS__, CUCO_FAIL_2, CUCO_FAIL_1)(__VA_ARGS__)

#define CUCO_ASSERT_CUDA_SUCCESS(expr)       \
  do {                                       \
    cudaError_t const cuco_status_ = (expr); \
    assert(cudaSuccess == cuco_status_);     \
    (void)cuco_status_;                      \
  } while (0)
