// Exception types of the drop-in surface (reference: include/cuco/utility/error.hpp:29-61).
#pragma once

#include <stdexcept>
#include <string>

namespace cuco {

/// Thrown when a documented precondition of the host API is violated.
struct logic_error : public std::logic_error {
  explicit logic_error(char const* what) : std::logic_error(what) {}
  explicit logic_error(std::string const& what) : std::logic_error(what) {}
};

/// Thrown when a CUDA runtime call made on behalf of the caller fails.
struct cuda_error : public std::runtime_error {
  explicit cuda_error(char const* what) : std::runtime_error(what) {}
  explicit cuda_error(std::string const& what) : std::runtime_error(what) {}
};

}  // namespace cuco
