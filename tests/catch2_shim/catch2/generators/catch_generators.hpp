// GENERATE of the Catch2 stand-in (see ../catch_test_macros.hpp): the test body re-runs once per
// combination of the values of all GENERATE calls it executes.
#pragma once
#include "../catch_test_macros.hpp"

#include <type_traits>

#define GENERATE(first, ...) \
  ::catch2_shim::generate<std::decay_t<decltype(first)>>({first, __VA_ARGS__})
