// TEMPLATE_TEST_CASE / TEMPLATE_TEST_CASE_SIG of the Catch2 stand-in (see catch_test_macros.hpp).
//   TEMPLATE_TEST_CASE_SIG(name, tags, ((template parameter list), names...), (arguments)...)
// declares a function template with that parameter list and registers one instantiation per
// argument tuple, named "<name> - <arguments>" like Catch2 does.
#pragma once
#include "catch_test_macros.hpp"

#define CATCH2_SHIM_UNPAREN(...) __VA_ARGS__
#define CATCH2_SHIM_SIG_OF(first, ...) CATCH2_SHIM_UNPAREN first
#define CATCH2_SHIM_EXPAND(x) x

#define CATCH2_SHIM_NTH(_1, _2, _3, _4, _5, _6, _7, _8, _9, _10, _11, _12, _13, _14, _15, _16, _17, _18, _19, _20, _21, _22, _23, _24, _25, _26, _27, _28, _29, _30, _31, _32, N, ...) N
#define CATCH2_SHIM_COUNT(...) CATCH2_SHIM_EXPAND(CATCH2_SHIM_NTH(__VA_ARGS__, 32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1))
#define CATCH2_SHIM_FE_1(m, a, b, x) m(a, b, x)
#define CATCH2_SHIM_FE_2(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_1(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_3(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_2(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_4(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_3(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_5(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_4(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_6(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_5(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_7(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_6(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_8(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_7(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_9(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_8(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_10(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_9(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_11(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_10(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_12(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_11(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_13(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_12(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_14(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_13(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_15(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_14(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_16(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_15(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_17(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_16(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_18(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_17(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_19(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_18(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_20(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_19(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_21(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_20(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_22(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_21(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_23(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_22(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_24(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_23(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_25(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_24(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_26(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_25(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_27(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_26(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_28(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_27(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_29(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_28(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_30(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_29(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_31(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_30(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FE_32(m, a, b, x, ...) m(a, b, x) CATCH2_SHIM_EXPAND(CATCH2_SHIM_FE_31(m, a, b, __VA_ARGS__))
#define CATCH2_SHIM_FOR_EACH(m, a, b, ...) \
  CATCH2_SHIM_EXPAND(CATCH2_SHIM_CAT(CATCH2_SHIM_FE_, CATCH2_SHIM_COUNT(__VA_ARGS__))(m, a, b, __VA_ARGS__))

#define CATCH2_SHIM_REGISTER_SIG(fn, name, tuple) \
  ::catch2_shim::registry().push_back({std::string(name) + " - " #tuple, &fn<CATCH2_SHIM_UNPAREN tuple>});
#define CATCH2_SHIM_REGISTER_TYPE(fn, name, type) \
  ::catch2_shim::registry().push_back({std::string(name) + " - " #type, &fn<type>});

#define CATCH2_SHIM_TTC_SIG(fn, name, sig, ...)                                              \
  template <CATCH2_SHIM_SIG_OF sig>                                                          \
  static void fn();                                                                          \
  namespace {                                                                                \
  struct CATCH2_SHIM_CAT(fn, _registrar) {                                                   \
    CATCH2_SHIM_CAT(fn, _registrar)()                                                        \
    {                                                                                        \
      CATCH2_SHIM_FOR_EACH(CATCH2_SHIM_REGISTER_SIG, fn, name, __VA_ARGS__)                  \
    }                                                                                        \
  } CATCH2_SHIM_CAT(fn, _registrar_instance);                                                \
  }                                                                                          \
  template <CATCH2_SHIM_SIG_OF sig>                                                          \
  static void fn()

#define TEMPLATE_TEST_CASE_SIG(name, tags, sig, ...) \
  CATCH2_SHIM_TTC_SIG(CATCH2_SHIM_UNIQUE(catch2_shim_template_test_), name, sig, __VA_ARGS__)

#define CATCH2_SHIM_TTC(fn, name, ...)                                                       \
  template <typename TestType>                                                               \
  static void fn();                                                                          \
  namespace {                                                                                \
  struct CATCH2_SHIM_CAT(fn, _registrar) {                                                   \
    CATCH2_SHIM_CAT(fn, _registrar)()                                                        \
    {                                                                                        \
      CATCH2_SHIM_FOR_EACH(CATCH2_SHIM_REGISTER_TYPE, fn, name, __VA_ARGS__)                 \
    }                                                                                        \
  } CATCH2_SHIM_CAT(fn, _registrar_instance);                                                \
  }                                                                                          \
  template <typename TestType>                                                               \
  static void fn()

#define TEMPLATE_TEST_CASE(name, tags, ...) \
  CATCH2_SHIM_TTC(CATCH2_SHIM_UNIQUE(catch2_shim_template_test_), name, __VA_ARGS__)
