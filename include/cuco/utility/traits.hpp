// `is_bitwise_comparable` customisation point (reference: include/cuco/utility/traits.hpp:39-69).
// Slots are matched and claimed by comparing raw bits (CAS), so a key/payload type must either have
// unique object representations or be opted in by the user.
#pragma once

#include <type_traits>

namespace cuco {

template <typename T, typename = void>
struct is_bitwise_comparable : std::false_type {};

template <typename T>
struct is_bitwise_comparable<T, std::enable_if_t<std::has_unique_object_representations_v<T>>>
  : std::true_type {};

template <typename T>
inline constexpr bool is_bitwise_comparable_v = is_bitwise_comparable<T>::value;

/// A bool that depends on template arguments, for `static_assert(dependent_false<T>)` branches.
template <bool Value, typename... Ts>
inline constexpr bool dependent_bool_value = Value;
template <typename... Ts>
inline constexpr bool dependent_false = dependent_bool_value<false, Ts...>;

}  // namespace cuco

/// Opt a type with padding/non-unique representation in to bitwise comparison.
#define CUCO_DECLARE_BITWISE_COMPARABLE(Type)           \
  namespace cuco {                                      \
  template <>                                           \
  struct is_bitwise_comparable<Type> : std::true_type { \
  };                                                    \
  }
