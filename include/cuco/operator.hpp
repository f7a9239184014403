// Operator tags that select which device-side operations a container ref exposes
// (reference: include/cuco/operator.hpp:20-83, detail/operator.inl:36-54). A ref type lists tags
// as template arguments and inherits one mixin per tag; `ref(cuco::insert, cuco::find)` passes the
// inline constexpr tag objects below.
#pragma once

#include <type_traits>

namespace cuco {
inline namespace op {

struct insert_tag {};
struct insert_and_find_tag {};
struct insert_or_assign_tag {};
struct insert_or_apply_tag {};
struct erase_tag {};
struct contains_tag {};
struct count_tag {};
struct find_tag {};
struct retrieve_tag {};
struct for_each_tag {};

inline constexpr insert_tag insert{};
inline constexpr insert_and_find_tag insert_and_find{};
inline constexpr insert_or_assign_tag insert_or_assign{};
inline constexpr insert_or_apply_tag insert_or_apply{};
inline constexpr erase_tag erase{};
inline constexpr contains_tag contains{};
inline constexpr count_tag count{};
inline constexpr find_tag find{};
inline constexpr retrieve_tag retrieve{};
inline constexpr for_each_tag for_each{};

}  // namespace op

namespace detail {

/// Primary template of the per-operator mixin; each container ref header specialises it.
template <typename OperatorTag, typename Ref>
class operator_impl;

/// True when `Tag` is among `Tags...`.
template <typename Tag, typename... Tags>
inline constexpr bool has_operator_v = (std::is_same_v<Tag, Tags> || ...);

/// Same query in the function form the reference exposes: has_operator<op::find_tag, Ops...>().
template <typename Tag, typename... Tags>
constexpr bool has_operator() noexcept
{
  return has_operator_v<Tag, Tags...>;
}

}  // namespace detail
}  // namespace cuco
