// cuco::static_multimap_ref — non-owning, trivially-copyable device handle of an
// experimental::static_multimap.
//
// Same template signature, aliases, constructors, accessors and rebind/make_copy/initialize members
// as the reference (include/cuco/static_multimap_ref.cuh:52-320), operators insert, contains, count
// and for_each (detail/static_multimap/static_multimap_ref.inl:405-690). Probing is delegated to
// cuco::b200::probe_engine with AllowsDuplicates = true.
#pragma once

#include <cuco/b200/probe_engine.cuh>
#include <cuco/b200/ref_mixins.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/operator.hpp>
#include <cuco/pair.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/storage.cuh>
#include <cuco/types.cuh>
#include <cuco/utility/cuda_thread_scope.cuh>

#include <cuda/atomic>
#include <cuda/std/type_traits>

#include <utility>

namespace cuco {

template <typename Key,
          typename T,
          cuda::thread_scope Scope,
          typename KeyEqual,
          typename ProbingScheme,
          typename StorageRef,
          typename... Operators>
class static_multimap_ref
  : public detail::operator_impl<
      Operators,
      static_multimap_ref<Key, T, Scope, KeyEqual, ProbingScheme, StorageRef, Operators...>>... {
  static constexpr auto allows_duplicates = true;

  static_assert(sizeof(T) == 4 or sizeof(T) == 8,
                "sizeof(mapped_type) must be either 4 bytes or 8 bytes.");
  static_assert(
    cuco::is_bitwise_comparable_v<Key>,
    "Key type must have unique object representations or have been explicitly declared as safe for "
    "bitwise comparison via specialization of cuco::is_bitwise_comparable_v<Key>.");

 public:
  /// The probe engine all operations run on (b200-specific, used by the bulk launchers).
  using engine_type =
    b200::probe_engine<Key, Scope, KeyEqual, ProbingScheme, StorageRef, allows_duplicates>;

  using key_type            = Key;
  using mapped_type         = T;
  using probing_scheme_type = ProbingScheme;
  using hasher              = typename probing_scheme_type::hasher;
  using storage_ref_type    = StorageRef;
  using window_type         = typename storage_ref_type::window_type;
  using value_type          = typename storage_ref_type::value_type;
  using extent_type         = typename storage_ref_type::extent_type;
  using size_type           = typename storage_ref_type::size_type;
  using key_equal           = KeyEqual;
  using iterator            = typename storage_ref_type::iterator;
  using const_iterator      = typename storage_ref_type::const_iterator;

  static constexpr auto cg_size      = probing_scheme_type::cg_size;
  static constexpr auto window_size  = storage_ref_type::window_size;
  static constexpr auto thread_scope = engine_type::thread_scope;

  __host__ __device__ explicit constexpr static_multimap_ref(cuco::empty_key<Key> empty_key_sentinel,
                                                        cuco::empty_value<T> empty_value_sentinel,
                                                        KeyEqual const& predicate,
                                                        ProbingScheme const& probing_scheme,
                                                        cuda_thread_scope<Scope>,
                                                        StorageRef storage_ref) noexcept
    : engine_{cuco::pair<Key, T>{empty_key_sentinel, empty_value_sentinel},
              predicate,
              probing_scheme,
              storage_ref}
  {
  }

  __host__ __device__ explicit constexpr static_multimap_ref(cuco::empty_key<Key> empty_key_sentinel,
                                                        cuco::empty_value<T> empty_value_sentinel,
                                                        cuco::erased_key<Key> erased_key_sentinel,
                                                        KeyEqual const& predicate,
                                                        ProbingScheme const& probing_scheme,
                                                        cuda_thread_scope<Scope>,
                                                        StorageRef storage_ref) noexcept
    : engine_{cuco::pair<Key, T>{empty_key_sentinel, empty_value_sentinel},
              erased_key_sentinel,
              predicate,
              probing_scheme,
              storage_ref}
  {
  }

  /// Re-types a ref with a different operator set (same table).
  template <typename... OtherOperators>
  __host__ __device__ explicit constexpr static_multimap_ref(
    static_multimap_ref<Key, T, Scope, KeyEqual, ProbingScheme, StorageRef, OtherOperators...>&&
      other) noexcept
    : engine_{std::move(other.engine_)}
  {
  }

  [[nodiscard]] __host__ __device__ constexpr auto capacity() const noexcept
  {
    return engine_.capacity();
  }
  [[nodiscard]] __host__ __device__ constexpr extent_type window_extent() const noexcept
  {
    return engine_.window_extent();
  }
  [[nodiscard]] __host__ __device__ constexpr key_type empty_key_sentinel() const noexcept
  {
    return engine_.empty_key_sentinel();
  }
  [[nodiscard]] __host__ __device__ constexpr mapped_type empty_value_sentinel() const noexcept
  {
    return engine_.empty_value_sentinel();
  }
  [[nodiscard]] __host__ __device__ constexpr key_type erased_key_sentinel() const noexcept
  {
    return engine_.erased_key_sentinel();
  }
  [[nodiscard]] __host__ __device__ constexpr key_equal key_eq() const noexcept
  {
    return engine_.key_eq();
  }
  [[nodiscard]] __host__ __device__ constexpr hasher hash_function() const noexcept
  {
    return engine_.hash_function();
  }
  [[nodiscard]] __device__ constexpr const_iterator end() const noexcept { return engine_.end(); }
  [[nodiscard]] __device__ constexpr iterator end() noexcept { return engine_.end(); }
  [[nodiscard]] __host__ __device__ constexpr auto storage_ref() const noexcept
  {
    return engine_.storage_ref();
  }
  [[nodiscard]] __host__ __device__ constexpr auto probing_scheme() const noexcept
  {
    return engine_.probing_scheme();
  }

  /// Same table, different operators.
  template <typename... NewOperators>
  [[nodiscard]] __host__ __device__ constexpr auto rebind_operators(
    NewOperators...) const noexcept
  {
    return static_multimap_ref<Key, T, Scope, KeyEqual, ProbingScheme, StorageRef, NewOperators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
      cuco::empty_value<T>{this->empty_value_sentinel()},
      cuco::erased_key<Key>{this->erased_key_sentinel()},
      this->key_eq(),
      this->probing_scheme(),
      {},
      this->storage_ref()};
  }

  /// Same table, different key predicate.
  template <typename NewKeyEqual>
  [[nodiscard]] __host__ __device__ constexpr auto rebind_key_eq(
    NewKeyEqual const& key_equal) const noexcept
  {
    return static_multimap_ref<Key, T, Scope, NewKeyEqual, ProbingScheme, StorageRef, Operators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
      cuco::empty_value<T>{this->empty_value_sentinel()},
      cuco::erased_key<Key>{this->erased_key_sentinel()},
      key_equal,
      this->probing_scheme(),
      {},
      this->storage_ref()};
  }

  /// Same table, different hash function(s).
  template <typename NewHash>
  [[nodiscard]] __host__ __device__ constexpr auto rebind_hash_function(NewHash const& hash) const
  {
    auto const scheme = this->probing_scheme().rebind_hash_function(hash);
    return static_multimap_ref<Key,
                          T,
                          Scope,
                          KeyEqual,
                          cuda::std::decay_t<decltype(scheme)>,
                          StorageRef,
                          Operators...>{cuco::empty_key<Key>{this->empty_key_sentinel()},
                                        cuco::empty_value<T>{this->empty_value_sentinel()},
                                        cuco::erased_key<Key>{this->erased_key_sentinel()},
                                        this->key_eq(),
                                        scheme,
                                        {},
                                        this->storage_ref()};
  }

  /// Copies the table into `memory_to_use` (e.g. shared memory) with the whole group and returns a
  /// ref over the copy, operating at `scope`.
  template <typename CG, cuda::thread_scope NewScope = thread_scope>
  [[nodiscard]] __device__ constexpr auto make_copy(
    CG const& tile,
    window_type* const memory_to_use,
    cuda_thread_scope<NewScope> scope = {}) const noexcept
  {
    engine_.make_copy(tile, memory_to_use);
    return static_multimap_ref<Key, T, NewScope, KeyEqual, ProbingScheme, StorageRef, Operators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
      cuco::empty_value<T>{this->empty_value_sentinel()},
      cuco::erased_key<Key>{this->erased_key_sentinel()},
      this->key_eq(),
      this->probing_scheme(),
      scope,
      storage_ref_type{this->window_extent(), memory_to_use}};
  }

  /// Fills the storage with the empty sentinel using the whole group (synchronises it).
  template <typename CG>
  __device__ constexpr void initialize(CG const& tile) noexcept
  {
    engine_.initialize(tile);
  }

  /// b200-specific: the probe engine behind this ref.
  [[nodiscard]] __host__ __device__ constexpr engine_type& engine() noexcept { return engine_; }
  [[nodiscard]] __host__ __device__ constexpr engine_type const& engine() const noexcept
  {
    return engine_;
  }

 private:
  engine_type engine_;

  template <typename Key_,
            typename T_,
            cuda::thread_scope Scope_,
            typename KeyEqual_,
            typename ProbingScheme_,
            typename StorageRef_,
            typename... Operators_>
  friend class static_multimap_ref;
};

namespace detail {

#define CUCO_B200_MULTIMAP_REF \
  static_multimap_ref<Key, T, Scope, KeyEqual, ProbingScheme, StorageRef, Operators...>
#define CUCO_B200_MULTIMAP_REF_TEMPLATE        \
  template <typename Key,                 \
            typename T,                   \
            cuda::thread_scope Scope,     \
            typename KeyEqual,            \
            typename ProbingScheme,       \
            typename StorageRef,          \
            typename... Operators>

CUCO_B200_MULTIMAP_REF_TEMPLATE
class operator_impl<op::insert_tag, CUCO_B200_MULTIMAP_REF>
  : public b200::mixin_insert<CUCO_B200_MULTIMAP_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTIMAP_REF_TEMPLATE
class operator_impl<op::contains_tag, CUCO_B200_MULTIMAP_REF>
  : public b200::mixin_contains<CUCO_B200_MULTIMAP_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTIMAP_REF_TEMPLATE
class operator_impl<op::count_tag, CUCO_B200_MULTIMAP_REF>
  : public b200::mixin_count<CUCO_B200_MULTIMAP_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTIMAP_REF_TEMPLATE
class operator_impl<op::for_each_tag, CUCO_B200_MULTIMAP_REF>
  : public b200::mixin_for_each<CUCO_B200_MULTIMAP_REF, ProbingScheme::cg_size> {};

#undef CUCO_B200_MULTIMAP_REF
#undef CUCO_B200_MULTIMAP_REF_TEMPLATE

}  // namespace detail
}  // namespace cuco
