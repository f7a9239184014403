// cuco::static_multiset_ref — non-owning, trivially-copyable device handle of a static_multiset.
//
// Same template signature, aliases, constructors, accessors, rebind/make_copy/initialize members and
// operators (insert, contains, find, count, retrieve, for_each) as the reference
// (include/cuco/static_multiset_ref.cuh:57-262, detail/static_multiset/static_multiset_ref.inl:
// 25-763). Equal keys may be stored any number of times: insertion never compares keys, lookups that
// enumerate matches walk to the first empty slot. Probing is delegated to cuco::b200::probe_engine
// (AllowsDuplicates = true); operator bodies live in cuco/b200/ref_mixins.cuh.
#pragma once

#include <cuco/b200/probe_engine.cuh>
#include <cuco/b200/ref_mixins.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/operator.hpp>
#include <cuco/probing_scheme.cuh>
#include <cuco/storage.cuh>
#include <cuco/types.cuh>
#include <cuco/utility/cuda_thread_scope.cuh>

#include <cuda/atomic>
#include <cuda/std/type_traits>

#include <utility>

namespace cuco {

template <typename Key,
          cuda::thread_scope Scope,
          typename KeyEqual,
          typename ProbingScheme,
          typename StorageRef,
          typename... Operators>
class static_multiset_ref
  : public detail::operator_impl<
      Operators,
      static_multiset_ref<Key, Scope, KeyEqual, ProbingScheme, StorageRef, Operators...>>... {
  static constexpr auto allows_duplicates = true;

 public:
  /// The probe engine all operations run on (b200-specific, used by the bulk launchers).
  using engine_type =
    b200::probe_engine<Key, Scope, KeyEqual, ProbingScheme, StorageRef, allows_duplicates>;

  using key_type            = Key;
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

  __host__ __device__ explicit constexpr static_multiset_ref(cuco::empty_key<Key> empty_key_sentinel,
                                                        KeyEqual const& predicate,
                                                        ProbingScheme const& probing_scheme,
                                                        cuda_thread_scope<Scope>,
                                                        StorageRef storage_ref) noexcept
    : engine_{static_cast<Key>(empty_key_sentinel), predicate, probing_scheme, storage_ref}
  {
  }

  __host__ __device__ explicit constexpr static_multiset_ref(cuco::empty_key<Key> empty_key_sentinel,
                                                        cuco::erased_key<Key> erased_key_sentinel,
                                                        KeyEqual const& predicate,
                                                        ProbingScheme const& probing_scheme,
                                                        cuda_thread_scope<Scope>,
                                                        StorageRef storage_ref) noexcept
    : engine_{static_cast<Key>(empty_key_sentinel),
              static_cast<Key>(erased_key_sentinel),
              predicate,
              probing_scheme,
              storage_ref}
  {
  }

  /// Re-types a ref with a different operator set (same table).
  template <typename... OtherOperators>
  __host__ __device__ explicit constexpr static_multiset_ref(
    static_multiset_ref<Key, Scope, KeyEqual, ProbingScheme, StorageRef, OtherOperators...>&&
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
    return static_multiset_ref<Key, Scope, KeyEqual, ProbingScheme, StorageRef, NewOperators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
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
    return static_multiset_ref<Key, Scope, NewKeyEqual, ProbingScheme, StorageRef, Operators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
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
    return static_multiset_ref<Key,
                          Scope,
                          KeyEqual,
                          cuda::std::decay_t<decltype(scheme)>,
                          StorageRef,
                          Operators...>{cuco::empty_key<Key>{this->empty_key_sentinel()},
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
    return static_multiset_ref<Key, NewScope, KeyEqual, ProbingScheme, StorageRef, Operators...>{
      cuco::empty_key<Key>{this->empty_key_sentinel()},
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
            cuda::thread_scope Scope_,
            typename KeyEqual_,
            typename ProbingScheme_,
            typename StorageRef_,
            typename... Operators_>
  friend class static_multiset_ref;
};

namespace detail {

#define CUCO_B200_MULTISET_REF static_multiset_ref<Key, Scope, KeyEqual, ProbingScheme, StorageRef, Operators...>
#define CUCO_B200_MULTISET_REF_TEMPLATE    \
  template <typename Key,             \
            cuda::thread_scope Scope, \
            typename KeyEqual,        \
            typename ProbingScheme,   \
            typename StorageRef,      \
            typename... Operators>

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::insert_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_insert<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::contains_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_contains<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::count_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_count<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::find_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_find<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::for_each_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_for_each<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

CUCO_B200_MULTISET_REF_TEMPLATE
class operator_impl<op::retrieve_tag, CUCO_B200_MULTISET_REF>
  : public b200::mixin_retrieve<CUCO_B200_MULTISET_REF, ProbingScheme::cg_size> {};

#undef CUCO_B200_MULTISET_REF
#undef CUCO_B200_MULTISET_REF_TEMPLATE

}  // namespace detail
}  // namespace cuco
