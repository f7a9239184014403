// One explicit instantiation of the cuco::static_map / static_set / static_multiset surface behind
// the C ABI.
// Compile with -DCUCO_SHIM_KIND=<k>; the include root on the command line decides whether this is
// the b200-native table (-I include) or cuco's own build (-I /root/reference/include). Only the
// public container API is used, so both compile from this one source.
#include "cabi_table.hpp"

#if defined(CUCO_SHIM_STUB)
// Development builds (cucollections_b200/build.py, CUCO_B200_DEV_KINDS) compile only the kinds under
// work; the others get this stub so that the library still links and says why a kind is missing.
#include <stdexcept>
#define CUCO_SHIM_CAT2(a, b) a##b
#define CUCO_SHIM_CAT(a, b)  CUCO_SHIM_CAT2(a, b)
cuco_b200_table* CUCO_SHIM_CAT(cuco_shim_make_kind_, CUCO_SHIM_KIND)(
  std::int64_t, double, std::int64_t, std::int64_t, int, std::int64_t, void*)
{
  throw std::invalid_argument("this kind is not part of the development build of the library");
}
#else

#include <cuco/static_map.cuh>
#include <cuco/static_multimap.cuh>
#include <cuco/static_multiset.cuh>
#include <cuco/static_set.cuh>
#include <cuco/utility/reduction_functors.cuh>

#include <cuda/std/tuple>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <cstdint>
#include <stdexcept>

#ifndef CUCO_SHIM_KIND
#error "define CUCO_SHIM_KIND"
#endif

namespace {

using i32 = std::int32_t;
using i64 = std::int64_t;

template <typename K>
using eq = thrust::equal_to<K>;
template <typename K, typename V, typename Probe, int W>
using map_t = cuco::static_map<K,
                               V,
                               cuco::extent<std::size_t>,
                               cuda::thread_scope_device,
                               eq<K>,
                               Probe,
                               cuco::cuda_allocator<cuco::pair<K, V>>,
                               cuco::storage<W>>;
template <typename K, typename Probe, int W>
using set_t = cuco::static_set<K,
                               cuco::extent<std::size_t>,
                               cuda::thread_scope_device,
                               eq<K>,
                               Probe,
                               cuco::cuda_allocator<K>,
                               cuco::storage<W>>;
template <typename K, typename Probe, int W>
using multiset_t = cuco::static_multiset<K,
                                         cuco::extent<std::size_t>,
                                         cuda::thread_scope_device,
                                         eq<K>,
                                         Probe,
                                         cuco::cuda_allocator<K>,
                                         cuco::storage<W>>;
template <typename K, typename V, typename Probe, int W>
using multimap_t = cuco::experimental::static_multimap<K,
                                                       V,
                                                       cuco::extent<std::size_t>,
                                                       cuda::thread_scope_device,
                                                       eq<K>,
                                                       Probe,
                                                       cuco::cuda_allocator<cuco::pair<K, V>>,
                                                       cuco::storage<W>>;
template <typename C>
struct is_multimap : std::false_type {};
template <typename K, typename V, typename E, cuda::thread_scope S, typename Q, typename P, typename A, typename St>
struct is_multimap<cuco::experimental::static_multimap<K, V, E, S, Q, P, A, St>> : std::true_type {};
template <typename C>
struct is_multiset : std::false_type {};
template <typename K, typename E, cuda::thread_scope S, typename Q, typename P, typename A, typename St>
struct is_multiset<cuco::static_multiset<K, E, S, Q, P, A, St>> : std::true_type {};

#if CUCO_SHIM_KIND == 0
using container_t = set_t<i32, cuco::double_hashing<4, cuco::default_hash_function<i32>>, 1>;
#elif CUCO_SHIM_KIND == 1
using container_t = map_t<i64, i64, cuco::linear_probing<1, cuco::default_hash_function<i64>>, 1>;
#elif CUCO_SHIM_KIND == 2
using container_t = map_t<i64, i64, cuco::double_hashing<8, cuco::default_hash_function<i64>>, 1>;
#elif CUCO_SHIM_KIND == 3
using container_t = map_t<i32, i32, cuco::linear_probing<4, cuco::default_hash_function<i32>>, 1>;
#elif CUCO_SHIM_KIND == 4
using container_t = map_t<i64, i64, cuco::linear_probing<4, cuco::default_hash_function<i64>>, 1>;
#elif CUCO_SHIM_KIND == 5
using container_t = set_t<i64, cuco::double_hashing<4, cuco::default_hash_function<i64>>, 1>;
#elif CUCO_SHIM_KIND == 6
using container_t = map_t<i64, i64, cuco::linear_probing<1, cuco::default_hash_function<i64>>, 2>;
#elif CUCO_SHIM_KIND == 7
using container_t = map_t<i32, i32, cuco::double_hashing<2, cuco::murmurhash3_32<i32>>, 2>;
#elif CUCO_SHIM_KIND == 8
using container_t = map_t<i32, i64, cuco::linear_probing<1, cuco::default_hash_function<i32>>, 1>;
#elif CUCO_SHIM_KIND == 9
using container_t = map_t<i64, i64, cuco::double_hashing<8, cuco::xxhash_64<i64>>, 1>;
#elif CUCO_SHIM_KIND == 10
using container_t = multiset_t<i32, cuco::double_hashing<4, cuco::default_hash_function<i32>>, 2>;
#elif CUCO_SHIM_KIND == 11
using container_t = multiset_t<i64, cuco::linear_probing<1, cuco::default_hash_function<i64>>, 2>;
#elif CUCO_SHIM_KIND == 12
using container_t = multimap_t<i64, i64, cuco::linear_probing<4, cuco::default_hash_function<i64>>, 1>;
#elif CUCO_SHIM_KIND == 13
using container_t = map_t<i64, i64, cuco::linear_probing<1, cuco::xxhash_64<i64>>, 1>;
#else
#error "unknown CUCO_SHIM_KIND"
#endif
using selected_container = container_t;

template <typename C, typename = void>
struct mapped_of {
  using type = void;
};
template <typename C>
struct mapped_of<C, std::void_t<typename C::mapped_type>> {
  using type = typename C::mapped_type;
};

/// (keys[i], values[i]) -> cuco::pair, for the separate-arrays input layout.
template <typename K, typename V>
struct zip_to_pair {
  K const* keys;
  V const* values;
  __host__ __device__ cuco::pair<K, V> operator()(std::int64_t i) const
  {
    return cuco::pair<K, V>{keys[i], values[i]};
  }
};

struct nonzero {
  __host__ __device__ bool operator()(std::uint8_t b) const { return b != 0; }
};

template <typename container_t>
class table_impl final : public cuco_b200_table {
  using key_type    = typename container_t::key_type;
  using mapped_type = typename mapped_of<container_t>::type;
  static constexpr bool is_map   = !std::is_void_v<mapped_type>;
  static constexpr bool is_multimap_v = is_multimap<container_t>::value;
  static constexpr bool is_multi      = is_multiset<container_t>::value || is_multimap_v;  // duplicates
  using payload_t   = std::conditional_t<is_map, mapped_type, key_type>;  // what find() writes
  using slot_type   = typename container_t::value_type;

 public:
  table_impl(i64 size, double lf, i64 ek, i64 ev, int has_erased, i64 erased, void* stream)
    : c_{make(size, lf, ek, ev, has_erased, erased, stream)}
  {
  }

  int kind() const override { return CUCO_SHIM_KIND; }
  int key_bytes() const override { return sizeof(key_type); }
  int value_bytes() const override { return is_map ? sizeof(payload_t) : 0; }
  i64 capacity() const override { return static_cast<i64>(c_.capacity()); }
  i64 size(void* s) override
  {
    if constexpr (is_multimap_v) {
      // the reference class has no size(); count over the whole key domain is not expressible either
      (void)s;
      throw std::invalid_argument("size is not an experimental::static_multimap operation");
    } else {
      return static_cast<i64>(c_.size(sref(s)));
    }
  }
  void clear(void* s) override { c_.clear_async(sref(s)); }

  void insert(const void* keys, const void* values, i64 n, void* s, i64* num) override
  {
    with_input(keys, values, n, [&](auto first, auto last) {
      if (num) {
        if constexpr (is_multi) {
          c_.insert(first, last, sref(s));  // returns void: every element is stored
          *num = n;
        } else {
          *num = static_cast<i64>(c_.insert(first, last, sref(s)));
        }
      } else {
        c_.insert_async(first, last, sref(s));
      }
    });
  }

  void insert_if(const void* keys, const void* values, const std::uint8_t* st, i64 n, void* s, i64* num) override
  {
    with_input(keys, values, n, [&](auto first, auto last) {
      if (num) {
        *num = static_cast<i64>(c_.insert_if(first, last, st, nonzero{}, sref(s)));
      } else {
        c_.insert_if_async(first, last, st, nonzero{}, sref(s));
      }
    });
  }

  void find(const void* keys, void* out, i64 n, void* s) override
  {
    if constexpr (is_multimap_v) {
      throw std::invalid_argument("find is not an experimental::static_multimap operation");
    } else {
      auto const* k = static_cast<key_type const*>(keys);
      c_.find_async(k, k + n, static_cast<payload_t*>(out), sref(s));
    }
  }

  void contains(const void* keys, std::uint8_t* out, i64 n, void* s) override
  {
    auto const* k = static_cast<key_type const*>(keys);
    c_.contains_async(k, k + n, reinterpret_cast<bool*>(out), sref(s));
  }

  void contains_if(const void* keys, const std::uint8_t* st, std::uint8_t* out, i64 n, void* s) override
  {
    auto const* k = static_cast<key_type const*>(keys);
    c_.contains_if_async(k, k + n, st, nonzero{}, reinterpret_cast<bool*>(out), sref(s));
  }

  void insert_and_find(const void* keys, const void* values, void* found, std::uint8_t* inserted, i64 n, void* s) override
  {
#if defined(CUCO_SHIM_REFERENCE) && (CUCO_SHIM_KIND == 6)
    // nvcc 12.9 aborts ("Broken module found") on the reference's insert_and_find for 16-byte slots
    // with storage<2>; that one operation is left out of the reference build of this kind.
    (void)keys, (void)values, (void)found, (void)inserted, (void)n, (void)s;
    throw std::invalid_argument("insert_and_find: not compilable in the reference build of kind 6");
#else
    if constexpr (is_multi) {
      throw std::invalid_argument("insert_and_find is not a static_multiset operation");
    } else {
      with_input(keys, values, n, [&](auto first, auto last) {
        c_.insert_and_find_async(
          first, last, static_cast<payload_t*>(found), reinterpret_cast<bool*>(inserted), sref(s));
      });
    }
#endif
  }

  void insert_or_assign(const void* keys, const void* values, i64 n, void* s) override
  {
    if constexpr (is_map && !is_multi) {
      with_input(keys, values, n, [&](auto first, auto last) {
        c_.insert_or_assign_async(first, last, sref(s));
      });
    } else {
      throw std::invalid_argument("insert_or_assign is a static_map operation");
    }
  }

  void insert_or_apply(const void* keys, const void* values, i64 n, int op, int has_init, i64 init, void* s) override
  {
    if constexpr (is_map && !is_multi) {
      with_input(keys, values, n, [&](auto first, auto last) {
        auto run = [&](auto functor) {
          if (has_init) {
            c_.insert_or_apply_async(first, last, static_cast<payload_t>(init), functor, sref(s));
          } else {
            c_.insert_or_apply_async(first, last, functor, sref(s));
          }
        };
        switch (op) {
          case 0: run(cuco::reduce::plus{}); break;
          case 1: run(cuco::reduce::min{}); break;
          case 2: run(cuco::reduce::max{}); break;
          default: throw std::invalid_argument("unknown reduce op");
        }
      });
    } else {
      throw std::invalid_argument("insert_or_apply is a static_map operation");
    }
  }

  void erase(const void* keys, i64 n, void* s) override
  {
    if constexpr (is_multi) {
      throw std::invalid_argument("erase is not a static_multiset operation");
    } else {
      auto const* k = static_cast<key_type const*>(keys);
      c_.erase_async(k, k + n, sref(s));
    }
  }

  i64 retrieve_all(void* keys_out, void* values_out, void* s) override
  {
    auto* k = static_cast<key_type*>(keys_out);
    if constexpr (is_multi) {
      throw std::invalid_argument("retrieve_all is not a static_multiset operation");
    } else if constexpr (is_map) {
      auto const ends = c_.retrieve_all(k, static_cast<payload_t*>(values_out), sref(s));
      return static_cast<i64>(ends.first - k);
    } else {
      return static_cast<i64>(c_.retrieve_all(k, sref(s)) - k);
    }
  }

  void rehash(i64 capacity, void* s) override
  {
    if constexpr (is_multi) {
      throw std::invalid_argument("rehash is not a static_multiset operation");
    } else if (capacity < 0) {
      c_.rehash(sref(s));
    } else {
      c_.rehash(static_cast<typename container_t::size_type>(capacity), sref(s));
    }
  }

  i64 count(const void* keys, i64 n, bool outer, void* s) override
  {
    if constexpr (is_multimap_v) {
      auto const* k = static_cast<key_type const*>(keys);
      if (outer) { throw std::invalid_argument("count_outer is a static_multiset operation"); }
      return static_cast<i64>(c_.count(k, k + n, sref(s)));
    } else if constexpr (is_multi) {
      auto const* k = static_cast<key_type const*>(keys);
      if (outer) {
        return static_cast<i64>(c_.count_outer(k, k + n, c_.key_eq(), c_.hash_function(), sref(s)));
      }
      return static_cast<i64>(c_.count(k, k + n, sref(s)));
    } else {
      throw std::invalid_argument("count is a static_multiset operation");
    }
  }

  i64 retrieve(const void* keys, i64 n, bool outer, void* probe_out, void* match_out, void* s) override
  {
    auto const* k = static_cast<key_type const*>(keys);
    auto* p       = static_cast<key_type*>(probe_out);
    auto* m       = static_cast<key_type*>(match_out);
    if constexpr (is_multimap_v) {
      (void)k, (void)p, (void)m, (void)n, (void)outer, (void)s;
      throw std::invalid_argument("retrieve is a static_set / static_multiset operation");
    } else if constexpr (is_multi) {
      if (outer) {
        return static_cast<i64>(
          c_.retrieve_outer(k, k + n, c_.key_eq(), c_.hash_function(), p, m, sref(s)).first - p);
      }
      return static_cast<i64>(c_.retrieve(k, k + n, p, m, sref(s)).first - p);
    } else if constexpr (!is_map) {
      if (outer) { throw std::invalid_argument("retrieve_outer is a static_multiset operation"); }
      return static_cast<i64>(c_.retrieve(k, k + n, p, m, sref(s)).first - p);
    } else {
      throw std::invalid_argument("retrieve is a static_set / static_multiset operation");
    }
  }

  // ---- exchange path (no reference counterpart: our engine only) ----------------------------
#if defined(CUCO_SHIM_REFERENCE) || (CUCO_SHIM_KIND >= 10 && CUCO_SHIM_KIND <= 12)
  exchange_shape exchange_plan(i64, int) override { throw unsupported(); }
  void exchange_route(const void*, const void*, i64, bool, exchange_shape, int, int, std::uint64_t,
                      void* const*, void* const*, void* const*, void*, void*, void*, void*, void*,
                      void*) override
  {
    throw unsupported();
  }
  void exchange_mutate(const void*, const void*, exchange_shape, int, int, void*) override
  {
    throw unsupported();
  }
  void exchange_lookup(const void*, const void*, void* const*, exchange_shape, int, int, int, void*) override
  {
    throw unsupported();
  }
  void exchange_unpermute(const void*, const void*, i64, void*, int, void*) override { throw unsupported(); }
  exchange_shape stage_plan(i64, int, int) override { throw unsupported(); }
  void exchange_stage(const void*, const void*, i64, bool, exchange_shape, int, int, std::uint64_t, void*, void*,
                      void*, void*, void*, void*, void*) override
  {
    throw unsupported();
  }
  void exchange_apply(const void*, const void*, std::uint32_t, int, int, int, int, void*) override
  {
    throw unsupported();
  }
  void exchange_lookup_local(const void*, const void*, void*, std::uint32_t, int, int, void*) override
  {
    throw unsupported();
  }
  std::uint32_t exchange_fine_regions(int) override { throw unsupported(); }
  void exchange_probe(const void*, const void*, std::uint32_t, std::uint32_t, int, std::uint32_t, std::uint32_t, int,
                      void*) override
  {
    throw unsupported();
  }
  static std::invalid_argument unsupported()
  {
    return std::invalid_argument("the exchange path exists in the native build of maps and sets only");
  }
#else
  using engine_t = std::decay_t<decltype(std::declval<container_t&>().b200_engine())>;
  using plan_t   = typename engine_t::exchange_plan;
  static plan_t to_plan(exchange_shape s) { return plan_t{s.num_regions, s.segment_capacity, s.spill_capacity}; }
  static cuco::b200::exchange_peers to_peers(void* const* ptrs, int n)
  {
    cuco::b200::exchange_peers p{};
    for (int i = 0; i < n; ++i) { p.base[i] = ptrs[i]; }
    return p;
  }

  exchange_shape exchange_plan(i64 n_max, int num_ranks) override
  {
    auto const p = c_.b200_engine().plan_exchange(n_max, num_ranks);
    return exchange_shape{p.num_regions, p.segment_capacity, p.spill_capacity};
  }

  void exchange_route(const void* keys, const void* values, i64 n, bool keys_only, exchange_shape shape,
                      int num_ranks, int my_rank, std::uint64_t salt, void* const* peer_segments,
                      void* const* peer_counts, void* const* peer_flags, void* counts_local,
                      void* position_local, void* spill, void* spill_index, void* spill_count,
                      void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto run          = [&](auto keys_only_tag, auto first) {
      eng.template exchange_route_async<decltype(keys_only_tag)::value>(
        first, n, to_plan(shape), num_ranks, my_rank, salt, to_peers(peer_segments, num_ranks),
        to_peers(peer_counts, num_ranks), to_peers(peer_flags, num_ranks),
        static_cast<unsigned int*>(counts_local), static_cast<std::uint32_t*>(position_local), spill,
        static_cast<std::uint32_t*>(spill_index), static_cast<unsigned int*>(spill_count), handle, sref(s));
    };
    if (keys_only) {
      run(std::true_type{}, static_cast<key_type const*>(keys));
    } else {
      with_input(keys, values, n, [&](auto first, auto) { run(std::false_type{}, first); });
    }
  }

  void exchange_mutate(const void* segments, const void* counts_recv, exchange_shape shape, int num_ranks,
                       int op, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto const* segs  = static_cast<slot_type const*>(segments);
    auto const* cnts  = static_cast<unsigned int const*>(counts_recv);
    if (op < 0) {
      eng.exchange_mutate_async(segs, cnts, to_plan(shape), num_ranks, handle, cuco::b200::action_insert{}, sref(s));
      return;
    }
    if constexpr (is_map) {
      auto run = [&](auto functor) {
        eng.exchange_mutate_async(segs, cnts, to_plan(shape), num_ranks, handle,
                                  cuco::b200::action_apply<decltype(functor), false>{functor}, sref(s));
      };
      switch (op) {
        case 0: run(cuco::reduce::plus{}); break;
        case 1: run(cuco::reduce::min{}); break;
        case 2: run(cuco::reduce::max{}); break;
        default: throw std::invalid_argument("unknown reduce op");
      }
    } else {
      throw std::invalid_argument("insert_or_apply is a static_map operation");
    }
  }

  void exchange_lookup(const void* segments, const void* counts_recv, void* const* peer_results,
                       exchange_shape shape, int num_ranks, int my_rank, int what, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto const* segs  = static_cast<key_type const*>(segments);
    auto const* cnts  = static_cast<unsigned int const*>(counts_recv);
    if (what == 0) {
      eng.template exchange_lookup_async<payload_t>(segs, cnts, to_peers(peer_results, num_ranks), to_plan(shape),
                                                    num_ranks, my_rank, handle, eng.make_find_emit(), sref(s));
    } else {
      eng.template exchange_lookup_async<std::uint8_t>(segs, cnts, to_peers(peer_results, num_ranks),
                                                       to_plan(shape), num_ranks, my_rank, handle,
                                                       cuco::b200::emit_present{}, sref(s));
    }
  }

  exchange_shape stage_plan(i64 n_max, int num_ranks, int groups) override
  {
    auto const p = c_.b200_engine().plan_stage(n_max, num_ranks, groups);
    return exchange_shape{p.num_regions, p.segment_capacity, p.spill_capacity};
  }

  void exchange_stage(const void* keys, const void* values, i64 n, bool keys_only, exchange_shape shape,
                      int num_ranks, int my_rank, std::uint64_t salt, void* stage, void* counts_local,
                      void* position_local, void* spill, void* spill_index, void* spill_count, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    // owner o's block of the staging buffer: [slice][cap] elements
    std::size_t const elem = keys_only ? sizeof(key_type) : sizeof(slot_type);
    cuco::b200::exchange_peers blocks{};
    for (int o = 0; o < num_ranks; ++o) {
      blocks.base[o] = static_cast<char*>(stage) +
                       static_cast<std::size_t>(o) * shape.num_regions * shape.segment_capacity * elem;
    }
    auto run = [&](auto keys_only_tag, auto first) {
      eng.template exchange_route_async<decltype(keys_only_tag)::value>(
        first, n, to_plan(shape), num_ranks, my_rank, salt, blocks, cuco::b200::exchange_peers{},
        cuco::b200::exchange_peers{}, static_cast<unsigned int*>(counts_local),
        static_cast<std::uint32_t*>(position_local), spill, static_cast<std::uint32_t*>(spill_index),
        static_cast<unsigned int*>(spill_count), handle, sref(s), true);
    };
    if (keys_only) {
      run(std::true_type{}, static_cast<key_type const*>(keys));
    } else {
      with_input(keys, values, n, [&](auto first, auto) { run(std::false_type{}, first); });
    }
  }

  void exchange_apply(const void* segments, const void* counts_recv, std::uint32_t segment_capacity, int num_ranks,
                      int group, int groups, int op, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto const* segs  = static_cast<slot_type const*>(segments);
    auto const* cnts  = static_cast<unsigned int const*>(counts_recv);
    auto const g      = static_cast<std::uint32_t>(group);
    auto const gs     = static_cast<std::uint32_t>(groups);
    if (op < 0) {
      eng.exchange_apply_async(segs, cnts, segment_capacity, num_ranks, g, gs, handle, cuco::b200::action_insert{},
                               sref(s));
      return;
    }
    if constexpr (is_map) {
      auto run = [&](auto functor) {
        eng.exchange_apply_async(segs, cnts, segment_capacity, num_ranks, g, gs, handle,
                                 cuco::b200::action_apply<decltype(functor), false>{functor}, sref(s));
      };
      switch (op) {
        case 0: run(cuco::reduce::plus{}); break;
        case 1: run(cuco::reduce::min{}); break;
        case 2: run(cuco::reduce::max{}); break;
        default: throw std::invalid_argument("unknown reduce op");
      }
    } else {
      throw std::invalid_argument("insert_or_apply is a static_map operation");
    }
  }

  std::uint32_t exchange_fine_regions(int num_ranks) override
  {
    return c_.b200_engine().exchange_fine_regions(num_ranks);
  }

  void exchange_probe(const void* segments, const void* counts_recv, std::uint32_t num_regions,
                      std::uint32_t segment_capacity, int num_ranks, std::uint32_t region_begin,
                      std::uint32_t region_count, int op, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto const* segs  = static_cast<slot_type const*>(segments);
    auto const* cnts  = static_cast<unsigned int const*>(counts_recv);
    if (op < 0) {
      eng.exchange_probe_async(segs, cnts, num_regions, segment_capacity, num_ranks, region_begin, region_count,
                               handle, cuco::b200::action_insert{}, sref(s));
      return;
    }
    if constexpr (is_map) {
      auto run = [&](auto functor) {
        eng.exchange_probe_async(segs, cnts, num_regions, segment_capacity, num_ranks, region_begin, region_count,
                                 handle, cuco::b200::action_apply<decltype(functor), false>{functor}, sref(s));
      };
      switch (op) {
        case 0: run(cuco::reduce::plus{}); break;
        case 1: run(cuco::reduce::min{}); break;
        case 2: run(cuco::reduce::max{}); break;
        default: throw std::invalid_argument("unknown reduce op");
      }
    } else {
      throw std::invalid_argument("insert_or_apply is a static_map operation");
    }
  }

  void exchange_lookup_local(const void* segments, const void* counts_recv, void* results,
                             std::uint32_t segment_capacity, int num_ranks, int what, void* s) override
  {
    auto& eng         = c_.b200_engine();
    auto const handle = typename engine_t::engine_handle{eng.make_engine()};
    auto const* segs  = static_cast<key_type const*>(segments);
    auto const* cnts  = static_cast<unsigned int const*>(counts_recv);
    if (what == 0) {
      eng.exchange_lookup_local_async(segs, cnts, static_cast<payload_t*>(results), segment_capacity, num_ranks,
                                      handle, eng.make_find_emit(), sref(s));
    } else {
      eng.exchange_lookup_local_async(segs, cnts, static_cast<std::uint8_t*>(results), segment_capacity, num_ranks,
                                      handle, cuco::b200::emit_present{}, sref(s));
    }
  }

  void exchange_unpermute(const void* results, const void* position_local, i64 n, void* out, int what,
                          void* s) override
  {
    auto const* pos = static_cast<std::uint32_t const*>(position_local);
    if (what == 0) {
      engine_t::exchange_unpermute_async(static_cast<payload_t const*>(results), pos, n,
                                         static_cast<payload_t*>(out), sref(s));
    } else {
      engine_t::exchange_unpermute_async(static_cast<std::uint8_t const*>(results), pos, n,
                                         static_cast<std::uint8_t*>(out), sref(s));
    }
  }
#endif

 private:
  static cuda::stream_ref sref(void* s) { return cuda::stream_ref{static_cast<cudaStream_t>(s)}; }

  static container_t make(i64 size, double lf, i64 ek, i64 ev, int has_erased, i64 erased, void* s)
  {
    auto const extent = cuco::extent<std::size_t>{static_cast<std::size_t>(size < 0 ? 0 : size)};
    auto const ekey   = cuco::empty_key<key_type>{static_cast<key_type>(ek)};
    if constexpr (is_map) {
      auto const eval = cuco::empty_value<payload_t>{static_cast<payload_t>(ev)};
      if (has_erased) {
        return container_t{extent, ekey, eval, cuco::erased_key<key_type>{static_cast<key_type>(erased)},
                           {}, {}, {}, {}, {}, sref(s)};
      }
      if (lf > 0.0 || lf < 0.0) { return container_t{extent, lf, ekey, eval, {}, {}, {}, {}, {}, sref(s)}; }
      return container_t{extent, ekey, eval, {}, {}, {}, {}, {}, sref(s)};
    } else {
      (void)ev;
      if (has_erased) {
        return container_t{extent, ekey, cuco::erased_key<key_type>{static_cast<key_type>(erased)},
                           {}, {}, {}, {}, {}, sref(s)};
      }
      if (lf > 0.0 || lf < 0.0) { return container_t{extent, lf, ekey, {}, {}, {}, {}, {}, sref(s)}; }
      return container_t{extent, ekey, {}, {}, {}, {}, {}, sref(s)};
    }
  }

  /// Presents the caller's buffers as the [first, last) range the container expects.
  template <typename F>
  void with_input(const void* keys, const void* values, i64 n, F&& f)
  {
    if constexpr (is_map) {
      if (values == nullptr) {
        auto const* p = static_cast<slot_type const*>(keys);  // AoS cuco::pair<Key,T>
        f(p, p + n);
      } else {
        auto const zip = zip_to_pair<key_type, payload_t>{static_cast<key_type const*>(keys),
                                                          static_cast<payload_t const*>(values)};
        auto const first =
          thrust::make_transform_iterator(thrust::counting_iterator<std::int64_t>{0}, zip);
        f(first, first + n);
      }
    } else {
      auto const* k = static_cast<key_type const*>(keys);
      f(k, k + n);
    }
  }

  container_t c_;
};

}  // namespace

#define CUCO_SHIM_CAT2(a, b) a##b
#define CUCO_SHIM_CAT(a, b)  CUCO_SHIM_CAT2(a, b)

cuco_b200_table* CUCO_SHIM_CAT(cuco_shim_make_kind_, CUCO_SHIM_KIND)(std::int64_t size,
                                                                      double load_factor,
                                                                      std::int64_t empty_key,
                                                                      std::int64_t empty_value,
                                                                      int has_erased,
                                                                      std::int64_t erased_key,
                                                                      void* stream)
{
  return new table_impl<selected_container>(size, load_factor, empty_key, empty_value, has_erased, erased_key, stream);
}
#endif  // CUCO_SHIM_STUB
