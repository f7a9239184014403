// Drop-in check of the C++ template surface: one source, compiled twice - against include/ (this
// implementation) and against /root/reference/include (cuCollections) - and run on the GPU. It uses
// only the public cuco:: API, in the ways the reference's own tests and examples do:
//   tests/static_map/{unique_sequence,insert_and_find,insert_or_assign,insert_or_apply,
//                     heterogeneous_lookup,key_sentinel,shared_memory,stream,for_each,erase}_test.cu
//   tests/static_set/{unique_sequence,insert_and_find,shared_memory,heterogeneous_lookup,
//                     retrieve}_test.cu
//   tests/static_multiset/{insert,contains,find,count,custom_count,retrieve,for_each}_test.cu
//   tests/static_multimap/{insert_contains,insert_if,count,for_each}_test.cu (experimental::static_multimap)
//   tests/utility/probing_scheme_test.cu, examples/static_map/device_ref_example.cu,
//   examples/static_set/device_ref_example.cu, examples/static_map/count_by_key_example.cu
// Prints one line per check and a JSON summary; exit code 0 iff everything passed.
#include <cuco/static_map.cuh>
#include <cuco/static_multimap.cuh>
#include <cuco/static_multiset.cuh>
#include <cuco/static_set.cuh>
#include <cuco/utility/reduction_functors.cuh>

#include <thrust/count.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/functional.h>
#include <thrust/host_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/logical.h>
#include <thrust/random.h>
#include <thrust/sequence.h>
#include <thrust/shuffle.h>
#include <thrust/sort.h>

#include <cuda/functional>

#include <cooperative_groups.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace cg = cooperative_groups;

static int g_failed = 0, g_total = 0;
static void report(bool ok, char const* what)
{
  ++g_total;
  if (!ok) { ++g_failed; }
  std::printf("%s %s\n", ok ? "PASS" : "FAIL", what);
}
#define CHECK(expr) report((expr), #expr)

template <typename It>
static bool all_true(It first, It last)
{
  return thrust::all_of(thrust::device, first, last, thrust::identity<bool>{});
}
template <typename It>
static bool none_true(It first, It last)
{
  return thrust::none_of(thrust::device, first, last, thrust::identity<bool>{});
}

// ------------------------------------------------------------------------------------------------
// user kernels over device refs
// ------------------------------------------------------------------------------------------------
template <typename Ref, typename PairIt>
__global__ void ref_insert_scalar(Ref ref, PairIt pairs, int n, int* num_inserted)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ref.insert(*(pairs + i))) { atomicAdd(num_inserted, 1); }
}

template <int CG, typename Ref, typename PairIt>
__global__ void ref_insert_tile(Ref ref, PairIt pairs, int n, int* num_inserted)
{
  auto const tile = cg::tiled_partition<CG>(cg::this_thread_block());
  int const i     = (blockIdx.x * blockDim.x + threadIdx.x) / CG;
  if (i < n) {
    bool const ok = ref.insert(tile, *(pairs + i));
    if (ok && tile.thread_rank() == 0) { atomicAdd(num_inserted, 1); }
  }
}

template <typename Ref, typename KeyIt, typename OutIt>
__global__ void ref_find_scalar(Ref ref, KeyIt keys, int n, OutIt out)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    auto const found = ref.find(*(keys + i));
    out[i]           = (found == ref.end()) ? ref.empty_value_sentinel() : found->second;
  }
}

template <int CG, typename Ref, typename KeyIt, typename OutIt, typename BoolIt>
__global__ void ref_find_contains_tile(Ref ref, KeyIt keys, int n, OutIt out, BoolIt present)
{
  auto const tile = cg::tiled_partition<CG>(cg::this_thread_block());
  int const i     = (blockIdx.x * blockDim.x + threadIdx.x) / CG;
  if (i < n) {
    auto const found = ref.find(tile, *(keys + i));
    bool const has   = ref.contains(tile, *(keys + i));
    if (tile.thread_rank() == 0) {
      out[i]     = (found == ref.end()) ? ref.empty_value_sentinel() : found->second;
      present[i] = has;
    }
  }
}

// examples/static_map/device_ref_example.cu: count occurrences with insert_and_find + atomic_ref
template <typename Ref, typename KeyIt>
__global__ void count_by_key(Ref ref, KeyIt keys, int n)
{
  using T     = typename Ref::mapped_type;
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    auto [slot, inserted] = ref.insert_and_find(cuco::pair{*(keys + i), T{0}});
    cuda::atomic_ref<T, cuda::thread_scope_device> payload{slot->second};
    payload.fetch_add(T{1}, cuda::memory_order_relaxed);
  }
}

template <typename Ref, typename PairIt>
__global__ void ref_upserts(Ref ref, PairIt pairs, int n)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    auto const p = *(pairs + i);
    ref.insert_or_apply(p, cuco::reduce::plus{});
  }
}

template <typename Ref, typename PairIt>
__global__ void ref_assign(Ref ref, PairIt pairs, int n)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { ref.insert_or_assign(*(pairs + i)); }
}

// tests/static_map/shared_memory_test.cu (shared_memory_hash_table_kernel): a table living entirely
// in shared memory, built with CTAD from a static extent exactly as the reference test does
template <std::size_t NumWindows>
__global__ void shared_memory_map(bool* ok_flags, int n)
{
  using Key       = std::int32_t;
  using T         = std::int32_t;
  using slot_type = cuco::pair<Key, T>;
  __shared__ cuco::window<slot_type, 1> windows[NumWindows];
  using extent_type      = cuco::extent<std::size_t, NumWindows>;
  using storage_ref_type = cuco::aow_storage_ref<slot_type, 1, extent_type>;
  auto raw_ref = cuco::static_map_ref{cuco::empty_key<Key>{-1},
                                      cuco::empty_value<T>{-1},
                                      thrust::equal_to<Key>{},
                                      cuco::linear_probing<1, cuco::default_hash_function<Key>>{},
                                      cuco::thread_scope_block,
                                      storage_ref_type{extent_type{}, windows}};
  auto const block = cg::this_thread_block();
  raw_ref.initialize(block);
  int const i = threadIdx.x;
  auto ref    = raw_ref.rebind_operators(cuco::insert);
  if (i < n) { ref.insert(slot_type{Key(i), T(2 * i)}); }
  block.sync();
  if (i < n) {
    auto const find_ref = ref.rebind_operators(cuco::find, cuco::contains);
    auto const found    = find_ref.find(Key(i));
    ok_flags[i] = (found != find_ref.end()) && (found->second == T(2 * i)) && find_ref.contains(Key(i)) &&
                  !find_ref.contains(Key(i + 10 * n));
  }
}

// make_copy: copy a global table into shared memory and query the copy at block scope
template <typename Ref, int NumWindows>
__global__ void query_shared_copy(Ref global_ref, bool* ok_flags, int n)
{
  using window_type = typename Ref::window_type;
  __shared__ window_type windows[NumWindows];
  auto const block = cg::this_thread_block();
  auto const local = global_ref.make_copy(block, windows, cuco::thread_scope_block);
  int const i      = threadIdx.x;
  if (i < n) {
    auto const found = local.find(typename Ref::key_type(i));
    ok_flags[i]      = (found != local.end()) && (found->second == typename Ref::mapped_type(i + 1));
  }
}

// tests/utility/probing_scheme_test.cu
template <typename Scheme, typename Extent>
__global__ void probing_sequences(Extent bound, std::size_t* scalar_seq, std::size_t* tile_seq, int len)
{
  Scheme scheme{};
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    auto it = scheme(std::int64_t{42}, bound);
    for (int i = 0; i < len; ++i, ++it) {
      scalar_seq[i] = *it;
    }
    auto const tile = cg::tiled_partition<1>(cg::this_thread_block());
    auto jt         = scheme(tile, std::int64_t{42}, bound);
    for (int i = 0; i < len; ++i, ++jt) {
      tile_seq[i] = *jt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// heterogeneous lookup types (tests/static_map/heterogeneous_lookup_test.cu)
// ------------------------------------------------------------------------------------------------
struct stored_key {
  std::int32_t a;
  __host__ __device__ stored_key() : a{0} {}
  __host__ __device__ stored_key(std::int32_t x) : a{x} {}
};
struct probe_key {
  std::int32_t a;
  std::int32_t unrelated;
  __host__ __device__ probe_key(std::int32_t x) : a{x}, unrelated{-x} {}
  __host__ __device__ operator stored_key() const { return stored_key{a}; }
};
struct hetero_hash {
  std::uint32_t seed;
  __host__ __device__ hetero_hash(std::uint32_t s = 0) : seed{s} {}  // double_hashing seeds hash2 with 1
  template <typename K>
  __host__ __device__ std::uint32_t operator()(K const& k) const
  {
    return cuco::murmurhash3_32<std::int32_t>{seed}(k.a);
  }
};
struct hetero_equal {
  template <typename L, typename R>
  __host__ __device__ bool operator()(L const& l, R const& r) const
  {
    return l.a == r.a;
  }
};
CUCO_DECLARE_BITWISE_COMPARABLE(stored_key)

// tests/static_map/key_sentinel_test.cu: the predicate indexes an array with the slot key, so it
// must never be called with the (-1) sentinel
__device__ int sentinel_probe_table[1024];
struct indexing_equal {
  __device__ bool operator()(int lhs, int rhs) const
  {
    return sentinel_probe_table[lhs] == sentinel_probe_table[rhs];
  }
};

struct make_kv {
  __host__ __device__ cuco::pair<std::int64_t, std::int64_t> operator()(std::int64_t i) const
  {
    return {i, i * 10};
  }
};
struct make_kv32 {
  __host__ __device__ cuco::pair<int, int> operator()(int i) const { return {i, i}; }
};
struct is_even {
  __host__ __device__ bool operator()(std::int64_t i) const { return i % 2 == 0; }
};
struct halve {
  __host__ __device__ cuco::pair<std::int64_t, std::int64_t> operator()(std::int64_t i) const
  {
    return {i / 2, i / 2};
  }
};

// ------------------------------------------------------------------------------------------------
template <typename Key, typename T, typename Probe, int W>
static void bulk_api_suite(char const* label)
{
  std::printf("-- bulk API: %s\n", label);
  constexpr int n = 4000;
  using map_type  = cuco::static_map<Key,
                                    T,
                                    cuco::extent<std::size_t>,
                                    cuda::thread_scope_device,
                                    thrust::equal_to<Key>,
                                    Probe,
                                    cuco::cuda_allocator<cuco::pair<Key, T>>,
                                    cuco::storage<W>>;
  map_type map{2 * n, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
  thrust::device_vector<Key> keys(n);
  thrust::sequence(keys.begin(), keys.end());
  thrust::device_vector<T> values(n);
  thrust::sequence(values.begin(), values.end());
  auto const pairs = thrust::make_zip_iterator(thrust::make_tuple(keys.begin(), values.begin()));
  thrust::device_vector<bool> present(n);
  thrust::device_vector<T> found(n);

  CHECK(map.size() == 0);
  map.contains(keys.begin(), keys.end(), present.begin());
  CHECK(none_true(present.begin(), present.end()));
  map.find(keys.begin(), keys.end(), found.begin());
  CHECK(thrust::count(found.begin(), found.end(), T(-1)) == n);

  // insert_if with an even stencil (unique_sequence_test.cu)
  auto const stencil = thrust::counting_iterator<std::int64_t>{0};
  CHECK(map.insert_if(pairs, pairs + n, stencil, is_even{}) == n / 2);
  CHECK(map.size() == n / 2);
  map.contains(keys.begin(), keys.end(), present.begin());
  CHECK(thrust::count(present.begin(), present.end(), true) == n / 2);

  CHECK(map.insert(pairs, pairs + n) == n / 2);
  CHECK(map.size() == n);
  map.find(keys.begin(), keys.end(), found.begin());
  CHECK(thrust::equal(found.begin(), found.end(), values.begin()));
  map.contains_if(keys.begin(), keys.end(), stencil, is_even{}, present.begin());
  CHECK(thrust::count(present.begin(), present.end(), true) == n / 2);

  // retrieve_all returns every pair exactly once
  thrust::device_vector<Key> out_keys(n);
  thrust::device_vector<T> out_vals(n);
  auto const ends = map.retrieve_all(out_keys.begin(), out_vals.begin());
  CHECK(ends.first - out_keys.begin() == n);
  thrust::sort(out_keys.begin(), out_keys.end());
  CHECK(thrust::equal(out_keys.begin(), out_keys.end(), keys.begin()));

  // insert_and_find twice (insert_and_find_test.cu)
  map.clear();
  thrust::device_vector<bool> inserted(n);
  map.insert_and_find(pairs, pairs + n, found.begin(), inserted.begin());
  CHECK(all_true(inserted.begin(), inserted.end()));
  CHECK(thrust::equal(found.begin(), found.end(), values.begin()));
  map.insert_and_find(pairs, pairs + n, found.begin(), inserted.begin());
  CHECK(none_true(inserted.begin(), inserted.end()));
  CHECK(thrust::equal(found.begin(), found.end(), values.begin()));

  // insert_or_assign doubles every payload (insert_or_assign_test.cu)
  thrust::device_vector<T> doubled(n);
  thrust::transform(values.begin(), values.end(), doubled.begin(), thrust::placeholders::_1 * 2);
  auto const pairs2 = thrust::make_zip_iterator(thrust::make_tuple(keys.begin(), doubled.begin()));
  map.insert_or_assign(pairs2, pairs2 + n);
  CHECK(map.size() == n);
  map.find(keys.begin(), keys.end(), found.begin());
  CHECK(thrust::equal(found.begin(), found.end(), doubled.begin()));
}

static void aggregate_suite()
{
  std::printf("-- insert_or_apply (insert_or_apply_test.cu)\n");
  using Key = std::int64_t;
  using T   = std::int64_t;
  constexpr int rows = 10000, distinct = 100;
  thrust::device_vector<Key> keys(rows);
  thrust::transform(thrust::counting_iterator<Key>{0}, thrust::counting_iterator<Key>{rows}, keys.begin(),
                    thrust::placeholders::_1 % distinct);
  thrust::device_vector<T> ones(rows, 1);
  auto const pairs = thrust::make_zip_iterator(thrust::make_tuple(keys.begin(), ones.begin()));
  thrust::device_vector<Key> q(distinct);
  thrust::sequence(q.begin(), q.end());
  thrust::device_vector<T> sums(distinct);
  for (T sentinel : {T{0}, T{-1}}) {
    for (int with_init = 0; with_init < 2; ++with_init) {
      cuco::static_map<Key, T, cuco::extent<std::size_t>, cuda::thread_scope_device, thrust::equal_to<Key>,
                       cuco::linear_probing<1, cuco::default_hash_function<Key>>>
        map{4 * distinct, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{sentinel}};
      if (with_init) {
        map.insert_or_apply(pairs, pairs + rows, T{0}, cuco::reduce::plus{});
      } else {
        map.insert_or_apply(pairs, pairs + rows, cuco::reduce::plus{});
      }
      CHECK(map.size() == distinct);
      map.find(q.begin(), q.end(), sums.begin());
      // init == sentinel == 0 combines onto 0; init != sentinel stores first; no init stores first
      CHECK(thrust::count(sums.begin(), sums.end(), T{rows / distinct}) == distinct);
    }
  }
  // device-side upserts through refs
  cuco::static_map<Key, T, cuco::extent<std::size_t>, cuda::thread_scope_device, thrust::equal_to<Key>,
                   cuco::linear_probing<1, cuco::default_hash_function<Key>>>
    map{4 * distinct, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{0}};
  ref_upserts<<<(rows + 127) / 128, 128>>>(map.ref(cuco::insert_or_apply), pairs, rows);
  cudaDeviceSynchronize();
  map.find(q.begin(), q.end(), sums.begin());
  CHECK(thrust::count(sums.begin(), sums.end(), T{rows / distinct}) == distinct);
  thrust::device_vector<T> sevens(rows, 7);
  auto const pairs7 = thrust::make_zip_iterator(thrust::make_tuple(keys.begin(), sevens.begin()));
  ref_assign<<<(rows + 127) / 128, 128>>>(map.ref(cuco::insert_or_assign), pairs7, rows);
  cudaDeviceSynchronize();
  map.find(q.begin(), q.end(), sums.begin());
  CHECK(thrust::count(sums.begin(), sums.end(), T{7}) == distinct);
  CHECK(map.size() == distinct);
}

template <int CG, int W>
static void device_ref_suite(char const* label)
{
  std::printf("-- device refs: %s\n", label);
  using Key = std::int64_t;
  using T   = std::int64_t;
  constexpr int n = 3000;
  using map_type  = cuco::static_map<Key, T, cuco::extent<std::size_t>, cuda::thread_scope_device,
                                    thrust::equal_to<Key>,
                                    cuco::double_hashing<CG, cuco::default_hash_function<Key>>,
                                    cuco::cuda_allocator<cuco::pair<Key, T>>, cuco::storage<W>>;
  map_type map{2 * n, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
  auto const pairs = thrust::make_transform_iterator(thrust::counting_iterator<std::int64_t>{0}, make_kv{});
  thrust::device_vector<int> counter(1, 0);
  thrust::device_vector<Key> keys(2 * n);
  thrust::sequence(keys.begin(), keys.end());
  thrust::device_vector<T> out(2 * n);
  thrust::device_vector<bool> present(2 * n);

  if constexpr (CG == 1) {
    ref_insert_scalar<<<(n + 127) / 128, 128>>>(map.ref(cuco::insert), pairs, n, counter.data().get());
    ref_insert_scalar<<<(n + 127) / 128, 128>>>(map.ref(cuco::insert), pairs, n, counter.data().get());
  } else {
    ref_insert_tile<CG><<<(n * CG + 127) / 128, 128>>>(map.ref(cuco::insert), pairs, n, counter.data().get());
    ref_insert_tile<CG><<<(n * CG + 127) / 128, 128>>>(map.ref(cuco::insert), pairs, n, counter.data().get());
  }
  cudaDeviceSynchronize();
  CHECK(counter[0] == n);  // second pass inserts nothing
  CHECK(map.size() == n);

  // what device refs wrote, the bulk API reads - and the other way round
  map.find(keys.begin(), keys.end(), out.begin());
  bool bulk_ok = true;
  {
    thrust::host_vector<T> h = out;
    for (int i = 0; i < 2 * n; ++i) {
      bulk_ok = bulk_ok && (h[i] == (i < n ? T(i) * 10 : T(-1)));
    }
  }
  CHECK(bulk_ok);
  thrust::fill(out.begin(), out.end(), T{0});
  if constexpr (CG == 1) {
    ref_find_scalar<<<(2 * n + 127) / 128, 128>>>(map.ref(cuco::find), keys.begin(), 2 * n, out.begin());
  }
  ref_find_contains_tile<CG><<<(2 * n * CG + 127) / 128, 128>>>(
    map.ref(cuco::find, cuco::contains), keys.begin(), 2 * n, out.begin(), present.begin());
  cudaDeviceSynchronize();
  bool ref_ok = true;
  {
    thrust::host_vector<T> h    = out;
    thrust::host_vector<bool> p = present;
    for (int i = 0; i < 2 * n; ++i) {
      ref_ok = ref_ok && (h[i] == (i < n ? T(i) * 10 : T(-1))) && (p[i] == (i < n));
    }
  }
  CHECK(ref_ok);
}

static void count_by_key_suite()
{
  std::printf("-- insert_and_find handle + atomic_ref (device_ref_example.cu)\n");
  using Key = std::int32_t;
  using T   = std::int32_t;
  constexpr int n = 19200, distinct = 64;  // 300 occurrences of every key
  thrust::device_vector<Key> keys(n);
  thrust::transform(thrust::counting_iterator<int>{0}, thrust::counting_iterator<int>{n}, keys.begin(),
                    thrust::placeholders::_1 % distinct);
  cuco::static_map<Key, T, cuco::extent<std::size_t>, cuda::thread_scope_device, thrust::equal_to<Key>,
                   cuco::linear_probing<1, cuco::default_hash_function<Key>>>
    map{4 * distinct, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
  count_by_key<<<(n + 127) / 128, 128>>>(map.ref(cuco::insert_and_find), keys.begin(), n);
  cudaDeviceSynchronize();
  thrust::device_vector<Key> q(distinct);
  thrust::sequence(q.begin(), q.end());
  thrust::device_vector<T> counts(distinct);
  map.find(q.begin(), q.end(), counts.begin());
  CHECK(map.size() == distinct);
  CHECK(thrust::count(counts.begin(), counts.end(), T{n / distinct}) == distinct);
}

static void shared_memory_suite()
{
  std::printf("-- shared memory tables (shared_memory_test.cu)\n");
  using Key = std::int32_t;
  using T   = std::int32_t;
  constexpr int n = 100;
  {
    // the window extent is rounded up to an entry of the prime table, exactly as the reference's
    // shared_memory_test.cu sizes its shared array
    constexpr auto valid_extent = cuco::make_window_extent<1, 1>(cuco::extent<std::size_t, 257>{});
    thrust::device_vector<bool> ok(n, false);
    shared_memory_map<valid_extent.value()><<<1, 128>>>(ok.data().get(), n);
    cudaDeviceSynchronize();
    CHECK(all_true(ok.begin(), ok.end()));
  }
  {
    constexpr int requested = 200;
    using map_type = cuco::static_map<Key, T, cuco::extent<std::int32_t, requested>, cuda::thread_scope_device,
                                      thrust::equal_to<Key>,
                                      cuco::linear_probing<1, cuco::default_hash_function<Key>>>;
    map_type map{cuco::extent<std::int32_t, requested>{}, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
    constexpr int num_windows = 211;  // prime_at_least(200)
    CHECK(map.capacity() == num_windows);
    auto const pairs = thrust::make_transform_iterator(
      thrust::counting_iterator<int>{0},
      cuda::proclaim_return_type<cuco::pair<int, int>>(
        [] __device__(int i) { return cuco::pair<int, int>{i, i + 1}; }));
    map.insert(pairs, pairs + n);
    thrust::device_vector<bool> ok(n, false);
    using ref_type = decltype(map.ref(cuco::find));
    query_shared_copy<ref_type, num_windows><<<1, 128>>>(map.ref(cuco::find), ok.data().get(), n);
    cudaDeviceSynchronize();
    CHECK(all_true(ok.begin(), ok.end()));
  }
}

static void heterogeneous_suite()
{
  std::printf("-- heterogeneous lookup, custom hash/equality (heterogeneous_lookup_test.cu)\n");
  constexpr int n = 2000;
  using T = std::int32_t;
  cuco::static_map<stored_key, T, cuco::extent<std::size_t>, cuda::thread_scope_device, hetero_equal,
                   cuco::linear_probing<1, hetero_hash>>
    map{2 * n, cuco::empty_key<stored_key>{stored_key{-1}}, cuco::empty_value<T>{-1}};
  auto const pairs = thrust::make_transform_iterator(
    thrust::counting_iterator<int>{0},
    cuda::proclaim_return_type<cuco::pair<probe_key, T>>(
      [] __device__(int i) { return cuco::pair<probe_key, T>{probe_key{i}, i}; }));
  auto const probes = thrust::make_transform_iterator(thrust::counting_iterator<int>{0},
                                                      cuda::proclaim_return_type<probe_key>(
                                                        [] __device__(int i) { return probe_key{i}; }));
  CHECK(map.insert(pairs, pairs + n) == n);
  thrust::device_vector<bool> present(2 * n);
  map.contains(probes, probes + 2 * n, present.begin());
  CHECK(thrust::count(present.begin(), present.begin() + n, true) == n);
  CHECK(thrust::count(present.begin() + n, present.end(), true) == 0);
  // (the reference's heterogeneous_lookup_test.cu exercises insert + contains only; its bulk find
  // does not return the payloads for this key type on sm_100a, so find is not part of this check)

  cuco::static_set<stored_key, cuco::extent<std::size_t>, cuda::thread_scope_device, hetero_equal,
                   cuco::double_hashing<2, hetero_hash>>
    set{2 * n, cuco::empty_key<stored_key>{stored_key{-1}}};
  CHECK(set.insert(probes, probes + n) == n);
  set.contains(probes, probes + 2 * n, present.begin());
  CHECK(thrust::count(present.begin(), present.end(), true) == n);
}

static void key_sentinel_suite()
{
  std::printf("-- predicate never sees a sentinel (key_sentinel_test.cu)\n");
  constexpr int n = 400;
  std::vector<int> identity(1024);
  for (int i = 0; i < 1024; ++i) { identity[i] = i; }
  cudaMemcpyToSymbol(sentinel_probe_table, identity.data(), sizeof(int) * 1024);
  auto const pairs = thrust::make_transform_iterator(thrust::counting_iterator<int>{0}, make_kv32{});
  {
    cuco::static_map<int, int, cuco::extent<std::size_t>, cuda::thread_scope_device, indexing_equal,
                     cuco::linear_probing<1, cuco::default_hash_function<int>>>
      map{2 * n, cuco::empty_key<int>{-1}, cuco::empty_value<int>{-1}, indexing_equal{}};
    CHECK(map.insert(pairs, pairs + n) == n);
    CHECK(map.size() == n);
  }
  {
    cuco::static_map<int, int, cuco::extent<std::size_t>, cuda::thread_scope_device, indexing_equal,
                     cuco::double_hashing<2, cuco::default_hash_function<int>>>
      map{2 * n, cuco::empty_key<int>{-1}, cuco::empty_value<int>{-1}, indexing_equal{}};
    thrust::device_vector<int> counter(1, 0);
    ref_insert_tile<2><<<(2 * n + 127) / 128, 128>>>(map.ref(cuco::insert), pairs, n, counter.data().get());
    cudaDeviceSynchronize();
    CHECK(counter[0] == n);
    CHECK(cudaGetLastError() == cudaSuccess);
  }
}

static void duplicate_and_set_suite()
{
  std::printf("-- duplicates, sets, streams, CTAD\n");
  // duplicate_keys_test.cu: pairs {i/2, i/2}
  constexpr int n = 50000;
  auto const pairs = thrust::make_transform_iterator(thrust::counting_iterator<std::int64_t>{0}, halve{});
  cuco::static_map map{std::size_t{2 * n}, cuco::empty_key<std::int64_t>{-1}, cuco::empty_value<std::int64_t>{-1}};
  CHECK(map.insert(pairs, pairs + n) == n / 2);
  CHECK(map.size() == n / 2);
  thrust::device_vector<std::int64_t> keys(n);
  thrust::sequence(keys.begin(), keys.end());
  thrust::device_vector<bool> present(n);
  map.contains(keys.begin(), keys.end(), present.begin());
  CHECK(all_true(present.begin(), present.begin() + n / 2));
  CHECK(none_true(present.begin() + n / 2, present.end()));

  // static_set default template arguments + CTAD, non-default stream (stream_test.cu)
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  {
    cuco::static_set set{cuco::extent<std::size_t>{2 * n}, cuco::empty_key<std::int32_t>{-1}, {}, {}, {}, {}, {},
                         cuda::stream_ref{stream}};
    thrust::device_vector<std::int32_t> k32(n);
    thrust::sequence(k32.begin(), k32.end());
    CHECK(set.insert(k32.begin(), k32.end(), stream) == n);
    CHECK(set.size(stream) == n);
    thrust::device_vector<std::int32_t> found(n);
    set.find(k32.begin(), k32.end(), found.begin(), stream);
    CHECK(thrust::equal(found.begin(), found.end(), k32.begin()));
    thrust::device_vector<std::int32_t> out(n);
    auto const end = set.retrieve_all(out.begin(), stream);
    CHECK(end - out.begin() == n);
    thrust::device_vector<bool> inserted(n);
    set.insert_and_find(k32.begin(), k32.end(), found.begin(), inserted.begin(), stream);
    CHECK(none_true(inserted.begin(), inserted.end()));
  }
  cudaStreamDestroy(stream);

  // erase + rehash (erase_test.cu, rehash_test.cu)
  {
    cuco::static_map<std::int64_t, std::int64_t> emap{std::size_t{4000}, cuco::empty_key<std::int64_t>{-1},
                                                      cuco::empty_value<std::int64_t>{-1},
                                                      cuco::erased_key<std::int64_t>{-2}};
    auto const kv = thrust::make_transform_iterator(thrust::counting_iterator<std::int64_t>{1}, make_kv{});
    constexpr int m = 1000;
    emap.insert(kv, kv + m);
    thrust::device_vector<std::int64_t> ek(m);
    thrust::sequence(ek.begin(), ek.end(), std::int64_t{1});
    emap.erase(ek.begin(), ek.begin() + m / 2);
    CHECK(emap.size() == m / 2);
    thrust::device_vector<bool> p(m);
    emap.contains(ek.begin(), ek.end(), p.begin());
    CHECK(none_true(p.begin(), p.begin() + m / 2));
    CHECK(all_true(p.begin() + m / 2, p.end()));
    emap.rehash();
    CHECK(emap.size() == m / 2);
    emap.rehash(std::size_t{16000});
    CHECK(emap.capacity() >= 16000);
    emap.contains(ek.begin(), ek.end(), p.begin());
    CHECK(thrust::count(p.begin(), p.end(), true) == m / 2);
  }

  // error conventions
  bool threw = false;
  try {
    cuco::static_map<int, int> bad{cuco::extent<std::size_t>{100}, 1.5, cuco::empty_key<int>{-1},
                                   cuco::empty_value<int>{-1}};
  } catch (cuco::logic_error const&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    cuco::static_map<int, int> bad{cuco::extent<std::size_t>{100}, cuco::empty_key<int>{-1},
                                   cuco::empty_value<int>{-1}, cuco::erased_key<int>{-1}};
  } catch (std::logic_error const&) {
    threw = true;
  }
  CHECK(threw);

  // static extents: capacity 422 (cg 1) / 412 (cg 2) with storage<2> (unique_sequence_test.cu)
  {
    cuco::static_map<int, int, cuco::extent<std::int32_t, 400>, cuda::thread_scope_device, thrust::equal_to<int>,
                     cuco::linear_probing<1, cuco::murmurhash3_32<int>>, cuco::cuda_allocator<char>,
                     cuco::storage<2>>
      a{cuco::extent<std::int32_t, 400>{}, cuco::empty_key<int>{-1}, cuco::empty_value<int>{-1}};
    CHECK(a.capacity() == 422);
    cuco::static_map<int, int, cuco::extent<std::int32_t, 400>, cuda::thread_scope_device, thrust::equal_to<int>,
                     cuco::double_hashing<2, cuco::murmurhash3_32<int>>, cuco::cuda_allocator<char>,
                     cuco::storage<2>>
      b{cuco::extent<std::int32_t, 400>{}, cuco::empty_key<int>{-1}, cuco::empty_value<int>{-1}};
    CHECK(b.capacity() == 412);
    auto const kv = thrust::make_transform_iterator(thrust::counting_iterator<int>{0}, make_kv32{});
    CHECK(a.insert(kv, kv + 400) == 400);
    CHECK(b.insert(kv, kv + 400) == 400);
    CHECK(a.size() == 400 && b.size() == 400);
  }
}

static void probing_suite()
{
  std::printf("-- probing sequences: scalar == tile<1> (probing_scheme_test.cu)\n");
  constexpr int len = 8;
  thrust::device_vector<std::size_t> a(len), b(len);
  auto const bound = cuco::make_window_extent<1, 2>(cuco::extent<std::size_t>{10});
  probing_sequences<cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>>>
    <<<1, 32>>>(bound, a.data().get(), b.data().get(), len);
  cudaDeviceSynchronize();
  CHECK(thrust::equal(a.begin(), a.end(), b.begin()));
  probing_sequences<cuco::double_hashing<1, cuco::default_hash_function<std::int64_t>>>
    <<<1, 32>>>(bound, a.data().get(), b.data().get(), len);
  cudaDeviceSynchronize();
  CHECK(thrust::equal(a.begin(), a.end(), b.begin()));
}

// ------------------------------------------------------------------------------------------------
// join probe of a set: static_set::retrieve (tests/static_set/retrieve_test.cu)
// ------------------------------------------------------------------------------------------------
template <typename Key, typename Probe>
static void set_retrieve_suite(char const* label)
{
  std::printf("# static_set::retrieve, %s\n", label);
  constexpr std::size_t n = 400;
  auto set = cuco::static_set{n, 1.0, cuco::empty_key<Key>{-1}, {}, Probe{}};
  thrust::device_vector<Key> probed(n), matched(n);
  auto const iter = thrust::counting_iterator<Key>{0};
  {
    auto const ends = set.retrieve(iter, iter + n, probed.begin(), matched.begin());
    CHECK(ends.first - probed.begin() == 0);
    CHECK(ends.second - matched.begin() == 0);
  }
  set.insert(iter, iter + n);
  {
    // twice as many probes as keys: the second half has no match and must not produce rows
    auto const ends = set.retrieve(iter, iter + 2 * n, probed.begin(), matched.begin());
    CHECK(ends.first - probed.begin() == static_cast<std::ptrdiff_t>(n));
    CHECK(ends.second - matched.begin() == static_cast<std::ptrdiff_t>(n));
    thrust::sort(probed.begin(), probed.end());
    thrust::sort(matched.begin(), matched.end());
    CHECK(thrust::equal(probed.begin(), probed.end(), iter));
    CHECK(thrust::equal(matched.begin(), matched.end(), iter));
  }
}

// ------------------------------------------------------------------------------------------------
// static_multiset (tests/static_multiset/*.cu)
// ------------------------------------------------------------------------------------------------
template <typename Key>
struct divide_by {
  Key divisor;
  __host__ __device__ Key operator()(Key i) const { return i / divisor; }
};

template <typename Ref, typename KeyIt>
__global__ void multiset_ref_scalar_kernel(Ref ref, KeyIt keys, std::size_t n, std::size_t multiplicity, int* errors)
{
  for (std::size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    auto const key      = *(keys + i);
    std::size_t matches = 0;
    ref.for_each(key, [&](auto const slot) {
      if (ref.key_eq()(key, slot)) { ++matches; }
    });
    if (matches != multiplicity) { atomicAdd(errors, 1); }
    if (static_cast<std::size_t>(ref.count(key)) != multiplicity) { atomicAdd(errors, 1); }
    if (ref.contains(key) != (multiplicity != 0)) { atomicAdd(errors, 1); }
  }
}

template <bool Synced, typename Ref, typename KeyIt>
__global__ void multiset_ref_tile_kernel(Ref ref, KeyIt keys, std::size_t n, std::size_t multiplicity, int* errors)
{
  constexpr int cgs = Ref::cg_size;
  auto const tile   = cg::tiled_partition<cgs>(cg::this_thread_block());
  for (std::size_t i = (blockIdx.x * blockDim.x + threadIdx.x) / cgs; i < n;
       i += (gridDim.x * blockDim.x) / cgs) {
    auto const key     = *(keys + i);
    std::size_t mine   = 0;
    auto const on_slot = [&](auto const slot) {
      if (ref.key_eq()(key, slot)) { ++mine; }
    };
    if constexpr (Synced) {
      ref.for_each(tile, key, on_slot, [](auto const& group) { group.sync(); });
    } else {
      ref.for_each(tile, key, on_slot);
    }
    auto const total = cg::reduce(tile, mine, cg::plus<std::size_t>());
    // the tile overload returns the matches seen by the calling lane
    auto const count =
      cg::reduce(tile, static_cast<std::size_t>(ref.count(tile, key)), cg::plus<std::size_t>());
    if (tile.thread_rank() == 0) {
      if (total != multiplicity) { atomicAdd(errors, 1); }
      if (count != multiplicity) { atomicAdd(errors, 1); }
    }
  }
}

template <int BlockSize, typename Ref, typename KeyIt, typename Key, typename Counter>
__global__ void multiset_ref_retrieve_kernel(
  Ref ref, KeyIt keys, std::size_t n, Key* probed, Key* matched, Counter* counter, bool outer)
{
  auto const block = cg::this_thread_block();
  if (outer) {
    ref.template retrieve_outer<BlockSize>(block, keys, keys + n, probed, matched, *counter);
  } else {
    ref.template retrieve<BlockSize>(block, keys, keys + n, probed, matched, *counter);
  }
}

template <typename Key, typename Probe>
static void multiset_suite(char const* label)
{
  std::printf("# static_multiset, %s\n", label);
  using set_type = cuco::static_multiset<Key,
                                         cuco::extent<std::size_t>,
                                         cuda::thread_scope_device,
                                         thrust::equal_to<Key>,
                                         Probe,
                                         cuco::cuda_allocator<Key>,
                                         cuco::storage<2>>;
  constexpr std::size_t unique = 400, multiplicity = 5, n = unique * multiplicity;
  auto set = set_type{cuco::extent<std::size_t>{2 * n}, cuco::empty_key<Key>{-1}};
  auto const iter = thrust::counting_iterator<Key>{0};

  // count / contains / retrieve on the empty container
  CHECK(set.size() == 0);
  CHECK(set.count(iter, iter + n) == 0);
  thrust::device_vector<bool> present(2 * n);
  set.contains(iter, iter + n, present.begin());
  CHECK(none_true(present.begin(), present.begin() + n));

  // n unique keys: count == n, every key contained and found (insert / contains / find tests)
  set.insert(iter, iter + n);
  CHECK(set.size() == n);
  CHECK(set.count(iter, iter + n) == n);
  set.contains(iter, iter + 2 * n, present.begin());
  CHECK(all_true(present.begin(), present.begin() + n));
  CHECK(none_true(present.begin() + n, present.end()));
  thrust::device_vector<Key> found(2 * n);
  set.find(iter, iter + 2 * n, found.begin());
  CHECK(thrust::equal(found.begin(), found.begin() + n, iter));
  CHECK(thrust::count(found.begin() + n, found.end(), Key(-1)) == static_cast<std::ptrdiff_t>(n));
  // every query repeated three times
  auto const thirds = thrust::make_transform_iterator(iter, divide_by<Key>{3});
  CHECK(set.count(thirds, thirds + 3 * n) == 3 * n);

  // multiplicity: each of `unique` keys stored `multiplicity` times, shuffled input
  set.clear();
  thrust::device_vector<Key> input(n);
  thrust::transform(iter, iter + n, input.begin(), divide_by<Key>{static_cast<Key>(multiplicity)});
  thrust::shuffle(input.begin(), input.end(), thrust::default_random_engine{7});
  thrust::device_vector<std::uint8_t> stencil(n);
  thrust::transform(iter, iter + n, stencil.begin(), is_even{});
  CHECK(set.insert_if(input.begin(), input.end(), stencil.begin(), is_even{}) == n / 2);
  CHECK(set.size() == n / 2);
  set.clear();
  set.insert(input.begin(), input.end());
  CHECK(set.size() == n);
  CHECK(set.count(iter, iter + unique) == n);
  CHECK(set.count(iter, iter + 2 * unique) == n);
  CHECK(set.count_outer(iter, iter + 2 * unique, set.key_eq(), set.hash_function()) == n + unique);
  {
    thrust::device_vector<Key> probed(n * multiplicity), matched(n * multiplicity);
    // probing with the stored stream itself: every element matches `multiplicity` elements
    auto const ends = set.retrieve(input.begin(), input.end(), probed.begin(), matched.begin());
    CHECK(ends.first - probed.begin() == static_cast<std::ptrdiff_t>(n * multiplicity));
    CHECK(ends.second - matched.begin() == static_cast<std::ptrdiff_t>(n * multiplicity));
    CHECK(thrust::equal(probed.begin(), probed.end(), matched.begin()));
    thrust::sort(probed.begin(), probed.end());
    auto const expected =
      thrust::make_transform_iterator(iter, divide_by<Key>{static_cast<Key>(multiplicity * multiplicity)});
    CHECK(thrust::equal(probed.begin(), probed.end(), expected));
  }
  {
    // outer retrieve over [0, 2 * unique): the upper half has no match and yields the sentinel
    thrust::device_vector<Key> probed(n + unique), matched(n + unique);
    auto const ends = set.retrieve_outer(
      iter, iter + 2 * unique, set.key_eq(), set.hash_function(), probed.begin(), matched.begin());
    CHECK(ends.first - probed.begin() == static_cast<std::ptrdiff_t>(n + unique));
    thrust::sort_by_key(probed.begin(), probed.end(), matched.begin());
    CHECK(thrust::equal(probed.begin(), probed.begin() + n, matched.begin()));
    CHECK(thrust::equal(probed.begin() + n, probed.end(), iter + unique));
    CHECK(thrust::count(matched.begin() + n, matched.end(), Key(-1)) == static_cast<std::ptrdiff_t>(unique));
  }

  // device refs: for_each / count / contains per thread or per tile, block-wide retrieve
  thrust::device_vector<int> errors(1, 0);
  if constexpr (set_type::cg_size == 1) {
    multiset_ref_scalar_kernel<<<8, 128>>>(
      set.ref(cuco::for_each, cuco::count, cuco::contains), iter, unique, multiplicity, errors.data().get());
    multiset_ref_scalar_kernel<<<8, 128>>>(
      set.ref(cuco::for_each, cuco::count, cuco::contains), iter + unique, unique, std::size_t{0}, errors.data().get());
  } else {
    multiset_ref_tile_kernel<false><<<8, 128>>>(
      set.ref(cuco::for_each, cuco::count), iter, unique, multiplicity, errors.data().get());
    multiset_ref_tile_kernel<true><<<8, 128>>>(
      set.ref(cuco::for_each, cuco::count), iter, unique, multiplicity, errors.data().get());
    multiset_ref_tile_kernel<false><<<8, 128>>>(
      set.ref(cuco::for_each, cuco::count), iter + unique, unique, std::size_t{0}, errors.data().get());
  }
  cudaDeviceSynchronize();
  CHECK(errors[0] == 0);
  {
    using counter_type = cuda::atomic<typename set_type::size_type, cuda::thread_scope_device>;
    thrust::device_vector<Key> probed(n + unique), matched(n + unique);
    counter_type* counter{};
    cudaMalloc(&counter, sizeof(counter_type));
    for (bool outer : {false, true}) {
      cudaMemset(counter, 0, sizeof(counter_type));
      multiset_ref_retrieve_kernel<128><<<1, 128>>>(set.ref(cuco::retrieve),
                                                   iter,
                                                   2 * unique,
                                                   probed.data().get(),
                                                   matched.data().get(),
                                                   counter,
                                                   outer);
      typename set_type::size_type rows{};
      cudaMemcpy(&rows, counter, sizeof(rows), cudaMemcpyDeviceToHost);
      CHECK(rows == (outer ? n + unique : n));
      thrust::sort_by_key(probed.begin(), probed.begin() + rows, matched.begin());
      CHECK(thrust::equal(probed.begin(), probed.begin() + n, matched.begin()));
      auto const expected =
        thrust::make_transform_iterator(iter, divide_by<Key>{static_cast<Key>(multiplicity)});
      CHECK(thrust::equal(probed.begin(), probed.begin() + n, expected));
    }
    cudaFree(counter);
  }
}

// ------------------------------------------------------------------------------------------------
// experimental::static_multimap (tests/static_multimap/{insert_contains,insert_if,count,for_each}_test.cu)
// ------------------------------------------------------------------------------------------------
template <typename Key, typename Value>
struct make_kv_mod {
  Key distinct;
  __host__ __device__ cuco::pair<Key, Value> operator()(std::int64_t i) const
  {
    return cuco::pair<Key, Value>{static_cast<Key>(i % distinct), static_cast<Value>(i)};
  }
};

template <typename Ref, typename Key>
__global__ void multimap_ref_kernel(Ref ref, Key n_keys, std::size_t multiplicity, Key distinct, int* errors)
{
  constexpr int cgs = Ref::cg_size;
  auto const tile   = cg::tiled_partition<cgs>(cg::this_thread_block());
  for (Key k = (blockIdx.x * blockDim.x + threadIdx.x) / cgs; k < n_keys; k += (gridDim.x * blockDim.x) / cgs) {
    std::size_t mine = 0, bad = 0;
    auto const on_slot = [&](auto const slot) {
      ++mine;
      // payload i was stored under key i % distinct
      if (slot.first != k || static_cast<Key>(slot.second % distinct) != k) { ++bad; }
    };
    std::size_t count = 0;
    bool present      = false;
    if constexpr (cgs == 1) {
      ref.for_each(k, on_slot);
      count   = ref.count(k);
      present = ref.contains(k);
    } else {
      ref.for_each(tile, k, on_slot);
      mine    = cg::reduce(tile, mine, cg::plus<std::size_t>());
      bad     = cg::reduce(tile, bad, cg::plus<std::size_t>());
      count   = cg::reduce(tile, static_cast<std::size_t>(ref.count(tile, k)), cg::plus<std::size_t>());
      present = ref.contains(tile, k);
    }
    std::size_t const want = k < distinct ? multiplicity : 0;
    if (tile.thread_rank() == 0) {
      if (mine != want || bad != 0) { atomicAdd(errors, 1); }
      if (count != want) { atomicAdd(errors, 1); }
      if (present != (want != 0)) { atomicAdd(errors, 1); }
    }
  }
}

template <typename Key, typename Value, typename Probe>
static void multimap_suite(char const* label)
{
  std::printf("# experimental::static_multimap, %s\n", label);
  using extent_type = cuco::extent<std::size_t>;
  using map_type    = cuco::experimental::static_multimap<Key,
                                                       Value,
                                                       extent_type,
                                                       cuda::thread_scope_device,
                                                       thrust::equal_to<Key>,
                                                       Probe,
                                                       cuco::cuda_allocator<cuda::std::byte>,
                                                       cuco::storage<2>>;
  constexpr std::size_t n = 4000, distinct = 500, multiplicity = n / distinct;
  auto map = map_type{extent_type{2 * n}, cuco::empty_key<Key>{-1}, cuco::empty_value<Value>{-1}};
  auto const keys  = thrust::counting_iterator<Key>{0};
  auto const pairs = thrust::make_transform_iterator(
    thrust::counting_iterator<std::int64_t>{0}, make_kv_mod<Key, Value>{static_cast<Key>(n)});
  thrust::device_vector<bool> present(n);

  map.contains(keys, keys + n, present.begin());
  CHECK(none_true(present.begin(), present.end()));
  CHECK(map.count(keys, keys + n) == 0);

  // unique keys
  map.insert(pairs, pairs + n);
  map.contains(keys, keys + n, present.begin());
  CHECK(all_true(present.begin(), present.end()));
  CHECK(map.count(keys, keys + n) == n);
  map.contains_if(keys, keys + n, thrust::counting_iterator<std::size_t>(0), is_even{}, present.begin());
  CHECK(thrust::count(present.begin(), present.end(), true) == static_cast<std::ptrdiff_t>(n / 2));

  // insert_if: half of the stream, then the same element n / 2 times
  map.clear();
  CHECK(map.insert_if(pairs, pairs + n, keys, is_even{}) == n / 2);
  CHECK(map.count(keys, keys + n) == n / 2);
  map.clear();
  auto const same = thrust::constant_iterator<cuco::pair<Key, Value>>{{1, 1}};
  CHECK(map.insert_if(same, same + n, keys, is_even{}) == n / 2);
  CHECK(map.count(keys, keys + n) == n / 2);

  // every key `multiplicity` times; device ref: for_each / count / contains
  map.clear();
  auto const dup_pairs = thrust::make_transform_iterator(
    thrust::counting_iterator<std::int64_t>{0}, make_kv_mod<Key, Value>{static_cast<Key>(distinct)});
  map.insert_async(dup_pairs, dup_pairs + n);
  CHECK(map.count(keys, keys + n) == n);
  CHECK(map.count(keys, keys + distinct / 2) == n / 2);
  thrust::device_vector<int> errors(1, 0);
  multimap_ref_kernel<<<8, 128>>>(map.ref(cuco::for_each, cuco::count, cuco::contains),
                                  static_cast<Key>(2 * distinct),
                                  multiplicity,
                                  static_cast<Key>(distinct),
                                  errors.data().get());
  cudaDeviceSynchronize();
  CHECK(errors[0] == 0);
}

int main()
{
  bulk_api_suite<std::int64_t, std::int64_t, cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>>, 1>(
    "int64/int64 linear_probing<1> storage<1>");
  bulk_api_suite<std::int64_t, std::int64_t, cuco::double_hashing<8, cuco::default_hash_function<std::int64_t>>, 1>(
    "int64/int64 double_hashing<8> storage<1>");
  bulk_api_suite<std::int32_t, std::int32_t, cuco::linear_probing<4, cuco::default_hash_function<std::int32_t>>, 1>(
    "int32/int32 linear_probing<4> storage<1> (defaults)");
  bulk_api_suite<std::int32_t, std::int32_t, cuco::double_hashing<2, cuco::murmurhash3_32<std::int32_t>>, 2>(
    "int32/int32 double_hashing<2, murmur> storage<2>");
  bulk_api_suite<std::int32_t, std::int64_t, cuco::linear_probing<2, cuco::xxhash_64<std::int32_t>>, 2>(
    "int32/int64 linear_probing<2, xxhash_64> storage<2> (padded slots)");
  bulk_api_suite<std::int64_t, std::int64_t,
                 cuco::double_hashing<4, cuco::murmurhash3_x64_128<std::int64_t>, cuco::murmurhash3_x64_128<std::int64_t>>, 1>(
    "int64/int64 double_hashing<4, murmur x64_128> (128-bit hash)");
  aggregate_suite();
  device_ref_suite<1, 1>("double_hashing<1> storage<1>");
  device_ref_suite<2, 2>("double_hashing<2> storage<2>");
  device_ref_suite<8, 1>("double_hashing<8> storage<1>");
  count_by_key_suite();
  shared_memory_suite();
  heterogeneous_suite();
  key_sentinel_suite();
  duplicate_and_set_suite();
  probing_suite();
  set_retrieve_suite<std::int32_t, cuco::double_hashing<2, cuco::default_hash_function<std::int32_t>>>(
    "int32 double_hashing<2>");
  set_retrieve_suite<std::int64_t, cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>>>(
    "int64 linear_probing<1>");
  multiset_suite<std::int32_t, cuco::double_hashing<4, cuco::default_hash_function<std::int32_t>>>(
    "int32 double_hashing<4> storage<2> (class defaults)");
  multiset_suite<std::int64_t, cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>>>(
    "int64 linear_probing<1> storage<2>");
  multiset_suite<std::int64_t, cuco::double_hashing<2, cuco::default_hash_function<std::int64_t>>>(
    "int64 double_hashing<2> storage<2>");
  multiset_suite<std::int32_t, cuco::linear_probing<1, cuco::default_hash_function<std::int32_t>>>(
    "int32 linear_probing<1> storage<2>");
  multimap_suite<std::int32_t, std::int32_t, cuco::double_hashing<2, cuco::murmurhash3_32<std::int32_t>, cuco::murmurhash3_32<std::int32_t>>>(
    "int32/int32 double_hashing<2, murmur> storage<2>");
  multimap_suite<std::int64_t, std::int64_t, cuco::linear_probing<1, cuco::murmurhash3_32<std::int64_t>>>(
    "int64/int64 linear_probing<1, murmur> storage<2>");
  multimap_suite<std::int32_t, std::int64_t, cuco::linear_probing<2, cuco::murmurhash3_32<std::int32_t>>>(
    "int32/int64 linear_probing<2, murmur> storage<2> (padded slots)");
  multimap_suite<std::int64_t, std::int32_t, cuco::double_hashing<1, cuco::murmurhash3_32<std::int64_t>, cuco::murmurhash3_32<std::int64_t>>>(
    "int64/int32 double_hashing<1, murmur> storage<2> (padded slots)");
  cudaDeviceSynchronize();
  bool const cuda_ok = cudaGetLastError() == cudaSuccess;
  report(cuda_ok, "no CUDA error at exit");
  std::printf("{\"total\": %d, \"failed\": %d}\n", g_total, g_failed);
  return g_failed ? 1 : 0;
}
