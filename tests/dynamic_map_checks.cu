// Drop-in check of cuco::experimental::dynamic_map: one source, compiled against include/ (this
// implementation) and against /root/reference/include (cuCollections), run on the GPU. It follows
// tests/dynamic_map/unique_sequence_test_experimental.cu and adds the growth case (a stream much
// larger than the first submap, queried in insertion order - the only order for which the
// reference's positional `contains` is defined). Prints one line per check and a JSON summary.
#include <cuco/dynamic_map.cuh>

#include <thrust/count.h>
#include <thrust/device_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/sequence.h>

#include <cstdint>
#include <cstdio>

static int g_failed = 0, g_total = 0;
static void report(bool ok, char const* what)
{
  ++g_total;
  if (!ok) { ++g_failed; }
  std::printf("%s %s\n", ok ? "PASS" : "FAIL", what);
}
#define CHECK(expr) report((expr), #expr)

template <typename Key, typename T>
struct make_pair_of {
  __host__ __device__ cuco::pair<Key, T> operator()(std::int64_t i) const
  {
    return cuco::pair<Key, T>{static_cast<Key>(i), static_cast<T>(i)};
  }
};

template <typename Key, typename T>
static void suite(char const* label, std::size_t initial_capacity, std::size_t num_keys)
{
  std::printf("# experimental::dynamic_map %s, initial capacity %zu, %zu keys\n", label, initial_capacity, num_keys);
  thrust::device_vector<Key> keys(num_keys);
  thrust::sequence(keys.begin(), keys.end());
  auto const pairs = thrust::make_transform_iterator(thrust::counting_iterator<std::int64_t>{0},
                                                     make_pair_of<Key, T>{});
  thrust::device_vector<bool> present(num_keys);
  {
    cuco::experimental::dynamic_map<Key, T> map{
      initial_capacity, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
    map.contains(keys.begin(), keys.end(), present.begin());
    CHECK(thrust::count(present.begin(), present.end(), true) == 0);
  }
  {
    cuco::experimental::dynamic_map<Key, T> map{
      initial_capacity, cuco::empty_key<Key>{-1}, cuco::empty_value<T>{-1}};
    map.insert(pairs, pairs + num_keys);
    map.contains(keys.begin(), keys.end(), present.begin());
    CHECK(thrust::count(present.begin(), present.end(), true) == static_cast<std::ptrdiff_t>(num_keys));
    // a second batch of new keys on top (the chain keeps growing), still queried in insertion order
    map.insert(pairs + num_keys, pairs + 2 * num_keys);
    thrust::device_vector<Key> all(2 * num_keys);
    thrust::sequence(all.begin(), all.end());
    thrust::device_vector<bool> both(2 * num_keys);
    map.contains(all.begin(), all.end(), both.begin());
    CHECK(thrust::count(both.begin(), both.end(), true) == static_cast<std::ptrdiff_t>(2 * num_keys));
  }
}

int main()
{
  suite<std::int32_t, std::int32_t>("int32/int32", 3'000'000, 1'000'000);  // one submap is enough
  suite<std::int64_t, std::int64_t>("int64/int64", 100'000, 1'000'000);    // grows through several submaps
  suite<std::int32_t, std::int64_t>("int32/int64", 50'000, 200'000);
  cudaDeviceSynchronize();
  report(cudaGetLastError() == cudaSuccess, "no CUDA error at exit");
  std::printf("{\"total\": %d, \"failed\": %d}\n", g_total, g_failed);
  return g_failed ? 1 : 0;
}
