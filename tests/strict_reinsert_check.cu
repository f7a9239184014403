// Opt-in insert rule, built with -DCUCO_B200_TOMBSTONE_AWARE_INSERT=1 (include/cuco/b200/probe_engine.cuh):
// an erased slot is only a candidate while the cluster is searched for the key, so re-inserting a key
// that is still present never stores it twice, and a table with no empty slot left (only tombstones)
// still accepts a new key after at most one full cycle. The DEFAULT build follows cuco's rule instead
// and is pinned against cuco in tests/test_baseline_configs_gpu.py::test_insert_after_erase_follows_cuco.
#include <cuco/static_map.cuh>
#include <cuco/static_set.cuh>

#include <thrust/device_vector.h>

#include <cstdint>
#include <cstdio>
#include <vector>

static int failures = 0;
static void report(bool ok, char const* what)
{
  std::printf("%s %s\n", ok ? "PASS" : "FAIL", what);
  if (!ok) { ++failures; }
}
#define CHECK(expr) report((expr), #expr)

template <typename Probe>
static void reinsert_never_duplicates(char const* label)
{
  std::printf("-- %s\n", label);
  using K = std::int64_t;
  cuco::static_map<K, K, cuco::extent<std::size_t>, cuda::thread_scope_device, thrust::equal_to<K>, Probe> map{
    64, cuco::empty_key<K>{-1}, cuco::empty_value<K>{-1}, cuco::erased_key<K>{-2}};
  std::vector<cuco::pair<K, K>> host;
  for (K k = 1; k <= 40; ++k) { host.push_back({k, 2 * k}); }
  thrust::device_vector<cuco::pair<K, K>> pairs(host.begin(), host.end());
  std::size_t fresh = 0;
  for (int i = 0; i < 40; ++i) { fresh += map.insert(pairs.begin() + i, pairs.begin() + i + 1); }
  CHECK(fresh == 40);
  std::vector<K> gone, kept;
  for (K k = 1; k <= 40; ++k) { (k % 2 ? gone : kept).push_back(k); }
  thrust::device_vector<K> d_gone(gone.begin(), gone.end());
  map.erase(d_gone.begin(), d_gone.end());
  CHECK(map.size() == 20);
  std::size_t again = 0;
  for (int i = 1; i < 40; i += 2) { again += map.insert(pairs.begin() + i, pairs.begin() + i + 1); }
  CHECK(again == 0);           // every kept key was found behind the tombstones
  CHECK(map.size() == 20);     // ... and none was stored a second time
  CHECK(map.insert(pairs.begin(), pairs.end()) == 20);  // the erased half comes back, bulk
  CHECK(map.size() == 40);
}

static void full_table_of_tombstones_accepts_a_key()
{
  std::printf("-- no empty slot left\n");
  using K = std::int32_t;
  cuco::static_set<K, cuco::extent<std::size_t>, cuda::thread_scope_device, thrust::equal_to<K>,
                   cuco::linear_probing<1, cuco::default_hash_function<K>>>
    set{7, cuco::empty_key<K>{-1}, cuco::erased_key<K>{-2}};
  auto const cap = static_cast<int>(set.capacity());
  std::vector<K> host;
  for (K k = 0; k < cap; ++k) { host.push_back(100 + k); }
  thrust::device_vector<K> keys(host.begin(), host.end());
  for (int i = 0; i < cap; ++i) { set.insert(keys.begin() + i, keys.begin() + i + 1); }
  CHECK(static_cast<int>(set.size()) == cap);  // load factor 1.0
  set.erase(keys.begin() + 2, keys.begin() + 3);
  thrust::device_vector<K> fresh(1, K{9999});
  CHECK(set.insert(fresh.begin(), fresh.end()) == 1);  // terminates, takes the tombstone
  CHECK(static_cast<int>(set.size()) == cap);
  thrust::device_vector<bool> present(1);
  set.contains(fresh.begin(), fresh.end(), present.begin());
  CHECK(present[0]);
}

int main()
{
  reinsert_never_duplicates<cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>>>("linear_probing<1>");
  reinsert_never_duplicates<cuco::double_hashing<1, cuco::default_hash_function<std::int64_t>>>("double_hashing<1>");
  full_table_of_tombstones_accepts_a_key();
  CHECK(cudaDeviceSynchronize() == cudaSuccess);
  std::printf("%d failures\n", failures);
  return failures == 0 ? 0 : 1;
}
