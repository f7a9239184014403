// Development harness for the blocked insert path (measurement tool, not product code): one
// static_map<int64,int64> with linear_probing<1>, the C2 workload, timed with CUDA events through the
// public cuco:: API, variants selected through the CUCO_B200_* environment (this file is compiled with
// CUCO_B200_TUNABLE). Prints one JSON line per configuration and checks every payload with find.
//   insert_lab [n = 100000000] [load_factor = 0.5] [reps = 5] [unique = 0]
#define CUCO_B200_TUNABLE 1
#include <cuco/static_map.cuh>

#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/count.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/transform.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

using Key   = std::int64_t;
using Value = std::int64_t;

__host__ __device__ inline std::uint64_t mix64(std::uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

struct make_pair_uniform {
  std::uint64_t n;
  bool unique;
  __device__ cuco::pair<Key, Value> operator()(std::int64_t i) const
  {
    // uniform_int[1, n] (~63 % distinct) or a bijective scramble of 1..n
    std::uint64_t const k = unique ? (static_cast<std::uint64_t>(i) * 0x9e3779b97f4a7c15ull) >> 1
                                   : 1 + mix64(static_cast<std::uint64_t>(i) + 42) % n;
    return {static_cast<Key>(k | (unique ? 1 : 0)), static_cast<Value>(k | (unique ? 1 : 0))};
  }
};

struct first_of {
  __device__ Key operator()(cuco::pair<Key, Value> const& p) const { return p.first; }
};

struct mismatch {
  __device__ bool operator()(thrust::tuple<Key, Value> const& t) const { return thrust::get<0>(t) != thrust::get<1>(t); }
};

int main(int argc, char** argv)
{
  std::int64_t const n = argc > 1 ? std::atoll(argv[1]) : 100'000'000;
  double const lf      = argc > 2 ? std::atof(argv[2]) : 0.5;
  int const reps       = argc > 3 ? std::atoi(argv[3]) : 5;
  bool const unique    = argc > 4 ? std::atoi(argv[4]) != 0 : false;

  thrust::device_vector<cuco::pair<Key, Value>> pairs(n);
  thrust::transform(thrust::device,
                    thrust::counting_iterator<std::int64_t>{0},
                    thrust::counting_iterator<std::int64_t>{n},
                    pairs.begin(),
                    make_pair_uniform{static_cast<std::uint64_t>(n), unique});
  thrust::device_vector<Key> keys(n);
  thrust::transform(thrust::device, pairs.begin(), pairs.end(), keys.begin(), first_of{});
  thrust::device_vector<Value> found(n);

  using map_t = cuco::static_map<Key,
                                 Value,
                                 cuco::extent<std::size_t>,
                                 cuda::thread_scope_device,
                                 thrust::equal_to<Key>,
                                 cuco::linear_probing<1, cuco::default_hash_function<Key>>>;
  map_t map{cuco::extent<std::size_t>{static_cast<std::size_t>(n)}, lf, cuco::empty_key<Key>{-1}, cuco::empty_value<Value>{-1}};

  cudaEvent_t a, b, c;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventCreate(&c);
  std::vector<float> ins, fnd;
  for (int r = 0; r < reps + 1; ++r) {
    map.clear_async();
    cudaEventRecord(a);
    map.insert_async(pairs.begin(), pairs.end());
    cudaEventRecord(b);
    map.find_async(keys.begin(), keys.end(), found.begin());
    cudaEventRecord(c);
    cudaEventSynchronize(c);
    float ti, tf;
    cudaEventElapsedTime(&ti, a, b);
    cudaEventElapsedTime(&tf, b, c);
    if (r > 0) {
      ins.push_back(ti);
      fnd.push_back(tf);
    }
  }
  auto const err = cudaGetLastError();
  auto const wrong = thrust::count_if(
    thrust::device,
    thrust::make_zip_iterator(thrust::make_tuple(keys.begin(), found.begin())),
    thrust::make_zip_iterator(thrust::make_tuple(keys.end(), found.end())),
    mismatch{});
  auto const size = map.size();
  std::sort(ins.begin(), ins.end());
  std::sort(fnd.begin(), fnd.end());
  auto const& t = cuco::b200::tuning();
  std::printf(
    "{\"test\": \"insert_lab\", \"n\": %lld, \"lf\": %.2f, \"unique\": %d, \"blocked\": %d, \"tile_route\": %d, "
    "\"stream_probe\": %d, \"slots\": %d, \"scout\": %d, \"kpt\": %d, \"prefetch\": %d, \"region_mib\": %zu, "
    "\"insert_ms_median\": %.3f, \"insert_ms_best\": %.3f, \"insert_gops\": %.2f, \"find_ms_median\": %.3f, "
    "\"size\": %llu, \"wrong\": %lld, \"cuda\": \"%s\"}\n",
    (long long)n, lf, (int)unique, t.blocked, (int)t.blocked_tile_route, (int)t.blocked_stream_probe, t.stream_slots, (int)t.stream_scout,
    t.blocked_keys_per_thread, (int)t.blocked_prefetch, t.region_bytes >> 20, ins[ins.size() / 2], ins[0],
    n / ins[ins.size() / 2] / 1e6, fnd[fnd.size() / 2], (unsigned long long)size, (long long)wrong,
    cudaGetErrorString(err));
  return wrong == 0 && err == cudaSuccess ? 0 : 1;
}
