// Host baseline for bench.py's `cpu_baseline` / `--impl reference` fallback legs.
// TEST/BENCH INFRASTRUCTURE ONLY - never linked into the product.
//
// cuCollections has no CPU path, so BASELINE.json's north_star asks for a plain host baseline
// "written for the benchmark and reported only": std::unordered_map / std::unordered_set with
// reserve(n / load_factor), keys hash-sharded over T threads, one private container per thread and
// no locks (SURVEY.md §8d "CPU baseline"). Every thread scans the whole input and keeps the keys it
// owns; timing is steady_clock around the parallel region.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

inline std::uint64_t mix(std::uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

template <typename F>
double run_parallel(int threads, F&& body)
{
  std::vector<std::thread> pool;
  auto const t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t] { body(t); });
  }
  for (auto& th : pool) {
    th.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

extern "C" {

/// insert n (key,value) pairs, then look up nq queries; writes seconds for each phase and a checksum
/// (sum of found values + number of distinct keys) so the work cannot be optimised away.
int cpu_baseline_map_i64(const std::int64_t* keys,
                         const std::int64_t* values,
                         std::int64_t n,
                         const std::int64_t* queries,
                         std::int64_t nq,
                         int threads,
                         double load_factor,
                         double* insert_seconds,
                         double* find_seconds,
                         std::int64_t* checksum)
{
  if (threads < 1) { threads = 1; }
  std::vector<std::unordered_map<std::int64_t, std::int64_t>> shards(threads);
  for (auto& s : shards) {
    s.max_load_factor(static_cast<float>(load_factor));
    s.reserve(static_cast<std::size_t>(n / threads + 1));
  }
  *insert_seconds = run_parallel(threads, [&](int t) {
    auto& m = shards[t];
    for (std::int64_t i = 0; i < n; ++i) {
      if (mix(static_cast<std::uint64_t>(keys[i])) % threads == static_cast<std::uint64_t>(t)) {
        m.emplace(keys[i], values ? values[i] : 0);
      }
    }
  });
  std::atomic<std::int64_t> sum{0};
  *find_seconds = run_parallel(threads, [&](int t) {
    auto const& m      = shards[t];
    std::int64_t local = 0;
    for (std::int64_t i = 0; i < nq; ++i) {
      if (mix(static_cast<std::uint64_t>(queries[i])) % threads == static_cast<std::uint64_t>(t)) {
        auto const it = m.find(queries[i]);
        local += it == m.end() ? -1 : it->second;
      }
    }
    sum += local;
  });
  std::int64_t distinct = 0;
  for (auto const& s : shards) {
    distinct += static_cast<std::int64_t>(s.size());
  }
  *checksum = sum.load() + distinct;
  return 0;
}

/// static_set<int32> flavour: insert n keys, then contains on nq queries.
int cpu_baseline_set_i32(const std::int32_t* keys,
                         std::int64_t n,
                         const std::int32_t* queries,
                         std::int64_t nq,
                         int threads,
                         double load_factor,
                         double* insert_seconds,
                         double* find_seconds,
                         std::int64_t* checksum)
{
  if (threads < 1) { threads = 1; }
  std::vector<std::unordered_set<std::int32_t>> shards(threads);
  for (auto& s : shards) {
    s.max_load_factor(static_cast<float>(load_factor));
    s.reserve(static_cast<std::size_t>(n / threads + 1));
  }
  *insert_seconds = run_parallel(threads, [&](int t) {
    auto& m = shards[t];
    for (std::int64_t i = 0; i < n; ++i) {
      if (mix(static_cast<std::uint32_t>(keys[i])) % threads == static_cast<std::uint64_t>(t)) {
        m.insert(keys[i]);
      }
    }
  });
  std::atomic<std::int64_t> hits{0};
  *find_seconds = run_parallel(threads, [&](int t) {
    auto const& m      = shards[t];
    std::int64_t local = 0;
    for (std::int64_t i = 0; i < nq; ++i) {
      if (mix(static_cast<std::uint32_t>(queries[i])) % threads == static_cast<std::uint64_t>(t)) {
        local += m.count(queries[i]);
      }
    }
    hits += local;
  });
  std::int64_t distinct = 0;
  for (auto const& s : shards) {
    distinct += static_cast<std::int64_t>(s.size());
  }
  *checksum = hits.load() + distinct;
  return 0;
}

int cpu_baseline_hardware_threads(void)
{
  auto const n = std::thread::hardware_concurrency();
  return n ? static_cast<int>(n) : 1;
}

}  // extern "C"
