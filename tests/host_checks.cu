// Host-side known-answer checks of the drop-in headers (no GPU needed: everything here is the
// __host__ half of __host__ __device__ code). Built by __graft_entry__.build() into
// tests/_build/host_checks; tests/test_host_checks.py runs it and asserts on the JSON it prints.
//
// Vectors are the ones the reference's own tests pin:
//   tests/utility/hash_test.cu:57-67,102-121,166-184,251-323   (identity, xxhash_64, xxhash_32, murmur3 128)
//   tests/utility/extent_test.cu:26-55                          (1234 -> 314 windows for cg 2, window 4)
//   tests/utility/fast_int_test.cu:27-61                        (div/mod identities)
//   tests/static_map/capacity_test.cu, static_set/capacity_test.cu (rounding golds via valid_num_windows)
#include <cuco/b200/bulk_kernels.cuh>
#include <cuco/extent.cuh>
#include <cuco/hash_functions.cuh>
#include <cuco/pair.cuh>
#include <cuco/probing_scheme.cuh>
#include <cuco/utility/fast_int.cuh>

#include <cuda/std/array>

#include <cstdint>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

namespace {

int g_failed = 0;
int g_total  = 0;
std::vector<std::string> g_failures;

void check(bool ok, char const* what)
{
  ++g_total;
  if (!ok) {
    ++g_failed;
    g_failures.emplace_back(what);
  }
}
#define CHECK(expr) check((expr), #expr)

template <int Words>
struct large_key {
  constexpr large_key(std::int32_t value) noexcept
  {
    for (int i = 0; i < Words; ++i) {
      data_[i] = value;
    }
  }
  std::int32_t data_[Words];
};

template <typename Hash, typename... Seed>
bool hashes_to(typename Hash::argument_type const& key, typename Hash::result_type expected, Seed... seed)
{
  return Hash{seed...}(key) == expected;
}

void hash_vectors()
{
  using cuco::identity_hash;
  CHECK((hashes_to<identity_hash<signed char>>(0, 0)));
  CHECK((hashes_to<identity_hash<std::int32_t>>(std::numeric_limits<std::int32_t>::max(),
                                                std::numeric_limits<std::int32_t>::max())));
  CHECK((hashes_to<identity_hash<std::int64_t>>(std::numeric_limits<std::int64_t>::max(),
                                                std::numeric_limits<std::int64_t>::max())));

  using cuco::xxhash_64;
  CHECK((hashes_to<xxhash_64<char>>(0, 16804241149081757544ull, 0)));
  CHECK((hashes_to<xxhash_64<char>>(42, 765293966243412708ull, 0)));
  CHECK((hashes_to<xxhash_64<char>>(0, 9486749600008296231ull, 42)));
  CHECK((hashes_to<xxhash_64<std::int32_t>>(0, 4246796580750024372ull, 0)));
  CHECK((hashes_to<xxhash_64<std::int32_t>>(0, 3614696996920510707ull, 42)));
  CHECK((hashes_to<xxhash_64<std::int32_t>>(42, 15516826743637085169ull, 0)));
  CHECK((hashes_to<xxhash_64<std::int32_t>>(123456789, 9462334144942111946ull, 0)));
  CHECK((hashes_to<xxhash_64<std::int64_t>>(0, 3803688792395291579ull, 0)));
  CHECK((hashes_to<xxhash_64<std::int64_t>>(0, 13194218611613725804ull, 42)));
  CHECK((hashes_to<xxhash_64<std::int64_t>>(42, 13066772586158965587ull, 0)));
  CHECK((hashes_to<xxhash_64<std::int64_t>>(123456789, 14662639848940634189ull, 0)));
  CHECK((hashes_to<xxhash_64<__int128>>(123456789, 7986913354431084250ull, 0)));
  CHECK((hashes_to<xxhash_64<large_key<32>>>(123456789, 2031761887105658523ull, 0)));

  using cuco::xxhash_32;
  CHECK((hashes_to<xxhash_32<char>>(0, 3479547966u, 0)));
  CHECK((hashes_to<xxhash_32<char>>(42, 3774771295u, 0)));
  CHECK((hashes_to<xxhash_32<char>>(0, 2099223482u, 42)));
  CHECK((hashes_to<xxhash_32<std::int32_t>>(0, 148298089u, 0)));
  CHECK((hashes_to<xxhash_32<std::int32_t>>(0, 2132181312u, 42)));
  CHECK((hashes_to<xxhash_32<std::int32_t>>(42, 1161967057u, 0)));
  CHECK((hashes_to<xxhash_32<std::int32_t>>(123456789, 2987034094u, 0)));
  CHECK((hashes_to<xxhash_32<std::int64_t>>(0, 3736311059u, 0)));
  CHECK((hashes_to<xxhash_32<std::int64_t>>(0, 1076387279u, 42)));
  CHECK((hashes_to<xxhash_32<std::int64_t>>(42, 2332451213u, 0)));
  CHECK((hashes_to<xxhash_32<std::int64_t>>(123456789, 1561711919u, 0)));
  CHECK((hashes_to<xxhash_32<__int128>>(123456789, 1846633701u, 0)));
  CHECK((hashes_to<xxhash_32<large_key<32>>>(123456789, 3715432378u, 0)));

  // static vs dynamic key size give the same value (hash_test.cu:224-246)
  {
    std::int32_t key = 42;
    CHECK(cuco::murmurhash3_32<std::int32_t>{}(key) ==
          cuco::murmurhash3_32<std::int32_t>{}.compute_hash(
            reinterpret_cast<cuda::std::byte const*>(&key), sizeof(key)));
    CHECK(cuco::xxhash_32<std::int32_t>{}(key) ==
          cuco::xxhash_32<std::int32_t>{}.compute_hash(reinterpret_cast<std::byte const*>(&key),
                                                       sizeof(key)));
    CHECK(cuco::xxhash_64<std::int32_t>{}(key) ==
          cuco::xxhash_64<std::int32_t>{}.compute_hash(reinterpret_cast<std::byte const*>(&key),
                                                       sizeof(key)));
    char c = 42;
    CHECK(cuco::murmurhash3_32<char>{}(c) ==
          cuco::murmurhash3_32<char>{}.compute_hash(reinterpret_cast<std::byte const*>(&c), 1));
  }

  using x64  = cuda::std::array<std::uint64_t, 2>;
  using x86  = cuda::std::array<std::uint32_t, 4>;
  using a32_2 = cuda::std::array<std::int32_t, 2>;
  using a32_3 = cuda::std::array<std::int32_t, 3>;
  using a32_4 = cuda::std::array<std::int32_t, 4>;
  using a32_16 = cuda::std::array<std::int32_t, 16>;
  using a64_2 = cuda::std::array<std::int64_t, 2>;
  using a64_3 = cuda::std::array<std::int64_t, 3>;
  using a64_4 = cuda::std::array<std::int64_t, 4>;
  using a64_16 = cuda::std::array<std::int64_t, 16>;
  using cuco::murmurhash3_x64_128;
  using cuco::murmurhash3_x86_128;
  a32_16 const seq32{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  a64_16 const seq64{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

  CHECK((hashes_to<murmurhash3_x64_128<std::int32_t>>(0, x64{14961230494313510588ull, 6383328099726337777ull}, 0)));
  CHECK((hashes_to<murmurhash3_x64_128<std::int32_t>>(9, x64{1779292183511753683ull, 16298496441448380334ull}, 0)));
  CHECK((hashes_to<murmurhash3_x64_128<std::int32_t>>(42, x64{2913627637088662735ull, 16344193523890567190ull}, 0)));
  CHECK((hashes_to<murmurhash3_x64_128<std::int32_t>>(42, x64{2248879576374326886ull, 18006515275339376488ull}, 42)));
  CHECK((hashes_to<murmurhash3_x64_128<a32_2>>(a32_2{2, 2}, x64{12221386834995143465ull, 6690950894782946573ull}, 0)));
  CHECK((hashes_to<murmurhash3_x64_128<a32_3>>(a32_3{1, 4, 9}, x64{299140022350411792ull, 9891903873182035274ull}, 42)));
  CHECK((hashes_to<murmurhash3_x64_128<a32_4>>(a32_4{42, 64, 108, 1024}, x64{4333511168876981289ull, 4659486988434316416ull}, 63)));
  CHECK((hashes_to<murmurhash3_x64_128<a32_16>>(seq32, x64{3302412811061286680ull, 7070355726356610672ull}, 1024)));
  CHECK((hashes_to<murmurhash3_x64_128<a64_2>>(a64_2{2, 2}, x64{8554944597931919519ull, 14938998000509429729ull}, 0)));
  CHECK((hashes_to<murmurhash3_x64_128<a64_3>>(a64_3{1, 4, 9}, x64{13442629947720186435ull, 7061727494178573325ull}, 42)));
  CHECK((hashes_to<murmurhash3_x64_128<a64_4>>(a64_4{42, 64, 108, 1024}, x64{8786399719555989948ull, 14954183901757012458ull}, 63)));
  CHECK((hashes_to<murmurhash3_x64_128<a64_16>>(seq64, x64{15409921801541329777ull, 10546487400963404004ull}, 1024)));

  CHECK((hashes_to<murmurhash3_x86_128<std::int32_t>>(0, x86{3422973727u, 2656139328u, 2656139328u, 2656139328u}, 0)));
  CHECK((hashes_to<murmurhash3_x86_128<std::int32_t>>(9, x86{2808089785u, 314604614u, 314604614u, 314604614u}, 0)));
  CHECK((hashes_to<murmurhash3_x86_128<std::int32_t>>(42, x86{3611919118u, 1962256489u, 1962256489u, 1962256489u}, 0)));
  CHECK((hashes_to<murmurhash3_x86_128<std::int32_t>>(42, x86{3399017053u, 732469929u, 732469929u, 732469929u}, 42)));
  CHECK((hashes_to<murmurhash3_x86_128<a32_2>>(a32_2{2, 2}, x86{1234494082u, 1431451587u, 431049201u, 431049201u}, 0)));
  CHECK((hashes_to<murmurhash3_x86_128<a32_3>>(a32_3{1, 4, 9}, x86{2516796247u, 2757675829u, 778406919u, 2453259553u}, 42)));
  CHECK((hashes_to<murmurhash3_x86_128<a32_4>>(a32_4{42, 64, 108, 1024}, x86{2686265656u, 591236665u, 3797082165u, 2731908938u}, 63)));
  CHECK((hashes_to<murmurhash3_x86_128<a32_16>>(seq32, x86{3918256832u, 4205523739u, 1707810111u, 1625952473u}, 1024)));
  CHECK((hashes_to<murmurhash3_x86_128<a64_2>>(a64_2{2, 2}, x86{3811075945u, 727160712u, 3510740342u, 235225510u}, 0)));
  CHECK((hashes_to<murmurhash3_x86_128<a64_3>>(a64_3{1, 4, 9}, x86{2817194959u, 206796677u, 3391242768u, 248681098u}, 42)));
  CHECK((hashes_to<murmurhash3_x86_128<a64_4>>(a64_4{42, 64, 108, 1024}, x86{2335912146u, 1566515912u, 760710030u, 452077451u}, 63)));
  CHECK((hashes_to<murmurhash3_x86_128<a64_16>>(seq64, x86{1101169764u, 1758958147u, 2406511780u, 2903571412u}, 1024)));
}

template <typename T>
void fast_int_identities()
{
  T const values[] = {1, 2, 9, 32, 4123, 8192, 4312456};
  constexpr T top  = std::numeric_limits<T>::max();
  for (T d : values) {
    cuco::utility::fast_int fast{d};
    CHECK(static_cast<T>(fast) == d);
    for (T n : values) {
      CHECK(n / fast == n / d);
      CHECK(n % fast == n % d);
    }
    CHECK(top / fast == top / d);
    CHECK(top % fast == top % d);
  }
  cuco::utility::fast_int fast_top{top};
  for (T n : values) {
    CHECK(n / fast_top == n / top);
    CHECK(n % fast_top == n % top);
  }
  // dense sweep with awkward divisors around powers of two and table-sized primes
  std::uint64_t state = 88172645463325252ull;
  auto next           = [&] {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
  };
  T const divisors[] = {3, 5, 7, 127, 128, 129, 65535, 65537, 200039789 % top, 1000081013 % top, top - 1};
  for (T d : divisors) {
    if (d <= 0) { continue; }
    cuco::utility::fast_int fast{d};
    bool ok = true;
    for (int i = 0; i < 20000; ++i) {
      T const n = static_cast<T>(next() & static_cast<std::uint64_t>(top));
      ok        = ok && (n / fast == n / d) && (n % fast == n % d);
    }
    CHECK(ok);
  }
}

void extents()
{
  // static extents are evaluated at compile time (extent_test.cu:35-47)
  constexpr auto s32 = cuco::make_window_extent<2, 4>(cuco::extent<std::int32_t, 1234>{});
  static_assert(s32.value() == 314);
  constexpr auto s64 = cuco::make_window_extent<2, 4>(cuco::extent<std::int64_t, 1234>{});
  static_assert(s64.value() == 314);
  constexpr auto ssz = cuco::make_window_extent<2, 4>(cuco::extent<std::size_t, 1234>{});
  static_assert(ssz.value() == 314);
  static_assert(cuco::extent<std::size_t, 1234>{} == 1234);
  CHECK(cuco::extent(std::size_t{1234}) == 1234u);
  CHECK((cuco::make_window_extent<2, 4>(cuco::extent<std::int32_t>{1234}).value() == 314));
  CHECK((cuco::make_window_extent<2, 4>(cuco::extent<std::int64_t>{1234}).value() == 314));
  CHECK((cuco::make_window_extent<2, 4>(cuco::extent<std::size_t>{1234}).value() == 314));
  CHECK((cuco::make_window_extent<2, 4>(std::size_t{1234}).value() == 314));
  // capacity golds = windows * window_size (capacity_test.cu): cg1 w2 400 -> 211 windows, cg2 w2 -> 206
  CHECK((cuco::make_window_extent<1, 2>(std::size_t{400}).value() * 2 == 422));
  CHECK((cuco::make_window_extent<2, 2>(std::size_t{400}).value() * 2 == 412));
  CHECK((cuco::make_window_extent<1, 2>(std::size_t{500}).value() * 2 == 502));
  CHECK((cuco::make_window_extent<1, 2>(std::size_t{0}).value() * 2 == 4));
  constexpr auto st400 = cuco::make_window_extent<1, 2>(cuco::extent<std::int32_t, 400>{});
  static_assert(st400.value() * 2 == 422);
  constexpr auto st400cg2 = cuco::make_window_extent<2, 2>(cuco::extent<std::int32_t, 400>{});
  static_assert(st400cg2.value() * 2 == 412);
  // too large an extent is an error
  bool threw = false;
  try {
    (void)cuco::make_window_extent<1, 1>(std::size_t{1} << 40);
  } catch (cuco::logic_error const&) {
    threw = true;
  }
  CHECK(threw);
}

void print_primes()
{
  // sample of prime_at_least for the Python side to cross-check against the oracle's rule walk
  std::uint64_t const probes[] = {1, 2, 3, 4, 6, 8, 100, 211, 212, 1000, 65536, 131071, 131072, 131101,
                                  200000, 1000003, 25037357, 200039789, 1000081013, 4000095551ull,
                                  17177758133ull};
  std::printf("\"primes\": {");
  bool first = true;
  for (auto p : probes) {
    std::printf("%s\"%llu\": %llu", first ? "" : ", ", (unsigned long long)p,
                (unsigned long long)cuco::b200::prime_at_least(p));
    first = false;
  }
  std::printf("}, ");
}

void probing_host_sequences()
{
  // scalar sequences on the host: linear walks consecutive windows, double hashing a fixed stride
  auto const bound = cuco::make_window_extent<1, 1>(std::size_t{10});
  cuco::linear_probing<1, cuco::default_hash_function<std::int64_t>> lp{};
  auto it          = lp(std::int64_t{42}, bound);
  auto const start = *it;
  bool ok          = true;
  for (std::size_t i = 0; i < 8; ++i, ++it) {
    ok = ok && (*it == (start + i) % bound.value());
  }
  CHECK(ok);
  cuco::double_hashing<1, cuco::default_hash_function<std::int64_t>> dh{};
  auto jt            = dh(std::int64_t{42}, bound);
  auto const first   = *jt;
  ++jt;
  auto const stride  = (*jt + bound.value() - first) % bound.value();
  CHECK(stride >= 1 && stride < bound.value());
  CHECK(first == cuco::xxhash_32<std::int64_t>{}(42) % bound.value());
  CHECK(stride == cuco::xxhash_32<std::int64_t>{1}(42) % (bound.value() - 1) + 1);
}

void pair_layout()
{
  static_assert(alignof(cuco::pair<std::int64_t, std::int64_t>) == 16);
  static_assert(sizeof(cuco::pair<std::int64_t, std::int64_t>) == 16);
  static_assert(alignof(cuco::pair<std::int32_t, std::int32_t>) == 8);
  static_assert(sizeof(cuco::pair<std::int32_t, std::int64_t>) == 16);
  static_assert(cuco::detail::is_packable<cuco::pair<std::int32_t, std::int32_t>>());
  static_assert(!cuco::detail::is_packable<cuco::pair<std::int64_t, std::int64_t>>());
  static_assert(cuco::is_tuple_like<cuco::pair<int, int>>::value);
  cuco::pair<int, long> p{1, 2};
  auto const q = cuco::make_pair(1, 2L);
  CHECK(p == q);
  cuco::pair<long, long> from_std{std::pair<int, int>{3, 4}};
  CHECK(from_std.first == 3 && from_std.second == 4);
  CHECK(cuda::std::get<0>(p) == 1 && cuda::std::get<1>(p) == 2);
}

}  // namespace

/// Host half of the partitioned / blocked paths: region_map (slot -> region and its exact inverse, also
/// over a table slice) and exchange_owner (key -> owner rank), the two functions every rank must agree on.
void exchange_host_logic()
{
  using cuco::b200::region_map;
  std::uint64_t state = 0x243f6a8885a308d3ull;
  auto next = [&] {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
  };
  for (std::uint64_t capacity : {std::uint64_t{4}, std::uint64_t{1000}, std::uint64_t{200039789}, std::uint64_t{1030100641},
                                 std::uint64_t{4120042477ull}}) {
    for (std::uint32_t regions : {2u, 7u, 99u, 197u, 1024u}) {
      if (capacity < regions) { continue; }  // callers never ask for more regions than slots
      for (std::uint64_t first : {std::uint64_t{0}, std::uint64_t{12345}}) {
        auto const map = region_map::over(capacity, regions, first);
        bool in_range = true, inverse = true, monotone = true;
        std::uint64_t previous = first;
        for (std::uint32_t r = 0; r < regions; ++r) {
          auto const begin = map.region_begin(r);
          monotone         = monotone && begin >= previous;
          previous         = begin;
          if (capacity >= 16ull * regions) {  // every region non-empty: its first slot maps to it, the slot before to r - 1
            inverse = inverse && map(begin) == r && (r == 0 || map(begin - 1) + 1 == r);
          }
        }
        for (int i = 0; i < 4000; ++i) {
          std::uint64_t const slot = first + next() % capacity;
          auto const r             = map(slot);
          in_range                 = in_range && r < regions;
          inverse                  = inverse && map.region_begin(r) <= slot;
          if (r + 1 < regions) { inverse = inverse && slot < map.region_begin(r + 1); }
        }
        check(map.region_begin(0) == first, "region_map: region 0 starts at the slice");
        check(in_range, "region_map: every slot maps below num_regions");
        check(inverse, "region_map: region_begin is the exact inverse of the map");
        check(monotone, "region_map: region starts are monotone");
        check(map(first + capacity - 1) == regions - 1 || capacity < regions, "region_map: the last slot is in the last region");
      }
    }
  }
  for (std::uint32_t ranks : {1u, 2u, 3u, 8u, 16u}) {
    std::vector<std::uint64_t> load(ranks, 0);
    bool below = true;
    constexpr std::uint64_t keys = 1000000;
    for (std::uint64_t k = 0; k < keys; ++k) {
      auto const owner = cuco::b200::exchange_owner(k, 0x9E3779B97F4A7C15ull, ranks);
      below            = below && owner < ranks;
      if (owner < ranks) { ++load[owner]; }
    }
    check(below, "exchange_owner: owner below the number of ranks");
    bool balanced = true;
    for (auto l : load) {
      double const share = static_cast<double>(l) * ranks / keys;
      balanced           = balanced && share > 0.98 && share < 1.02;
    }
    check(balanced, "exchange_owner: sequential keys spread within 2 % over the ranks");
  }
}

void print_owners()
{
  std::printf("\"owners\": {");
  bool first = true;
  for (std::int64_t key : {std::int64_t{0}, std::int64_t{1}, std::int64_t{42}, std::int64_t{-7}, std::int64_t{1} << 40,
                           std::int64_t{123456789012345}}) {
    for (std::uint32_t ranks : {2u, 8u}) {
      std::printf("%s\"%lld/%u\": %u", first ? "" : ", ", static_cast<long long>(key), ranks,
                  cuco::b200::exchange_owner(static_cast<std::uint64_t>(key), 0x9E3779B97F4A7C15ull, ranks));
      first = false;
    }
  }
  std::printf("}, ");
}

int main()
{
  hash_vectors();
  exchange_host_logic();
  fast_int_identities<std::int32_t>();
  fast_int_identities<std::uint32_t>();
  fast_int_identities<std::int64_t>();
  fast_int_identities<std::uint64_t>();
  extents();
  probing_host_sequences();
  pair_layout();
  std::printf("{");
  print_primes();
  print_owners();
  std::printf("\"total\": %d, \"failed\": %d, \"failures\": [", g_total, g_failed);
  for (std::size_t i = 0; i < g_failures.size(); ++i) {
    std::printf("%s\"%s\"", i ? ", " : "", g_failures[i].c_str());
  }
  std::printf("]}\n");
  return g_failed ? 1 : 0;
}
