// Second set of hardware measurements on B200 (measurement tool, not product code): does blocking
// the table by L2-sized windows pay, and how fast can a batch be partitioned by window?
//   * windowed_read / windowed_cas: the accesses of one launch sweep the table window by window
//     (every access is a random 32-byte sector *inside* the current window, about one access per
//     sector), optionally with the next window streamed into L2 ahead of use
//     (prefetch.global.L2 per line, or one cp.async.bulk.prefetch.L2 per CTA)
//   * partition: 16-byte elements scattered into R contiguous segments through a shared-memory
//     staging tile (one global atomic per segment per tile, line-sized coalesced runs out)
// Output: one JSON object per line. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
      std::exit(1);                                                                  \
    }                                                                                \
  } while (0)

__host__ __device__ inline std::uint64_t mix64(std::uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__device__ inline std::uint64_t load32(char const* p)
{
  unsigned long long a, b, c, d;
  asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(p)
               : "memory");
  return a ^ b ^ c ^ d;
}

__device__ inline std::uint64_t cas128(char* p, std::uint64_t e0, std::uint64_t e1, std::uint64_t d0, std::uint64_t d1)
{
  std::uint64_t lo, hi;
  asm volatile(
    "{\n\t.reg .b128 e, d, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
    "atom.relaxed.gpu.global.cas.b128 o, [%6], e, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
    : "=l"(lo), "=l"(hi)
    : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "l"(p)
    : "memory");
  return lo ^ hi;
}

constexpr int kBlock = 256;
constexpr int kMlp   = 4;

// MODE 0: 32-byte read; 1: cas128 (succeeds on a table of all-ones); 2: read then cas128 on the
// same slot (the insert sequence). PREFETCH 0: none, 1: per-line prefetch.global.L2 of the next
// window, 2: one bulk L2 prefetch per CTA, 3: explicit 16-byte loads of the next window (discarded)
template <int MODE, int PREFETCH>
__global__ void __launch_bounds__(kBlock) windowed(char* buf,
                                                   std::uint64_t window_bytes,
                                                   std::uint64_t tiles_per_window,
                                                   std::uint64_t num_windows,
                                                   std::uint64_t* sink)
{
  std::uint64_t const tile   = blockIdx.x;
  std::uint64_t const window = tile / tiles_per_window;
  std::uint64_t const within = tile - window * tiles_per_window;
  char* const base           = buf + window * window_bytes;
  std::uint64_t acc          = 0;

  if constexpr (PREFETCH != 0) {
    if (window + 1 < num_windows) {
      std::uint64_t const share = window_bytes / tiles_per_window;  // multiple of 128 by construction
      char const* next          = base + window_bytes + within * share;
      if constexpr (PREFETCH == 1) {
        for (std::uint64_t off = threadIdx.x * 128ull; off < share; off += kBlock * 128ull) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(next + off));
        }
      } else if constexpr (PREFETCH == 2) {
        if (threadIdx.x == 0) {
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(next), "r"((unsigned)share) : "memory");
        }
      } else {
        for (std::uint64_t off = threadIdx.x * 16ull; off < share; off += kBlock * 16ull) {
          unsigned a, b, c, d;
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                       : "l"(next + off));
          acc += (a == 0x12345u) + (b == 0x7777u) + (c == 0x999u) + (d == 0x31337u);
        }
      }
    }
  }

  std::uint64_t const sectors = window_bytes / 32;
  std::uint64_t v[kMlp];
  char* p[kMlp];
#pragma unroll
  for (int j = 0; j < kMlp; ++j) {
    std::uint64_t const h = mix64((tile * kMlp + j) * kBlock + threadIdx.x + 0x9e37ull);
    std::uint64_t const u = __umul64hi(h, sectors);
    p[j]                  = base + u * 32 + ((h & 1) ? 16 : 0);
    if constexpr (MODE == 0) {
      v[j] = load32(base + u * 32);
    } else if constexpr (MODE == 1) {
      v[j] = cas128(p[j], ~0ull, ~0ull, h | 1, tile);
    } else {
      v[j] = load32(base + u * 32);
    }
  }
#pragma unroll
  for (int j = 0; j < kMlp; ++j) {
    if constexpr (MODE == 2) {
      // the CAS depends on the load having returned, like an insert
      std::uint64_t const e = (v[j] == 0x5555aaaaull) ? 1ull : ~0ull;
      acc ^= cas128(p[j], e, ~0ull, mix64(v[j] + j) | 1, tile);
    } else {
      acc ^= v[j];
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

struct elem16 {
  std::uint64_t k, v;
};

constexpr int kItems = 16;  // per thread; tile = 4096 elements = 64 KB of staging

// Scatter n elements into R segments of fixed capacity (seg_cap elements each).
template <int R_MAX>
__global__ void __launch_bounds__(kBlock) partition(elem16 const* __restrict__ in,
                                                    std::uint64_t n,
                                                    elem16* __restrict__ out,
                                                    unsigned int* cursors,
                                                    unsigned int R,
                                                    std::uint64_t seg_cap)
{
  extern __shared__ __align__(16) unsigned char smem[];
  elem16* const stage         = reinterpret_cast<elem16*>(smem);                       // [tile]
  unsigned short* const owner = reinterpret_cast<unsigned short*>(stage + kBlock * kItems);  // [tile]
  __shared__ unsigned int hist[R_MAX];
  __shared__ unsigned int local_base[R_MAX];
  __shared__ unsigned int global_base[R_MAX];
  __shared__ unsigned int warp_sums[kBlock / 32];

  constexpr std::uint64_t tile = std::uint64_t{kBlock} * kItems;
  std::uint64_t const base     = blockIdx.x * tile;

  for (unsigned r = threadIdx.x; r < R; r += kBlock) { hist[r] = 0; }
  __syncthreads();

  elem16 e[kItems];
  unsigned int bucket[kItems];
  unsigned int rank[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    std::uint64_t const idx = base + std::uint64_t{j} * kBlock + threadIdx.x;
    bucket[j]               = 0xffffffffu;
    if (idx < n) {
      uint4 const raw = __ldcs(reinterpret_cast<uint4 const*>(in + idx));
      e[j].k          = (std::uint64_t)raw.x | ((std::uint64_t)raw.y << 32);
      e[j].v          = (std::uint64_t)raw.z | ((std::uint64_t)raw.w << 32);
      bucket[j]       = (unsigned)__umul64hi(mix64(e[j].k), (std::uint64_t)R);
    }
  }
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    if (bucket[j] != 0xffffffffu) { rank[j] = atomicAdd(&hist[bucket[j]], 1u); }
  }
  __syncthreads();
  // exclusive scan of hist (R <= R_MAX <= 4 * kBlock): each thread owns up to 4 consecutive entries
  {
    constexpr int per = (R_MAX + kBlock - 1) / kBlock;
    unsigned int mine[per];
    unsigned int sum = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) {
      unsigned const r = threadIdx.x * per + i;
      mine[i]          = r < R ? hist[r] : 0u;
      sum += mine[i];
    }
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned int const up = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) { incl += up; }
    }
    if ((threadIdx.x & 31) == 31) { warp_sums[threadIdx.x >> 5] = incl; }
    __syncthreads();
    unsigned int offset = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) { offset += warp_sums[w]; }
    unsigned int running = offset + incl - sum;
#pragma unroll
    for (int i = 0; i < per; ++i) {
      unsigned const r = threadIdx.x * per + i;
      if (r < R) {
        local_base[r]  = running;
        global_base[r] = mine[i] ? atomicAdd(&cursors[r], mine[i]) : 0u;
        running += mine[i];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    if (bucket[j] != 0xffffffffu) {
      unsigned const pos = local_base[bucket[j]] + rank[j];
      stage[pos]         = e[j];
      owner[pos]         = (unsigned short)bucket[j];
    }
  }
  __syncthreads();
  std::uint64_t const count = (n - base) < tile ? (n - base) : tile;
  for (unsigned pos = threadIdx.x; pos < count; pos += kBlock) {
    unsigned const b          = owner[pos];
    std::uint64_t const where = (std::uint64_t)global_base[b] + (pos - local_base[b]);
    if (where < seg_cap) { out[b * seg_cap + where] = stage[pos]; }
  }
}

__global__ void fill_keys(elem16* in, std::uint64_t n)
{
  for (std::uint64_t i = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (std::uint64_t)gridDim.x * blockDim.x) {
    in[i] = elem16{mix64(i + 12345), i};
  }
}

template <typename F>
float time_once(F&& launch)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a));
  launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  CK(cudaEventDestroy(a));
  CK(cudaEventDestroy(b));
  return ms;
}

int main(int argc, char** argv)
{
  std::uint64_t const gib   = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 3;
  std::uint64_t const bytes = gib << 30;
  char* buf;
  std::uint64_t* sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 8));

  {
    std::size_t now = 0;
    cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
    std::size_t persist = 0;
    cudaDeviceGetLimit(&persist, cudaLimitPersistingL2CacheSize);
    std::printf("{\"test\": \"defaults\", \"l2_fetch_granularity\": %zu, \"persisting_l2\": %zu}\n", now, persist);
  }
  auto run_windowed = [&](auto mode_tag, auto pf_tag, std::uint64_t window_mib, double density) {
    constexpr int MODE = decltype(mode_tag)::value;
    constexpr int PF   = decltype(pf_tag)::value;
    std::uint64_t const window_bytes = window_mib << 20;
    std::uint64_t const num_windows  = bytes / window_bytes;
    std::uint64_t const per_tile     = std::uint64_t{kBlock} * kMlp;
    // tiles per window: density accesses per sector, rounded to a power of two so the prefetch
    // share stays a multiple of 128 bytes
    std::uint64_t want = (std::uint64_t)(window_bytes / 32 * density) / per_tile;
    std::uint64_t tiles_per_window = 1;
    while (tiles_per_window * 2 <= want) { tiles_per_window *= 2; }
    float best = 1e30f;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(buf, 0xff, bytes));
      CK(cudaDeviceSynchronize());
      float const ms = time_once([&] {
        windowed<MODE, PF><<<(unsigned)(tiles_per_window * num_windows), kBlock>>>(
          buf, window_bytes, tiles_per_window, num_windows, sink);
      });
      best = ms < best ? ms : best;
    }
    double const ops = (double)tiles_per_window * num_windows * per_tile;
    char const* modes[] = {"read32", "cas128", "read32+cas128"};
    char const* pfs[]   = {"none", "prefetch.L2 per line", "bulk prefetch.L2 per CTA", "plain loads"};
    std::printf("{\"test\": \"windowed\", \"mode\": \"%s\", \"prefetch\": \"%s\", \"window_MiB\": %llu, "
                "\"accesses_per_sector\": %.3f, \"table_GiB\": %llu, \"ms\": %.3f, \"Gops_s\": %.2f, "
                "\"table_stream_GBps\": %.1f}\n",
                modes[MODE], pfs[PF], (unsigned long long)window_mib, ops / (bytes / 32.0),
                (unsigned long long)gib, best, ops / best / 1e6, bytes / best / 1e6);
    std::fflush(stdout);
  };
#define RW(M, P, W, D) run_windowed(std::integral_constant<int, M>{}, std::integral_constant<int, P>{}, W, D)
  for (std::uint64_t w : {16ull, 32ull, 64ull}) {
    RW(0, 0, w, 1.0);
    RW(0, 1, w, 1.0);
    RW(0, 2, w, 1.0);
    RW(0, 3, w, 1.0);
    RW(1, 0, w, 1.0);
    RW(1, 1, w, 1.0);
    RW(1, 2, w, 1.0);
    RW(2, 0, w, 1.0);
    RW(2, 1, w, 1.0);
    RW(2, 2, w, 1.0);
    RW(2, 3, w, 1.0);
  }
  // whole table as one window = the unblocked pattern, at each L2 fetch granularity the driver
  // accepts (ncu shows 4 DRAM sectors per random 32-byte read with the default setting)
  for (std::size_t g : {std::size_t{128}, std::size_t{64}, std::size_t{32}}) {
    cudaError_t const rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
    std::size_t now = 0;
    cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
    std::printf("{\"test\": \"set_l2_fetch_granularity\", \"requested\": %zu, \"rc\": \"%s\", \"now\": %zu}\n",
                g, cudaGetErrorName(rc), now);
    RW(0, 0, gib << 10, 1.0);
    RW(1, 0, gib << 10, 1.0);
    RW(2, 0, gib << 10, 1.0);
    RW(2, 2, 32, 1.0);
  }
  CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128));
  // fewer accesses per sector (load factor 0.5 with half the batch being duplicates ~ 0.5)
  RW(2, 1, 32, 0.5);
  RW(2, 2, 32, 0.5);
  RW(2, 2, 32, 2.0);

  // ---- partition ----
  {
    std::uint64_t const n = 100'000'000ull;
    elem16 *in, *out;
    unsigned int* cursors;
    CK(cudaMalloc(&in, n * sizeof(elem16)));
    fill_keys<<<1024, 256>>>(in, n);
    CK(cudaDeviceSynchronize());
    for (unsigned R : {64u, 128u, 256u, 512u, 1024u}) {
      std::uint64_t const seg_cap = (n / R) + (n / R) / 16 + 4096;
      CK(cudaMalloc(&out, R * seg_cap * sizeof(elem16)));
      CK(cudaMalloc(&cursors, R * sizeof(unsigned int)));
      std::size_t const smem = kBlock * kItems * (sizeof(elem16) + sizeof(unsigned short));
      CK(cudaFuncSetAttribute(partition<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      unsigned const grid = (unsigned)((n + kBlock * kItems - 1) / (kBlock * kItems));
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(cursors, 0, R * sizeof(unsigned int)));
        float const ms = time_once([&] { partition<1024><<<grid, kBlock, smem>>>(in, n, out, cursors, R, seg_cap); });
        best = ms < best ? ms : best;
      }
      std::vector<unsigned int> h(R);
      CK(cudaMemcpy(h.data(), cursors, R * sizeof(unsigned int), cudaMemcpyDeviceToHost));
      std::uint64_t total = 0, mx = 0;
      for (auto c : h) {
        total += c;
        mx = c > mx ? c : mx;
      }
      std::printf("{\"test\": \"partition\", \"segments\": %u, \"n\": %llu, \"ms\": %.3f, \"Gelem_s\": %.2f, "
                  "\"GBps_read_plus_write\": %.1f, \"routed\": %llu, \"max_segment\": %llu, \"segment_capacity\": %llu}\n",
                  R, (unsigned long long)n, best, n / best / 1e6, 2.0 * n * 16 / best / 1e6,
                  (unsigned long long)total, (unsigned long long)mx, (unsigned long long)seg_cap);
      std::fflush(stdout);
      CK(cudaFree(out));
      CK(cudaFree(cursors));
    }
    CK(cudaFree(in));
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
