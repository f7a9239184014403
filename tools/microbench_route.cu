// Fourth hardware probe (measurement tool, not product code; prepared for the next tuning pass):
// what bounds pass 1 of the L2-blocked insert (`route_kernel`, 0.98 ms per 100 M 16-byte pairs =
// 3.5 TB/s against a 6.5 TB/s copy peak)? The product kernel ranks every element of a 4096-element
// tile with one shared-memory atomic, stages the tile in shared memory grouped by region and copies
// it out in line-sized runs. Variants of that one kernel, same tile geometry (512 threads x 8):
//   rank = smem_atomic   one ATOMS per element (what the product does)
//   rank = match_any     one ATOMS per distinct region per warp (__match_any_sync + popc)
//   rank = none          no ranking at all: elements are staged in input order (lower bound of the
//                        staging + copy-out cost; output is NOT a partition)
//   out  = staged        shared-memory staging, contiguous runs out (product)
//   out  = direct        each element stored straight to its reserved position (16-byte scattered
//                        stores, no staging, one sync less)
//   out  = none          ranking only, nothing written (lower bound of the ranking cost)
// plus the region count R (64 / 200 / 1024) and 8 vs 16 items per thread.
// Output: one JSON object per line. Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench_route.cu -o tools/_build/microbench_route
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

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

struct elem16 {
  std::uint64_t k, v;
};

constexpr int kBlock = 512;
constexpr int kRMax  = 1024;

enum rank_mode { rank_smem_atomic = 0, rank_match_any = 1, rank_none = 2 };
enum out_mode { out_staged = 0, out_direct = 1, out_none = 2 };

template <int RANK, int OUT, int ITEMS>
__global__ void __launch_bounds__(kBlock) route(elem16 const* __restrict__ in,
                                                std::uint64_t n,
                                                elem16* __restrict__ out,
                                                unsigned int* cursors,
                                                unsigned int R,
                                                std::uint64_t seg_cap,
                                                unsigned long long* sink)
{
  extern __shared__ __align__(16) unsigned char smem[];
  elem16* const stage         = reinterpret_cast<elem16*>(smem);                             // [tile]
  unsigned short* const owner = reinterpret_cast<unsigned short*>(stage + kBlock * ITEMS);   // [tile]
  __shared__ unsigned int hist[kRMax];
  __shared__ unsigned int local_base[kRMax];
  __shared__ unsigned int global_base[kRMax];
  __shared__ unsigned int warp_sums[kBlock / 32];

  constexpr std::uint64_t tile = std::uint64_t{kBlock} * ITEMS;
  std::uint64_t const base     = blockIdx.x * tile;
  unsigned const lane          = threadIdx.x & 31;

  for (unsigned r = threadIdx.x; r < R; r += kBlock) { hist[r] = 0; }
  __syncthreads();

  elem16 e[ITEMS];
  unsigned int bucket[ITEMS];
  unsigned int rank[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    std::uint64_t const idx = base + static_cast<std::uint64_t>(j) * kBlock + threadIdx.x;
    bucket[j]               = 0xffffffffu;
    if (idx < n) {
      uint4 const raw = __ldcs(reinterpret_cast<uint4 const*>(in + idx));
      e[j].k          = (std::uint64_t)raw.x | ((std::uint64_t)raw.y << 32);
      e[j].v          = (std::uint64_t)raw.z | ((std::uint64_t)raw.w << 32);
    }
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    std::uint64_t const idx = base + static_cast<std::uint64_t>(j) * kBlock + threadIdx.x;
    if (idx < n) { bucket[j] = (unsigned)__umul64hi(mix64(e[j].k), (std::uint64_t)R); }
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if constexpr (RANK == rank_smem_atomic) {
      if (bucket[j] != 0xffffffffu) { rank[j] = atomicAdd(&hist[bucket[j]], 1u); }
    } else if constexpr (RANK == rank_match_any) {
      unsigned const peers  = __match_any_sync(0xffffffffu, bucket[j]);
      unsigned const leader = __ffs(peers) - 1;
      unsigned start        = 0;
      if (lane == leader && bucket[j] != 0xffffffffu) { start = atomicAdd(&hist[bucket[j]], __popc(peers)); }
      start   = __shfl_sync(0xffffffffu, start, leader);
      rank[j] = start + __popc(peers & ((1u << lane) - 1));
    } else {
      rank[j] = j * kBlock + threadIdx.x;  // input order
    }
  }
  __syncthreads();

  if constexpr (RANK != rank_none) {
    // exclusive scan of hist; thread t owns entries [t * per, (t + 1) * per)
    constexpr int per = (kRMax + kBlock - 1) / kBlock;
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
      if (lane >= (unsigned)d) { incl += up; }
    }
    if (lane == 31) { warp_sums[threadIdx.x >> 5] = incl; }
    __syncthreads();
    unsigned int offset = 0;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) { offset += warp_sums[w]; }
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
    __syncthreads();
  }

  if constexpr (OUT == out_none) {
    unsigned long long acc = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (bucket[j] != 0xffffffffu) { acc += rank[j] + e[j].v; }
    }
    if (acc == 0x123456789abcdefull) { *sink = acc; }
  } else if constexpr (OUT == out_direct) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (bucket[j] != 0xffffffffu) {
        std::uint64_t where;
        if constexpr (RANK == rank_none) {
          where = base + rank[j];
          out[where] = e[j];
        } else {
          where = (std::uint64_t)global_base[bucket[j]] + rank[j];
          if (where < seg_cap) { out[bucket[j] * seg_cap + where] = e[j]; }
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (bucket[j] != 0xffffffffu) {
        unsigned const pos = RANK == rank_none ? rank[j] : local_base[bucket[j]] + rank[j];
        stage[pos]         = e[j];
        owner[pos]         = (unsigned short)bucket[j];
      }
    }
    __syncthreads();
    std::uint64_t const count = (n - base) < tile ? (n - base) : tile;
    for (unsigned pos = threadIdx.x; pos < count; pos += kBlock) {
      if constexpr (RANK == rank_none) {
        out[base + pos] = stage[pos];
      } else {
        unsigned const b          = owner[pos];
        std::uint64_t const where = (std::uint64_t)global_base[b] + (pos - local_base[b]);
        if (where < seg_cap) { out[b * seg_cap + where] = stage[pos]; }
      }
    }
  }
}

__global__ void fill_keys(elem16* in, std::uint64_t n)
{
  for (std::uint64_t i = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (std::uint64_t)gridDim.x * blockDim.x) {
    in[i] = elem16{mix64(i + 12345), i};
  }
}

template <int RANK, int OUT, int ITEMS>
void run(elem16 const* in, std::uint64_t n, elem16* out, unsigned int* cursors, unsigned R, unsigned long long* sink)
{
  char const* ranks[] = {"smem_atomic", "match_any", "none"};
  char const* outs[]  = {"staged", "direct", "none"};
  std::uint64_t const seg_cap = n / R + n / R / 8 + 4096;
  std::size_t const smem      = std::size_t{kBlock} * ITEMS * (sizeof(elem16) + sizeof(unsigned short));
  auto const kernel           = route<RANK, OUT, ITEMS>;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned const grid = (unsigned)((n + std::uint64_t{kBlock} * ITEMS - 1) / (std::uint64_t{kBlock} * ITEMS));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaMemset(cursors, 0, kRMax * sizeof(unsigned int)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    kernel<<<grid, kBlock, smem>>>(in, n, out, cursors, R, seg_cap, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
    if (rep > 0 && ms < best) { best = ms; }
  }
  double const bytes = (OUT == out_none ? 16.0 : 32.0) * (double)n;
  std::printf("{\"test\": \"route\", \"rank\": \"%s\", \"out\": \"%s\", \"items_per_thread\": %d, \"regions\": %u, "
              "\"n\": %llu, \"ms\": %.3f, \"Gelem_s\": %.2f, \"GBps\": %.1f}\n",
              ranks[RANK], outs[OUT], ITEMS, R, (unsigned long long)n, best, n / best / 1e6, bytes / best / 1e6);
  std::fflush(stdout);
}

int main(int argc, char** argv)
{
  std::uint64_t const n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 100000000ull;
  elem16 *in, *out;
  unsigned int* cursors;
  unsigned long long* sink;
  // segments: R * (n / R * 1.125 + 4096) elements
  std::uint64_t const out_elems = n + n / 8 + std::uint64_t{kRMax} * 4096 + 4096;
  CK(cudaMalloc(&in, n * sizeof(elem16)));
  CK(cudaMalloc(&out, out_elems * sizeof(elem16)));
  CK(cudaMalloc(&cursors, kRMax * sizeof(unsigned int)));
  CK(cudaMalloc(&sink, 8));
  fill_keys<<<148 * 8, 256>>>(in, n);
  CK(cudaDeviceSynchronize());

  for (unsigned R : {200u, 64u, 1024u}) {
    run<rank_smem_atomic, out_staged, 8>(in, n, out, cursors, R, sink);  // the product's scheme
    run<rank_match_any, out_staged, 8>(in, n, out, cursors, R, sink);
    run<rank_smem_atomic, out_direct, 8>(in, n, out, cursors, R, sink);
    run<rank_match_any, out_direct, 8>(in, n, out, cursors, R, sink);
    run<rank_smem_atomic, out_none, 8>(in, n, out, cursors, R, sink);
    run<rank_match_any, out_none, 8>(in, n, out, cursors, R, sink);
  }
  run<rank_none, out_staged, 8>(in, n, out, cursors, 200, sink);   // staging + copy-out alone
  run<rank_none, out_direct, 8>(in, n, out, cursors, 200, sink);   // plain streaming copy through registers
  run<rank_smem_atomic, out_staged, 16>(in, n, out, cursors, 200, sink);
  run<rank_match_any, out_staged, 16>(in, n, out, cursors, 200, sink);
  return 0;
}
