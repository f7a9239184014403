// Round-2 hardware probes (measurement tool, not product code).
//
// A. What bounds pass 2 of the L2-blocked insert? Claims confined to one L2-resident window of the
//    table at a time (like tools/microbench2.cu `windowed`), comparing the claim protocols:
//      cas128              one 128-bit CAS (what the product does for pair<int64,int64>)
//      cas64               64-bit CAS on the key half only
//      cas64 + st64        key CAS, then a plain 8-byte payload store (cuco's cas_dependent_write)
//      cas64 + cas64       cuco's back_to_back_cas
//      cas32               32-bit CAS (4/8-byte slots of int32 tables)
//      exch128 / st128     upper bounds: unconditional atomic / plain store of the slot
//    each with and without a bulk L2 prefetch of the next window.
// B. ATOMS cost of the route pass: shared-memory atomicAdd ranking with 64 / 200 / 1024 bins, against
//    a two-digit ballot ranking that uses no shared-memory atomics.
// C. Small cp.async.bulk shared->global copies (the per-run copy-out of a routed tile): how many
//    run-sized (256 B - 1 KB) bulk stores per second does the TMA unit sustain?
// Output: one JSON object per line.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

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

__device__ inline std::uint64_t exch128(char* p, std::uint64_t d0, std::uint64_t d1)
{
  std::uint64_t lo, hi;
  asm volatile(
    "{\n\t.reg .b128 d, o;\n\tmov.b128 d, {%2, %3};\n\t"
    "atom.relaxed.gpu.global.exch.b128 o, [%4], d;\n\tmov.b128 {%0, %1}, o;\n\t}"
    : "=l"(lo), "=l"(hi)
    : "l"(d0), "l"(d1), "l"(p)
    : "memory");
  return lo ^ hi;
}

constexpr int kBlock = 256;
constexpr int kMlp   = 4;

enum { C_CAS128 = 0, C_CAS64 = 1, C_CAS64_ST = 2, C_CAS64X2 = 3, C_CAS32 = 4, C_EXCH128 = 5, C_ST128 = 6, C_READ32 = 7 };

template <int CLAIM, bool PREFETCH>
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

  if constexpr (PREFETCH) {
    if (window + 1 < num_windows && threadIdx.x == 0) {
      std::uint64_t const share = window_bytes / tiles_per_window;
      char const* next          = base + window_bytes + within * share;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(next), "r"((unsigned)share) : "memory");
    }
  }

  std::uint64_t const sectors = window_bytes / 32;
#pragma unroll
  for (int j = 0; j < kMlp; ++j) {
    std::uint64_t const h = mix64((tile * kMlp + j) * kBlock + threadIdx.x + 0x9e37ull);
    std::uint64_t const u = __umul64hi(h, sectors);
    char* const p         = base + u * 32 + ((h & 1) ? 16 : 0);
    auto* const q         = reinterpret_cast<unsigned long long*>(p);
    if constexpr (CLAIM == C_CAS128) {
      acc ^= cas128(p, ~0ull, ~0ull, h | 1, tile);
    } else if constexpr (CLAIM == C_CAS64) {
      acc ^= atomicCAS(q, ~0ull, (unsigned long long)(h | 1));
    } else if constexpr (CLAIM == C_CAS64_ST) {
      acc ^= atomicCAS(q, ~0ull, (unsigned long long)(h | 1));
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(q + 1), "l"(tile) : "memory");
    } else if constexpr (CLAIM == C_CAS64X2) {
      acc ^= atomicCAS(q, ~0ull, (unsigned long long)(h | 1));
      acc ^= atomicCAS(q + 1, ~0ull, (unsigned long long)tile);
    } else if constexpr (CLAIM == C_CAS32) {
      acc ^= atomicCAS(reinterpret_cast<unsigned int*>(p), ~0u, (unsigned)(h | 1));
    } else if constexpr (CLAIM == C_EXCH128) {
      acc ^= exch128(p, h | 1, tile);
    } else if constexpr (CLAIM == C_ST128) {
      asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(h | 1), "l"(tile) : "memory");
    } else {
      unsigned long long a, b, c, d;
      asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                   : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                   : "l"(base + u * 32)
                   : "memory");
      acc ^= a ^ b ^ c ^ d;
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
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

// ---------------------------------------------------------------------------------------------
// B. ranking of a 4096-element tile by region (no copy-out): shared atomics vs two-digit ballots
// ---------------------------------------------------------------------------------------------
struct elem16 {
  std::uint64_t k, v;
};

__global__ void fill_keys(elem16* in, std::uint64_t n)
{
  for (std::uint64_t i = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (std::uint64_t)gridDim.x * blockDim.x) {
    in[i] = elem16{mix64(i + 12345), i};
  }
}

constexpr int kRouteBlock = 512;
constexpr int kItems      = 8;

template <int R_MAX>
__global__ void __launch_bounds__(kRouteBlock) rank_atomics(elem16 const* __restrict__ in, std::uint64_t n, unsigned R, unsigned* sink)
{
  __shared__ unsigned hist[R_MAX];
  constexpr std::uint64_t tile = std::uint64_t{kRouteBlock} * kItems;
  unsigned acc                 = 0;
  for (std::uint64_t base = blockIdx.x * tile; base < n; base += gridDim.x * tile) {
    for (unsigned r = threadIdx.x; r < R; r += kRouteBlock) { hist[r] = 0; }
    __syncthreads();
    uint4 raw[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      std::uint64_t const idx = base + std::uint64_t{j} * kRouteBlock + threadIdx.x;
      raw[j] = idx < n ? __ldcs(reinterpret_cast<uint4 const*>(in + idx)) : uint4{0, 0, 0, 0};
    }
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      std::uint64_t const k = (std::uint64_t)raw[j].x | ((std::uint64_t)raw[j].y << 32);
      unsigned const b      = (unsigned)__umul64hi(mix64(k), (std::uint64_t)R);
      acc += atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
  }
  if (acc == 0x12345u) { *sink = acc; }
}

// Two-digit ranking: stable 16-way split by the low digit, then by the high digit. Within a warp a
// row of 32 elements is ranked by four ballots (peers = lanes with the same digit); lane d < 16 of
// every warp carries the warp's running count of digit d. No shared-memory atomics.
template <bool SECOND_PASS>
__global__ void __launch_bounds__(kRouteBlock) rank_ballots(elem16 const* __restrict__ in, std::uint64_t n, unsigned R, unsigned* sink)
{
  constexpr int warps = kRouteBlock / 32;
  __shared__ unsigned warp_digit[warps][16];   // count of digit d in warp w (pass A / pass B)
  __shared__ unsigned digit_base[warps][16];   // position of the first element of (d, w)
  __shared__ unsigned short order_a[kRouteBlock * kItems];  // pass-A order: element ids
  __shared__ unsigned short region_of[kRouteBlock * kItems];
  constexpr std::uint64_t tile = std::uint64_t{kRouteBlock} * kItems;
  unsigned const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned acc        = 0;
  for (std::uint64_t base = blockIdx.x * tile; base < n; base += gridDim.x * tile) {
    uint4 raw[kItems];
    unsigned region[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      std::uint64_t const idx = base + (std::uint64_t{warp} * kItems + j) * 32 + lane;  // warp-contiguous rows
      raw[j] = idx < n ? __ldcs(reinterpret_cast<uint4 const*>(in + idx)) : uint4{0, 0, 0, 0};
    }
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      std::uint64_t const k = (std::uint64_t)raw[j].x | ((std::uint64_t)raw[j].y << 32);
      region[j]             = (unsigned)__umul64hi(mix64(k), (std::uint64_t)R);
    }
    // ---- pass A: by low digit ----
    unsigned running = 0;  // lane d: elements of digit d seen so far in this warp
    unsigned rank_a[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      unsigned const d  = region[j] & 15u;
      unsigned const b0 = __ballot_sync(0xffffffffu, d & 1u), b1 = __ballot_sync(0xffffffffu, d & 2u);
      unsigned const b2 = __ballot_sync(0xffffffffu, d & 4u), b3 = __ballot_sync(0xffffffffu, d & 8u);
      unsigned const peers = ((d & 1u) ? b0 : ~b0) & ((d & 2u) ? b1 : ~b1) & ((d & 4u) ? b2 : ~b2) & ((d & 8u) ? b3 : ~b3);
      unsigned const mine  = ((lane & 1u) ? b0 : ~b0) & ((lane & 2u) ? b1 : ~b1) & ((lane & 4u) ? b2 : ~b2) & ((lane & 8u) ? b3 : ~b3);
      unsigned const before = __shfl_sync(0xffffffffu, running, d);
      rank_a[j]             = before + __popc(peers & ((1u << lane) - 1u));
      running += __popc(mine);  // lanes >= 16 carry junk that nobody reads
    }
    if (lane < 16) { warp_digit[warp][lane] = running; }
    __syncthreads();
    if (threadIdx.x < 16) {
      // digit-major exclusive scan over (digit, warp): 256 values, one thread per digit then a fix-up
      unsigned sum = 0;
      for (int w = 0; w < warps; ++w) {
        unsigned const c              = warp_digit[w][threadIdx.x];
        digit_base[w][threadIdx.x]    = sum;
        sum += c;
      }
      warp_digit[0][threadIdx.x] = sum;  // total of the digit
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned run = 0;
      for (int d = 0; d < 16; ++d) {
        unsigned const c = warp_digit[0][d];
        warp_digit[0][d] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      unsigned const d   = region[j] & 15u;
      unsigned const pos = warp_digit[0][d] + digit_base[warp][d] + rank_a[j];
      unsigned const id  = (warp * kItems + j) * 32 + lane;
      if constexpr (SECOND_PASS) {
        order_a[pos]   = (unsigned short)id;
        region_of[pos] = (unsigned short)region[j];
      } else {
        acc += pos;
      }
    }
    __syncthreads();
    if constexpr (SECOND_PASS) {
      // ---- pass B: by high digit, reading in pass-A order ----
      running = 0;
      unsigned rank_b[kItems], reg_b[kItems];
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        unsigned const at = (warp * kItems + j) * 32 + lane;
        reg_b[j]          = region_of[at];
        unsigned const d  = (reg_b[j] >> 4) & 15u;
        unsigned const b0 = __ballot_sync(0xffffffffu, d & 1u), b1 = __ballot_sync(0xffffffffu, d & 2u);
        unsigned const b2 = __ballot_sync(0xffffffffu, d & 4u), b3 = __ballot_sync(0xffffffffu, d & 8u);
        unsigned const peers = ((d & 1u) ? b0 : ~b0) & ((d & 2u) ? b1 : ~b1) & ((d & 4u) ? b2 : ~b2) & ((d & 8u) ? b3 : ~b3);
        unsigned const mine  = ((lane & 1u) ? b0 : ~b0) & ((lane & 2u) ? b1 : ~b1) & ((lane & 4u) ? b2 : ~b2) & ((lane & 8u) ? b3 : ~b3);
        unsigned const before = __shfl_sync(0xffffffffu, running, d);
        rank_b[j]             = before + __popc(peers & ((1u << lane) - 1u));
        running += __popc(mine);
      }
      if (lane < 16) { warp_digit[warp][lane] = running; }
      __syncthreads();
      if (threadIdx.x < 16) {
        unsigned sum = 0;
        for (int w = 0; w < warps; ++w) {
          unsigned const c           = warp_digit[w][threadIdx.x];
          digit_base[w][threadIdx.x] = sum;
          sum += c;
        }
        warp_digit[0][threadIdx.x] = sum;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int d = 0; d < 16; ++d) {
          unsigned const c = warp_digit[0][d];
          warp_digit[0][d] = run;
          run += c;
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        unsigned const d = (reg_b[j] >> 4) & 15u;
        acc += warp_digit[0][d] + digit_base[warp][d] + rank_b[j] + order_a[(warp * kItems + j) * 32 + lane];
      }
      __syncthreads();
    }
  }
  if (acc == 0x12345u) { *sink = acc; }
}

// ---------------------------------------------------------------------------------------------
// C. small bulk shared->global copies
// ---------------------------------------------------------------------------------------------
// Every CTA owns a 64 KB shared tile and repeatedly stores it to global memory as runs of `run`
// bytes with cp.async.bulk (issued by `issuers` threads), or with ordinary 16-byte thread stores.
template <bool BULK>
__global__ void __launch_bounds__(512) copy_out(char* out, std::uint64_t out_bytes, unsigned run, int reps, int issuers)
{
  extern __shared__ __align__(128) unsigned char tile[];
  constexpr unsigned tile_bytes = 64 * 1024;
  for (unsigned i = threadIdx.x * 16; i < tile_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(tile + i) = uint4{i, blockIdx.x, 3u, 4u};
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  unsigned const runs = tile_bytes / run;
  for (int r = 0; r < reps; ++r) {
    std::uint64_t const slab = (std::uint64_t(blockIdx.x) * reps + r) * tile_bytes % (out_bytes - tile_bytes);
    char* const dst          = out + (slab & ~std::uint64_t{127});
    if constexpr (BULK) {
      if ((int)threadIdx.x < issuers) {
        for (unsigned q = threadIdx.x; q < runs; q += issuers) {
          // scatter the runs over the slab like segment runs (still inside it)
          unsigned const where = (q * 2654435761u) % runs;
          unsigned const src   = static_cast<unsigned>(__cvta_generic_to_shared(tile + q * run));
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + std::uint64_t{where} * run),
                       "r"(src), "r"(run)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    } else {
      for (unsigned i = threadIdx.x * 16; i < tile_bytes; i += blockDim.x * 16) {
        unsigned const q     = i / run;
        unsigned const where = (q * 2654435761u) % runs;
        *reinterpret_cast<uint4*>(dst + std::uint64_t{where} * run + (i - q * run)) = *reinterpret_cast<uint4 const*>(tile + i);
      }
      __syncthreads();
    }
  }
}

int main(int argc, char** argv)
{
  std::uint64_t const gib   = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 3;
  std::uint64_t const bytes = gib << 30;
  char* buf;
  std::uint64_t* sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 8));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  // ---- A ----
  auto run_windowed = [&](auto claim_tag, auto pf_tag, std::uint64_t window_mib, double density) {
    constexpr int CLAIM = decltype(claim_tag)::value;
    constexpr bool PF   = decltype(pf_tag)::value;
    std::uint64_t const window_bytes = window_mib << 20;
    std::uint64_t const num_windows  = bytes / window_bytes;
    std::uint64_t const per_tile     = std::uint64_t{kBlock} * kMlp;
    std::uint64_t want               = (std::uint64_t)(window_bytes / 32 * density) / per_tile;
    std::uint64_t tiles_per_window   = 1;
    while (tiles_per_window * 2 <= want) { tiles_per_window *= 2; }
    float best = 1e30f;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(buf, 0xff, bytes));
      CK(cudaDeviceSynchronize());
      float const ms = time_once([&] {
        windowed<CLAIM, PF><<<(unsigned)(tiles_per_window * num_windows), kBlock>>>(
          buf, window_bytes, tiles_per_window, num_windows, sink);
      });
      best = ms < best ? ms : best;
    }
    double const ops     = (double)tiles_per_window * num_windows * per_tile;
    char const* claims[] = {"cas128", "cas64", "cas64+st64", "cas64+cas64", "cas32", "exch128", "st128", "read32"};
    std::printf("{\"test\": \"windowed_claim\", \"claim\": \"%s\", \"prefetch\": %d, \"window_MiB\": %llu, "
                "\"accesses_per_sector\": %.3f, \"ms\": %.3f, \"Gops_s\": %.2f}\n",
                claims[CLAIM], (int)PF, (unsigned long long)window_mib, ops / (bytes / 32.0), best, ops / best / 1e6);
    std::fflush(stdout);
  };
#define RW(C, P, W, D) run_windowed(std::integral_constant<int, C>{}, std::integral_constant<bool, P>{}, W, D)
  for (std::uint64_t w : {16ull, 32ull}) {
    RW(C_CAS128, false, w, 1.0);
    RW(C_CAS128, true, w, 1.0);
    RW(C_CAS64, false, w, 1.0);
    RW(C_CAS64, true, w, 1.0);
    RW(C_CAS64_ST, false, w, 1.0);
    RW(C_CAS64_ST, true, w, 1.0);
    RW(C_CAS64X2, true, w, 1.0);
    RW(C_CAS32, true, w, 1.0);
    RW(C_EXCH128, true, w, 1.0);
    RW(C_ST128, false, w, 1.0);
    RW(C_ST128, true, w, 1.0);
    RW(C_READ32, false, w, 1.0);
    RW(C_READ32, true, w, 1.0);
  }
  RW(C_CAS128, true, 16, 0.5);
  RW(C_CAS64_ST, true, 16, 0.5);
  RW(C_CAS128, true, 8, 1.0);
  RW(C_CAS64_ST, true, 8, 1.0);

  // ---- B ----
  {
    std::uint64_t const n = 100'000'000ull;
    elem16* in;
    unsigned* sink32;
    CK(cudaMalloc(&in, n * sizeof(elem16)));
    CK(cudaMalloc(&sink32, 4));
    fill_keys<<<1024, 256>>>(in, n);
    CK(cudaDeviceSynchronize());
    unsigned const grid = (unsigned)((n + kRouteBlock * kItems - 1) / (kRouteBlock * kItems));
    auto report = [&](char const* name, unsigned R, float ms) {
      std::printf("{\"test\": \"rank\", \"scheme\": \"%s\", \"regions\": %u, \"ms\": %.3f, \"Gelem_s\": %.2f}\n", name, R,
                  ms, n / ms / 1e6);
      std::fflush(stdout);
    };
    for (unsigned R : {16u, 64u, 200u, 256u, 1024u}) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        float const ms = time_once([&] { rank_atomics<1024><<<grid, kRouteBlock>>>(in, n, R, sink32); });
        best           = ms < best ? ms : best;
      }
      report("shared atomics", R, best);
    }
    {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        float const ms = time_once([&] { rank_ballots<false><<<grid, kRouteBlock>>>(in, n, 16, sink32); });
        best           = ms < best ? ms : best;
      }
      report("ballots, one digit (16 regions)", 16, best);
      best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        float const ms = time_once([&] { rank_ballots<true><<<grid, kRouteBlock>>>(in, n, 256, sink32); });
        best           = ms < best ? ms : best;
      }
      report("ballots, two digits (256 regions)", 256, best);
    }
    CK(cudaFree(in));
    CK(cudaFree(sink32));
  }

  // ---- C ----
  {
    std::size_t const smem = 64 * 1024;
    CK(cudaFuncSetAttribute(copy_out<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(copy_out<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int const reps      = 16;
    unsigned const grid = sms * 2 * 8;  // 2 CTAs per SM resident, 8 waves
    for (unsigned run : {128u, 256u, 512u, 1024u, 4096u, 65536u}) {
      for (int issuers : {1, 32, 128}) {
        float best = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {
          float const ms = time_once([&] { copy_out<true><<<grid, 512, smem>>>(buf, bytes, run, reps, issuers); });
          best           = ms < best ? ms : best;
        }
        double const moved = (double)grid * reps * 65536.0;
        std::printf("{\"test\": \"copy_out\", \"how\": \"cp.async.bulk\", \"run_bytes\": %u, \"issuers\": %d, \"ms\": %.3f, "
                    "\"GBps\": %.1f, \"Mruns_s\": %.1f}\n",
                    run, issuers, best, moved / best / 1e6, moved / run / best / 1e3);
        std::fflush(stdout);
      }
      float best = 1e30f;
      for (int rep = 0; rep < 2; ++rep) {
        float const ms = time_once([&] { copy_out<false><<<grid, 512, smem>>>(buf, bytes, run, reps, 0); });
        best           = ms < best ? ms : best;
      }
      double const moved = (double)grid * reps * 65536.0;
      std::printf("{\"test\": \"copy_out\", \"how\": \"thread stores\", \"run_bytes\": %u, \"issuers\": 0, \"ms\": %.3f, "
                  "\"GBps\": %.1f}\n",
                  run, best, moved / best / 1e6);
      std::fflush(stdout);
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
