// Hardware ceilings for the hash-table access pattern on B200 (measurement tool, not product code):
//   * random aligned reads of 16/32/64/128 B from a large buffer, at several levels of
//     memory-level parallelism per thread  -> the real "random sector" roofline
//   * random 64-bit / 128-bit CAS (all succeeding, and all failing) -> atomic throughput at L2/DRAM
//   * load-then-CAS vs CAS-first on cold lines
//   * streaming copy as the sequential reference
// Output: one JSON object per line. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
      std::exit(1);                                                                \
    }                                                                              \
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

template <int BYTES>
__device__ inline std::uint64_t load_bytes(char const* p)
{
  if constexpr (BYTES == 16) {
    unsigned a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "l"(p));
    return (std::uint64_t)(a ^ c) | ((std::uint64_t)(b ^ d) << 32);
  } else if constexpr (BYTES == 32) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    return a ^ b ^ c ^ d;
  } else {
    std::uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < BYTES / 32; ++i) {
      acc ^= load_bytes<32>(p + 32 * i);
    }
    return acc;
  }
}

// every thread performs `iters` rounds of MLP independent random reads of BYTES
template <int BYTES, int MLP>
__global__ void __launch_bounds__(256) random_read(char const* buf, std::uint64_t n_units, int iters, std::uint64_t* sink)
{
  std::uint64_t const tid = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x;
  std::uint64_t acc       = 0;
  for (int it = 0; it < iters; ++it) {
    std::uint64_t v[MLP];
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      std::uint64_t const h = mix64(tid * 1000003ull + (std::uint64_t)it * MLP + j);
      std::uint64_t const u = __umul64hi(h, n_units);
      v[j]                  = load_bytes<BYTES>(buf + u * BYTES);
    }
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      acc ^= v[j];
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

// dependent chain: next address depends on loaded data (latency probe), one chain per thread
__global__ void __launch_bounds__(256) chase(char const* buf, std::uint64_t n_units, int iters, std::uint64_t* sink)
{
  std::uint64_t const tid = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x;
  std::uint64_t h         = mix64(tid + 1);
  for (int it = 0; it < iters; ++it) {
    std::uint64_t const u = __umul64hi(h, n_units);
    h                     = mix64(h + load_bytes<16>(buf + u * 32) + it);
  }
  if (h == 42) { *sink = h; }
}

template <int MLP, bool WIDE>
__global__ void __launch_bounds__(256) random_cas(char* buf, std::uint64_t n_slots, int iters, std::uint64_t expect, std::uint64_t* sink)
{
  std::uint64_t const tid = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x;
  std::uint64_t acc       = 0;
  for (int it = 0; it < iters; ++it) {
    std::uint64_t lo[MLP], hi[MLP];
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      std::uint64_t const h = mix64(tid * 1000003ull + (std::uint64_t)it * MLP + j);
      std::uint64_t const u = __umul64hi(h, n_slots);
      char* p               = buf + u * 16;
      if constexpr (WIDE) {
        asm volatile(
          "{\n\t.reg .b128 e, d, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
          "atom.relaxed.gpu.global.cas.b128 o, [%6], e, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
          : "=l"(lo[j]), "=l"(hi[j])
          : "l"(expect), "l"(expect), "l"(h | 1), "l"(tid), "l"(p)
          : "memory");
      } else {
        lo[j] = atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)expect, (unsigned long long)(h | 1));
        hi[j] = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      acc ^= lo[j] ^ hi[j];
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

// load 16 B, then two 64-bit CAS (the reference's 16-byte slot protocol) vs one 128-bit CAS
template <int MODE>  // 0: load + cas64 + cas64, 1: load + cas128, 2: cas128 only
__global__ void __launch_bounds__(256) insert_like(char* buf, std::uint64_t n_slots, int iters, std::uint64_t* sink)
{
  std::uint64_t const tid = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x;
  std::uint64_t acc       = 0;
  for (int it = 0; it < iters; ++it) {
    std::uint64_t const h = mix64(tid * 1000003ull + it);
    std::uint64_t const u = __umul64hi(h, n_slots);
    char* p               = buf + u * 16;
    std::uint64_t e0 = ~0ull, e1 = ~0ull;
    if (MODE != 2) {
      unsigned long long a, b;
      asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
      e0 = a;
      e1 = b;
    }
    if (MODE == 0) {
      acc ^= atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)e0, (unsigned long long)(h | 1));
      acc ^= atomicCAS(reinterpret_cast<unsigned long long*>(p) + 1, (unsigned long long)e1, (unsigned long long)tid);
    } else {
      std::uint64_t lo, hi;
      asm volatile(
        "{\n\t.reg .b128 e, d, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], e, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(lo), "=l"(hi)
        : "l"(e0), "l"(e1), "l"(h | 1), "l"(tid), "l"(p)
        : "memory");
      acc ^= lo ^ hi;
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

__global__ void __launch_bounds__(256) stream_copy(uint4 const* in, uint4* out, std::uint64_t n)
{
  for (std::uint64_t i = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (std::uint64_t)gridDim.x * blockDim.x) {
    out[i] = in[i];
  }
}

template <typename F>
float time_ms(F&& launch, int reps = 3)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv)
{
  std::uint64_t const gib = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4;
  std::uint64_t const bytes = gib << 30;
  char* buf;
  std::uint64_t* sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(buf, 0xff, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int const block = 256;

  // streaming reference
  {
    std::uint64_t const n = bytes / 2 / 16;
    float ms = time_ms([&] { stream_copy<<<sms * 8, block>>>((uint4 const*)buf, (uint4*)(buf + bytes / 2), n); });
    std::printf("{\"test\": \"stream_copy\", \"GiB\": %llu, \"ms\": %.3f, \"GBps\": %.1f}\n",
                (unsigned long long)gib, ms, 2.0 * n * 16 / ms / 1e6);
  }

  auto run_read = [&](auto bytes_tag, auto mlp_tag, std::uint64_t span_bytes, int blocks_per_sm) {
    constexpr int B = decltype(bytes_tag)::value;
    constexpr int M = decltype(mlp_tag)::value;
    int const grid  = sms * blocks_per_sm;
    std::uint64_t const threads = (std::uint64_t)grid * block;
    int const iters = (int)((1ull << 28) / (threads * M)) + 1;  // ~268 M accesses
    float ms = time_ms([&] { random_read<B, M><<<grid, block>>>(buf, span_bytes / B, iters, sink); });
    double const acc = (double)threads * M * iters;
    std::printf("{\"test\": \"random_read\", \"bytes\": %d, \"mlp\": %d, \"blocks_per_sm\": %d, \"span_MiB\": %llu, "
                "\"ms\": %.3f, \"Gacc_s\": %.2f, \"GBps\": %.1f}\n",
                B, M, blocks_per_sm, (unsigned long long)(span_bytes >> 20), ms, acc / ms / 1e6, acc * B / ms / 1e6);
    std::fflush(stdout);
  };
  using I = std::integral_constant<int, 0>;
  (void)sizeof(I);
#define RR(B, M, SPAN, BPS) run_read(std::integral_constant<int, B>{}, std::integral_constant<int, M>{}, SPAN, BPS)
  for (int bps : {4, 8}) {
    RR(16, 1, bytes, bps);
    RR(16, 4, bytes, bps);
    RR(32, 1, bytes, bps);
    RR(32, 2, bytes, bps);
    RR(32, 4, bytes, bps);
    RR(32, 8, bytes, bps);
    RR(64, 4, bytes, bps);
    RR(128, 4, bytes, bps);
  }
  // footprint sweep at 32 B, MLP 4 (L2-resident -> TLB reach -> full)
  for (std::uint64_t mib : {8ull, 64ull, 128ull, 256ull, 512ull, 1024ull, 2048ull}) {
    if ((mib << 20) <= bytes) { RR(32, 4, mib << 20, 8); }
  }

  // latency probe: one dependent chain per thread, few threads
  {
    int const iters = 2000;
    float ms = time_ms([&] { chase<<<1, 32>>>(buf, bytes / 32, iters, sink); }, 2);
    std::printf("{\"test\": \"dependent_chain\", \"ns_per_access\": %.1f}\n", ms * 1e6 / iters);
  }

  auto run_cas = [&](auto mlp_tag, auto wide_tag, bool succeed, int blocks_per_sm) {
    constexpr int M  = decltype(mlp_tag)::value;
    constexpr bool W = decltype(wide_tag)::value;
    int const grid   = sms * blocks_per_sm;
    std::uint64_t const threads = (std::uint64_t)grid * block;
    int const iters  = (int)((1ull << 27) / (threads * M)) + 1;  // ~134 M CAS over bytes/16 slots
    CK(cudaMemset(buf, 0xff, bytes));
    std::uint64_t const expect = succeed ? ~0ull : 0x1234ull;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    random_cas<M, W><<<grid, block>>>(buf, bytes / 16, iters, expect, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    double const ops = (double)threads * M * iters;
    std::printf("{\"test\": \"random_cas\", \"bits\": %d, \"mlp\": %d, \"blocks_per_sm\": %d, \"mostly_succeed\": %s, "
                "\"ms\": %.3f, \"Gops_s\": %.2f}\n",
                W ? 128 : 64, M, blocks_per_sm, succeed ? "true" : "false", ms, ops / ms / 1e6);
    std::fflush(stdout);
  };
#define RC(M, W, S, BPS) run_cas(std::integral_constant<int, M>{}, std::bool_constant<W>{}, S, BPS)
  for (int bps : {4, 8}) {
    RC(1, false, true, bps);
    RC(1, true, true, bps);
    RC(2, false, true, bps);
    RC(2, true, true, bps);
    RC(4, false, true, bps);
    RC(4, true, true, bps);
    RC(4, false, false, bps);
    RC(4, true, false, bps);
  }

  auto run_insert = [&](auto mode_tag, int blocks_per_sm) {
    constexpr int MODE = decltype(mode_tag)::value;
    int const grid     = sms * blocks_per_sm;
    std::uint64_t const threads = (std::uint64_t)grid * block;
    int const iters    = (int)((1ull << 27) / threads) + 1;
    CK(cudaMemset(buf, 0xff, bytes));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    insert_like<MODE><<<grid, block>>>(buf, bytes / 16, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    double const ops = (double)threads * iters;
    char const* names[] = {"load+cas64+cas64", "load+cas128", "cas128_first"};
    std::printf("{\"test\": \"insert_like\", \"mode\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"Gops_s\": %.2f}\n",
                names[MODE], blocks_per_sm, ms, ops / ms / 1e6);
    std::fflush(stdout);
  };
  for (int bps : {4, 8}) {
    run_insert(std::integral_constant<int, 0>{}, bps);
    run_insert(std::integral_constant<int, 1>{}, bps);
    run_insert(std::integral_constant<int, 2>{}, bps);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
