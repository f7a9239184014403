// Does cudaLimitMaxL2FetchGranularity change the cost of a random 32-byte read on B200?
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); std::exit(1); } } while (0)
__host__ __device__ inline std::uint64_t mix64(std::uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int MODE>  // 0: nc 32B, 1: nc 16B + L2::64B, 2: nc 16B + L2::128B, 3: nc 16B, 4: plain 8B
__global__ void __launch_bounds__(256) random_read(char const* buf, std::uint64_t n_units, int iters, std::uint64_t* sink)
{
  std::uint64_t const tid = blockIdx.x * (std::uint64_t)blockDim.x + threadIdx.x;
  std::uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    std::uint64_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      std::uint64_t const h = mix64(tid * 1000003ull + (std::uint64_t)it * 4 + j);
      char const* p = buf + __umul64hi(h, n_units) * 32;
      unsigned long long a = 0, b = 0, c = 0, d = 0;
      if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
      if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
      if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
      if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
      if (MODE == 4) a = *reinterpret_cast<unsigned long long const*>(p);
      v[j] = a ^ b ^ c ^ d;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc ^= v[j];
  }
  if (acc == 0x123456789abcdefull) *sink = acc;
}

template <int MODE>
void run(char const* name, char* buf, std::uint64_t bytes, std::uint64_t* sink, int sms)
{
  int const grid = sms * 8, block = 256;
  std::uint64_t const threads = (std::uint64_t)grid * block;
  int const iters = (int)((1ull << 28) / (threads * 4)) + 1;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  random_read<MODE><<<grid, block>>>(buf, bytes / 32, iters, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  random_read<MODE><<<grid, block>>>(buf, bytes / 32, iters, sink);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  std::printf("  {\"load\": \"%s\", \"Gacc_s\": %.2f}\n", name, (double)threads * 4 * iters / ms / 1e6);
}

int main()
{
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  std::uint64_t const bytes = 4ull << 30;
  for (int gran : {0, 32, 64, 128}) {
    if (gran) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
      std::printf("set granularity %d -> %s\n", gran, cudaGetErrorString(e));
    }
    size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
    std::printf("granularity limit now %zu\n", cur);
    char* buf; std::uint64_t* sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(buf, 1, bytes));
    run<0>("nc 32B", buf, bytes, sink, sms);
    run<3>("nc 16B", buf, bytes, sink, sms);
    run<1>("nc 16B L2::64B", buf, bytes, sink, sms);
    run<2>("nc 16B L2::128B", buf, bytes, sink, sms);
    run<4>("plain 8B", buf, bytes, sink, sms);
    CK(cudaFree(buf)); CK(cudaFree(sink));
  }
  return 0;
}
