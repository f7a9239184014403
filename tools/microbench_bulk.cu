// Does the fetch granularity of a random 32-byte read depend on HOW it is issued? (round 2)
// profiles/r01_hardware_probes.md: an ordinary 32-byte load that misses L2 moves a whole 128-byte line
// from DRAM (3.9 sectors per read), whatever cudaLimitMaxL2FetchGranularity or the L2::64B / L2::128B
// qualifiers say - which is what pins find / contains at 0.31 of the sector roofline. This probe issues the
// same random 32-byte reads four ways and lets ncu count dram__sectors_read per launch:
//   0  ld.global.nc.L1::no_allocate 256-bit          (what lookup_kernel does)
//   1  cp.async.bulk.shared.global, 32 bytes per thread, one mbarrier per CTA round (bulk-copy engine)
//   2  cp.async.ca.shared.global 16 bytes x 2         (LDGSTS)
//   3  ld.global.nc 256-bit with an L2 evict_first cache-hint policy
//   4  ld.relaxed.gpu 256-bit   5  ld.volatile 2 x 128-bit   6  ld.global.cg 2 x 128-bit
//   7  two 128-bit compare-and-swaps that never succeed (atomic reads)
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I include tools/microbench_bulk.cu
#include <cuco/b200/stream_kernels.cuh>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace ptx = cuco::b200::ptx;

__device__ __forceinline__ std::uint64_t mix(std::uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

template <int Mode>
__global__ void __launch_bounds__(256) probe(char const* table, std::uint64_t sectors, std::uint64_t reads,
                                             unsigned long long* sink)
{
  __shared__ __align__(128) char landing[256 * 32];
  __shared__ __align__(8) std::uint64_t arrived;
  if (threadIdx.x == 0) {
    ptx::mbarrier_init(&arrived, 256);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  unsigned long long acc = 0;
  unsigned round         = 0;
  std::uint64_t policy   = 0;
  if (Mode == 3) { policy = ptx::policy_evict_first(); }
  for (std::uint64_t i = std::uint64_t{blockIdx.x} * 256 + threadIdx.x; i < reads;
       i += std::uint64_t{gridDim.x} * 256, ++round) {
    char const* p = table + (mix(i) % sectors) * 32;
    if (Mode == 0) {
      std::uint64_t a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 3) {
      std::uint64_t a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                   : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                   : "l"(p), "l"(policy));
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 4) {
      std::uint64_t a, b, c, d;
      asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 5) {
      std::uint64_t a, b, c, d;
      asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
      asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16) : "memory");
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 6) {
      std::uint64_t a, b, c, d;
      asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
      asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16));
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 7) {
      // atomic reads: two 128-bit compare-and-swaps that can never succeed
      std::uint64_t a, b, c, d;
      asm volatile(
        "{\n\t.reg .b128 cmp, val, old;\n\t"
        "mov.b128 cmp, {%4, %4};\n\tmov.b128 val, {%4, %4};\n\t"
        "atom.relaxed.gpu.global.cas.b128 old, [%5], cmp, val;\n\t"
        "mov.b128 {%0, %1}, old;\n\t"
        "atom.relaxed.gpu.global.cas.b128 old, [%5+16], cmp, val;\n\t"
        "mov.b128 {%2, %3}, old;\n\t}"
        : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
        : "l"(0x7ff1e2d3c4b5a697ull), "l"(p)
        : "memory");
      acc += a ^ b ^ c ^ d;
    } else if (Mode == 1) {
      // every thread announces and issues its own 32-byte bulk copy; the CTA waits for all of them
      ptx::mbarrier_arrive_expect(&arrived, 32);
      ptx::bulk_load(landing + threadIdx.x * 32, p, 32, &arrived);
      ptx::mbarrier_wait(&arrived, round & 1);
      auto const* q = reinterpret_cast<std::uint64_t const*>(landing + threadIdx.x * 32);
      acc += q[0] ^ q[1] ^ q[2] ^ q[3];
      __syncthreads();  // nobody re-arms the barrier before everyone has seen this phase
    } else {
      unsigned const dst = ptx::shared_address(landing + threadIdx.x * 32);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(p + 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      auto const* q = reinterpret_cast<std::uint64_t const*>(landing + threadIdx.x * 32);
      acc += q[0] ^ q[1] ^ q[2] ^ q[3];
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

template <int Mode>
static void run(char const* name, char const* table, std::uint64_t sectors, std::uint64_t reads, unsigned long long* sink)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  int const grid = 148 * 8;
  probe<Mode><<<grid, 256>>>(table, sectors, reads, sink);  // warm-up
  cudaEventRecord(a);
  probe<Mode><<<grid, 256>>>(table, sectors, reads, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  std::printf("{\"test\": \"random 32 B reads\", \"how\": \"%s\", \"reads\": %llu, \"span_gib\": %.1f, \"ms\": %.3f, "
              "\"greads_per_s\": %.2f, \"cuda\": \"%s\"}\n",
              name, (unsigned long long)reads, sectors * 32.0 / (1ull << 30), ms, reads / ms / 1e6,
              cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
  std::uint64_t const bytes = (argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4096ull) << 20;
  std::uint64_t reads = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 100000000ull;
  reads -= reads % (148ull * 8 * 256);  // every thread runs the same number of rounds (mode 1 has a CTA-wide barrier per round)
  char* table;
  unsigned long long* sink;
  cudaMalloc(&table, bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(table, 1, bytes);
  run<0>("ld.global.nc 256-bit", table, bytes / 32, reads, sink);
  run<3>("ld.global.nc 256-bit + L2 evict_first hint", table, bytes / 32, reads, sink);
  run<2>("cp.async.ca 2 x 16 B (LDGSTS)", table, bytes / 32, reads, sink);
  run<1>("cp.async.bulk 32 B per thread", table, bytes / 32, reads, sink);
  run<4>("ld.relaxed.gpu 256-bit", table, bytes / 32, reads, sink);
  run<5>("ld.volatile 2 x 128-bit", table, bytes / 32, reads, sink);
  run<6>("ld.global.cg 2 x 128-bit", table, bytes / 32, reads, sink);
  run<7>("atom.cas.b128 x 2 (never succeeds)", table, bytes / 32, reads, sink);
  cudaDeviceSynchronize();
  return 0;
}
