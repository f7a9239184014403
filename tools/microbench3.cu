// Third hardware probe (measurement tool, not product code): why does "load a chunk, then CAS the
// slot" move more DRAM sectors than cuco's "load the key, CAS key, CAS payload"? Times (and, under
// ncu, counts DRAM sectors of) random insert-like sequences over a table much larger than L2 with
//   load width   8 / 16 / 32 bytes, weak / relaxed.gpu / non-coherent
//   claim        one 128-bit CAS, or 64-bit CAS + 64-bit CAS, or 64-bit CAS + plain payload store
// Kernel names carry the variant so an ncu launch list is self-describing.
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

enum { LD_WEAK = 0, LD_RELAXED = 1, LD_NC = 2, LD_EVICT_LAST = 3, LD_EVICT_FIRST = 4 };
enum { CLAIM_CAS128 = 0, CLAIM_CAS64X2 = 1, CLAIM_CAS64_ST = 2, CLAIM_NONE = 3 };

template <int BYTES, int KIND>
__device__ inline std::uint64_t load_first_word(char const* slot)
{
  // loads the BYTES-aligned chunk holding `slot` and returns the slot's first 8 bytes
  char const* p = reinterpret_cast<char const*>(reinterpret_cast<std::uintptr_t>(slot) & ~std::uintptr_t(BYTES - 1));
  unsigned long long w[4] = {0, 0, 0, 0};
#define LD(PFX)                                                                                                    \
  if constexpr (BYTES == 8) {                                                                                     \
    asm volatile(PFX ".u64 %0, [%1];" : "=l"(w[0]) : "l"(p) : "memory");                                          \
  } else if constexpr (BYTES == 16) {                                                                             \
    asm volatile(PFX ".v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");                      \
  } else {                                                                                                        \
    asm volatile(PFX ".v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p) : "memory"); \
  }
  if constexpr (KIND == LD_WEAK) {
    LD("ld.global.L1::no_allocate")
  } else if constexpr (KIND == LD_RELAXED) {
    LD("ld.relaxed.gpu.global.L1::no_allocate")
  } else if constexpr (KIND == LD_NC) {
    LD("ld.global.nc.L1::no_allocate")
  } else {
    // L2 eviction-priority hint on the table line: does the CAS that follows still find it in L2?
    unsigned long long policy;
    if constexpr (KIND == LD_EVICT_LAST) {
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    } else {
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    }
    if constexpr (BYTES == 8) {
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(w[0]) : "l"(p), "l"(policy) : "memory");
    } else if constexpr (BYTES == 16) {
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(w[0]), "=l"(w[1]) : "l"(p), "l"(policy) : "memory");
    } else {
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                   : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p), "l"(policy) : "memory");
    }
  }
#undef LD
  int const idx = BYTES == 32 ? (int)((reinterpret_cast<std::uintptr_t>(slot) & 16) ? 2 : 0) : 0;
  return idx == 0 ? w[0] : w[2];
}

__device__ inline void cas128(char* p, std::uint64_t e0, std::uint64_t e1, std::uint64_t d0, std::uint64_t d1,
                              std::uint64_t& o0, std::uint64_t& o1)
{
  asm volatile(
    "{\n\t.reg .b128 e, d, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
    "atom.relaxed.gpu.global.cas.b128 o, [%6], e, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
    : "=l"(o0), "=l"(o1)
    : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "l"(p)
    : "memory");
}

constexpr int kBlock = 256;

// one op per thread per iteration; grid-stride so the launch shape matches the product kernels
template <int LOAD_BYTES, int LOAD_KIND, int CLAIM, int KPT>
__global__ void __launch_bounds__(kBlock) insert_seq(char* table, std::uint64_t n_slots, std::uint64_t n_ops, std::uint64_t* sink)
{
  std::uint64_t acc = 0;
  std::uint64_t const tile = std::uint64_t{kBlock} * KPT;
  for (std::uint64_t base = blockIdx.x * tile; base < n_ops; base += gridDim.x * tile) {
    char* slot[KPT];
    std::uint64_t seen[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      std::uint64_t const i = base + std::uint64_t{j} * kBlock + threadIdx.x;
      std::uint64_t const h = mix64(i + 0x51ed27ull);
      slot[j]               = table + __umul64hi(h, n_slots) * 16;
      seen[j]               = ~0ull;
      if constexpr (LOAD_BYTES != 0) { seen[j] = load_first_word<LOAD_BYTES ? LOAD_BYTES : 8, LOAD_KIND>(slot[j]); }
    }
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      std::uint64_t const i = base + std::uint64_t{j} * kBlock + threadIdx.x;
      if (i >= n_ops) { continue; }
      std::uint64_t const key = mix64(i) | 1;
      if constexpr (CLAIM == CLAIM_CAS128) {
        std::uint64_t o0, o1;
        cas128(slot[j], seen[j], ~0ull, key, i, o0, o1);
        acc ^= o0 ^ o1;
      } else if constexpr (CLAIM == CLAIM_CAS64X2) {
        auto* p = reinterpret_cast<unsigned long long*>(slot[j]);
        acc ^= atomicCAS(p, (unsigned long long)seen[j], (unsigned long long)key);
        acc ^= atomicCAS(p + 1, ~0ull, (unsigned long long)i);
      } else if constexpr (CLAIM == CLAIM_CAS64_ST) {
        auto* p = reinterpret_cast<unsigned long long*>(slot[j]);
        acc ^= atomicCAS(p, (unsigned long long)seen[j], (unsigned long long)key);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + 1), "l"(i) : "memory");
      } else {
        acc ^= seen[j];
      }
    }
  }
  if (acc == 0x123456789abcdefull) { *sink = acc; }
}

template <int LB, int LK, int CL, int KPT>
void run(char* buf, std::uint64_t bytes, std::uint64_t n_ops, std::uint64_t* sink, int sms, char const* name)
{
  float best = 1e30f;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(buf, 0xff, bytes));
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    std::uint64_t const tiles = (n_ops + kBlock * KPT - 1) / (kBlock * KPT);
    insert_seq<LB, LK, CL, KPT><<<(unsigned)tiles, kBlock>>>(buf, bytes / 16, n_ops, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    best = ms < best ? ms : best;
  }
  std::printf("{\"test\": \"insert_seq\", \"variant\": \"%s\", \"load_bytes\": %d, \"kpt\": %d, \"ms\": %.3f, \"Gops_s\": %.2f}\n",
              name, LB, KPT, best, n_ops / best / 1e6);
  std::fflush(stdout);
}

int main(int argc, char** argv)
{
  std::uint64_t const gib   = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 3;
  std::uint64_t const n_ops = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 100'000'000ull;
  std::uint64_t const bytes = gib << 30;
  char* buf;
  std::uint64_t* sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 8));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  run<8, LD_WEAK, CLAIM_NONE, 1>(buf, bytes, n_ops, sink, sms, "ld8 only");
  run<32, LD_WEAK, CLAIM_NONE, 1>(buf, bytes, n_ops, sink, sms, "ld32 only");
  run<0, LD_WEAK, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "cas128 only");
  run<8, LD_WEAK, CLAIM_CAS64X2, 1>(buf, bytes, n_ops, sink, sms, "ld8 + cas64 + cas64 (cuco)");
  run<8, LD_WEAK, CLAIM_CAS64_ST, 1>(buf, bytes, n_ops, sink, sms, "ld8 + cas64 + st64");
  run<8, LD_WEAK, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld8 + cas128");
  run<16, LD_WEAK, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld16 + cas128");
  run<32, LD_WEAK, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld32 + cas128");
  run<32, LD_RELAXED, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld32.relaxed + cas128");
  run<32, LD_NC, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld32.nc + cas128");
  run<16, LD_RELAXED, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld16.relaxed + cas128");
  run<32, LD_WEAK, CLAIM_CAS64X2, 1>(buf, bytes, n_ops, sink, sms, "ld32 + cas64 + cas64");
  run<32, LD_WEAK, CLAIM_CAS128, 2>(buf, bytes, n_ops, sink, sms, "ld32 + cas128, 2 keys/thread");
  run<8, LD_WEAK, CLAIM_CAS64X2, 2>(buf, bytes, n_ops, sink, sms, "ld8 + cas64 + cas64, 2 keys/thread");
  run<32, LD_WEAK, CLAIM_CAS128, 4>(buf, bytes, n_ops, sink, sms, "ld32 + cas128, 4 keys/thread");
  run<32, LD_EVICT_LAST, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld32.evict_last + cas128");
  run<32, LD_EVICT_LAST, CLAIM_CAS128, 2>(buf, bytes, n_ops, sink, sms, "ld32.evict_last + cas128, 2 keys/thread");
  run<32, LD_EVICT_LAST, CLAIM_CAS128, 4>(buf, bytes, n_ops, sink, sms, "ld32.evict_last + cas128, 4 keys/thread");
  run<16, LD_EVICT_LAST, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld16.evict_last + cas128");
  run<8, LD_EVICT_LAST, CLAIM_CAS64X2, 1>(buf, bytes, n_ops, sink, sms, "ld8.evict_last + cas64 + cas64");
  run<32, LD_EVICT_FIRST, CLAIM_CAS128, 1>(buf, bytes, n_ops, sink, sms, "ld32.evict_first + cas128");
  run<32, LD_EVICT_LAST, CLAIM_NONE, 1>(buf, bytes, n_ops, sink, sms, "ld32.evict_last only");
  run<32, LD_EVICT_FIRST, CLAIM_NONE, 1>(buf, bytes, n_ops, sink, sms, "ld32.evict_first only");
  CK(cudaDeviceSynchronize());
  return 0;
}
