// extern "C" entry points of include/cuco_b200.h: argument checking, exception -> status code
// translation, dispatch to the per-kind tables (cabi_kind.cu), launch tuning, and the routing
// kernels of the hash-partitioned multi-GPU table. Compiled into both libcuco_b200.so (native) and
// oracle/_ref/libcuco_ref.so (reference build of the same shim; -DCUCO_SHIM_REFERENCE).
#include "cabi_table.hpp"

#include "../../include/cuco_b200.h"

#if !defined(CUCO_SHIM_REFERENCE)
#include <cuco/b200/bulk_engine.cuh>
#endif

#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstring>
#include <exception>
#include <new>
#include <stdexcept>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) noexcept
{
  try {
    f();
    g_last_error.clear();
    return 0;
  } catch (std::bad_alloc const& e) {
    g_last_error = std::string{"out of memory: "} + e.what();
    return 2;
  } catch (std::logic_error const& e) {  // cuco::logic_error, std::invalid_argument
    g_last_error = e.what();
    return 3;
  } catch (std::exception const& e) {  // cuco::cuda_error and anything else
    g_last_error = e.what();
    return 4;
  } catch (...) {
    g_last_error = "unknown exception";
    return 5;
  }
}

void require(bool ok, char const* what)
{
  if (!ok) { throw std::invalid_argument(what); }
}

cuco_shim_factory const g_factories[CUCO_B200_NUM_KINDS] = {
  cuco_shim_make_kind_0, cuco_shim_make_kind_1, cuco_shim_make_kind_2, cuco_shim_make_kind_3,
  cuco_shim_make_kind_4, cuco_shim_make_kind_5, cuco_shim_make_kind_6, cuco_shim_make_kind_7,
  cuco_shim_make_kind_8, cuco_shim_make_kind_9, cuco_shim_make_kind_10, cuco_shim_make_kind_11,
  cuco_shim_make_kind_12, cuco_shim_make_kind_13};

void check_launch()
{
  auto const status = cudaPeekAtLastError();
  if (status != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error(std::string{"kernel launch failed: "} + cudaGetErrorString(status));
  }
}

// ================================================================================================
// routing kernels for the hash-partitioned table
// ================================================================================================
constexpr int route_block = 256;
constexpr int route_items = 8;  // elements per thread per tile
constexpr int max_parts   = 64;

__host__ __device__ inline std::uint64_t mix64(std::uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

/// Owner rank of a key: high bits of a 64-bit mix that shares nothing with the in-table xxhash.
__device__ inline int owner_of(std::uint64_t key, std::uint64_t salt, int num_parts)
{
  return static_cast<int>(__umul64hi(mix64(key ^ salt), static_cast<std::uint64_t>(num_parts)));
}

template <typename Key>
__device__ inline std::uint64_t key_bits(Key k)
{
  if constexpr (sizeof(Key) == 4) {
    return static_cast<std::uint64_t>(static_cast<std::uint32_t>(k));
  } else {
    return static_cast<std::uint64_t>(k);
  }
}

template <typename Key, int Stride>
__global__ __launch_bounds__(route_block) void partition_count_kernel(
  Key const* keys, std::int64_t n, int num_parts, std::uint64_t salt, unsigned long long* counts)
{
  __shared__ unsigned int hist[max_parts];
  for (int p = threadIdx.x; p < num_parts; p += route_block) {
    hist[p] = 0;
  }
  __syncthreads();

  constexpr std::int64_t tile = std::int64_t{route_block} * route_items;
  for (std::int64_t base = std::int64_t{blockIdx.x} * tile; base < n;
       base += std::int64_t{gridDim.x} * tile) {
#pragma unroll
    for (int j = 0; j < route_items; ++j) {
      std::int64_t const i = base + std::int64_t{j} * route_block + threadIdx.x;
      bool const live      = i < n;
      int const owner      = live ? owner_of(key_bits(keys[i * Stride]), salt, num_parts) : -1;
      // one shared-memory add per distinct owner per warp
      unsigned const peers = __match_any_sync(0xffffffffu, owner);
      if (live && (__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31)) {
        atomicAdd(&hist[owner], __popc(peers));
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < num_parts; p += route_block) {
    if (hist[p]) { atomicAdd(&counts[p], static_cast<unsigned long long>(hist[p])); }
  }
}

/// Element = Key (+ Payload). AoS: one struct {Key, Payload}; SoA: two arrays.
template <typename Key, typename Payload, bool HasPayload, bool AoS>
__global__ __launch_bounds__(route_block) void partition_scatter_kernel(Key const* keys,
                                                                        Payload const* values,
                                                                        std::int64_t n,
                                                                        int num_parts,
                                                                        std::uint64_t salt,
                                                                        unsigned long long* cursors,
                                                                        Key* keys_out,
                                                                        Payload* values_out,
                                                                        std::int64_t* src_index)
{
  __shared__ unsigned int hist[max_parts];
  __shared__ unsigned long long seg_base[max_parts];
  constexpr int stride      = AoS ? 2 : 1;
  constexpr std::int64_t tile = std::int64_t{route_block} * route_items;

  for (std::int64_t base = std::int64_t{blockIdx.x} * tile; base < n;
       base += std::int64_t{gridDim.x} * tile) {
    for (int p = threadIdx.x; p < num_parts; p += route_block) {
      hist[p] = 0;
    }
    __syncthreads();

    Key k[route_items];
    Payload v[route_items];
    int owner[route_items];
    unsigned rank[route_items];
#pragma unroll
    for (int j = 0; j < route_items; ++j) {
      std::int64_t const i = base + std::int64_t{j} * route_block + threadIdx.x;
      bool const live      = i < n;
      owner[j]             = -1;
      if (live) {
        k[j] = keys[i * stride];
        if constexpr (HasPayload) { v[j] = AoS ? reinterpret_cast<Payload const*>(keys)[i * 2 + 1] : values[i]; }
        owner[j] = owner_of(key_bits(k[j]), salt, num_parts);
      }
      unsigned const peers  = __match_any_sync(0xffffffffu, owner[j]);
      int const leader      = __ffs(peers) - 1;
      unsigned const before = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
      unsigned start        = 0;
      if (live && leader == static_cast<int>(threadIdx.x & 31)) {
        start = atomicAdd(&hist[owner[j]], __popc(peers));
      }
      start   = __shfl_sync(0xffffffffu, start, leader);
      rank[j] = start + before;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < num_parts; p += route_block) {
      seg_base[p] = hist[p] ? atomicAdd(&cursors[p], static_cast<unsigned long long>(hist[p])) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < route_items; ++j) {
      if (owner[j] >= 0) {
        auto const dst = seg_base[owner[j]] + rank[j];
        if constexpr (AoS) {
          keys_out[dst * 2]                                  = k[j];
          reinterpret_cast<Payload*>(keys_out)[dst * 2 + 1] = v[j];
        } else {
          keys_out[dst] = k[j];
          if constexpr (HasPayload) { values_out[dst] = v[j]; }
        }
        if (src_index != nullptr) {
          src_index[dst] = base + std::int64_t{j} * route_block + threadIdx.x;
        }
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ __launch_bounds__(route_block) void scatter_by_index_kernel(T const* in,
                                                                       std::int64_t const* index,
                                                                       T* out,
                                                                       std::int64_t n)
{
  for (std::int64_t i = std::int64_t{blockIdx.x} * route_block + threadIdx.x; i < n;
       i += std::int64_t{gridDim.x} * route_block) {
    out[index[i]] = in[i];
  }
}

unsigned route_grid(std::int64_t n, std::int64_t per_block)
{
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  std::int64_t const blocks = (n + per_block - 1) / per_block;
  std::int64_t const cap    = std::int64_t{sms} * 8;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}


// ================================================================================================
// host-buffer pipeline: H2D of chunk i+1, table kernels on chunk i and D2H of chunk i-1 overlap
// ================================================================================================
// PCIe is full duplex and the copy engines run beside the SMs, so a bulk call on host buffers costs
// max(upload, kernels, download) instead of their sum when it is cut into chunks. Uploads go on one
// internal stream, downloads on another, the table kernels stay on the caller's stream; events tie
// them together so the whole call remains ordered on the caller's stream.
void cuda_ok(cudaError_t status, char const* what)
{
  if (status != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error(std::string{what} + ": " + cudaGetErrorString(status));
  }
}

class host_pipeline {
 public:
  static constexpr int depth = 3;

  static host_pipeline& for_current_device()
  {
    static std::mutex guard;
    static std::map<int, std::unique_ptr<host_pipeline>> per_device;
    int device = 0;
    cuda_ok(cudaGetDevice(&device), "cudaGetDevice");
    std::lock_guard<std::mutex> lock{guard};
    auto& slot = per_device[device];
    if (!slot) { slot.reset(new host_pipeline{}); }
    return *slot;
  }

  std::mutex mutex;  ///< one host-buffer call at a time per device (they share the staging buffers)
  cudaStream_t upload{}, download{};
  cudaEvent_t forked{}, uploaded[depth]{}, consumed[depth]{}, produced[depth]{}, downloaded[depth]{};
  char* in[depth]{};
  char* out[depth]{};
  std::size_t in_bytes{0}, out_bytes{0};

  void reserve(std::size_t in_need, std::size_t out_need)
  {
    if (in_need > in_bytes) {
      cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
      for (auto& b : in) {
        if (b) { cudaFree(b); }
        b = nullptr;
        cuda_ok(cudaMalloc(reinterpret_cast<void**>(&b), in_need), "cudaMalloc(host pipeline input)");
      }
      in_bytes = in_need;
    }
    if (out_need > out_bytes) {
      cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
      for (auto& b : out) {
        if (b) { cudaFree(b); }
        b = nullptr;
        cuda_ok(cudaMalloc(reinterpret_cast<void**>(&b), out_need), "cudaMalloc(host pipeline output)");
      }
      out_bytes = out_need;
    }
  }

 private:
  host_pipeline()
  {
    cuda_ok(cudaStreamCreateWithFlags(&upload, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_ok(cudaStreamCreateWithFlags(&download, cudaStreamNonBlocking), "cudaStreamCreate");
    auto make = [](cudaEvent_t& e) {
      cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    };
    make(forked);
    for (int i = 0; i < depth; ++i) {
      make(uploaded[i]);
      make(consumed[i]);
      make(produced[i]);
      make(downloaded[i]);
    }
  }
};

std::int64_t host_chunk_elements()
{
  static std::int64_t const value = [] {
    if (char const* s = std::getenv("CUCO_B200_HOST_CHUNK")) {
      auto const v = std::atoll(s);
      if (v > 0) { return static_cast<std::int64_t>(v); }
    }
    return std::int64_t{1} << 22;  // 4 Mi elements: 64 MB of int64 pairs, ~1.2 ms of PCIe 5 x16
  }();
  return value;
}

std::size_t align256(std::size_t bytes) { return (bytes + 255) / 256 * 256; }

/// Runs `op(device_keys, device_values, device_out, count)` chunk by chunk over host buffers.
/// in_a / in_b: host arrays with `a_bytes` / `b_bytes` per element (in_b may be null);
/// host_out: host array with `out_elem` bytes per element (may be null: no download).
template <typename Op>
void run_host_pipeline(const void* in_a,
                       std::size_t a_bytes,
                       const void* in_b,
                       std::size_t b_bytes,
                       void* host_out,
                       std::size_t out_elem,
                       void* host_out2,
                       std::size_t out2_elem,
                       std::int64_t n,
                       cudaStream_t stream,
                       Op&& op)
{
  if (n == 0) { return; }
  auto& pipe = host_pipeline::for_current_device();
  std::lock_guard<std::mutex> lock{pipe.mutex};
  std::int64_t const chunk = std::min<std::int64_t>(host_chunk_elements(), n);
  std::size_t const b_off   = align256(static_cast<std::size_t>(chunk) * a_bytes);
  std::size_t const out2_off = align256(static_cast<std::size_t>(chunk) * out_elem);
  pipe.reserve(b_off + align256(static_cast<std::size_t>(chunk) * b_bytes),
               out2_off + align256(static_cast<std::size_t>(chunk) * out2_elem));

  // uploads start after whatever the caller has already queued on `stream`
  cuda_ok(cudaEventRecord(pipe.forked, stream), "cudaEventRecord");
  cuda_ok(cudaStreamWaitEvent(pipe.upload, pipe.forked, 0), "cudaStreamWaitEvent");
  bool const downloads = host_out != nullptr;

  std::int64_t index = 0;
  for (std::int64_t begin = 0; begin < n; begin += chunk, ++index) {
    int const b             = static_cast<int>(index % host_pipeline::depth);
    std::int64_t const count = std::min<std::int64_t>(chunk, n - begin);
    // staging buffer b is free once the kernels of its previous chunk have run
    cuda_ok(cudaStreamWaitEvent(pipe.upload, pipe.consumed[b], 0), "cudaStreamWaitEvent");
    cuda_ok(cudaMemcpyAsync(pipe.in[b],
                            static_cast<char const*>(in_a) + static_cast<std::size_t>(begin) * a_bytes,
                            static_cast<std::size_t>(count) * a_bytes,
                            cudaMemcpyHostToDevice,
                            pipe.upload),
            "cudaMemcpyAsync(H2D)");
    if (in_b != nullptr) {
      cuda_ok(cudaMemcpyAsync(pipe.in[b] + b_off,
                              static_cast<char const*>(in_b) + static_cast<std::size_t>(begin) * b_bytes,
                              static_cast<std::size_t>(count) * b_bytes,
                              cudaMemcpyHostToDevice,
                              pipe.upload),
              "cudaMemcpyAsync(H2D)");
    }
    cuda_ok(cudaEventRecord(pipe.uploaded[b], pipe.upload), "cudaEventRecord");

    cuda_ok(cudaStreamWaitEvent(stream, pipe.uploaded[b], 0), "cudaStreamWaitEvent");
    if (downloads) {
      cuda_ok(cudaStreamWaitEvent(stream, pipe.downloaded[b], 0), "cudaStreamWaitEvent");
    }
    op(pipe.in[b], in_b != nullptr ? pipe.in[b] + b_off : nullptr, pipe.out[b], pipe.out[b] + out2_off, count);
    cuda_ok(cudaEventRecord(pipe.consumed[b], stream), "cudaEventRecord");

    if (downloads) {
      cuda_ok(cudaStreamWaitEvent(pipe.download, pipe.consumed[b], 0), "cudaStreamWaitEvent");
      cuda_ok(cudaMemcpyAsync(static_cast<char*>(host_out) + static_cast<std::size_t>(begin) * out_elem,
                              pipe.out[b],
                              static_cast<std::size_t>(count) * out_elem,
                              cudaMemcpyDeviceToHost,
                              pipe.download),
              "cudaMemcpyAsync(D2H)");
      if (host_out2 != nullptr) {
        cuda_ok(cudaMemcpyAsync(static_cast<char*>(host_out2) + static_cast<std::size_t>(begin) * out2_elem,
                                pipe.out[b] + out2_off,
                                static_cast<std::size_t>(count) * out2_elem,
                                cudaMemcpyDeviceToHost,
                                pipe.download),
                "cudaMemcpyAsync(D2H)");
      }
      cuda_ok(cudaEventRecord(pipe.downloaded[b], pipe.download), "cudaEventRecord");
    }
  }
  if (downloads) {
    // the call is complete on `stream` only when every result has reached the host
    for (int b = 0; b < host_pipeline::depth; ++b) {
      cuda_ok(cudaStreamWaitEvent(stream, pipe.downloaded[b], 0), "cudaStreamWaitEvent");
    }
  }
}

}  // namespace

// everything else in the library is built with -fvisibility=hidden; the C ABI is the export list
#pragma GCC visibility push(default)
extern "C" {

const char* cuco_b200_build_info(void)
{
#if defined(CUCO_SHIM_REFERENCE)
  return "reference";
#else
  return "native";
#endif
}

const char* cuco_b200_last_error(void) { return g_last_error.c_str(); }

int cuco_b200_create(int kind,
                     int64_t size,
                     double load_factor,
                     int64_t empty_key,
                     int64_t empty_value,
                     int has_erased,
                     int64_t erased_key,
                     void* stream,
                     cuco_b200_table** out)
{
  return guarded([&] {
    require(out != nullptr, "out must not be NULL");
    require(kind >= 0 && kind < CUCO_B200_NUM_KINDS, "unknown table kind");
    require(!(has_erased && load_factor != 0.0),
            "the erased-key constructor takes a capacity, not a load factor");
    *out = g_factories[kind](size, load_factor, empty_key, empty_value, has_erased, erased_key, stream);
  });
}

int cuco_b200_destroy(cuco_b200_table* t)
{
  return guarded([&] { delete t; });
}

int cuco_b200_kind_of(const cuco_b200_table* t) { return t ? t->kind() : -1; }
int cuco_b200_key_bytes(const cuco_b200_table* t) { return t ? t->key_bytes() : -1; }
int cuco_b200_value_bytes(const cuco_b200_table* t) { return t ? t->value_bytes() : -1; }
int64_t cuco_b200_capacity(const cuco_b200_table* t) { return t ? t->capacity() : -1; }

int cuco_b200_size(cuco_b200_table* t, void* stream, int64_t* out)
{
  return guarded([&] {
    require(t && out, "NULL argument");
    *out = t->size(stream);
  });
}

int cuco_b200_clear(cuco_b200_table* t, void* stream)
{
  return guarded([&] {
    require(t, "NULL table");
    t->clear(stream);
    check_launch();
  });
}

int cuco_b200_insert(cuco_b200_table* t,
                     const void* keys,
                     const void* values,
                     int64_t n,
                     void* stream,
                     int64_t* num_inserted)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0), "bad argument");
    t->insert(keys, values, n, stream, num_inserted);
    check_launch();
  });
}

int cuco_b200_insert_if(cuco_b200_table* t,
                        const void* keys,
                        const void* values,
                        const uint8_t* stencil,
                        int64_t n,
                        void* stream,
                        int64_t* num_inserted)
{
  return guarded([&] {
    require(t && n >= 0 && ((keys && stencil) || n == 0), "bad argument");
    t->insert_if(keys, values, stencil, n, stream, num_inserted);
    check_launch();
  });
}

int cuco_b200_find(cuco_b200_table* t, const void* keys, void* out, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((keys && out) || n == 0), "bad argument");
    t->find(keys, out, n, stream);
    check_launch();
  });
}

int cuco_b200_contains(cuco_b200_table* t, const void* keys, uint8_t* out, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((keys && out) || n == 0), "bad argument");
    t->contains(keys, out, n, stream);
    check_launch();
  });
}

int cuco_b200_contains_if(cuco_b200_table* t,
                          const void* keys,
                          const uint8_t* stencil,
                          uint8_t* out,
                          int64_t n,
                          void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((keys && out && stencil) || n == 0), "bad argument");
    t->contains_if(keys, stencil, out, n, stream);
    check_launch();
  });
}

int cuco_b200_insert_and_find(cuco_b200_table* t,
                              const void* keys,
                              const void* values,
                              void* found,
                              uint8_t* inserted,
                              int64_t n,
                              void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((keys && found && inserted) || n == 0), "bad argument");
    t->insert_and_find(keys, values, found, inserted, n, stream);
    check_launch();
  });
}

int cuco_b200_insert_or_assign(
  cuco_b200_table* t, const void* keys, const void* values, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0), "bad argument");
    t->insert_or_assign(keys, values, n, stream);
    check_launch();
  });
}

int cuco_b200_insert_or_apply(cuco_b200_table* t,
                              const void* keys,
                              const void* values,
                              int64_t n,
                              int reduce_op,
                              int has_init,
                              int64_t init,
                              void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0), "bad argument");
    t->insert_or_apply(keys, values, n, reduce_op, has_init, init, stream);
    check_launch();
  });
}

int cuco_b200_erase(cuco_b200_table* t, const void* keys, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0), "bad argument");
    t->erase(keys, n, stream);
    check_launch();
  });
}

int cuco_b200_retrieve_all(
  cuco_b200_table* t, void* keys_out, void* values_out, int64_t* n_out, void* stream)
{
  return guarded([&] {
    require(t && keys_out && n_out, "NULL argument");
    *n_out = t->retrieve_all(keys_out, values_out, stream);
  });
}

int cuco_b200_count(
  cuco_b200_table* t, const void* keys, int64_t n, int outer, void* stream, int64_t* out)
{
  return guarded([&] {
    require(t && out && n >= 0 && (keys || n == 0), "bad argument");
    *out = t->count(keys, n, outer != 0, stream);
  });
}

int cuco_b200_retrieve(cuco_b200_table* t,
                       const void* keys,
                       int64_t n,
                       int outer,
                       void* probe_out,
                       void* match_out,
                       int64_t* n_out,
                       void* stream)
{
  return guarded([&] {
    require(t && n_out && n >= 0 && ((keys && probe_out && match_out) || n == 0), "bad argument");
    *n_out = t->retrieve(keys, n, outer != 0, probe_out, match_out, stream);
  });
}

int cuco_b200_rehash(cuco_b200_table* t, int64_t capacity, void* stream)
{
  return guarded([&] {
    require(t, "NULL table");
    t->rehash(capacity, stream);
  });
}

int cuco_b200_insert_host(cuco_b200_table* t,
                          const void* host_keys,
                          const void* host_values,
                          int64_t n,
                          void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (host_keys || n == 0), "bad argument");
    auto const kb = static_cast<std::size_t>(t->key_bytes());
    auto const vb = static_cast<std::size_t>(t->value_bytes());
    bool const aos = vb != 0 && host_values == nullptr;
    run_host_pipeline(host_keys, aos ? kb + vb : kb, host_values, vb, nullptr, 0, nullptr, 0, n,
                      static_cast<cudaStream_t>(stream),
                      [&](void* keys, void* values, void*, void*, std::int64_t count) {
                        t->insert(keys, values, count, stream, nullptr);
                        check_launch();
                      });
  });
}

int cuco_b200_find_host(
  cuco_b200_table* t, const void* host_keys, void* host_out, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((host_keys && host_out) || n == 0), "bad argument");
    auto const kb  = static_cast<std::size_t>(t->key_bytes());
    auto const out = static_cast<std::size_t>(t->value_bytes() ? t->value_bytes() : t->key_bytes());
    run_host_pipeline(host_keys, kb, nullptr, 0, host_out, out, nullptr, 0, n,
                      static_cast<cudaStream_t>(stream),
                      [&](void* keys, void*, void* found, void*, std::int64_t count) {
                        t->find(keys, found, count, stream);
                        check_launch();
                      });
  });
}

int cuco_b200_contains_host(
  cuco_b200_table* t, const void* host_keys, uint8_t* host_out, int64_t n, void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((host_keys && host_out) || n == 0), "bad argument");
    auto const kb = static_cast<std::size_t>(t->key_bytes());
    run_host_pipeline(host_keys, kb, nullptr, 0, host_out, 1, nullptr, 0, n,
                      static_cast<cudaStream_t>(stream),
                      [&](void* keys, void*, void* present, void*, std::int64_t count) {
                        t->contains(keys, static_cast<std::uint8_t*>(present), count, stream);
                        check_launch();
                      });
  });
}

int cuco_b200_insert_and_find_host(cuco_b200_table* t,
                                   const void* host_keys,
                                   const void* host_values,
                                   void* host_found,
                                   uint8_t* host_inserted,
                                   int64_t n,
                                   void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((host_keys && host_found && host_inserted) || n == 0), "bad argument");
    auto const kb  = static_cast<std::size_t>(t->key_bytes());
    auto const vb  = static_cast<std::size_t>(t->value_bytes());
    bool const aos = vb != 0 && host_values == nullptr;
    auto const out = vb ? vb : kb;
    run_host_pipeline(host_keys, aos ? kb + vb : kb, host_values, vb, host_found, out, host_inserted, 1, n,
                      static_cast<cudaStream_t>(stream),
                      [&](void* keys, void* values, void* found, void* inserted, std::int64_t count) {
                        t->insert_and_find(keys, values, found, static_cast<std::uint8_t*>(inserted), count, stream);
                        check_launch();
                      });
  });
}

int cuco_b200_exchange_plan(cuco_b200_table* t,
                            int64_t n_max,
                            int num_ranks,
                            uint32_t* num_regions,
                            uint32_t* segment_capacity,
                            uint32_t* spill_capacity)
{
  return guarded([&] {
    require(t && num_regions && segment_capacity && spill_capacity, "NULL argument");
    auto const shape  = t->exchange_plan(n_max, num_ranks);
    *num_regions      = shape.num_regions;
    *segment_capacity = shape.segment_capacity;
    *spill_capacity   = shape.spill_capacity;
  });
}

int cuco_b200_exchange_route(cuco_b200_table* t,
                             const void* keys,
                             const void* values,
                             int64_t n,
                             int keys_only,
                             uint32_t num_regions,
                             uint32_t segment_capacity,
                             uint32_t spill_capacity,
                             int num_ranks,
                             int my_rank,
                             uint64_t salt,
                             void* const* peer_segments,
                             void* const* peer_counts,
                             void* const* peer_flags,
                             void* counts_local,
                             void* position_local,
                             void* spill,
                             void* spill_index,
                             void* spill_count,
                             void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0) && peer_segments && peer_counts && peer_flags &&
              counts_local && spill && spill_count,
            "bad argument");
    require(!keys_only || (position_local && spill_index), "lookups need the position buffers");
    require(my_rank >= 0 && my_rank < num_ranks, "rank out of range");
    t->exchange_route(keys, values, n, keys_only != 0,
                      cuco_b200_table::exchange_shape{num_regions, segment_capacity, spill_capacity},
                      num_ranks, my_rank, salt, peer_segments, peer_counts, peer_flags, counts_local,
                      position_local, spill, spill_index, spill_count, stream);
    check_launch();
  });
}

int cuco_b200_exchange_mutate(cuco_b200_table* t,
                              const void* segments,
                              const void* counts_recv,
                              uint32_t num_regions,
                              uint32_t segment_capacity,
                              int num_ranks,
                              int reduce_op,
                              void* stream)
{
  return guarded([&] {
    require(t && segments && counts_recv, "NULL argument");
    t->exchange_mutate(segments, counts_recv,
                       cuco_b200_table::exchange_shape{num_regions, segment_capacity, 0}, num_ranks,
                       reduce_op, stream);
    check_launch();
  });
}

int cuco_b200_exchange_lookup(cuco_b200_table* t,
                              const void* segments,
                              const void* counts_recv,
                              void* const* peer_results,
                              uint32_t num_regions,
                              uint32_t segment_capacity,
                              int num_ranks,
                              int my_rank,
                              int what,
                              void* stream)
{
  return guarded([&] {
    require(t && segments && counts_recv && peer_results, "NULL argument");
    t->exchange_lookup(segments, counts_recv, peer_results,
                       cuco_b200_table::exchange_shape{num_regions, segment_capacity, 0}, num_ranks,
                       my_rank, what, stream);
    check_launch();
  });
}

int cuco_b200_exchange_stage_plan(cuco_b200_table* t,
                                  int64_t n_max,
                                  int num_ranks,
                                  int slices,
                                  uint32_t* segment_capacity,
                                  uint32_t* spill_capacity)
{
  return guarded([&] {
    require(t && segment_capacity && spill_capacity, "NULL argument");
    auto const shape  = t->stage_plan(n_max, num_ranks, slices);
    *segment_capacity = shape.segment_capacity;
    *spill_capacity   = shape.spill_capacity;
  });
}

int cuco_b200_exchange_stage(cuco_b200_table* t,
                             const void* keys,
                             const void* values,
                             int64_t n,
                             int keys_only,
                             int slices,
                             uint32_t segment_capacity,
                             uint32_t spill_capacity,
                             int num_ranks,
                             int my_rank,
                             uint64_t salt,
                             void* stage,
                             void* counts_local,
                             void* position_local,
                             void* spill,
                             void* spill_index,
                             void* spill_count,
                             void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && (keys || n == 0) && stage && counts_local && spill && spill_count, "bad argument");
    require(!keys_only || (position_local && spill_index), "lookups need the position buffers");
    require(my_rank >= 0 && my_rank < num_ranks && slices >= 1, "rank or slice count out of range");
    t->exchange_stage(keys, values, n, keys_only != 0,
                      cuco_b200_table::exchange_shape{static_cast<uint32_t>(slices), segment_capacity, spill_capacity},
                      num_ranks, my_rank, salt, stage, counts_local, position_local, spill, spill_index,
                      spill_count, stream);
    check_launch();
  });
}

int cuco_b200_exchange_fine_regions(cuco_b200_table* t, int num_ranks, uint32_t* num_regions)
{
  return guarded([&] {
    require(t && num_regions && num_ranks >= 1, "bad argument");
    *num_regions = t->exchange_fine_regions(num_ranks);
  });
}

int cuco_b200_exchange_probe(cuco_b200_table* t,
                             const void* segments,
                             const void* counts_recv,
                             uint32_t num_regions,
                             uint32_t segment_capacity,
                             int num_ranks,
                             uint32_t region_begin,
                             uint32_t region_count,
                             int reduce_op,
                             void* stream)
{
  return guarded([&] {
    require(t && segments && counts_recv, "NULL argument");
    t->exchange_probe(segments, counts_recv, num_regions, segment_capacity, num_ranks, region_begin, region_count,
                      reduce_op, stream);
    check_launch();
  });
}

int cuco_b200_exchange_publish(const void* counts_local,
                               const void* spill_count,
                               void* const* peer_counts,
                               void* const* peer_flags,
                               int slices,
                               uint32_t segment_capacity,
                               int num_ranks,
                               int my_rank,
                               int source_major,
                               void* stream)
{
  return guarded([&] {
#if defined(CUCO_SHIM_REFERENCE)
    (void)counts_local, (void)spill_count, (void)peer_counts, (void)peer_flags, (void)slices;
    (void)segment_capacity, (void)num_ranks, (void)my_rank, (void)stream, (void)source_major;
    throw std::invalid_argument("the exchange path exists in the native build only");
#else
    require(counts_local && spill_count && peer_counts && peer_flags, "NULL argument");
    require(num_ranks >= 1 && num_ranks <= cuco::b200::exchange_max_ranks && my_rank >= 0 && my_rank < num_ranks &&
              slices >= 1,
            "rank or slice count out of range");
    cuco::b200::exchange_peers counts{}, flags{};
    for (int r = 0; r < num_ranks; ++r) {
      counts.base[r] = peer_counts[r];
      flags.base[r]  = peer_flags[r];
    }
    cuco::b200::exchange_geometry const geometry{static_cast<std::uint32_t>(num_ranks),
                                                 static_cast<std::uint32_t>(my_rank),
                                                 static_cast<std::uint32_t>(slices),
                                                 segment_capacity,
                                                 0,
                                                 source_major ? 1u : static_cast<std::uint32_t>(num_ranks),
                                                 source_major ? static_cast<std::uint32_t>(my_rank) * slices
                                                              : static_cast<std::uint32_t>(my_rank)};
    auto const buckets = static_cast<unsigned>(num_ranks) * static_cast<unsigned>(slices);
    cuco::b200::exchange_publish_kernel<<<(buckets + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned int const*>(counts_local), static_cast<unsigned int const*>(spill_count), counts, flags,
      geometry);
    check_launch();
#endif
  });
}

#if !defined(CUCO_SHIM_REFERENCE)
namespace {
constexpr int push_max_copies = 16;
struct push_job {
  uint4* dst[push_max_copies];
  uint4 const* src[push_max_copies];
  unsigned long long vectors[push_max_copies];  // 16-byte units per copy
};

/// blockIdx.y = copy, the CTAs of a row stride over that copy: plain 16-byte loads and stores, four in
/// flight per thread. The destinations may be peer mappings (NVLink): a handful of CTAs keeps the links
/// busy, and one launch replaces the fixed cost of a cudaMemcpyAsync per peer.
__global__ void __launch_bounds__(256) push_kernel(push_job job)
{
  auto const* src       = job.src[blockIdx.y];
  auto* dst             = job.dst[blockIdx.y];
  auto const n          = job.vectors[blockIdx.y];
  auto const stride     = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  unsigned long long i  = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 const a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i]              = a;
    dst[i + stride]     = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) {
    dst[i] = src[i];
  }
}
}  // namespace
#endif

int cuco_b200_push_async(void* const* dst,
                         const void* const* src,
                         const int64_t* bytes,
                         int num_copies,
                         int ctas_per_copy,
                         void* stream)
{
  return guarded([&] {
#if defined(CUCO_SHIM_REFERENCE)
    (void)dst, (void)src, (void)bytes, (void)num_copies, (void)ctas_per_copy, (void)stream;
    throw std::invalid_argument("the exchange path exists in the native build only");
#else
    require(dst && src && bytes && num_copies >= 0 && num_copies <= push_max_copies && ctas_per_copy >= 1,
            "bad argument");
    if (num_copies == 0) { return; }
    push_job job{};
    for (int i = 0; i < num_copies; ++i) {
      require(bytes[i] >= 0 && bytes[i] % 16 == 0 && reinterpret_cast<std::uintptr_t>(dst[i]) % 16 == 0 &&
                reinterpret_cast<std::uintptr_t>(src[i]) % 16 == 0,
              "copies must be 16-byte aligned multiples of 16 bytes");
      job.dst[i]     = static_cast<uint4*>(dst[i]);
      job.src[i]     = static_cast<uint4 const*>(src[i]);
      job.vectors[i] = static_cast<unsigned long long>(bytes[i]) / 16;
    }
    push_kernel<<<dim3{static_cast<unsigned>(ctas_per_copy), static_cast<unsigned>(num_copies)}, 256, 0,
                  static_cast<cudaStream_t>(stream)>>>(job);
    check_launch();
#endif
  });
}

int cuco_b200_copy_async(void* dst, const void* src, int64_t bytes, void* stream)
{
  return guarded([&] {
    require(bytes >= 0 && ((dst && src) || bytes == 0), "bad argument");
    if (bytes == 0) { return; }
    auto const status =
      cudaMemcpyAsync(dst, src, static_cast<std::size_t>(bytes), cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
    if (status != cudaSuccess) { throw std::runtime_error(cudaGetErrorString(status)); }
  });
}

int cuco_b200_exchange_apply(cuco_b200_table* t,
                             const void* segments,
                             const void* counts_recv,
                             uint32_t segment_capacity,
                             int num_ranks,
                             int slice,
                             int slices,
                             int reduce_op,
                             void* stream)
{
  return guarded([&] {
    require(t && segments && counts_recv, "NULL argument");
    require(slice >= 0 && slice < slices && num_ranks >= 1, "slice out of range");
    t->exchange_apply(segments, counts_recv, segment_capacity, num_ranks, slice, slices, reduce_op, stream);
    check_launch();
  });
}

int cuco_b200_exchange_lookup_local(cuco_b200_table* t,
                                    const void* segments,
                                    const void* counts_recv,
                                    void* results,
                                    uint32_t segment_capacity,
                                    int num_ranks,
                                    int what,
                                    void* stream)
{
  return guarded([&] {
    require(t && segments && counts_recv && results, "NULL argument");
    t->exchange_lookup_local(segments, counts_recv, results, segment_capacity, num_ranks, what, stream);
    check_launch();
  });
}

int cuco_b200_exchange_unpermute(cuco_b200_table* t,
                                 const void* results,
                                 const void* position_local,
                                 int64_t n,
                                 void* out,
                                 int what,
                                 void* stream)
{
  return guarded([&] {
    require(t && n >= 0 && ((results && position_local && out) || n == 0), "bad argument");
    t->exchange_unpermute(results, position_local, n, out, what, stream);
    check_launch();
  });
}

int cuco_b200_set_tuning(int keys_per_thread,
                         int cas_first,
                         int sector_chunks,
                         int waves,
                         int force_generic,
                         int l2_window,
                         int coherent_loads)
{
#if defined(CUCO_SHIM_REFERENCE)
  (void)keys_per_thread, (void)cas_first, (void)sector_chunks, (void)waves, (void)force_generic,
    (void)l2_window, (void)coherent_loads;
  return 1;
#else
  auto& t = cuco::b200::tuning();
  // keys_per_thread = lookups + 10 * mutations when >= 10 (e.g. 12 = lookups 2, mutations 1)
  int const lookup_kpt = keys_per_thread % 10, mutate_kpt = keys_per_thread / 10;
  if (lookup_kpt == 1 || lookup_kpt == 2 || lookup_kpt == 4) { t.keys_per_thread = lookup_kpt; }
  if (mutate_kpt == 1 || mutate_kpt == 2 || mutate_kpt == 4) {
    t.mutate_keys_per_thread = mutate_kpt;
  } else if (keys_per_thread > 0 && keys_per_thread < 10) {
    t.mutate_keys_per_thread = lookup_kpt;  // single digit sets both
  }
  if (cas_first >= 0) { t.cas_first = cas_first != 0; }
  if (sector_chunks >= 0) { t.sector_chunks = sector_chunks != 0; }
  if (waves >= 0) { t.waves = waves; }
  if (force_generic >= 0) { t.force_generic = force_generic != 0; }
  if (l2_window >= 0) { t.l2_window = l2_window != 0; }
  if (coherent_loads >= 0) { t.coherent_loads = coherent_loads != 0; }
  return 0;
#endif
}

int cuco_b200_set_blocking(int mode, int region_mib)
{
#if defined(CUCO_SHIM_REFERENCE)
  (void)mode, (void)region_mib;
  return 1;
#else
  auto& t   = cuco::b200::tuning();
  t.blocked = mode < 0 ? -1 : (mode > 0 ? 1 : 0);
  if (region_mib > 0) { t.region_bytes = static_cast<std::size_t>(region_mib) << 20; }
  if (region_mib < 0) { t.region_bytes = static_cast<std::size_t>(-region_mib) << 10; }  // KiB (tests)
  return 0;
#endif
}

int cuco_b200_set_blocking_variant(int keys_per_thread, int cas_first, int prefetch)
{
#if defined(CUCO_SHIM_REFERENCE)
  (void)keys_per_thread, (void)cas_first, (void)prefetch;
  return 1;
#else
  auto& t = cuco::b200::tuning();
  if (keys_per_thread == 1 || keys_per_thread == 2 || keys_per_thread == 4) {
    t.blocked_keys_per_thread = keys_per_thread;
  }
  if (cas_first >= 0) { t.blocked_cas_first = cas_first != 0; }
  if (prefetch >= 0) { t.blocked_prefetch = prefetch != 0; }
  return 0;
#endif
}

int cuco_b200_set_stream_variant(int tile_route, int stream_probe, int slots)
{
#if defined(CUCO_SHIM_REFERENCE)
  (void)tile_route, (void)stream_probe, (void)slots;
  return 1;
#else
  auto& t = cuco::b200::tuning();
  if (tile_route >= 0) { t.blocked_tile_route = tile_route != 0; }
  if (stream_probe >= 0) { t.blocked_stream_probe = stream_probe != 0; }
  if (slots == 1 || slots == 2) { t.stream_slots = slots; }
  return 0;
#endif
}

int cuco_b200_partition_count(const void* keys,
                              int key_bytes,
                              int pair_aos,
                              int64_t n,
                              int num_parts,
                              uint64_t salt,
                              int64_t* counts,
                              void* stream)
{
  return guarded([&] {
    require(n >= 0 && (keys || n == 0) && counts, "bad argument");
    require(num_parts >= 1 && num_parts <= max_parts, "num_parts must be in [1, 64]");
    require(key_bytes == 4 || key_bytes == 8, "key_bytes must be 4 or 8");
    if (n == 0) { return; }
    auto const grid = route_grid(n, std::int64_t{route_block} * route_items);
    auto* c         = reinterpret_cast<unsigned long long*>(counts);
    auto s          = static_cast<cudaStream_t>(stream);
    if (key_bytes == 8) {
      auto const* k = static_cast<std::int64_t const*>(keys);
      if (pair_aos) {
        partition_count_kernel<std::int64_t, 2><<<grid, route_block, 0, s>>>(k, n, num_parts, salt, c);
      } else {
        partition_count_kernel<std::int64_t, 1><<<grid, route_block, 0, s>>>(k, n, num_parts, salt, c);
      }
    } else {
      auto const* k = static_cast<std::int32_t const*>(keys);
      if (pair_aos) {
        partition_count_kernel<std::int32_t, 2><<<grid, route_block, 0, s>>>(k, n, num_parts, salt, c);
      } else {
        partition_count_kernel<std::int32_t, 1><<<grid, route_block, 0, s>>>(k, n, num_parts, salt, c);
      }
    }
    check_launch();
  });
}

int cuco_b200_partition_scatter(const void* keys,
                                const void* values,
                                int key_bytes,
                                int value_bytes,
                                int pair_aos,
                                int64_t n,
                                int num_parts,
                                uint64_t salt,
                                int64_t* cursors,
                                void* keys_out,
                                void* values_out,
                                int64_t* src_index,
                                void* stream)
{
  return guarded([&] {
    require(n >= 0 && ((keys && keys_out) || n == 0) && cursors, "bad argument");
    require(num_parts >= 1 && num_parts <= max_parts, "num_parts must be in [1, 64]");
    require(key_bytes == 4 || key_bytes == 8, "key_bytes must be 4 or 8");
    bool const has_payload = pair_aos || values != nullptr;
    require(!has_payload || pair_aos || (value_bytes == key_bytes && values_out),
            "separate value arrays must have the key width and an output");
    if (n == 0) { return; }
    auto const grid = route_grid(n, std::int64_t{route_block} * route_items);
    auto* c         = reinterpret_cast<unsigned long long*>(cursors);
    auto s          = static_cast<cudaStream_t>(stream);
    auto launch     = [&](auto key_tag) {
      using K       = decltype(key_tag);
      auto const* k = static_cast<K const*>(keys);
      auto const* v = static_cast<K const*>(values);
      auto* ko      = static_cast<K*>(keys_out);
      auto* vo      = static_cast<K*>(values_out);
      if (pair_aos) {
        partition_scatter_kernel<K, K, true, true>
          <<<grid, route_block, 0, s>>>(k, v, n, num_parts, salt, c, ko, vo, src_index);
      } else if (has_payload) {
        partition_scatter_kernel<K, K, true, false>
          <<<grid, route_block, 0, s>>>(k, v, n, num_parts, salt, c, ko, vo, src_index);
      } else {
        partition_scatter_kernel<K, K, false, false>
          <<<grid, route_block, 0, s>>>(k, v, n, num_parts, salt, c, ko, vo, src_index);
      }
    };
    if (key_bytes == 8) {
      launch(std::int64_t{});
    } else {
      launch(std::int32_t{});
    }
    check_launch();
  });
}

int cuco_b200_scatter_by_index(
  const void* in, const int64_t* index, void* out, int elem_bytes, int64_t n, void* stream)
{
  return guarded([&] {
    require(n >= 0 && ((in && index && out) || n == 0), "bad argument");
    if (n == 0) { return; }
    auto const grid = route_grid(n, route_block);
    auto s          = static_cast<cudaStream_t>(stream);
    switch (elem_bytes) {
      case 1:
        scatter_by_index_kernel<std::uint8_t><<<grid, route_block, 0, s>>>(
          static_cast<std::uint8_t const*>(in), index, static_cast<std::uint8_t*>(out), n);
        break;
      case 4:
        scatter_by_index_kernel<std::uint32_t><<<grid, route_block, 0, s>>>(
          static_cast<std::uint32_t const*>(in), index, static_cast<std::uint32_t*>(out), n);
        break;
      case 8:
        scatter_by_index_kernel<std::uint64_t><<<grid, route_block, 0, s>>>(
          static_cast<std::uint64_t const*>(in), index, static_cast<std::uint64_t*>(out), n);
        break;
      default: throw std::invalid_argument("elem_bytes must be 1, 4 or 8");
    }
    check_launch();
  });
}

}  // extern "C"
#pragma GCC visibility pop
