// table_engine — host side of the hot path: owns the slot storage and launches the bulk kernels.
//
// Counterpart of the reference's `detail::open_addressing_impl`
// (include/cuco/detail/open_addressing/open_addressing_impl.cuh:69-1168) plus the map-only
// launchers (detail/static_map/static_map.inl:279-358, detail/static_map/helpers.cuh:50-114).
// Same stream-ordered contract: `*_async` never synchronise, the counting variants return after one
// device->host copy. Differences in how the work is issued:
//   * thread-per-key kernels with several probes in flight per thread, one CTA per tile (measured
//     10-30 % faster than a persistent grid for random probes), instead of one 128-thread block per
//     128/cg_size keys;
//   * contiguous iterators are unwrapped to raw pointers so inputs stream through 128-bit
//     non-allocating loads;
//   * mutations of tables far beyond L2 take the L2-blocked path when the batch is dense enough:
//     route the batch by 16 MB table region, then probe region by region (bulk_kernels.cuh);
//   * the success counter of the synchronous insert lives with the container (the reference
//     cudaMallocs and frees one per call, impl.cuh:337-347), as does the staging buffer of the
//     blocked path (grow-only, ~1.07 x batch bytes);
//   * tables small enough to live in L2 get an access-policy window on the launch (persisting
//     lines for the table, streaming for everything else);
//   * hash-partitioned multi-GPU tables: exchange_* members route a batch straight into the owner
//     ranks' memory and probe / answer it there (no reference counterpart).
// Run-time tuning (keys per thread, CAS-first, chunk width, blocked-path variants) is compiled in only
// when CUCO_B200_TUNABLE is defined (the C-ABI library used by bench.py does, for the benchmarked
// instantiations); otherwise the defaults below are the only instantiations.
#pragma once

#include <cuco/b200/blocked_match.cuh>
#include <cuco/b200/bulk_kernels.cuh>
#include <cuco/b200/match_kernels.cuh>
#include <cuco/b200/probe_engine.cuh>
#include <cuco/b200/stream_kernels.cuh>
#include <cuco/detail/error.hpp>
#include <cuco/detail/utility/cuda.hpp>
#include <cuco/detail/utils.hpp>
#include <cuco/extent.cuh>
#include <cuco/operator.hpp>
#include <cuco/probing_scheme.cuh>
#include <cuco/storage.cuh>

#include <cuda/stream_ref>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/type_traits/is_contiguous_iterator.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <type_traits>
#include <utility>

namespace cuco::b200 {

/// Knobs of the bulk launchers. Defaults are what the header-only build uses.
struct tuning_t {
  int keys_per_thread  = 2;     ///< independent probes in flight per thread (1, 2 or 4)
  bool cas_first       = false; ///< inserts start with the CAS instead of a load (pays off only
                                ///< when nearly every key is new: a failed CAS still dirties the sector)
  bool sector_chunks   = true;  ///< 32-byte chunk loads (else one window per load)
  int waves            = 0;     ///< 0: one CTA per tile; k > 0: persistent grid of k * resident CTAs.
                                ///< Measured on B200: one tile per CTA is 10-30 % faster than a
                                ///< one-wave persistent grid on every op (profiles/r01_sweep.md)
  int mutate_keys_per_thread = 1;  ///< insert-type kernels: 2 dependent memory operations per key,
                                   ///< so occupancy (fewer registers) beats per-thread MLP
  bool force_generic   = false; ///< route everything through the one-key-per-thread fallback
  bool l2_window       = true;  ///< persisting-L2 access window for tables <= l2_window_bytes
  bool coherent_loads  = false; ///< mutating kernels read the table with relaxed.gpu loads
  int blocked          = -1;    ///< L2-blocked mutations: -1 auto (tables much larger than L2), 0 off, 1 on
  std::size_t region_bytes          = std::size_t{16} << 20;   ///< table slice kept L2-resident
  std::size_t blocked_min_table     = std::size_t{256} << 20;  ///< auto mode: table at least this big
  std::int64_t blocked_min_elements = std::int64_t{1} << 22;   ///< auto mode: batch at least this big
  int blocked_keys_per_thread = 4;  ///< pass 2 of the blocked path: probes in flight per thread
  bool blocked_cas_first      = true;   ///< pass 2 starts with the CAS (table slice is L2-resident;
                                        ///< measured 34.9 vs 30.6 Gops/s, profiles/r01_insert_probe_v4.jsonl)
  bool blocked_prefetch       = true;   ///< pass 2 streams the next region into L2 ahead of use
  bool blocked_tile_route     = true;   ///< pass 1 = tile_route_kernel (bulk-copy input, persistent) when
                                        ///< the batch is a 16-byte aligned array of slot images
  bool blocked_stream_probe   = false;  ///< pass 2 = stream_mutate_kernel (warp-persistent, refilling); measured
                                        ///< SLOWER than blocked_mutate_kernel on B200 (5.1 vs 2.6 ms per 100 M,
                                        ///< profiles/r02_insert_lab_v2_rows_parked.jsonl): kept for sweeps only
  int stream_slots            = 2;      ///< rows of 32 keys in flight per warp of stream_mutate_kernel (1, 2)
  bool stream_scout           = true;   ///< every key's home line is prefetched into L2 one chunk ahead of its CAS
  int exchange_lookup_keys_per_thread = 2;  ///< owner side of routed lookups (1, 2 or 4)
  std::size_t l2_window_bytes = std::size_t{48} << 20;
  int match_ahead = 1;  ///< retrieve on tables with duplicates: chunks of the probe sequence loaded
                        ///< together while they stay inside one 128-byte line (1, 2 or 4)
  int count_ahead = 2;  ///< the same for count, on tables of at least count_ahead_min_table bytes.
                        ///< Measured on B200 (profiles/r01_matches_ahead*.jsonl, multiplicity 4) for
                        ///< depth 1 / 2 / 4 on an 800 MB table: count 13.6 / 15.0 / 13.6 G probes/s,
                        ///< retrieve 11.6 / 10.5 / 6.1 G rows/s; on a 320 MB table depth 2 LOSES
                        ///< (count 17.7 -> 14.9): the extra registers cost occupancy and a third of
                        ///< the table already sits in L2
  std::size_t count_ahead_min_table = std::size_t{512} << 20;
};

inline tuning_t tuning_from_env()
{
  tuning_t t{};
  if (char const* s = std::getenv("CUCO_B200_KPT")) { t.keys_per_thread = std::atoi(s); }
  if (char const* s = std::getenv("CUCO_B200_CAS_FIRST")) { t.cas_first = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_SECTOR")) { t.sector_chunks = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_WAVES")) { t.waves = std::max(0, std::atoi(s)); }
  if (char const* s = std::getenv("CUCO_B200_MUTATE_KPT")) { t.mutate_keys_per_thread = std::atoi(s); }
  if (char const* s = std::getenv("CUCO_B200_GENERIC")) { t.force_generic = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_L2_WINDOW")) { t.l2_window = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_COHERENT")) { t.coherent_loads = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_BLOCKED")) { t.blocked = std::atoi(s); }
  if (char const* s = std::getenv("CUCO_B200_BLOCKED_KPT")) { t.blocked_keys_per_thread = std::atoi(s); }
  if (char const* s = std::getenv("CUCO_B200_BLOCKED_CAS_FIRST")) { t.blocked_cas_first = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_BLOCKED_PREFETCH")) { t.blocked_prefetch = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_TILE_ROUTE")) { t.blocked_tile_route = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_STREAM_PROBE")) { t.blocked_stream_probe = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_STREAM_SLOTS")) { t.stream_slots = std::atoi(s); }
  if (char const* s = std::getenv("CUCO_B200_STREAM_SCOUT")) { t.stream_scout = std::atoi(s) != 0; }
  if (char const* s = std::getenv("CUCO_B200_EXCHANGE_LOOKUP_KPT")) {
    t.exchange_lookup_keys_per_thread = std::atoi(s);
  }
  if (char const* s = std::getenv("CUCO_B200_MATCH_AHEAD")) {
    t.match_ahead = t.count_ahead = std::atoi(s);  // one switch for sweeps, at every table size
    t.count_ahead_min_table = 0;
  }
  if (char const* s = std::getenv("CUCO_B200_REGION_MIB")) {
    t.region_bytes = static_cast<std::size_t>(std::max(1, std::atoi(s))) << 20;
  }
  return t;
}

/// Process-wide tuning state (mutable so a harness can sweep it between launches).
inline tuning_t& tuning()
{
  static tuning_t t = tuning_from_env();
  return t;
}

/// Grid for a persistent kernel: resident CTAs on the device times `waves`, capped by the tiles.
template <typename Kernel>
inline unsigned persistent_grid(Kernel kernel, int block_size, cuco::detail::index_type tiles)
{
  static int resident = 0;  // one per kernel instantiation; devices in a process are identical
  if (resident == 0) { resident = cuco::detail::max_occupancy_grid_size(block_size, kernel); }
  auto const waves = tuning().waves;
  auto const want  = waves > 0 ? static_cast<cuco::detail::index_type>(resident) * waves
                               : cuco::detail::index_type{0x7fffffff};
  return static_cast<unsigned>(std::max<cuco::detail::index_type>(1, std::min(tiles, want)));
}

/// Launches `kernel` on `stream`; when `window_bytes` is non-zero the launch carries an L2 access
/// policy window over [window_base, +window_bytes) marking those lines persisting.
template <typename Kernel, typename... Args>
inline void launch(Kernel kernel,
                   unsigned grid,
                   unsigned block,
                   cudaStream_t stream,
                   void* window_base,
                   std::size_t window_bytes,
                   Args... args)
{
  if (window_bytes == 0) {
    kernel<<<grid, block, 0, stream>>>(args...);
    return;
  }
  cudaLaunchConfig_t config{};
  config.gridDim          = dim3{grid};
  config.blockDim         = dim3{block};
  config.dynamicSmemBytes = 0;
  config.stream           = stream;
  cudaLaunchAttribute attr{};
  attr.id                                 = cudaLaunchAttributeAccessPolicyWindow;
  attr.val.accessPolicyWindow.base_ptr    = window_base;
  attr.val.accessPolicyWindow.num_bytes   = window_bytes;
  attr.val.accessPolicyWindow.hitRatio    = 1.0f;
  attr.val.accessPolicyWindow.hitProp     = cudaAccessPropertyPersisting;
  attr.val.accessPolicyWindow.missProp    = cudaAccessPropertyStreaming;
  config.attrs                            = &attr;
  config.numAttrs                         = 1;
  cudaLaunchKernelEx(&config, kernel, args...);
}

/// Makes sure the device has a persisting-L2 carve-out (once per process); returns its size.
inline std::size_t persisting_l2_bytes()
{
  static std::size_t bytes = [] {
    int dev = 0, max_persist = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { return std::size_t{0}; }
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (max_persist <= 0) { return std::size_t{0}; }
    std::size_t current = 0;
    cudaDeviceGetLimit(&current, cudaLimitPersistingL2CacheSize);
    if (current == 0) {
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, static_cast<std::size_t>(max_persist)) !=
          cudaSuccess) {
        cudaGetLastError();
        return std::size_t{0};
      }
      current = static_cast<std::size_t>(max_persist);
    }
    return current;
  }();
  return bytes;
}

/// Opts `kernel` into `bytes` of dynamic shared memory on the CURRENT device (the attribute is per
/// device and per function, so a process driving several GPUs must set it on each). Cached per
/// (kernel address, device); kernels of different instantiations may share one C++ type, so the
/// cache is keyed by the pointer value, never by the type. Returns false if the device refuses.
template <typename Kernel>
inline bool opt_in_dynamic_smem(Kernel kernel, std::size_t bytes)
{
  static std::mutex guard;
  static std::map<std::pair<void const*, int>, bool> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { return false; }
  auto const key = std::make_pair(reinterpret_cast<void const*>(kernel), dev);
  std::lock_guard<std::mutex> lock{guard};
  auto const it = done.find(key);
  if (it != done.end()) { return it->second; }
  bool const ok = cudaFuncSetAttribute(kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes)) == cudaSuccess;
  if (!ok) { cudaGetLastError(); }
  done.emplace(key, ok);
  return ok;
}

template <typename It>
inline auto unwrap(It it)
{
  return thrust::try_unwrap_contiguous_iterator(it);
}

/// Key type of an input element: the element itself for sets; `first` / `get<0>` of a pair-like for
/// maps (host-side mirror of probe_engine::heterogeneous_value + key_of).
template <typename T, typename = void>
struct has_member_first : std::false_type {};
template <typename T>
struct has_member_first<T, std::void_t<decltype(std::declval<T const&>().first)>> : std::true_type {};

template <typename T>
struct type_box {
  using type = T;
};

template <typename T, bool HasPayload>
struct input_key_of {
  static auto probe()
  {
    if constexpr (!HasPayload) {
      return type_box<cuda::std::remove_cv_t<T>>{};
    } else if constexpr (has_member_first<T>::value) {
      return type_box<
        cuda::std::remove_cv_t<cuda::std::remove_reference_t<decltype(std::declval<T const&>().first)>>>{};
    } else {
      return type_box<cuda::std::remove_cv_t<cuda::std::tuple_element_t<0, T>>>{};
    }
  }
  using type = typename decltype(probe())::type;
};

template <class Key,
          class Value,
          class Extent,
          cuda::thread_scope Scope,
          class KeyEqual,
          class ProbingScheme,
          class Allocator,
          class Storage,
          bool AllowsDuplicates = false>
class table_engine {
  static_assert(sizeof(Key) <= 8, "Container does not support key types larger than 8 bytes.");
  static_assert(sizeof(Value) <= 16, "Container does not support slot types larger than 16 bytes.");
  static_assert(
    cuco::is_bitwise_comparable_v<Key>,
    "Key type must have unique object representations or have been explicitly declared as safe for "
    "bitwise comparison via specialization of cuco::is_bitwise_comparable_v<Key>.");
  static_assert(
    std::is_base_of_v<cuco::detail::probing_scheme_base<ProbingScheme::cg_size>, ProbingScheme>,
    "ProbingScheme must inherit from cuco::detail::probing_scheme_base");

 public:
  static constexpr auto has_payload  = !std::is_same_v<Key, Value>;
  static constexpr auto cg_size      = ProbingScheme::cg_size;
  static constexpr auto window_size  = Storage::window_size;
  static constexpr auto thread_scope = Scope;

  using key_type    = Key;
  using value_type  = Value;  ///< slot type
  using extent_type = decltype(make_window_extent<table_engine>(std::declval<Extent>()));
  using size_type   = typename extent_type::value_type;
  using key_equal   = KeyEqual;
  using storage_type        = cuco::detail::storage<Storage, value_type, extent_type, Allocator>;
  using allocator_type      = typename storage_type::allocator_type;
  using storage_ref_type    = typename storage_type::ref_type;
  using probing_scheme_type = ProbingScheme;
  using hasher              = typename probing_scheme_type::hasher;
  static constexpr bool allows_duplicates = AllowsDuplicates;  ///< multiset / multimap semantics
  using engine_type =
    probe_engine<key_type, Scope, key_equal, probing_scheme_type, storage_ref_type, AllowsDuplicates>;

  static constexpr int block_size = 256;

  // ------------------------------------------------------------------------------------------
  // construction
  // ------------------------------------------------------------------------------------------
  constexpr table_engine(Extent capacity,
                         Value empty_slot_sentinel,
                         KeyEqual const& pred,
                         ProbingScheme const& probing_scheme,
                         Allocator const& alloc,
                         cuda::stream_ref stream)
    : empty_slot_sentinel_{empty_slot_sentinel},
      erased_key_sentinel_{key_of(empty_slot_sentinel)},
      predicate_{pred},
      probing_scheme_{probing_scheme},
      storage_{make_window_extent<table_engine>(capacity), alloc}
  {
    this->clear_async(stream);
  }

  constexpr table_engine(Extent n,
                         double desired_load_factor,
                         Value empty_slot_sentinel,
                         KeyEqual const& pred,
                         ProbingScheme const& probing_scheme,
                         Allocator const& alloc,
                         cuda::stream_ref stream)
    : empty_slot_sentinel_{empty_slot_sentinel},
      erased_key_sentinel_{key_of(empty_slot_sentinel)},
      predicate_{pred},
      probing_scheme_{probing_scheme},
      storage_{make_window_extent<table_engine>(checked_capacity(n, desired_load_factor)), alloc}
  {
    this->clear_async(stream);
  }

  constexpr table_engine(Extent capacity,
                         Value empty_slot_sentinel,
                         Key erased_key_sentinel,
                         KeyEqual const& pred,
                         ProbingScheme const& probing_scheme,
                         Allocator const& alloc,
                         cuda::stream_ref stream)
    : empty_slot_sentinel_{empty_slot_sentinel},
      erased_key_sentinel_{erased_key_sentinel},
      predicate_{pred},
      probing_scheme_{probing_scheme},
      storage_{make_window_extent<table_engine>(capacity), alloc}
  {
    CUCO_EXPECTS(this->empty_key_sentinel() != this->erased_key_sentinel(),
                 "The empty key sentinel and erased key sentinel cannot be the same value.",
                 std::logic_error);
    this->clear_async(stream);
  }

  ~table_engine() = default;
  table_engine(table_engine const&)            = delete;
  table_engine& operator=(table_engine const&) = delete;

  void clear(cuda::stream_ref stream) { storage_.initialize(empty_slot_sentinel_, stream); }
  void clear_async(cuda::stream_ref stream) noexcept
  {
    storage_.initialize_async(empty_slot_sentinel_, stream);
  }

  // ------------------------------------------------------------------------------------------
  // insert family
  // ------------------------------------------------------------------------------------------
  template <typename InputIt, typename Ref>
  size_type insert(InputIt first, InputIt last, Ref ref, cuda::stream_ref stream)
  {
    return this->insert_if(
      first, last, thrust::constant_iterator<bool>{true}, always_true{}, ref, stream);
  }

  template <typename InputIt, typename Ref>
  void insert_async(InputIt first, InputIt last, Ref ref, cuda::stream_ref stream) noexcept
  {
    this->insert_if_async(
      first, last, thrust::constant_iterator<bool>{true}, always_true{}, ref, stream);
  }

  template <typename InputIt, typename StencilIt, typename Predicate, typename Ref>
  size_type insert_if(InputIt first,
                      InputIt last,
                      StencilIt stencil,
                      Predicate pred,
                      Ref ref,
                      cuda::stream_ref stream)
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return 0; }
    auto* counter = this->zeroed_counter(stream);
    this->mutate<true>(first, n, stencil, pred, counter, ref, action_insert{}, stream);
    return this->read_counter(counter, stream);
  }

  template <typename InputIt, typename StencilIt, typename Predicate, typename Ref>
  void insert_if_async(InputIt first,
                       InputIt last,
                       StencilIt stencil,
                       Predicate pred,
                       Ref ref,
                       cuda::stream_ref stream) noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    this->mutate<false>(
      first, n, stencil, pred, static_cast<size_type*>(nullptr), ref, action_insert{}, stream);
  }

  template <typename InputIt, typename FoundIt, typename InsertedIt, typename Ref>
  void insert_and_find_async(InputIt first,
                             InputIt last,
                             FoundIt found_begin,
                             InsertedIt inserted_begin,
                             Ref ref,
                             cuda::stream_ref stream) noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    auto found    = unwrap(found_begin);
    auto inserted = unwrap(inserted_begin);
    this->mutate<false>(first,
                        n,
                        thrust::constant_iterator<bool>{true},
                        always_true{},
                        static_cast<size_type*>(nullptr),
                        ref,
                        action_insert_and_find<decltype(found), decltype(inserted)>{found, inserted},
                        stream);
  }

  template <typename InputIt, typename Ref>
  void insert_or_assign_async(InputIt first, InputIt last, Ref ref, cuda::stream_ref stream) noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    this->mutate<false>(first,
                        n,
                        thrust::constant_iterator<bool>{true},
                        always_true{},
                        static_cast<size_type*>(nullptr),
                        ref,
                        action_assign{},
                        stream);
  }

  /// `direct_apply`: the caller passed an init equal to the empty payload, so first arrivals combine
  /// onto the sentinel instead of storing (decided on the host; same rule as
  /// static_map_ref.inl:788-829 applies per element on the device).
  template <typename InputIt, typename Op, typename Ref>
  void insert_or_apply_async(
    InputIt first, InputIt last, bool direct_apply, Op op, Ref ref, cuda::stream_ref stream) noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    auto const all = thrust::constant_iterator<bool>{true};
    auto* no_count = static_cast<size_type*>(nullptr);
    if (direct_apply) {
      this->mutate<false>(
        first, n, all, always_true{}, no_count, ref, action_apply<Op, true>{op}, stream);
    } else {
      this->mutate<false>(
        first, n, all, always_true{}, no_count, ref, action_apply<Op, false>{op}, stream);
    }
  }

  template <typename InputIt, typename Ref>
  void erase_async(InputIt first, InputIt last, Ref ref, cuda::stream_ref stream = {})
  {
    CUCO_EXPECTS(this->empty_key_sentinel() != this->erased_key_sentinel(),
                 "The empty key sentinel and erased key sentinel cannot be the same value.",
                 std::logic_error);
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    auto in           = unwrap(first);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    // two keys per thread, one CTA per 512 keys (random probes: the hardware scheduler balances them)
    auto const grid = static_cast<unsigned>(std::min<cuco::detail::index_type>(
      cuco::detail::int_div_ceil(n, cuco::detail::index_type{block_size} * 2), 0x7fffffff));
    if (this->fast_path_ok(false)) {
      erase_kernel<block_size, engine_t::sector_chunk_slots>
        <<<grid, block_size, 0, stream.get()>>>(in, n, engine);
    } else {
      erase_kernel<block_size, engine_t::window_chunk_slots>
        <<<grid, block_size, 0, stream.get()>>>(in, n, engine);
    }
  }

  // ------------------------------------------------------------------------------------------
  // lookups
  // ------------------------------------------------------------------------------------------
  template <typename InputIt, typename OutputIt, typename Ref>
  void contains_async(
    InputIt first, InputIt last, OutputIt output_begin, Ref ref, cuda::stream_ref stream) const noexcept
  {
    this->contains_if_async(
      first, last, thrust::constant_iterator<bool>{true}, always_true{}, output_begin, ref, stream);
  }

  template <typename InputIt,
            typename StencilIt,
            typename Predicate,
            typename OutputIt,
            typename Ref>
  void contains_if_async(InputIt first,
                         InputIt last,
                         StencilIt stencil,
                         Predicate pred,
                         OutputIt output_begin,
                         Ref ref,
                         cuda::stream_ref stream) const noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    this->lookup(first, n, stencil, pred, output_begin, ref, emit_present{}, stream);
  }

  template <typename InputIt, typename OutputIt, typename Ref>
  void find_async(
    InputIt first, InputIt last, OutputIt output_begin, Ref ref, cuda::stream_ref stream) const noexcept
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return; }
    this->lookup(first,
                 n,
                 thrust::constant_iterator<bool>{true},
                 always_true{},
                 output_begin,
                 ref,
                 emit_found<typename Ref::engine_type>{empty_slot_sentinel_},
                 stream);
  }

  // ------------------------------------------------------------------------------------------
  // all-matches queries: count / retrieve (match_kernels.cuh)
  // ------------------------------------------------------------------------------------------
  /// Sum over [first, last) of the number of stored elements matching each key; with `IsOuter` a key
  /// without matches counts as one (reference open_addressing_impl.cuh:677-706). Synchronises.
  template <bool IsOuter, typename InputIt, typename Ref>
  [[nodiscard]] size_type count(InputIt first, InputIt last, Ref ref, cuda::stream_ref stream) const
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return 0; }
    auto in           = unwrap(first);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    auto* counter     = this->zeroed_counter(stream);
    auto const grid   = generic_grid(n);
    if (this->blocked_matches<IsOuter, true>(in, n, static_cast<value_type*>(nullptr),
                                             static_cast<value_type*>(nullptr), counter, engine, stream)) {
      return this->read_counter(counter, stream);
    }
    if (this->fast_path_ok(false)) {
      auto const table_bytes = static_cast<std::size_t>(storage_.capacity()) * sizeof(value_type);
      auto const depth = table_bytes >= tuning().count_ahead_min_table ? tuning().count_ahead : 1;
      with_match_ahead<engine_t>(depth, [&](auto ahead) {
        count_kernel<IsOuter, block_size, engine_t::sector_chunk_slots, decltype(ahead)::value>
          <<<grid, block_size, 0, stream.get()>>>(in, n, counter, engine);
      });
    } else {
      count_kernel<IsOuter, block_size, engine_t::window_chunk_slots, 1>
        <<<grid, block_size, 0, stream.get()>>>(in, n, counter, engine);
    }
    return this->read_counter(counter, stream);
  }

  /// For every key of [first, last) and every stored element matching it, writes the key to
  /// `output_probe` and the element to `output_match` (same position, unspecified order); with
  /// `IsOuter` a key without matches yields {key, empty slot sentinel}. Returns the number of rows
  /// (reference open_addressing_impl.cuh:604-660, static_set.inl:349-373). Synchronises.
  template <bool IsOuter, typename InputIt, typename OutputProbeIt, typename OutputMatchIt, typename Ref>
  size_type retrieve(InputIt first,
                     InputIt last,
                     OutputProbeIt output_probe,
                     OutputMatchIt output_match,
                     Ref ref,
                     cuda::stream_ref stream) const
  {
    auto const n = cuco::detail::distance(first, last);
    if (n == 0) { return 0; }
    auto in           = unwrap(first);
    auto out_probe    = unwrap(output_probe);
    auto out_match    = unwrap(output_match);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    auto* counter     = this->zeroed_counter(stream);
    if (this->blocked_matches<IsOuter, false>(in, n, out_probe, out_match, counter, engine, stream)) {
      return this->read_counter(counter, stream);
    }
    // one CTA per round of 256 keys (like the other random-probe kernels)
    auto const grid = static_cast<unsigned>(std::min<cuco::detail::index_type>(
      cuco::detail::int_div_ceil(n, cuco::detail::index_type{block_size}), 0x7fffffff));
    if (this->fast_path_ok(false)) {
      with_match_ahead<engine_t>(tuning().match_ahead, [&](auto ahead) {
        retrieve_kernel<IsOuter, block_size, engine_t::sector_chunk_slots, decltype(ahead)::value>
          <<<grid, block_size, 0, stream.get()>>>(in, n, out_probe, out_match, counter, engine);
      });
    } else {
      retrieve_kernel<IsOuter, block_size, engine_t::window_chunk_slots, 1>
        <<<grid, block_size, 0, stream.get()>>>(in, n, out_probe, out_match, counter, engine);
    }
    return this->read_counter(counter, stream);
  }

  /// L2-blocked count / retrieve (blocked_match.cuh) for probe batches that are a plain, 16-byte
  /// aligned array of keys over a table much larger than L2: stage the keys grouped by table region,
  /// then probe region by region. Returns false - nothing launched, the caller takes the direct
  /// kernels - when the path does not apply or scratch memory is not to be had. `IsCount` selects
  /// the total-only flavour (the output iterators are ignored).
  template <bool IsOuter,
            bool IsCount,
            typename InputIt,
            typename OutputProbeIt,
            typename OutputMatchIt,
            typename EngineT>
  [[nodiscard]] bool blocked_matches(InputIt in,
                                     cuco::detail::index_type n,
                                     OutputProbeIt out_probe,
                                     OutputMatchIt out_match,
                                     size_type* counter,
                                     EngineT const& engine,
                                     cuda::stream_ref stream) const
  {
    using cuco::detail::index_type;
    if constexpr (!std::is_pointer_v<InputIt>) {
      return false;
    } else if constexpr (!std::is_same_v<std::remove_cv_t<std::remove_pointer_t<InputIt>>, key_type> ||
                         !(sizeof(key_type) == 4 || sizeof(key_type) == 8)) {
      return false;
    } else {
      auto const& t          = tuning();
      auto const capacity    = static_cast<std::uint64_t>(storage_.capacity());
      auto const table_bytes = capacity * sizeof(value_type);
      if (t.blocked == 0 || !this->fast_path_ok(false) || (reinterpret_cast<std::uintptr_t>(in) % 16) != 0) {
        return false;
      }
      if (table_bytes / route_max_regions > (std::size_t{64} << 20) || n >= (index_type{1} << 31)) { return false; }
      // auto mode: every probe fetches a 128-byte line of a table well beyond L2; grouping pays once
      // the batch touches the table about once per line or denser
      if (t.blocked < 0 && !(table_bytes >= t.blocked_min_table && n >= t.blocked_min_elements &&
                             static_cast<std::uint64_t>(n) * 128 >= table_bytes)) {
        return false;
      }
      auto const num_regions = static_cast<std::uint32_t>(std::min<std::uint64_t>(
        route_max_regions, std::max<std::uint64_t>(2, (table_bytes + t.region_bytes - 1) / t.region_bytes)));
      auto const mean = (static_cast<std::uint64_t>(n) + num_regions - 1) / num_regions;
      auto const cap  = static_cast<std::uint32_t>((mean + mean / 16 + 1024 + 15) / 16 * 16);
      auto const spill_capacity =
        static_cast<std::uint32_t>(std::max<std::uint64_t>(65536, static_cast<std::uint64_t>(n) / 8));
      std::size_t const counts_bytes = (((num_regions + 1) * sizeof(unsigned int)) + 255) / 256 * 256;
      std::size_t const staged_bytes = (static_cast<std::size_t>(num_regions) * cap * sizeof(key_type) + 255) / 256 * 256;
      auto* base = static_cast<char*>(this->scratch_alloc(
        counts_bytes + staged_bytes + static_cast<std::size_t>(spill_capacity) * sizeof(key_type), stream.get()));
      if (base == nullptr) { return false; }
      auto* counts      = reinterpret_cast<unsigned int*>(base);  // [num_regions] fills, then the spill count
      auto* spill_count = counts + num_regions;
      auto* segments    = reinterpret_cast<key_type*>(base + counts_bytes);
      auto* spill       = reinterpret_cast<key_type*>(base + counts_bytes + staged_bytes);
      constexpr int chunk = EngineT::sector_chunk_slots;

      // ---- pass 1: group the probe keys by region ----
      using stencil_t   = thrust::constant_iterator<bool>;
      auto const router = tile_route_kernel<tile_route_block_size,
                                            chunk,
                                            false,
                                            stencil_t,
                                            always_true,
                                            size_type,
                                            EngineT,
                                            action_insert,
                                            region_spill_router,
                                            key_type>;
      auto const smem = tile_route_smem_bytes<tile_route_block_size, key_type>(num_regions);
      if (!opt_in_dynamic_smem(router, std::size_t{200} << 10)) {
        this->scratch_free(base, stream.get());
        return false;
      }
      cudaMemsetAsync(counts, 0, (num_regions + 1) * sizeof(unsigned int), stream.get());
      {
        auto const tiles = cuco::detail::int_div_ceil(n, index_type{tile_route_block_size} * route_items_per_thread);
        int per_sm       = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, router, tile_route_block_size, smem);
        auto const resident = std::max(1, per_sm) * cuco::detail::multiprocessor_count();
        auto const grid     = static_cast<unsigned>(std::min<index_type>(tiles, resident));
        region_spill_router route{};
        route.regions = region_map::over(capacity, num_regions);
        router<<<grid, tile_route_block_size, smem, stream.get()>>>(
          in, n, stencil_t{true}, always_true{}, segments, counts, route, cap, static_cast<size_type*>(nullptr),
          engine, action_insert{}, route_spill{spill, spill_count, spill_capacity});
      }

      // ---- pass 2: probe region by region ----
      auto const region_slots = (capacity + num_regions - 1) / num_regions;
      auto const ctas         = static_cast<unsigned>(
        std::max<index_type>(1, cuco::detail::int_div_ceil(index_type{cap}, index_type{block_size} * 4)));
      auto const share = (region_slots * sizeof(value_type) + ctas - 1) / ctas;
      blocked_layout const layout{counts,
                                  cap,
                                  region_slots,
                                  table_bytes,
                                  t.blocked_prefetch ? static_cast<std::uint32_t>((share + 127) / 128 * 128) : 0u,
                                  1u};
      auto direct = [&](key_type const* keys, index_type count) {
        if constexpr (IsCount) {
          count_kernel<IsOuter, block_size, chunk, 1>
            <<<generic_grid(count), block_size, 0, stream.get()>>>(keys, count, counter, engine);
        } else {
          auto const grid = static_cast<unsigned>(std::min<index_type>(
            cuco::detail::int_div_ceil(count, index_type{block_size}), 0x7fffffff));
          retrieve_kernel<IsOuter, block_size, chunk, 1>
            <<<grid, block_size, 0, stream.get()>>>(keys, count, out_probe, out_match, counter, engine);
        }
      };
      if constexpr (IsCount) {
        blocked_count_kernel<IsOuter, block_size, chunk, 1>
          <<<dim3{ctas, num_regions}, block_size, 0, stream.get()>>>(segments, layout, counter, engine);
      } else {
        blocked_retrieve_kernel<IsOuter, block_size, chunk, 1>
          <<<dim3{ctas, num_regions}, block_size, 0, stream.get()>>>(
            segments, layout, out_probe, out_match, counter, engine);
      }

      // ---- keys that did not fit their segment (skewed batches) ----
      unsigned int spilled = 0;
      cudaMemcpyAsync(&spilled, spill_count, sizeof(spilled), cudaMemcpyDeviceToHost, stream.get());
      cudaStreamSynchronize(stream.get());
      if (spilled > spill_capacity) {
        // more than the list holds: start over with the direct kernel on the whole batch
        cudaMemsetAsync(counter, 0, sizeof(size_type), stream.get());
        direct(in, n);
      } else if (spilled > 0) {
        direct(spill, static_cast<index_type>(spilled));
      }
      this->scratch_free(base, stream.get());
      return true;
    }
  }

  // ------------------------------------------------------------------------------------------
  // whole-table queries
  // ------------------------------------------------------------------------------------------
  [[nodiscard]] size_type size(cuda::stream_ref stream) const
  {
    auto* counter     = this->zeroed_counter(stream);
    auto const engine = this->make_engine();
    auto const chunks = cuco::detail::int_div_ceil(
      static_cast<cuco::detail::index_type>(storage_.capacity()),
      cuco::detail::index_type{engine_type::sector_chunk_slots});
    auto const kernel = size_kernel<block_size, engine_type, size_type>;
    // persistent: (SM count x 8) CTAs stream the table, one atomic per warp at the end
    auto const tiles = cuco::detail::int_div_ceil(chunks, cuco::detail::index_type{block_size} * 4);
    auto const grid  = static_cast<unsigned>(std::max<cuco::detail::index_type>(
      1, std::min<cuco::detail::index_type>(tiles, cuco::detail::index_type{cuco::detail::multiprocessor_count()} * 8)));
    kernel<<<grid, block_size, 0, stream.get()>>>(engine, counter);
    return this->read_counter(counter, stream);
  }

  [[nodiscard]] constexpr auto capacity() const noexcept { return storage_.capacity(); }
  [[nodiscard]] constexpr key_type empty_key_sentinel() const noexcept
  {
    return key_of(empty_slot_sentinel_);
  }
  [[nodiscard]] constexpr value_type empty_slot_sentinel() const noexcept
  {
    return empty_slot_sentinel_;
  }
  [[nodiscard]] constexpr key_type erased_key_sentinel() const noexcept
  {
    return erased_key_sentinel_;
  }
  [[nodiscard]] constexpr key_equal key_eq() const noexcept { return predicate_; }
  [[nodiscard]] constexpr probing_scheme_type const& probing_scheme() const noexcept
  {
    return probing_scheme_;
  }
  [[nodiscard]] constexpr hasher hash_function() const noexcept
  {
    return probing_scheme_.hash_function();
  }
  [[nodiscard]] constexpr allocator_type allocator() const noexcept { return storage_.allocator(); }
  [[nodiscard]] constexpr storage_ref_type storage_ref() const noexcept { return storage_.ref(); }

  /// Engine over the container's own storage (what the bulk kernels run on).
  [[nodiscard]] engine_type make_engine() const noexcept
  {
    return engine_type{
      empty_slot_sentinel_, erased_key_sentinel_, predicate_, probing_scheme_, storage_.ref()};
  }

  /// Minimal stand-in for a container ref when the engine is built from the table itself.
  struct engine_handle {
    using engine_type = typename table_engine::engine_type;
    engine_type e;
    [[nodiscard]] __host__ __device__ engine_type const& engine() const noexcept { return e; }
  };

  /// insert_if over the table's current storage without going through a container ref (rehash).
  template <typename InputIt, typename StencilIt, typename Predicate>
  void insert_slots_if(
    InputIt first, InputIt last, StencilIt stencil, Predicate pred, cuda::stream_ref stream) noexcept
  {
    this->insert_if_async(first, last, stencil, pred, engine_handle{this->make_engine()}, stream);
  }

  /// Replaces the storage by a fresh one of `extent` windows; returns the old storage.
  storage_type exchange_storage(extent_type extent, cuda::stream_ref stream)
  {
    storage_type old = std::move(storage_);
    new (&storage_) storage_type{extent, old.allocator()};
    this->clear_async(stream);
    return old;
  }

 private:
  [[nodiscard]] static constexpr key_type const& key_of(value_type const& slot) noexcept
  {
    if constexpr (has_payload) {
      return slot.first;
    } else {
      return slot;
    }
  }

  static Extent checked_capacity(Extent n, double desired_load_factor)
  {
    CUCO_EXPECTS(desired_load_factor > 0., "Desired occupancy must be larger than zero");
    CUCO_EXPECTS(desired_load_factor <= 1., "Desired occupancy must be no larger than one");
    // `Extent` may be a cuco::extent or a plain integer (CTAD from `static_map{n, ...}`)
    using raw_size = typename extent_type::value_type;
    return Extent{static_cast<raw_size>(
      std::ceil(static_cast<double>(static_cast<raw_size>(n)) / desired_load_factor))};
  }

  static unsigned generic_grid(cuco::detail::index_type n)
  {
    auto const blocks = cuco::detail::int_div_ceil(n, cuco::detail::index_type{block_size});
    auto const cap    = static_cast<cuco::detail::index_type>(cuco::detail::multiprocessor_count()) * 16;
    return static_cast<unsigned>(std::max<cuco::detail::index_type>(1, std::min(blocks, cap)));
  }

  /// Look-ahead depth of the all-matches walks: only tables with duplicates have long walks.
  template <typename EngineT, typename Run>
  static void with_match_ahead(int depth, Run&& run)
  {
    if constexpr (!EngineT::allows_duplicates) {
      run(std::integral_constant<int, 1>{});
    } else {
      switch (depth) {
        case 4: run(std::integral_constant<int, 4>{}); break;
        case 2: run(std::integral_constant<int, 2>{}); break;
        default: run(std::integral_constant<int, 1>{}); break;
      }
    }
  }

  /// Can the sector-chunk, single-CAS kernels run on this table right now?
  [[nodiscard]] bool fast_path_ok(bool mutating) const noexcept
  {
    if (tuning().force_generic) { return false; }
    if (!engine_type::pow2_slot) { return false; }
    if ((reinterpret_cast<std::uintptr_t>(storage_.data()) % 32) != 0) { return false; }
    if (mutating) {
      // A configured erased-key sentinel does not rule the fast kernels out: a tombstone is never
      // bit-identical to the empty slot image, so the claim CAS on it fails, the slot classifies as
      // "available but not empty" and the key is finished by the tombstone-aware general driver
      // (mutate_slow_path). Keys that meet no tombstone never leave the fast path.
      if (!engine_type::single_cas) { return false; }
    }
    return true;
  }

  /// Bytes of L2 window to request for this table (0 = none).
  [[nodiscard]] std::size_t window_bytes() const noexcept
  {
    auto const bytes = static_cast<std::size_t>(storage_.capacity()) * sizeof(value_type);
    if (!tuning().l2_window || bytes > tuning().l2_window_bytes) { return 0; }
    auto const carve_out = persisting_l2_bytes();
    return bytes <= carve_out ? bytes : 0;
  }

  template <typename InputIt,
            typename StencilIt,
            typename Predicate,
            typename OutputIt,
            typename Ref,
            typename Emit>
  void lookup(InputIt first,
              cuco::detail::index_type n,
              StencilIt stencil,
              Predicate pred,
              OutputIt output_begin,
              Ref ref,
              Emit emit,
              cuda::stream_ref stream) const noexcept
  {
    auto in           = unwrap(first);
    auto st           = unwrap(stencil);
    auto out          = unwrap(output_begin);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    void* base        = storage_.data();
    auto const window = this->window_bytes();

    if (!this->fast_path_ok(false)) {
      auto const kernel = generic_lookup_kernel<block_size,
                                                decltype(in),
                                                decltype(st),
                                                Predicate,
                                                decltype(out),
                                                engine_t,
                                                Emit>;
      launch(kernel, generic_grid(n), block_size, stream.get(), base, window, in, n, st, pred, out, engine, emit);
      return;
    }

    auto run = [&](auto kpt, auto chunk) {
      constexpr int KPT   = decltype(kpt)::value;
      constexpr int Chunk = decltype(chunk)::value;
      auto const kernel   = lookup_kernel<block_size,
                                        KPT,
                                        Chunk,
                                        decltype(in),
                                        decltype(st),
                                        Predicate,
                                        decltype(out),
                                        engine_t,
                                        Emit>;
      auto const tiles =
        cuco::detail::int_div_ceil(n, cuco::detail::index_type{block_size} * KPT);
      launch(kernel,
             persistent_grid(kernel, block_size, tiles),
             block_size,
             stream.get(),
             base,
             window,
             in,
             n,
             st,
             pred,
             out,
             engine,
             emit);
    };
    dispatch_variant<engine_t>(run);
  }

  template <bool Counted,
            typename InputIt,
            typename StencilIt,
            typename Predicate,
            typename Ref,
            typename Action>
  void mutate(InputIt first,
              cuco::detail::index_type n,
              StencilIt stencil,
              Predicate pred,
              size_type* counter,
              Ref ref,
              Action action,
              cuda::stream_ref stream) noexcept
  {
    auto in           = unwrap(first);
    auto st           = unwrap(stencil);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;

    // The blocked path stages slot images and probes with the STORED key, so it is only taken when
    // the batch's key type is the table's key type: a heterogeneous insert key must be hashed and
    // compared in its own type (reference open_addressing_ref_impl.cuh:1506-1521), which only the
    // direct kernels do.
    using input_value_type = typename cuda::std::iterator_traits<decltype(in)>::value_type;
    constexpr bool native_keys =
      std::is_same_v<typename input_key_of<input_value_type, has_payload>::type, key_type>;
    if constexpr (engine_t::single_cas && engine_t::pow2_slot && Action::blockable && native_keys) {
      if (this->fast_path_ok(true) && this->blocking_pays(n)) {
        // false: no scratch memory for the staged batch (or the route kernel could not be
        // configured) - nothing has been launched yet and the direct path below takes the batch
        if (this->blocked_mutate<Counted>(in, n, st, pred, counter, engine, action, stream)) { return; }
      }
    }
    this->direct_mutate<Counted>(in, n, st, pred, counter, engine, action, stream);
  }

  /// The un-regrouped mutation kernels: thread-per-key fast path, else the general fallback.
  template <bool Counted,
            typename InputIt,
            typename StencilIt,
            typename Predicate,
            typename EngineT,
            typename Action>
  void direct_mutate(InputIt in,
                     cuco::detail::index_type n,
                     StencilIt st,
                     Predicate pred,
                     size_type* counter,
                     EngineT const& engine,
                     Action action,
                     cuda::stream_ref stream) noexcept
  {
    using engine_t    = EngineT;
    void* base        = storage_.data();
    auto const window = this->window_bytes();
    if constexpr (engine_t::single_cas && engine_t::pow2_slot) {
      if (this->fast_path_ok(true)) {
        auto run = [&](auto kpt, auto chunk) {
          constexpr int KPT   = decltype(kpt)::value;
          constexpr int Chunk = decltype(chunk)::value;
          auto go             = [&](auto cas_first, auto coherent) {
            constexpr bool CasFirst = decltype(cas_first)::value;
            constexpr auto Policy =
              decltype(coherent)::value ? load_policy::coherent : load_policy::streaming;
            auto const kernel = mutate_kernel<block_size,
                                              KPT,
                                              Chunk,
                                              CasFirst,
                                              Counted,
                                              Policy,
                                              InputIt,
                                              StencilIt,
                                              Predicate,
                                              size_type,
                                              engine_t,
                                              Action>;
            auto const tiles =
              cuco::detail::int_div_ceil(n, cuco::detail::index_type{block_size} * KPT);
            launch(kernel,
                   persistent_grid(kernel, block_size, tiles),
                   block_size,
                   stream.get(),
                   base,
                   window,
                   in,
                   n,
                   st,
                   pred,
                   counter,
                   engine,
                   action);
          };
#if defined(CUCO_B200_TUNABLE)
          auto const& t = tuning();
          if (t.cas_first) {
            t.coherent_loads ? go(std::true_type{}, std::true_type{})
                             : go(std::true_type{}, std::false_type{});
          } else {
            t.coherent_loads ? go(std::false_type{}, std::true_type{})
                             : go(std::false_type{}, std::false_type{});
          }
#else
          go(std::false_type{}, std::false_type{});
#endif
        };
        dispatch_variant<engine_t, true>(run);
        return;
      }
    }
    auto const kernel = generic_mutate_kernel<block_size,
                                              Counted,
                                              InputIt,
                                              StencilIt,
                                              Predicate,
                                              size_type,
                                              engine_t,
                                              Action>;
    launch(kernel, generic_grid(n), block_size, stream.get(), base, window, in, n, st, pred, counter, engine, action);
  }

  /// Should this batch take the L2-blocked path? Auto mode wants a table well beyond L2, a batch
  /// that touches it about once per 64 bytes or denser (otherwise there is no line reuse to win),
  /// and regions that still fit L2 with the region count capped.
  [[nodiscard]] bool blocking_pays(cuco::detail::index_type n) const noexcept
  {
    auto const& t = tuning();
    if (t.blocked == 0 || n <= 0) { return false; }
    auto const bytes = static_cast<std::size_t>(storage_.capacity()) * sizeof(value_type);
    // at most route_max_regions regions, and a region must stay L2-resident while it is probed
    // (measured: 64 MiB windows still run at 0.8x the rate of 16 MiB ones, profiles/r01_hardware_probes.md)
    if (bytes / route_max_regions > (std::size_t{64} << 20)) { return false; }
    if (t.blocked > 0) { return true; }
    return bytes >= t.blocked_min_table && n >= t.blocked_min_elements &&
           static_cast<std::size_t>(n) * 64 >= bytes;
  }

  /// Stream-ordered scratch memory for one bulk call (nullptr if the device is out of memory; the
  /// callers then take a path that needs none). Returned with `scratch_free` on the same stream.
  [[nodiscard]] void* scratch_alloc(std::size_t bytes, cudaStream_t stream) const noexcept
  {
    auto const pool = scratch_pool();
    if (pool == nullptr) { return nullptr; }
    void* p = nullptr;
    if (cudaMallocFromPoolAsync(&p, bytes, pool, stream) != cudaSuccess) {
      cudaGetLastError();  // clear the sticky-free error state of the runtime call
      return nullptr;
    }
    return p;
  }

  /// See `device_scratch_pool()` (bulk_kernels.cuh): one stream-ordered pool per device, shared by all
  /// containers.
  [[nodiscard]] static cudaMemPool_t scratch_pool() noexcept { return device_scratch_pool(); }

  void scratch_free(void* p, cudaStream_t stream) const noexcept
  {
    if (p != nullptr) { cudaFreeAsync(p, stream); }
  }

  /// L2-blocked mutation: route the batch by table region (pass 1), then probe the regions in
  /// order with the region's slots resident in L2 (pass 2). See bulk_kernels.cuh.
  template <bool Counted,
            typename InputIt,
            typename StencilIt,
            typename Predicate,
            typename EngineT,
            typename Action>
  [[nodiscard]] bool blocked_mutate(InputIt in,
                      cuco::detail::index_type n,
                      StencilIt stencil,
                      Predicate pred,
                      size_type* counter,
                      EngineT const& engine,
                      Action action,
                      cuda::stream_ref stream,
                      std::uint64_t first_slot  = 0,
                      std::uint64_t slice_slots = 0)
  {
    using cuco::detail::index_type;
    // region counters and staged positions are 32-bit: larger batches go through in slices
    constexpr index_type slice = index_type{1} << 31;
    if (n > slice) {
      for (index_type done = 0; done < n; done += slice) {
        auto const len = std::min<index_type>(slice, n - done);
        if (!this->blocked_mutate<Counted>(
              in + done, len, stencil + done, pred, counter, engine, action, stream, first_slot, slice_slots)) {
          if (done == 0) { return false; }
          // later slices without scratch memory: the direct kernel takes them (same results)
          this->direct_mutate<Counted>(in + done, n - done, stencil + done, pred, counter, engine, action, stream);
          return true;
        }
      }
      return true;
    }
    auto const& t          = tuning();
    // the slots the batch can hash to: the whole table, or the slice the caller vouches for (batches
    // that arrive grouped by table slice, e.g. from the multi-GPU exchange)
    auto const capacity    = slice_slots != 0 ? slice_slots : static_cast<std::uint64_t>(storage_.capacity());
    auto const table_bytes = capacity * sizeof(value_type);
    auto const num_regions = static_cast<std::uint32_t>(std::min<std::uint64_t>(
      route_max_regions, std::max<std::uint64_t>(2, (table_bytes + t.region_bytes - 1) / t.region_bytes)));
    // expected elements per region plus 1/16 slack and a constant for tiny batches; a multiple of
    // 16 elements so that every segment (and every 128-key chunk of it) starts 16-byte aligned
    auto const mean             = (static_cast<std::uint64_t>(n) + num_regions - 1) / num_regions;
    auto const segment_capacity = static_cast<std::uint32_t>((mean + mean / 16 + 1024 + 15) / 16 * 16);
    auto const staged           = static_cast<std::uint64_t>(num_regions) * segment_capacity;

    std::size_t const counts_bytes = ((num_regions * sizeof(unsigned int)) + 255) / 256 * 256;
    auto* base = static_cast<char*>(
      this->scratch_alloc(counts_bytes + staged * sizeof(value_type), stream.get()));
    if (base == nullptr) { return false; }
    auto* counts   = reinterpret_cast<unsigned int*>(base);
    auto* segments = reinterpret_cast<value_type*>(base + counts_bytes);
    cudaMemsetAsync(counts, 0, num_regions * sizeof(unsigned int), stream.get());

    auto const regions = region_map::over(capacity, num_regions, first_slot);

    constexpr int chunk = EngineT::sector_chunk_slots;
    // ---- pass 1 ----
    bool routed = false;
    if constexpr (std::is_pointer_v<InputIt>) {
      using input_value = std::remove_cv_t<std::remove_pointer_t<InputIt>>;
      if constexpr (std::is_same_v<input_value, value_type>) {
        if (t.blocked_tile_route && (reinterpret_cast<std::uintptr_t>(in) % 16) == 0) {
          auto const kernel = tile_route_kernel<tile_route_block_size,
                                                chunk,
                                                Counted,
                                                StencilIt,
                                                Predicate,
                                                size_type,
                                                EngineT,
                                                Action>;
          auto const smem = tile_route_smem_bytes<tile_route_block_size, value_type>(num_regions);
          if (opt_in_dynamic_smem(kernel, std::size_t{200} << 10)) {
            auto const tiles =
              cuco::detail::int_div_ceil(n, index_type{tile_route_block_size} * route_items_per_thread);
            int per_sm = 0;  // depends on the number of regions (bucket arrays live in shared memory)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tile_route_block_size, smem);
            auto const resident = std::max(1, per_sm) * cuco::detail::multiprocessor_count();
            auto const grid     = static_cast<unsigned>(std::min<index_type>(tiles, resident));
            kernel<<<grid, tile_route_block_size, smem, stream.get()>>>(
              in, n, stencil, pred, segments, counts, region_router{regions}, segment_capacity, counter, engine,
              action, route_spill{});
            routed = true;
          }
        }
      }
    }
    if (!routed) {
      auto const kernel = route_kernel<route_block_size,
                                       chunk,
                                       Counted,
                                       InputIt,
                                       StencilIt,
                                       Predicate,
                                       size_type,
                                       EngineT,
                                       Action>;
      constexpr std::size_t smem = route_smem_bytes<route_block_size, value_type>();
      if (!opt_in_dynamic_smem(kernel, smem)) {
        this->scratch_free(base, stream.get());
        return false;
      }
      auto const tiles =
        cuco::detail::int_div_ceil(n, index_type{route_block_size} * route_items_per_thread);
      auto const grid = static_cast<unsigned>(std::min<index_type>(tiles, index_type{0x7fffffff}));
      kernel<<<grid, route_block_size, smem, stream.get()>>>(
        in, n, stencil, pred, segments, counts, regions, segment_capacity, counter, engine, action);
    }
    // ---- pass 2 ----
    this->probe_segments<Counted>(
      segments, counts, num_regions, 1u, segment_capacity, counter, engine, action, stream, first_slot, slice_slots);
    this->scratch_free(base, stream.get());
    return true;
  }

  /// Pass 2 of the blocked path (also the owner side of an exchanged batch, `sources` = ranks):
  /// probes the staged segments region by region with the region's slots resident in L2.
  template <bool Counted, typename EngineT, typename Action>
  void probe_segments(value_type const* segments,
                      unsigned int const* counts,
                      std::uint32_t num_regions,
                      std::uint32_t sources,
                      std::uint32_t segment_capacity,
                      size_type* counter,
                      EngineT const& engine,
                      Action action,
                      cuda::stream_ref stream,
                      std::uint64_t first_slot  = 0,
                      std::uint64_t slice_slots = 0,
                      std::uint32_t region_begin = 0,
                      std::uint32_t region_count = 0,
                      bool source_major          = false)
  {
    using cuco::detail::index_type;
    auto const& t           = tuning();
    bool const partial      = region_count != 0 || source_major;
    if (region_count == 0) { region_count = num_regions; }
    auto const capacity     = slice_slots != 0 ? slice_slots : static_cast<std::uint64_t>(storage_.capacity());
    auto const table_bytes  = static_cast<std::uint64_t>(storage_.capacity()) * sizeof(value_type);
    auto const region_slots = (capacity + num_regions - 1) / num_regions;
    constexpr int chunk     = EngineT::sector_chunk_slots;

    if (t.blocked_stream_probe && !partial && first_slot == 0 && slice_slots == 0 && segment_capacity % 16 == 0 &&
        (reinterpret_cast<std::uintptr_t>(segments) % 16) == 0) {
      auto run = [&](auto rows_tag, auto blocks_tag) {
        constexpr int Rows      = decltype(rows_tag)::value;
        constexpr int MinBlocks = decltype(blocks_tag)::value;
        auto const kernel =
          stream_mutate_kernel<block_size, Rows, chunk, MinBlocks, Counted, size_type, EngineT, Action>;
        constexpr std::size_t smem = stream_mutate_smem_bytes<block_size, value_type>();
        if (!opt_in_dynamic_smem(kernel, smem)) { return false; }
        stream_layout const layout{
          counts,
          segment_capacity,
          static_cast<std::uint32_t>((segment_capacity + stream_chunk_keys - 1) / stream_chunk_keys),
          num_regions * sources,
          sources,
          region_slots,
          table_bytes,
          (t.blocked_prefetch ? 1u : 0u) | (t.stream_scout ? 2u : 0u)};
        static int const resident = [&] {
          int per_sm = 0;
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_size, smem);
          return std::max(1, per_sm) * cuco::detail::multiprocessor_count();
        }();
        auto const items = static_cast<std::uint64_t>(layout.num_segments) * layout.chunks_per_segment;
        auto const ctas  = (items + block_size / 32 - 1) / (block_size / 32);
        auto const grid  = static_cast<unsigned>(std::max<std::uint64_t>(1, std::min<std::uint64_t>(ctas, resident)));
        // the ticket counter that hands the chunks out in region order
        auto* ticket = static_cast<unsigned long long*>(this->scratch_alloc(sizeof(unsigned long long), stream.get()));
        if (ticket == nullptr) { return false; }
        cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), stream.get());
        kernel<<<grid, block_size, smem, stream.get()>>>(segments, layout, ticket, counter, engine, action);
        this->scratch_free(ticket, stream.get());
        return true;
      };
      bool launched = false;
#if defined(CUCO_B200_TUNABLE)
      switch (t.stream_slots) {
        case 1: launched = run(std::integral_constant<int, 1>{}, std::integral_constant<int, 6>{}); break;
        default: launched = run(std::integral_constant<int, 2>{}, std::integral_constant<int, 4>{}); break;
      }
#else
      launched = run(std::integral_constant<int, 2>{}, std::integral_constant<int, 4>{});
#endif
      if (launched) { return; }
    }

    auto run = [&](auto kpt, auto cas_first) {
      constexpr int KPT       = decltype(kpt)::value;
      constexpr bool CasFirst = decltype(cas_first)::value;
      auto const kernel =
        blocked_mutate_kernel<block_size, KPT, chunk, CasFirst, Counted, size_type, EngineT, Action>;
      auto const tiles_per_segment = static_cast<unsigned>(
        cuco::detail::int_div_ceil(index_type{segment_capacity}, index_type{block_size} * KPT));
      auto const region_bytes = region_slots * sizeof(value_type);
      auto const shares       = static_cast<std::uint64_t>(tiles_per_segment) * sources;
      auto const share        = (region_bytes + shares - 1) / shares;
      blocked_layout const layout{
        counts,
        segment_capacity,
        region_slots,
        table_bytes,
        t.blocked_prefetch ? static_cast<std::uint32_t>((share + 127) / 128 * 128) : 0u,
        sources,
        first_slot,
        region_begin,
        num_regions,
        source_major ? 1u : 0u};
      kernel<<<dim3{tiles_per_segment, region_count * sources}, block_size, 0, stream.get()>>>(
        segments, layout, counter, engine, action);
    };
#if defined(CUCO_B200_TUNABLE)
    auto with_kpt = [&](auto cas_first) {
      switch (t.blocked_keys_per_thread) {
        case 1: run(std::integral_constant<int, 1>{}, cas_first); break;
        case 4: run(std::integral_constant<int, 4>{}, cas_first); break;
        default: run(std::integral_constant<int, 2>{}, cas_first); break;
      }
    };
    t.blocked_cas_first ? with_kpt(std::true_type{}) : with_kpt(std::false_type{});
#else
    run(std::integral_constant<int, 4>{}, std::true_type{});
#endif
  }

  // ------------------------------------------------------------------------------------------
  // exchange path of hash-partitioned tables (see bulk_kernels.cuh, "Exchange path")
  // ------------------------------------------------------------------------------------------
 public:
  /// Buffer geometry for batches of at most `n_max` elements per rank over `num_ranks` shards.
  struct exchange_plan {
    std::uint32_t num_regions;       ///< R: L2 regions per shard (num_ranks * R <= 1024)
    std::uint32_t segment_capacity;  ///< elements per (owner, region, source) segment
    std::uint32_t spill_capacity;    ///< elements of the local spill list
  };

  [[nodiscard]] exchange_plan plan_exchange(cuco::detail::index_type n_max, int num_ranks) const
  {
    CUCO_EXPECTS(num_ranks >= 1 && num_ranks <= exchange_max_ranks, "unsupported number of ranks");
    CUCO_EXPECTS(n_max >= 0 && n_max < (cuco::detail::index_type{1} << 32),
                 "exchange batches are limited to 2^32 - 1 elements per rank");
    auto const& t          = tuning();
    auto const table_bytes = static_cast<std::uint64_t>(storage_.capacity()) * sizeof(value_type);
    // Fine-grained exchange (the router also groups by L2 region of the owner's shard) while the
    // runs stay long enough for NVLink: a tile of 4096 elements over P * R buckets. Beyond 512
    // buckets the router groups by owner only (16 KB runs) and the owner regroups locally.
    auto const local_regions = std::max<std::uint64_t>(1, (table_bytes + t.region_bytes - 1) / t.region_bytes);
    auto const regions       = local_regions * num_ranks <= 512 ? local_regions : std::uint64_t{1};
    auto const segments = regions * static_cast<std::uint64_t>(num_ranks);
    auto const mean     = (static_cast<std::uint64_t>(n_max) + segments - 1) / segments;
    return exchange_plan{static_cast<std::uint32_t>(regions),
                         static_cast<std::uint32_t>((mean + mean / 16 + 1024 + 15) / 16 * 16),
                         static_cast<std::uint32_t>(
                           std::max<std::uint64_t>(65536, static_cast<std::uint64_t>(n_max) / 8))};
  }

  /// Source side: groups [first, first + n) by (owner, region) and stores it into the owners'
  /// segment buffers; then publishes the fill counts and this rank's spill count to the peers.
  /// The caller zeroes nothing: `counts_local` and `spill_count` are reset here, on `stream`.
  template <bool KeysOnly, typename InputIt, typename Ref>
  void exchange_route_async(InputIt first,
                            cuco::detail::index_type n,
                            exchange_plan plan,
                            int num_ranks,
                            int my_rank,
                            std::uint64_t salt,
                            exchange_peers segments,
                            exchange_peers counts_recv,
                            exchange_peers spill_flags,
                            unsigned int* counts_local,
                            std::uint32_t* position_local,
                            void* spill,
                            std::uint32_t* spill_index,
                            unsigned int* spill_count,
                            Ref ref,
                            cuda::stream_ref stream,
                            bool staging = false)
  {
    using cuco::detail::index_type;
    auto in           = unwrap(first);
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    using elem_type   = std::conditional_t<KeysOnly, key_type, value_type>;
    auto const capacity = static_cast<std::uint64_t>(storage_.capacity());
    // staging: `segments.base[owner]` is the owner's block [region][cap] of a LOCAL buffer that the
    // copy engines deliver later; otherwise it is the owner's own buffer, region-major, source-minor
    exchange_geometry const geometry{static_cast<std::uint32_t>(num_ranks),
                                     static_cast<std::uint32_t>(my_rank),
                                     plan.num_regions,
                                     plan.segment_capacity,
                                     salt,
                                     staging ? 1u : static_cast<std::uint32_t>(num_ranks),
                                     staging ? 0u : static_cast<std::uint32_t>(my_rank)};
    auto const buckets = static_cast<std::size_t>(num_ranks) * plan.num_regions;
    CUCO_CUDA_TRY(cudaMemsetAsync(counts_local, 0, buckets * sizeof(unsigned int), stream.get()));
    CUCO_CUDA_TRY(cudaMemsetAsync(spill_count, 0, sizeof(unsigned int), stream.get()));
    if constexpr (std::is_pointer_v<decltype(in)>) {
      // staging a contiguous, 16-byte aligned array of slot images (mutations) or keys (lookups): the
      // bulk-copy fed persistent router of the single-GPU blocked path, with (owner, bucket) buckets, a
      // spill list and - for lookups - the staged position of every input element
      using input_value = std::remove_cv_t<std::remove_pointer_t<decltype(in)>>;
      if constexpr (std::is_same_v<input_value, elem_type> && sizeof(elem_type) >= 8) {
        auto const& t = tuning();
        if (staging && n > 0 && t.blocked_tile_route && (reinterpret_cast<std::uintptr_t>(in) % 16) == 0 &&
            buckets <= 4096) {
          using stencil_t   = thrust::constant_iterator<bool>;
          auto const kernel = tile_route_kernel<tile_route_block_size,
                                                engine_t::sector_chunk_slots,
                                                false,
                                                stencil_t,
                                                always_true,
                                                size_type,
                                                engine_t,
                                                action_insert,
                                                exchange_router,
                                                elem_type,
                                                KeysOnly>;
          auto const smem = tile_route_smem_bytes<tile_route_block_size, elem_type, KeysOnly>(
            static_cast<std::uint32_t>(buckets));
          if (opt_in_dynamic_smem(kernel, std::size_t{200} << 10) && smem <= (std::size_t{200} << 10)) {
            auto const tiles =
              cuco::detail::int_div_ceil(n, index_type{tile_route_block_size} * route_items_per_thread);
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tile_route_block_size, smem);
            auto const resident = std::max(1, per_sm) * cuco::detail::multiprocessor_count();
            auto const grid     = static_cast<unsigned>(std::min<index_type>(tiles, resident));
            exchange_router const router{region_map::over(capacity, plan.num_regions),
                                         static_cast<std::uint32_t>(num_ranks),
                                         salt};
            kernel<<<grid, tile_route_block_size, smem, stream.get()>>>(
              in, n, stencil_t{true}, always_true{}, static_cast<elem_type*>(segments.base[0]), counts_local, router,
              plan.segment_capacity, static_cast<size_type*>(nullptr), engine, action_insert{},
              route_spill{spill, spill_count, plan.spill_capacity, position_local, spill_index});
            return;
          }
        }
      }
    }
    if (n > 0) {
      auto const regions = region_map::over(capacity, plan.num_regions);
      auto const kernel  = exchange_route_kernel<route_block_size, KeysOnly, decltype(in), engine_t>;
      constexpr std::size_t smem = exchange_smem_bytes<route_block_size, elem_type, KeysOnly>();
      CUCO_EXPECTS(opt_in_dynamic_smem(kernel, smem),
                   "the device refused the dynamic shared memory the exchange router needs");
      auto const tiles = cuco::detail::int_div_ceil(n, index_type{route_block_size} * route_items_per_thread);
      auto const grid  = static_cast<unsigned>(std::min<index_type>(tiles, index_type{0x7fffffff}));
      kernel<<<grid, route_block_size, smem, stream.get()>>>(in,
                                                             n,
                                                             segments,
                                                             counts_local,
                                                             position_local,
                                                             spill,
                                                             spill_index,
                                                             spill_count,
                                                             plan.spill_capacity,
                                                             regions,
                                                             geometry,
                                                             engine);
    }
    if (staging) { return; }  // counts and spill flags travel with the staged segments
    auto const publish_grid = static_cast<unsigned>((buckets + 255) / 256);
    exchange_publish_kernel<<<publish_grid, 256, 0, stream.get()>>>(
      counts_local, spill_count, counts_recv, spill_flags, geometry);
  }

  /// Geometry of the staged exchange: every rank groups a batch of at most `n_max` elements by
  /// (owner, table slice) - `groups` slices per shard - into a local buffer [owner][slice][cap].
  [[nodiscard]] exchange_plan plan_stage(cuco::detail::index_type n_max, int num_ranks, int groups) const
  {
    CUCO_EXPECTS(num_ranks >= 1 && num_ranks <= exchange_max_ranks, "unsupported number of ranks");
    CUCO_EXPECTS(groups >= 1 && groups * num_ranks <= route_max_regions, "too many (owner, slice) buckets");
    CUCO_EXPECTS(n_max >= 0 && n_max < (cuco::detail::index_type{1} << 32),
                 "exchange batches are limited to 2^32 - 1 elements per rank");
    auto const segments = static_cast<std::uint64_t>(groups) * static_cast<std::uint64_t>(num_ranks);
    auto const mean     = (static_cast<std::uint64_t>(n_max) + segments - 1) / segments;
    // owner and slice are both hash-uniform: 1/32 of slack is > 20 sigma from 2^17 elements per bucket up
    return exchange_plan{static_cast<std::uint32_t>(groups),
                         static_cast<std::uint32_t>((mean + mean / 32 + 1024 + 15) / 16 * 16),
                         static_cast<std::uint32_t>(
                           std::max<std::uint64_t>(65536, static_cast<std::uint64_t>(n_max) / 8))};
  }

  /// The slots [first, first + count) of table slice `group` of `groups` (same map on every rank).
  [[nodiscard]] std::pair<std::uint64_t, std::uint64_t> slice_of(std::uint32_t group, std::uint32_t groups) const noexcept
  {
    auto const capacity = static_cast<std::uint64_t>(storage_.capacity());
    auto const map      = region_map::over(capacity, groups);
    auto const first    = map.region_begin(group);
    auto const last     = group + 1 == groups ? capacity : map.region_begin(group + 1);
    return {first, last - first};
  }

  /// Owner side of the staged exchange: applies the `num_ranks` received segments of ONE table slice
  /// (contiguous [source][cap], fill counts in `counts_recv[source]`) - an ordinary gappy batch whose
  /// keys all hash into that slice, so the L2-blocked path divides just the slice into regions.
  template <typename Ref, typename Action>
  void exchange_apply_async(value_type const* segments,
                            unsigned int const* counts_recv,
                            std::uint32_t segment_capacity,
                            int num_ranks,
                            std::uint32_t group,
                            std::uint32_t groups,
                            Ref ref,
                            Action action,
                            cuda::stream_ref stream)
  {
    using cuco::detail::index_type;
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    if constexpr (!(engine_t::single_cas && engine_t::pow2_slot)) {
      CUCO_FAIL("the exchange path needs slots that one CAS can claim (4, 8 or packed 16 bytes)");
    } else {
      CUCO_EXPECTS(this->fast_path_ok(true), "the exchange path needs container-owned storage without tombstones");
      auto const virtual_n     = index_type{segment_capacity} * num_ranks;
      auto const live          = segment_live{counts_recv, segment_capacity};
      auto const [first, span] = this->slice_of(group, groups);
      auto const slice_bytes   = span * sizeof(value_type);
      auto const& t            = tuning();
      auto const in            = segments;
      auto const index         = thrust::counting_iterator<index_type>{0};
      if constexpr (Action::blockable) {
        // same rule as blocking_pays(), applied to the slice the batch is confined to
        bool const dense = static_cast<std::uint64_t>(virtual_n) * 64 >= slice_bytes;
        bool const pays  = t.blocked > 0 || (t.blocked < 0 && slice_bytes >= (std::size_t{2} * t.region_bytes) &&
                                             virtual_n >= (index_type{1} << 16) && dense);
        if (pays && slice_bytes / route_max_regions <= (std::size_t{64} << 20)) {
          if (this->template blocked_mutate<false>(in, virtual_n, index, live, static_cast<size_type*>(nullptr),
                                                   engine, action, stream, first, span)) {
            return;
          }
        }
      }
      this->template direct_mutate<false>(in, virtual_n, index, live, static_cast<size_type*>(nullptr), engine,
                                          action, stream);
    }
  }

  /// Owner side of a routed mutation: probes the received segments region by region.
  template <typename Ref, typename Action>
  void exchange_mutate_async(value_type const* segments,
                             unsigned int const* counts_recv,
                             exchange_plan plan,
                             int num_ranks,
                             Ref ref,
                             Action action,
                             cuda::stream_ref stream)
  {
    using cuco::detail::index_type;
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    if constexpr (!(engine_t::single_cas && engine_t::pow2_slot)) {
      CUCO_FAIL("the exchange path needs slots that one CAS can claim (4, 8 or packed 16 bytes)");
    } else {
    CUCO_EXPECTS(this->fast_path_ok(true), "the exchange path needs container-owned storage without tombstones");
    if (plan.num_regions == 1 && num_ranks > 1) {
      // owner-only routing: the received segments are an ordinary (gappy) batch; the bulk path
      // regroups it by L2 region locally when that pays
      auto const virtual_n = index_type{plan.segment_capacity} * num_ranks;
      this->template mutate<false>(segments,
                                   virtual_n,
                                   thrust::counting_iterator<index_type>{0},
                                   segment_live{counts_recv, plan.segment_capacity},
                                   static_cast<size_type*>(nullptr),
                                   ref,
                                   action,
                                   stream);
      return;
    }
    this->template probe_segments<false>(segments,
                                         counts_recv,
                                         plan.num_regions,
                                         static_cast<std::uint32_t>(num_ranks),
                                         plan.segment_capacity,
                                         static_cast<size_type*>(nullptr),
                                         engine,
                                         action,
                                         stream);
    }
  }

  /// Owner side of the staged exchange, fine mode: the sources already grouped their batches by the
  /// L2 regions of this shard (`num_regions` of them), segments arrive source-major
  /// [source][region][cap]; probes regions [region_begin, region_begin + region_count) with the
  /// region's slots resident in L2 - no local regrouping pass.
  template <typename Ref, typename Action>
  void exchange_probe_async(value_type const* segments,
                            unsigned int const* counts_recv,
                            std::uint32_t num_regions,
                            std::uint32_t segment_capacity,
                            int num_ranks,
                            std::uint32_t region_begin,
                            std::uint32_t region_count,
                            Ref ref,
                            Action action,
                            cuda::stream_ref stream)
  {
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    if constexpr (!(engine_t::single_cas && engine_t::pow2_slot && Action::blockable)) {
      CUCO_FAIL("the exchange path needs slots that one CAS can claim (4, 8 or packed 16 bytes)");
    } else {
      CUCO_EXPECTS(this->fast_path_ok(true), "the exchange path needs container-owned storage without tombstones");
      CUCO_EXPECTS(region_begin + region_count <= num_regions && region_count > 0, "region range out of bounds");
      this->template probe_segments<false>(segments,
                                           counts_recv,
                                           num_regions,
                                           static_cast<std::uint32_t>(num_ranks),
                                           segment_capacity,
                                           static_cast<size_type*>(nullptr),
                                           engine,
                                           action,
                                           stream,
                                           0,
                                           0,
                                           region_begin,
                                           region_count,
                                           true);
    }
  }

  /// Owner side of a staged lookup: the received keys [source][cap] (fill counts per source) are an
  /// ordinary gappy batch for the single-GPU lookup kernel, answered in arrival order into a local
  /// buffer of the same shape (the copy engines return it to the sources).
  template <typename Result, typename Ref, typename Emit>
  void exchange_lookup_local_async(key_type const* segments,
                                   unsigned int const* counts_recv,
                                   Result* results,
                                   std::uint32_t segment_capacity,
                                   int num_ranks,
                                   Ref ref,
                                   Emit emit,
                                   cuda::stream_ref stream)
  {
    using cuco::detail::index_type;
    this->lookup(segments,
                 index_type{segment_capacity} * num_ranks,
                 thrust::counting_iterator<index_type>{0},
                 segment_live{counts_recv, segment_capacity},
                 results,
                 ref,
                 emit,
                 stream);
  }

  /// L2 regions a shard is divided into for the fine mode of the staged exchange (the same rule as the
  /// single-GPU blocked path), or 0 when `num_ranks` x regions would exceed the router's bucket limit.
  [[nodiscard]] std::uint32_t exchange_fine_regions(int num_ranks) const noexcept
  {
    auto const& t          = tuning();
    auto const table_bytes = static_cast<std::uint64_t>(storage_.capacity()) * sizeof(value_type);
    // Budget of (owner, region) buckets for the source's router. Measured on B200 (tools/stage_probe.py,
    // profiles/r02_stage_probe.jsonl, 100 M pairs): 0.75 - 0.79 ms up to ~200 buckets, 1.32 ms at 400,
    // 2.2 ms at 800 - a tile of 2048 elements then holds 2 - 3 per bucket and the copy-out degenerates to
    // one 16-byte store per sector. Beyond the budget the coarse mode (source groups by owner and slice,
    // 0.76 ms; the owner regroups, 0.75 ms) is cheaper.
    std::uint64_t budget = 256;
    if (char const* s = std::getenv("CUCO_B200_FINE_BUCKETS")) { budget = static_cast<std::uint64_t>(std::max(1, std::atoi(s))); }
    std::uint64_t const limit =
      std::min<std::uint64_t>(route_max_regions, budget) / static_cast<std::uint64_t>(num_ranks);
    // 16 MiB regions when they fit the bucket budget, else up to 64 MiB ones (0.8x the probe rate)
    for (std::uint64_t bytes = t.region_bytes; bytes <= (std::uint64_t{64} << 20); bytes *= 2) {
      auto const regions = std::max<std::uint64_t>(2, (table_bytes + bytes - 1) / bytes);
      if (regions <= limit) { return static_cast<std::uint32_t>(regions); }
    }
    return 0;
  }

  /// Owner side of a routed lookup: results go straight into the sources' result buffers.
  template <typename Result, typename Ref, typename Emit>
  void exchange_lookup_async(key_type const* segments,
                             unsigned int const* counts_recv,
                             exchange_peers results,
                             exchange_plan plan,
                             int num_ranks,
                             int my_rank,
                             Ref ref,
                             Emit emit,
                             cuda::stream_ref stream,
                             bool local_results = false) const
  {
    using cuco::detail::index_type;
    auto const engine = ref.engine();
    using engine_t    = std::decay_t<decltype(engine)>;
    CUCO_EXPECTS(this->fast_path_ok(false), "the exchange path needs container-owned storage");
    constexpr int chunk = engine_t::sector_chunk_slots;
    // local_results: `results.base[source]` is that source's block of a LOCAL buffer (the copy
    // engines return it); otherwise the source's own buffer, written through peer stores
    exchange_geometry const geometry{static_cast<std::uint32_t>(num_ranks),
                                     static_cast<std::uint32_t>(my_rank),
                                     plan.num_regions,
                                     plan.segment_capacity,
                                     0,
                                     static_cast<std::uint32_t>(num_ranks),
                                     local_results ? 0u : static_cast<std::uint32_t>(my_rank)};
    auto run = [&](auto kpt_tag) {
      constexpr int kpt            = decltype(kpt_tag)::value;
      auto const tiles_per_segment = static_cast<unsigned>(
        cuco::detail::int_div_ceil(index_type{plan.segment_capacity}, index_type{block_size} * kpt));
      exchange_lookup_kernel<block_size, kpt, chunk, Result>
        <<<dim3{tiles_per_segment, plan.num_regions * static_cast<unsigned>(num_ranks)},
           block_size,
           0,
           stream.get()>>>(segments, counts_recv, results, geometry, engine, emit);
    };
    switch (tuning().exchange_lookup_keys_per_thread) {
      case 1: run(std::integral_constant<int, 1>{}); break;
      case 4: run(std::integral_constant<int, 4>{}); break;
      default: run(std::integral_constant<int, 2>{}); break;
    }
  }

  /// Source side of a routed lookup: out[i] = result of key i, for the n keys this rank routed.
  template <typename Result, typename OutputIt>
  static void exchange_unpermute_async(Result const* results,
                                       std::uint32_t const* position_local,
                                       cuco::detail::index_type n,
                                       OutputIt out,
                                       cuda::stream_ref stream)
  {
    if (n == 0) { return; }
    auto const blocks = cuco::detail::int_div_ceil(n, cuco::detail::index_type{256} * 4);
    exchange_unpermute_kernel<<<static_cast<unsigned>(std::min<cuco::detail::index_type>(blocks, 0x7fffffff)),
                                256,
                                0,
                                stream.get()>>>(results, position_local, n, unwrap(out));
  }

  [[nodiscard]] auto make_find_emit() const noexcept
  {
    return emit_found<engine_type>{empty_slot_sentinel_};
  }

 private:
  /// Picks the (keys per thread, chunk width) instantiation.
  template <typename EngineT, bool Mutating = false, typename Run>
  static void dispatch_variant(Run&& run)
  {
    constexpr int sector = EngineT::sector_chunk_slots;
#if defined(CUCO_B200_TUNABLE)
    constexpr int window = EngineT::window_chunk_slots;
    auto const& t        = tuning();
    auto with_chunk      = [&](auto kpt) {
      if (t.sector_chunks || window == sector) {
        run(kpt, std::integral_constant<int, sector>{});
      } else {
        run(kpt, std::integral_constant<int, window>{});
      }
    };
    switch (Mutating ? t.mutate_keys_per_thread : t.keys_per_thread) {
      case 1: with_chunk(std::integral_constant<int, 1>{}); break;
      case 4: with_chunk(std::integral_constant<int, 4>{}); break;
      default: with_chunk(std::integral_constant<int, 2>{}); break;
    }
#else
    run(std::integral_constant<int, Mutating ? 1 : 2>{}, std::integral_constant<int, sector>{});
#endif
  }

  /// Device counter of ONE bulk call, taken from the stream-ordered pool and zeroed on `stream`.
  size_type* zeroed_counter(cuda::stream_ref stream) const
  {
    auto* counter = static_cast<size_type*>(this->scratch_alloc(sizeof(size_type), stream.get()));
    if (counter == nullptr) { CUCO_CUDA_TRY(cudaErrorMemoryAllocation); }
    CUCO_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(size_type), stream.get()));
    return counter;
  }

  /// Reads the call's counter back, returns it to the pool and waits for the stream; also the place
  /// where a failed launch of the synchronous entry points surfaces as cuco::cuda_error.
  size_type read_counter(size_type* counter, cuda::stream_ref stream) const
  {
    size_type host{};
    auto const launched = cudaGetLastError();
    auto const copied =
      cudaMemcpyAsync(&host, counter, sizeof(size_type), cudaMemcpyDeviceToHost, stream.get());
    this->scratch_free(counter, stream.get());
    CUCO_CUDA_TRY(launched);
    CUCO_CUDA_TRY(copied);
    stream.wait();
    return host;
  }

 protected:
  value_type empty_slot_sentinel_;
  key_type erased_key_sentinel_;
  key_equal predicate_;
  probing_scheme_type probing_scheme_;
  storage_type storage_;
  /// Stream-ordered pool behind the per-call counter and the staging buffer of the blocked path.
  /// Every bulk call takes its own allocation ON ITS STREAM and returns it there, so calls on
  /// different streams never share scratch state (the reference allocates a counter per call too,
  /// impl.cuh:337-347); the pool keeps freed memory (release threshold = max), so a steady stream of
  /// calls costs no cudaMalloc after the first.
};

}  // namespace cuco::b200
