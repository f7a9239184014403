/* cuco_b200.h — C ABI of the b200-native open-addressing hash table.
 *
 * The reference (NVIDIA/cuCollections) is a header-only C++ template library with no FFI; its
 * boundary for this path is the `cuco::static_map` / `cuco::static_set` class templates
 * (reference include/cuco/static_map.cuh:88-986, include/cuco/static_set.cuh:82-798; next row:
 * `cuco::static_multiset`, include/cuco/static_multiset.cuh:81-729, and
 * `cuco::experimental::static_multimap`, include/cuco/static_multimap.cuh:45-549), which this
 * repository re-implements under include/cuco/. This header is the plain-C view of explicit
 * instantiations of that surface, so that non-C++ hosts (ctypes, cgo, JNI, ...) and the parity /
 * benchmark harness can drive it. Every entry point names the reference member it forwards to.
 *
 * The same shim source (cucollections_b200/csrc/cabi_*.cu) is compiled twice:
 *   libcuco_b200.so        against include/cuco           (this implementation)
 *   oracle/_ref/libcuco_ref.so against /root/reference/include (cuco's own sm_100a build)
 * so both run through identical glue and can be compared call for call.
 *
 * Conventions: all pointers are device pointers unless stated otherwise; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream); keys/values/sentinels are passed as
 * int64_t and narrowed to the table's key/payload type; functions return 0 on success and a
 * non-zero code otherwise, with a message available from cuco_b200_last_error() (thread local).
 * "_async" behaviour: calls that do not return a count or size only enqueue work on `stream`.
 */
#ifndef CUCO_B200_H
#define CUCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cuco_b200_table cuco_b200_table; /* opaque */

/* Explicit instantiations. cg = cooperative-group size of the probing scheme (bucket width in
 * windows), w = slots per window (cuco::storage<w>). Hash is cuco::default_hash_function
 * (xxhash_32) unless noted. */
enum cuco_b200_kind {
  CUCO_B200_SET_I32_DH4       = 0, /* static_set<int32>            double_hashing<4> w1 (class default) */
  CUCO_B200_MAP_I64_LP1       = 1, /* static_map<int64,int64>      linear_probing<1> w1 */
  CUCO_B200_MAP_I64_DH8       = 2, /* static_map<int64,int64>      double_hashing<8> w1 */
  CUCO_B200_MAP_I32_LP4       = 3, /* static_map<int32,int32>      linear_probing<4> w1 (class default) */
  CUCO_B200_MAP_I64_LP4       = 4, /* static_map<int64,int64>      linear_probing<4> w1 (class default) */
  CUCO_B200_SET_I64_DH4       = 5, /* static_set<int64>            double_hashing<4> w1 (class default) */
  CUCO_B200_MAP_I64_LP1_W2    = 6, /* static_map<int64,int64>      linear_probing<1> w2 (window = 32 B sector) */
  CUCO_B200_MAP_I32_DH2_W2_MM = 7, /* static_map<int32,int32>      double_hashing<2, murmurhash3_32> w2 (reference test matrix) */
  CUCO_B200_MAP_I32I64_LP1    = 8, /* static_map<int32,int64>      linear_probing<1> w1 (padded 16 B slot: two-step claim path) */
  CUCO_B200_MAP_I64_DH8_X64   = 9, /* static_map<int64,int64>      double_hashing<8, xxhash_64> w1 (tables near 2^32 windows) */
  CUCO_B200_MULTISET_I32_DH4_W2 = 10, /* static_multiset<int32>     double_hashing<4> w2 (class default) */
  CUCO_B200_MULTISET_I64_LP1_W2 = 11, /* static_multiset<int64>     linear_probing<1> w2 */
  CUCO_B200_MULTIMAP_I64_LP4    = 12, /* experimental::static_multimap<int64,int64> linear_probing<4> w1 (class default):
                                         insert, insert_if, contains, contains_if, count only */
  CUCO_B200_MAP_I64_LP1_X64     = 13, /* static_map<int64,int64>      linear_probing<1, xxhash_64> w1: the hash-partitioned
                                         benchmark table; a 32-bit hash folds unevenly onto shards of 1 - 4 G slots */
  CUCO_B200_NUM_KINDS         = 14
};

/* Reduction selector for cuco_b200_insert_or_apply (cuco::reduce::plus / min / max,
 * reference include/cuco/utility/reduction_functors.cuh:25-82). */
enum cuco_b200_reduce_op { CUCO_B200_PLUS = 0, CUCO_B200_MIN = 1, CUCO_B200_MAX = 2 };

/* "native" for this implementation, "reference" for the cuco build of the same shim. */
const char* cuco_b200_build_info(void);
const char* cuco_b200_last_error(void);

/* Constructors (reference static_map.cuh:160-260 / static_set.cuh:148-240).
 *   load_factor == 0 : `size` is a capacity           -> ctor(capacity, sentinels...)
 *   load_factor  > 0 : `size` is a number of keys     -> ctor(n, desired_load_factor, sentinels...)
 *   has_erased  != 0 : ctor(capacity, sentinels..., erased_key)  (load_factor must be 0)
 * Errors: load factor outside (0,1], erased == empty, extent too large, allocation failure. */
int cuco_b200_create(int kind,
                     int64_t size,
                     double load_factor,
                     int64_t empty_key,
                     int64_t empty_value,
                     int has_erased,
                     int64_t erased_key,
                     void* stream,
                     cuco_b200_table** out);
int cuco_b200_destroy(cuco_b200_table* t);

int cuco_b200_kind_of(const cuco_b200_table* t);
int cuco_b200_key_bytes(const cuco_b200_table* t);   /* 4 or 8 */
int cuco_b200_value_bytes(const cuco_b200_table* t); /* 0 for sets */
int64_t cuco_b200_capacity(const cuco_b200_table* t); /* capacity()  */
int cuco_b200_size(cuco_b200_table* t, void* stream, int64_t* out); /* size(stream), synchronises */
int cuco_b200_clear(cuco_b200_table* t, void* stream);              /* clear_async(stream) */

/* insert / insert_async (static_map.cuh:302,316; static_set.cuh:272,286).
 * Maps: `values == NULL` means `keys` is an array of cuco::pair<Key,T> (AoS, the layout the
 * reference benchmarks use); otherwise keys and values are separate arrays (exercises the generic
 * iterator path through a transform iterator). Sets ignore `values`.
 * `num_inserted == NULL` -> insert_async; else synchronous insert and *num_inserted (host) = #new keys. */
int cuco_b200_insert(cuco_b200_table* t,
                     const void* keys,
                     const void* values,
                     int64_t n,
                     void* stream,
                     int64_t* num_inserted);

/* insert_if / insert_if_async with stencil of bytes, predicate "stencil[i] != 0"
 * (static_map.cuh:344,373). */
int cuco_b200_insert_if(cuco_b200_table* t,
                        const void* keys,
                        const void* values,
                        const uint8_t* stencil,
                        int64_t n,
                        void* stream,
                        int64_t* num_inserted);

/* find_async (static_map.cuh:765; static_set.cuh:588): out[i] = payload (maps) / stored key (sets),
 * or the empty value / empty key sentinel. `out` has the payload (maps) or key (sets) type. */
int cuco_b200_find(cuco_b200_table* t, const void* keys, void* out, int64_t n, void* stream);

/* contains_async (static_map.cuh:661; static_set.cuh:484): out[i] = 0/1 as bool bytes. */
int cuco_b200_contains(cuco_b200_table* t, const void* keys, uint8_t* out, int64_t n, void* stream);

/* contains_if_async (static_map.cuh:722): out[i] = stencil[i] != 0 ? contains(keys[i]) : false. */
int cuco_b200_contains_if(cuco_b200_table* t,
                          const void* keys,
                          const uint8_t* stencil,
                          uint8_t* out,
                          int64_t n,
                          void* stream);

/* insert_and_find_async (static_map.cuh:395; static_set.cuh:365): found[i] = resident payload/key,
 * inserted[i] = this element created the entry. `keys`/`values` as in cuco_b200_insert. */
int cuco_b200_insert_and_find(cuco_b200_table* t,
                              const void* keys,
                              const void* values,
                              void* found,
                              uint8_t* inserted,
                              int64_t n,
                              void* stream);

/* insert_or_assign_async (static_map.cuh:467). Maps only. */
int cuco_b200_insert_or_assign(
  cuco_b200_table* t, const void* keys, const void* values, int64_t n, void* stream);

/* insert_or_apply_async (static_map.cuh:541 without init, :573 with init). Maps only. */
int cuco_b200_insert_or_apply(cuco_b200_table* t,
                              const void* keys,
                              const void* values,
                              int64_t n,
                              int reduce_op,
                              int has_init,
                              int64_t init,
                              void* stream);

/* erase_async (static_map.cuh:620; static_set.cuh:449). Requires an erased-key sentinel. */
int cuco_b200_erase(cuco_b200_table* t, const void* keys, int64_t n, void* stream);

/* retrieve_all (static_map.cuh:893; static_set.cuh:682): unordered dump; *n_out (host) = count.
 * `values_out` is ignored for sets. Outputs must hold `capacity` elements. Synchronises. */
int cuco_b200_retrieve_all(
  cuco_b200_table* t, void* keys_out, void* values_out, int64_t* n_out, void* stream);

/* count / count_outer (static_multiset.cuh:615,661; experimental::static_multimap::count,
 * static_multimap.cuh:469): *out (host) = total number of stored elements matching the n keys;
 * outer != 0 (multisets only) counts a key without matches as one. Multisets and multimaps only.
 * Synchronises. */
int cuco_b200_count(
  cuco_b200_table* t, const void* keys, int64_t n, int outer, void* stream, int64_t* out);

/* retrieve (static_set.cuh:620; static_multiset.cuh:506) / retrieve_outer (static_multiset.cuh:593):
 * for every key and every stored element equal to it writes the key to probe_out and the element to
 * match_out (same position, unspecified order); outer != 0 (multisets only) adds {key, empty key}
 * for keys without matches. *n_out (host) = rows written; size the outputs with cuco_b200_count
 * (multisets) or n (sets). Sets and multisets only. Synchronises. */
int cuco_b200_retrieve(cuco_b200_table* t,
                       const void* keys,
                       int64_t n,
                       int outer,
                       void* probe_out,
                       void* match_out,
                       int64_t* n_out,
                       void* stream);

/* rehash (static_map.cuh:911,931): capacity < 0 keeps the current extent. Synchronises. */
int cuco_b200_rehash(cuco_b200_table* t, int64_t capacity, void* stream);

/* ---- host-buffer entry points ------------------------------------------------------------------
 * Same operations as above with HOST pointers (pinned memory for full PCIe speed; pageable memory
 * works but copies synchronously). There is no reference counterpart - a cuco user copies the batch
 * to the device, calls the bulk API and copies the results back; these calls do exactly that, cut
 * into chunks (CUCO_B200_HOST_CHUNK elements, default 4 Mi) so that the upload of chunk i+1, the
 * table kernels on chunk i and the download of chunk i-1 overlap. Stream-ordered on `stream`: work
 * queued before the call is seen, and the results are in host memory once `stream` reaches the end
 * of the call. `host_values == NULL` for maps means AoS pairs, as in cuco_b200_insert. */
int cuco_b200_insert_host(
  cuco_b200_table* t, const void* host_keys, const void* host_values, int64_t n, void* stream);
int cuco_b200_find_host(
  cuco_b200_table* t, const void* host_keys, void* host_out, int64_t n, void* stream);
int cuco_b200_contains_host(
  cuco_b200_table* t, const void* host_keys, uint8_t* host_out, int64_t n, void* stream);
int cuco_b200_insert_and_find_host(cuco_b200_table* t,
                                   const void* host_keys,
                                   const void* host_values,
                                   void* host_found,
                                   uint8_t* host_inserted,
                                   int64_t n,
                                   void* stream);

/* Launch tuning of the native build (no-op returning 1 in the reference build).
 * keys_per_thread in {1,2,4} sets lookups and mutations alike, or L + 10*M sets them separately
 * (12 = lookups 2, mutations 1); waves = 0 launches one CTA per tile, k > 0 a persistent grid of
 * k waves; the others are booleans. Negative = leave unchanged. */
int cuco_b200_set_tuning(int keys_per_thread,
                         int cas_first,
                         int sector_chunks,
                         int waves,
                         int force_generic,
                         int l2_window,
                         int coherent_loads);

/* L2-blocked mutation control of the native build: mode -1 auto (tables much larger than L2),
 * 0 off, 1 always; region_mib = size of the table slice kept L2-resident (0: unchanged; negative: the
 * size in KiB, so tests can exercise many regions on small tables). */
int cuco_b200_set_blocking(int mode, int region_mib);
/* Pass 2 of the blocked path: probes in flight per thread (1, 2, 4), whether a probe starts with the
 * CAS, whether the next region is prefetched into L2. Negative / other = leave unchanged. */
int cuco_b200_set_blocking_variant(int keys_per_thread, int cas_first, int prefetch);
/* Kernels of the blocked path: tile_route != 0 -> pass 1 is the persistent router fed by the bulk-copy
 * engine (contiguous, 16-byte aligned batches of slot images; anything else keeps the round-1
 * router); stream_probe != 0 -> pass 2 is the warp-persistent refilling probe stream (else the
 * round-1 one-tile-per-CTA kernel); slots in {2,3,4} = probes in flight per lane of that stream
 * (instantiations with run-time tuning only). Negative / other = leave unchanged. */
int cuco_b200_set_stream_variant(int tile_route, int stream_probe, int slots);

/* ---- hash-partitioned multi-GPU support (no reference counterpart) --------------
 * owner(key) = mulhi64(murmur_fmix64(key ^ salt), num_parts): high bits of a mix that is independent
 * of the in-table hash. Both calls are stream-ordered and only enqueue work.
 *
 * cuco_b200_partition_count: counts[p] += #{i : owner(keys[i]) == p}   (counts: int64[num_parts], device)
 * cuco_b200_partition_scatter: given exclusive offsets (device int64[num_parts], consumed as running
 *   cursors), writes keys (and values if non-NULL, and the source index if src_index != NULL) of each
 *   element into its owner's segment. key_bytes/value_bytes in {4,8}. pair_aos != 0: `keys` is an
 *   array of {key,value} structs of 2*key_bytes and the output is the same AoS layout. */
int cuco_b200_partition_count(const void* keys,
                              int key_bytes,
                              int pair_aos,
                              int64_t n,
                              int num_parts,
                              uint64_t salt,
                              int64_t* counts,
                              void* stream);
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
                                void* stream);
/* ---- fused exchange path (native build only; no reference counterpart) --------------------------
 * One routing kernel groups a rank's batch by (owner rank, L2 region of the owner's shard) and
 * stores it straight into the owners' segment buffers through peer pointers (NVLink P2P), so the
 * all-to-all and the region grouping of the owner's probe pass are the same kernel; lookups return
 * the same way (the owner's lookup kernel stores results into the source's result buffer).
 * All shards must have the same capacity. Buffers (per rank, P = num_ranks, R = num_regions,
 * cap = segment_capacity; "peer" = addressable from every rank, e.g. torch symmetric memory):
 *   segments  peer, P*R*cap elements (slot size for mutations, key size for lookups);
 *             owner-side index (region*P + source)*cap + pos
 *   counts    peer, uint32[R*P]   fill of each received segment, written by its source
 *   flags     peer, uint32[P]     flags[s] = elements rank s spilled in this call
 *   results   peer, P*R*cap results; source-side index (owner*R + region)*cap + pos
 *   counts_local uint32[P*R], position_local uint32[n_max] (where each routed key's result will
 *   arrive, source-side index; 0xffffffff = spilled), spill / spill_index [spill_capacity],
 *   spill_count uint32[1]        local scratch of the source
 * Call order per bulk operation (every rank, same stream): barrier; exchange_route; barrier;
 * exchange_mutate or exchange_lookup; (lookups) barrier; exchange_unpermute. A caller that finds a
 * non-zero flag must finish the spilled elements through another path (cucollections_b200/
 * partitioned.py uses the all_to_all fallback). */
int cuco_b200_exchange_plan(cuco_b200_table* t,
                            int64_t n_max,
                            int num_ranks,
                            uint32_t* num_regions,
                            uint32_t* segment_capacity,
                            uint32_t* spill_capacity);
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
                             void* stream);
/* reduce_op < 0: insert; otherwise insert_or_apply with cuco_b200_reduce_op */
int cuco_b200_exchange_mutate(cuco_b200_table* t,
                              const void* segments,
                              const void* counts_recv,
                              uint32_t num_regions,
                              uint32_t segment_capacity,
                              int num_ranks,
                              int reduce_op,
                              void* stream);
/* what: 0 find (results have the payload / key type), 1 contains (bytes) */
int cuco_b200_exchange_lookup(cuco_b200_table* t,
                              const void* segments,
                              const void* counts_recv,
                              void* const* peer_results,
                              uint32_t num_regions,
                              uint32_t segment_capacity,
                              int num_ranks,
                              int my_rank,
                              int what,
                              void* stream);
/* out[i] = results[position_local[i]] for the n keys routed by the matching exchange_route call */
int cuco_b200_exchange_unpermute(cuco_b200_table* t,
                                 const void* results,
                                 const void* position_local,
                                 int64_t n,
                                 void* out,
                                 int what,
                                 void* stream);

/* ---- staged exchange (no reference counterpart) -----------------------------------
 * The transfer itself is left to the copy engines, so that it overlaps the owners' kernels:
 *   1. exchange_stage   groups the local batch by (owner, table slice) into a LOCAL buffer
 *                       stage[owner][slice][segment_capacity] (elements: slot images, or keys for
 *                       lookups) and fills counts_local[owner * slices + slice]; position_local /
 *                       spill* as for exchange_route. `slices` cuts every shard into equal slot ranges.
 *   2. the caller copies block (owner, slice) of `stage` into the owner's receive buffer
 *                       recv[slice][source][segment_capacity] and the matching fill counts into
 *                       recv_counts[slice][source] (cudaMemcpyAsync over NVLink peer mappings; slice by
 *                       slice, so that slice s + 1 travels while slice s is applied).
 *   3. exchange_apply   (mutations) applies the num_ranks received segments of ONE slice - all their
 *                       keys hash into that slice of the table, so the L2-blocked path regroups and
 *                       probes just that slice. reduce_op as for exchange_mutate.
 *      exchange_lookup_local (lookups) answers the received keys [source][segment_capacity] into a LOCAL
 *                       result buffer of the same shape, which the caller copies back to the sources'
 *                       results[owner][segment_capacity]; exchange_unpermute then restores input order
 *                       (lookups are staged with slices = 1: position = owner * segment_capacity + i). */
int cuco_b200_exchange_stage_plan(cuco_b200_table* t,
                                  int64_t n_max,
                                  int num_ranks,
                                  int slices,
                                  uint32_t* segment_capacity,
                                  uint32_t* spill_capacity);
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
                             void* stream);
/* Step 2, small part: stores counts_local[owner * slices + slice] into the owner's
 * recv_counts[slice * num_ranks + my_rank] and *spill_count into every peer's flags[my_rank] (peer
 * stores; the data blocks themselves go through cuco_b200_copy_async). */
int cuco_b200_exchange_publish(const void* counts_local,
                               const void* spill_count,
                               void* const* peer_counts,
                               void* const* peer_flags,
                               int slices,
                               uint32_t segment_capacity,
                               int num_ranks,
                               int my_rank,
                               int source_major, /* != 0: recv_counts[my_rank * slices + slice] (fine mode) */
                               void* stream);
/* Fine mode of the staged exchange, for shards small enough that num_ranks x (L2 regions of the shard)
 * stays within the router's bucket budget: the sources stage by (owner, L2 REGION) directly
 * (exchange_stage with slices = *num_regions), blocks travel source-major into
 * recv[source][region][segment_capacity], and exchange_probe probes a range of regions with their slots
 * resident in L2 - the owner needs no regrouping pass of its own. *num_regions = 0: use exchange_apply. */
int cuco_b200_exchange_fine_regions(cuco_b200_table* t, int num_ranks, uint32_t* num_regions);
int cuco_b200_exchange_probe(cuco_b200_table* t,
                             const void* segments,
                             const void* counts_recv,
                             uint32_t num_regions,
                             uint32_t segment_capacity,
                             int num_ranks,
                             uint32_t region_begin,
                             uint32_t region_count,
                             int reduce_op,
                             void* stream);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): device-to-device, local or through an
 * NVLink peer mapping; runs on the copy engines, next to whatever kernels the SMs are running. */
int cuco_b200_copy_async(void* dst, const void* src, int64_t bytes, void* stream);
/* The same for up to 16 copies at once with ONE small kernel (ctas_per_copy CTAs of 256 threads per copy,
 * 16-byte vectors; every pointer 16-byte aligned, every size a multiple of 16): for the small blocks of
 * the lookup chunks, where eight peer copies cost their fixed latency rather than bandwidth. */
int cuco_b200_push_async(void* const* dst,
                         const void* const* src,
                         const int64_t* bytes,
                         int num_copies,
                         int ctas_per_copy,
                         void* stream);
int cuco_b200_exchange_apply(cuco_b200_table* t,
                             const void* segments,
                             const void* counts_recv,
                             uint32_t segment_capacity,
                             int num_ranks,
                             int slice,
                             int slices,
                             int reduce_op,
                             void* stream);
int cuco_b200_exchange_lookup_local(cuco_b200_table* t,
                                    const void* segments,
                                    const void* counts_recv,
                                    void* results,
                                    uint32_t segment_capacity,
                                    int num_ranks,
                                    int what,
                                    void* stream);

/* out[index[i]] = in[i] for i < n (elem_bytes in {1,4,8}); un-permutes routed lookup results. */
int cuco_b200_scatter_by_index(
  const void* in, const int64_t* index, void* out, int elem_bytes, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUCO_B200_H */
