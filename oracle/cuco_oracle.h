/* cuco_oracle.h — CPU restatement of cuCollections' open-addressing hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only
 * as the checker or the reported baseline. The product (cucollections_b200/libcuco_b200.so) never
 * links, loads or falls back to it.
 *
 * What it restates, sequentially and in plain C, with the reference file:line it follows:
 *   hashes            include/cuco/detail/hash_functions/xxhash.cuh:65-205,248-422; murmurhash3.cuh:38-616
 *   capacity rounding include/cuco/detail/extent/extent.inl:90-116 + detail/prime.hpp:30 (rule form)
 *   probe sequences   include/cuco/detail/probing_scheme/probing_scheme_impl.inl:63-69,106-128,167-193
 *   slot comparison   include/cuco/detail/equal_wrapper.cuh:29-109
 *   insert/find/...   include/cuco/detail/open_addressing/open_addressing_ref_impl.cuh:374-979
 *   upserts           include/cuco/detail/static_map/static_map_ref.inl:486-620,850-1052
 *   bulk semantics    include/cuco/detail/open_addressing/kernels.cuh:64-667
 *   count / retrieve  include/cuco/detail/open_addressing/open_addressing_ref_impl.cuh:834-892,1009-1282
 * A cooperative group of cg lanes is executed as "look at all cg windows of the step, then act",
 * which is what the ballots in the reference compute. Elements are processed in input order, so
 * where the reference says "one unspecified element wins" the oracle's answer is the first; parity
 * tests therefore use inputs whose result does not depend on the winner (value = f(key), or a
 * commutative reduction).
 *
 * Parity pinning: tests/test_oracle_golden.py replays every known-answer vector the reference's own
 * tests hold for this path (hash values, capacities, extents) against this library, and
 * tests/test_parity_gpu.py compares it call-for-call with cuco itself (oracle/_ref/libcuco_ref.so)
 * on the GPU.
 */
#ifndef CUCO_ORACLE_H
#define CUCO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- hashes (byte-stream form; keys are hashed over their 4 or 8 byte object representation) --- */
uint32_t oracle_xxhash32(const void* data, uint64_t len, uint32_t seed);
uint64_t oracle_xxhash64(const void* data, uint64_t len, uint64_t seed);
uint32_t oracle_murmur3_32(const void* data, uint64_t len, uint32_t seed);
void oracle_murmur3_x64_128(const void* data, uint64_t len, uint64_t seed, uint64_t out[2]);
void oracle_murmur3_x86_128(const void* data, uint64_t len, uint32_t seed, uint32_t out[4]);
uint32_t oracle_murmur3_fmix32(uint32_t key, uint32_t seed);
uint64_t oracle_murmur3_fmix64(uint64_t key, uint64_t seed);

/* ---- capacity rounding ---------------------------------------------------------------------- */
/* Smallest member >= n of the reference's prime sequence (2,3,5,7, then next prime >= last+6 below
 * 2^17, then next prime >= last+2^17); 0 if n exceeds the last entry 17177758133. */
uint64_t oracle_prime_at_least(uint64_t n);
/* make_window_extent<cg, w>(requested).value(): number of windows; 0 on "Invalid input extent". */
uint64_t oracle_num_windows(int64_t requested, int cg, int w);
/* fast_int identities are checked against plain C division by the tests; exposed for completeness */
uint64_t oracle_ceil_div_lf(uint64_t n, double load_factor); /* ceil(n / lf) as in impl.cuh:181-183 */

/* ---- tables --------------------------------------------------------------------------------- */
enum oracle_hash { ORACLE_XXHASH32 = 0, ORACLE_XXHASH64 = 1, ORACLE_MURMUR3_32 = 2 };
enum oracle_probing { ORACLE_LINEAR = 0, ORACLE_DOUBLE = 1 };
enum oracle_reduce { ORACLE_PLUS = 0, ORACLE_MIN = 1, ORACLE_MAX = 2 };

typedef struct oracle_table oracle_table;

/* value_bytes == 0 makes a set. load_factor == 0: `size` is a capacity, else a key count.
 * Returns NULL on invalid arguments (load factor outside (0,1], erased == empty, extent too large). */
oracle_table* oracle_create(int key_bytes,
                            int value_bytes,
                            int cg,
                            int w,
                            int probing,
                            int hash,
                            int64_t size,
                            double load_factor,
                            int64_t empty_key,
                            int64_t empty_value,
                            int has_erased,
                            int64_t erased_key);
void oracle_destroy(oracle_table* t);
int64_t oracle_capacity(const oracle_table* t);
int64_t oracle_size(const oracle_table* t);
void oracle_clear(oracle_table* t);

/* All key/value/output arrays are int64 on this side (narrow keys are sign-extended by the caller);
 * stencil/bool arrays are bytes. `values` may be NULL for sets. Return value of insert*: #new keys. */
int64_t oracle_insert(oracle_table* t, const int64_t* keys, const int64_t* values, int64_t n);
int64_t oracle_insert_if(
  oracle_table* t, const int64_t* keys, const int64_t* values, const uint8_t* stencil, int64_t n);
void oracle_find(const oracle_table* t, const int64_t* keys, int64_t* out, int64_t n);
void oracle_contains(const oracle_table* t, const int64_t* keys, uint8_t* out, int64_t n);
void oracle_contains_if(
  const oracle_table* t, const int64_t* keys, const uint8_t* stencil, uint8_t* out, int64_t n);
void oracle_insert_and_find(oracle_table* t,
                            const int64_t* keys,
                            const int64_t* values,
                            int64_t* found,
                            uint8_t* inserted,
                            int64_t n);
void oracle_insert_or_assign(oracle_table* t, const int64_t* keys, const int64_t* values, int64_t n);
void oracle_insert_or_apply(oracle_table* t,
                            const int64_t* keys,
                            const int64_t* values,
                            int64_t n,
                            int reduce_op,
                            int has_init,
                            int64_t init);
void oracle_erase(oracle_table* t, const int64_t* keys, int64_t n);
int64_t oracle_retrieve_all(const oracle_table* t, int64_t* keys_out, int64_t* values_out);
/* static_multiset semantics (include/cuco/static_multiset.cuh:81-729): inserts never compare keys,
 * so equal keys are stored repeatedly. Set right after oracle_create. */
void oracle_set_allows_duplicates(oracle_table* t, int allows);
/* count / count_outer (open_addressing_impl.cuh:677-706): total matches of the n probe keys. */
int64_t oracle_count(const oracle_table* t, const int64_t* keys, int64_t n, int outer);
/* retrieve / retrieve_outer (open_addressing_impl.cuh:604-660; static_set.inl:349-373): one row
 * {probe key, matched key[, matched payload]} per match, in input order (the reference's order is
 * unspecified: compare sorted). Returns the number of rows; outputs sized by oracle_count. */
int64_t oracle_retrieve(const oracle_table* t,
                        const int64_t* keys,
                        int64_t n,
                        int outer,
                        int64_t* probe_out,
                        int64_t* match_keys_out,
                        int64_t* match_values_out);
/* First `len` window indices of the probe sequence of `key` as rank `rank` of a cg-wide group
 * (rank 0 of cg 1 = scalar sequence); for tests/utility/probing_scheme_test.cu style checks. */
void oracle_probe_sequence(const oracle_table* t, int64_t key, int rank, int64_t* out, int len);

#ifdef __cplusplus
}
#endif
#endif
