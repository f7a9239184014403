/* cuco_oracle.c — see cuco_oracle.h. TEST INFRASTRUCTURE ONLY (never linked into the product). */
#include "cuco_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===============================================================================================
 * hashes
 * =============================================================================================== */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint32_t rd32(const uint8_t* p)
{
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
static uint64_t rd64(const uint8_t* p)
{
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}

/* reference: detail/hash_functions/xxhash.cuh:113-170 (compute_hash) and :194-202 (finalize) */
uint32_t oracle_xxhash32(const void* data, uint64_t len, uint32_t seed)
{
  const uint32_t P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu,
                 P5 = 0x165667B1u;
  const uint8_t* p  = (const uint8_t*)data;
  uint64_t off      = 0;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = rotl32(v1 + rd32(p + off) * P2, 13) * P1;
      v2 = rotl32(v2 + rd32(p + off + 4) * P2, 13) * P1;
      v3 = rotl32(v3 + rd32(p + off + 8) * P2, 13) * P1;
      v4 = rotl32(v4 + rd32(p + off + 12) * P2, 13) * P1;
      off += 16;
    } while (off + 16 <= len);
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else {
    h = seed + P5;
  }
  h += (uint32_t)len;
  while (off + 4 <= len) {
    h = rotl32(h + rd32(p + off) * P3, 17) * P4;
    off += 4;
  }
  while (off < len) {
    h = rotl32(h + p[off] * P5, 11) * P1;
    off++;
  }
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

static uint64_t xxh64_round(uint64_t acc, uint64_t in)
{
  return rotl64(acc + in * 0xC2B2AE3D27D4EB4Full, 31) * 0x9E3779B185EBCA87ull;
}

/* reference: detail/hash_functions/xxhash.cuh:296-422 */
uint64_t oracle_xxhash64(const void* data, uint64_t len, uint64_t seed)
{
  const uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full,
                 P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull,
                 P5 = 0x27D4EB2F165667C5ull;
  const uint8_t* p  = (const uint8_t*)data;
  uint64_t off      = 0;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = xxh64_round(v1, rd64(p + off));
      v2 = xxh64_round(v2, rd64(p + off + 8));
      v3 = xxh64_round(v3, rd64(p + off + 16));
      v4 = xxh64_round(v4, rd64(p + off + 24));
      off += 32;
    } while (off + 32 <= len);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = (h ^ xxh64_round(0, v1)) * P1 + P4;
    h = (h ^ xxh64_round(0, v2)) * P1 + P4;
    h = (h ^ xxh64_round(0, v3)) * P1 + P4;
    h = (h ^ xxh64_round(0, v4)) * P1 + P4;
  } else {
    h = seed + P5;
  }
  h += len;
  while (off + 8 <= len) {
    h ^= xxh64_round(0, rd64(p + off));
    h = rotl64(h, 27) * P1 + P4;
    off += 8;
  }
  if (off + 4 <= len) {
    h ^= (uint64_t)rd32(p + off) * P1;
    h = rotl64(h, 23) * P2 + P3;
    off += 4;
  }
  while (off < len) {
    h ^= p[off] * P5;
    h = rotl64(h, 11) * P1;
    off++;
  }
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}

static uint32_t fmix32(uint32_t h)
{
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
static uint64_t fmix64(uint64_t h)
{
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}
/* reference: detail/hash_functions/murmurhash3.cuh:57-66, 99-108 */
uint32_t oracle_murmur3_fmix32(uint32_t key, uint32_t seed) { return fmix32(key ^ seed); }
uint64_t oracle_murmur3_fmix64(uint64_t key, uint64_t seed) { return fmix64(key ^ seed); }

/* reference: detail/hash_functions/murmurhash3.cuh:163-205 */
uint32_t oracle_murmur3_32(const void* data, uint64_t len, uint32_t seed)
{
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  const uint8_t* p  = (const uint8_t*)data;
  uint64_t nblocks  = len / 4;
  uint32_t h        = seed;
  for (uint64_t i = 0; i < nblocks; i++) {
    uint32_t k = rd32(p + 4 * i);
    k *= c1;
    k = rotl32(k, 15);
    k *= c2;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  const uint8_t* tail = p + nblocks * 4;
  uint32_t k          = 0;
  switch (len & 3) {
    case 3: k ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1:
      k ^= tail[0];
      k *= c1;
      k = rotl32(k, 15);
      k *= c2;
      h ^= k;
  }
  h ^= (uint32_t)len;
  return fmix32(h);
}

/* reference: detail/hash_functions/murmurhash3.cuh:288-373 */
void oracle_murmur3_x64_128(const void* data, uint64_t len, uint64_t seed, uint64_t out[2])
{
  const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
  const uint8_t* p  = (const uint8_t*)data;
  uint64_t nblocks  = len / 16;
  uint64_t h1 = seed, h2 = seed;
  for (uint64_t i = 0; i < nblocks; i++) {
    uint64_t k1 = rd64(p + 16 * i), k2 = rd64(p + 16 * i + 8);
    k1 *= c1;
    k1 = rotl64(k1, 31);
    k1 *= c2;
    h1 ^= k1;
    h1 = rotl64(h1, 27);
    h1 += h2;
    h1 = h1 * 5 + 0x52dce729;
    k2 *= c2;
    k2 = rotl64(k2, 33);
    k2 *= c1;
    h2 ^= k2;
    h2 = rotl64(h2, 31);
    h2 += h1;
    h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t* tail = p + nblocks * 16;
  uint64_t k1 = 0, k2 = 0;
  int rem = (int)(len & 15);
  for (int i = rem - 1; i >= 8; i--) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
  if (rem > 8) {
    k2 *= c2;
    k2 = rotl64(k2, 33);
    k2 *= c1;
    h2 ^= k2;
  }
  for (int i = (rem < 8 ? rem : 8) - 1; i >= 0; i--) k1 ^= (uint64_t)tail[i] << (8 * i);
  if (rem > 0) {
    k1 *= c1;
    k1 = rotl64(k1, 31);
    k1 *= c2;
    h1 ^= k1;
  }
  h1 ^= len;
  h2 ^= len;
  h1 += h2;
  h2 += h1;
  h1 = fmix64(h1);
  h2 = fmix64(h2);
  h1 += h2;
  h2 += h1;
  out[0] = h1;
  out[1] = h2;
}

/* reference: detail/hash_functions/murmurhash3.cuh:451-590 */
void oracle_murmur3_x86_128(const void* data, uint64_t len, uint32_t seed, uint32_t out[4])
{
  const uint32_t c[4]    = {0x239b961bu, 0xab0e9789u, 0x38b34ae5u, 0xa1e38b93u};
  const int krot[4]      = {15, 16, 17, 18};
  const int hrot[4]      = {19, 17, 15, 13};
  const uint32_t hadd[4] = {0x561ccd1bu, 0x0bcaa747u, 0x96cd1c35u, 0x32ac3b17u};
  const uint8_t* p       = (const uint8_t*)data;
  uint64_t nblocks       = len / 16;
  uint32_t h[4]          = {seed, seed, seed, seed};
  for (uint64_t i = 0; i < nblocks; i++) {
    for (int j = 0; j < 4; j++) {
      uint32_t k = rd32(p + 16 * i + 4 * j);
      k *= c[j];
      k = rotl32(k, krot[j]);
      k *= c[(j + 1) & 3];
      h[j] ^= k;
      h[j] = rotl32(h[j], hrot[j]);
      h[j] += h[(j + 1) & 3];
      h[j] = h[j] * 5 + hadd[j];
    }
  }
  const uint8_t* tail = p + nblocks * 16;
  int rem             = (int)(len & 15);
  for (int j = 3; j >= 0; j--) {
    if (rem > 4 * j) {
      uint32_t k = 0;
      int hi     = rem < 4 * j + 4 ? rem : 4 * j + 4;
      for (int t = hi - 1; t >= 4 * j; t--) k ^= (uint32_t)tail[t] << (8 * (t - 4 * j));
      k *= c[j];
      k = rotl32(k, krot[j]);
      k *= c[(j + 1) & 3];
      h[j] ^= k;
    }
  }
  for (int j = 0; j < 4; j++) h[j] ^= (uint32_t)len;
  h[0] += h[1]; h[0] += h[2]; h[0] += h[3];
  h[1] += h[0]; h[2] += h[0]; h[3] += h[0];
  for (int j = 0; j < 4; j++) h[j] = fmix32(h[j]);
  h[0] += h[1]; h[0] += h[2]; h[0] += h[3];
  h[1] += h[0]; h[2] += h[0]; h[3] += h[0];
  memcpy(out, h, sizeof(h));
}

/* ===============================================================================================
 * capacity rounding (reference: detail/extent/extent.inl:90-116; prime.hpp:30 regenerated by rule)
 * =============================================================================================== */
static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m)
{
  return (uint64_t)((unsigned __int128)a * b % m);
}
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t m)
{
  uint64_t r = 1;
  a %= m;
  while (e) {
    if (e & 1) r = mulmod(r, a, m);
    a = mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static int is_prime_u64(uint64_t n)
{
  static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (int i = 0; i < 12; i++) {
    if (n % bases[i] == 0) return n == bases[i];
  }
  uint64_t d = n - 1;
  int s      = 0;
  while ((d & 1) == 0) {
    d >>= 1;
    s++;
  }
  for (int i = 0; i < 12; i++) {
    uint64_t x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int composite = 1;
    for (int r = 1; r < s; r++) {
      x = mulmod(x, x, n);
      if (x == n - 1) {
        composite = 0;
        break;
      }
    }
    if (composite) return 0;
  }
  return 1;
}

#define ORACLE_LAST_PRIME 17177758133ull

uint64_t oracle_prime_at_least(uint64_t n)
{
  /* the sequence is path dependent, so walk it from the start; memoise the position reached */
  static uint64_t cached = 7;
  if (n > ORACLE_LAST_PRIME) return 0;
  if (n <= 2) return 2;
  if (n <= 3) return 3;
  if (n <= 5) return 5;
  if (n <= 7) return 7;
  uint64_t p = (cached <= n) ? cached : 7;
  /* `cached` only ever holds sequence members, and the walk from a smaller member reaches n */
  while (p < n) {
    uint64_t c = p + (p < (1ull << 17) ? 6 : (1ull << 17));
    while (!is_prime_u64(c)) c++;
    if (c <= n) cached = c;
    p = c;
  }
  return p;
}

uint64_t oracle_num_windows(int64_t requested, int cg, int w)
{
  uint64_t req   = requested < 1 ? 1 : (uint64_t)requested;
  uint64_t group = (uint64_t)cg * (uint64_t)w;
  uint64_t size  = (req + group - 1) / group;
  uint64_t prime = oracle_prime_at_least(size);
  return prime ? prime * (uint64_t)cg : 0;
}

uint64_t oracle_ceil_div_lf(uint64_t n, double load_factor)
{
  return (uint64_t)ceil((double)n / load_factor);
}

/* ===============================================================================================
 * table
 * =============================================================================================== */
struct oracle_table {
  int key_bytes, value_bytes, cg, w, probing, hash;
  uint64_t num_windows; /* N */
  uint64_t capacity;    /* N * w */
  int64_t empty_key, empty_value, erased_key;
  int64_t* keys;
  int64_t* values; /* NULL for sets */
  int allows_duplicates; /* static_multiset / static_multimap semantics (ref_impl.cuh:98-99) */
};

enum { UNEQUAL = 0, EQUAL = 1, EMPTY = 2, AVAILABLE = 3 }; /* equal_wrapper.cuh:28-35 */

static int64_t narrow(int64_t v, int bytes) { return bytes == 4 ? (int64_t)(int32_t)v : v; }

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
                            int64_t erased_key)
{
  if ((key_bytes != 4 && key_bytes != 8) || (value_bytes != 0 && value_bytes != 4 && value_bytes != 8)) return NULL;
  if (cg < 1 || w < 1) return NULL;
  int64_t requested = size;
  if (load_factor != 0.0) {
    /* impl.cuh:181-186 */
    if (!(load_factor > 0.0) || !(load_factor <= 1.0)) return NULL;
    requested = (int64_t)oracle_ceil_div_lf((uint64_t)(size < 0 ? 0 : size), load_factor);
  }
  empty_key  = narrow(empty_key, key_bytes);
  erased_key = has_erased ? narrow(erased_key, key_bytes) : empty_key;
  if (has_erased && erased_key == empty_key) return NULL; /* impl.cuh:225-227 */
  uint64_t n = oracle_num_windows(requested, cg, w);
  if (n == 0) return NULL;
  oracle_table* t = (oracle_table*)calloc(1, sizeof(*t));
  t->key_bytes    = key_bytes;
  t->value_bytes  = value_bytes;
  t->cg           = cg;
  t->w            = w;
  t->probing      = probing;
  t->hash         = hash;
  t->num_windows  = n;
  t->capacity     = n * (uint64_t)w;
  t->empty_key    = empty_key;
  t->empty_value  = value_bytes ? narrow(empty_value, value_bytes) : 0;
  t->erased_key   = erased_key;
  t->keys         = (int64_t*)malloc(sizeof(int64_t) * t->capacity);
  t->values       = value_bytes ? (int64_t*)malloc(sizeof(int64_t) * t->capacity) : NULL;
  oracle_clear(t);
  return t;
}

void oracle_destroy(oracle_table* t)
{
  if (!t) return;
  free(t->keys);
  free(t->values);
  free(t);
}

int64_t oracle_capacity(const oracle_table* t) { return (int64_t)t->capacity; }

void oracle_clear(oracle_table* t)
{
  for (uint64_t i = 0; i < t->capacity; i++) {
    t->keys[i] = t->empty_key;
    if (t->values) t->values[i] = t->empty_value;
  }
}

/* kernels.cuh:642-667 + functors.cuh:66-107 */
int64_t oracle_size(const oracle_table* t)
{
  int64_t n = 0;
  for (uint64_t i = 0; i < t->capacity; i++) {
    n += !(t->keys[i] == t->empty_key || t->keys[i] == t->erased_key);
  }
  return n;
}

/* hash of a key over its object representation; second hash of double hashing is seeded with 1
 * (probing_scheme.cuh:139) */
static uint64_t hash_key(const oracle_table* t, int64_t key, uint64_t seed)
{
  uint8_t buf[8];
  if (t->key_bytes == 4) {
    int32_t k = (int32_t)key;
    memcpy(buf, &k, 4);
  } else {
    memcpy(buf, &key, 8);
  }
  switch (t->hash) {
    case ORACLE_XXHASH64: return oracle_xxhash64(buf, (uint64_t)t->key_bytes, seed);
    case ORACLE_MURMUR3_32: return oracle_murmur3_32(buf, (uint64_t)t->key_bytes, (uint32_t)seed);
    default: return oracle_xxhash32(buf, (uint64_t)t->key_bytes, (uint32_t)seed);
  }
}

/* detail/utils.cuh:133-142: hash + rank without overflowing size_t */
static uint64_t add_rank(uint64_t base, uint64_t rank)
{
  uint64_t top = UINT64_MAX;
  return (base > top - rank) ? rank - (top - base) : base + rank;
}

typedef struct {
  uint64_t start; /* window index of rank 0 ... see lane_window */
  uint64_t step;
  uint64_t h1;
} probe_t;

/* probing_scheme_impl.inl:106-128 (linear), :167-193 (double) */
static probe_t make_probe(const oracle_table* t, int64_t key)
{
  probe_t p;
  uint64_t n = t->num_windows;
  p.h1       = hash_key(t, key, 0);
  p.start    = p.h1 % n;
  if (t->probing == ORACLE_LINEAR) {
    p.step = (uint64_t)t->cg; /* scalar form uses 1 == cg when cg == 1 */
  } else {
    uint64_t h2 = hash_key(t, key, 1);
    p.step      = (h2 % (n / (uint64_t)t->cg - 1) + 1) * (uint64_t)t->cg;
  }
  return p;
}

/* window visited by lane `rank` at step `k` */
static uint64_t lane_window(const oracle_table* t, const probe_t* p, int rank, uint64_t k)
{
  uint64_t n     = t->num_windows;
  uint64_t start = add_rank(p->h1, (uint64_t)rank) % n;
  return (uint64_t)(((unsigned __int128)start + (unsigned __int128)k * p->step) % n);
}

void oracle_probe_sequence(const oracle_table* t, int64_t key, int rank, int64_t* out, int len)
{
  probe_t p = make_probe(t, narrow(key, t->key_bytes));
  for (int i = 0; i < len; i++) out[i] = (int64_t)lane_window(t, &p, rank, (uint64_t)i);
}

/* equal_wrapper.cuh:97-108 */
static int classify(const oracle_table* t, int64_t probe, int64_t slot_key, int for_insert)
{
  if (for_insert) {
    if (slot_key == t->empty_key || slot_key == t->erased_key) return AVAILABLE;
    /* multi-containers never look for an equal key when inserting (ref_impl.cuh:389-391, 444-446) */
    if (t->allows_duplicates) return UNEQUAL;
  } else {
    if (slot_key == t->empty_key) return EMPTY;
  }
  return probe == slot_key ? EQUAL : UNEQUAL;
}

/* One probe step as a cg-wide group sees it: every lane reports the first slot of its window whose
 * state is not UNEQUAL (ref_impl.cuh:433-452, 952-961). Returns the group decision:
 *   *slot = index of the slot of the lowest lane reporting `EQUAL` (if any lane does) else of the
 *   lowest lane reporting AVAILABLE/EMPTY; return value = that state, or UNEQUAL if no lane has one. */
static int group_step(const oracle_table* t, const probe_t* p, uint64_t k, int64_t key, int for_insert, uint64_t* slot)
{
  int best_state     = UNEQUAL;
  uint64_t best_slot = 0;
  for (int r = 0; r < t->cg; r++) {
    uint64_t wdx = lane_window(t, p, r, k);
    for (int i = 0; i < t->w; i++) {
      uint64_t s = wdx * (uint64_t)t->w + (uint64_t)i;
      int st     = classify(t, key, t->keys[s], for_insert);
      if (st == UNEQUAL) continue;
      /* this lane's report; EQUAL from any lane outranks AVAILABLE/EMPTY from lower lanes */
      if (st == EQUAL) {
        if (best_state != EQUAL) {
          best_state = EQUAL;
          best_slot  = s;
        }
      } else if (best_state == UNEQUAL) {
        best_state = st;
        best_slot  = s;
      }
      break;
    }
  }
  *slot = best_slot;
  return best_state;
}

/* ref_impl.cuh:374-485 (insert), :502-655 (insert_and_find): returns slot, *is_new */
static uint64_t insert_one(oracle_table* t, int64_t key, int64_t value, int* is_new)
{
  probe_t p = make_probe(t, key);
  for (uint64_t k = 0;; k++) {
    uint64_t s;
    int st = group_step(t, &p, k, key, 1, &s);
    if (st == EQUAL) {
      *is_new = 0;
      return s;
    }
    if (st == AVAILABLE) {
      t->keys[s] = key;
      if (t->values) t->values[s] = value;
      *is_new = 1;
      return s;
    }
  }
}

/* ref_impl.cuh:766-822 (contains), :907-979 (find): returns 1 and *slot when present */
static int find_one(const oracle_table* t, int64_t key, uint64_t* slot)
{
  probe_t p = make_probe(t, key);
  for (uint64_t k = 0;; k++) {
    int st = group_step(t, &p, k, key, 0, slot);
    if (st == EQUAL) return 1;
    if (st == EMPTY) return 0;
  }
}

int64_t oracle_insert_if(
  oracle_table* t, const int64_t* keys, const int64_t* values, const uint8_t* stencil, int64_t n)
{
  int64_t added = 0;
  for (int64_t i = 0; i < n; i++) {
    if (stencil && !stencil[i]) continue; /* kernels.cuh:78-80 */
    int is_new;
    insert_one(t, narrow(keys[i], t->key_bytes), values ? narrow(values[i], t->value_bytes) : 0, &is_new);
    added += is_new;
  }
  return added;
}

int64_t oracle_insert(oracle_table* t, const int64_t* keys, const int64_t* values, int64_t n)
{
  return oracle_insert_if(t, keys, values, NULL, n);
}

/* kernels.cuh:341-400: payload (map) / stored key (set), sentinel on miss */
void oracle_find(const oracle_table* t, const int64_t* keys, int64_t* out, int64_t n)
{
  for (int64_t i = 0; i < n; i++) {
    uint64_t s;
    if (find_one(t, narrow(keys[i], t->key_bytes), &s)) {
      out[i] = t->values ? t->values[s] : t->keys[s];
    } else {
      out[i] = t->values ? t->empty_value : t->empty_key;
    }
  }
}

/* kernels.cuh:257-297 */
void oracle_contains_if(
  const oracle_table* t, const int64_t* keys, const uint8_t* stencil, uint8_t* out, int64_t n)
{
  for (int64_t i = 0; i < n; i++) {
    uint64_t s;
    out[i] = (!stencil || stencil[i]) ? (uint8_t)find_one(t, narrow(keys[i], t->key_bytes), &s) : 0;
  }
}

void oracle_contains(const oracle_table* t, const int64_t* keys, uint8_t* out, int64_t n)
{
  oracle_contains_if(t, keys, NULL, out, n);
}

/* kernels.cuh:504-564 */
void oracle_insert_and_find(oracle_table* t,
                            const int64_t* keys,
                            const int64_t* values,
                            int64_t* found,
                            uint8_t* inserted,
                            int64_t n)
{
  for (int64_t i = 0; i < n; i++) {
    int is_new;
    uint64_t s  = insert_one(t, narrow(keys[i], t->key_bytes), values ? narrow(values[i], t->value_bytes) : 0, &is_new);
    found[i]    = t->values ? t->values[s] : t->keys[s];
    inserted[i] = (uint8_t)is_new;
  }
}

/* static_map_ref.inl:486-620: key present afterwards, payload = this element's value */
void oracle_insert_or_assign(oracle_table* t, const int64_t* keys, const int64_t* values, int64_t n)
{
  for (int64_t i = 0; i < n; i++) {
    int is_new;
    int64_t v  = narrow(values[i], t->value_bytes);
    uint64_t s = insert_one(t, narrow(keys[i], t->key_bytes), v, &is_new);
    if (!is_new) t->values[s] = v;
  }
}

static int64_t apply_op(int op, int64_t a, int64_t b, int bytes)
{
  switch (op) {
    case ORACLE_MIN: return a < b ? a : b;
    case ORACLE_MAX: return a > b ? a : b;
    default: return narrow((int64_t)((uint64_t)a + (uint64_t)b), bytes);
  }
}

/* static_map_ref.inl:788-829 (dispatch on init), :850-1052 (impl + attempt_insert_or_apply):
 *   slot <= 8 bytes: first arrival stores its value (packed CAS of the whole pair)
 *   slot 16 bytes, init == empty_value: first arrival applies op onto the sentinel payload
 *   slot 16 bytes otherwise: first arrival stores its value
 *   later arrivals always apply op */
void oracle_insert_or_apply(oracle_table* t,
                            const int64_t* keys,
                            const int64_t* values,
                            int64_t n,
                            int reduce_op,
                            int has_init,
                            int64_t init)
{
  int slot_bytes = t->key_bytes + t->value_bytes;
  int direct     = has_init && narrow(init, t->value_bytes) == t->empty_value && slot_bytes > 8;
  for (int64_t i = 0; i < n; i++) {
    int is_new;
    int64_t key = narrow(keys[i], t->key_bytes);
    int64_t v   = narrow(values[i], t->value_bytes);
    /* probe first so a new entry can start from the sentinel payload in direct mode */
    probe_t p = make_probe(t, key);
    uint64_t s;
    int st;
    for (uint64_t k = 0;; k++) {
      st = group_step(t, &p, k, key, 1, &s);
      if (st == EQUAL || st == AVAILABLE) break;
    }
    is_new = st == AVAILABLE;
    if (is_new) {
      t->keys[s]   = key;
      t->values[s] = direct ? apply_op(reduce_op, t->empty_value, v, t->value_bytes) : v;
    } else {
      t->values[s] = apply_op(reduce_op, t->values[s], v, t->value_bytes);
    }
  }
}

/* ref_impl.cuh:667-751: tombstone = {erased_key, empty_value} */
void oracle_erase(oracle_table* t, const int64_t* keys, int64_t n)
{
  for (int64_t i = 0; i < n; i++) {
    uint64_t s;
    if (find_one(t, narrow(keys[i], t->key_bytes), &s)) {
      t->keys[s] = t->erased_key;
      if (t->values) t->values[s] = t->empty_value;
    }
  }
}

/* impl.cuh:726-785 (order unspecified there; slot order here) */
void oracle_set_allows_duplicates(oracle_table* t, int allows) { t->allows_duplicates = allows != 0; }

/* ref_impl.cuh:834-892 (count) and :1046-1282 (retrieve): per probe step every lane walks its window
 * up to the first EMPTY slot, every EQUAL slot before it is a match; the walk ends after the step
 * in which any lane saw EMPTY. Containers without duplicates stop at the first match (:836-837).
 * `emit` != NULL receives the slot index of every match. */
static int64_t matches_of(const oracle_table* t, int64_t key, uint64_t* emit, int64_t emit_cap)
{
  probe_t p     = make_probe(t, key);
  int64_t count = 0;
  for (uint64_t k = 0;; k++) {
    int saw_empty = 0;
    for (int r = 0; r < t->cg; r++) {
      uint64_t wdx = lane_window(t, &p, r, k);
      for (int i = 0; i < t->w; i++) {
        uint64_t s = wdx * (uint64_t)t->w + (uint64_t)i;
        int st     = classify(t, key, t->keys[s], 0);
        if (st == EMPTY) {
          saw_empty = 1;
          break;
        }
        if (st == EQUAL) {
          if (emit && count < emit_cap) emit[count] = s;
          count++;
          if (!t->allows_duplicates) return count;
        }
      }
    }
    if (saw_empty) return count;
  }
}

/* kernels.cuh:587-627 + impl.cuh:677-706: sum of matches; outer: a key without matches counts 1 */
int64_t oracle_count(const oracle_table* t, const int64_t* keys, int64_t n, int outer)
{
  int64_t total = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t c = matches_of(t, narrow(keys[i], t->key_bytes), NULL, 0);
    total += (outer && c == 0) ? 1 : c;
  }
  return total;
}

/* kernels.cuh:437-471, impl.cuh:604-660, static_set.inl:349-373: rows {probe key, matched slot
 * content}; outer: {key, empty sentinel} for keys without matches. The reference leaves the row
 * order unspecified; here rows follow input order, matches in probe order. Returns the row count;
 * match_values_out may be NULL (sets / multisets). */
int64_t oracle_retrieve(const oracle_table* t,
                        const int64_t* keys,
                        int64_t n,
                        int outer,
                        int64_t* probe_out,
                        int64_t* match_keys_out,
                        int64_t* match_values_out)
{
  int64_t rows = 0;
  uint64_t* hits = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(t->capacity ? t->capacity : 1));
  for (int64_t i = 0; i < n; i++) {
    int64_t key = narrow(keys[i], t->key_bytes);
    int64_t c   = matches_of(t, key, hits, (int64_t)t->capacity);
    if (c == 0 && outer) {
      probe_out[rows]      = key;
      match_keys_out[rows] = t->empty_key;
      if (match_values_out) match_values_out[rows] = t->values ? t->empty_value : 0;
      rows++;
    }
    for (int64_t j = 0; j < c; j++) {
      probe_out[rows]      = key;
      match_keys_out[rows] = t->keys[hits[j]];
      if (match_values_out) match_values_out[rows] = t->values ? t->values[hits[j]] : 0;
      rows++;
    }
  }
  free(hits);
  return rows;
}

int64_t oracle_retrieve_all(const oracle_table* t, int64_t* keys_out, int64_t* values_out)
{
  int64_t n = 0;
  for (uint64_t i = 0; i < t->capacity; i++) {
    if (t->keys[i] == t->empty_key || t->keys[i] == t->erased_key) continue;
    keys_out[n] = t->keys[i];
    if (t->values && values_out) values_out[n] = t->values[i];
    n++;
  }
  return n;
}
