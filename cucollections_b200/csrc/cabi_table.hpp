// Type-erased table interface behind the C ABI (include/cuco_b200.h). One concrete subclass per
// `cuco_b200_kind`, each compiled in its own translation unit (cabi_kind.cu with
// -DCUCO_SHIM_KIND=<k>) so the instantiations build in parallel. This file has no dependency on
// either cuco include root: the same shim is compiled against ours and against the reference's.
#pragma once

#include <cstdint>
#include <string>

struct cuco_b200_table {
  virtual ~cuco_b200_table() = default;

  virtual int kind() const        = 0;
  virtual int key_bytes() const   = 0;
  virtual int value_bytes() const = 0;  // 0 for sets
  virtual std::int64_t capacity() const = 0;

  virtual std::int64_t size(void* stream) = 0;
  virtual void clear(void* stream)        = 0;

  virtual void insert(const void* keys, const void* values, std::int64_t n, void* stream, std::int64_t* num) = 0;
  virtual void insert_if(const void* keys, const void* values, const std::uint8_t* stencil, std::int64_t n, void* stream, std::int64_t* num) = 0;
  virtual void find(const void* keys, void* out, std::int64_t n, void* stream) = 0;
  virtual void contains(const void* keys, std::uint8_t* out, std::int64_t n, void* stream) = 0;
  virtual void contains_if(const void* keys, const std::uint8_t* stencil, std::uint8_t* out, std::int64_t n, void* stream) = 0;
  virtual void insert_and_find(const void* keys, const void* values, void* found, std::uint8_t* inserted, std::int64_t n, void* stream) = 0;
  virtual void insert_or_assign(const void* keys, const void* values, std::int64_t n, void* stream) = 0;
  virtual void insert_or_apply(const void* keys, const void* values, std::int64_t n, int op, int has_init, std::int64_t init, void* stream) = 0;
  virtual void erase(const void* keys, std::int64_t n, void* stream) = 0;
  virtual std::int64_t retrieve_all(void* keys_out, void* values_out, void* stream) = 0;
  virtual void rehash(std::int64_t capacity, void* stream) = 0;
  virtual std::int64_t count(const void* keys, std::int64_t n, bool outer, void* stream) = 0;
  virtual std::int64_t retrieve(const void* keys, std::int64_t n, bool outer, void* probe_out, void* match_out, void* stream) = 0;

  // ---- exchange path of hash-partitioned tables (native build only; see include/cuco_b200.h) ----
  struct exchange_shape {
    std::uint32_t num_regions, segment_capacity, spill_capacity;
  };
  virtual exchange_shape exchange_plan(std::int64_t n_max, int num_ranks) = 0;
  virtual void exchange_route(const void* keys, const void* values, std::int64_t n, bool keys_only,
                              exchange_shape shape, int num_ranks, int my_rank, std::uint64_t salt,
                              void* const* peer_segments, void* const* peer_counts,
                              void* const* peer_flags, void* counts_local, void* position_local,
                              void* spill, void* spill_index, void* spill_count, void* stream) = 0;
  /// op: -1 insert, 0/1/2 insert_or_apply plus/min/max
  virtual void exchange_mutate(const void* segments, const void* counts_recv, exchange_shape shape,
                               int num_ranks, int op, void* stream) = 0;
  /// what: 0 find (payload / key), 1 contains (bytes)
  virtual void exchange_lookup(const void* segments, const void* counts_recv,
                               void* const* peer_results, exchange_shape shape, int num_ranks,
                               int my_rank, int what, void* stream) = 0;
  virtual void exchange_unpermute(const void* results, const void* position_local, std::int64_t n,
                                  void* out, int what, void* stream) = 0;
  // ---- staged exchange: local grouping, copy-engine transfers, slice-wise apply ----
  virtual exchange_shape stage_plan(std::int64_t n_max, int num_ranks, int groups) = 0;
  virtual void exchange_stage(const void* keys, const void* values, std::int64_t n, bool keys_only,
                              exchange_shape shape, int num_ranks, int my_rank, std::uint64_t salt,
                              void* stage, void* counts_local, void* position_local, void* spill,
                              void* spill_index, void* spill_count, void* stream) = 0;
  virtual void exchange_apply(const void* segments, const void* counts_recv, std::uint32_t segment_capacity,
                              int num_ranks, int group, int groups, int op, void* stream) = 0;
  virtual std::uint32_t exchange_fine_regions(int num_ranks) = 0;
  virtual void exchange_probe(const void* segments, const void* counts_recv, std::uint32_t num_regions,
                              std::uint32_t segment_capacity, int num_ranks, std::uint32_t region_begin,
                              std::uint32_t region_count, int op, void* stream) = 0;
  virtual void exchange_lookup_local(const void* segments, const void* counts_recv, void* results,
                                     std::uint32_t segment_capacity, int num_ranks, int what,
                                     void* stream) = 0;
};

// Factory signature every cabi_kind.cu instance exports (C++ linkage, hidden from the C ABI).
using cuco_shim_factory = cuco_b200_table* (*)(std::int64_t size,
                                               double load_factor,
                                               std::int64_t empty_key,
                                               std::int64_t empty_value,
                                               int has_erased,
                                               std::int64_t erased_key,
                                               void* stream);

#define CUCO_SHIM_DECLARE_FACTORY(K)                                                         \
  cuco_b200_table* cuco_shim_make_kind_##K(std::int64_t, double, std::int64_t, std::int64_t, \
                                           int, std::int64_t, void*)
CUCO_SHIM_DECLARE_FACTORY(0);
CUCO_SHIM_DECLARE_FACTORY(1);
CUCO_SHIM_DECLARE_FACTORY(2);
CUCO_SHIM_DECLARE_FACTORY(3);
CUCO_SHIM_DECLARE_FACTORY(4);
CUCO_SHIM_DECLARE_FACTORY(5);
CUCO_SHIM_DECLARE_FACTORY(6);
CUCO_SHIM_DECLARE_FACTORY(7);
CUCO_SHIM_DECLARE_FACTORY(8);
CUCO_SHIM_DECLARE_FACTORY(9);
CUCO_SHIM_DECLARE_FACTORY(10);
CUCO_SHIM_DECLARE_FACTORY(11);
CUCO_SHIM_DECLARE_FACTORY(12);
CUCO_SHIM_DECLARE_FACTORY(13);
