/*
 * join_oracle.h -- CPU restatement of dwarf_bench's Join hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (dwarf_bench_b200/, the
 * C-ABI library, the dwarf_bench CLI) may include, link or call this.  It is
 * used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED.  The functions below are checked in tests/ against
 *   - the reference's golden vector  tests/join_tests.cpp:10-19  (8 rows),
 *   - the slot-layout known answers  tests/hash_table_tests.cpp:50-54,112-113,175-180,
 *   - the reference's own headers compiled unmodified into oracle/_ref/
 *     (join_helpers.hpp seq_join, hashfunctions.hpp MurmurHash3_x86_32,
 *      hashtable.hpp SimpleNonOwningHashTable) on identical seeded inputs.
 *
 * All file:line citations are relative to the reference checkout.
 */
#ifndef DWARF_JOIN_ORACLE_H
#define DWARF_JOIN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- hash functions (common/dpcpp/hashfunctions.hpp) -------------------- */

/* MurmurHash3_x86_32 of one 4-byte block, BEFORE the `% _sz` step
 * (hashfunctions.hpp:87-129).  len is the byte length the reference passes
 * (always sizeof(uint32_t) == 4 on the Join path, join/join.cpp:32). */
uint32_t dwo_murmur3_x86_32(uint32_t v, uint32_t seed, int len);

/* The full functor: h1 % sz (hashfunctions.hpp:130). */
uint64_t dwo_murmur3_slot(uint32_t v, uint32_t seed, int len, uint64_t sz);

/* ---- join_helpers (join/join_helpers/join_helpers.hpp) ------------------ */

/* seq_join, O(na*nb), emission order i-outer / j-inner (join_helpers.hpp:85-104).
 * Writes at most cap rows and returns the TOTAL number of matches. */
uint64_t dwo_seq_join_u32(const uint32_t *a_keys, const uint32_t *a_vals, uint64_t na,
                          const uint32_t *b_keys, const uint32_t *b_vals, uint64_t nb,
                          uint32_t *out_key, uint32_t *out_va, uint32_t *out_vb,
                          uint64_t cap);
uint64_t dwo_seq_join_u64(const uint64_t *a_keys, const uint64_t *a_vals, uint64_t na,
                          const uint64_t *b_keys, const uint64_t *b_vals, uint64_t nb,
                          uint64_t *out_key, uint64_t *out_va, uint64_t *out_vb,
                          uint64_t cap);

/* Same result multiset as seq_join in O(n log n): sort both sides by key and
 * emit the per-key cross product.  Output rows come out sorted by
 * (key, va, vb) -- the canonical form join_helpers.hpp:106-114 compares in. */
uint64_t dwo_sort_join_u32(const uint32_t *a_keys, const uint32_t *a_vals, uint64_t na,
                           const uint32_t *b_keys, const uint32_t *b_vals, uint64_t nb,
                           uint32_t *out_key, uint32_t *out_va, uint32_t *out_vb,
                           uint64_t cap);
uint64_t dwo_sort_join_u64(const uint64_t *a_keys, const uint64_t *a_vals, uint64_t na,
                           const uint64_t *b_keys, const uint64_t *b_vals, uint64_t nb,
                           uint64_t *out_key, uint64_t *out_va, uint64_t *out_vb,
                           uint64_t cap);

/* operator== on ColJoinedTableTy (join_helpers.hpp:106-125): sort both row
 * lists, compare.  Returns 1 when equal as multisets, 0 otherwise. */
int dwo_rows_equal_u32(const uint32_t *k1, const uint32_t *a1, const uint32_t *b1, uint64_t n1,
                       const uint32_t *k2, const uint32_t *a2, const uint32_t *b2, uint64_t n2);
int dwo_rows_equal_u64(const uint64_t *k1, const uint64_t *a1, const uint64_t *b1, uint64_t n1,
                       const uint64_t *k2, const uint64_t *a2, const uint64_t *b2, uint64_t n2);

/* ---- SimpleNonOwningHashTable (common/dpcpp/hashtable.hpp:5-93) --------- */

enum { DWO_HASH_MURMUR = 0, /* MurmurHash3_x86_32(sz, 4, seed)          */
       DWO_HASH_MODULO = 1  /* StaticSimpleHasher<Size>: v % Size
                               (hashfunctions.hpp:33-35; used by
                               tests/hash_table_tests.cpp)                */ };

typedef struct {
  uint32_t *keys;     /* [size]                                   */
  uint32_t *vals;     /* [size]                                   */
  uint32_t *bitmask;  /* [bitmask_sz]                             */
  uint64_t size;      /* _size        (join.cpp:30: 2*buf_size)   */
  uint64_t bitmask_sz;/* _bitmask_sz  (join.cpp:31: ceil(size/32))*/
  int hash_kind;
  uint32_t seed;
} dwo_table;

/* insert(): claim a slot through update_bitmask, then store key and value
 * (hashtable.hpp:15-21,70-92).  Thread-safe in the same sense as the
 * reference (atomic fetch_or on the bitmask word).  Returns the slot. */
uint32_t dwo_table_insert(dwo_table *t, uint32_t key, uint32_t val);
/* at(): hashtable.hpp:23-40.  Returns 1 and writes *val on a hit. */
int dwo_table_at(const dwo_table *t, uint32_t key, uint32_t *val);
/* has(): hashtable.hpp:42-58. */
int dwo_table_has(const dwo_table *t, uint32_t key);

/* ---- Join::_run timed region (join/join.cpp:30-113) --------------------- */

typedef struct {
  double build_us;   /* join.cpp:112 build_time = build_end - host_start */
  double probe_us;   /* join.cpp:113 probe_time = host_end - build_end   */
  double host_us;    /* join.cpp:111 host_time                           */
  int threads;       /* OpenMP threads used for the two parallel loops   */
} dwo_join_timing;

/* One iteration of the reference's loop body: fresh table (T = 2n slots,
 * bitmask ceil(T/32) words, keys = 0xFFFFFFFF, data = 0), the join_build
 * parallel loop, the join_probe parallel loop writing PROBE-ALIGNED,
 * 0xFFFFFFFF-filled out arrays (join.cpp:36-104).  The SYCL parallel_for is
 * restated as an OpenMP parallel for.  na build rows, nb probe rows (the
 * reference always has na == nb == buf_size).  Returns 0, or -1 on OOM. */
int dwo_join_build_probe_u32(const uint32_t *a_keys, const uint32_t *a_vals, uint64_t na,
                             const uint32_t *b_keys, const uint32_t *b_vals, uint64_t nb,
                             uint32_t murmur_seed,
                             uint32_t *out_key, uint32_t *out_present, uint32_t *out_val,
                             dwo_join_timing *timing);

/* Host compaction of the probe-aligned arrays (join.cpp:119-129): keep row i
 * where out_key[i] != 0xFFFFFFFF.  Returns the number of rows kept. */
uint64_t dwo_compact_u32(const uint32_t *out_key, const uint32_t *out_present,
                         const uint32_t *out_val, uint64_t n,
                         uint32_t *res_k, uint32_t *res_present, uint32_t *res_val);

/* ---- HashBuild::_run (hash/hash_build.cpp:19-81) ------------------------ */

/* Build with val = key, then has() for every source key; returns the number
 * of keys reported present (== n when correct).  build_us receives the timed
 * build (hash_build.cpp:35-57). */
uint64_t dwo_hash_build_check_u32(const uint32_t *src, uint64_t n, uint32_t murmur_seed,
                                  double *build_us, int *threads);

/* ---- JoinOmnisci one-to-many table (common/dpcpp/omnisci_hashtable.hpp) -- */

/* Restates build_table (:80-108) + build_count_buffer (:223-248) +
 * build_pos_buffer (:250-261) + build_id_buffer (:110-147) + lookup (:149-192)
 * with SimpleHasher (v % ht_size, hashfunctions.hpp:43-49), ht_size =
 * 2*distinct (join_omnisci.cpp:69).  For each probe row j writes
 * match_off[j], match_cnt[j] into ids[] (match_cnt 0 when absent).  ids must
 * hold na entries.  Within one key's id list the order is unspecified in the
 * reference (atomic fetch_add race); here it is ascending build row. */
int dwo_omnisci_join_u32(const uint32_t *a_keys, uint64_t na,
                         const uint32_t *b_keys, uint64_t nb,
                         uint64_t *ids, uint64_t *match_off, uint64_t *match_cnt);

/* ---- input generators (common/common.cpp:7-36) -------------------------- */

/* make_unique_random(size): `size` sorted, unique uint32 drawn from
 * [0, 10*size) (common.cpp:7-20).  The reference seeds from
 * std::random_device; here the seed is explicit (xorshift-style generator,
 * same distribution, not the same stream -- the reference's stream is not
 * reproducible by construction). */
void dwo_make_unique_random(uint64_t size, uint64_t seed, uint32_t *out);
/* make_random<uint32_t>(size, 1, 10000) (common.hpp:31-40). */
void dwo_make_random_u32(uint64_t size, uint64_t seed, uint32_t lo, uint32_t hi, uint32_t *out);

int dwo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
