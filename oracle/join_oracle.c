/*
 * join_oracle.c -- CPU restatement of dwarf_bench's Join hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see join_oracle.h.  Parity status: PINNED
 * (golden vectors + the reference's own headers compiled into oracle/_ref).
 *
 * Build: `make -C oracle` -> oracle/_build/libjoin_oracle.so
 */
#include "join_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* MurmurHash3_x86_32 -- hashfunctions.hpp:64-137                            */
/* ------------------------------------------------------------------------ */

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline uint32_t fmix32(uint32_t h) {   /* hashfunctions.hpp:77-85 */
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

uint32_t dwo_murmur3_x86_32(uint32_t v, uint32_t seed, int len) {
  /* hashfunctions.hpp:87-129.  The functor hashes the bytes of ONE uint32_t;
   * nblocks = len/4, the tail switch covers len&3 bytes of the same word. */
  const uint8_t *data = (const uint8_t *)&v;
  const int nblocks = len / 4;
  uint32_t h1 = seed;
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  for (int i = 0; i < nblocks; ++i) {          /* body, :100-110 */
    uint32_t k1;
    memcpy(&k1, data + 4 * i, 4);
    k1 *= c1;
    k1 = rotl32(k1, 15);
    k1 *= c2;
    h1 ^= k1;
    h1 = rotl32(h1, 13);
    h1 = h1 * 5 + 0xe6546b64u;
  }
  const uint8_t *tail = data + nblocks * 4;    /* tail, :112-126 */
  uint32_t k1 = 0;
  switch (len & 3) {
  case 3: k1 ^= (uint32_t)tail[2] << 16; /* fallthrough */
  case 2: k1 ^= (uint32_t)tail[1] << 8;  /* fallthrough */
  case 1:
    k1 ^= tail[0];
    k1 *= c1;
    k1 = rotl32(k1, 15);
    k1 *= c2;
    h1 ^= k1;
  }
  h1 ^= (uint32_t)len;                          /* :128 */
  return fmix32(h1);                            /* :129 */
}

uint64_t dwo_murmur3_slot(uint32_t v, uint32_t seed, int len, uint64_t sz) {
  return (uint64_t)dwo_murmur3_x86_32(v, seed, len) % sz;  /* :130 */
}

/* ------------------------------------------------------------------------ */
/* seq_join / sort join / row equality -- join_helpers.hpp:85-125            */
/* ------------------------------------------------------------------------ */

#define DEFINE_JOIN_FUNCS(SUF, T)                                                        \
  typedef struct { T k, a, b; } row_##SUF;                                               \
  typedef struct { T k, v; } kv_##SUF;                                                   \
                                                                                         \
  static int cmp_row_##SUF(const void *x, const void *y) {                               \
    const row_##SUF *p = (const row_##SUF *)x, *q = (const row_##SUF *)y;                \
    if (p->k != q->k) return p->k < q->k ? -1 : 1;                                       \
    if (p->a != q->a) return p->a < q->a ? -1 : 1;                                       \
    if (p->b != q->b) return p->b < q->b ? -1 : 1;                                       \
    return 0;                                                                            \
  }                                                                                      \
  static int cmp_kv_##SUF(const void *x, const void *y) {                                \
    const kv_##SUF *p = (const kv_##SUF *)x, *q = (const kv_##SUF *)y;                   \
    if (p->k != q->k) return p->k < q->k ? -1 : 1;                                       \
    if (p->v != q->v) return p->v < q->v ? -1 : 1;                                       \
    return 0;                                                                            \
  }                                                                                      \
                                                                                         \
  uint64_t dwo_seq_join_##SUF(const T *a_keys, const T *a_vals, uint64_t na,             \
                              const T *b_keys, const T *b_vals, uint64_t nb,             \
                              T *out_key, T *out_va, T *out_vb, uint64_t cap) {          \
    uint64_t m = 0;                                                                      \
    for (uint64_t i = 0; i < na; ++i)          /* join_helpers.hpp:93 */                 \
      for (uint64_t j = 0; j < nb; ++j)        /* :94 */                                 \
        if (a_keys[i] == b_keys[j]) {          /* :95 */                                 \
          if (m < cap) {                                                                 \
            out_key[m] = a_keys[i];            /* :96 */                                 \
            out_va[m] = a_vals[i];             /* :97 */                                 \
            out_vb[m] = b_vals[j];             /* :98 */                                 \
          }                                                                              \
          ++m;                                                                           \
        }                                                                                \
    return m;                                                                            \
  }                                                                                      \
                                                                                         \
  uint64_t dwo_sort_join_##SUF(const T *a_keys, const T *a_vals, uint64_t na,            \
                               const T *b_keys, const T *b_vals, uint64_t nb,            \
                               T *out_key, T *out_va, T *out_vb, uint64_t cap) {         \
    kv_##SUF *A = (kv_##SUF *)malloc((na ? na : 1) * sizeof(kv_##SUF));                  \
    kv_##SUF *B = (kv_##SUF *)malloc((nb ? nb : 1) * sizeof(kv_##SUF));                  \
    if (!A || !B) { free(A); free(B); return UINT64_MAX; }                               \
    for (uint64_t i = 0; i < na; ++i) { A[i].k = a_keys[i]; A[i].v = a_vals[i]; }        \
    for (uint64_t j = 0; j < nb; ++j) { B[j].k = b_keys[j]; B[j].v = b_vals[j]; }        \
    qsort(A, na, sizeof(kv_##SUF), cmp_kv_##SUF);                                        \
    qsort(B, nb, sizeof(kv_##SUF), cmp_kv_##SUF);                                        \
    uint64_t m = 0, i = 0, j = 0;                                                        \
    while (i < na && j < nb) {                                                           \
      if (A[i].k < B[j].k) { ++i; continue; }                                            \
      if (B[j].k < A[i].k) { ++j; continue; }                                            \
      uint64_t ie = i, je = j;                                                           \
      while (ie < na && A[ie].k == A[i].k) ++ie;                                         \
      while (je < nb && B[je].k == B[j].k) ++je;                                         \
      for (uint64_t x = i; x < ie; ++x)        /* the per-key cross product that */      \
        for (uint64_t y = j; y < je; ++y) {    /* seq_join's double loop yields  */      \
          if (m < cap) { out_key[m] = A[x].k; out_va[m] = A[x].v; out_vb[m] = B[y].v; }  \
          ++m;                                                                           \
        }                                                                                \
      i = ie; j = je;                                                                    \
    }                                                                                    \
    free(A); free(B);                                                                    \
    return m;                                                                            \
  }                                                                                      \
                                                                                         \
  int dwo_rows_equal_##SUF(const T *k1, const T *a1, const T *b1, uint64_t n1,           \
                           const T *k2, const T *a2, const T *b2, uint64_t n2) {         \
    if (n1 != n2) return 0;                    /* vector== compares sizes first */       \
    row_##SUF *r1 = (row_##SUF *)malloc((n1 ? n1 : 1) * sizeof(row_##SUF));              \
    row_##SUF *r2 = (row_##SUF *)malloc((n2 ? n2 : 1) * sizeof(row_##SUF));              \
    if (!r1 || !r2) { free(r1); free(r2); return -1; }                                   \
    for (uint64_t i = 0; i < n1; ++i) {        /* to_row_store, :33-44 */                \
      r1[i].k = k1[i]; r1[i].a = a1[i]; r1[i].b = b1[i];                                 \
      r2[i].k = k2[i]; r2[i].a = a2[i]; r2[i].b = b2[i];                                 \
    }                                                                                    \
    qsort(r1, n1, sizeof(row_##SUF), cmp_row_##SUF);   /* eq(), :106-114 */              \
    qsort(r2, n2, sizeof(row_##SUF), cmp_row_##SUF);                                     \
    int same = 1;                                                                        \
    for (uint64_t i = 0; i < n1 && same; ++i)                                            \
      same = r1[i].k == r2[i].k && r1[i].a == r2[i].a && r1[i].b == r2[i].b;             \
    free(r1); free(r2);                                                                  \
    return same;                                                                         \
  }

DEFINE_JOIN_FUNCS(u32, uint32_t)
DEFINE_JOIN_FUNCS(u64, uint64_t)

/* ------------------------------------------------------------------------ */
/* SimpleNonOwningHashTable -- hashtable.hpp:5-93                            */
/* ------------------------------------------------------------------------ */

#define ELEM_SZ 32u   /* hashtable.hpp:68: CHAR_BIT * sizeof(uint32_t) */

static inline uint32_t table_hash(const dwo_table *t, uint32_t key) {
  if (t->hash_kind == DWO_HASH_MODULO) return (uint32_t)(key % t->size);
  return (uint32_t)dwo_murmur3_slot(key, t->seed, 4, t->size);
}

/* sycl::ext::intel::ctz<uint32_t>: trailing zero count, 32 for x == 0. */
static inline uint32_t ctz32(uint32_t x) { return x ? (uint32_t)__builtin_ctz(x) : 32u; }

static uint32_t update_bitmask(dwo_table *t, uint32_t at) {   /* hashtable.hpp:70-92 */
  uint32_t major_idx = at / ELEM_SZ;
  uint8_t minor_idx = (uint8_t)(at % ELEM_SZ);
  for (;;) {
    uint32_t mask = (uint32_t)1 << minor_idx;
    uint32_t present = __atomic_fetch_or(&t->bitmask[major_idx], mask, __ATOMIC_SEQ_CST);
    if (!(present & mask)) return major_idx * ELEM_SZ + minor_idx;      /* :78-80 */
    uint32_t occupied = ctz32(~(present >> minor_idx));                 /* :82-83 */
    if (occupied + minor_idx >= ELEM_SZ ||
        (uint64_t)major_idx * ELEM_SZ + minor_idx >= t->size) {         /* :84-85 */
      major_idx = (uint32_t)(((uint64_t)major_idx + 1) % t->bitmask_sz); /* :86 */
      minor_idx = 0;                                                    /* :87 */
    } else {
      minor_idx = (uint8_t)(minor_idx + occupied);                      /* :89 */
    }
  }
}

uint32_t dwo_table_insert(dwo_table *t, uint32_t key, uint32_t val) {  /* :15-21 */
  uint32_t pos = update_bitmask(t, table_hash(t, key));
  t->keys[pos] = key;
  t->vals[pos] = val;
  return pos;
}

static inline int bit_present(const dwo_table *t, uint32_t pos) {
  return (t->bitmask[pos / ELEM_SZ] & ((uint32_t)1 << (pos % ELEM_SZ))) != 0;
}

int dwo_table_at(const dwo_table *t, uint32_t key, uint32_t *val) {    /* :23-40 */
  uint32_t pos = table_hash(t, key);
  const uint32_t start = pos;
  int present = bit_present(t, pos);
  while (present) {
    if (t->keys[pos] == key) { *val = t->vals[pos]; return 1; }        /* first hit wins */
    pos = (uint32_t)(((uint64_t)pos + 1) % t->size);
    if (pos == start) break;
    present = bit_present(t, pos);
  }
  return 0;
}

int dwo_table_has(const dwo_table *t, uint32_t key) {                  /* :42-58 */
  uint32_t dummy;
  return dwo_table_at(t, key, &dummy);
}

/* ------------------------------------------------------------------------ */
/* Join::_run timed region -- join/join.cpp:30-113                           */
/* ------------------------------------------------------------------------ */

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);   /* std::chrono::steady_clock */
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int dwo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int table_alloc(dwo_table *t, uint64_t n, uint32_t seed, uint32_t key_fill) {
  t->size = n * 2;                                   /* join.cpp:30 */
  t->bitmask_sz = (t->size + 31) / 32;               /* join.cpp:31 ceil(ht_size/32) */
  if (t->bitmask_sz == 0) t->bitmask_sz = 1;
  t->hash_kind = DWO_HASH_MURMUR;                    /* join.cpp:32 */
  t->seed = seed;
  t->bitmask = (uint32_t *)calloc(t->bitmask_sz, 4); /* join.cpp:36 */
  t->vals = (uint32_t *)calloc(t->size ? t->size : 1, 4);   /* join.cpp:37 */
  t->keys = (uint32_t *)malloc((t->size ? t->size : 1) * 4);
  if (!t->bitmask || !t->vals || !t->keys) return -1;
  memset(t->keys, (int)(key_fill & 0xff), t->size * 4);     /* join.cpp:38 / hash_build.cpp:28 */
  return 0;
}

static void table_free(dwo_table *t) { free(t->bitmask); free(t->vals); free(t->keys); }

int dwo_join_build_probe_u32(const uint32_t *a_keys, const uint32_t *a_vals, uint64_t na,
                             const uint32_t *b_keys, const uint32_t *b_vals, uint64_t nb,
                             uint32_t murmur_seed,
                             uint32_t *out_key, uint32_t *out_present, uint32_t *out_val,
                             dwo_join_timing *timing) {
  dwo_table t;
  if (na == 0) {                                /* nothing to build: every probe misses */
    memset(out_key, 0xff, nb * 4); memset(out_present, 0xff, nb * 4); memset(out_val, 0xff, nb * 4);
    if (timing) { timing->build_us = timing->probe_us = timing->host_us = 0; timing->threads = dwo_max_threads(); }
    return 0;
  }
  if (table_alloc(&t, na, murmur_seed, 0xff) != 0) { table_free(&t); return -1; }
  memset(out_key, 0xff, nb * 4);                /* join.cpp:41-43, vector(buf_size, -1) */
  memset(out_present, 0xff, nb * 4);
  memset(out_val, 0xff, nb * 4);

  double t0 = now_us();                         /* join.cpp:59 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)na; ++i)     /* join_build, join.cpp:69-75 */
    dwo_table_insert(&t, a_keys[i], a_vals[i]);
  double t1 = now_us();                         /* join.cpp:78 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nb; ++i) {   /* join_probe, join.cpp:93-103 */
    uint32_t v;
    if (dwo_table_at(&t, b_keys[i], &v)) {
      out_key[i] = b_keys[i];
      out_present[i] = v;
      out_val[i] = b_vals[i];
    }
  }
  double t2 = now_us();                         /* join.cpp:105 */
  if (timing) {
    timing->build_us = t1 - t0;
    timing->probe_us = t2 - t1;
    timing->host_us = t2 - t0;
    timing->threads = dwo_max_threads();
  }
  table_free(&t);
  return 0;
}

uint64_t dwo_compact_u32(const uint32_t *out_key, const uint32_t *out_present,
                         const uint32_t *out_val, uint64_t n,
                         uint32_t *res_k, uint32_t *res_present, uint32_t *res_val) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i)              /* join.cpp:123-129 */
    if (out_key[i] != 0xFFFFFFFFu) {
      res_k[m] = out_key[i];
      res_present[m] = out_present[i];
      res_val[m] = out_val[i];
      ++m;
    }
  return m;
}

/* ------------------------------------------------------------------------ */
/* HashBuild::_run -- hash/hash_build.cpp:19-81                              */
/* ------------------------------------------------------------------------ */

uint64_t dwo_hash_build_check_u32(const uint32_t *src, uint64_t n, uint32_t murmur_seed,
                                  double *build_us, int *threads) {
  dwo_table t;
  if (n == 0) { if (build_us) *build_us = 0; if (threads) *threads = dwo_max_threads(); return 0; }
  if (table_alloc(&t, n, murmur_seed, 0x00) != 0) { table_free(&t); return UINT64_MAX; }
  double t0 = now_us();                         /* hash_build.cpp:35 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i)      /* hash_build.cpp:43-49 */
    dwo_table_insert(&t, src[i], src[i]);
  double t1 = now_us();                         /* hash_build.cpp:52 */
  uint64_t found = 0;
#pragma omp parallel for schedule(static) reduction(+ : found)
  for (int64_t i = 0; i < (int64_t)n; ++i)      /* hash_build.cpp:69-75 */
    found += (uint64_t)dwo_table_has(&t, src[i]);
  if (build_us) *build_us = t1 - t0;
  if (threads) *threads = dwo_max_threads();
  table_free(&t);
  return found;
}

/* ------------------------------------------------------------------------ */
/* OmniSci one-to-many table -- omnisci_hashtable.hpp                        */
/* ------------------------------------------------------------------------ */

static int cmp_u32(const void *x, const void *y) {
  uint32_t p = *(const uint32_t *)x, q = *(const uint32_t *)y;
  return p < q ? -1 : p > q;
}

int dwo_omnisci_join_u32(const uint32_t *a_keys, uint64_t na,
                         const uint32_t *b_keys, uint64_t nb,
                         uint64_t *ids, uint64_t *match_off, uint64_t *match_cnt) {
  const uint32_t empty_key = 0xFFFFFFFFu;        /* join_omnisci.cpp:78 */
  /* count_distinct, join_omnisci.cpp:10-13 */
  uint64_t distinct = 0;
  {
    uint32_t *tmp = (uint32_t *)malloc((na ? na : 1) * 4);
    if (!tmp) return -1;
    memcpy(tmp, a_keys, na * 4);
    qsort(tmp, na, 4, cmp_u32);
    for (uint64_t i = 0; i < na; ++i) distinct += (i == 0 || tmp[i] != tmp[i - 1]);
    free(tmp);
  }
  const uint64_t ht_size = distinct * 2;         /* join_omnisci.cpp:69 */
  for (uint64_t j = 0; j < nb; ++j) { match_off[j] = 0; match_cnt[j] = 0; }
  if (ht_size == 0) return 0;
  uint32_t *ht = (uint32_t *)malloc(ht_size * 4);
  uint64_t *cnt = (uint64_t *)calloc(ht_size, 8);
  uint64_t *pos = (uint64_t *)calloc(ht_size, 8);
  if (!ht || !cnt || !pos) { free(ht); free(cnt); free(pos); return -1; }
  for (uint64_t s = 0; s < ht_size; ++s) ht[s] = empty_key;   /* ctor kernel :58-77 */

  for (uint64_t i = 0; i < na; ++i) {            /* build_table :80-108 */
    uint64_t h = a_keys[i] % ht_size, p = h;     /* SimpleHasher, hashfunctions.hpp:43-49 */
    do {
      if (ht[p] == empty_key) { ht[p] = a_keys[i]; break; }   /* CAS succeeded */
      if (ht[p] == a_keys[i]) break;                           /* expected_key == ks[i] */
      p = (p + 1) % ht_size;
    } while (p != h);
  }
  for (uint64_t i = 0; i < na; ++i) {            /* build_count_buffer :223-248 */
    uint64_t h = a_keys[i] % ht_size, p = h;
    do {
      if (ht[p] == a_keys[i]) { cnt[p]++; break; }
      p = (p + 1) % ht_size;
    } while (p != h);
  }
  {                                              /* build_pos_buffer :250-261 */
    uint64_t run = 0;
    for (uint64_t s = 0; s < ht_size; ++s) { pos[s] = run; run += cnt[s]; }  /* exclusive_scan */
    for (uint64_t s = 0; s < ht_size; ++s) cnt[s] = 0;                       /* host zeroing loop */
  }
  for (uint64_t i = 0; i < na; ++i) {            /* build_id_buffer :110-147 */
    uint64_t h = a_keys[i] % ht_size, p = h;
    do {
      if (ht[p] == a_keys[i]) { ids[pos[p] + cnt[p]++] = i; break; }
      p = (p + 1) % ht_size;
    } while (p != h);
  }
  for (uint64_t j = 0; j < nb; ++j) {            /* lookup :149-192 */
    uint64_t h = b_keys[j] % ht_size, p = h;
    int found = ht[p] == b_keys[j];
    if (!found) {
      p = (h + 1) % ht_size;
      for (;;) {
        if (ht[p] == b_keys[j]) { found = 1; break; }
        if (p == h || ht[p] == empty_key) break;
        p = (p + 1) % ht_size;
      }
    }
    if (found) { match_off[j] = pos[p]; match_cnt[j] = cnt[p]; }
  }
  free(ht); free(cnt); free(pos);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* generators -- common/common.cpp:7-36                                      */
/* ------------------------------------------------------------------------ */

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

void dwo_make_unique_random(uint64_t size, uint64_t seed, uint32_t *out) {
  /* common.cpp:7-20 draws dist(gen) % (size*10) into a std::set until it holds
   * `size` values, then returns them in set (ascending) order.  Same outcome
   * here with a membership bitmap over [0, 10*size). */
  if (size == 0) return;
  const uint64_t domain = size * 10;             /* size_multiplier = 10, :8 */
  uint8_t *seen = (uint8_t *)calloc((domain + 7) / 8, 1);
  uint64_t s = seed * 0x2545f4914f6cdd1dull + 1, have = 0;
  while (have < size) {
    uint64_t v = splitmix64(&s) % domain;
    if (!(seen[v >> 3] & (1u << (v & 7)))) { seen[v >> 3] |= (uint8_t)(1u << (v & 7)); ++have; }
  }
  uint64_t w = 0;
  for (uint64_t v = 0; v < domain; ++v)
    if (seen[v >> 3] & (1u << (v & 7))) out[w++] = (uint32_t)v;
  free(seen);
}

void dwo_make_random_u32(uint64_t size, uint64_t seed, uint32_t lo, uint32_t hi, uint32_t *out) {
  uint64_t s = seed * 0x9e3779b97f4a7c15ull + 7;     /* common.hpp:31-40 */
  const uint64_t span = (uint64_t)hi - lo + 1;
  for (uint64_t i = 0; i < size; ++i) out[i] = lo + (uint32_t)(splitmix64(&s) % span);
}
