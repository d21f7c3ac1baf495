// aggregate.cuh -- hash aggregation (GROUP BY key, SUM(value)) into the join table's bucket layout.
//
// Replaces kernel `hash_build` of the GroupBy dwarf (groupby/groupby.cpp:60-72) = NonOwningHashTableNonBitmask::add
// (common/dpcpp/hashtable.hpp:136-153): claim the key's slot with a CAS on the key, then fetch_add the value.  Sums wrap
// at the value width, as the reference's uint32_t arithmetic does.
//
// The reference adds every row straight into the global table.  With few groups (the library facade uses 20,
// bench.cpp:80) that is hundreds of millions of atomics on a handful of addresses.  Here aggregation is two-level:
//   1. every CTA keeps a small open-addressing table in SHARED memory (AGG_SLOTS entries) and adds its rows there
//      (atomicCAS on the key, atomicAdd on the sum); a row whose key finds no place within AGG_PROBES steps goes to
//      level 2 directly, so high-cardinality inputs degrade to the reference's behaviour instead of failing;
//   2. at the end the CTA adds its partial sums to the global table: groups x CTAs atomics instead of one per row.
// The global table is the join's own bucket table (table.cuh): a slot holds (key, sum); keys are claimed in slot order
// with a CAS, so the occupied slots of a bucket stay a prefix and every probe kernel (dwj_probe_aligned / _contains) reads
// the result as SimpleNonOwningHashTable::at would (the GroupBy dwarf's second kernel, groupby.cpp:84-92).  A slot is
// born all-ones: the thread that claims a key adds (value + 1), which makes the sum start from zero without a second
// store racing with other adders.
#pragma once
#include <type_traits>

#include "table.cuh"

namespace dwj {

template <int W> struct AggArgs {
  const typename KeyT<W>::type *keys;
  const typename KeyT<W>::type *vals;
  uint64_t n;
  void *table;
  uint64_t bucket_mask;
  uint64_t seed;
};

constexpr int AGG_SLOTS = 2048, AGG_PROBES = 8, AGG_THREADS = 256, AGG_ROWS = 8;

DWJ_D unsigned long long agg_cas(unsigned long long *p, unsigned long long cmp, unsigned long long val) { return atomicCAS(p, cmp, val); }
DWJ_D uint32_t agg_cas(uint32_t *p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
DWJ_D void agg_add(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
DWJ_D void agg_add(uint32_t *p, uint32_t v) { atomicAdd(p, v); }

// table[key] += val in a global bucket table (level 2 of the aggregation; the count pass of the one-to-many build, csr.cuh).
template <int W> DWJ_D void table_add(void *table, uint64_t bucket_mask, uint64_t seed, typename KeyT<W>::type key, typename KeyT<W>::type val) {
  using K = typename KeyT<W>::type;
  using A = typename std::conditional<W == 4, uint32_t, unsigned long long>::type;
  constexpr int SLOTS = Bucket<W>::SLOTS;
  constexpr K EMPTY = ~(K)0;
  if (key == EMPTY) return;                                   // the reserved key (table.cuh)
  uint64_t b = slot_hash(key, seed) & bucket_mask;
  for (;;) {
    A *slot = reinterpret_cast<A *>((char *)table + (b << 5));
#pragma unroll 1
    for (int i = 0; i < SLOTS; ++i) {
      A *kp = slot + 2 * i;
      A cur = *reinterpret_cast<volatile A *>(kp);
      if (cur == (A)EMPTY) {
        cur = agg_cas(kp, (A)EMPTY, (A)key);
        if (cur == (A)EMPTY) { agg_add(kp + 1, (A)(val + 1)); return; }      // claimed: all-ones + val + 1 == val
      }
      if (cur == (A)key) { agg_add(kp + 1, (A)val); return; }
    }
    b = (b + 1) & bucket_mask;                                // bucket full of other keys: next sector
  }
}
template <int W> DWJ_D void global_add(const AggArgs<W> &a, typename KeyT<W>::type key, typename KeyT<W>::type val) {
  table_add<W>(a.table, a.bucket_mask, a.seed, key, val);
}

template <int W>
__global__ void __launch_bounds__(AGG_THREADS) aggregate_kernel(AggArgs<W> a) {
  using K = typename KeyT<W>::type;
  using A = typename std::conditional<W == 4, uint32_t, unsigned long long>::type;
  constexpr K EMPTY = ~(K)0;
  __shared__ A s_key[AGG_SLOTS];
  __shared__ A s_sum[AGG_SLOTS];
  for (int i = threadIdx.x; i < AGG_SLOTS; i += AGG_THREADS) { s_key[i] = (A)EMPTY; s_sum[i] = 0; }
  __syncthreads();
  constexpr uint64_t TILE = (uint64_t)AGG_THREADS * AGG_ROWS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE + threadIdx.x;
    K k[AGG_ROWS], v[AGG_ROWS];
#pragma unroll
    for (int j = 0; j < AGG_ROWS; ++j) {
      const uint64_t i = base + (uint64_t)j * AGG_THREADS;
      k[j] = i < a.n ? load_stream(a.keys + i) : EMPTY;
      v[j] = i < a.n ? load_stream(a.vals + i) : (K)0;
    }
#pragma unroll
    for (int j = 0; j < AGG_ROWS; ++j) {
      if (k[j] == EMPTY) continue;
      uint32_t h = (uint32_t)(slot_hash(k[j], a.seed) >> 7) & (AGG_SLOTS - 1);       // bits the bucket index does not use first
      bool placed = false;
#pragma unroll 1
      for (int step = 0; step < AGG_PROBES && !placed; ++step) {
        A cur = s_key[h];
        if (cur == (A)EMPTY) cur = agg_cas(&s_key[h], (A)EMPTY, (A)k[j]);
        if (cur == (A)EMPTY || cur == (A)k[j]) { agg_add(&s_sum[h], (A)v[j]); placed = true; }
        h = (h + 1) & (AGG_SLOTS - 1);
      }
      if (!placed) global_add<W>(a, k[j], v[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AGG_SLOTS; i += AGG_THREADS)
    if (s_key[i] != (A)EMPTY) global_add<W>(a, (K)s_key[i], (K)s_sum[i]);
}

}  // namespace dwj
