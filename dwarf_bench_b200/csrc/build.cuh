// build.cuh -- hash-table build kernel.
//
// Replaces kernel `join_build` (join/join.cpp:60-77) / `hash_build` (hash/hash_build.cpp:36-50),
// i.e. SimpleNonOwningHashTable::insert + update_bitmask (common/dpcpp/hashtable.hpp:15-21,70-92):
// claim the first free slot at or after hash(key); duplicate keys each take their own slot.
// Here the claim and the (key, payload) store are ONE atomic compare-and-swap on the whole slot
// (64-bit CAS for 4-byte keys, 128-bit CAS for 8-byte keys), so there is no separate bitmask and
// no window in which a slot is claimed but not yet filled.
#pragma once
#include "table.cuh"

namespace dwj {

template <int W> struct BuildArgs {
  const typename KeyT<W>::type *keys;
  const typename KeyT<W>::type *vals;
  uint64_t n;
  void *table;
  uint64_t bucket_mask;
  uint64_t seed;
};

// Try every slot the snapshot shows as empty; a failed CAS means another row took it meanwhile.
// Slots are never emptied again, so when this returns false the bucket is full.
DWJ_D bool insert_into_bucket(void *table, uint64_t b, const Bucket<4> &bk, uint32_t k, uint32_t v) {
  unsigned long long *bp = (unsigned long long *)table + (b << 2);
  const unsigned long long mine = (unsigned long long)k | ((unsigned long long)v << 32);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (bk.empty(i) && atomicCAS(bp + i, ~0ull, mine) == ~0ull) return true;
  return false;
}
DWJ_D bool insert_into_bucket(void *table, uint64_t b, const Bucket<8> &bk, uint64_t k, uint64_t v) {
  unsigned __int128 *bp = (unsigned __int128 *)table + (b << 1);
  const unsigned __int128 empty = ~(unsigned __int128)0;
  const unsigned __int128 mine = (unsigned __int128)k | ((unsigned __int128)v << 64);
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (bk.empty(i) && atomicCAS(bp + i, empty, mine) == empty) return true;
  return false;
}

template <int W, class K>
DWJ_D void insert_row(void *table, uint64_t mask, uint64_t b, Bucket<W> bk, K k, K v) {
  if (k == ~(K)0) return;               // reserved empty marker (table.cuh)
  for (;;) {
    if (insert_into_bucket(table, b, bk, k, v)) return;
    b = (b + 1) & mask;                 // linear probing at sector granularity
    bk = load_bucket_cg<W>(table, b);
  }
}

// A CTA takes tiles of 256*ROWS CONSECUTIVE rows (tile = blockIdx, grid-stride over tiles), so the rows in
// flight across the GPU form one contiguous window of the input: when the engine has pre-partitioned the input by
// table region (dwj_api.cu) that window touches one L2-resident slice of the table.  ROWS independent rows per
// thread: all home-bucket sectors are requested first, then the CAS phase runs, so ROWS random-sector latencies
// overlap per thread.
template <int W, int ROWS>
__global__ void __launch_bounds__(256) build_kernel(BuildArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr uint64_t TILE = 256ull * ROWS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE + threadIdx.x;
    K k[ROWS], v[ROWS];
    uint64_t b[ROWS];
    Bucket<W> bk[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const uint64_t i = base + (uint64_t)r * 256;
      const bool live = i < a.n;
      k[r] = live ? load_stream(a.keys + i) : ~(K)0;       // the reserved key is skipped by insert_row
      v[r] = live ? load_stream(a.vals + i) : ~(K)0;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      b[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
      bk[r] = load_bucket_cg<W>(a.table, b[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) insert_row<W, K>(a.table, a.bucket_mask, b[r], bk[r], k[r], v[r]);
  }
}

}  // namespace dwj
