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

// One CAS attempt on slot i of bucket b; returns true when the row now owns the slot.
DWJ_D bool cas_slot(void *table, uint64_t b, int i, uint32_t k, uint32_t v) {
  unsigned long long *bp = (unsigned long long *)table + (b << 2);
  return atomicCAS(bp + i, ~0ull, (unsigned long long)k | ((unsigned long long)v << 32)) == ~0ull;
}
DWJ_D bool cas_slot(void *table, uint64_t b, int i, uint64_t k, uint64_t v) {
  unsigned __int128 *bp = (unsigned __int128 *)table + (b << 1);
  const unsigned __int128 empty = ~(unsigned __int128)0;
  return atomicCAS(bp + i, empty, (unsigned __int128)k | ((unsigned __int128)v << 64)) == empty;
}

// First slot the snapshot shows as empty (SLOTS when the bucket looks full).  Occupied slots form a prefix.
template <int W> DWJ_D int first_empty(const Bucket<W> &bk) {
  int i = Bucket<W>::SLOTS;
#pragma unroll
  for (int s = Bucket<W>::SLOTS - 1; s >= 0; --s) i = bk.empty(s) ? s : i;
  return i;
}

// Cold path: the optimistic CAS lost its slot to another row (or the home bucket was full).  Re-read and walk on;
// a failed CAS means another row took the slot meanwhile, and slots are never emptied again.
template <int W, class K>
DWJ_D void insert_slow(void *table, uint64_t mask, uint64_t b, K k, K v) {
  for (;;) {
    const Bucket<W> bk = load_bucket_cg<W>(table, b);
#pragma unroll
    for (int i = 0; i < Bucket<W>::SLOTS; ++i)
      if (bk.empty(i) && cas_slot(table, b, i, k, v)) return;
    b = (b + 1) & mask;                 // linear probing at sector granularity
  }
}

// A CTA takes tiles of 256*ROWS CONSECUTIVE rows (tile = blockIdx, grid-stride over tiles), so the rows in
// flight across the GPU form one contiguous window of the input: when the engine has pre-partitioned the input by
// table region (dwj_api.cu) that window touches one L2-resident slice of the table.  ROWS independent rows per
// thread, in three phases so that the long latencies overlap instead of adding up: all home-bucket sectors are
// requested, then ONE optimistic CAS per row is issued on the first slot its snapshot shows as empty (the CAS
// round trips of the ROWS rows are in flight together -- the first version checked each result before issuing
// the next and spent 87 % of its stall samples there), and only rows that lost their slot take the retry loop.
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
      k[r] = live ? load_stream(a.keys + i) : ~(K)0;       // the reserved key is never stored
      v[r] = live ? load_stream(a.vals + i) : ~(K)0;
    }
    bool done[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      b[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
      bk[r] = load_bucket_cg<W>(a.table, b[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int slot = first_empty<W>(bk[r]);
      done[r] = k[r] == ~(K)0;                              // the reserved key (and rows past the end) is not stored
      if (!done[r] && slot < Bucket<W>::SLOTS) done[r] = cas_slot(a.table, b[r], slot, k[r], v[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (!done[r]) insert_slow<W, K>(a.table, a.bucket_mask, b[r], k[r], v[r]);
  }
}

}  // namespace dwj
