// build.cuh -- hash-table build kernel.
//
// Replaces kernel `join_build` (join/join.cpp:60-77) / `hash_build` (hash/hash_build.cpp:36-50),
// i.e. SimpleNonOwningHashTable::insert + update_bitmask (common/dpcpp/hashtable.hpp:15-21,70-92):
// claim the first free slot at or after hash(key); duplicate keys each take their own slot.
// Here the claim and the (key, payload) store are ONE atomic compare-and-swap on the whole slot
// (64-bit CAS for 4-byte keys, 128-bit CAS for 8-byte keys), so there is no separate bitmask and
// no window in which a slot is claimed but not yet filled.
#pragma once
#include <cooperative_groups.h>

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
template <int W, int ROWS, int MODE>
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
    if constexpr (MODE == 0) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        b[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
        bk[r] = load_bucket_cg<W>(a.table, b[r]);
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int slot = first_empty<W>(bk[r]);
        done[r] = k[r] == ~(K)0;                            // the reserved key (and rows past the end) is not stored
        if (!done[r] && slot < Bucket<W>::SLOTS) done[r] = cas_slot(a.table, b[r], slot, k[r], v[r]);
      }
    } else {                                                // blind CAS on slot 0: one round trip when the bucket is empty
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        b[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
        done[r] = k[r] == ~(K)0;
        if (!done[r]) done[r] = cas_slot(a.table, b[r], 0, k[r], v[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (!done[r]) insert_slow<W, K>(a.table, a.bucket_mask, b[r], k[r], v[r]);
  }
}

// ---- region-fused build (persistent, cooperative launch) ---------------------------------------------------------
// For a table split into L2-sized regions (input pre-partitioned by region, dwj_api.cu): clear a region's slice,
// grid-sync, insert that region's rows, move on.  The slice is written (all-ones) while it is being brought into L2
// and is still there when its rows arrive, so the inserts are L2 hits and the table costs ONE write-back per line
// instead of memset write-back + first-touch read + write-back (tools/gather_bench atomics: random 64-bit CAS runs
// at 138 G/s on an L2-resident table and 24 G/s on an HBM-resident one; the separate memset + build_kernel path got 40).
template <int W> struct RegionBuildArgs {
  const typename KeyT<W>::type *keys;          // region-major
  const typename KeyT<W>::type *vals;
  const unsigned long long *offsets;           // [regions + 1] row offsets of the regions (device)
  void *table;
  uint64_t bucket_mask;
  uint64_t seed;
  uint64_t slice_bytes;                        // table bytes per region
  uint32_t regions;
};

template <int W, int ROWS>
__global__ void __launch_bounds__(256) build_regions_kernel(RegionBuildArgs<W> a) {
  using K = typename KeyT<W>::type;
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr uint64_t TILE = 256ull * ROWS;
  const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x, gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
  auto clear_slice = [&](uint32_t r) {
    uint4 *slice = reinterpret_cast<uint4 *>((char *)a.table + (uint64_t)r * a.slice_bytes);
    for (uint64_t i = gtid; i < a.slice_bytes / 16; i += gthreads) slice[i] = ones;
  };
  clear_slice(0);
  for (uint32_t r = 0; r < a.regions; ++r) {
    // Linear probing may carry a row of region r over the slice boundary into the first buckets of slice r+1 (and
    // the last region wraps into slice 0, built long before): slice r+1 is therefore cleared BEFORE region r is
    // built, never after.
    if (r + 1 < a.regions) clear_slice(r + 1);
    grid.sync();                               // slices r and r+1 are clear (and L2-resident); all earlier regions are built
    const uint64_t row0 = a.offsets[r], row1 = a.offsets[r + 1];
    const uint64_t tiles = (row1 - row0 + TILE - 1) / TILE;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint64_t base = row0 + tile * TILE + threadIdx.x;
      K k[ROWS], v[ROWS];
      uint64_t b[ROWS];
      Bucket<W> bk[ROWS];
      bool done[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const uint64_t i = base + (uint64_t)q * 256;
        const bool live = i < row1;
        k[q] = live ? load_stream(a.keys + i) : ~(K)0;
        v[q] = live ? load_stream(a.vals + i) : ~(K)0;
      }
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        b[q] = slot_hash(k[q], a.seed) & a.bucket_mask;
        bk[q] = load_bucket_cg<W>(a.table, b[q]);
      }
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const int slot = first_empty<W>(bk[q]);
        done[q] = k[q] == ~(K)0;
        if (!done[q] && slot < Bucket<W>::SLOTS) done[q] = cas_slot(a.table, b[q], slot, k[q], v[q]);
      }
#pragma unroll
      for (int q = 0; q < ROWS; ++q)
        if (!done[q]) insert_slow<W, K>(a.table, a.bucket_mask, b[q], k[q], v[q]);
    }
    // No barrier here: the next iteration clears slice r+2, which region r cannot reach.
  }
}

}  // namespace dwj
