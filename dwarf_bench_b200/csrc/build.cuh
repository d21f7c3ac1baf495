// build.cuh -- hash-table build kernel.
//
// Replaces kernel `join_build` (join/join.cpp:60-77) / `hash_build` (hash/hash_build.cpp:36-50),
// i.e. SimpleNonOwningHashTable::insert + update_bitmask (common/dpcpp/hashtable.hpp:15-21,70-92):
// claim the first free slot at or after hash(key); duplicate keys each take their own slot.
//
// The reference claims a slot with fetch_or on a bitmask word and then fills keys[] / vals[].  Here a row takes a
// TICKET: one atomicAdd on a per-bucket fill counter returns the slot index inside the 32-byte bucket, and the
// (key, payload) pair is then written with one plain fire-and-forget store.  One L2 round trip per row, no bucket
// snapshot, no retry; only rows whose home bucket is already full (ticket >= SLOTS, 3.7 % of the rows at load 0.5)
// hop to the next bucket.  Slots are handed out in order, so the occupied slots of a bucket always form a prefix
// (probe.cuh relies on that).  The counters live in a scratch array (4 B per bucket) that only the build touches.
//
// What bounds it (tools/build_bench.cu, 268 M rows into a 4 GB table, input grouped into 128 table regions):
//   tickets alone 209 G rows/s; slot stores alone 54 G rows/s -- an 8-byte store into a sector that is not in L2
//   goes through the write-miss path, which is far slower than the read-miss path; with the region's slice
//   brought into L2 by `prefetch.global.L2` while the PREVIOUS region is being built the same stores run at
//   130 G rows/s and the whole insert at 77 G rows/s (3.5 ms instead of 5.7 ms).  Hence the look-ahead below.
//   (The first version -- bucket snapshot, then one 64/128-bit CAS per row -- ran at 28-46 G rows/s everywhere.)
#pragma once
#include "table.cuh"

namespace dwj {

template <int W> struct BuildArgs {
  const typename KeyT<W>::type *keys;
  const typename KeyT<W>::type *vals;
  uint64_t n;
  void *table;
  unsigned int *fill;                  // [buckets] ticket counters, zeroed before the launch
  uint64_t bucket_mask;
  uint64_t seed;
  // Region look-ahead (input grouped by table region, dwj_api.cu): rows [offsets[r], offsets[r+1]) go to the table
  // slice [r * slice_bytes, (r+1) * slice_bytes).  regions <= 1: no look-ahead.
  const unsigned long long *offsets;   // [regions + 1], device
  uint32_t regions;
  uint64_t slice_bytes;
  // Segmented input (table.cuh): n_segs > 0 => rows come from segs[], `n` is the number of TILES, and segments
  // [r * segs_per_region, (r+1) * segs_per_region) belong to table region r (look-ahead).
  const Seg *segs;
  uint32_t n_segs;
  uint32_t segs_per_region;
  PassFilter filter;                   // rows of other key classes are skipped (table.cuh)
  const unsigned long long *n_dev;     // non-null: the row count lives on the device (a filtered partition pass produced
                                       // the input); `n` is then only an upper bound that sizes the grid
};

DWJ_D void store_slot(void *table, uint64_t b, uint32_t i, uint32_t k, uint32_t v) {
  unsigned long long *bp = (unsigned long long *)table + (b << 2);
  bp[i] = (unsigned long long)k | ((unsigned long long)v << 32);
}
DWJ_D void store_slot(void *table, uint64_t b, uint32_t i, uint64_t k, uint64_t v) {
  ulonglong2 *bp = (ulonglong2 *)table + (b << 1);
  bp[i] = make_ulonglong2(k, v);
}

DWJ_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// While the tiles of region r are being inserted, each of them pulls its share of region r+1's table slice (and of
// its ticket counters) into L2.  Only a hint: a wrong guess costs bandwidth, never correctness.
template <int W>
DWJ_D void prefetch_next_region(const BuildArgs<W> &a, uint64_t row0, uint64_t tile_rows) {
  uint32_t lo = 0, hi = a.regions;                  // offsets[lo] <= row0 < offsets[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(a.offsets + mid) <= row0) lo = mid; else hi = mid;
  }
  if (lo + 1 >= a.regions) return;
  const uint64_t start = __ldg(a.offsets + lo), end = __ldg(a.offsets + lo + 1);
  const uint64_t len = max(end - start, tile_rows), lines = a.slice_bytes >> 7;
  const uint64_t l0 = (row0 - start) * lines / len, l1 = min((row0 - start + tile_rows) * lines / len, lines);
  const char *tb = (const char *)a.table + (uint64_t)(lo + 1) * a.slice_bytes;
  const char *fb = (const char *)a.fill + (uint64_t)(lo + 1) * (a.slice_bytes >> 3);
  for (uint64_t l = l0 + threadIdx.x; l < l1; l += blockDim.x) prefetch_l2(tb + (l << 7));
  for (uint64_t l = (l0 >> 3) + threadIdx.x; l < ((l1 + 7) >> 3); l += blockDim.x) prefetch_l2(fb + (l << 7));
}

// Same look-ahead for segmented input: tile `tile` of segment `si` lies in region si / segs_per_region, whose tiles are
// [first_unit of its first segment, first_unit of the next region's first segment).
template <int W>
DWJ_D void prefetch_next_region_segs(const BuildArgs<W> &a, uint32_t si, uint64_t tile) {
  const uint32_t region = si / a.segs_per_region;
  if (region + 1 >= a.regions) return;
  const uint64_t t0 = __ldg(&a.segs[region * a.segs_per_region].first_unit);
  const uint64_t t1 = __ldg(&a.segs[(region + 1) * a.segs_per_region].first_unit);
  const uint64_t len = max(t1 - t0, (uint64_t)1), lines = a.slice_bytes >> 7;
  const uint64_t l0 = (tile - t0) * lines / len, l1 = min((tile - t0 + 1) * lines / len, lines);
  const char *tb = (const char *)a.table + (uint64_t)(region + 1) * a.slice_bytes;
  const char *fb = (const char *)a.fill + (uint64_t)(region + 1) * (a.slice_bytes >> 3);
  for (uint64_t l = l0 + threadIdx.x; l < l1; l += blockDim.x) prefetch_l2(tb + (l << 7));
  for (uint64_t l = (l0 >> 3) + threadIdx.x; l < ((l1 + 7) >> 3); l += blockDim.x) prefetch_l2(fb + (l << 7));
}

// A CTA takes one tile of 256*ROWS CONSECUTIVE rows, so the rows in flight across the GPU form one contiguous window
// of the input: when the engine has grouped the input by table region that window touches one L2-resident slice of
// the table.  The tickets of a thread's ROWS rows are in flight together, and so are the hops of the rows that
// overflow (one round trip per hop level, not per row).  From the third hop on a bucket's counter is read before it
// is bumped, so long chains of full buckets (heavily duplicated keys) are walked with plain loads and the counters
// cannot run away.
template <int W, int ROWS>
__global__ void __launch_bounds__(256) build_kernel(BuildArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr uint32_t SLOTS = Bucket<W>::SLOTS;
  constexpr uint64_t TILE = 256ull * ROWS;
  const uint64_t n_rows = a.n_dev ? min((uint64_t)__ldg(a.n_dev), a.n) : a.n;
  const uint64_t tiles = a.n_segs ? a.n : (n_rows + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint64_t base = tile * TILE + threadIdx.x, limit = n_rows;
    const K *kp = a.keys, *vp = a.vals;
    if (a.n_segs) {                                        // the tile's rows live in one segment (possibly peer memory)
      const uint32_t si = find_segment(a.segs, a.n_segs, tile);
      const Seg sg = a.segs[si];
      kp = (const K *)sg.keys;
      vp = (const K *)sg.vals;
      base = (tile - sg.first_unit) * TILE + threadIdx.x;
      limit = sg.rows;
      if (a.regions > 1) prefetch_next_region_segs<W>(a, si, tile);
    } else if (a.regions > 1) {
      prefetch_next_region<W>(a, tile * TILE, TILE);
    }
    K k[ROWS], v[ROWS];
    uint64_t b[ROWS];
    uint32_t t[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const uint64_t i = base + (uint64_t)r * 256;
      const bool live = i < limit;
      k[r] = live ? load_stream(kp + i) : ~(K)0;           // the reserved key (and a row past the end) is never stored
      v[r] = live ? load_stream(vp + i) : ~(K)0;
    }
    if (a.filter.mask) {                                   // grid-uniform: rows of other key classes count as absent
#pragma unroll
      for (int r = 0; r < ROWS; ++r) k[r] = pass_ok(k[r], a.seed, a.filter) ? k[r] : ~(K)0;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      b[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
      t[r] = k[r] != ~(K)0 ? atomicAdd(a.fill + b[r], 1u) : 0u;
    }
    uint32_t pending = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (k[r] == ~(K)0) continue;
      if (t[r] < SLOTS) store_slot(a.table, b[r], t[r], k[r], v[r]);
      else pending |= 1u << r;
    }
    for (uint32_t hop = 1; pending; ++hop) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (pending >> r & 1u) {
          b[r] = (b[r] + 1) & a.bucket_mask;               // linear probing at sector granularity
          t[r] = hop > 2 && __ldcg(a.fill + b[r]) >= SLOTS ? SLOTS : atomicAdd(a.fill + b[r], 1u);
        }
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if ((pending >> r & 1u) && t[r] < SLOTS) {
          store_slot(a.table, b[r], t[r], k[r], v[r]);
          pending &= ~(1u << r);
        }
      }
    }
  }
}

}  // namespace dwj
