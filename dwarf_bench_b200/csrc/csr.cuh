// csr.cuh -- one-to-many build for engines WITHOUT DWJ_FLAG_UNIQUE_BUILD_KEYS: distinct keys in the bucket table, the
// payloads of every key in one contiguous RUN.
//
// Replaces OmniSci::HashTable::{build_table, build_count_buffer, build_pos_buffer, build_id_buffer}
// (common/dpcpp/omnisci_hashtable.hpp:80-192): distinct-key CAS table -> per-slot counts -> exclusive scan ->
// positions -> id buffer grouped by key.  The same three steps here, each one kernel, no host round trip:
//   1. csr_count_kernel    every build row: table[key] += 1 (CAS on the key, atomicAdd on the payload field: aggregate.cuh)
//   2. csr_offsets_kernel  every bucket: its keys' counts become run offsets -- a warp scans the space its 32 buckets
//                          need and reserves it with ONE atomicAdd on a cursor (the order of the runs is irrelevant, so no
//                          global scan); the slot payload is replaced by the offset and the run's header is zeroed
//   3. csr_fill_kernel     every build row: look the key up, take the next position of its run (atomicAdd on the header,
//                          which thereby ends up holding the count) and store the payload there
// Run layout (array `runs` of K words, K = the engine's key/payload type): runs start on 32-byte GRANULES (one sector),
//   run = [ count | payload 0 | payload 1 | ... ]    occupying ceil((1 + count) * sizeof(K) / 32) granules
// and a slot's payload field holds the run's granule index.  A probe row therefore costs ONE bucket sector plus ONE more
// 32-byte load that carries the count and the first 7 (4-byte keys) / 3 (8-byte keys) payloads -- two gathers per row, and
// the L1TEX pipe serves one scattered sector per clock and SM.  The old layout -- every duplicate in its own slot, chains
// continued into the next bucket -- made every probe walk to the first bucket with a free slot: 2.5 dependent sectors on
// average at BASELINE config 3 (x4 duplicates fill whole buckets) in a divergent loop (12.6 ms; with runs on 16-byte
// granules, i.e. a third gather for the fourth payload, 5.8 ms).
// Space: at most (1 + 32/sizeof(K)) words per build row (all keys distinct), allocated with the engine.
#pragma once
#include "aggregate.cuh"
#include "build.cuh"
#include "probe.cuh"

namespace dwj {

// Rows of one build tile (256 * ROWS consecutive rows of the relation or of one segment), filter applied: a row that is
// past the end, carries the reserved key or belongs to another key class comes back as the reserved key.
template <int W, int ROWS>
DWJ_D void csr_load_tile(const BuildArgs<W> &a, uint64_t tile, uint64_t n_rows, typename KeyT<W>::type (&k)[ROWS],
                         typename KeyT<W>::type (&v)[ROWS], bool with_vals) {
  using K = typename KeyT<W>::type;
  constexpr uint64_t TILE = 256ull * ROWS;
  uint64_t base = tile * TILE + threadIdx.x, limit = n_rows;
  const K *kp = a.keys, *vp = a.vals;
  if (a.n_segs) {
    const Seg sg = a.segs[find_segment(a.segs, a.n_segs, tile)];
    kp = (const K *)sg.keys;
    vp = (const K *)sg.vals;
    base = (tile - sg.first_unit) * TILE + threadIdx.x;
    limit = sg.rows;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    const bool live = i < limit;
    k[r] = live ? load_stream(kp + i) : ~(K)0;
    v[r] = live && with_vals ? load_stream(vp + i) : ~(K)0;
    if (a.filter.mask) k[r] = pass_ok(k[r], a.seed, a.filter) ? k[r] : ~(K)0;
  }
}

template <int W, int ROWS>
__global__ void __launch_bounds__(256) csr_count_kernel(BuildArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr uint64_t TILE = 256ull * ROWS;
  const uint64_t n_rows = a.n_dev ? min((uint64_t)__ldg(a.n_dev), a.n) : a.n;
  const uint64_t tiles = a.n_segs ? a.n : (n_rows + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    K k[ROWS], v[ROWS];
    csr_load_tile<W, ROWS>(a, tile, n_rows, k, v, false);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (k[r] != ~(K)0) table_add<W>(a.table, a.bucket_mask, a.seed, k[r], (K)1);
  }
}

template <int W> struct CsrOffsetArgs {
  void *table;
  uint64_t buckets;
  typename KeyT<W>::type *runs;
  unsigned long long *cursor;        // granules handed out so far (zeroed before the launch)
  uint64_t cap_granules;
  unsigned int *overflow;            // raised when the runs do not fit (more rows than the engine was created for)
};

// One thread per bucket.  The occupied slots of a bucket are a prefix (aggregate.cuh claims them in order).
template <int W>
__global__ void __launch_bounds__(256) csr_offsets_kernel(CsrOffsetArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr int SLOTS = Bucket<W>::SLOTS;
  constexpr uint32_t G = CsrGeom<W>::G;
  constexpr K EMPTY = ~(K)0;
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const bool live = b < a.buckets;
  K *slot = reinterpret_cast<K *>((char *)a.table + ((live ? b : 0) << 5));
  K key[SLOTS], cnt[SLOTS];
  unsigned long long need = 0;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    key[i] = live ? slot[2 * i] : EMPTY;
    cnt[i] = live ? slot[2 * i + 1] : (K)0;
    if (key[i] != EMPTY) need += CsrGeom<W>::granules(cnt[i]);
  }
  unsigned long long incl = need;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 31 && total) base = atomicAdd(a.cursor, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  const bool fits = base + total <= a.cap_granules;          // warp-uniform
  if (!fits && lane == 31) atomicExch(a.overflow, 1u);
  unsigned long long off = base + incl - need;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    if (key[i] == EMPTY) continue;
    if (fits) {
      slot[2 * i + 1] = (K)off;
      a.runs[off * G] = 0;                                   // header: the fill kernel counts it up
      off += CsrGeom<W>::granules(cnt[i]);
    } else {
      slot[2 * i + 1] = EMPTY;                               // "no run": the key then joins with nothing (error raised on the host)
    }
  }
}

DWJ_D uint32_t csr_take(uint32_t *header) { return atomicAdd(header, 1u); }
DWJ_D unsigned long long csr_take(unsigned long long *header) { return atomicAdd(header, 1ull); }

template <int W, int ROWS>
__global__ void __launch_bounds__(256) csr_fill_kernel(BuildArgs<W> a, typename KeyT<W>::type *runs) {
  using K = typename KeyT<W>::type;
  using A = typename std::conditional<W == 4, uint32_t, unsigned long long>::type;
  constexpr uint32_t G = CsrGeom<W>::G;
  constexpr K EMPTY = ~(K)0;
  constexpr uint64_t TILE = 256ull * ROWS;
  const uint64_t n_rows = a.n_dev ? min((uint64_t)__ldg(a.n_dev), a.n) : a.n;
  const uint64_t tiles = a.n_segs ? a.n : (n_rows + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    K k[ROWS], v[ROWS];
    csr_load_tile<W, ROWS>(a, tile, n_rows, k, v, true);
    uint64_t hb[ROWS];
    Bucket<W> bk[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      hb[r] = slot_hash(k[r], a.seed) & a.bucket_mask;
      bk[r] = load_bucket_ro<W>(a.table, hb[r]);             // written by the previous kernels, immutable in this one
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      K off = EMPTY;
      if (k[r] == EMPTY || !find_first<W, K>(a.table, a.bucket_mask, hb[r], bk[r], k[r], off) || off == EMPTY) continue;
      K *run = runs + (uint64_t)off * G;
      const A pos = csr_take(reinterpret_cast<A *>(run));
      run[1 + pos] = v[r];
    }
  }
}

}  // namespace dwj
