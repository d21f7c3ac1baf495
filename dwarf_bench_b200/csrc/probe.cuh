// probe.cuh -- streaming probe kernels and device-side output materialisation.
//
// Replaces kernel `join_probe` (join/join.cpp:80-104) = SimpleNonOwningHashTable::at
// (common/dpcpp/hashtable.hpp:23-40), `hash_build_check` (hash/hash_build.cpp:61-76) = ::has
// (hashtable.hpp:42-58), and the HOST compaction loop join/join.cpp:119-129, which here runs on
// the device: per-tile match counts -> single-pass decoupled-lookback scan -> compacted rows
// written in probe-row order (the order the reference's host loop produces).
//
// Work decomposition: a CTA takes TILES of THREADS*ITEMS consecutive probe rows in striped
// order (row = tile_base + j*THREADS + t), so every column load and every output store of a
// warp is one contiguous 128 B (4-byte keys) / 256 B (8-byte keys) request, and each thread has
// ITEMS independent random-sector loads in flight.
#pragma once
#include <type_traits>

#include "table.cuh"

namespace dwj {

enum ProbeMode { PROBE_ALIGNED = 0, PROBE_PAIRS = 1, PROBE_COUNT = 2, PROBE_CONTAINS = 3 };

template <int W> struct ProbeArgs {
  using K = typename KeyT<W>::type;
  const K *keys;
  const K *vals;           // may be null (CONTAINS / COUNT)
  uint64_t n;
  const void *table;
  uint64_t bucket_mask;
  uint64_t seed;
  K *out_key;              // may be null in PAIRS mode
  K *out_build_val;
  K *out_probe_val;
  uint32_t *out_flags;     // CONTAINS
  uint64_t capacity;       // PAIRS
  unsigned long long *n_matches;   // PAIRS / COUNT (device)
  unsigned long long *tile_state;  // PAIRS: [0] = ticket counter, [1..] = lookback descriptors
  uint64_t num_tiles;
  const Seg *segs;                 // segmented input (table.cuh), staged PAIRS kernel only; n_segs == 0: rows [0, n)
  uint32_t n_segs;
  const unsigned long long *n_dev; // non-null: the row count lives on the device (input produced by a filtered
                                   // partition pass); `n` is then an upper bound that sizes the grid
  const unsigned int *hot_keys;    // staged kernel: non-null and *hot_keys != 0 => the probe keys are skewed
                                   // (probe_skew_sample_kernel): bucket sectors are then allowed into L1
  const K *runs;                   // non-null: one-to-many table (csr.cuh) -- the table holds distinct keys and a slot's
                                   // payload is the granule index of the key's run [count | payloads...] in runs[]
};
template <int W> DWJ_D uint64_t probe_rows(const ProbeArgs<W> &a) {
  return a.n_dev ? min((uint64_t)__ldg(a.n_dev), a.n) : a.n;
}

// ---- decoupled look-back ----------------------------------------------------------------------
// One 64-bit descriptor per tile: status in the top two bits, value in the low 62.
constexpr unsigned long long LB_AGGREGATE = 1ull << 62;   // value = this tile's own count
constexpr unsigned long long LB_PREFIX = 2ull << 62;      // value = inclusive prefix up to this tile
constexpr unsigned long long LB_VALUE_MASK = (1ull << 62) - 1;

DWJ_D unsigned long long lb_load(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DWJ_D void lb_store(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by all 32 lanes of one warp.  Tiles are numbered by a ticket counter, so every
// predecessor has already started and the spin below cannot deadlock.
DWJ_D unsigned long long lookback_exclusive_prefix(unsigned long long *desc, uint64_t tile,
                                                   unsigned long long tile_total) {
  const unsigned lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) lb_store(desc, LB_PREFIX | tile_total);
    return 0;
  }
  if (lane == 0) lb_store(desc + tile, LB_AGGREGATE | tile_total);
  unsigned long long exclusive = 0;
  int64_t window_end = (int64_t)tile - 1;
  for (;;) {
    const int64_t idx = window_end - lane;
    unsigned long long d = idx >= 0 ? lb_load(desc + idx) : LB_PREFIX;   // before tile 0: prefix 0
    while (__any_sync(0xffffffffu, (d >> 62) == 0)) {
      if ((d >> 62) == 0) d = lb_load(desc + idx);
    }
    const unsigned prefix_lanes = __ballot_sync(0xffffffffu, (d >> 62) == 2);
    const unsigned first = prefix_lanes ? __ffs(prefix_lanes) - 1 : 32;   // nearest tile with a full prefix
    unsigned long long v = lane <= first ? (d & LB_VALUE_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    exclusive += v;
    if (prefix_lanes) break;
    window_end -= 32;
  }
  if (lane == 0) lb_store(desc + tile, LB_PREFIX | ((exclusive + tile_total) & LB_VALUE_MASK));
  return exclusive;
}

// ---- chain walk ---------------------------------------------------------------------------------
// The home bucket is tested branch-free (4 compares + selects: straight-line code the compiler keeps in
// registers); only a FULL home bucket without a hit continues into the next sector, in a cold loop.
template <int W, class K>
DWJ_D bool find_first_overflow(const void *table, uint64_t mask, uint64_t b, K key, K &payload) {
  for (;;) {
    b = (b + 1) & mask;
    const Bucket<W> bk = load_bucket_ro<W>(table, b);
    bool hit = false;
#pragma unroll
    for (int i = Bucket<W>::SLOTS - 1; i >= 0; --i) {
      const bool m = bk.match(i, key);
      payload = m ? bk.payload(i) : payload;
      hit |= m;
    }
    if (hit || bk.any_empty()) return hit;      // chains never skip a bucket with a free slot
  }
}

// First hit (SimpleNonOwningHashTable::at): returns true and the payload of the lowest matching slot.
template <int W, class K>
DWJ_D bool find_first(const void *table, uint64_t mask, uint64_t b, const Bucket<W> &bk, K key, K &payload) {
  bool hit = false;
#pragma unroll
  for (int i = Bucket<W>::SLOTS - 1; i >= 0; --i) {     // descending, so the lowest slot wins
    const bool m = bk.match(i, key);
    payload = m ? bk.payload(i) : payload;
    hit |= m;
  }
  if (hit || bk.any_empty()) return hit;
  return find_first_overflow<W, K>(table, mask, b, key, payload);
}

// ---- warp-centric kernel: ALIGNED / CONTAINS / COUNT ---------------------------------------------------
// No inter-warp communication is needed for these shapes, so there are no CTA barriers at all: a warp takes
// tiles of 32*ITEMS consecutive rows (every column access is one contiguous 128 B / 256 B request) and the
// grid strides over tiles in order.  COUNT accumulates in registers and issues one atomic per warp.
template <int W, int MODE, bool UNIQUE, int ITEMS, bool FULL>
DWJ_D void simple_round(const ProbeArgs<W> &a, uint64_t base, unsigned lane, unsigned long long &local_count) {
  using K = typename KeyT<W>::type;
  constexpr K SENTINEL = ~(K)0;
  K key[ITEMS], pval[ITEMS];
  Bucket<W> bk[ITEMS];
  uint64_t hb[ITEMS];
  const uint32_t rows = FULL ? 0u : (uint32_t)min((uint64_t)(32 * ITEMS), a.n - base);
  const K *kp = a.keys + base + lane, *vp = a.vals + base + lane;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * 32 + lane < rows;
    key[j] = live ? load_stream(kp + j * 32) : SENTINEL;
    if constexpr (MODE == PROBE_ALIGNED) pval[j] = live ? load_stream(vp + j * 32) : SENTINEL;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    hb[j] = slot_hash(key[j], a.seed) & a.bucket_mask;
    bk[j] = load_bucket_ro<W>(a.table, hb[j]);
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint64_t row = base + (uint64_t)j * 32 + lane;
    K payload = SENTINEL;
    if constexpr (MODE == PROBE_COUNT) {
      if (key[j] != SENTINEL) {
        if (UNIQUE) local_count += find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], payload) ? 1u : 0u;
        else local_count += find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], payload) && payload != SENTINEL
                                ? (unsigned long long)__ldg(a.runs + (uint64_t)payload * CsrGeom<W>::G) : 0ull;   // the run's header
      }
    } else {
      bool hit = find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], payload) && key[j] != SENTINEL;
      if (MODE == PROBE_ALIGNED && a.runs) {        // one-to-many table: ONE of the key's build rows (the first of its run)
        hit = hit && payload != SENTINEL;
        if (hit) payload = __ldg(a.runs + (uint64_t)payload * CsrGeom<W>::G + 1);
      }
      if (!(FULL || j * 32 + lane < rows)) continue;
      if constexpr (MODE == PROBE_CONTAINS) {
        store_stream(a.out_flags + row, hit ? 1u : 0u);
      } else {                                    // join/join.cpp:97-101, sentinel elsewhere (:41-43)
        store_stream(a.out_key + row, hit ? key[j] : SENTINEL);
        store_stream(a.out_build_val + row, hit ? payload : SENTINEL);
        store_stream(a.out_probe_val + row, hit ? pval[j] : SENTINEL);
      }
    }
  }
}

template <int W, int MODE, bool UNIQUE, int ITEMS>
__global__ void __launch_bounds__(256) probe_simple_kernel(ProbeArgs<W> a) {
  constexpr uint64_t WTILE = 32ull * ITEMS;
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
  a.n = probe_rows(a);
  const uint64_t tiles = (a.n + WTILE - 1) / WTILE;
  unsigned long long local_count = 0;
  for (uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < tiles; tile += warps_total) {
    const uint64_t base = tile * WTILE;
    if (base + WTILE <= a.n) simple_round<W, MODE, UNIQUE, ITEMS, true>(a, base, lane, local_count);
    else simple_round<W, MODE, UNIQUE, ITEMS, false>(a, base, lane, local_count);
  }
  if constexpr (MODE == PROBE_COUNT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, o);
    if (lane == 0 && local_count) atomicAdd(a.n_matches, local_count);
  }
}

// ---- CTA-tile PAIRS kernel for NON-unique build keys (every equal build row is emitted) -----------------
// The table is a one-to-many table (csr.cuh): a probe row finds its key once (first hit, as with unique keys) and reads
// the count and the first payloads of the key's run with one 16-byte load.  One look-back descriptor (or one atomic)
// per tile.  A tile whose result rows fit four per probe row is staged in shared memory and written with contiguous
// stores; a tile with more (high multiplicity) is emitted run by run, a warp copying each run cooperatively.
template <int W, bool ORDERED, int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) probe_pairs_multi_kernel(ProbeArgs<W> a) {
  using K = typename KeyT<W>::type;
  using V = typename std::conditional<W == 4, uint4, ulonglong2>::type;      // 16 bytes of staged rows
  constexpr int TILE = THREADS * ITEMS;
  constexpr int WARPS = THREADS / 32;
  constexpr int VEC = 16 / W;
  constexpr K SENTINEL = ~(K)0;
  __shared__ unsigned long long s_tile;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_warp_sums[ITEMS][WARPS];
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_tile = ORDERED ? atomicAdd(a.tile_state, 1ull) : (unsigned long long)blockIdx.x;   // ticket: look-back needs predecessors running
  __syncthreads();
  const uint64_t tile = s_tile;
  const uint64_t base = tile * TILE;
  a.n = probe_rows(a);
  const bool full = base + TILE <= a.n;

  K key[ITEMS], pval[ITEMS], pl[ITEMS][4], g[ITEMS];      // g: granule index of the key's run
  uint32_t cnt[ITEMS], off[ITEMS];
  {
    Bucket<W> bk[ITEMS];
    uint64_t hb[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint64_t row = base + (uint64_t)j * THREADS + t;
      const bool live = full || row < a.n;
      key[j] = live ? load_stream(a.keys + row) : SENTINEL;
      pval[j] = live ? load_stream(a.vals + row) : SENTINEL;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      hb[j] = slot_hash(key[j], a.seed) & a.bucket_mask;
      bk[j] = load_bucket_ro<W>(a.table, hb[j]);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      g[j] = SENTINEL;
      if (key[j] == SENTINEL || !find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], g[j])) g[j] = SENTINEL;
    }
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    pl[j][0] = pl[j][1] = pl[j][2] = pl[j][3] = SENTINEL;
    cnt[j] = g[j] != SENTINEL ? csr_head<K>(a.runs + (uint64_t)g[j] * CsrGeom<W>::G, pl[j]) : 0u;
    uint32_t incl = cnt[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    off[j] = incl - cnt[j];
    if (lane == 31) s_warp_sums[j][warp] = incl;
  }
  __syncthreads();
  uint32_t tile_total = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      if (w == (int)warp) off[j] += tile_total;      // everything before (j, warp) in row order
      tile_total += s_warp_sums[j][w];
    }
  }
  if (warp == 0) {
    unsigned long long excl = 0;
    if constexpr (ORDERED) {
      excl = lookback_exclusive_prefix(a.tile_state + 1, tile, tile_total);
      if (lane == 0 && tile == a.num_tiles - 1) *a.n_matches = excl + tile_total;
    } else {
      if (lane == 0 && tile_total) excl = atomicAdd(a.n_matches, (unsigned long long)tile_total);
    }
    if (lane == 0) s_base = excl;
  }
  __syncthreads();
  const unsigned long long out_base = s_base;
  const bool fits = out_base + tile_total <= a.capacity;
  // Common case -- the tile's result rows fit the staging buffer (4 per probe row): every thread drops its matches at
  // their tile-relative positions in shared memory and the CTA then writes the tile's rows with contiguous stores.
  // Straight from registers, lane l's i-th match lands 4 rows from lane l+1's: every store instruction of a warp
  // touched four times the sectors it filled.
  extern __shared__ __align__(16) unsigned char s_stage_raw[];
  constexpr uint32_t STAGE = TILE * 4;
  if (tile_total <= STAGE) {
    K *s_ob = reinterpret_cast<K *>(s_stage_raw), *s_op = s_ob + STAGE, *s_ok = s_op + STAGE;      // s_ok only exists with a key column
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (cnt[j] == 0) continue;
      const uint32_t o = off[j];
      if (cnt[j] == 4 && (o & (VEC - 1)) == 0) {
        // Exactly four matches at a 16-byte boundary (BASELINE config 3: every key four times): 16-byte shared stores.
        // Word by word, lane l writes word 4 l + i -- four lanes per bank on every store instruction.
        if constexpr (W == 4) {
          *reinterpret_cast<V *>(s_ob + o) = make_uint4(pl[j][0], pl[j][1], pl[j][2], pl[j][3]);
          *reinterpret_cast<V *>(s_op + o) = make_uint4(pval[j], pval[j], pval[j], pval[j]);
          if (a.out_key) *reinterpret_cast<V *>(s_ok + o) = make_uint4(key[j], key[j], key[j], key[j]);
        } else {
          *reinterpret_cast<V *>(s_ob + o) = make_ulonglong2(pl[j][0], pl[j][1]);
          *reinterpret_cast<V *>(s_ob + o + 2) = make_ulonglong2(pl[j][2], pl[j][3]);
          *reinterpret_cast<V *>(s_op + o) = make_ulonglong2(pval[j], pval[j]);
          *reinterpret_cast<V *>(s_op + o + 2) = make_ulonglong2(pval[j], pval[j]);
          if (a.out_key) {
            *reinterpret_cast<V *>(s_ok + o) = make_ulonglong2(key[j], key[j]);
            *reinterpret_cast<V *>(s_ok + o + 2) = make_ulonglong2(key[j], key[j]);
          }
        }
      } else if (cnt[j] <= 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < (int)cnt[j]) {
            s_ob[o + i] = pl[j][i];
            s_op[o + i] = pval[j];
            if (a.out_key) s_ok[o + i] = key[j];
          }
      } else {                                        // a longer run: read it where it lies
        const K *run = a.runs + (uint64_t)g[j] * CsrGeom<W>::G + 1;
        for (uint32_t i = 0; i < cnt[j]; ++i) {
          s_ob[o + i] = __ldg(run + i);
          s_op[o + i] = pval[j];
          if (a.out_key) s_ok[o + i] = key[j];
        }
      }
    }
    __syncthreads();
    for (uint32_t i = t; i < tile_total; i += THREADS) {
      if (fits || out_base + i < a.capacity) {
        store_stream(a.out_build_val + out_base + i, s_ob[i]);
        store_stream(a.out_probe_val + out_base + i, s_op[i]);
        if (a.out_key) store_stream(a.out_key + out_base + i, s_ok[i]);
      }
    }
    return;
  }
  // High multiplicity (more than 4 result rows per probe row on average): the warp emits its rows' runs one after the
  // other, 32 lanes copying a run with contiguous loads and stores -- possible because a key's payloads are contiguous.
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned long long run = (unsigned long long)(a.runs + (uint64_t)(cnt[j] ? g[j] : 0) * CsrGeom<W>::G + 1);
    unsigned m = __ballot_sync(0xffffffffu, cnt[j] != 0);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t c = __shfl_sync(0xffffffffu, cnt[j], src);
      const unsigned long long o = __shfl_sync(0xffffffffu, out_base + off[j], src);
      const K *rp = reinterpret_cast<const K *>(__shfl_sync(0xffffffffu, run, src));
      const K kk = (K)__shfl_sync(0xffffffffu, (unsigned long long)key[j], src);
      const K pv = (K)__shfl_sync(0xffffffffu, (unsigned long long)pval[j], src);
      for (uint32_t i = lane; i < c; i += 32) {
        if (fits || o + i < a.capacity) {
          if (a.out_key) store_stream(a.out_key + o + i, kk);
          store_stream(a.out_build_val + o + i, __ldg(rp + i));
          store_stream(a.out_probe_val + o + i, pv);
        }
      }
    }
  }
}

// ---- staged PAIRS kernel (unique build keys) -------------------------------------------------------------
// At B200 probe rates (> 100 rows/ns) one look-back descriptor per 1-2 K rows means > 100 descriptors per
// microsecond -- more than one 32-wide look-back window per L2 round trip, so the look-back distance grows until
// it throttles the kernel (profiles/r1_probe.md: the first PAIRS kernel ran 2.5x slower than ALIGNED, which does
// the same loads and stores without a scan).  Here a CTA owns a CHUNK of WARPS*SUB*ITEMS*32 consecutive rows and
// publishes ONE descriptor for it.  Inside the chunk the warps are fully decoupled: warp w probes its own
// contiguous slice in SUB rounds of ITEMS*32 rows and compacts its hits into its own slice of shared memory
// (ballot + popc, no barrier).  Two barriers per chunk frame the look-back; then every warp streams its staged
// rows out with contiguous 128 B stores.  At most one match per probe row (UNIQUE), so the staging never
// overflows.
//   ORDERED = true : chunk base from the decoupled look-back -> output in probe-row order (reference order)
//   ORDERED = false: chunk base from one atomicAdd on the match counter -> same multiset, chunk order arbitrary
// One round of a warp: 32*ITEMS consecutive rows starting at `base`.  FULL = no row of the round is past the end
// of the relation: no bounds predicates at all (interior rounds; the 64-bit compares and predicated loads of the
// guarded version were a third of the instruction stream).
template <int W, bool WITH_KEY, int ITEMS, bool FULL, bool HOT>
DWJ_D void staged_round(const ProbeArgs<W> &a, uint64_t base, uint64_t limit, unsigned lane, unsigned lt,
                        typename KeyT<W>::type *wb, typename KeyT<W>::type *wp, typename KeyT<W>::type *wk, uint32_t &staged) {
  using K = typename KeyT<W>::type;
  constexpr K SENTINEL = ~(K)0;
  K key[ITEMS], pval[ITEMS];
  Bucket<W> bk[ITEMS];
  uint64_t hb[ITEMS];
  const K *kp = a.keys + base + lane, *vp = a.vals + base + lane;
  const uint32_t rows = FULL ? 0u : (uint32_t)min((uint64_t)(32 * ITEMS), limit - base);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * 32 + lane < rows;
    key[j] = live ? load_stream(kp + j * 32) : SENTINEL;       // SENTINEL never matches (reserved key)
    pval[j] = live ? load_stream(vp + j * 32) : SENTINEL;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    hb[j] = slot_hash(key[j], a.seed) & a.bucket_mask;
    // HOT (grid-uniform, decided on the device): with skewed probe keys -- a few percent of all rows ask for the same
    // sectors -- every SM serves the hot buckets from its own L1; with uniform keys the staging buffers keep the carve-out
    if constexpr (HOT) bk[j] = load_bucket_ro<W>(a.table, hb[j]);
    else bk[j] = load_bucket_stream<W>(a.table, hb[j]);
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    K payload = SENTINEL;
    const bool hit = find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], payload) && key[j] != SENTINEL;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const uint32_t o = staged + __popc(m & lt);
      wb[o] = payload;
      wp[o] = pval[j];
      if constexpr (WITH_KEY) wk[o] = key[j];
    }
    staged += __popc(m);
  }
}

template <int W, bool ORDERED, bool WITH_KEY, int WARPS, int ITEMS, int SUB, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) probe_pairs_staged_kernel(ProbeArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr int WROWS = 32 * ITEMS * SUB;           // rows per warp
  constexpr int CHUNK = WARPS * WROWS;
  extern __shared__ __align__(16) unsigned char s_raw[];
  K *s_build = reinterpret_cast<K *>(s_raw);
  K *s_probe = s_build + CHUNK;
  K *s_key = s_probe + CHUNK;                       // only touched when WITH_KEY
  __shared__ unsigned long long s_chunk;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_wtot[WARPS];

  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_chunk = ORDERED ? atomicAdd(a.tile_state, 1ull) : (unsigned long long)blockIdx.x;
  __syncthreads();
  const uint64_t chunk = s_chunk;
  uint64_t chunk_base = chunk * CHUNK, limit = probe_rows(a);
  if (a.n_segs) {                                   // CTA-uniform: the chunk's rows live in one segment (possibly peer memory)
    const Seg sg = a.segs[find_segment(a.segs, a.n_segs, chunk)];
    a.keys = (const K *)sg.keys;
    a.vals = (const K *)sg.vals;
    chunk_base = (chunk - sg.first_unit) * CHUNK;
    limit = sg.rows;
  }
  const uint64_t warp_base = chunk_base + (uint64_t)warp * WROWS;
  K *wb = s_build + warp * WROWS, *wp = s_probe + warp * WROWS, *wk = s_key + warp * WROWS;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t staged = 0;                              // warp-uniform running count
  const bool hot = a.hot_keys && __ldg(a.hot_keys) != 0;

  // One branch per CTA, not per load: the two flavours of the loop are separate code (a select in front of every
  // bucket load cost the uniform case 23 %).
  if (warp_base + WROWS <= limit && !hot) {         // warp-uniform: the whole slice is inside the relation / segment
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub)
      staged_round<W, WITH_KEY, ITEMS, true, false>(a, warp_base + (uint64_t)sub * (32 * ITEMS), limit, lane, lt, wb, wp, wk, staged);
  } else if (warp_base + WROWS <= limit) {
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub)
      staged_round<W, WITH_KEY, ITEMS, true, true>(a, warp_base + (uint64_t)sub * (32 * ITEMS), limit, lane, lt, wb, wp, wk, staged);
  } else {
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub) {
      const uint64_t base = warp_base + (uint64_t)sub * (32 * ITEMS);
      if (base >= limit) break;
      if (hot) staged_round<W, WITH_KEY, ITEMS, false, true>(a, base, limit, lane, lt, wb, wp, wk, staged);
      else staged_round<W, WITH_KEY, ITEMS, false, false>(a, base, limit, lane, lt, wb, wp, wk, staged);
    }
  }
  if constexpr (!ORDERED) {
    // Unordered output: nothing ties the warps of a chunk together -- every warp reserves its own output range with
    // one atomicAdd and streams its rows out; no CTA barrier at all (29 % of the ordered kernel's stall samples).
    unsigned long long wbase = 0;
    if (lane == 0 && staged) wbase = atomicAdd(a.n_matches, (unsigned long long)staged);
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    __syncwarp();
    K *ob = a.out_build_val + wbase, *op = a.out_probe_val + wbase, *ok = a.out_key + wbase;
    if (wbase + staged <= a.capacity) {
      for (uint32_t i = lane; i < staged; i += 32) {
        store_stream(ob + i, wb[i]);
        store_stream(op + i, wp[i]);
        if constexpr (WITH_KEY) store_stream(ok + i, wk[i]);
      }
    } else {
      for (uint32_t i = lane; i < staged; i += 32) {
        if (wbase + i < a.capacity) {
          store_stream(ob + i, wb[i]);
          store_stream(op + i, wp[i]);
          if constexpr (WITH_KEY) store_stream(ok + i, wk[i]);
        }
      }
    }
    return;
  }
  if (lane == 0) s_wtot[warp] = staged;
  __syncthreads();

  // One descriptor per chunk.
  if (warp == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) total += s_wtot[w];
    unsigned long long excl = 0;
    if constexpr (ORDERED) {
      excl = lookback_exclusive_prefix(a.tile_state + 1, chunk, total);
      if (lane == 0 && chunk == a.num_tiles - 1) *a.n_matches = excl + total;
    } else {
      if (lane == 0 && total) excl = atomicAdd(a.n_matches, (unsigned long long)total);
    }
    if (lane == 0) s_base = excl;
  }
  __syncthreads();
  unsigned long long out_base = s_base;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) out_base += w < (int)warp ? s_wtot[w] : 0u;
  K *ob = a.out_build_val + out_base, *op = a.out_probe_val + out_base, *ok = a.out_key + out_base;
  if (out_base + staged <= a.capacity) {            // warp-uniform: no per-row capacity checks
    for (uint32_t i = lane; i < staged; i += 32) {
      store_stream(ob + i, wb[i]);
      store_stream(op + i, wp[i]);
      if constexpr (WITH_KEY) store_stream(ok + i, wk[i]);
    }
  } else {
    for (uint32_t i = lane; i < staged; i += 32) {
      if (out_base + i < a.capacity) {
        store_stream(ob + i, wb[i]);
        store_stream(op + i, wp[i]);
        if constexpr (WITH_KEY) store_stream(ok + i, wk[i]);
      }
    }
  }
}

// ---- hot probe keys ----------------------------------------------------------------------------------------------------
// One CTA samples 4096 probe keys (evenly spaced) into a sketch of 8192 bucket counters: with uniform keys no counter
// passes a handful, with Zipf(1.0) keys the hottest bucket collects several percent of the samples.  The verdict stays on
// the device (*flag), where the staged kernel reads it: no host synchronisation.
template <int W>
__global__ void __launch_bounds__(1024) probe_skew_sample_kernel(const typename KeyT<W>::type *keys, uint64_t n, const unsigned long long *n_dev,
                                                                 uint64_t bucket_mask, uint64_t seed, unsigned int *flag) {
  __shared__ unsigned int s_cnt[8192];
  __shared__ unsigned int s_max;
  for (int i = threadIdx.x; i < 8192; i += 1024) s_cnt[i] = 0;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const uint64_t rows = n_dev ? min((uint64_t)__ldg(n_dev), n) : n;
  const uint64_t stride = max(rows / 4096, (uint64_t)1);
  for (int i = 0; i < 4; ++i) {
    const uint64_t r = ((uint64_t)i * 1024 + threadIdx.x) * stride;
    if (r < rows) {
      const uint64_t bucket = slot_hash(__ldg(keys + r), seed) & bucket_mask;
      atomicAdd(&s_cnt[((uint32_t)bucket ^ (uint32_t)(bucket >> 32)) * 0x9E3779B1u >> 19], 1u);       // 13-bit sketch index
    }
  }
  __syncthreads();
  unsigned int m = 0;
  for (int i = threadIdx.x; i < 8192; i += 1024) m = max(m, s_cnt[i]);
  atomicMax(&s_max, m);
  __syncthreads();
  if (threadIdx.x == 0) *flag = s_max >= 24 ? 1u : 0u;     // >= 0.6 % of the samples in one bucket
}

}  // namespace dwj
