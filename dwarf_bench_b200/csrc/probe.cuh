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
};

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
// First hit (SimpleNonOwningHashTable::at): returns true and the payload.  `bk` is the home
// bucket, already loaded by the caller so that ITEMS loads are in flight together.
template <int W, class K>
DWJ_D bool find_first(const void *table, uint64_t mask, uint64_t b, Bucket<W> bk, K key, K &payload) {
  for (;;) {
#pragma unroll
    for (int i = 0; i < Bucket<W>::SLOTS; ++i)
      if (bk.match(i, key)) { payload = bk.payload(i); return true; }
    if (bk.any_empty()) return false;      // chains never skip a bucket with a free slot
    b = (b + 1) & mask;
    bk = load_bucket_ro<W>(table, b);
  }
}

// All hits (seq_join semantics): calls emit(payload) for every equal build row, returns the count.
template <int W, class K, class F>
DWJ_D uint32_t for_each_match(const void *table, uint64_t mask, uint64_t b, Bucket<W> bk, K key, F emit) {
  uint32_t c = 0;
  for (;;) {
#pragma unroll
    for (int i = 0; i < Bucket<W>::SLOTS; ++i)
      if (bk.match(i, key)) { emit(bk.payload(i)); ++c; }
    if (bk.any_empty()) return c;
    b = (b + 1) & mask;
    bk = load_bucket_ro<W>(table, b);
  }
}

// ---- the probe kernel -------------------------------------------------------------------------------
template <int W, int MODE, bool UNIQUE, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) probe_kernel(ProbeArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr int TILE = THREADS * ITEMS;
  constexpr int WARPS = THREADS / 32;
  constexpr K SENTINEL = ~(K)0;
  __shared__ unsigned long long s_tile;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_warp_sums[ITEMS][WARPS];

  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;

  // Tile id: scheduling order for PAIRS (look-back needs every predecessor to be running),
  // plain block index otherwise.
  uint64_t tile;
  if constexpr (MODE == PROBE_PAIRS) {
    if (t == 0) s_tile = atomicAdd(a.tile_state, 1ull);
    __syncthreads();
    tile = s_tile;
  } else {
    tile = blockIdx.x;
  }
  const uint64_t base = tile * TILE;

  K key[ITEMS];
  Bucket<W> bk[ITEMS];
  uint64_t hb[ITEMS];
  bool live[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint64_t row = base + (uint64_t)j * THREADS + t;
    live[j] = row < a.n;
    key[j] = live[j] ? load_stream(a.keys + row) : SENTINEL;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    hb[j] = slot_hash(key[j], a.seed) & a.bucket_mask;
    if (live[j]) bk[j] = load_bucket_ro<W>(a.table, hb[j]);
  }

  if constexpr (MODE == PROBE_ALIGNED || MODE == PROBE_CONTAINS) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (!live[j]) continue;
      const uint64_t row = base + (uint64_t)j * THREADS + t;
      K payload = SENTINEL;
      const bool hit = find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], payload);
      if constexpr (MODE == PROBE_CONTAINS) {
        store_stream(a.out_flags + row, hit ? 1u : 0u);
      } else {                                    // join/join.cpp:97-101, sentinel elsewhere (:41-43)
        store_stream(a.out_key + row, hit ? key[j] : SENTINEL);
        store_stream(a.out_build_val + row, payload);
        store_stream(a.out_probe_val + row, hit ? load_stream(a.vals + row) : SENTINEL);
      }
    }
  } else {
  // ---- PAIRS / COUNT: per-item match counts ------------------------------------------------------
  uint32_t cnt[ITEMS];
  K first_payload[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    cnt[j] = 0;
    first_payload[j] = SENTINEL;
    if (!live[j]) continue;
    if (UNIQUE) {
      cnt[j] = find_first<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j], first_payload[j]) ? 1u : 0u;
    } else {
      cnt[j] = for_each_match<W, K>(a.table, a.bucket_mask, hb[j], bk[j], key[j],
                                    [&](K p) { if (first_payload[j] == SENTINEL) first_payload[j] = p; });
    }
  }

  // Exclusive scan over the tile in row order (j major, t minor).
  uint32_t off[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    uint32_t incl;
    if (UNIQUE) {
      const unsigned m = __ballot_sync(0xffffffffu, cnt[j] != 0);
      incl = __popc(m & (0xffffffffu >> (31 - lane)));
    } else {
      incl = cnt[j];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
    }
    off[j] = incl - cnt[j];
    if (lane == 31) s_warp_sums[j][warp] = incl;
  }
  __syncthreads();
  uint32_t running = 0, tile_total;
  {
    // Every thread folds the ITEMS*WARPS partial sums it needs (tiny: <= 64 values).
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const uint32_t s = s_warp_sums[j][w];
        if (w == (int)warp) off[j] += acc;       // sums of everything before (j, warp)
        acc += s;
      }
    }
    tile_total = acc;
    (void)running;
  }

  if constexpr (MODE == PROBE_COUNT) {
    if (t == 0 && tile_total) atomicAdd(a.n_matches, (unsigned long long)tile_total);
    return;
  }

  // ---- look-back: where does this tile's output start? --------------------------------------------
  if (warp == 0) {
    const unsigned long long excl = lookback_exclusive_prefix(a.tile_state + 1, tile, tile_total);
    if (lane == 0) {
      s_base = excl;
      if (tile == a.num_tiles - 1) *a.n_matches = excl + tile_total;
    }
  }
  __syncthreads();
  const unsigned long long out_base = s_base;

  // ---- compacted writes, probe-row order ---------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (cnt[j] == 0) continue;
    const uint64_t row = base + (uint64_t)j * THREADS + t;
    const K pv = load_stream(a.vals + row);
    unsigned long long o = out_base + off[j];
    if (UNIQUE || cnt[j] == 1) {
      if (o < a.capacity) {
        if (a.out_key) store_stream(a.out_key + o, key[j]);
        store_stream(a.out_build_val + o, first_payload[j]);
        store_stream(a.out_probe_val + o, pv);
      }
    } else {
      // Re-walk the (now cache-warm) chain and emit every equal build row.
      for_each_match<W, K>(a.table, a.bucket_mask, hb[j], load_bucket_ro<W>(a.table, hb[j]), key[j], [&](K p) {
        if (o < a.capacity) {
          if (a.out_key) store_stream(a.out_key + o, key[j]);
          store_stream(a.out_build_val + o, p);
          store_stream(a.out_probe_val + o, pv);
        }
        ++o;
      });
    }
  }
  }  // PAIRS / COUNT
}

}  // namespace dwj
