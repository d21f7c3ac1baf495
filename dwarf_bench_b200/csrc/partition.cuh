// partition.cuh -- radix partition of (key, payload) columns on bits of the key hash.
//
// No reference counterpart (the reference is single-device and never partitions).  Two users:
//   * the multi-GPU exchange (dwj_partition): partition id = top bits of an INDEPENDENT hash, so rows with equal
//     keys meet on one GPU and the local bucket index stays uniformly distributed;
//   * the engine's own L2-locality pass (csrc/dwj_api.cu, "regions"): partition id = top bits of the BUCKET INDEX
//     itself, so a partition's rows touch one contiguous slice of the table that fits in L2.  On B200 an L2 miss
//     fills a 128-byte line (profiles/r1_gather_microbench.md), so a random probe of an HBM-resident table costs
//     ~100 B of DRAM traffic per row; partitioning first costs 4 + 16 B per row of pure streaming instead.
//
// Three launches: histogram -> offsets (one block) -> scatter.  Neither hot kernel uses shared-memory atomics with a
// result: those cost ~2 cycles per LANE on this machine (64 cycles per warp instruction per SM), i.e. one per row caps
// a kernel at ~140 G rows/s whatever else it does.
//   histogram: <= 8 partitions: eight 8-bit counters packed in a 64-bit register per thread;
//              more: a private 8-bit counter per (thread, partition) in shared memory, laid out so lane l always
//              hits bank l; a row is LDS.U8 / IADD / STS.U8.  Both fold into 32-bit totals before a byte can overflow.
//   scatter  : the ranking of CUB's onesweep -- the lanes of a warp that go to the same partition find each other
//              with one ballot per partition-id bit, the lowest of them bumps a WARP-PRIVATE shared-memory counter by
//              the group size with a plain read-modify-write, the others take their rank from the ballot.  Per tile
//              of 4096 (4-byte keys) / 2048 rows: per-warp counts -> exclusive scan over (partition, warp) -> one
//              global reservation per partition -> rows staged in shared memory grouped by partition -> streamed out,
//              consecutive threads writing consecutive addresses of a run.
//   Measured (tools/partition_sweep.py, 268 M rows of 4+4 bytes): histogram 0.22 ms at 8..128 partitions (4.9 TB/s),
//   0.66 ms at 512; scatter 1.26 ms at 16 partitions (3.4 TB/s), 1.50 ms at 128, 2.0 ms at 512; 8+8-byte rows reach
//   5.0 TB/s at 16 partitions.  The scatter replaced a shared-memory-atomic version (1.65 ms at 128, 2.7 ms at 512)
//   and, for <= 8 partitions, a register-only ballot kernel that stored rows unstaged (1.26-1.8 ms).
#pragma once
#include "table.cuh"

namespace dwj {

constexpr int PART_MAX = 512;
constexpr int PART_THREADS = 256;

template <int W> struct PartitionArgs {
  using K = typename KeyT<W>::type;
  const K *keys;
  const K *vals;       // may be null
  uint64_t n;
  uint32_t log2_parts;
  uint64_t seed;
  uint32_t mode;         // PartMode, see part_id
  uint32_t rank_bits;    // PART_BY_BOTH: bits of the destination rank (log2_parts = rank_bits + region bits)
  uint32_t bucket_shift;
  uint64_t bucket_mask;
  K *out_keys;
  K *out_vals;         // may be null
  unsigned long long *hist;     // [PART_MAX] zeroed before the histogram kernel
  unsigned long long *cursor;   // [PART_MAX] running write positions
  unsigned long long *offsets;  // [parts + 1] result
  // Segmented input (table.cuh; the many-way scatter only): n_segs > 0 => the rows come from segs[] -- possibly peer
  // memory, which makes the scatter the receiving end of the multi-GPU exchange -- and `n` is the number of TILES.
  const Seg *segs;
  uint32_t n_segs;
  PassFilter filter;            // rows of other key classes are dropped (histograms do not count them)
};

// Partition id of a key.  `mode` is uniform across the grid, so the branch costs one predicate.
//   PART_BY_HASH   : top bits of the independent partition hash (multi-GPU exchange: destination rank)
//   PART_BY_BUCKET : top bits of the bucket index (engine regions: the rows of a partition touch one table slice)
//   PART_BY_BOTH   : destination rank (rank_bits) in the high bits, table region of the DESTINATION's table in the low
//                    bits -- one pass that serves the exchange and the receiver's region grouping (dwj_xpart_*)
enum PartMode : uint32_t { PART_BY_HASH = 0, PART_BY_BUCKET = 1, PART_BY_BOTH = 2 };
template <int W, uint32_t MODE> DWJ_D uint32_t part_id_of(const PartitionArgs<W> &a, typename KeyT<W>::type key) {
  if constexpr (MODE == PART_BY_BUCKET) return (uint32_t)((slot_hash(key, a.seed) & a.bucket_mask) >> a.bucket_shift);
  if constexpr (MODE == PART_BY_HASH) return partition_of(key, a.log2_parts, a.seed);
  const uint32_t region_bits = a.log2_parts - a.rank_bits;
  return partition_of(key, a.rank_bits, a.seed) << region_bits | (uint32_t)((slot_hash(key, a.seed) & a.bucket_mask) >> a.bucket_shift);
}
// Run-time mode (the scatter: one id per row beside ~100 other instructions); the histograms, which do little else,
// are compiled per mode.
template <int W> DWJ_D uint32_t part_id(const PartitionArgs<W> &a, typename KeyT<W>::type key) {
  if (a.mode == PART_BY_BUCKET) return part_id_of<W, PART_BY_BUCKET>(a, key);
  if (a.mode == PART_BY_HASH) return part_id_of<W, PART_BY_HASH>(a, key);
  return part_id_of<W, PART_BY_BOTH>(a, key);
}

constexpr uint32_t PART_DEAD = 0xFFFFFFFFu;   // partition id of a lane past the end of the input

__global__ void stage_value_kernel(unsigned long long *dst, unsigned long long v) { *dst = v; }

template <int W> __global__ void partition_offsets_kernel(PartitionArgs<W> a) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const uint32_t parts = 1u << a.log2_parts;
    unsigned long long run = 0;
    for (uint32_t p = 0; p < parts; ++p) {
      a.offsets[p] = run;
      a.cursor[p] = run;
      run += a.hist[p];
    }
    a.offsets[parts] = run;
  }
}

// ---- scatter (2 .. 512 partitions) and the many-way histogram ----------------------------------------------------
template <int BITS> DWJ_D unsigned match_partition(uint32_t p, unsigned alive) {
  unsigned peers = alive;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const bool bit = (p >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

template <int W, uint32_t MODE, int THREADS, int HROWS>
__global__ void __launch_bounds__(THREADS) partition_hist_private_kernel(PartitionArgs<W> a) {
  using K = typename KeyT<W>::type;
  extern __shared__ __align__(16) unsigned char s_cnt[];      // [parts][THREADS] bytes
  const uint32_t parts = 1u << a.log2_parts;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // byte of (this thread, partition 0): lane l -> word l of a 128-byte group (bank l for every partition), the four
  // warps of a group take the four bytes of the word
  unsigned char *mine = s_cnt + (warp >> 2) * 128 + lane * 4 + (warp & 3);
  constexpr int WORDS = THREADS / 4;                           // 32-bit words per partition row
  constexpr int PER = PART_MAX / THREADS;                      // partitions folded by one thread
  uint32_t *words = reinterpret_cast<uint32_t *>(s_cnt);
  for (uint32_t i = threadIdx.x; i < parts * WORDS; i += THREADS) words[i] = 0;
  __syncthreads();
  unsigned long long acc[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) acc[q] = 0;
  auto fold = [&]() {      // all threads: sum and clear the THREADS byte counters of `my` partitions
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const uint32_t p = threadIdx.x + q * THREADS;
      if (p < parts) {
        uint32_t sum = 0;
#pragma unroll 8
        for (int k = 0; k < WORDS; ++k) {
          const uint32_t idx = p * WORDS + ((k + lane) & (WORDS - 1));   // rotated: the lanes of a warp hit 32 banks
          sum = __dp4a(words[idx], 0x01010101u, sum);
          words[idx] = 0;
        }
        acc[q] += sum;
      }
    }
    __syncthreads();
  };
  constexpr uint64_t TILE = (uint64_t)THREADS * HROWS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  uint32_t pending = 0;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE + threadIdx.x;
    K k[HROWS];
    if (tile * TILE + TILE <= a.n) {
#pragma unroll
      for (int j = 0; j < HROWS; ++j) k[j] = load_stream(a.keys + base + (uint64_t)j * THREADS);
#pragma unroll
      for (int j = 0; j < HROWS; ++j) {
        unsigned char *c = mine + part_id_of<W, MODE>(a, k[j]) * THREADS;
        *c = (unsigned char)(*c + 1);
      }
    } else {
#pragma unroll
      for (int j = 0; j < HROWS; ++j) {
        const uint64_t i = base + (uint64_t)j * THREADS;
        if (i < a.n) {
          unsigned char *c = mine + part_id_of<W, MODE>(a, load_stream(a.keys + i)) * THREADS;
          *c = (unsigned char)(*c + 1);
        }
      }
    }
    pending += HROWS;
    if (pending > 255 - HROWS) { fold(); pending = 0; }       // block-uniform: every thread has seen the same tiles
  }
  fold();
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const uint32_t p = threadIdx.x + q * THREADS;
    if (p < parts && acc[q]) atomicAdd(a.hist + p, acc[q]);
  }
}

template <int W> struct PairT;
template <> struct PairT<4> {
  using type = uint2;
  static DWJ_D uint2 make(uint32_t k, uint32_t v) { return make_uint2(k, v); }
  static DWJ_D uint32_t key(uint2 r) { return r.x; }
  static DWJ_D uint32_t val(uint2 r) { return r.y; }
};
template <> struct PairT<8> {
  using type = ulonglong2;
  static DWJ_D ulonglong2 make(uint64_t k, uint64_t v) { return make_ulonglong2(k, v); }
  static DWJ_D uint64_t key(ulonglong2 r) { return r.x; }
  static DWJ_D uint64_t val(ulonglong2 r) { return r.y; }
};

template <int W, int THREADS, int ITEMS> struct ScatterManySmem {
  using K = typename KeyT<W>::type;
  static constexpr uint32_t TILE = THREADS * ITEMS;
  static constexpr int WARPS = THREADS / 32;
  // dynamic shared memory: rows[TILE] (pairs, or keys | payloads) | delta[parts] (int64) | wc[WARPS][parts] (uint32) | part[TILE] (uint16, 8-byte rows)
  static size_t bytes(uint32_t parts) { return (size_t)TILE * (2 * sizeof(K) + (W == 4 ? 2 : 0)) + (size_t)parts * (8 + 4 * WARPS); }
};

template <int W, int BITS, int THREADS, int ITEMS, bool FULL, bool MATCH>
DWJ_D void scatter_many_tile(const PartitionArgs<W> &a, const typename KeyT<W>::type *kp, const typename KeyT<W>::type *vp, uint32_t rows,
                             unsigned char *smem, unsigned int *s_scan) {
  using K = typename KeyT<W>::type;
  constexpr uint32_t TILE = THREADS * ITEMS;
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t PARTS = 1u << BITS;
  constexpr int PER = (PARTS + THREADS - 1) / THREADS;
  // 16-byte rows are staged as (key, payload) PAIRS -- one shared-memory store and one load per row instead of two of
  // each -- and the partition of a staged row is recomputed from its key when the row leaves (a dozen ALU instructions
  // against a store and a load): that kernel is bound by the shared-memory pipe (r2 ncu: mio + short-scoreboard 42 % of
  // the stall samples, issue slots 41 % busy; +3..8 %).  8-byte rows keep separate columns and a staged partition id:
  // their kernel is bound by issue slots, and the re-hash cost 20 % (profiles/r2_partition_sweep.txt).
  constexpr bool PAIRS = W == 8;
  using Pair = typename PairT<W>::type;
  Pair *s_rows = reinterpret_cast<Pair *>(smem);
  K *s_keys = reinterpret_cast<K *>(smem);
  K *s_vals = s_keys + TILE;
  long long *s_delta = reinterpret_cast<long long *>(smem + (size_t)TILE * 2 * sizeof(K));
  unsigned int *s_wc = reinterpret_cast<unsigned int *>(s_delta + PARTS);
  unsigned short *s_part = reinterpret_cast<unsigned short *>(s_wc + WARPS * PARTS);   // 8-byte rows: partition of every staged row
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const bool with_vals = a.vals != nullptr;
  unsigned int *mywc = s_wc + warp * PARTS;
#pragma unroll
  for (uint32_t p = lane; p < PARTS; p += 32) mywc[p] = 0;

  K k[ITEMS], v[ITEMS];
  uint32_t pr[ITEMS];                               // partition << 16 | rank inside (warp, partition); PART_DEAD = dead row
  uint32_t livemask = 0;                            // FULL = false: rows inside the relation AND of the current key class
  kp += threadIdx.x;
  vp += threadIdx.x;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool in = FULL || j * THREADS + threadIdx.x < rows;
    k[j] = in ? load_stream(kp + j * THREADS) : (K)0;
    const bool live = FULL || (in && pass_ok(k[j], a.seed, a.filter));
    livemask |= (live ? 1u : 0u) << j;
    v[j] = live && with_vals ? load_stream(vp + j * THREADS) : (K)0;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || (livemask >> j & 1u);
    const uint32_t p = part_id<W>(a, k[j]);
    const unsigned alive = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, live);
    // lanes of the warp bound for the same partition: one MATCH instruction or one ballot per partition-id bit
    const unsigned peers = MATCH ? (__match_any_sync(0xffffffffu, p) & alive) : match_partition<BITS>(p, alive);
    uint32_t old = 0;
    if (live && (peers & lt) == 0) {                // lowest lane of the group: plain RMW on the warp's own counter
      old = mywc[p];
      mywc[p] = old + __popc(peers);
    }
    __syncwarp();
    old = __shfl_sync(0xffffffffu, old, (__ffs(peers) - 1) & 31);
    pr[j] = live ? (p << 16 | (old + __popc(peers & lt))) : PART_DEAD;
  }
  __syncthreads();
  // (partition, warp) exclusive scan + one global reservation per partition.  Thread t owns partitions t*PER .. +PER.
  {
    unsigned local[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const uint32_t p = threadIdx.x * PER + q;
      unsigned total = 0;
      if (p < PARTS) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) total += s_wc[w * PARTS + p];
      }
      local[q] = total;
      sum += total;
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += n;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned before = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) before += w < (int)warp ? s_scan[w] : 0u;
    if (!FULL && threadIdx.x == THREADS - 1) s_scan[WARPS] = before + incl;      // rows of the tile that are staged at all
    unsigned run = before + incl - sum;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const uint32_t p = threadIdx.x * PER + q;
      if (p < PARTS) {
        const unsigned long long g = local[q] ? atomicAdd(a.cursor + p, (unsigned long long)local[q]) : 0ull;
        s_delta[p] = (long long)g - (long long)run;
        unsigned at = run;                          // warp w's rows of partition p are staged at s_wc[w][p] + rank
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
          const unsigned c = s_wc[w * PARTS + p];
          s_wc[w * PARTS + p] = at;
          at += c;
        }
      }
      run += local[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (FULL || pr[j] != PART_DEAD) {
      const uint32_t s = mywc[pr[j] >> 16] + (pr[j] & 0xFFFFu);
      if constexpr (PAIRS) {
        s_rows[s] = PairT<W>::make(k[j], v[j]);
      } else {
        s_keys[s] = k[j];
        s_vals[s] = v[j];
        s_part[s] = (unsigned short)(pr[j] >> 16);
      }
    }
  }
  __syncthreads();
  const uint32_t staged_rows = FULL ? (uint32_t)TILE : s_scan[WARPS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t s = j * THREADS + threadIdx.x;
    if (FULL || s < staged_rows) {
      if constexpr (PAIRS) {
        const Pair row = s_rows[s];
        const long long dst = (long long)s + s_delta[part_id<W>(a, PairT<W>::key(row))];
        store_stream(a.out_keys + dst, PairT<W>::key(row));
        if (with_vals) store_stream(a.out_vals + dst, PairT<W>::val(row));
      } else {
        const long long dst = (long long)s + s_delta[s_part[s]];
        store_stream(a.out_keys + dst, s_keys[s]);
        if (with_vals) store_stream(a.out_vals + dst, s_vals[s]);
      }
    }
  }
}

template <int W, int BITS, int THREADS, int ITEMS, int MINB, bool MATCH>
__global__ void __launch_bounds__(THREADS, MINB) partition_scatter_many_kernel(PartitionArgs<W> a) {
  constexpr uint32_t TILE = THREADS * ITEMS;
  using K = typename KeyT<W>::type;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ unsigned int s_scan[THREADS / 32 + 1];
  const uint64_t num_tiles = a.n_segs ? a.n : (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const K *kp = a.keys, *vp = a.vals;
    uint64_t base = tile * TILE, limit = a.n;
    if (a.n_segs) {                                 // CTA-uniform: the tile's rows live in one segment (possibly peer memory)
      const Seg sg = a.segs[find_segment(a.segs, a.n_segs, tile)];
      kp = (const K *)sg.keys;
      vp = (const K *)sg.vals;
      base = (tile - sg.first_unit) * TILE;
      limit = sg.rows;
    }
    const uint32_t rows = (uint32_t)min((uint64_t)TILE, limit - base);
    if (rows == TILE && !a.filter.mask) scatter_many_tile<W, BITS, THREADS, ITEMS, true, MATCH>(a, kp + base, vp + base, rows, s_dyn, s_scan);
    else scatter_many_tile<W, BITS, THREADS, ITEMS, false, MATCH>(a, kp + base, vp + base, rows, s_dyn, s_scan);
    __syncthreads();                                // shared memory is reused by the next tile
  }
}

// ---- histogram with shared-memory reductions (any partition count up to 4096, pass filter) -----------------------------
// One 32-bit counter per partition in shared memory, bumped with RED.shared (no result: the cost that rules ATOMS out of
// the scatter does not apply).  Used where the private-counter kernels do not reach: the (destination rank x table
// region) histogram of the multi-GPU exchange when ranks x regions exceeds 512 -- the senders count for the receivers,
// so that a receiver can lay out its region-grouped buffer before it pulls a single row -- and every histogram under a
// pass filter.
template <int W, uint32_t MODE, int HROWS>
__global__ void __launch_bounds__(PART_THREADS) partition_hist_red_kernel(PartitionArgs<W> a) {
  using K = typename KeyT<W>::type;
  extern __shared__ unsigned int s_bins[];
  const uint32_t parts = 1u << a.log2_parts;
  for (uint32_t i = threadIdx.x; i < parts; i += PART_THREADS) s_bins[i] = 0;
  __syncthreads();
  constexpr uint64_t TILE = (uint64_t)PART_THREADS * HROWS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE + threadIdx.x;
    K k[HROWS];
#pragma unroll
    for (int j = 0; j < HROWS; ++j) {
      const uint64_t i = base + (uint64_t)j * PART_THREADS;
      k[j] = i < a.n ? load_stream(a.keys + i) : (K)0;
    }
#pragma unroll
    for (int j = 0; j < HROWS; ++j) {
      const uint64_t i = base + (uint64_t)j * PART_THREADS;
      if (i < a.n && pass_ok(k[j], a.seed, a.filter)) atomicAdd(&s_bins[part_id_of<W, MODE>(a, k[j])], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < parts; i += PART_THREADS)
    if (s_bins[i]) atomicAdd(a.hist + i, (unsigned long long)s_bins[i]);
}

// ---- class compaction (dwj_filter_rows) -------------------------------------------------------------------------------------
// Keeps the rows of the pass filter's key class: ballot + popc give every kept row its place inside the warp's output,
// a shared-memory prefix over the warps and ONE global reservation per tile of THREADS * ITEMS rows place the warps, and
// every store instruction writes one contiguous run.  Output order is arbitrary.  No ranking, no staging: this is a
// stream compaction, not a partition (the many-way kernel with one partition ran it at 3.2 TB/s).
// With a.hist set the kernel also counts the kept rows per table region (partition id of mode a.mode, 2^log2_parts <= 512
// bins, RED.shared) -- the histogram the region scatter of the compact copy needs, without another pass over it.
template <int W, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) filter_rows_kernel(PartitionArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr int WARPS = THREADS / 32;
  constexpr uint64_t TILE = (uint64_t)THREADS * ITEMS;
  __shared__ unsigned int s_cnt[WARPS];
  __shared__ unsigned long long s_base;
  __shared__ unsigned int s_bins[PART_MAX];
  const uint32_t bins = a.hist ? 1u << a.log2_parts : 0u;
  for (uint32_t i = threadIdx.x; i < bins; i += THREADS) s_bins[i] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
  const bool with_vals = a.vals != nullptr;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE + (uint64_t)warp * (32 * ITEMS) + lane;       // a warp owns 32 * ITEMS consecutive rows
    K k[ITEMS], v[ITEMS];
    unsigned mask[ITEMS];
    unsigned total = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint64_t i = base + (uint64_t)j * 32;
      const bool in = i < a.n;
      k[j] = in ? load_stream(a.keys + i) : (K)0;
      v[j] = in && with_vals ? load_stream(a.vals + i) : (K)0;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const bool live = base + (uint64_t)j * 32 < a.n && pass_ok(k[j], a.seed, a.filter);
      mask[j] = __ballot_sync(0xffffffffu, live);
      total += __popc(mask[j]);
    }
    if (lane == 0) s_cnt[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned sum = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) sum += s_cnt[w];
      s_base = sum ? atomicAdd(a.cursor, (unsigned long long)sum) : 0ull;
    }
    __syncthreads();
    unsigned long long at = s_base;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) at += w < (int)warp ? s_cnt[w] : 0u;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (mask[j] >> lane & 1u) {
        const unsigned long long pos = at + __popc(mask[j] & lt);
        store_stream(a.out_keys + pos, k[j]);
        if (with_vals) store_stream(a.out_vals + pos, v[j]);
        if (bins) atomicAdd(&s_bins[part_id<W>(a, k[j])], 1u);
      }
      at += __popc(mask[j]);
    }
    __syncthreads();                                // s_cnt / s_base are reused by the next tile
  }
  for (uint32_t i = threadIdx.x; i < bins; i += THREADS)
    if (s_bins[i]) atomicAdd(a.hist + i, (unsigned long long)s_bins[i]);
}

// ---- histogram for at most 8 partitions ------------------------------------------------------------------------
// Per-partition counters are PACKED into one 64-bit register per thread: 8 fields x 8 bits, flushed to shared memory
// before a field can overflow.  Interior tiles run a FULL = true instantiation of the tile body without any bounds
// predicate (the 64-bit compares and the per-row branches they cause were most of the instruction stream at first).
template <int W, uint32_t MODE, int HROWS, bool FULL>
DWJ_D void hist8_tile(const PartitionArgs<W> &a, uint64_t tile_base, unsigned long long &acc) {
  using K = typename KeyT<W>::type;
  const K *kp = a.keys + tile_base + threadIdx.x;
  const uint32_t rows = FULL ? 0u : (uint32_t)(a.n - tile_base);
  K k[HROWS];
#pragma unroll
  for (int j = 0; j < HROWS; ++j) k[j] = (FULL || j * PART_THREADS + threadIdx.x < rows) ? load_stream(kp + j * PART_THREADS) : (K)0;
#pragma unroll
  for (int j = 0; j < HROWS; ++j) {
    const unsigned long long one = 1ull << (8 * part_id_of<W, MODE>(a, k[j]));
    acc += (FULL || j * PART_THREADS + threadIdx.x < rows) ? one : 0ull;
  }
}

template <int W, uint32_t MODE, int HROWS>
__global__ void __launch_bounds__(PART_THREADS) partition_hist8_kernel(PartitionArgs<W> a) {
  __shared__ unsigned int s_hist[8];
  if (threadIdx.x < 8) s_hist[threadIdx.x] = 0;
  __syncthreads();
  constexpr uint64_t TILE = (uint64_t)PART_THREADS * HROWS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  unsigned long long acc = 0;      // 8 x 8-bit counters
  uint32_t pending = 0;
  auto flush = [&]() {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const unsigned c = (unsigned)(acc >> (8 * p)) & 0xFFu;
      if (c) atomicAdd(&s_hist[p], c);
    }
    acc = 0;
    pending = 0;
  };
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (tile * TILE + TILE <= a.n) hist8_tile<W, MODE, HROWS, true>(a, tile * TILE, acc);
    else hist8_tile<W, MODE, HROWS, false>(a, tile * TILE, acc);
    pending += HROWS;
    if (pending > 255 - HROWS) flush();
  }
  flush();
  __syncthreads();
  if (threadIdx.x < 8 && s_hist[threadIdx.x]) atomicAdd(a.hist + threadIdx.x, (unsigned long long)s_hist[threadIdx.x]);
}

}  // namespace dwj
