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
  // Per-partition destinations (<= 8 partitions, dwj_partition_scatter_to): partition p is written to
  // dst_keys[p] / dst_vals[p] -- which may be PEER GPU memory mapped over NVLink -- instead of out_keys/out_vals.
  uint32_t use_dst;
  K *dst_keys[8];
  K *dst_vals[8];
  unsigned long long *hist;     // [PART_MAX] zeroed before the histogram kernel
  unsigned long long *cursor;   // [PART_MAX] running write positions
  unsigned long long *offsets;  // [parts + 1] result
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

template <int W, int THREADS, int ITEMS> struct ScatterManySmem {
  using K = typename KeyT<W>::type;
  static constexpr uint32_t TILE = THREADS * ITEMS;
  static constexpr int WARPS = THREADS / 32;
  // dynamic shared memory: keys[TILE] | vals[TILE] | delta[parts] (int64) | wc[WARPS][parts] (uint32) | part[TILE] (uint16)
  static size_t bytes(uint32_t parts) { return (size_t)TILE * (2 * sizeof(K) + 2) + (size_t)parts * (8 + 4 * WARPS); }
};

template <int W, int BITS, int THREADS, int ITEMS, bool FULL>
DWJ_D void scatter_many_tile(const PartitionArgs<W> &a, uint64_t base, uint32_t rows, unsigned char *smem, unsigned int *s_scan) {
  using K = typename KeyT<W>::type;
  constexpr uint32_t TILE = THREADS * ITEMS;
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t PARTS = 1u << BITS;
  constexpr int PER = (PARTS + THREADS - 1) / THREADS;
  K *s_keys = reinterpret_cast<K *>(smem);
  K *s_vals = s_keys + TILE;
  long long *s_delta = reinterpret_cast<long long *>(s_vals + TILE);
  unsigned int *s_wc = reinterpret_cast<unsigned int *>(s_delta + PARTS);
  unsigned short *s_part = reinterpret_cast<unsigned short *>(s_wc + WARPS * PARTS);   // partition of every staged row: two
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                     // shared-memory ops instead of a re-hash
  const unsigned lt = (1u << lane) - 1u;
  const bool with_vals = a.vals != nullptr;
  unsigned int *mywc = s_wc + warp * PARTS;
#pragma unroll
  for (uint32_t p = lane; p < PARTS; p += 32) mywc[p] = 0;

  K k[ITEMS], v[ITEMS];
  uint32_t pr[ITEMS];                               // partition << 16 | rank inside (warp, partition); PART_DEAD = dead row
  const K *kp = a.keys + base + threadIdx.x, *vp = a.vals + base + threadIdx.x;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * THREADS + threadIdx.x < rows;
    k[j] = live ? load_stream(kp + j * THREADS) : (K)0;
    v[j] = live && with_vals ? load_stream(vp + j * THREADS) : (K)0;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * THREADS + threadIdx.x < rows;
    const uint32_t p = part_id<W>(a, k[j]);
    const unsigned alive = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, live);
    const unsigned peers = match_partition<BITS>(p, alive);
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
      s_keys[s] = k[j];
      s_vals[s] = v[j];
      s_part[s] = (unsigned short)(pr[j] >> 16);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t s = j * THREADS + threadIdx.x;
    if (FULL || s < rows) {
      const long long dst = (long long)s + s_delta[s_part[s]];
      store_stream(a.out_keys + dst, s_keys[s]);
      if (with_vals) store_stream(a.out_vals + dst, s_vals[s]);
    }
  }
}

template <int W, int BITS, int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) partition_scatter_many_kernel(PartitionArgs<W> a) {
  constexpr uint32_t TILE = THREADS * ITEMS;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ unsigned int s_scan[THREADS / 32];
  const uint64_t num_tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE;
    const uint32_t rows = (uint32_t)min((uint64_t)TILE, a.n - base);
    if (rows == TILE) scatter_many_tile<W, BITS, THREADS, ITEMS, true>(a, base, rows, s_dyn, s_scan);
    else scatter_many_tile<W, BITS, THREADS, ITEMS, false>(a, base, rows, s_dyn, s_scan);
    __syncthreads();                                // shared memory is reused by the next tile
  }
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

// Scatter for <= 8 partitions with one DESTINATION POINTER per partition, for destinations on the far side of NVLink
// (dwj_partition_scatter_to): three ballots per 32 rows and warp-distributed counters (lane q keeps the warp's running
// count of partition q) rank the rows; the tile is first grouped by partition in shared memory and then
// streamed out so that every warp-level store is one contiguous 128-byte (4-byte keys) / 256-byte (8-byte keys)
// piece of ONE destination.  Storing straight from registers hands each peer ~16-byte pieces per instruction, which
// NVLink moves at a fraction of its bandwidth (8 GPUs: 26.8 ms per step against a 2.6 ms transfer floor).
template <int W, int ITEMS> struct Scatter8StagedSmem {
  using K = typename KeyT<W>::type;
  K keys[PART_THREADS * ITEMS];
  K vals[PART_THREADS * ITEMS];
  unsigned char part[PART_THREADS * ITEMS];
  unsigned int wcnt[PART_THREADS / 32][8];    // per-warp rows per partition, then per-warp prefix
  unsigned int start[8];                      // staging offset of each partition inside the tile
  long long delta[8];                         // global row of the partition's run minus its staging offset
  K *dstk[8];
  K *dstv[8];
};

template <int W, int ITEMS, bool FULL, bool WITH_VALS>
DWJ_D void scatter8_staged_tile(const PartitionArgs<W> &a, uint64_t base, uint32_t rows, Scatter8StagedSmem<W, ITEMS> &sm) {
  using K = typename KeyT<W>::type;
  constexpr int WARPS = PART_THREADS / 32;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned q0 = (lane & 1) ? 0xffffffffu : 0u, q1 = (lane & 2) ? 0xffffffffu : 0u, q2 = (lane & 4) ? 0xffffffffu : 0u;
  K k[ITEMS], v[ITEMS];
  uint32_t pr[ITEMS];                              // rank inside the warp << 4 | partition (8 = dead row)
  const K *kp = a.keys + base + threadIdx.x, *vp = a.vals + base + threadIdx.x;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * PART_THREADS + threadIdx.x < rows;
    k[j] = live ? load_stream(kp + j * PART_THREADS) : (K)0;
    if constexpr (WITH_VALS) v[j] = live ? load_stream(vp + j * PART_THREADS) : (K)0;
  }
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool live = FULL || j * PART_THREADS + threadIdx.x < rows;
    const uint32_t p = part_id<W>(a, k[j]);
    const unsigned alive = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, live);
    const unsigned b0 = __ballot_sync(0xffffffffu, p & 1u), b1 = __ballot_sync(0xffffffffu, p & 2u), b2 = __ballot_sync(0xffffffffu, p & 4u);
    const unsigned m0 = (p & 1u) ? b0 : ~b0, m1 = (p & 2u) ? b1 : ~b1, m2 = (p & 4u) ? b2 : ~b2;
    const unsigned peers = m0 & m1 & m2 & alive;
    const uint32_t before = __shfl_sync(0xffffffffu, run, p);
    pr[j] = live ? ((before + __popc(peers & lt)) << 4 | p) : 8u;
    run += __popc(~(b0 ^ q0) & ~(b1 ^ q1) & ~(b2 ^ q2) & alive);
  }
  if (lane < 8) sm.wcnt[warp][lane] = run;
  __syncthreads();
  if (threadIdx.x < 8) {                           // partition q: prefix over warps, staging offset, global reservation
    const unsigned q = threadIdx.x;
    unsigned total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { const unsigned c = sm.wcnt[w][q]; sm.wcnt[w][q] = total; total += c; }
    unsigned start = 0;
    // exclusive scan over the 8 partitions through shuffles inside this (partial) warp
    unsigned incl = total;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned n = __shfl_up_sync(0xffu, incl, o);
      if (q >= (unsigned)o) incl += n;
    }
    start = incl - total;
    sm.start[q] = start;
    const unsigned long long g = total ? atomicAdd(a.cursor + q, (unsigned long long)total) : 0ull;
    sm.delta[q] = (long long)g - (long long)start;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (FULL || pr[j] != 8u) {
      const uint32_t p = pr[j] & 15u;
      const uint32_t s = sm.start[p] + sm.wcnt[warp][p] + (pr[j] >> 4);
      sm.keys[s] = k[j];
      if constexpr (WITH_VALS) sm.vals[s] = v[j];
      sm.part[s] = (unsigned char)p;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t s = j * PART_THREADS + threadIdx.x;
    if (FULL || s < rows) {
      const uint32_t p = sm.part[s];
      const long long dst = (long long)s + sm.delta[p];
      store_stream(sm.dstk[p] + dst, sm.keys[s]);
      if constexpr (WITH_VALS) store_stream(sm.dstv[p] + dst, sm.vals[s]);
    }
  }
  __syncthreads();
}

template <int W, int ITEMS>
__global__ void __launch_bounds__(PART_THREADS, 3) partition_scatter8_staged_kernel(PartitionArgs<W> a) {
  constexpr uint32_t TILE = PART_THREADS * ITEMS;
  __shared__ Scatter8StagedSmem<W, ITEMS> sm;
  if (threadIdx.x < 8) {
    sm.dstk[threadIdx.x] = a.use_dst ? a.dst_keys[threadIdx.x] : a.out_keys;
    sm.dstv[threadIdx.x] = a.use_dst ? a.dst_vals[threadIdx.x] : a.out_vals;
  }
  __syncthreads();
  const bool with_vals = a.vals != nullptr;
  const uint64_t num_tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE;
    const uint32_t rows = (uint32_t)min((uint64_t)TILE, a.n - base);
    if (rows == TILE) {
      if (with_vals) scatter8_staged_tile<W, ITEMS, true, true>(a, base, rows, sm);
      else scatter8_staged_tile<W, ITEMS, true, false>(a, base, rows, sm);
    } else {
      if (with_vals) scatter8_staged_tile<W, ITEMS, false, true>(a, base, rows, sm);
      else scatter8_staged_tile<W, ITEMS, false, false>(a, base, rows, sm);
    }
  }
}

// ---- exchange: push runs into peer memory with a FEW CTAs -------------------------------------------------------------
// The folded exchange (dwj_xpart_*) leaves ranks x regions contiguous runs per relation, each bound for its own place
// in some peer's receive area.  On this machine the copy engines serialise such a list at ~27 us per copy whatever
// its size (112 copies of 8 MB: 3 ms, 310 GB/s), so the runs are pushed by a small kernel instead: a few dozen CTAs
// stream all runs through registers with coalesced 128-byte (4-byte rows) / 256-byte warp stores -- the piece size
// NVLink wants -- while the rest of the SMs go on partitioning and probing.  Work unit: one block of 256 x ELEMS
// rows of one run; blocks are dealt round-robin to the CTAs.
struct PushRun {
  void *dst;
  const void *src;
  unsigned long long rows;
  unsigned long long first_block;      // blocks of all earlier runs
};

template <int W, int ELEMS>
__global__ void __launch_bounds__(256) push_runs_kernel(const PushRun *runs, uint32_t n_runs, unsigned long long total_blocks) {
  using K = typename KeyT<W>::type;
  constexpr unsigned long long BLOCK_ROWS = 256ull * ELEMS;
  for (unsigned long long blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
    uint32_t lo = 0, hi = n_runs;                   // runs[lo].first_block <= blk < runs[hi].first_block
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(&runs[mid].first_block) <= blk) lo = mid; else hi = mid;
    }
    const PushRun r = runs[lo];
    const unsigned long long row0 = (blk - r.first_block) * BLOCK_ROWS + threadIdx.x;
    const K *src = (const K *)r.src + row0;
    K *dst = (K *)r.dst + row0;
    K v[ELEMS];
    if (row0 - threadIdx.x + BLOCK_ROWS <= r.rows) {
#pragma unroll
      for (int j = 0; j < ELEMS; ++j) v[j] = load_stream(src + j * 256);
#pragma unroll
      for (int j = 0; j < ELEMS; ++j) dst[j * 256] = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < ELEMS; ++j) if (row0 + j * 256 < r.rows) v[j] = load_stream(src + j * 256);
#pragma unroll
      for (int j = 0; j < ELEMS; ++j) if (row0 + j * 256 < r.rows) dst[j * 256] = v[j];
    }
  }
}

}  // namespace dwj
