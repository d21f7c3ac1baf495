// dwj_api.cu -- C ABI (include/dwj.h) over the sm_100a join kernels.  No CPU fallback.
#include "../../include/dwj.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "aggregate.cuh"
#include "build.cuh"
#include "csr.cuh"
#include "partition.cuh"
#include "probe.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(_e == cudaErrorMemoryAllocation ? DWJ_ERR_OOM : DWJ_ERR_CUDA, "%s: %s (%s:%d)",  \
                  #call, cudaGetErrorString(_e), __FILE__, __LINE__);                              \
  } while (0)

struct DeviceGuard {   // callers (torch, other engines) may have another device current
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

uint64_t next_pow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Staged PAIRS kernel (unique build keys): warps per CTA, rows per thread per round, rounds, min CTAs per SM.
// DWJ_STAGED_SHAPE=0..4 picks one at run time for tuning sweeps (tools/probe_sweep.py); the default is what
// profiles/ shows to be fastest.
struct StagedShape { int warps, items, sub, minb; };
constexpr StagedShape STAGED_SHAPES_4[] = {{8, 4, 4, 4}, {8, 4, 8, 2}, {8, 2, 8, 4}, {4, 4, 4, 8}, {8, 8, 2, 2}};
constexpr StagedShape STAGED_SHAPES_8[] = {{8, 2, 4, 4}, {8, 2, 8, 2}, {8, 4, 2, 4}, {4, 2, 4, 8}, {8, 4, 4, 2}};
constexpr int DEFAULT_STAGED_SHAPE_4 = 2, DEFAULT_STAGED_SHAPE_8 = 3;   // gpurun sweep, profiles/r1_probe.md
constexpr bool DEFAULT_SCATTER_MATCH = false;
// log2(partitions) from which the shared-memory RED histogram replaces the private-counter ones (profiles/r2_partition_sweep.txt:
// 8-byte keys 0.35 ms per 2^28 rows at every partition count against 0.35 .. 0.79; 4-byte keys 0.29 against 0.22 / 0.24 / 0.35 / 0.65
// at 32 / 128 / 256 / 512 partitions)
constexpr int DEFAULT_HIST_RED_FROM_4 = 8, DEFAULT_HIST_RED_FROM_8 = 4;
constexpr int SIMPLE_ITEMS_4 = 4, SIMPLE_ITEMS_8 = 4;   // rows per thread per round, warp-centric kernel
constexpr uint64_t HOST_CHUNK_BYTES = 32ull << 20;   // per column per pipeline stage in dwj_join_host
constexpr int HOST_STAGES = 4;                       // staging slots of the dwj_join_host pipeline

}  // namespace

struct dwj_engine {
  dwj_config cfg{};
  int W = 4;
  cudaDeviceProp prop{};
  void *table = nullptr;
  unsigned int *fill = nullptr;                // per-bucket ticket counters of the build (build.cuh)
  uint64_t slots = 0, buckets = 0, table_bytes = 0;
  // one-to-many engines (no DWJ_FLAG_UNIQUE_BUILD_KEYS): the payload runs of csr.cuh; `csr` says the last build made them
  void *runs = nullptr;
  uint64_t runs_granules = 0;
  bool csr = false;
  uint64_t build_rows = 0;
  bool built = false;
  bool l2_window = false;
  cudaAccessPolicyWindow window{};
  // events
  cudaEvent_t ev_build[2]{}, ev_probe[2]{}, ev_part[2]{}, ev_probek[2]{}, ev_buildk[2]{};
  bool have_build = false, have_probe = false, have_part = false;
  // scan / counters scratch
  unsigned long long *tile_state = nullptr;
  uint64_t tile_state_cap = 0;                 // in descriptors
  unsigned long long *counter = nullptr;       // device uint64 used when the caller passes no d_n_matches
  unsigned long long *part_scratch = nullptr;  // hist[PART_MAX] + cursor[PART_MAX] + region offsets[PART_MAX + 1]
  // segmented input (dwj_*_segments): ring of segment tables, and the caller's list while a call is in flight
  dwj::Seg *seg_tables = nullptr, *seg_tables_host = nullptr;
  cudaEvent_t seg_done[32]{};
  uint32_t seg_calls = 0;
  const void *const *pending_seg_keys = nullptr, *const *pending_seg_vals = nullptr;
  const uint64_t *pending_seg_rows = nullptr;
  uint32_t pending_segs = 0, pending_segs_per_region = 0;
  // Small host arrays (scatter start rows) travel through a ring of PINNED staging slots: an asynchronous copy from
  // pageable memory synchronises the stream first, and a stream of the multi-GPU join may be parked behind a kernel that
  // waits for work this very host thread has yet to enqueue.
  unsigned long long *pin_ring = nullptr;
  cudaEvent_t pin_done[64]{};
  uint32_t pin_calls = 0;
  dwj::PassFilter filter{};                    // DWJ_OPT_PASS_FILTER
  bool append_output = false;                  // DWJ_OPT_APPEND_OUTPUT
  bool pre_cleared = false;                    // dwj_clear_table ran: the next build skips its own clear
  cudaEvent_t ev_cleared = nullptr;
  unsigned long long *xpart_cursor = nullptr;  // PART_MAX cursors of dwj_xpart_scatter (own scratch: may run beside a local join)
  unsigned long long *pull_cursor = nullptr;   // PART_MAX cursors of dwj_region_scatter_segments: its own scratch, so the
                                               // receiving scatter on one stream can overlap a sending one on another
  // L2-locality regions: inputs are radix-partitioned on the top `region_bits` bits of the bucket index first
  uint32_t region_bits = 0;
  void *region_build = nullptr, *region_probe = nullptr;   // partitioned copies (keys then payloads)
  uint64_t region_build_rows = 0, region_probe_rows = 0;   // capacities in rows
  uint32_t launches_build = 0, launches_probe = 0;
  int staged_shape = -1;   // -1: per-key-width default
  // dwj_join_host staging
  void *stage = nullptr;
  uint64_t stage_bytes = 0;
  cudaStream_t hs[3]{};                        // H2D, compute, D2H
  cudaEvent_t hev[3 * HOST_STAGES]{};          // per slot: input landed, probe done, slot free
  cudaEvent_t htime[4]{};
  unsigned long long *h_counts = nullptr;      // pinned: per-slot match counts
  dwj_timing last_host{};
};

namespace {

template <class Kern, class Args>
cudaError_t launch(dwj_engine *e, Kern kern, dim3 grid, dim3 block, cudaStream_t s, Args args, bool table_window, size_t smem = 0) {
  if (smem > 40 * 1024) {        // dynamic + the kernel's static shared memory may pass the 48 KB default together
    cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ae != cudaSuccess) return ae;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  if (table_window && e->l2_window) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow = e->window;
    lc.attrs = attr;
    lc.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&lc, kern, args);
}

int ensure_tile_state(dwj_engine *e, uint64_t tiles, cudaStream_t s) {
  if (tiles + 1 <= e->tile_state_cap) return DWJ_OK;
  if (e->tile_state) {
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(e->tile_state));
    e->tile_state = nullptr;
    e->tile_state_cap = 0;
  }
  const uint64_t cap = std::max<uint64_t>(tiles + 1, 1024);
  CU(cudaMalloc(&e->tile_state, cap * sizeof(unsigned long long)));
  e->tile_state_cap = cap;
  return DWJ_OK;
}

// Small host arrays (scatter start rows, segment tables) reach the device through a KERNEL that reads a pinned, mapped
// staging slot -- not through cudaMemcpyAsync.  Two reasons, both met on hardware: an asynchronous copy from pageable
// memory synchronises the stream first; and copy-engine jobs of different streams share hardware queues, so a copy
// that waits behind a flag-polling kernel of the multi-GPU join (dwj_xj.cu) holds up another stream's copy that the
// polled flag depends on -- a dependency cycle the streams themselves do not contain.  Kernels have their own queues.
__global__ void stage_words_kernel(unsigned long long *dst, const unsigned long long *src, uint32_t words) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
int stage_words(unsigned long long *dst, const unsigned long long *pinned_src, uint32_t words, cudaStream_t s) {
  if (!words) return DWJ_OK;
  stage_words_kernel<<<(words + 255) / 256 > 8 ? 8 : (words + 255) / 256, 256, 0, s>>>(dst, pinned_src, words);
  CU(cudaGetLastError());
  return DWJ_OK;
}
// dst[0 .. words) = src[0 .. words) (words <= PART_MAX) on stream s; src may be reused at once.
constexpr uint32_t PIN_SLOTS = 64;
int upload_words(dwj_engine *e, unsigned long long *dst, const uint64_t *src, uint32_t words, cudaStream_t s) {
  const uint32_t slot = e->pin_calls++ % PIN_SLOTS;
  if (e->pin_calls > PIN_SLOTS) CU(cudaEventSynchronize(e->pin_done[slot]));     // the copy PIN_SLOTS calls ago has left the slot
  unsigned long long *h = e->pin_ring + (size_t)slot * dwj::PART_MAX;
  memcpy(h, src, words * sizeof(unsigned long long));
  if (int rc = stage_words(dst, h, words, s)) return rc;
  CU(cudaEventRecord(e->pin_done[slot], s));
  return DWJ_OK;
}

// Thread-private byte counters (histogram, > 8 partitions) and the ballot-ranked, shared-memory-staged scatter.
template <int W, uint32_t MODE, int THREADS>
int hist_many_launch(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  constexpr int HROWS = 8;
  auto kern = dwj::partition_hist_private_kernel<W, MODE, THREADS, HROWS>;
  const size_t smem = (size_t)THREADS << a.log2_parts;
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  const uint64_t htiles = (a.n + (uint64_t)THREADS * HROWS - 1) / ((uint64_t)THREADS * HROWS);
  const unsigned grid = (unsigned)std::min<uint64_t>(htiles, (uint64_t)e->prop.multiProcessorCount * std::max(per_sm, 1));
  CU(launch(e, kern, dim3(grid), dim3(THREADS), s, a, false, smem));
  return DWJ_OK;
}
template <int W, uint32_t MODE> int hist_launch_mode(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  if (a.log2_parts <= 3) {      // <= 8 partitions: packed-register counters
    constexpr int HROWS = 8;
    const uint64_t htiles = (a.n + 256ull * HROWS - 1) / (256ull * HROWS);
    const dim3 hgrid((unsigned)std::min<uint64_t>(htiles, (uint64_t)e->prop.multiProcessorCount * 8));
    CU(launch(e, dwj::partition_hist8_kernel<W, MODE, HROWS>, hgrid, dim3(dwj::PART_THREADS), s, a, false));
    return DWJ_OK;
  }
  return a.log2_parts <= 8 ? hist_many_launch<W, MODE, 256>(e, a, s) : hist_many_launch<W, MODE, 128>(e, a, s);
}
// Shared-memory RED histogram: any partition count up to 4096 and the only one that honours the pass filter.
template <int W, uint32_t MODE> int hist_red_launch(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  constexpr int HROWS = 8;
  const uint64_t htiles = (a.n + 256ull * HROWS - 1) / (256ull * HROWS);
  const unsigned grid = (unsigned)std::min<uint64_t>(htiles, (uint64_t)e->prop.multiProcessorCount * 8);
  CU(launch(e, dwj::partition_hist_red_kernel<W, MODE, HROWS>, dim3(grid), dim3(dwj::PART_THREADS), s, a, false, sizeof(unsigned int) << a.log2_parts));
  return DWJ_OK;
}
// Histogram of a.keys into a.hist (zeroed by the caller).
template <int W> int hist_launch(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  if (!a.n) return DWJ_OK;
  static const int red_env = getenv("DWJ_HIST_RED_FROM") ? atoi(getenv("DWJ_HIST_RED_FROM")) : -1;   // tuning sweep
  const int red_from = red_env >= 0 ? red_env : (W == 4 ? DEFAULT_HIST_RED_FROM_4 : DEFAULT_HIST_RED_FROM_8);
  if (a.filter.mask || a.log2_parts > 9 || (int)a.log2_parts >= red_from) {
    if (a.mode == dwj::PART_BY_BUCKET) return hist_red_launch<W, dwj::PART_BY_BUCKET>(e, a, s);
    if (a.mode == dwj::PART_BY_HASH) return hist_red_launch<W, dwj::PART_BY_HASH>(e, a, s);
    return hist_red_launch<W, dwj::PART_BY_BOTH>(e, a, s);
  }
  if (a.mode == dwj::PART_BY_BUCKET) return hist_launch_mode<W, dwj::PART_BY_BUCKET>(e, a, s);
  if (a.mode == dwj::PART_BY_HASH) return hist_launch_mode<W, dwj::PART_BY_HASH>(e, a, s);
  return hist_launch_mode<W, dwj::PART_BY_BOTH>(e, a, s);
}

// Shape of the many-way scatter: threads per CTA, rows per thread, CTAs per SM.  The tile (threads x rows) decides the
// length of the runs a CTA writes per partition (tile / partitions rows) and how far the per-tile overhead -- counter
// scan, one global reservation per partition -- is amortised; profiles/r2_partition_sweep.txt.  DWJ_SCATTER_SHAPE /
// DWJ_SCATTER_MATCH pick a shape / the MATCH-instruction ranking at run time for tuning sweeps.
struct ScatterShape { int threads, items, minb; };
constexpr ScatterShape SCATTER_SHAPES_4[] = {{256, 16, 3}, {512, 16, 2}, {512, 32, 1}, {1024, 16, 1}, {256, 32, 2}};
constexpr ScatterShape SCATTER_SHAPES_8[] = {{256, 8, 3}, {512, 8, 2}, {512, 16, 1}, {1024, 8, 1}, {256, 16, 2}};
constexpr int N_SCATTER_SHAPES = 5, FILTER_SCATTER_SHAPE = 0;   // under a pass filter only a fraction of a tile's rows is staged and
                                                                // written; measured on the 2^31 x 2^31 int64 join, one GPU, 2 passes
                                                                // (ms per filtered 2^31-row scatter / per step): shape 0 28.7 / 247,
                                                                // shape 1 25.7 / 256, shape 2 37.6 / 284, shape 4 (twice the rows
                                                                // per thread) 33.7 / 267
template <int W> constexpr ScatterShape scatter_shape_of(int i) { return W == 4 ? SCATTER_SHAPES_4[i] : SCATTER_SHAPES_8[i]; }
template <int W> size_t scatter_smem(int shape, uint32_t parts) {
  const ScatterShape sh = scatter_shape_of<W>(shape);
  return (size_t)sh.threads * sh.items * (2 * W + (W == 4 ? 2 : 0)) + (size_t)parts * (8 + 4 * (sh.threads / 32));
}
// Shape for a scatter into 2^bits partitions (bits < 5: always the small tile -- long runs anyway).  Measured
// (profiles/r2_partition_sweep.txt): the small tile wins everywhere except 512-way with 16-byte rows, where a tile of
// 2048 rows leaves 32-byte runs (2.25 TB/s; 2.76 with 4096 rows).  MATCH-instruction ranking lost everywhere.
template <int W> int pick_scatter_shape(uint32_t bits, bool filtered = false) {
  if (bits < 5) return 0;
  static const int forced = getenv("DWJ_SCATTER_SHAPE") ? atoi(getenv("DWJ_SCATTER_SHAPE")) : -1;
  static const int forced_filtered = getenv("DWJ_SCATTER_SHAPE_FILTERED") ? atoi(getenv("DWJ_SCATTER_SHAPE_FILTERED")) : -1;
  int shape = W == 8 && bits >= 9 ? 1 : 0;
  if (filtered) shape = forced_filtered >= 0 && forced_filtered < N_SCATTER_SHAPES ? forced_filtered : FILTER_SCATTER_SHAPE;
  else if (forced >= 0 && forced < N_SCATTER_SHAPES) shape = forced;
  while (shape > 0 && scatter_smem<W>(shape, 1u << bits) > 227 * 1024) --shape;
  return shape;
}
template <int W> uint64_t scatter_tile_rows(uint32_t bits) {
  const ScatterShape sh = scatter_shape_of<W>(pick_scatter_shape<W>(bits));
  return (uint64_t)sh.threads * sh.items;
}

template <int W, int BITS, int SHAPE, bool MATCH>
int scatter_many_launch_shape(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  constexpr ScatterShape S = scatter_shape_of<W>(SHAPE);
  constexpr uint64_t TILE = (uint64_t)S.threads * S.items;
  const uint64_t tiles = a.n_segs ? a.n : (a.n + TILE - 1) / TILE;
  auto kern = dwj::partition_scatter_many_kernel<W, BITS, S.threads, S.items, S.minb, MATCH>;
  CU(launch(e, kern, dim3((unsigned)std::min<uint64_t>(tiles, 0x7fffffffull)), dim3(S.threads), s, a, false, scatter_smem<W>(SHAPE, 1u << BITS)));
  return DWJ_OK;
}
template <int W, int BITS>
int scatter_many_launch(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  static const bool match = getenv("DWJ_SCATTER_MATCH") ? atoi(getenv("DWJ_SCATTER_MATCH")) != 0 : DEFAULT_SCATTER_MATCH;
  if constexpr (BITS >= 5) {
    switch (pick_scatter_shape<W>(BITS, a.filter.mask != 0)) {
    case 4: return match ? scatter_many_launch_shape<W, BITS, 4, true>(e, a, s) : scatter_many_launch_shape<W, BITS, 4, false>(e, a, s);
    case 1: return match ? scatter_many_launch_shape<W, BITS, 1, true>(e, a, s) : scatter_many_launch_shape<W, BITS, 1, false>(e, a, s);
    case 2: return match ? scatter_many_launch_shape<W, BITS, 2, true>(e, a, s) : scatter_many_launch_shape<W, BITS, 2, false>(e, a, s);
    case 3: return match ? scatter_many_launch_shape<W, BITS, 3, true>(e, a, s) : scatter_many_launch_shape<W, BITS, 3, false>(e, a, s);
    default: break;
    }
  }
  return match ? scatter_many_launch_shape<W, BITS, 0, true>(e, a, s) : scatter_many_launch_shape<W, BITS, 0, false>(e, a, s);
}
// Scatter a.keys / a.vals to a.out_* at the running positions in a.cursor (1 .. 512 partitions).
template <int W> int scatter_launch(dwj_engine *e, const dwj::PartitionArgs<W> &a, cudaStream_t s) {
  if (!a.n) return DWJ_OK;
  switch (a.log2_parts) {
  case 0:                         // one partition: a copy -- unless rows are filtered or gathered from segments
    if (a.filter.mask || a.n_segs) return scatter_many_launch<W, 1>(e, a, s);     // every row has partition id 0
    CU(cudaMemcpyAsync(a.out_keys, a.keys, a.n * W, cudaMemcpyDeviceToDevice, s));
    if (a.vals) CU(cudaMemcpyAsync(a.out_vals, a.vals, a.n * W, cudaMemcpyDeviceToDevice, s));
    return DWJ_OK;
  case 1: return scatter_many_launch<W, 1>(e, a, s);
  case 2: return scatter_many_launch<W, 2>(e, a, s);
  case 3: return scatter_many_launch<W, 3>(e, a, s);
  case 4: return scatter_many_launch<W, 4>(e, a, s);
  case 5: return scatter_many_launch<W, 5>(e, a, s);
  case 6: return scatter_many_launch<W, 6>(e, a, s);
  case 7: return scatter_many_launch<W, 7>(e, a, s);
  case 8: return scatter_many_launch<W, 8>(e, a, s);
  default: return scatter_many_launch<W, 9>(e, a, s);
  }
}

template <int W>
dwj::PartitionArgs<W> partition_args(const dwj_engine *e, const void *keys, const void *vals, uint64_t n, uint32_t log2_parts, uint32_t mode,
                                     uint32_t rank_bits) {
  using K = typename dwj::KeyT<W>::type;
  dwj::PartitionArgs<W> a{};
  a.keys = (const K *)keys;
  a.vals = (const K *)vals;
  a.n = n;
  a.log2_parts = log2_parts;
  a.seed = e->cfg.hash_seed;
  a.mode = mode;
  a.rank_bits = rank_bits;
  a.bucket_mask = e->buckets - 1;
  a.filter = e->filter;
  uint32_t lgb = 0;
  while ((1ull << lgb) < e->buckets) ++lgb;
  const uint32_t region_bits = mode == dwj::PART_BY_BOTH ? log2_parts - rank_bits : log2_parts;
  a.bucket_shift = lgb >= region_bits ? lgb - region_bits : 0;
  return a;
}

// mode PART_BY_HASH  : partition id from the independent partition hash (multi-GPU exchange, dwj_partition)
// mode PART_BY_BUCKET: partition id = top log2_parts bits of the bucket index (engine regions)
template <int W>
int partition_impl(dwj_engine *e, const void *keys, const void *vals, uint64_t n, uint32_t log2_parts, uint32_t mode, void *ok,
                   void *ov, uint64_t *d_offsets, cudaStream_t s) {
  using K = typename dwj::KeyT<W>::type;
  dwj::PartitionArgs<W> a = partition_args<W>(e, keys, vals, n, log2_parts, mode, 0);
  a.out_keys = (K *)ok;
  a.out_vals = (K *)ov;
  a.hist = e->part_scratch;
  a.cursor = e->part_scratch + dwj::PART_MAX;
  a.offsets = (unsigned long long *)d_offsets;
  CU(cudaMemsetAsync(e->part_scratch, 0, 2 * dwj::PART_MAX * sizeof(unsigned long long), s));
  if (int rc = hist_launch<W>(e, a, s)) return rc;
  CU(launch(e, dwj::partition_offsets_kernel<W>, dim3(1), dim3(32), s, a, false));
  return scatter_launch<W>(e, a, s);
}

// Histogram only (exchange planning) -> d_counts[parts].  PART_BY_HASH: the independent partition hash; PART_BY_BOTH:
// the combined (destination rank, table region) id of dwj_xpart_*.
template <int W> int partition_hist_impl(dwj_engine *e, const void *keys, uint64_t n, uint32_t log2_parts, uint32_t mode, uint32_t rank_bits,
                                         uint64_t *d_counts, cudaStream_t s) {
  dwj::PartitionArgs<W> a = partition_args<W>(e, keys, nullptr, n, log2_parts, mode, rank_bits);
  a.hist = (unsigned long long *)d_counts;
  CU(cudaMemsetAsync(d_counts, 0, sizeof(uint64_t) << log2_parts, s));
  return hist_launch<W>(e, a, s);
}

// Scatter with the layout planned by the caller: the rows of partition p go to out[h_start_rows[p] ...] (host array).
// The second half of a partition whose histogram fed an exchange plan; partitions may land anywhere inside the
// allocation `ok` / `ov` point into (e.g. some in a send buffer, some straight in a receive buffer).
template <int W>
int partition_scatter_planned_impl(dwj_engine *e, const void *keys, const void *vals, uint64_t n, uint32_t log2_parts, uint32_t mode,
                                   uint32_t rank_bits, const uint64_t *h_start_rows, void *ok, void *ov, cudaStream_t s) {
  using K = typename dwj::KeyT<W>::type;
  dwj::PartitionArgs<W> a = partition_args<W>(e, keys, vals, n, log2_parts, mode, rank_bits);
  a.out_keys = (K *)ok;
  a.out_vals = (K *)ov;
  a.cursor = e->xpart_cursor;
  if (int rc = upload_words(e, a.cursor, h_start_rows, 1u << log2_parts, s)) return rc;     // the caller's array may be reused at once
  if (log2_parts == 0 && n && !a.filter.mask) {       // one partition: a copy to the planned position
    CU(cudaMemcpyAsync((K *)ok + h_start_rows[0], keys, n * W, cudaMemcpyDeviceToDevice, s));
    if (vals) CU(cudaMemcpyAsync((K *)ov + h_start_rows[0], vals, n * W, cudaMemcpyDeviceToDevice, s));
    return DWJ_OK;
  }
  return scatter_launch<W>(e, a, s);
}

// Grow-only scratch for the region-partitioned copy of a relation: [keys | payloads], rows each.
int ensure_region_buffer(void **buf, uint64_t *cap_rows, uint64_t rows, int W, cudaStream_t s) {
  if (rows <= *cap_rows) return DWJ_OK;
  if (*buf) {
    CU(cudaStreamSynchronize(s));
    CU(cudaFree(*buf));
    *buf = nullptr;
    *cap_rows = 0;
  }
  CU(cudaMalloc(buf, 2 * rows * (uint64_t)W));
  *cap_rows = rows;
  return DWJ_OK;
}

// Segment table of the pending dwj_*_segments call for a kernel whose work unit is `unit_rows` rows: uploads
// {first unit, first row, rows} per segment (+ a closing entry) through a ring of pinned / device tables.  The caller
// records e->seg_done[slot] after the kernel that reads the table.
constexpr uint32_t SEG_SLOTS = 32, SEG_MAX = dwj::PART_MAX + 1;
int upload_segments(dwj_engine *e, uint64_t unit_rows, cudaStream_t s, const dwj::Seg **d_out, uint64_t *total_units, uint32_t *slot_out) {
  const uint32_t slot = e->seg_calls++ % SEG_SLOTS;
  if (e->seg_calls > SEG_SLOTS) CU(cudaEventSynchronize(e->seg_done[slot]));
  dwj::Seg *h = e->seg_tables_host + slot * SEG_MAX, *d = e->seg_tables + slot * SEG_MAX;
  unsigned long long units = 0;
  for (uint32_t i = 0; i < e->pending_segs; ++i) {
    h[i] = dwj::Seg{units, e->pending_seg_keys[i], e->pending_seg_vals ? e->pending_seg_vals[i] : nullptr, e->pending_seg_rows[i]};
    units += (e->pending_seg_rows[i] + unit_rows - 1) / unit_rows;
  }
  h[e->pending_segs] = dwj::Seg{units, nullptr, nullptr, 0};
  static_assert(sizeof(dwj::Seg) % sizeof(unsigned long long) == 0, "Seg is staged word by word");
  if (int rc = stage_words((unsigned long long *)d, (const unsigned long long *)h, (e->pending_segs + 1) * (uint32_t)(sizeof(dwj::Seg) / 8), s)) return rc;
  *d_out = d;
  *total_units = units;
  *slot_out = slot;
  return DWJ_OK;
}

template <int W>
int region_scatter_segments_impl(dwj_engine *e, const uint64_t *h_start_rows, void *ok, void *ov, bool with_vals, cudaStream_t s) {
  using K = typename dwj::KeyT<W>::type;
  dwj::PartitionArgs<W> a = partition_args<W>(e, nullptr, nullptr, 0, e->region_bits, dwj::PART_BY_BUCKET, 0);
  a.filter = dwj::PassFilter{};                 // the sender already dropped the rows of other key classes
  a.out_keys = (K *)ok;
  a.out_vals = (K *)ov;
  a.cursor = e->pull_cursor;
  if (int rc = upload_words(e, a.cursor, h_start_rows, 1u << e->region_bits, s)) return rc;
  const uint64_t TILE = scatter_tile_rows<W>(std::max<uint32_t>(e->region_bits, 1));     // one partition runs the 2-way kernel
  uint64_t tiles = 0;
  uint32_t slot = 0;
  if (int rc = upload_segments(e, TILE, s, &a.segs, &tiles, &slot)) return rc;
  a.n = tiles;
  a.n_segs = e->pending_segs;
  if (with_vals) a.vals = (const K *)(uintptr_t)1;      // non-null: the kernel takes the payload pointers from the segments
  const int rc = scatter_launch<W>(e, a, s);
  CU(cudaEventRecord(e->seg_done[slot], s));
  return rc;
}

template <int W>
int filter_rows_impl(dwj_engine *e, const void *keys, const void *vals, uint64_t n, void *ok, void *ov, uint64_t *d_n_out,
                     uint64_t *d_region_counts, cudaStream_t s) {
  using K = typename dwj::KeyT<W>::type;
  // the partition id only serves the optional region histogram; a.filter = the pass filter
  dwj::PartitionArgs<W> a = partition_args<W>(e, keys, vals, n, d_region_counts ? e->region_bits : 0, dwj::PART_BY_BUCKET, 0);
  a.out_keys = (K *)ok;
  a.out_vals = (K *)ov;
  a.cursor = (unsigned long long *)d_n_out;                // the write cursor IS the row count
  a.hist = (unsigned long long *)d_region_counts;
  CU(cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), s));
  if (d_region_counts) CU(cudaMemsetAsync(d_region_counts, 0, sizeof(uint64_t) << e->region_bits, s));
  if (!n) return DWJ_OK;
  if (!a.filter.mask && !d_region_counts) {                // no filter, no histogram: a copy
    CU(cudaMemcpyAsync(ok, keys, n * W, cudaMemcpyDeviceToDevice, s));
    if (vals) CU(cudaMemcpyAsync(ov, vals, n * W, cudaMemcpyDeviceToDevice, s));
    dwj::stage_value_kernel<<<1, 1, 0, s>>>((unsigned long long *)d_n_out, (unsigned long long)n);
    CU(cudaGetLastError());
    return DWJ_OK;
  }
  constexpr int THREADS = 256, ITEMS = 8;
  const uint64_t tiles = (n + (uint64_t)THREADS * ITEMS - 1) / ((uint64_t)THREADS * ITEMS);
  const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)e->prop.multiProcessorCount * 8);
  CU(launch(e, dwj::filter_rows_kernel<W, THREADS, ITEMS>, dim3(grid), dim3(THREADS), s, a, false));
  return DWJ_OK;
}

// grouped: the caller's rows are already grouped by table region (dwj_build_grouped); `grouped_offsets` (device,
// regions + 1 entries, may be null) are the regions' row ranges for the look-ahead.
template <int W> int build_impl(dwj_engine *e, const void *keys, const void *vals, uint64_t n, cudaStream_t s, bool grouped = false,
                                const uint64_t *grouped_offsets = nullptr) {
  using K = typename dwj::KeyT<W>::type;
  CU(cudaEventRecord(e->ev_build[0], s));
  e->launches_build = 0;
  const bool partitioned = e->region_bits && n && !grouped;
  if (partitioned) {      // group the rows by table region first: the inserts then hit an L2-resident slice
    if (int rc = ensure_region_buffer(&e->region_build, &e->region_build_rows, std::max<uint64_t>(n, e->cfg.max_build_rows), W, s)) return rc;
    K *pk = (K *)e->region_build, *pv = pk + e->region_build_rows;
    if (int rc = partition_impl<W>(e, keys, vals, n, e->region_bits, dwj::PART_BY_BUCKET, pk, pv, (uint64_t *)(e->part_scratch + 2 * dwj::PART_MAX), s)) return rc;
    keys = pk;
    vals = pv;
    e->launches_build += 4;
  }
  if (e->pre_cleared) {                 // dwj_clear_table did both clears, possibly on another stream
    CU(cudaStreamWaitEvent(s, e->ev_cleared, 0));
  } else {
    CU(cudaMemsetAsync(e->table, 0xFF, e->table_bytes, s));
    e->launches_build++;
  }
  if (n) {
    if (!e->pre_cleared) CU(cudaMemsetAsync(e->fill, 0, e->buckets * sizeof(unsigned int), s));
    dwj::BuildArgs<W> a{};
    a.keys = (const K *)keys;
    a.vals = (const K *)vals;
    a.n = n;
    a.table = e->table;
    a.fill = e->fill;
    a.bucket_mask = e->buckets - 1;
    a.seed = e->cfg.hash_seed;
    if (partitioned && e->filter.mask) a.n_dev = e->part_scratch + 2 * dwj::PART_MAX + (1u << e->region_bits);   // rows that passed the filter
    else a.filter = e->filter;          // unpartitioned input: the insert kernel filters
    static const bool no_ahead = getenv("DWJ_BUILD_NO_AHEAD") && atoi(getenv("DWJ_BUILD_NO_AHEAD"));   // A/B switch (development)
    if ((partitioned || (grouped && grouped_offsets && e->region_bits)) && !no_ahead) {   // region look-ahead: offsets stay on the device
      a.offsets = partitioned ? e->part_scratch + 2 * dwj::PART_MAX : (const unsigned long long *)grouped_offsets;
      a.regions = 1u << e->region_bits;
      a.slice_bytes = e->table_bytes >> e->region_bits;
    }
    constexpr int ROWS = 4;
    uint64_t tiles = (n + 256ull * ROWS - 1) / (256ull * ROWS);
    uint32_t seg_slot = 0;
    if (e->pending_segs) {              // dwj_build_segments: rows come from a segment list, region by region
      if (int rc = upload_segments(e, 256ull * ROWS, s, &a.segs, &tiles, &seg_slot)) return rc;
      a.n = tiles;
      a.n_segs = e->pending_segs;
      a.segs_per_region = e->pending_segs_per_region;
      if (e->region_bits && !no_ahead && a.segs_per_region && a.n_segs == (a.segs_per_region << e->region_bits)) {
        a.regions = 1u << e->region_bits;
        a.slice_bytes = e->table_bytes >> e->region_bits;
      }
    }
    CU(cudaEventRecord(e->ev_buildk[0], s));
    const dim3 grid((unsigned)std::min<uint64_t>(tiles, 0x7fffffffull));
    if (e->runs) {                      // one-to-many engine: count -> offsets -> fill (csr.cuh)
      unsigned long long *cursor = e->counter + 5;
      CU(cudaMemsetAsync(cursor, 0, 2 * sizeof(unsigned long long), s));      // cursor, overflow flag
      if (tiles) CU(launch(e, dwj::csr_count_kernel<W, ROWS>, grid, dim3(256), s, a, true));
      dwj::CsrOffsetArgs<W> oa{e->table, e->buckets, (K *)e->runs, cursor, e->runs_granules, (unsigned int *)(e->counter + 6)};
      CU(launch(e, dwj::csr_offsets_kernel<W>, dim3((unsigned)((e->buckets + 255) / 256)), dim3(256), s, oa, true));
      if (tiles) {
        dwj::csr_fill_kernel<W, ROWS><<<grid, dim3(256), 0, s>>>(a, (K *)e->runs);
        CU(cudaGetLastError());
      }
      e->launches_build += 3;
    } else if (tiles) {
      CU(launch(e, dwj::build_kernel<W, ROWS>, grid, dim3(256), s, a, true));
    }
    CU(cudaEventRecord(e->ev_buildk[1], s));
    if (e->pending_segs) CU(cudaEventRecord(e->seg_done[seg_slot], s));
    e->launches_build += 2;
  }
  e->pre_cleared = false;
  CU(cudaEventRecord(e->ev_build[1], s));
  e->have_build = true;
  e->build_rows = n;
  e->built = true;
  e->csr = e->runs != nullptr;
  return DWJ_OK;
}

// ALIGNED / CONTAINS / COUNT: warp-centric kernel, no barriers, grid sized to fill the machine.
template <int W, int MODE, bool UNIQUE>
int simple_launch(dwj_engine *e, dwj::ProbeArgs<W> a, cudaStream_t s) {
  constexpr int ITEMS = W == 4 ? SIMPLE_ITEMS_4 : SIMPLE_ITEMS_8;
  const uint64_t tiles = (a.n + 32ull * ITEMS - 1) / (32ull * ITEMS);
  e->launches_probe = 0;
  if (MODE == dwj::PROBE_COUNT) {
    CU(cudaMemsetAsync(a.n_matches, 0, sizeof(unsigned long long), s));
    e->launches_probe++;
  }
  if (tiles) {
    const uint64_t ctas = (tiles + 7) / 8;
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, (uint64_t)e->prop.multiProcessorCount * 16);
    CU(launch(e, dwj::probe_simple_kernel<W, MODE, UNIQUE, ITEMS>, dim3(grid), dim3(256), s, a, true));
    e->launches_probe++;
  }
  return DWJ_OK;
}

// PAIRS over a one-to-many table: one look-back descriptor (or one atomic) per CTA tile.  Shapes = threads per CTA, probe
// rows per thread, min CTAs per SM (DWJ_MULTI_SHAPE selects one; the default is the sweep's winner, profiles/r2_csr.md).
struct MultiShape { int threads, items, minb; };
constexpr MultiShape MULTI_SHAPES_4[] = {{256, 4, 4}, {256, 4, 3}, {256, 2, 5}, {256, 1, 8}};
constexpr MultiShape MULTI_SHAPES_8[] = {{256, 2, 4}, {256, 2, 3}, {256, 2, 4}, {256, 1, 8}};   // [2] = [0]: the sweep ran on 4-byte keys
constexpr int DEFAULT_MULTI_SHAPE = 2;

template <int W, bool ORDERED, int SHAPE>
int multi_launch_shape(dwj_engine *e, dwj::ProbeArgs<W> a, cudaStream_t s) {
  constexpr MultiShape S = W == 4 ? MULTI_SHAPES_4[SHAPE] : MULTI_SHAPES_8[SHAPE];
  constexpr int THREADS = S.threads, ITEMS = S.items, MINB = S.minb;
  constexpr uint64_t TILE = (uint64_t)THREADS * ITEMS;
  const uint64_t tiles = (a.n + TILE - 1) / TILE;
  a.num_tiles = tiles;
  e->launches_probe = 0;
  if (ORDERED) {
    int rc = ensure_tile_state(e, tiles, s);
    if (rc) return rc;
    a.tile_state = e->tile_state;
    CU(cudaMemsetAsync(e->tile_state, 0, (tiles + 1) * sizeof(unsigned long long), s));
    e->launches_probe++;
  }
  if (!e->append_output) {
    CU(cudaMemsetAsync(a.n_matches, 0, sizeof(unsigned long long), s));
    e->launches_probe++;
  }
  if (tiles) {
    if (tiles > 0x7fffffffull) return fail(DWJ_ERR_INVALID, "probe of %llu rows needs more than 2^31 tiles", (unsigned long long)a.n);
    const size_t stage = (size_t)TILE * 4 * W * (a.out_key ? 3 : 2);       // result rows of a tile, staged for contiguous stores
    CU(launch(e, dwj::probe_pairs_multi_kernel<W, ORDERED, THREADS, ITEMS, MINB>, dim3((unsigned)tiles), dim3(THREADS), s, a, true, stage));
    e->launches_probe++;
  }
  return DWJ_OK;
}

template <int W, bool ORDERED>
int multi_launch(dwj_engine *e, dwj::ProbeArgs<W> a, cudaStream_t s) {
  static const int shape = getenv("DWJ_MULTI_SHAPE") ? atoi(getenv("DWJ_MULTI_SHAPE")) : DEFAULT_MULTI_SHAPE;
  switch (shape) {
  case 1: return multi_launch_shape<W, ORDERED, 1>(e, a, s);
  case 3: return multi_launch_shape<W, ORDERED, 3>(e, a, s);
  case 0: return multi_launch_shape<W, ORDERED, 0>(e, a, s);
  default: return multi_launch_shape<W, ORDERED, 2>(e, a, s);
  }
}

template <int W, bool ORDERED, bool WITH_KEY, int SHAPE>
int staged_launch_shape(dwj_engine *e, dwj::ProbeArgs<W> a, cudaStream_t s) {
  constexpr StagedShape S = W == 4 ? STAGED_SHAPES_4[SHAPE] : STAGED_SHAPES_8[SHAPE];
  constexpr uint64_t CHUNK = (uint64_t)S.warps * 32 * S.items * S.sub;
  uint64_t chunks = (a.n + CHUNK - 1) / CHUNK;
  uint32_t seg_slot = 0;
  if (e->pending_segs) {                // dwj_probe_pairs_segments
    if (int rc = upload_segments(e, CHUNK, s, &a.segs, &chunks, &seg_slot)) return rc;
    a.n_segs = e->pending_segs;
  }
  a.num_tiles = chunks;
  e->launches_probe = 0;
  if (ORDERED) {
    int rc = ensure_tile_state(e, chunks, s);
    if (rc) return rc;
    a.tile_state = e->tile_state;
    CU(cudaMemsetAsync(e->tile_state, 0, (chunks + 1) * sizeof(unsigned long long), s));
    e->launches_probe++;
  }
  if (!e->append_output) {
    CU(cudaMemsetAsync(a.n_matches, 0, sizeof(unsigned long long), s));
    e->launches_probe++;
  }
  if (chunks) {
    if (chunks > 0x7fffffffull) return fail(DWJ_ERR_INVALID, "probe of %llu rows needs more than 2^31 chunks", (unsigned long long)a.n);
    unsigned int *hot = (unsigned int *)(e->counter + 4);
    if (!e->pending_segs && a.n >= 65536) {          // skewed probe keys? (the staged kernel bypasses L1 for table sectors otherwise)
      dwj::probe_skew_sample_kernel<W><<<1, 1024, 0, s>>>(a.keys, a.n, a.n_dev, a.bucket_mask, a.seed, hot);
      CU(cudaGetLastError());
      a.hot_keys = hot;
      e->launches_probe++;
    }
    auto kern = dwj::probe_pairs_staged_kernel<W, ORDERED, WITH_KEY, S.warps, S.items, S.sub, S.minb>;
    const size_t smem = CHUNK * W * (WITH_KEY ? 3 : 2);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)chunks);
    lc.blockDim = dim3(S.warps * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = s;
    cudaLaunchAttribute attr[1];
    if (e->l2_window) {
      attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[0].val.accessPolicyWindow = e->window;
      lc.attrs = attr;
      lc.numAttrs = 1;
    }
    CU(cudaLaunchKernelEx(&lc, kern, a));
    e->launches_probe++;
  }
  if (e->pending_segs) CU(cudaEventRecord(e->seg_done[seg_slot], s));
  return DWJ_OK;
}

template <int W, bool ORDERED, bool WITH_KEY>
int staged_launch(dwj_engine *e, dwj::ProbeArgs<W> a, cudaStream_t s) {
  const int shape = e->staged_shape >= 0 ? e->staged_shape : (W == 4 ? DEFAULT_STAGED_SHAPE_4 : DEFAULT_STAGED_SHAPE_8);
  switch (shape) {
  case 1: return staged_launch_shape<W, ORDERED, WITH_KEY, 1>(e, a, s);
  case 2: return staged_launch_shape<W, ORDERED, WITH_KEY, 2>(e, a, s);
  case 3: return staged_launch_shape<W, ORDERED, WITH_KEY, 3>(e, a, s);
  case 4: return staged_launch_shape<W, ORDERED, WITH_KEY, 4>(e, a, s);
  default: return staged_launch_shape<W, ORDERED, WITH_KEY, 0>(e, a, s);
  }
}

template <int W>
int probe_impl(dwj_engine *e, int mode, const void *keys, const void *vals, uint64_t n, void *ok, void *ob, void *op,
               uint32_t *flags, uint64_t capacity, uint64_t *d_n, uint64_t *h_n, cudaStream_t s, bool grouped = false) {
  using K = typename dwj::KeyT<W>::type;
  if (!e->built) return fail(DWJ_ERR_STATE, "probe before dwj_build");
  if (e->append_output && mode == dwj::PROBE_PAIRS && !d_n) return fail(DWJ_ERR_INVALID, "DWJ_OPT_APPEND_OUTPUT needs the device counter d_n_matches");
  dwj::ProbeArgs<W> a{};
  a.keys = (const K *)keys;
  a.vals = (const K *)vals;
  a.n = n;
  a.table = e->table;
  a.bucket_mask = e->buckets - 1;
  a.seed = e->cfg.hash_seed;
  a.out_key = (K *)ok;
  a.out_build_val = (K *)ob;
  a.out_probe_val = (K *)op;
  a.out_flags = flags;
  a.capacity = capacity;
  a.n_matches = d_n ? (unsigned long long *)d_n : e->counter;
  a.runs = e->csr ? (const K *)e->runs : nullptr;
  // one-to-many tables come from the builds of engines without the flag; what dwj_aggregate_sum leaves holds distinct keys
  const bool unique = (e->cfg.flags & DWJ_FLAG_UNIQUE_BUILD_KEYS) != 0 || !e->csr;
  CU(cudaEventRecord(e->ev_probe[0], s));
  int rc;
  uint32_t extra_launches = 0;
  if (e->region_bits && n && !grouped && (mode == dwj::PROBE_PAIRS || mode == dwj::PROBE_COUNT)) {
    // Same grouping for the probe relation: its rows then walk the table slice by slice (output order becomes
    // region-major; the row multiset is unchanged).
    if ((rc = ensure_region_buffer(&e->region_probe, &e->region_probe_rows, n, W, s))) return rc;
    K *pk = (K *)e->region_probe, *pv = pk + e->region_probe_rows;
    const bool with_vals = mode == dwj::PROBE_PAIRS;
    if ((rc = partition_impl<W>(e, keys, with_vals ? vals : nullptr, n, e->region_bits, dwj::PART_BY_BUCKET, pk, with_vals ? pv : nullptr,
                                (uint64_t *)(e->part_scratch + 2 * dwj::PART_MAX), s)))
      return rc;
    a.keys = pk;
    a.vals = pv;
    if (e->filter.mask) a.n_dev = e->part_scratch + 2 * dwj::PART_MAX + (1u << e->region_bits);   // rows that passed the filter
    extra_launches = 4;
  }
  CU(cudaEventRecord(e->ev_probek[0], s));
  switch (mode) {
  case dwj::PROBE_ALIGNED: rc = simple_launch<W, dwj::PROBE_ALIGNED, true>(e, a, s); break;
  case dwj::PROBE_CONTAINS: rc = simple_launch<W, dwj::PROBE_CONTAINS, true>(e, a, s); break;
  case dwj::PROBE_COUNT:
    rc = unique ? simple_launch<W, dwj::PROBE_COUNT, true>(e, a, s) : simple_launch<W, dwj::PROBE_COUNT, false>(e, a, s);
    break;
  default:
    {
      // Probe-row order only exists when the kernel sees the caller's rows in the caller's order: a region-partitioned,
      // grouped or segmented input is already permuted, so the order-preserving look-back would buy nothing.
      static const bool keep_lookback = getenv("DWJ_KEEP_LOOKBACK") && atoi(getenv("DWJ_KEEP_LOOKBACK"));   // A/B switch (development)
      const bool permuted = extra_launches || grouped;
      // Appending (DWJ_OPT_APPEND_OUTPUT) continues at the running count in *d_n_matches: the atomic variants do that by
      // construction, the look-back variants start from zero.
      const bool ordered = !(e->cfg.flags & DWJ_FLAG_UNORDERED_OUTPUT) && (!permuted || keep_lookback) && !e->append_output;
      if (unique) {      // rows staged in shared memory; one look-back per chunk (ordered) or one atomic per warp
        if (ordered) rc = ok ? staged_launch<W, true, true>(e, a, s) : staged_launch<W, true, false>(e, a, s);
        else rc = ok ? staged_launch<W, false, true>(e, a, s) : staged_launch<W, false, false>(e, a, s);
      } else {
        rc = ordered ? multi_launch<W, true>(e, a, s) : multi_launch<W, false>(e, a, s);
      }
    }
  }
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_probek[1], s));
  e->launches_probe += extra_launches;
  CU(cudaEventRecord(e->ev_probe[1], s));
  e->have_probe = true;
  if (h_n) {
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, a.n_matches, sizeof(total), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *h_n = total;
    if (e->csr) {                       // the last one-to-many build ran out of run space (csr_offsets_kernel): keys were dropped
      unsigned int overflow = 0;
      CU(cudaMemcpy(&overflow, e->counter + 6, sizeof(overflow), cudaMemcpyDeviceToHost));
      if (overflow) return fail(DWJ_ERR_CAPACITY, "the last build held more rows than the %llu the one-to-many engine was created for", (unsigned long long)e->cfg.max_build_rows);
    }
    if (mode == dwj::PROBE_PAIRS && total > capacity)
      return fail(DWJ_ERR_OVERFLOW, "join produced %llu rows, output capacity is %llu", total, (unsigned long long)capacity);
  }
  return DWJ_OK;
}

int check_engine(const dwj_engine *e) { return e ? DWJ_OK : fail(DWJ_ERR_INVALID, "null engine"); }

}  // namespace

extern "C" {

int dwj_abi_version(void) { return DWJ_ABI_VERSION; }
// dwj_xj.cu reports through the same per-thread buffer (hidden visibility: not part of the ABI)
void dwj_internal_set_error(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
const char *dwj_last_error(void) { return g_err; }

int dwj_create(const dwj_config *cfg, dwj_engine **out) {
  if (!cfg || !out) return fail(DWJ_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->key_bytes != 4 && cfg->key_bytes != 8) return fail(DWJ_ERR_INVALID, "key_bytes must be 4 or 8, got %d", cfg->key_bytes);
  if (cfg->payload_bytes != cfg->key_bytes) return fail(DWJ_ERR_INVALID, "payload_bytes must equal key_bytes");
  double lf = cfg->load_factor == 0.0 ? 0.5 : cfg->load_factor;
  if (!(lf > 0.0 && lf <= 0.9)) return fail(DWJ_ERR_INVALID, "load_factor must be in (0, 0.9], got %g", cfg->load_factor);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(DWJ_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(DWJ_ERR_INVALID, "device %d out of range (have %d)", cfg->device, ndev);

  dwj_engine *e = new (std::nothrow) dwj_engine();
  if (!e) return fail(DWJ_ERR_OOM, "host allocation failed");
  e->cfg = *cfg;
  e->cfg.load_factor = lf;
  e->W = cfg->key_bytes;
  if (const char *ps = getenv("DWJ_STAGED_SHAPE")) e->staged_shape = atoi(ps);
  DeviceGuard g(cfg->device);
  auto bail = [&](int rc) { dwj_destroy(e); return rc; };
  if (cudaGetDeviceProperties(&e->prop, cfg->device) != cudaSuccess) return bail(fail(DWJ_ERR_CUDA, "cudaGetDeviceProperties failed"));
  if (e->prop.major < 10) return bail(fail(DWJ_ERR_CUDA, "device %d is sm_%d%d; this library holds sm_100a code only", cfg->device, e->prop.major, e->prop.minor));

  const uint32_t spb = 32u / (2u * (uint32_t)e->W);
  uint64_t need = (uint64_t)((double)std::max<uint64_t>(cfg->max_build_rows, 1) / lf + 0.999999);
  e->slots = std::max<uint64_t>(next_pow2(need), spb);
  e->buckets = e->slots / spb;
  e->table_bytes = e->buckets * 32ull;
  cudaError_t me = cudaMalloc(&e->table, e->table_bytes);
  if (me != cudaSuccess) return bail(fail(DWJ_ERR_OOM, "cudaMalloc of a %llu-byte table failed: %s", (unsigned long long)e->table_bytes, cudaGetErrorString(me)));
  if (cudaMalloc((void **)&e->fill, e->buckets * sizeof(unsigned int)) != cudaSuccess)
    return bail(fail(DWJ_ERR_OOM, "cudaMalloc of the %llu-byte ticket array failed", (unsigned long long)(e->buckets * sizeof(unsigned int))));
  if (!(cfg->flags & DWJ_FLAG_UNIQUE_BUILD_KEYS)) {      // payload runs: header + payloads per key, 32-byte granules (csr.cuh)
    const uint64_t g = 32u / (uint32_t)e->W, rows = std::max<uint64_t>(cfg->max_build_rows, 1);
    e->runs_granules = ((1 + g) * rows + g - 1) / g + 1;
    if (e->W == 4 && e->runs_granules >= 0xFFFFFFFFull)
      return bail(fail(DWJ_ERR_INVALID, "a one-to-many engine with 4-byte keys holds at most %llu build rows", (unsigned long long)(0xFFFFFFFFull * g / (1 + g) - 2)));
    if (cudaMalloc(&e->runs, e->runs_granules * 32) != cudaSuccess)
      return bail(fail(DWJ_ERR_OOM, "cudaMalloc of %llu bytes for the one-to-many payload runs failed", (unsigned long long)(e->runs_granules * 32)));
  }
  if (cudaMalloc(&e->xpart_cursor, dwj::PART_MAX * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&e->pull_cursor, dwj::PART_MAX * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&e->counter, 64) != cudaSuccess || cudaMalloc(&e->part_scratch, (3 * dwj::PART_MAX + 1) * sizeof(unsigned long long)) != cudaSuccess)
    return bail(fail(DWJ_ERR_OOM, "scratch allocation failed"));
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreate(&e->ev_build[i]) != cudaSuccess || cudaEventCreate(&e->ev_probe[i]) != cudaSuccess ||
        cudaEventCreate(&e->ev_part[i]) != cudaSuccess || cudaEventCreate(&e->ev_probek[i]) != cudaSuccess ||
        cudaEventCreate(&e->ev_buildk[i]) != cudaSuccess)
      return bail(fail(DWJ_ERR_CUDA, "cudaEventCreate failed"));

  // staging rings of the segment / start-row uploads: allocated here, never inside a join (an allocation may synchronise
  // the device, and a multi-GPU join has kernels parked on flags only later host work will raise)
  if (cudaMalloc((void **)&e->seg_tables, SEG_SLOTS * SEG_MAX * sizeof(dwj::Seg)) != cudaSuccess ||
      cudaHostAlloc((void **)&e->seg_tables_host, SEG_SLOTS * SEG_MAX * sizeof(dwj::Seg), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
      cudaHostAlloc((void **)&e->pin_ring, (size_t)PIN_SLOTS * dwj::PART_MAX * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
    return bail(fail(DWJ_ERR_OOM, "staging ring allocation failed"));
  for (auto &ev : e->seg_done)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(fail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
  for (auto &ev : e->pin_done)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(fail(DWJ_ERR_CUDA, "cudaEventCreate failed"));

  {   // L2-locality regions (see partition.cuh).  DWJ_REGION_MB / DWJ_PARTITION_MIN_MB are tuning overrides.
    double region_mb = 32.0, min_mb = 192.0;
    if (const char *v = getenv("DWJ_REGION_MB")) region_mb = atof(v);
    if (const char *v = getenv("DWJ_PARTITION_MIN_MB")) min_mb = atof(v);
    if (!(cfg->flags & DWJ_FLAG_NO_PARTITION) && region_mb > 0 && (double)e->table_bytes >= min_mb * 1048576.0) {
      uint32_t bits = 0, lgb = 0;
      while ((double)(e->table_bytes >> bits) > region_mb * 1048576.0) ++bits;
      while ((1ull << lgb) < e->buckets) ++lgb;
      uint32_t max_bits = 0;
      while ((1u << (max_bits + 1)) <= (uint32_t)dwj::PART_MAX) ++max_bits;
      e->region_bits = std::min(std::min(bits, max_bits), lgb);
    }
  }

  if ((cfg->flags & DWJ_FLAG_L2_PERSIST) && e->prop.persistingL2CacheMaxSize > 0 && e->prop.accessPolicyMaxWindowSize > 0) {
    const size_t carve = std::min<size_t>((size_t)e->prop.persistingL2CacheMaxSize, (size_t)e->table_bytes);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
      e->window.base_ptr = e->table;
      e->window.num_bytes = std::min<size_t>((size_t)e->table_bytes, (size_t)e->prop.accessPolicyMaxWindowSize);
      e->window.hitRatio = (float)std::min(1.0, (double)carve / (double)e->window.num_bytes);
      e->window.hitProp = cudaAccessPropertyPersisting;
      e->window.missProp = cudaAccessPropertyStreaming;
      e->l2_window = true;
    } else {
      cudaGetLastError();
    }
  }
  cudaMemset(e->counter, 0, 64);
  *out = e;
  return DWJ_OK;
}

int dwj_destroy(dwj_engine *e) {
  if (!e) return DWJ_OK;
  DeviceGuard g(e->cfg.device);
  cudaDeviceSynchronize();
  if (e->l2_window) cudaCtxResetPersistingL2Cache();
  cudaFree(e->table);
  cudaFree(e->fill);
  cudaFree(e->runs);
  cudaFree(e->tile_state);
  cudaFree(e->counter);
  cudaFree(e->part_scratch);
  cudaFree(e->pull_cursor);
  if (e->ev_cleared) cudaEventDestroy(e->ev_cleared);
  cudaFree(e->xpart_cursor);
  cudaFree(e->seg_tables);
  if (e->seg_tables_host) cudaFreeHost(e->seg_tables_host);
  if (e->pin_ring) cudaFreeHost(e->pin_ring);
  for (auto &ev : e->pin_done)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : e->seg_done)
    if (ev) cudaEventDestroy(ev);
  cudaFree(e->region_build);
  cudaFree(e->region_probe);
  cudaFree(e->stage);
  for (int i = 0; i < 2; ++i) {
    if (e->ev_build[i]) cudaEventDestroy(e->ev_build[i]);
    if (e->ev_probe[i]) cudaEventDestroy(e->ev_probe[i]);
    if (e->ev_part[i]) cudaEventDestroy(e->ev_part[i]);
    if (e->ev_probek[i]) cudaEventDestroy(e->ev_probek[i]);
    if (e->ev_buildk[i]) cudaEventDestroy(e->ev_buildk[i]);
  }
  for (auto &st : e->hs)
    if (st) cudaStreamDestroy(st);
  for (auto &ev : e->hev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : e->htime)
    if (ev) cudaEventDestroy(ev);
  if (e->h_counts) cudaFreeHost(e->h_counts);
  cudaGetLastError();
  delete e;
  return DWJ_OK;
}

int dwj_get_info(const dwj_engine *e, dwj_info *info) {
  if (!e || !info) return fail(DWJ_ERR_INVALID, "null argument");
  info->slots = e->slots;
  info->table_bytes = e->table_bytes;
  info->build_rows = e->build_rows;
  info->slot_bytes = 2u * (uint32_t)e->W;
  info->slots_per_bucket = 32u / info->slot_bytes;
  info->l2_persist = e->l2_window ? 1u : 0u;
  info->sm_count = (uint32_t)e->prop.multiProcessorCount;
  info->l2_bytes = (uint64_t)e->prop.l2CacheSize;
  info->launches_build = e->launches_build;
  info->launches_probe = e->launches_probe;
  info->radix_parts = 1u << e->region_bits;
  info->probe_passes = 1u;
  info->flags = e->cfg.flags;
  info->device = e->cfg.device;
  info->hash_seed = e->cfg.hash_seed;
  info->max_build_rows = e->cfg.max_build_rows;
  info->hot_probe_keys = 0;
  info->reserved = 0;
  if (e->have_probe) {                  // verdict of the last staged probe's key sample (synchronises with the device)
    unsigned int hot = 0;
    DeviceGuard g(e->cfg.device);
    if (cudaMemcpy(&hot, (const unsigned int *)(e->counter + 4), sizeof(hot), cudaMemcpyDeviceToHost) == cudaSuccess) info->hot_probe_keys = hot;
    else cudaGetLastError();
  }
  return DWJ_OK;
}

int dwj_build(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals)) return fail(DWJ_ERR_INVALID, "null build column");
  // under a pass filter only one key class in 2^pass_bits is inserted (classes are hash bits: an even split)
  const uint64_t expect = e->filter.mask ? n_rows / ((uint64_t)e->filter.mask + 1) : n_rows;
  if (expect > e->cfg.max_build_rows && ((double)expect > 0.9 * (double)e->slots || e->runs))     // payload runs are sized for max_build_rows
    return fail(DWJ_ERR_CAPACITY, "%llu build rows exceed the table created for %llu", (unsigned long long)expect,
                (unsigned long long)e->cfg.max_build_rows);
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? build_impl<4>(e, d_keys, d_vals, n_rows, (cudaStream_t)stream)
                   : build_impl<8>(e, d_keys, d_vals, n_rows, (cudaStream_t)stream);
}

int dwj_probe_aligned(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_key,
                      void *d_out_build_val, void *d_out_probe_val, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals || !d_out_key || !d_out_build_val || !d_out_probe_val))
    return fail(DWJ_ERR_INVALID, "null probe column or output");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_ALIGNED, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream)
                   : probe_impl<8>(e, dwj::PROBE_ALIGNED, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int dwj_probe_contains(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t *d_out_flags, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_out_flags)) return fail(DWJ_ERR_INVALID, "null probe column or output");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_CONTAINS, d_keys, nullptr, n_rows, nullptr, nullptr, nullptr, d_out_flags, 0, nullptr, nullptr, (cudaStream_t)stream)
                   : probe_impl<8>(e, dwj::PROBE_CONTAINS, d_keys, nullptr, n_rows, nullptr, nullptr, nullptr, d_out_flags, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int dwj_probe_pairs(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_key,
                    void *d_out_build_val, void *d_out_probe_val, uint64_t capacity, uint64_t *d_n_matches,
                    uint64_t *n_matches, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals)) return fail(DWJ_ERR_INVALID, "null probe column");
  if (capacity && (!d_out_build_val || !d_out_probe_val)) return fail(DWJ_ERR_INVALID, "null output column");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_PAIRS, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream)
                   : probe_impl<8>(e, dwj::PROBE_PAIRS, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream);
}

int dwj_probe_count(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint64_t *d_n_matches, uint64_t *n_matches,
                    void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && !d_keys) return fail(DWJ_ERR_INVALID, "null probe column");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_COUNT, d_keys, nullptr, n_rows, nullptr, nullptr, nullptr, nullptr, 0, d_n_matches, n_matches, (cudaStream_t)stream)
                   : probe_impl<8>(e, dwj::PROBE_COUNT, d_keys, nullptr, n_rows, nullptr, nullptr, nullptr, nullptr, 0, d_n_matches, n_matches, (cudaStream_t)stream);
}

int dwj_timings(dwj_engine *e, dwj_timing *t) {
  if (!e || !t) return fail(DWJ_ERR_INVALID, "null argument");
  DeviceGuard g(e->cfg.device);
  *t = e->last_host;
  if (e->have_build) {
    CU(cudaEventSynchronize(e->ev_build[1]));
    CU(cudaEventElapsedTime(&t->build_ms, e->ev_build[0], e->ev_build[1]));
    if (e->build_rows) CU(cudaEventElapsedTime(&t->build_kernel_ms, e->ev_buildk[0], e->ev_buildk[1]));
  }
  if (e->have_probe) {
    CU(cudaEventSynchronize(e->ev_probe[1]));
    CU(cudaEventElapsedTime(&t->probe_ms, e->ev_probe[0], e->ev_probe[1]));
    CU(cudaEventElapsedTime(&t->probe_kernel_ms, e->ev_probek[0], e->ev_probek[1]));
  }
  if (e->have_part) {
    CU(cudaEventSynchronize(e->ev_part[1]));
    CU(cudaEventElapsedTime(&t->partition_ms, e->ev_part[0], e->ev_part[1]));
  }
  if (t->total_ms == 0.f) t->total_ms = t->build_ms + t->probe_ms;
  return DWJ_OK;
}

int dwj_partition(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, uint32_t n_parts,
                  void *d_out_keys, void *d_out_vals, uint64_t *d_offsets, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_parts == 0 || n_parts > (uint32_t)dwj::PART_MAX || (n_parts & (n_parts - 1)))
    return fail(DWJ_ERR_INVALID, "n_parts must be a power of two in [1, %d], got %u", dwj::PART_MAX, n_parts);
  if (!d_offsets || (n_rows && (!d_keys || !d_out_keys))) return fail(DWJ_ERR_INVALID, "null partition argument");
  if ((d_vals == nullptr) != (d_out_vals == nullptr)) return fail(DWJ_ERR_INVALID, "d_vals and d_out_vals must both be given or both be null");
  uint32_t lg = 0;
  while ((1u << lg) < n_parts) ++lg;
  DeviceGuard g(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaEventRecord(e->ev_part[0], s));
  const int rc = e->W == 4 ? partition_impl<4>(e, d_keys, d_vals, n_rows, lg, dwj::PART_BY_HASH, d_out_keys, d_out_vals, d_offsets, s)
                           : partition_impl<8>(e, d_keys, d_vals, n_rows, lg, dwj::PART_BY_HASH, d_out_keys, d_out_vals, d_offsets, s);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_part[1], s));
  e->have_part = true;
  return DWJ_OK;
}

int dwj_partition_hist(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_parts, uint64_t *d_counts, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_parts == 0 || n_parts > (uint32_t)dwj::PART_MAX || (n_parts & (n_parts - 1)))
    return fail(DWJ_ERR_INVALID, "n_parts must be a power of two in [1, %d], got %u", dwj::PART_MAX, n_parts);
  if (!d_counts || (n_rows && !d_keys)) return fail(DWJ_ERR_INVALID, "null partition argument");
  uint32_t lg = 0;
  while ((1u << lg) < n_parts) ++lg;
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? partition_hist_impl<4>(e, d_keys, n_rows, lg, dwj::PART_BY_HASH, 0, d_counts, (cudaStream_t)stream)
                   : partition_hist_impl<8>(e, d_keys, n_rows, lg, dwj::PART_BY_HASH, 0, d_counts, (cudaStream_t)stream);
}

// ---- exchange partition folded with the receiver's region grouping ------------------------------------------------------
namespace {
int xpart_bits(const dwj_engine *e, uint32_t n_ranks, uint32_t *rank_bits, uint32_t *fold_bits) {
  if (n_ranks == 0 || (n_ranks & (n_ranks - 1)) || n_ranks > (uint32_t)dwj::PART_MAX)
    return fail(DWJ_ERR_INVALID, "n_ranks must be a power of two in [1, %d], got %u", dwj::PART_MAX, n_ranks);
  uint32_t rb = 0, max_bits = 0;
  while ((1u << rb) < n_ranks) ++rb;
  while ((1u << (max_bits + 1)) <= (uint32_t)dwj::PART_MAX) ++max_bits;
  *rank_bits = rb;
  *fold_bits = rb + e->region_bits <= max_bits ? e->region_bits : 0;     // all region bits or none
  return DWJ_OK;
}
}  // namespace

uint32_t dwj_xpart_regions(const dwj_engine *e, uint32_t n_ranks) {
  uint32_t rb = 0, fb = 0;
  if (!e || xpart_bits(e, n_ranks, &rb, &fb)) return 0;
  return 1u << fb;
}

int dwj_xpart_hist(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_ranks, uint64_t *d_counts, void *stream) {
  if (int rc = check_engine(e)) return rc;
  uint32_t rb = 0, fb = 0;
  if (int rc = xpart_bits(e, n_ranks, &rb, &fb)) return rc;
  if (!d_counts || (n_rows && !d_keys)) return fail(DWJ_ERR_INVALID, "null partition argument");
  DeviceGuard g(e->cfg.device);
  const uint32_t mode = fb ? dwj::PART_BY_BOTH : dwj::PART_BY_HASH;
  return e->W == 4 ? partition_hist_impl<4>(e, d_keys, n_rows, rb + fb, mode, rb, d_counts, (cudaStream_t)stream)
                   : partition_hist_impl<8>(e, d_keys, n_rows, rb + fb, mode, rb, d_counts, (cudaStream_t)stream);
}

int dwj_xpart_scatter(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, uint32_t n_ranks, const uint64_t *start_rows,
                      void *d_out_keys, void *d_out_vals, void *stream) {
  if (int rc = check_engine(e)) return rc;
  uint32_t rb = 0, fb = 0;
  if (int rc = xpart_bits(e, n_ranks, &rb, &fb)) return rc;
  if (!start_rows || (n_rows && (!d_keys || !d_out_keys))) return fail(DWJ_ERR_INVALID, "null partition argument");
  if ((d_vals == nullptr) != (d_out_vals == nullptr)) return fail(DWJ_ERR_INVALID, "d_vals and d_out_vals must both be given or both be null");
  DeviceGuard g(e->cfg.device);
  const uint32_t mode = fb ? dwj::PART_BY_BOTH : dwj::PART_BY_HASH;
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaEventRecord(e->ev_part[0], s));
  const int rc = e->W == 4 ? partition_scatter_planned_impl<4>(e, d_keys, d_vals, n_rows, rb + fb, mode, rb, start_rows, d_out_keys, d_out_vals, s)
                           : partition_scatter_planned_impl<8>(e, d_keys, d_vals, n_rows, rb + fb, mode, rb, start_rows, d_out_keys, d_out_vals, s);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_part[1], s));
  e->have_part = true;
  return DWJ_OK;
}

int dwj_build_grouped(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, const uint64_t *d_region_offsets, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals)) return fail(DWJ_ERR_INVALID, "null build column");
  if (n_rows > e->cfg.max_build_rows && ((double)n_rows > 0.9 * (double)e->slots || e->runs))
    return fail(DWJ_ERR_CAPACITY, "%llu build rows exceed the table created for %llu", (unsigned long long)n_rows,
                (unsigned long long)e->cfg.max_build_rows);
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? build_impl<4>(e, d_keys, d_vals, n_rows, (cudaStream_t)stream, true, d_region_offsets)
                   : build_impl<8>(e, d_keys, d_vals, n_rows, (cudaStream_t)stream, true, d_region_offsets);
}

int dwj_probe_pairs_grouped(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_key, void *d_out_build_val,
                            void *d_out_probe_val, uint64_t capacity, uint64_t *d_n_matches, uint64_t *n_matches, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals)) return fail(DWJ_ERR_INVALID, "null probe column");
  if (capacity && (!d_out_build_val || !d_out_probe_val)) return fail(DWJ_ERR_INVALID, "null output column");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_PAIRS, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream, true)
                   : probe_impl<8>(e, dwj::PROBE_PAIRS, d_keys, d_vals, n_rows, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream, true);
}

namespace {
struct PendingSegs {       // scope guard: the segment list is only valid during one call
  dwj_engine *e;
  PendingSegs(dwj_engine *e_, uint32_t n, const void *const *keys, const void *const *vals, const uint64_t *rows, uint32_t per_region) : e(e_) {
    e->pending_segs = n;
    e->pending_seg_keys = keys;
    e->pending_seg_vals = vals;
    e->pending_seg_rows = rows;
    e->pending_segs_per_region = per_region;
  }
  ~PendingSegs() { e->pending_segs = 0; }
};
int check_segments(uint32_t n, const void *first, const uint64_t *rows, uint64_t *total) {
  if (n == 0 || n > (uint32_t)dwj::PART_MAX) return fail(DWJ_ERR_INVALID, "between 1 and %d segments, got %u", dwj::PART_MAX, n);
  if (!first || !rows) return fail(DWJ_ERR_INVALID, "null segment list");
  *total = 0;
  for (uint32_t i = 0; i < n; ++i) *total += rows[i];
  return DWJ_OK;
}
}  // namespace

int dwj_build_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals, const uint64_t *seg_rows,
                       uint32_t segments_per_region, void *stream) {
  if (int rc = check_engine(e)) return rc;
  uint64_t total = 0;
  if (int rc = check_segments(n_segments, seg_keys, seg_rows, &total)) return rc;
  if (!seg_vals) return fail(DWJ_ERR_INVALID, "null build payload segments");
  if (total > e->cfg.max_build_rows && ((double)total > 0.9 * (double)e->slots || e->runs))
    return fail(DWJ_ERR_CAPACITY, "%llu build rows exceed the table created for %llu", (unsigned long long)total,
                (unsigned long long)e->cfg.max_build_rows);
  DeviceGuard g(e->cfg.device);
  PendingSegs ps(e, total ? n_segments : 0, seg_keys, seg_vals, seg_rows, segments_per_region);
  return e->W == 4 ? build_impl<4>(e, nullptr, nullptr, total, (cudaStream_t)stream, true, nullptr)
                   : build_impl<8>(e, nullptr, nullptr, total, (cudaStream_t)stream, true, nullptr);
}

int dwj_probe_pairs_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals,
                             const uint64_t *seg_rows, void *d_out_key, void *d_out_build_val, void *d_out_probe_val, uint64_t capacity,
                             uint64_t *d_n_matches, uint64_t *n_matches, void *stream) {
  if (int rc = check_engine(e)) return rc;
  uint64_t total = 0;
  if (int rc = check_segments(n_segments, seg_keys, seg_rows, &total)) return rc;
  if (!seg_vals) return fail(DWJ_ERR_INVALID, "null probe payload segments");
  if (!(e->cfg.flags & DWJ_FLAG_UNIQUE_BUILD_KEYS))
    return fail(DWJ_ERR_INVALID, "segmented probe is implemented for DWJ_FLAG_UNIQUE_BUILD_KEYS engines");
  if (capacity && (!d_out_build_val || !d_out_probe_val)) return fail(DWJ_ERR_INVALID, "null output column");
  DeviceGuard g(e->cfg.device);
  PendingSegs ps(e, total ? n_segments : 0, seg_keys, seg_vals, seg_rows, 0);
  return e->W == 4 ? probe_impl<4>(e, dwj::PROBE_PAIRS, nullptr, nullptr, total, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream, true)
                   : probe_impl<8>(e, dwj::PROBE_PAIRS, nullptr, nullptr, total, d_out_key, d_out_build_val, d_out_probe_val, nullptr, capacity, d_n_matches, n_matches, (cudaStream_t)stream, true);
}

int dwj_region_scatter_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals,
                                const uint64_t *seg_rows, const uint64_t *start_rows, void *d_out_keys, void *d_out_vals, void *stream) {
  if (int rc = check_engine(e)) return rc;
  uint64_t total = 0;
  if (int rc = check_segments(n_segments, seg_keys, seg_rows, &total)) return rc;
  if (!start_rows || (total && !d_out_keys)) return fail(DWJ_ERR_INVALID, "null scatter argument");
  if ((seg_vals == nullptr) != (d_out_vals == nullptr)) return fail(DWJ_ERR_INVALID, "seg_vals and d_out_vals must both be given or both be null");
  if (!total) return DWJ_OK;
  DeviceGuard g(e->cfg.device);
  PendingSegs ps(e, n_segments, seg_keys, seg_vals, seg_rows, 0);
  return e->W == 4 ? region_scatter_segments_impl<4>(e, start_rows, d_out_keys, d_out_vals, seg_vals != nullptr, (cudaStream_t)stream)
                   : region_scatter_segments_impl<8>(e, start_rows, d_out_keys, d_out_vals, seg_vals != nullptr, (cudaStream_t)stream);
}

int dwj_xpart_hist2(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_ranks, uint64_t *d_counts, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_ranks == 0 || (n_ranks & (n_ranks - 1)) || n_ranks > 8) return fail(DWJ_ERR_INVALID, "n_ranks must be 1, 2, 4 or 8, got %u", n_ranks);
  if (!d_counts || (n_rows && !d_keys)) return fail(DWJ_ERR_INVALID, "null partition argument");
  uint32_t rb = 0;
  while ((1u << rb) < n_ranks) ++rb;
  DeviceGuard g(e->cfg.device);
  const uint32_t bits = rb + e->region_bits;
  const uint32_t mode = e->region_bits ? dwj::PART_BY_BOTH : dwj::PART_BY_HASH;
  return e->W == 4 ? partition_hist_impl<4>(e, d_keys, n_rows, bits, mode, rb, d_counts, (cudaStream_t)stream)
                   : partition_hist_impl<8>(e, d_keys, n_rows, bits, mode, rb, d_counts, (cudaStream_t)stream);
}

int dwj_filter_rows(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_keys, void *d_out_vals,
                    uint64_t *d_n_out, uint64_t *d_region_counts, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (!d_n_out || (n_rows && (!d_keys || !d_out_keys))) return fail(DWJ_ERR_INVALID, "null filter argument");
  if ((d_vals == nullptr) != (d_out_vals == nullptr)) return fail(DWJ_ERR_INVALID, "d_vals and d_out_vals must both be given or both be null");
  DeviceGuard g(e->cfg.device);
  return e->W == 4 ? filter_rows_impl<4>(e, d_keys, d_vals, n_rows, d_out_keys, d_out_vals, d_n_out, d_region_counts, (cudaStream_t)stream)
                   : filter_rows_impl<8>(e, d_keys, d_vals, n_rows, d_out_keys, d_out_vals, d_n_out, d_region_counts, (cudaStream_t)stream);
}

int dwj_aggregate_sum(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *stream) {
  if (int rc = check_engine(e)) return rc;
  if (n_rows && (!d_keys || !d_vals)) return fail(DWJ_ERR_INVALID, "null aggregation column");
  DeviceGuard g(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaEventRecord(e->ev_build[0], s));
  CU(cudaMemsetAsync(e->table, 0xFF, e->table_bytes, s));
  CU(cudaEventRecord(e->ev_buildk[0], s));
  if (n_rows) {
    const uint64_t tiles = (n_rows + (uint64_t)dwj::AGG_THREADS * dwj::AGG_ROWS - 1) / ((uint64_t)dwj::AGG_THREADS * dwj::AGG_ROWS);
    const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)e->prop.multiProcessorCount * 4);
    if (e->W == 4) {
      dwj::AggArgs<4> a{(const uint32_t *)d_keys, (const uint32_t *)d_vals, n_rows, e->table, e->buckets - 1, e->cfg.hash_seed};
      CU(launch(e, dwj::aggregate_kernel<4>, dim3(grid), dim3(dwj::AGG_THREADS), s, a, false));
    } else {
      dwj::AggArgs<8> a{(const uint64_t *)d_keys, (const uint64_t *)d_vals, n_rows, e->table, e->buckets - 1, e->cfg.hash_seed};
      CU(launch(e, dwj::aggregate_kernel<8>, dim3(grid), dim3(dwj::AGG_THREADS), s, a, false));
    }
  }
  CU(cudaEventRecord(e->ev_buildk[1], s));
  CU(cudaEventRecord(e->ev_build[1], s));
  e->launches_build = 2;
  e->have_build = true;
  e->build_rows = n_rows;
  e->built = true;
  e->csr = false;                       // slot payloads are sums, not run offsets
  e->pre_cleared = false;
  return DWJ_OK;
}

int dwj_clear_table(dwj_engine *e, void *stream) {
  if (int rc = check_engine(e)) return rc;
  DeviceGuard g(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!e->ev_cleared) CU(cudaEventCreateWithFlags(&e->ev_cleared, cudaEventDisableTiming));
  CU(cudaMemsetAsync(e->table, 0xFF, e->table_bytes, s));
  CU(cudaMemsetAsync(e->fill, 0, e->buckets * sizeof(unsigned int), s));
  CU(cudaEventRecord(e->ev_cleared, s));
  e->pre_cleared = true;
  return DWJ_OK;
}

int dwj_set_option(dwj_engine *e, int option, uint64_t value) {
  if (int rc = check_engine(e)) return rc;
  switch (option) {
  case DWJ_OPT_APPEND_OUTPUT: e->append_output = value != 0; return DWJ_OK;
  case DWJ_OPT_PASS_FILTER: {
    // value = rank_bits | pass_bits << 8 | pass_id << 16: the class of a key is the pass_bits bits of its partition
    // hash right below the rank_bits destination-rank bits; pass_bits == 0 switches the filter off.
    const uint32_t rank_bits = value & 0xff, pass_bits = (value >> 8) & 0xff, pass_id = (uint32_t)(value >> 16);
    if (pass_bits == 0) { e->filter = dwj::PassFilter{}; return DWJ_OK; }
    if (rank_bits + pass_bits > 16 || pass_id >= (1u << pass_bits)) return fail(DWJ_ERR_INVALID, "bad pass filter %llx", (unsigned long long)value);
    e->filter.shift = (e->W == 4 ? 32u : 64u) - rank_bits - pass_bits;
    e->filter.mask = (1u << pass_bits) - 1u;
    e->filter.want = pass_id;
    return DWJ_OK;
  }
  default: return fail(DWJ_ERR_INVALID, "unknown option %d", option);
  }
}

uint32_t dwj_partition_of(uint64_t key, int32_t key_bytes, uint32_t n_parts, uint64_t hash_seed) {
  uint32_t lg = 0;
  while ((1u << lg) < n_parts) ++lg;
  return key_bytes == 4 ? dwj::partition_of((uint32_t)key, lg, hash_seed) : dwj::partition_of((uint64_t)key, lg, hash_seed);
}

uint32_t dwj_region_of(uint64_t key, int32_t key_bytes, uint64_t buckets, uint32_t region_bits, uint64_t hash_seed) {
  uint32_t lgb = 0;
  while ((1ull << lgb) < buckets) ++lgb;
  if (!region_bits || region_bits > lgb) return 0;
  const uint64_t h = key_bytes == 4 ? dwj::slot_hash((uint32_t)key, hash_seed) : dwj::slot_hash((uint64_t)key, hash_seed);
  return (uint32_t)((h & (buckets - 1)) >> (lgb - region_bits));
}

// ---- host-buffer join ---------------------------------------------------------------------------------
// Three streams (H2D, compute, D2H) and HOST_STAGES staging slots: the copy engines run continuously while the
// host only ever waits for a chunk that is two behind the one it just enqueued.  PCIe is full duplex, so the H2D of
// chunk c+1.. overlaps the probe of chunk c and the D2H of chunk c-1...
static int join_host_impl(dwj_engine *e, const void *build_keys, const void *build_vals, uint64_t n_build, const void *probe_keys,
                          const void *probe_vals, uint64_t n_probe, int out_mode, void *out_key, void *out_build_val,
                          void *out_probe_val, uint64_t out_capacity, uint64_t *n_out, dwj_timing *timing);

int dwj_join_host(dwj_engine *e, const void *build_keys, const void *build_vals, uint64_t n_build, const void *probe_keys,
                  const void *probe_vals, uint64_t n_probe, int out_mode, void *out_key, void *out_build_val,
                  void *out_probe_val, uint64_t out_capacity, uint64_t *n_out, dwj_timing *timing) {
  const int rc = join_host_impl(e, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_mode, out_key, out_build_val,
                                out_probe_val, out_capacity, n_out, timing);
  if (rc && e) {
    // An error exit may leave copies queued that target the caller's buffers: nothing of this call is in flight once
    // it has returned.  The message of the failure is kept.
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    DeviceGuard g(e->cfg.device);
    for (auto &st : e->hs)
      if (st) cudaStreamSynchronize(st);
    cudaGetLastError();
    memcpy(g_err, keep, sizeof(keep));
  }
  return rc;
}

static int join_host_impl(dwj_engine *e, const void *build_keys, const void *build_vals, uint64_t n_build, const void *probe_keys,
                          const void *probe_vals, uint64_t n_probe, int out_mode, void *out_key, void *out_build_val,
                          void *out_probe_val, uint64_t out_capacity, uint64_t *n_out, dwj_timing *timing) {
  if (int rc = check_engine(e)) return rc;
  if (out_mode < DWJ_OUT_ALIGNED || out_mode > DWJ_OUT_COUNT) return fail(DWJ_ERR_INVALID, "bad out_mode %d", out_mode);
  if ((n_build && (!build_keys || !build_vals)) || (n_probe && (!probe_keys || !probe_vals)))
    return fail(DWJ_ERR_INVALID, "null input column");
  if (out_mode == DWJ_OUT_ALIGNED && n_probe && (!out_key || !out_build_val || !out_probe_val))
    return fail(DWJ_ERR_INVALID, "ALIGNED output needs all three columns");
  if (out_mode == DWJ_OUT_PAIRS && out_capacity && (!out_build_val || !out_probe_val))
    return fail(DWJ_ERR_INVALID, "PAIRS output needs the two payload columns");
  if (out_mode != DWJ_OUT_ALIGNED && !n_out) return fail(DWJ_ERR_INVALID, "n_out is required");
  if (n_build > e->cfg.max_build_rows && ((double)n_build > 0.9 * (double)e->slots || e->runs))
    return fail(DWJ_ERR_CAPACITY, "%llu build rows exceed the table created for %llu", (unsigned long long)n_build,
                (unsigned long long)e->cfg.max_build_rows);
  DeviceGuard g(e->cfg.device);
  constexpr int NS = HOST_STAGES, LAG = 2;
  const uint64_t W = (uint64_t)e->W;
  const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint64_t>(n_probe, 1), HOST_CHUNK_BYTES / W));
  // staging: build k,v | NS stages x (probe k,v + out k,b,p) | NS counters
  const uint64_t build_b = ((n_build * W + 255) / 256) * 256, chunk_b = ((chunk_rows * W + 255) / 256) * 256;
  const uint64_t need = 2 * build_b + (uint64_t)NS * 5 * chunk_b + 64 * NS;
  if (need > e->stage_bytes) {
    if (e->stage) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->stage)); e->stage = nullptr; e->stage_bytes = 0; }
    CU(cudaMalloc(&e->stage, need));
    e->stage_bytes = need;
  }
  for (auto &st : e->hs)
    if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto &ev : e->hev)
    if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : e->htime)
    if (!ev) CU(cudaEventCreate(&ev));
  if (!e->h_counts) CU(cudaHostAlloc((void **)&e->h_counts, NS * sizeof(unsigned long long), cudaHostAllocDefault));
  cudaStream_t s_in = e->hs[0], s_comp = e->hs[1], s_out = e->hs[2];
  cudaEvent_t *ev_in = e->hev, *ev_done = e->hev + NS, *ev_free = e->hev + 2 * NS;
  char *base = (char *)e->stage;
  char *d_bk = base, *d_bv = base + build_b;
  char *stage_base = base + 2 * build_b;
  unsigned long long *d_cnt = (unsigned long long *)(stage_base + (uint64_t)NS * 5 * chunk_b);
  auto stage_ptr = [&](int st, int col) { return stage_base + (uint64_t)(st * 5 + col) * chunk_b; };

  // htime: 0 start, 1 build h2d done, 2 build done, 3 end
  CU(cudaEventRecord(e->htime[0], s_in));
  if (n_build) {
    CU(cudaMemcpyAsync(d_bk, build_keys, n_build * W, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(d_bv, build_vals, n_build * W, cudaMemcpyHostToDevice, s_in));
  }
  CU(cudaEventRecord(e->htime[1], s_in));
  CU(cudaStreamWaitEvent(s_comp, e->htime[1], 0));
  if (int rc = dwj_build(e, d_bk, d_bv, n_build, s_comp)) return rc;
  CU(cudaEventRecord(e->htime[2], s_comp));

  uint64_t produced = 0;       // rows written to the host outputs so far (PAIRS) / matches (COUNT)
  bool overflow = false;

  if (out_mode == DWJ_OUT_ALIGNED) {
    const uint64_t n_chunks = (n_probe + chunk_rows - 1) / chunk_rows;
    for (uint64_t c = 0; c < n_chunks; ++c) {
      const int st = (int)(c % NS);
      const uint64_t row0 = c * chunk_rows, rows = std::min<uint64_t>(chunk_rows, n_probe - row0);
      if (c >= (uint64_t)NS) CU(cudaStreamWaitEvent(s_in, ev_free[st], 0));      // the slot's previous chunk has left the device
      CU(cudaMemcpyAsync(stage_ptr(st, 0), (const char *)probe_keys + row0 * W, rows * W, cudaMemcpyHostToDevice, s_in));
      CU(cudaMemcpyAsync(stage_ptr(st, 1), (const char *)probe_vals + row0 * W, rows * W, cudaMemcpyHostToDevice, s_in));
      CU(cudaEventRecord(ev_in[st], s_in));
      CU(cudaStreamWaitEvent(s_comp, ev_in[st], 0));
      if (int rc = dwj_probe_aligned(e, stage_ptr(st, 0), stage_ptr(st, 1), rows, stage_ptr(st, 2), stage_ptr(st, 3), stage_ptr(st, 4), s_comp)) return rc;
      CU(cudaEventRecord(ev_done[st], s_comp));
      CU(cudaStreamWaitEvent(s_out, ev_done[st], 0));
      CU(cudaMemcpyAsync((char *)out_key + row0 * W, stage_ptr(st, 2), rows * W, cudaMemcpyDeviceToHost, s_out));
      CU(cudaMemcpyAsync((char *)out_build_val + row0 * W, stage_ptr(st, 3), rows * W, cudaMemcpyDeviceToHost, s_out));
      CU(cudaMemcpyAsync((char *)out_probe_val + row0 * W, stage_ptr(st, 4), rows * W, cudaMemcpyDeviceToHost, s_out));
      CU(cudaEventRecord(ev_free[st], s_out));
    }
  } else {
    // PAIRS / COUNT.  A staging slot holds `chunk_rows` result rows.  With duplicate build keys a piece of the probe
    // relation may produce more: the probe kernel then keeps what fits and still reports the exact count, so the piece is
    // simply probed AGAIN in smaller pieces -- and the piece size stays reduced (it adapts to the multiplicity of the
    // data within a piece or two).  In flight: up to LAG + 1 pieces, each {first row, rows, slot}.
    struct Piece { uint64_t row0, rows; int st; };
    Piece inflight[NS];
    uint32_t head = 0, count = 0;                 // ring of pieces enqueued and not yet drained
    uint64_t next_row = 0, piece_rows = chunk_rows, issued = 0;
    auto enqueue = [&]() -> int {
      const int st = (int)(issued % NS);
      const uint64_t row0 = next_row, rows = std::min<uint64_t>(piece_rows, n_probe - row0);
      if (issued >= (uint64_t)NS) CU(cudaStreamWaitEvent(s_in, ev_free[st], 0));
      CU(cudaMemcpyAsync(stage_ptr(st, 0), (const char *)probe_keys + row0 * W, rows * W, cudaMemcpyHostToDevice, s_in));
      if (out_mode == DWJ_OUT_PAIRS) CU(cudaMemcpyAsync(stage_ptr(st, 1), (const char *)probe_vals + row0 * W, rows * W, cudaMemcpyHostToDevice, s_in));
      CU(cudaEventRecord(ev_in[st], s_in));
      CU(cudaStreamWaitEvent(s_comp, ev_in[st], 0));
      int rc;
      if (out_mode == DWJ_OUT_PAIRS)
        rc = dwj_probe_pairs(e, stage_ptr(st, 0), stage_ptr(st, 1), rows, out_key ? stage_ptr(st, 2) : nullptr, stage_ptr(st, 3),
                             stage_ptr(st, 4), chunk_rows, (uint64_t *)(d_cnt + st * 8), nullptr, s_comp);
      else
        rc = dwj_probe_count(e, stage_ptr(st, 0), rows, (uint64_t *)(d_cnt + st * 8), nullptr, s_comp);
      if (rc) return rc;
      CU(cudaMemcpyAsync(e->h_counts + st, d_cnt + st * 8, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_comp));
      CU(cudaEventRecord(ev_done[st], s_comp));
      inflight[(head + count++) % NS] = Piece{row0, rows, st};
      next_row = row0 + rows;
      ++issued;
      return DWJ_OK;
    };
    // Finish the oldest piece: wait for its probe, read its match count, enqueue its D2H at the running offset -- or, if
    // it produced more rows than a slot holds, drop everything in flight and restart from it with smaller pieces.
    auto drain = [&]() -> int {
      const Piece p = inflight[head];
      CU(cudaEventSynchronize(ev_done[p.st]));
      const unsigned long long cnt = e->h_counts[p.st];
      if (out_mode == DWJ_OUT_PAIRS && cnt > chunk_rows) {
        if (p.rows == 1)
          return fail(DWJ_ERR_OVERFLOW, "one probe row has %llu matches, more than the %llu rows dwj_join_host stages at a time -- "
                      "use dwj_probe_pairs with device buffers", cnt, (unsigned long long)chunk_rows);
        CU(cudaStreamSynchronize(s_comp));          // the younger pieces are discarded with this one
        for (uint32_t i = 0; i < count; ++i) CU(cudaEventRecord(ev_free[inflight[(head + i) % NS].st], s_comp));
        count = 0;
        next_row = p.row0;
        piece_rows = std::max<uint64_t>(1, p.rows / (2 * ((cnt + chunk_rows - 1) / chunk_rows)));
        return DWJ_OK;
      }
      if (out_mode == DWJ_OUT_PAIRS) {
        const uint64_t room = produced < out_capacity ? out_capacity - produced : 0;
        const uint64_t take = std::min<uint64_t>(cnt, room);
        if (take < cnt) overflow = true;
        CU(cudaStreamWaitEvent(s_out, ev_done[p.st], 0));
        if (take) {
          if (out_key) CU(cudaMemcpyAsync((char *)out_key + produced * W, stage_ptr(p.st, 2), take * W, cudaMemcpyDeviceToHost, s_out));
          CU(cudaMemcpyAsync((char *)out_build_val + produced * W, stage_ptr(p.st, 3), take * W, cudaMemcpyDeviceToHost, s_out));
          CU(cudaMemcpyAsync((char *)out_probe_val + produced * W, stage_ptr(p.st, 4), take * W, cudaMemcpyDeviceToHost, s_out));
        }
      }
      CU(cudaEventRecord(ev_free[p.st], s_out));
      produced += cnt;
      head = (head + 1) % NS;
      --count;
      return DWJ_OK;
    };
    while (next_row < n_probe || count) {
      if (next_row < n_probe && count <= (uint32_t)LAG) {       // LAG pieces stay queued behind the one the host waits for
        if (int rc = enqueue()) return rc;
        if (count <= (uint32_t)LAG && next_row < n_probe) continue;
      }
      if (int rc = drain()) return rc;
    }
  }
  CU(cudaStreamSynchronize(s_comp));
  CU(cudaStreamSynchronize(s_out));
  CU(cudaEventRecord(e->htime[3], s_out));
  CU(cudaEventSynchronize(e->htime[3]));

  dwj_timing t{};
  CU(cudaEventElapsedTime(&t.h2d_ms, e->htime[0], e->htime[1]));
  CU(cudaEventElapsedTime(&t.build_ms, e->htime[1], e->htime[2]));
  CU(cudaEventElapsedTime(&t.total_ms, e->htime[0], e->htime[3]));
  t.probe_ms = std::max(0.f, t.total_ms - t.h2d_ms - t.build_ms);   // the streamed probe phase, copies overlapped
  e->last_host = t;
  if (timing) *timing = t;
  if (n_out) *n_out = out_mode == DWJ_OUT_ALIGNED ? n_probe : produced;
  if (overflow) return fail(DWJ_ERR_OVERFLOW, "join produced %llu rows, output capacity is %llu", (unsigned long long)produced, (unsigned long long)out_capacity);
  return DWJ_OK;
}

}  // extern "C"
