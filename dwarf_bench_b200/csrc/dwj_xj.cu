// dwj_xj.cu -- the multi-GPU join behind the C ABI (include/dwj.h: dwj_xj_*, dwj_mg_*).
//
// No reference counterpart: the reference runs one queue on one device (join/join.cpp:23-24).
//
// One dwj_xj per GPU ("rank").  Equal keys must meet on one GPU, so every relation needs one exchange step.  Here the
// exchange is a PULL over peer memory, fused into the kernels that consume the rows:
//
//   sender    one partition pass per batch (the build relation, then the probe relation in chunks) groups the rank's
//             rows by destination rank -- and, when ranks x table regions <= 512, by the table region of the
//             destination's table as well -- inside its OWN memory (the "send slots" of a peer-mapped block), then
//             raises a flag in every peer's memory;
//   receiver  waits for the flags and runs its build / probe kernels with a segment list that points INTO THE SENDERS'
//             SLOTS: one segment per (table region, source rank), walked region by region.  The kernels read their rows
//             over NVLink exactly once, at the pace they consume them; there is no receive buffer, no copy-engine job
//             and no collective on the data path.  When the regions do not fit the sender's pass (ranks x regions > 512)
//             or the build keys are not unique, the receiver's own region scatter is the pulling kernel instead
//             (dwj_region_scatter_segments) and the local build / probe run on its region-grouped output;
//   counts    every sender counts its rows per (destination, table region) before anything moves (dwj_xpart_hist2) and
//             writes the counts into the peers' memory; one host synchronisation per step reads them and plans every
//             offset of the step.  Flags and counts live in the same peer-mapped block: no NCCL / MPI call anywhere.
//
// Synchronisation is by monotonic sequence numbers in peer memory (st.release.sys / ld.acquire.sys from one-CTA
// kernels on the streams): ready[slot][src] = "src has filled this slot for the n-th time", done[slot][dst] = "dst has
// pulled its rows of the n-th filling".  The waiting kernels give up after 20 s (DWJ_XJ_TIMEOUT_MS) and raise an error word the
// host reads at its next synchronisation, so a lost peer fails the join instead of hanging the GPU.
//
// The same object serves one process per GPU (torchrun: the block is torch symmetric memory, bench.py) and one process
// for all GPUs (dwj_mg_*: cudaMalloc + cudaDeviceEnablePeerAccess, one host thread per GPU -- the C++ host framework's
// `--gpus N`).  world == 1 degenerates to the single-GPU join with a chunked probe; `passes` > 1 runs the join once per
// key class (DWJ_OPT_PASS_FILTER) for working sets larger than the GPUs' memory.
#include "../../include/dwj.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

extern "C" void dwj_internal_set_error(const char *msg);      // dwj_api.cu (not exported)

namespace {

int xfail(int code, const char *fmt, ...) {          // one error buffer per thread for the whole library (dwj_last_error)
  char msg[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  dwj_internal_set_error(msg);
  return code;
}
#define XCU(call)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return xfail(_e == cudaErrorMemoryAllocation ? DWJ_ERR_OOM : DWJ_ERR_CUDA, "%s: %s (%s:%d)",   \
                   #call, cudaGetErrorString(_e), __FILE__, __LINE__);                               \
  } while (0)
#define XRC(call)                                                                                    \
  do {                                                                                               \
    if (int _rc = (call)) return _rc;                                                                \
  } while (0)

constexpr uint32_t MAX_WORLD = 8, MAX_SLOTS = 8;                 // slot 0: build relation; 1..: ring of probe chunks
constexpr uint32_t FLAG_CNT = 0, FLAG_ERR = 8, FLAG_READY = 16, FLAG_DONE = FLAG_READY + MAX_SLOTS * MAX_WORLD;
constexpr uint64_t CTRL_FLAG_BYTES = 4096, HDR_WORDS = 8;


struct Peers {
  unsigned long long *ctrl[MAX_WORLD];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// flag[index + me] = value in every peer's control block (this rank's earlier work on the stream is complete: the
// kernel boundary orders it, the fence + release publish it system-wide).
__global__ void xj_signal_kernel(Peers peers, uint32_t world, uint32_t me, uint32_t index, unsigned long long value) {
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.ctrl[threadIdx.x] + index + me, value);
  }
}
// Spin until flag[index + r] >= value for every rank r (local memory: the peers write, this GPU polls its own L2).
__global__ void xj_wait_kernel(unsigned long long *ctrl, uint32_t world, uint32_t index, unsigned long long value, unsigned long long timeout_ns) {
  if (threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(ctrl + index + threadIdx.x) < value) {
      if (global_ns() - t0 > timeout_ns) {
        atomicMax(ctrl + FLAG_ERR, (unsigned long long)(1 + index + threadIdx.x));
        break;
      }
      __nanosleep(200);
    }
  }
}
// ---- the transfer: a copy kernel that keeps every NVLink busy ------------------------------------------------------------
// Measured (profiles/r2_exchange.md): build / probe / scatter kernels that read their rows straight out of peer memory
// ("fused pull") move 215-285 GB/s per GPU -- a tile loads, then computes, and while it computes its SM has nothing in
// flight on a link with ~3 us of latency.  A kernel that does nothing but load and store keeps the SMs' miss queues full:
// every CTA streams 32 KB blocks (uint4 per lane, 8 deep) and the blocks are dealt ROUND-ROBIN over the runs -- one run
// per (source rank, column) -- so all peers are read at once and no source's egress is shared by several readers.
struct CopyRun {
  const char *src;          // first byte (any 4-byte alignment); the 16-byte aligned body starts `head` bytes in
  char *dst;
  unsigned long long bytes;
  unsigned int head;        // bytes before src becomes 16-byte aligned (< 16, <= bytes)
  unsigned int blocks;      // 32 KB blocks of the body (the last one partial)
};
constexpr int COPY_THREADS = 256, COPY_UNROLL = 8;
constexpr unsigned long long COPY_BLOCK = (unsigned long long)COPY_THREADS * COPY_UNROLL * 16;
__device__ __forceinline__ uint4 ld_stream16(const void *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(COPY_THREADS) xj_copy_runs_kernel(const CopyRun *runs, uint32_t n_runs, uint32_t max_blocks) {
  const unsigned long long total = (unsigned long long)max_blocks * n_runs;
  for (unsigned long long t = blockIdx.x; t < total; t += gridDim.x) {
    const CopyRun r = runs[t % n_runs];
    const unsigned int lb = (unsigned int)(t / n_runs);
    if (lb >= r.blocks) continue;
    if (lb == 0)                                       // the few bytes before the aligned body
      for (unsigned int i = threadIdx.x * 4; i < r.head; i += COPY_THREADS * 4) *(unsigned int *)(r.dst + i) = *(const unsigned int *)(r.src + i);
    const unsigned long long off = r.head + (unsigned long long)lb * COPY_BLOCK, len = min(COPY_BLOCK, r.bytes - off);
    const char *src = r.src + off;
    char *dst = r.dst + off;
    const bool dst16 = ((unsigned long long)dst & 15) == 0;
    if (len == COPY_BLOCK) {
      uint4 v[COPY_UNROLL];
#pragma unroll
      for (int j = 0; j < COPY_UNROLL; ++j) v[j] = ld_stream16(src + ((size_t)j * COPY_THREADS + threadIdx.x) * 16);
#pragma unroll
      for (int j = 0; j < COPY_UNROLL; ++j) {
        char *d = dst + ((size_t)j * COPY_THREADS + threadIdx.x) * 16;
        if (dst16) *(uint4 *)d = v[j];
        else { unsigned int *q = (unsigned int *)d; q[0] = v[j].x; q[1] = v[j].y; q[2] = v[j].z; q[3] = v[j].w; }
      }
    } else {
      const unsigned long long vec = len / 16;
      for (unsigned long long i = threadIdx.x; i < vec; i += COPY_THREADS) {
        const uint4 v = ld_stream16(src + i * 16);
        unsigned int *q = (unsigned int *)(dst + i * 16);
        q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
      }
      for (unsigned long long i = vec * 16 + threadIdx.x * 4; i < len; i += COPY_THREADS * 4) *(unsigned int *)(dst + i) = *(const unsigned int *)(src + i);
    }
  }
}

__global__ void xj_stage_kernel(unsigned long long *dst, const unsigned long long *pinned_src, uint32_t words) {
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = pinned_src[i];
}
// Counts of this rank for all batches, counts[batch][dst][region], go to every peer p as
//   gather_p[src = me][0..HDR)                         header: config hash, step
//   gather_p[me][HDR + batch * (world + regions) + d]   rows this rank sends to rank d           (d < world)
//   gather_p[me][... + world + g]                       rows it sends to p for p's table region g
// followed by the count flag.  One CTA per peer.
__global__ void xj_publish_kernel(Peers peers, uint32_t world, uint32_t me, const unsigned long long *counts, uint32_t batches,
                                  uint32_t regions, uint64_t gather_word0, uint64_t src_stride_words, unsigned long long cfg_hash,
                                  unsigned long long step) {
  const uint32_t p = blockIdx.x;
  unsigned long long *out = peers.ctrl[p] + gather_word0 + (uint64_t)me * src_stride_words;
  const uint32_t per_batch = world + regions;
  for (uint32_t i = threadIdx.x; i < batches * per_batch; i += blockDim.x) {
    const uint32_t b = i / per_batch, j = i % per_batch;
    unsigned long long v = 0;
    if (j < world) {
      const unsigned long long *row = counts + ((uint64_t)b * world + j) * regions;
      for (uint32_t g = 0; g < regions; ++g) v += row[g];
    } else {
      v = counts[((uint64_t)b * world + p) * regions + (j - world)];
    }
    out[HDR_WORDS + i] = v;
  }
  if (threadIdx.x == 0) {
    out[0] = cfg_hash;
    out[1] = step;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(peers.ctrl[p] + FLAG_CNT + me, step);
  }
}

uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
uint32_t log2u(uint32_t v) {
  uint32_t l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

}  // namespace

struct dwj_xj {
  dwj_engine *e = nullptr;
  dwj_xj_config cfg{};
  dwj_info info{};
  int device = 0;
  uint32_t W = 4, world = 1, me = 0, regions = 1, fold_regions = 1, chunks = 1, ring = 1, batches = 2, passes = 1;
  bool unique = false, direct = false;       // direct: build / probe kernels pull; else the region scatter pulls
  uint64_t chunk_rows = 0, slot_rows[2] = {0, 0}, rows_total = 0, ctrl_bytes = 0, gather_word0 = 0, src_stride_words = 0;
  uint64_t cap_recv_chunk = 0;
  unsigned long long cfg_hash = 0;
  Peers peers{};
  char *keys_base[MAX_WORLD]{}, *vals_base[MAX_WORLD]{};
  unsigned long long *ctrl = nullptr;        // = peers.ctrl[me]
  cudaStream_t s_part = nullptr, s_join = nullptr, s_pull = nullptr;
  cudaEvent_t ev_in = nullptr, ev_part_end = nullptr, ev_join_end = nullptr, ev_pulled[2]{}, ev_consumed[2]{}, ev_build_pulled = nullptr;
  cudaEvent_t ev_t[7]{};                     // timeline of the last step: start, counts, scattered, built, done, build pulled, last chunk pulled
  unsigned long long *d_counts = nullptr, *h_counts = nullptr, *h_gather = nullptr;
  unsigned long long *d_region_off = nullptr, *h_region_off = nullptr;   // ring of 4
  uint32_t region_off_calls = 0;
  cudaEvent_t ev_region_off[4]{};
  void *local_build = nullptr, *local_probe[2] = {nullptr, nullptr};     // fused scatter pull: region-grouped landing buffers
  uint64_t local_build_rows = 0;
  // copy pull (world > 1): the rows of a batch are first copied out of the senders' slots by xj_copy_runs_kernel.
  //   direct : arena_a = [build landing | probe landing 0 | probe landing 1], consumed in place through segment lists
  //   scatter: arena_a holds the raw rows (build, then the two probe halves), arena_b their region-grouped form
  bool copy_pull = false;
  void *arena_a = nullptr, *arena_b = nullptr;
  uint64_t arena_rows = 0, cap_build = 0;
  CopyRun *d_runs = nullptr, *h_runs = nullptr;   // ring of RUN_SLOTS x 2 * MAX_WORLD
  uint32_t run_calls = 0;
  cudaEvent_t ev_runs[32]{}, ev_scat_b = nullptr, ev_scat[2]{};
  int sm_count = 148;
  unsigned long long step = 0, slot_use[MAX_SLOTS]{};
  unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;      // DWJ_XJ_TIMEOUT_MS
  void (*host_barrier)(void *) = nullptr;      // dwj_mg: aligns the rank threads' enqueue order (ranks sharing one process)
  void *host_barrier_ctx = nullptr;
  dwj_xj_timing last{};
  uint64_t last_remote_bytes = 0;
};

namespace {

uint64_t slot_base_row(const dwj_xj *x, uint32_t slot) {
  return slot == 0 ? 0 : x->slot_rows[0] + (uint64_t)(slot - 1) * x->slot_rows[1];
}

int xj_layout(const dwj_engine *e, const dwj_xj_config *cfg, dwj_xj *x) {
  if (!e || !cfg) return xfail(DWJ_ERR_INVALID, "null argument");
  if (cfg->world < 1 || cfg->world > (int)MAX_WORLD || (cfg->world & (cfg->world - 1)))
    return xfail(DWJ_ERR_INVALID, "world must be 1, 2, 4 or 8, got %d", cfg->world);
  if (cfg->rank < 0 || cfg->rank >= cfg->world) return xfail(DWJ_ERR_INVALID, "rank %d out of range", cfg->rank);
  const uint32_t passes = cfg->passes ? cfg->passes : 1;
  if (passes & (passes - 1) || passes > 16) return xfail(DWJ_ERR_INVALID, "passes must be a power of two <= 16, got %u", passes);
  dwj_info info{};
  if (int rc = dwj_get_info(e, &info)) return rc;
  x->info = info;
  x->cfg = *cfg;
  x->world = (uint32_t)cfg->world;
  x->me = (uint32_t)cfg->rank;
  x->passes = passes;
  x->W = info.slot_bytes / 2;
  x->regions = info.radix_parts;
  x->fold_regions = dwj_xpart_regions(e, x->world);
  // Chunking the probe relation lets the senders partition piece c+1 while the receivers pull piece c -- but every piece
  // walks all table regions, so a table that does not stay in L2 is read from HBM once PER PIECE.  Small tables: pieces
  // of 2^26 rows.  Large tables: two pieces across GPUs (one overlap point), one piece on a single GPU (nothing to
  // overlap with; measured on the 2^31 x 2^31 join: 32 pieces 266 ms per step).
  const uint64_t probe_rows = std::max<uint64_t>(cfg->max_probe_rows, 1);
  if (cfg->chunk_rows) x->chunk_rows = cfg->chunk_rows;
  else if (info.table_bytes <= (512ull << 20)) x->chunk_rows = 1ull << 26;
  else x->chunk_rows = cfg->world == 1 ? probe_rows : (probe_rows + 1) / 2;
  x->chunks = (uint32_t)std::max<uint64_t>(1, (cfg->max_probe_rows + x->chunk_rows - 1) / x->chunk_rows);
  if (x->chunks > 255) return xfail(DWJ_ERR_INVALID, "%u probe chunks: raise chunk_rows", x->chunks);
  x->ring = std::min<uint32_t>(x->chunks, 3);
  x->batches = 1 + x->chunks;
  // A slot holds what one batch of this rank can be after the pass filter (an even split plus head-room: the classes
  // are hash bits; the plan checks the real counts and fails loudly).
  auto slot_cap = [&](uint64_t rows) { return round_up(passes == 1 ? rows : (uint64_t)((double)rows / passes * 1.05) + 4096, 64); };
  x->slot_rows[0] = std::max<uint64_t>(slot_cap(cfg->max_build_rows), 64);
  x->slot_rows[1] = std::max<uint64_t>(slot_cap(std::min<uint64_t>(x->chunk_rows, std::max<uint64_t>(cfg->max_probe_rows, 1))), 64);
  x->rows_total = x->slot_rows[0] + (uint64_t)x->ring * x->slot_rows[1];
  x->src_stride_words = HDR_WORDS + (uint64_t)x->batches * (x->world + x->regions);
  x->gather_word0 = CTRL_FLAG_BYTES / 8;
  x->ctrl_bytes = round_up(CTRL_FLAG_BYTES + 2 * x->world * x->src_stride_words * 8, 4096);
  return DWJ_OK;
}

int signal_all(dwj_xj *x, uint32_t index, unsigned long long value, cudaStream_t s) {
  xj_signal_kernel<<<1, 32, 0, s>>>(x->peers, x->world, x->me, index, value);
  XCU(cudaGetLastError());
  return DWJ_OK;
}
int wait_all(dwj_xj *x, uint32_t index, unsigned long long value, cudaStream_t s) {
  xj_wait_kernel<<<1, 32, 0, s>>>(x->ctrl, x->world, index, value, x->timeout_ns);
  XCU(cudaGetLastError());
  return DWJ_OK;
}

}  // namespace

extern "C" {

// ---- the plan of one batch, as pure host functions (exported so the layout logic is testable without a GPU) -------------
// Sender: mine[dst][region] = rows of this batch bound for rank dst / its table region.  The slot is laid out
// destination-major and, when the regions are folded into the sender's pass (fold_regions == regions), region-minor:
// start[] = first row of every partition of that pass (world * fold_regions entries).
int dwj_xj_plan_send(uint32_t world, uint32_t regions, uint32_t fold_regions, uint64_t slot_base_row, const uint64_t *mine, uint64_t *start) {
  if (!mine || !start || !world || !regions || (fold_regions != regions && fold_regions != 1)) return xfail(DWJ_ERR_INVALID, "bad plan argument");
  uint64_t run = slot_base_row;
  for (uint32_t d = 0; d < world; ++d) {
    if (fold_regions == regions) {
      for (uint32_t g = 0; g < regions; ++g) { start[(size_t)d * regions + g] = run; run += mine[(size_t)d * regions + g]; }
    } else {                                    // regions not folded into this pass: one run per destination
      start[d] = run;
      for (uint32_t g = 0; g < regions; ++g) run += mine[(size_t)d * regions + g];
    }
  }
  return DWJ_OK;
}
// Receiver `me`: tot[src][dst] = rows src sends to dst (so src's block for me starts sum(tot[src][0..me)) rows into its
// slot), reg[src][region] = of those for me, the rows of my table region.  Every list visits the sources in ROTATED order
// me, me+1, ... (mod world): if all ranks started with source 0, eight readers would share one GPU's NVLink egress while
// seven links idle (measured: 215 GB/s per GPU instead of ~700).
//   direct != 0: one segment per (region, source) in walking order -- region-major, rotated source-minor.
//   direct == 0: every source's block for me is cut into up to `pieces` pieces (of at least 2^16 rows) and the pieces are
//                dealt round-robin over the sources, so the pulling scatter reads from all peers at once;
//                region_start[g] = first row of region g in the landing buffer it fills.
// seg_src[i] = source rank of segment i; *n_segments = segments written (at most max(regions, pieces) * world);
// *total = rows this rank receives.
int dwj_xj_plan_recv(uint32_t world, uint32_t me, uint32_t regions, uint64_t slot_base_row, const uint64_t *tot, const uint64_t *reg, int direct,
                     uint32_t pieces, uint64_t *seg_first_row, uint64_t *seg_rows, uint32_t *seg_src, uint64_t *region_start, uint64_t *total,
                     uint32_t *n_segments) {
  if (!tot || !reg || !seg_first_row || !seg_rows || !seg_src || !total || !n_segments || !world || !regions || me >= world)
    return xfail(DWJ_ERR_INVALID, "bad plan argument");
  uint64_t sum = 0;
  uint32_t n = 0;
  std::vector<uint64_t> block(world);                       // first row of src's block for me inside src's slot
  for (uint32_t s = 0; s < world; ++s) {
    uint64_t r = slot_base_row;
    for (uint32_t d = 0; d < me; ++d) r += tot[(size_t)s * world + d];
    block[s] = r;
    sum += tot[(size_t)s * world + me];
  }
  if (direct) {
    std::vector<uint64_t> at(block);
    for (uint32_t g = 0; g < regions; ++g)
      for (uint32_t i = 0; i < world; ++i) {
        const uint32_t s = (me + i) % world;
        seg_first_row[n] = at[s];
        seg_rows[n] = reg[(size_t)s * regions + g];
        seg_src[n] = s;
        at[s] += reg[(size_t)s * regions + g];
        ++n;
      }
  } else {
    uint64_t most = 0;
    for (uint32_t s = 0; s < world; ++s) most = std::max<uint64_t>(most, tot[(size_t)s * world + me]);
    const uint32_t k = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(pieces ? pieces : 1, most >> 16));
    for (uint32_t p = 0; p < k; ++p)
      for (uint32_t i = 0; i < world; ++i) {
        const uint32_t s = (me + i) % world;
        const uint64_t rows = tot[(size_t)s * world + me], lo = rows * p / k, hi = rows * (p + 1) / k;
        seg_first_row[n] = block[s] + lo;
        seg_rows[n] = hi - lo;
        seg_src[n] = s;
        ++n;
      }
  }
  if (region_start) {
    uint64_t run = 0;
    for (uint32_t g = 0; g < regions; ++g) {
      region_start[g] = run;
      for (uint32_t s = 0; s < world; ++s) run += reg[(size_t)s * regions + g];
    }
  }
  *total = sum;
  *n_segments = n;
  return DWJ_OK;
}

int dwj_xj_block_bytes(const dwj_engine *e, const dwj_xj_config *cfg, uint64_t *bytes) {
  if (!bytes) return xfail(DWJ_ERR_INVALID, "null argument");
  dwj_xj tmp;
  if (int rc = xj_layout(e, cfg, &tmp)) return rc;
  *bytes = tmp.ctrl_bytes + 2 * round_up(tmp.rows_total * tmp.W, 4096);
  return DWJ_OK;
}

int dwj_xj_destroy(dwj_xj *x) {
  if (!x) return DWJ_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(x->device);
  for (cudaStream_t s : {x->s_part, x->s_join, x->s_pull})
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (cudaEvent_t ev : {x->ev_in, x->ev_part_end, x->ev_join_end, x->ev_build_pulled, x->ev_pulled[0], x->ev_pulled[1], x->ev_consumed[0], x->ev_consumed[1]})
    if (ev) cudaEventDestroy(ev);
  for (auto ev : x->ev_t) if (ev) cudaEventDestroy(ev);
  for (auto ev : x->ev_region_off) if (ev) cudaEventDestroy(ev);
  cudaFree(x->d_counts);
  cudaFree(x->d_region_off);
  cudaFree(x->local_build);
  cudaFree(x->local_probe[0]);
  cudaFree(x->local_probe[1]);
  cudaFree(x->arena_a);
  cudaFree(x->arena_b);
  cudaFree(x->d_runs);
  if (x->h_runs) cudaFreeHost(x->h_runs);
  for (auto ev : x->ev_runs) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : {x->ev_scat_b, x->ev_scat[0], x->ev_scat[1]}) if (ev) cudaEventDestroy(ev);
  if (x->h_counts) cudaFreeHost(x->h_counts);
  if (x->h_gather) cudaFreeHost(x->h_gather);
  if (x->h_region_off) cudaFreeHost(x->h_region_off);
  cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  delete x;
  return DWJ_OK;
}

int dwj_xj_create(dwj_engine *e, const dwj_xj_config *cfg, void *const *blocks, dwj_xj **out) {
  if (!out || !blocks) return xfail(DWJ_ERR_INVALID, "null argument");
  *out = nullptr;
  dwj_xj *x = new (std::nothrow) dwj_xj();
  if (!x) return xfail(DWJ_ERR_OOM, "host allocation failed");
  if (int rc = xj_layout(e, cfg, x)) { delete x; return rc; }
  x->e = e;
  if (const char *v = getenv("DWJ_XJ_TIMEOUT_MS")) x->timeout_ns = std::max(1ull, std::strtoull(v, nullptr, 10)) * 1000000ull;
  x->unique = (x->info.flags & DWJ_FLAG_UNIQUE_BUILD_KEYS) != 0;
  // Direct pull needs the regions folded into the sender's pass and the segmented probe (unique build keys).
  x->direct = x->fold_regions == x->regions && x->unique && !cfg->force_scatter_pull;
  const uint64_t col_bytes = round_up(x->rows_total * x->W, 4096);
  for (uint32_t r = 0; r < x->world; ++r) {
    if (!blocks[r]) { delete x; return xfail(DWJ_ERR_INVALID, "null block pointer for rank %u", r); }
    x->peers.ctrl[r] = (unsigned long long *)blocks[r];
    x->keys_base[r] = (char *)blocks[r] + x->ctrl_bytes;
    x->vals_base[r] = x->keys_base[r] + col_bytes;
  }
  x->ctrl = x->peers.ctrl[x->me];
  x->device = x->info.device;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(x->device);
  auto bail = [&](int rc) { dwj_xj_destroy(x); if (prev >= 0) cudaSetDevice(prev); return rc; };
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);      // hi = numerically lowest = highest priority
  // Partition and transfer streams above the join stream: everything downstream, on every rank, waits for them.
  // A/B on 8 GPUs (profiles/r2_exchange.md): config 5 29.9 ms with these priorities, 31.6 ms with one priority for all;
  // config 2 per GPU 7.26 against 8.10 ms.  DWJ_XJ_PRIORITY=0 gives all three streams one priority.
  if (getenv("DWJ_XJ_PRIORITY") && !atoi(getenv("DWJ_XJ_PRIORITY"))) hi = lo;
  if (cudaStreamCreateWithPriority(&x->s_part, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&x->s_pull, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&x->s_join, cudaStreamNonBlocking, lo) != cudaSuccess)
    return bail(xfail(DWJ_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError())));
  for (cudaEvent_t *ev : {&x->ev_in, &x->ev_part_end, &x->ev_join_end, &x->ev_build_pulled, &x->ev_pulled[0], &x->ev_pulled[1], &x->ev_consumed[0], &x->ev_consumed[1]})
    if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return bail(xfail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
  for (auto &ev : x->ev_region_off)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(xfail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
  for (auto &ev : x->ev_t)
    if (cudaEventCreate(&ev) != cudaSuccess) return bail(xfail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
  const uint64_t count_words = (uint64_t)x->batches * x->world * x->regions;
  if (cudaMalloc((void **)&x->d_counts, (count_words + 16) * 8) != cudaSuccess ||
      cudaMalloc((void **)&x->d_region_off, 4 * (uint64_t)(x->regions + 1) * 8) != cudaSuccess ||
      cudaHostAlloc((void **)&x->h_counts, (count_words + 16) * 8, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void **)&x->h_gather, (x->world * x->src_stride_words + 8) * 8, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void **)&x->h_region_off, 4 * (uint64_t)(x->regions + 1) * 8, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
    return bail(xfail(DWJ_ERR_OOM, "scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  x->copy_pull = x->world > 1 && !(getenv("DWJ_XJ_FUSED_PULL") && atoi(getenv("DWJ_XJ_FUSED_PULL")));
  cudaDeviceGetAttribute(&x->sm_count, cudaDevAttrMultiProcessorCount, x->device);
  if (x->copy_pull) {
    const double slack = cfg->recv_slack > 0 ? cfg->recv_slack : 1.25;
    x->cap_build = std::max<uint64_t>(x->info.max_build_rows, 64);
    x->cap_recv_chunk = round_up((uint64_t)((double)x->slot_rows[1] * slack) + 4096, 64);
    x->arena_rows = x->direct ? x->cap_build + 2 * x->cap_recv_chunk : std::max(x->cap_build, 2 * x->cap_recv_chunk);
    constexpr uint32_t RUN_SLOTS = 32;
    if (cudaMalloc(&x->arena_a, 2 * x->arena_rows * x->W) != cudaSuccess || (!x->direct && cudaMalloc(&x->arena_b, 2 * x->arena_rows * x->W) != cudaSuccess) ||
        cudaMalloc((void **)&x->d_runs, RUN_SLOTS * 2 * MAX_WORLD * sizeof(CopyRun)) != cudaSuccess ||
        cudaHostAlloc((void **)&x->h_runs, RUN_SLOTS * 2 * MAX_WORLD * sizeof(CopyRun), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
      return bail(xfail(DWJ_ERR_OOM, "landing buffers (%llu rows x %d): %s", (unsigned long long)x->arena_rows, x->direct ? 1 : 2,
                        cudaGetErrorString(cudaGetLastError())));
    for (auto &ev : x->ev_runs)
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(xfail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
    for (cudaEvent_t *ev : {&x->ev_scat_b, &x->ev_scat[0], &x->ev_scat[1]})
      if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return bail(xfail(DWJ_ERR_CUDA, "cudaEventCreate failed"));
  } else if (!x->direct) {
    // Landing buffers of the pulling scatter: what this rank can receive (table capacity; a probe chunk from every
    // source with head-room -- the plan checks the real counts).
    const double slack = cfg->recv_slack > 0 ? cfg->recv_slack : 1.25;
    x->local_build_rows = std::max<uint64_t>(x->info.max_build_rows, 64);
    x->cap_recv_chunk = round_up((uint64_t)((double)x->slot_rows[1] * slack) + 4096, 64);
    if (cudaMalloc(&x->local_build, 2 * x->local_build_rows * x->W) != cudaSuccess ||
        cudaMalloc(&x->local_probe[0], 2 * x->cap_recv_chunk * x->W) != cudaSuccess ||
        cudaMalloc(&x->local_probe[1], 2 * x->cap_recv_chunk * x->W) != cudaSuccess)
      return bail(xfail(DWJ_ERR_OOM, "landing buffers (%llu + 2 x %llu rows): %s", (unsigned long long)x->local_build_rows,
                        (unsigned long long)x->cap_recv_chunk, cudaGetErrorString(cudaGetLastError())));
  }
  // Everything two ranks must agree on for equal keys to meet: table geometry, hash seed, key width,
  // batch structure.  Checked against every peer's header at each step.
  unsigned long long h = 1469598103934665603ull;
  for (unsigned long long v : {(unsigned long long)x->info.slots, (unsigned long long)x->info.hash_seed, (unsigned long long)x->W,
                               (unsigned long long)x->world, (unsigned long long)x->batches, (unsigned long long)x->chunk_rows,
                               (unsigned long long)x->passes, (unsigned long long)x->regions, (unsigned long long)(x->unique ? 1 : 0),
                               (unsigned long long)x->slot_rows[0], (unsigned long long)x->slot_rows[1]}) {
    h ^= v;
    h *= 1099511628211ull;
  }
  x->cfg_hash = h;
  // The control block starts at zero: all sequence numbers count from 1.  (Every rank clears its own block; the caller
  // synchronises the ranks between create and the first join -- dwj_mg does, torch's rendezvous does.)
  if (cudaMemset(x->ctrl, 0, x->ctrl_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
    return bail(xfail(DWJ_ERR_CUDA, "clearing the control block failed: %s", cudaGetErrorString(cudaGetLastError())));
  if (prev >= 0) cudaSetDevice(prev);
  *out = x;
  return DWJ_OK;
}

int dwj_xj_describe(const dwj_xj *x, dwj_xj_info *info) {
  if (!x || !info) return xfail(DWJ_ERR_INVALID, "null argument");
  info->regions = x->regions;
  info->fold_regions = x->fold_regions;
  info->chunks = x->chunks;
  info->ring = x->ring;
  info->passes = x->passes;
  info->direct_pull = x->direct ? 1u : 0u;
  info->chunk_rows = x->chunk_rows;
  info->block_bytes = x->ctrl_bytes + 2 * round_up(x->rows_total * x->W, 4096);
  info->landing_bytes = x->copy_pull ? 2 * x->arena_rows * x->W * (x->direct ? 1 : 2)
                                     : x->direct ? 0 : 2 * (x->local_build_rows + 2 * x->cap_recv_chunk) * x->W;
  info->copy_pull = x->copy_pull ? 1u : 0u;
  info->compact_passes = (x->world == 1 && x->passes > 1 && x->chunks == 1 && x->direct && x->slot_rows[0] == x->slot_rows[1] &&
                          !(getenv("DWJ_XJ_NO_LOCAL_PASS") && atoi(getenv("DWJ_XJ_NO_LOCAL_PASS")))) ? 1u : 0u;
  info->reserved = 0;
  return DWJ_OK;
}

int dwj_xj_timings(dwj_xj *x, dwj_xj_timing *t) {
  if (!x || !t) return xfail(DWJ_ERR_INVALID, "null argument");
  *t = x->last;
  return DWJ_OK;
}

// One pass over one key class on ONE GPU with nothing to exchange (world == 1, passes > 1, one probe chunk): the class's
// rows of a relation are first compacted out of the input (dwj_filter_rows: one streaming pass that also counts the kept
// rows per table region) into the slot the other relation is not using, then partitioned by table region from the
// compact copy (full tiles, no filter in the 512-way scatter) and built / probed through one segment per region.  Everything on one stream: on one
// GPU the step is work-bound, there is nothing to overlap with.  Measured on the 2^31 x 2^31 int64 join: the
// pass-filtered 512-way scatter ran at 1.8 TB/s of real DRAM traffic (half of every tile dropped) and was 47 % of the step.
static int xj_pass_local(dwj_xj *x, uint32_t pass, const void *bk, const void *bv, uint64_t n_build, const void *pk, const void *pv,
                         uint64_t n_probe, void *ok, void *ob, void *op, uint64_t capacity, uint64_t *d_count, bool timed) {
  dwj_engine *e = x->e;
  const uint32_t G = x->regions, W = x->W;
  cudaStream_t s = x->s_join;
  const uint64_t filter_on = (uint64_t)log2u(x->passes) << 8 | (uint64_t)pass << 16;
  char *slot_k[2] = {x->keys_base[0] + slot_base_row(x, 0) * W, x->keys_base[0] + slot_base_row(x, 1) * W};
  char *slot_v[2] = {x->vals_base[0] + slot_base_row(x, 0) * W, x->vals_base[0] + slot_base_row(x, 1) * W};
  unsigned long long *d_live = x->d_counts + (uint64_t)x->batches * G, *h_live = x->h_counts + (uint64_t)x->batches * G;
  std::vector<uint64_t> start(G), rows(G);
  std::vector<const void *> sk(G), sv(G);
  ++x->step;
  if (pass == 0) XCU(cudaMemsetAsync(d_count, 0, 8, s));
  for (int rel = 0; rel < 2; ++rel) {                       // 0: build relation -> slot 0 via slot 1; 1: probe relation -> slot 1 via slot 0
    const void *k = rel ? pk : bk, *v = rel ? pv : bv;
    const uint64_t n = rel ? n_probe : n_build;
    char *tk = slot_k[1 - rel], *tv = slot_v[1 - rel];
    XRC(dwj_set_option(e, DWJ_OPT_PASS_FILTER, filter_on));
    XRC(dwj_filter_rows(e, k, v, n, tk, tv, (uint64_t *)d_live, (uint64_t *)x->d_counts, s));      // + the kept rows per table region
    XRC(dwj_set_option(e, DWJ_OPT_PASS_FILTER, 0));
    XCU(cudaMemcpyAsync(h_live, d_live, 8, cudaMemcpyDeviceToHost, s));
    XCU(cudaMemcpyAsync(x->h_counts, x->d_counts, (uint64_t)G * 8, cudaMemcpyDeviceToHost, s));
    XCU(cudaStreamSynchronize(s));
    const uint64_t live = *h_live;
    if (live > x->slot_rows[rel])
      return xfail(DWJ_ERR_CAPACITY, "key class %u of the %s relation has %llu rows, its slot holds %llu", pass, rel ? "probe" : "build",
                   (unsigned long long)live, (unsigned long long)x->slot_rows[rel]);
    if (!rel && live > x->info.max_build_rows && (double)live > 0.9 * (double)x->info.slots)
      return xfail(DWJ_ERR_CAPACITY, "key class %u holds %llu build rows, the table was created for %llu", pass, (unsigned long long)live,
                   (unsigned long long)x->info.max_build_rows);
    if (timed && !rel) XCU(cudaEventRecord(x->ev_t[1], s));
    dwj_xj_plan_send(1, G, G, slot_base_row(x, rel), (const uint64_t *)x->h_counts, start.data());
    if (!rel) XRC(dwj_clear_table(e, s));
    XRC(dwj_xpart_scatter(e, tk, tv, live, 1, start.data(), x->keys_base[0], x->vals_base[0], s));
    for (uint32_t g = 0; g < G; ++g) {
      sk[g] = x->keys_base[0] + start[g] * W;
      sv[g] = x->vals_base[0] + start[g] * W;
      rows[g] = x->h_counts[g];
    }
    if (!rel) {
      XRC(dwj_build_segments(e, G, sk.data(), sv.data(), rows.data(), 1, s));
      if (timed) XCU(cudaEventRecord(x->ev_t[3], s));
    } else {
      if (timed) XCU(cudaEventRecord(x->ev_t[2], s));
      XRC(dwj_set_option(e, DWJ_OPT_APPEND_OUTPUT, 1));
      const int rc = dwj_probe_pairs_segments(e, G, sk.data(), sv.data(), rows.data(), ok, ob, op, capacity, d_count, nullptr, s);
      dwj_set_option(e, DWJ_OPT_APPEND_OUTPUT, 0);
      if (rc) return rc;
    }
  }
  XCU(cudaEventRecord(x->ev_join_end, s));
  return DWJ_OK;
}

// One pass of the join over one key class (all keys when passes == 1).
static int xj_pass(dwj_xj *x, uint32_t pass, const void *bk, const void *bv, uint64_t n_build, const void *pk, const void *pv,
                   uint64_t n_probe, void *ok, void *ob, void *op, uint64_t capacity, uint64_t *d_count, bool timed) {
  dwj_engine *e = x->e;
  const uint32_t w = x->world, G = x->regions, B = x->batches, W = x->W;
  const uint32_t rank_bits = log2u(w);
  if (x->passes > 1) XRC(dwj_set_option(e, DWJ_OPT_PASS_FILTER, (uint64_t)rank_bits | (uint64_t)log2u(x->passes) << 8 | (uint64_t)pass << 16));
  const unsigned long long step = ++x->step;
  const uint32_t par = (uint32_t)(step & 1);
  const uint64_t gather0 = x->gather_word0 + (uint64_t)par * w * x->src_stride_words;
  struct Batch { const char *k, *v; uint64_t n; };
  std::vector<Batch> batch(B);
  batch[0] = {(const char *)bk, (const char *)bv, n_build};
  for (uint32_t c = 0; c < x->chunks; ++c) {
    const uint64_t r0 = std::min<uint64_t>((uint64_t)c * x->chunk_rows, n_probe), r1 = std::min<uint64_t>(r0 + x->chunk_rows, n_probe);
    batch[1 + c] = {(const char *)pk + r0 * W, (const char *)pv + r0 * W, r1 - r0};
  }

  // ---- 1. counts: every batch, per (destination rank, table region); published to the peers; one host sync ------------
  for (uint32_t b = 0; b < B; ++b) XRC(dwj_xpart_hist2(e, batch[b].k, batch[b].n, w, (uint64_t *)(x->d_counts + (uint64_t)b * w * G), x->s_part));
  xj_publish_kernel<<<w, 256, 0, x->s_part>>>(x->peers, w, x->me, x->d_counts, B, G, gather0, x->src_stride_words, x->cfg_hash, step);
  XCU(cudaGetLastError());
  XRC(wait_all(x, FLAG_CNT, step, x->s_part));
  XCU(cudaMemcpyAsync(x->h_gather, x->ctrl + gather0, w * x->src_stride_words * 8, cudaMemcpyDeviceToHost, x->s_part));
  XCU(cudaMemcpyAsync(x->h_counts, x->d_counts, (uint64_t)B * w * G * 8, cudaMemcpyDeviceToHost, x->s_part));
  unsigned long long *h_err = x->h_gather + (uint64_t)w * x->src_stride_words;     // the error word rides along
  XCU(cudaMemcpyAsync(h_err, x->ctrl + FLAG_ERR, 8, cudaMemcpyDeviceToHost, x->s_part));
  if (timed) XCU(cudaEventRecord(x->ev_t[1], x->s_part));
  XCU(cudaStreamSynchronize(x->s_part));
  const unsigned long long err_word = *h_err;
  if (err_word) return xfail(DWJ_ERR_STATE, "rank %u: timed out waiting for flag %llu of a peer (step %llu)", x->me, err_word - 1, step);
  for (uint32_t s = 0; s < w; ++s) {
    const unsigned long long *hd = x->h_gather + (uint64_t)s * x->src_stride_words;
    if (hd[0] != x->cfg_hash || hd[1] != step)
      return xfail(DWJ_ERR_STATE, "rank %u and rank %u disagree on the join configuration (table size, hash seed, key width, chunking) "
                                  "or are out of step (%llu vs %llu): equal keys would not meet", x->me, s, hd[1], step);
  }

  // ---- 2. plan (host) ---------------------------------------------------------------------------------------------------
  auto tot = [&](uint32_t s, uint32_t b, uint32_t d) { return x->h_gather[(uint64_t)s * x->src_stride_words + HDR_WORDS + (uint64_t)b * (w + G) + d]; };
  auto reg = [&](uint32_t s, uint32_t b, uint32_t g) { return x->h_gather[(uint64_t)s * x->src_stride_words + HDR_WORDS + (uint64_t)b * (w + G) + w + g]; };
  // capacity checks every rank can make for every rank (all totals are known everywhere): all ranks fail together
  for (uint32_t d = 0; d < w; ++d) {
    uint64_t nb = 0;
    for (uint32_t s = 0; s < w; ++s) nb += tot(s, 0, d);
    if (nb > x->info.max_build_rows && (double)nb > 0.9 * (double)x->info.slots)
      return xfail(DWJ_ERR_CAPACITY, "rank %u would receive %llu build rows, its table was created for %llu", d, (unsigned long long)nb,
                   (unsigned long long)x->info.max_build_rows);
    if (!x->direct || x->copy_pull)
      for (uint32_t b = 1; b < B; ++b) {
        uint64_t np_ = 0;
        for (uint32_t s = 0; s < w; ++s) np_ += tot(s, b, d);
        if (np_ > x->cap_recv_chunk)
          return xfail(DWJ_ERR_CAPACITY, "rank %u would receive %llu rows of probe chunk %u, its landing buffer holds %llu (raise recv_slack)",
                       d, (unsigned long long)np_, b - 1, (unsigned long long)x->cap_recv_chunk);
      }
  }
  for (uint32_t s = 0; s < w; ++s)
    for (uint32_t b = 0; b < B; ++b) {
      uint64_t rows = 0;
      for (uint32_t d = 0; d < w; ++d) rows += tot(s, b, d);
      if (rows > x->slot_rows[b ? 1 : 0])
        return xfail(DWJ_ERR_CAPACITY, "batch %u of rank %u has %llu rows after the pass filter, its send slot holds %llu", b, s,
                     (unsigned long long)rows, (unsigned long long)x->slot_rows[b ? 1 : 0]);
    }

  // ---- 3. sender: one scatter per batch into its slot, then the ready flag ------------------------------------------------
  const uint32_t fold = x->fold_regions, parts = w * fold;
  std::vector<uint64_t> start(parts);
  std::vector<unsigned long long> use(B);
  for (uint32_t b = 0; b < B; ++b) use[b] = ++x->slot_use[b == 0 ? 0 : 1 + (b - 1) % x->ring];
  auto send_batch = [&](uint32_t b) -> int {
    const uint32_t slot = b == 0 ? 0 : 1 + (b - 1) % x->ring;
    dwj_xj_plan_send(w, G, fold, slot_base_row(x, slot), (const uint64_t *)(x->h_counts + (uint64_t)b * w * G), start.data());
    if (use[b] > 1) XRC(wait_all(x, FLAG_DONE + slot * MAX_WORLD, use[b] - 1, x->s_part));    // every reader of the previous filling is done
    XRC(dwj_xpart_scatter(e, batch[b].k, batch[b].v, batch[b].n, w, start.data(), x->keys_base[x->me], x->vals_base[x->me], x->s_part));
    XRC(signal_all(x, FLAG_READY + slot * MAX_WORLD, use[b], x->s_part));
    if (b + 1 == B) {
      if (timed) XCU(cudaEventRecord(x->ev_t[2], x->s_part));
      XCU(cudaEventRecord(x->ev_part_end, x->s_part));
    }
    return DWJ_OK;
  };

  // ---- 4. receiver: pull --------------------------------------------------------------------------------------------------
  // where rank s keeps its rows for me (and, inside that block, for my region g) in the slot of batch b: dwj_xj_plan_recv
  std::vector<const void *> sk, sv;
  std::vector<uint64_t> sr, sfirst, rstart(G), tot_b((size_t)w * w), reg_b((size_t)w * G);
  uint64_t remote_bytes = 0;
  const uint32_t pieces = std::max<uint32_t>(1, 256 / w);        // scatter pull: pieces per source, dealt round-robin
  std::vector<uint32_t> ssrc;
  uint32_t nseg = 0;
  auto plan_recv = [&](uint32_t b, bool direct) {
    for (uint32_t s = 0; s < w; ++s) {
      for (uint32_t d = 0; d < w; ++d) tot_b[(size_t)s * w + d] = tot(s, b, d);
      for (uint32_t g = 0; g < G; ++g) reg_b[(size_t)s * G + g] = reg(s, b, g);
    }
    const size_t cap = (size_t)std::max(G, pieces) * w;
    sfirst.assign(cap, 0); sr.assign(cap, 0); ssrc.assign(cap, 0); sk.assign(cap, nullptr); sv.assign(cap, nullptr);
    uint64_t total = 0;
    dwj_xj_plan_recv(w, x->me, G, slot_base_row(x, b == 0 ? 0 : 1 + (b - 1) % x->ring), tot_b.data(), reg_b.data(), direct ? 1 : 0, pieces,
                     sfirst.data(), sr.data(), ssrc.data(), rstart.data(), &total, &nseg);
    for (uint32_t i = 0; i < nseg; ++i) {
      const uint32_t src = ssrc[i];
      sk[i] = x->keys_base[src] + sfirst[i] * W;
      sv[i] = x->vals_base[src] + sfirst[i] * W;
      if (src != x->me) remote_bytes += 2ull * sr[i] * W;
    }
    return total;
  };
  // ---- copy pull: rows of batch b, from every source's slot into a local landing area, source-major --------------------
  std::vector<uint64_t> lrow(w + 1);
  auto block_row = [&](uint32_t s, uint32_t b) {
    uint64_t r = slot_base_row(x, b == 0 ? 0 : 1 + (b - 1) % x->ring);
    for (uint32_t d = 0; d < x->me; ++d) r += tot(s, b, d);
    return r;
  };
  auto copy_batch = [&](uint32_t b, char *lk, char *lv, cudaStream_t st) -> int {
    constexpr uint32_t RUN_SLOTS = 32, PER = 2 * MAX_WORLD;
    const uint32_t slot = x->run_calls++ % RUN_SLOTS;
    if (x->run_calls > RUN_SLOTS) XCU(cudaEventSynchronize(x->ev_runs[slot]));
    CopyRun *h = x->h_runs + (size_t)slot * PER, *d = x->d_runs + (size_t)slot * PER;
    lrow[0] = 0;
    for (uint32_t s = 0; s < w; ++s) lrow[s + 1] = lrow[s] + tot(s, b, x->me);
    uint32_t n = 0, max_blocks = 0;
    for (uint32_t i = 0; i < w; ++i) {                        // rotated: at any time every source is read by one reader
      const uint32_t src = (x->me + i) % w;
      const uint64_t bytes = tot(src, b, x->me) * W, from = block_row(src, b) * W, to = lrow[src] * W;
      if (!bytes) continue;
      for (int col = 0; col < 2; ++col) {
        CopyRun r{};
        r.src = (col ? x->vals_base[src] : x->keys_base[src]) + from;
        r.dst = (col ? lv : lk) + to;
        r.bytes = bytes;
        r.head = (unsigned int)std::min<uint64_t>(bytes, (16 - ((uintptr_t)r.src & 15)) & 15);
        r.blocks = (unsigned int)((bytes - r.head + COPY_BLOCK - 1) / COPY_BLOCK);
        if (!r.blocks) r.blocks = 1;                          // the head alone
        max_blocks = std::max(max_blocks, r.blocks);
        h[n++] = r;
      }
      if (src != x->me) remote_bytes += 2 * bytes;
    }
    if (n) {
      static_assert(sizeof(CopyRun) % 8 == 0, "CopyRun is staged word by word");
      xj_stage_kernel<<<1, 256, 0, st>>>((unsigned long long *)d, (const unsigned long long *)h, n * (uint32_t)(sizeof(CopyRun) / 8));
      const unsigned long long total = (unsigned long long)max_blocks * n;
      // two CTAs per SM keep 64 KB per SM in flight -- 9 MB on the chip against the ~3 MB that 770 GB/s x 3.5 us needs --
      // and leave the SMs' thread slots to the kernels that run beside the transfer
      static const int per_sm = getenv("DWJ_XJ_COPY_CTAS") ? std::max(1, atoi(getenv("DWJ_XJ_COPY_CTAS"))) : 2;
      xj_copy_runs_kernel<<<(unsigned)std::min<unsigned long long>(total, (unsigned long long)x->sm_count * per_sm), COPY_THREADS, 0, st>>>(d, n, max_blocks);
      XCU(cudaGetLastError());
    }
    XCU(cudaEventRecord(x->ev_runs[slot], st));
    return DWJ_OK;
  };
  // segments of a landed batch for the direct path: (region, source), all local now
  auto landed_segments = [&](uint32_t b, char *lk, char *lv) {
    sk.assign((size_t)G * w, nullptr); sv.assign((size_t)G * w, nullptr); sr.assign((size_t)G * w, 0);
    for (uint32_t s = 0; s < w; ++s) {
      uint64_t r = lrow[s];
      for (uint32_t g = 0; g < G; ++g) {
        sk[(size_t)g * w + s] = lk + r * W;
        sv[(size_t)g * w + s] = lv + r * W;
        sr[(size_t)g * w + s] = reg(s, b, g);
        r += reg(s, b, g);
      }
    }
    nseg = G * w;
  };
  auto region_starts = [&](uint32_t b) {
    uint64_t run = 0;
    for (uint32_t g = 0; g < G; ++g) {
      rstart[g] = run;
      for (uint32_t s = 0; s < w; ++s) run += reg(s, b, g);
    }
    return run;
  };
  char *const A_k = (char *)x->arena_a, *const A_v = A_k + x->arena_rows * W;
  char *const B_k = (char *)x->arena_b, *const B_v = x->arena_b ? B_k + x->arena_rows * W : nullptr;
  auto copy_recv_build = [&]() -> int {
    XRC(dwj_clear_table(e, x->s_join));
    XRC(wait_all(x, FLAG_READY, use[0], x->s_pull));
    XCU(cudaStreamWaitEvent(x->s_pull, x->ev_join_end, 0));          // the previous pass / step has left the landing areas
    XRC(copy_batch(0, A_k, A_v, x->s_pull));
    XRC(signal_all(x, FLAG_DONE, use[0], x->s_pull));
    XCU(cudaEventRecord(x->ev_build_pulled, x->s_pull));
    if (timed) XCU(cudaEventRecord(x->ev_t[5], x->s_pull));
    XCU(cudaStreamWaitEvent(x->s_join, x->ev_build_pulled, 0));
    if (x->direct) {
      landed_segments(0, A_k, A_v);
      XRC(dwj_build_segments(e, nseg, sk.data(), sv.data(), sr.data(), w, x->s_join));
    } else {
      const uint64_t nb = region_starts(0);
      const void *k1[1] = {A_k}, *v1[1] = {A_v};
      const uint64_t r1[1] = {nb};
      XRC(dwj_region_scatter_segments(e, 1, k1, v1, r1, rstart.data(), B_k, B_v, x->s_join));
      XCU(cudaEventRecord(x->ev_scat_b, x->s_join));                 // arena A may take the first probe chunks
      const uint32_t ro = x->region_off_calls++ % 4;
      if (x->region_off_calls > 4) XCU(cudaEventSynchronize(x->ev_region_off[ro]));
      unsigned long long *h_ro = x->h_region_off + (uint64_t)ro * (G + 1), *d_ro = x->d_region_off + (uint64_t)ro * (G + 1);
      for (uint32_t g = 0; g < G; ++g) h_ro[g] = rstart[g];
      h_ro[G] = nb;
      xj_stage_kernel<<<1, 256, 0, x->s_join>>>(d_ro, h_ro, G + 1);
      XCU(cudaGetLastError());
      XRC(dwj_build_grouped(e, B_k, B_v, nb, G > 1 ? (const uint64_t *)d_ro : nullptr, x->s_join));
      XCU(cudaEventRecord(x->ev_region_off[ro], x->s_join));
    }
    if (timed) XCU(cudaEventRecord(x->ev_t[3], x->s_join));
    return DWJ_OK;
  };
  auto copy_recv_chunk = [&](uint32_t c) -> int {
    const uint32_t b = 1 + c, slot = 1 + c % x->ring, h = c & 1;
    // landing of this chunk: direct -- behind the build landing; scatter -- one half of arena A (raw), grouped into
    // the same half of arena B
    const uint64_t row0 = x->direct ? x->cap_build + (uint64_t)h * x->cap_recv_chunk : (uint64_t)h * (x->arena_rows / 2);
    char *lk = A_k + row0 * W, *lv = A_v + row0 * W;
    XRC(wait_all(x, FLAG_READY + slot * MAX_WORLD, use[b], x->s_pull));
    if (x->direct) {
      if (c >= 2) XCU(cudaStreamWaitEvent(x->s_pull, x->ev_consumed[h], 0));      // the probe that last read this landing
    } else {
      if (c < 2) XCU(cudaStreamWaitEvent(x->s_pull, x->ev_scat_b, 0));            // arena A held the raw build rows
      else XCU(cudaStreamWaitEvent(x->s_pull, x->ev_scat[h], 0));                 // the scatter that last read this half
    }
    XRC(copy_batch(b, lk, lv, x->s_pull));
    XRC(signal_all(x, FLAG_DONE + slot * MAX_WORLD, use[b], x->s_pull));
    XCU(cudaEventRecord(x->ev_pulled[h], x->s_pull));
    if (timed && c + 1 == x->chunks) XCU(cudaEventRecord(x->ev_t[6], x->s_pull));
    XCU(cudaStreamWaitEvent(x->s_join, x->ev_pulled[h], 0));
    if (x->direct) {
      landed_segments(b, lk, lv);
      XRC(dwj_probe_pairs_segments(e, nseg, sk.data(), sv.data(), sr.data(), ok, ob, op, capacity, d_count, nullptr, x->s_join));
      XCU(cudaEventRecord(x->ev_consumed[h], x->s_join));
    } else {
      const uint64_t np_ = region_starts(b);
      char *gk = B_k + row0 * W, *gv = B_v + row0 * W;                // B is free: the build and the probe of chunk c-2 precede on this stream
      const void *k1[1] = {lk}, *v1[1] = {lv};
      const uint64_t r1[1] = {np_};
      XRC(dwj_region_scatter_segments(e, 1, k1, v1, r1, rstart.data(), gk, gv, x->s_join));
      XCU(cudaEventRecord(x->ev_scat[h], x->s_join));
      XRC(dwj_probe_pairs_grouped(e, gk, gv, np_, ok, ob, op, capacity, d_count, nullptr, x->s_join));
    }
    return DWJ_OK;
  };

  auto recv_build = [&]() -> int {
    if (x->copy_pull) return copy_recv_build();
    // The clear of the table needs nothing from the peers: it runs (on the join stream, behind the previous probes)
    // while the senders still partition and the first rows are pulled.
    XRC(dwj_clear_table(e, x->s_join));
    if (x->direct) {
      plan_recv(0, true);
      XRC(wait_all(x, FLAG_READY, use[0], x->s_join));
      XRC(dwj_build_segments(e, nseg, sk.data(), sv.data(), sr.data(), w, x->s_join));
      XRC(signal_all(x, FLAG_DONE, use[0], x->s_join));
    } else {
      const uint64_t nb = plan_recv(0, false);
      XRC(wait_all(x, FLAG_READY, use[0], x->s_pull));
      XCU(cudaStreamWaitEvent(x->s_pull, x->ev_join_end, 0));        // the previous step's build may still read the landing buffer
      char *lk = (char *)x->local_build, *lv = lk + x->local_build_rows * W;
      XRC(dwj_region_scatter_segments(e, nseg, sk.data(), sv.data(), sr.data(), rstart.data(), lk, lv, x->s_pull));
      XRC(signal_all(x, FLAG_DONE, use[0], x->s_pull));
      XCU(cudaEventRecord(x->ev_build_pulled, x->s_pull));
      // region offsets of the landing buffer for the build's look-ahead (ring of 4 pinned / device arrays)
      const uint32_t ro = x->region_off_calls++ % 4;
      if (x->region_off_calls > 4) XCU(cudaEventSynchronize(x->ev_region_off[ro]));
      unsigned long long *h_ro = x->h_region_off + (uint64_t)ro * (G + 1), *d_ro = x->d_region_off + (uint64_t)ro * (G + 1);
      for (uint32_t g = 0; g < G; ++g) h_ro[g] = rstart[g];
      h_ro[G] = nb;
      xj_stage_kernel<<<1, 256, 0, x->s_join>>>(d_ro, h_ro, G + 1);      // a kernel, not a copy-engine job: see stage_words (dwj_api.cu)
      XCU(cudaGetLastError());
      XCU(cudaStreamWaitEvent(x->s_join, x->ev_build_pulled, 0));
      XRC(dwj_build_grouped(e, lk, lv, nb, G > 1 ? (const uint64_t *)d_ro : nullptr, x->s_join));
      XCU(cudaEventRecord(x->ev_region_off[ro], x->s_join));
    }
    if (timed) XCU(cudaEventRecord(x->ev_t[3], x->s_join));
    return DWJ_OK;
  };
  auto recv_chunk = [&](uint32_t c) -> int {
    if (x->copy_pull) return copy_recv_chunk(c);
    const uint32_t b = 1 + c, slot = 1 + c % x->ring;
    if (x->direct) {
      plan_recv(b, true);
      XRC(wait_all(x, FLAG_READY + slot * MAX_WORLD, use[b], x->s_join));
      XRC(dwj_probe_pairs_segments(e, nseg, sk.data(), sv.data(), sr.data(), ok, ob, op, capacity, d_count, nullptr, x->s_join));
      XRC(signal_all(x, FLAG_DONE + slot * MAX_WORLD, use[b], x->s_join));
    } else {
      const uint64_t np_ = plan_recv(b, false);
      const uint32_t lb = c & 1;
      char *lk = (char *)x->local_probe[lb], *lv = lk + x->cap_recv_chunk * W;
      XRC(wait_all(x, FLAG_READY + slot * MAX_WORLD, use[b], x->s_pull));
      if (c >= 2 || step > 1) XCU(cudaStreamWaitEvent(x->s_pull, x->ev_consumed[lb], 0));   // the probe that last read this landing buffer
      XRC(dwj_region_scatter_segments(e, nseg, sk.data(), sv.data(), sr.data(), rstart.data(), lk, lv, x->s_pull));
      XRC(signal_all(x, FLAG_DONE + slot * MAX_WORLD, use[b], x->s_pull));
      XCU(cudaEventRecord(x->ev_pulled[lb], x->s_pull));
      XCU(cudaStreamWaitEvent(x->s_join, x->ev_pulled[lb], 0));
      XRC(dwj_probe_pairs_grouped(e, lk, lv, np_, ok, ob, op, capacity, d_count, nullptr, x->s_join));
      XCU(cudaEventRecord(x->ev_consumed[lb], x->s_join));
    }
    return DWJ_OK;
  };

  // ---- 5. enqueue, in dependency order -------------------------------------------------------------------------------------
  // A batch that reuses a slot waits (inside a kernel, on the partition stream) until every reader of the slot's previous
  // filling is done.  This rank's own reader is therefore enqueued BEFORE that wait: hardware work queues are shared
  // between streams, and work submitted behind a polling kernel may not start until the kernel ends -- measured: with all
  // sender batches enqueued first, the pulls they waited for sat behind the waits until those timed out.
  if (pass == 0) XCU(cudaMemsetAsync(d_count, 0, 8, x->s_join));
  XRC(dwj_set_option(e, DWJ_OPT_APPEND_OUTPUT, 1));
  // Ranks that share one process (dwj_mg) may even share one GPU (tests): there the same rule must hold ACROSS ranks --
  // whatever a polling kernel waits for has been submitted before it -- so the rank threads meet at a host barrier
  // between every send and the receives that poll for it.  Ranks in separate processes own their GPU's queues.
  auto hb = [&]() { if (x->host_barrier) x->host_barrier(x->host_barrier_ctx); };
  XRC(send_batch(0));
  for (uint32_t b = 1; b <= std::min(x->ring, x->chunks); ++b) XRC(send_batch(b));
  hb();
  XRC(recv_build());
  for (uint32_t c = 0; c < x->chunks; ++c) {
    XRC(recv_chunk(c));
    if (c + x->ring + 1 < B) {
      hb();                                                          // every reader of the slot is enqueued ...
      XRC(send_batch(c + x->ring + 1));                              // ... before the batch that refills it
      hb();
    }
  }
  XRC(dwj_set_option(e, DWJ_OPT_APPEND_OUTPUT, 0));
  XCU(cudaEventRecord(x->ev_join_end, x->s_join));
  x->last_remote_bytes += remote_bytes;
  return DWJ_OK;
}

int dwj_xj_join(dwj_xj *x, const void *d_build_keys, const void *d_build_vals, uint64_t n_build, const void *d_probe_keys,
                const void *d_probe_vals, uint64_t n_probe, void *d_out_key, void *d_out_build_val, void *d_out_probe_val,
                uint64_t capacity, uint64_t *d_n_matches, void *stream) {
  if (!x) return xfail(DWJ_ERR_INVALID, "null exchange join");
  if (!d_n_matches) return xfail(DWJ_ERR_INVALID, "d_n_matches (device counter) is required");
  if ((n_build && (!d_build_keys || !d_build_vals)) || (n_probe && (!d_probe_keys || !d_probe_vals))) return xfail(DWJ_ERR_INVALID, "null input column");
  if (capacity && (!d_out_build_val || !d_out_probe_val)) return xfail(DWJ_ERR_INVALID, "null output column");
  // the send slots were sized from these; a larger input would run one slot into the next
  if (n_build > x->cfg.max_build_rows || n_probe > x->cfg.max_probe_rows)
    return xfail(DWJ_ERR_CAPACITY, "%llu build / %llu probe rows exceed the %llu / %llu this exchange join was created for",
                 (unsigned long long)n_build, (unsigned long long)n_probe, (unsigned long long)x->cfg.max_build_rows,
                 (unsigned long long)x->cfg.max_probe_rows);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(x->device);
  cudaStream_t cs = (cudaStream_t)stream;
  int rc = DWJ_OK;
  auto body = [&]() -> int {
    XCU(cudaEventRecord(x->ev_in, cs));
    for (cudaStream_t s : {x->s_part, x->s_join, x->s_pull}) XCU(cudaStreamWaitEvent(s, x->ev_in, 0));
    XCU(cudaEventRecord(x->ev_t[0], x->s_part));
    x->last_remote_bytes = 0;
    for (uint32_t pass = 0; pass < x->passes; ++pass) {
      // one GPU, nothing to exchange, the slots can serve as each other's scratch: compact-then-partition
      const bool local = x->world == 1 && x->passes > 1 && x->chunks == 1 && x->direct && x->slot_rows[0] == x->slot_rows[1] &&
                         !(getenv("DWJ_XJ_NO_LOCAL_PASS") && atoi(getenv("DWJ_XJ_NO_LOCAL_PASS")));
      if (int r = (local ? xj_pass_local : xj_pass)(x, pass, d_build_keys, d_build_vals, n_build, d_probe_keys, d_probe_vals, n_probe, d_out_key,
                                                    d_out_build_val, d_out_probe_val, capacity, d_n_matches, pass == 0))
        return r;
    }
    XCU(cudaEventRecord(x->ev_t[4], x->s_join));
    XCU(cudaStreamWaitEvent(cs, x->ev_join_end, 0));
    XCU(cudaStreamWaitEvent(cs, x->ev_part_end, 0));
    return DWJ_OK;
  };
  rc = body();
  if (x->passes > 1) dwj_set_option(x->e, DWJ_OPT_PASS_FILTER, 0);
  dwj_set_option(x->e, DWJ_OPT_APPEND_OUTPUT, 0);
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

// Timeline of the last dwj_xj_join on this rank (synchronises on its end).  Pass 0 only for the inner marks.
int dwj_xj_sync_timings(dwj_xj *x, dwj_xj_timing *t) {
  if (!x || !t) return xfail(DWJ_ERR_INVALID, "null argument");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(x->device);
  auto body = [&]() -> int {
    XCU(cudaEventSynchronize(x->ev_t[4]));
    dwj_xj_timing r{};
    XCU(cudaEventElapsedTime(&r.counts_ms, x->ev_t[0], x->ev_t[1]));
    XCU(cudaEventElapsedTime(&r.scattered_ms, x->ev_t[0], x->ev_t[2]));
    XCU(cudaEventElapsedTime(&r.built_ms, x->ev_t[0], x->ev_t[3]));
    XCU(cudaEventElapsedTime(&r.total_ms, x->ev_t[0], x->ev_t[4]));
    if (x->copy_pull) {
      XCU(cudaEventElapsedTime(&r.build_pulled_ms, x->ev_t[0], x->ev_t[5]));
      XCU(cudaEventElapsedTime(&r.last_pulled_ms, x->ev_t[0], x->ev_t[6]));
    }
    r.remote_bytes = x->last_remote_bytes;
    unsigned long long err_word = 0;
    XCU(cudaMemcpy(&err_word, x->ctrl + FLAG_ERR, 8, cudaMemcpyDeviceToHost));
    if (err_word) {
      unsigned long long fl[FLAG_DONE + MAX_SLOTS * MAX_WORLD];
      XCU(cudaMemcpy(fl, x->ctrl, sizeof(fl), cudaMemcpyDeviceToHost));
      char buf[400] = "";
      for (uint32_t sl = 0; sl <= x->ring; ++sl) {
        const size_t at = strlen(buf);
        snprintf(buf + at, sizeof(buf) - at, " slot%u use %llu ready[", sl, x->slot_use[sl]);
        for (uint32_t q = 0; q < x->world; ++q) { const size_t a2 = strlen(buf); snprintf(buf + a2, sizeof(buf) - a2, "%llu ", fl[FLAG_READY + sl * MAX_WORLD + q]); }
        { const size_t a2 = strlen(buf); snprintf(buf + a2, sizeof(buf) - a2, "] done["); }
        for (uint32_t q = 0; q < x->world; ++q) { const size_t a2 = strlen(buf); snprintf(buf + a2, sizeof(buf) - a2, "%llu ", fl[FLAG_DONE + sl * MAX_WORLD + q]); }
        { const size_t a2 = strlen(buf); snprintf(buf + a2, sizeof(buf) - a2, "]"); }
      }
      return xfail(DWJ_ERR_STATE, "rank %u: timed out waiting for flag %llu of a peer;%s", x->me, err_word - 1, buf);
    }
    x->last = r;
    *t = r;
    return DWJ_OK;
  };
  const int rc = body();
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------
// dwj_mg_*: all GPUs of one box from ONE process -- what the C++ host framework (`dwarf_bench Join --gpus N`) calls.
// One engine + one dwj_xj per GPU, peer access enabled both ways, blocks from cudaMalloc; dwj_mg_join runs the ranks on
// one host thread each (every rank blocks once per step on the count exchange, which needs all of them enqueued).
// ---------------------------------------------------------------------------------------------------------------------------
namespace {
// Reusable barrier of the rank threads of dwj_mg_join; abort() releases everybody for good (a rank failed).
struct HostBarrier {
  std::mutex mu;
  std::condition_variable cv;
  uint32_t n = 1, waiting = 0;
  uint64_t generation = 0;
  bool aborted = false;
  void arrive() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted || n <= 1) return;
    const uint64_t gen = generation;
    if (++waiting == n) {
      waiting = 0;
      ++generation;
      cv.notify_all();
      return;
    }
    cv.wait(lk, [&] { return generation != gen || aborted; });
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
  void reset(uint32_t ranks) {
    std::lock_guard<std::mutex> lk(mu);
    n = ranks; waiting = 0; aborted = false;
  }
};
void host_barrier_arrive(void *ctx) { static_cast<HostBarrier *>(ctx)->arrive(); }
}  // namespace

struct dwj_mg {
  HostBarrier barrier;
  dwj_mg_config cfg{};
  uint32_t n = 0;
  dwj_engine *eng[MAX_WORLD]{};
  dwj_xj *xj[MAX_WORLD]{};
  void *block[MAX_WORLD]{};
  cudaStream_t stream[MAX_WORLD]{};
  unsigned long long *d_count[MAX_WORLD]{};
  // dwj_mg_join_host staging
  void *stage[MAX_WORLD]{};
  uint64_t stage_rows_b = 0, stage_rows_p = 0, stage_rows_o = 0;
};

extern "C" {

int dwj_mg_destroy(dwj_mg *m) {
  if (!m) return DWJ_OK;
  for (uint32_t r = 0; r < m->n; ++r) {
    cudaSetDevice(m->cfg.devices[r]);
    cudaDeviceSynchronize();
  }
  for (uint32_t r = 0; r < m->n; ++r) {
    cudaSetDevice(m->cfg.devices[r]);
    dwj_xj_destroy(m->xj[r]);
    dwj_destroy(m->eng[r]);
    cudaFree(m->block[r]);
    cudaFree(m->d_count[r]);
    cudaFree(m->stage[r]);
    if (m->stream[r]) cudaStreamDestroy(m->stream[r]);
  }
  cudaGetLastError();
  delete m;
  return DWJ_OK;
}

int dwj_mg_create(const dwj_mg_config *cfg, dwj_mg **out) {
  if (!cfg || !out) return xfail(DWJ_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_gpus < 1 || cfg->n_gpus > (int)MAX_WORLD || (cfg->n_gpus & (cfg->n_gpus - 1)))
    return xfail(DWJ_ERR_INVALID, "n_gpus must be 1, 2, 4 or 8, got %d", cfg->n_gpus);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return xfail(DWJ_ERR_CUDA, "no CUDA device available; this engine has no CPU fallback");
  for (int r = 0; r < cfg->n_gpus; ++r)
    if (cfg->devices[r] < 0 || cfg->devices[r] >= ndev) return xfail(DWJ_ERR_INVALID, "device %d (rank %d) out of range: %d visible", cfg->devices[r], r, ndev);
  dwj_mg *m = new (std::nothrow) dwj_mg();
  if (!m) return xfail(DWJ_ERR_OOM, "host allocation failed");
  m->cfg = *cfg;
  m->n = (uint32_t)cfg->n_gpus;
  auto bail = [&](int rc) { dwj_mg_destroy(m); return rc; };
  // peer access, both ways, between distinct devices (a repeated ordinal = several ranks on one GPU, for tests)
  for (uint32_t a = 0; a < m->n; ++a)
    for (uint32_t b = 0; b < m->n; ++b) {
      const int da = cfg->devices[a], db = cfg->devices[b];
      if (da == db) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, da, db);
      if (!can) return bail(xfail(DWJ_ERR_CUDA, "GPU %d cannot map GPU %d's memory (no NVLink / PCIe P2P)", da, db));
      cudaSetDevice(da);
      const cudaError_t pe = cudaDeviceEnablePeerAccess(db, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return bail(xfail(DWJ_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", da, db, cudaGetErrorString(pe)));
      cudaGetLastError();
    }
  // what a rank receives: an even share of the global build relation plus head-room
  const double slack = cfg->recv_slack > 0 ? cfg->recv_slack : 1.25;
  const uint32_t passes = cfg->passes ? cfg->passes : 1;
  const uint64_t eng_rows = (uint64_t)((double)cfg->max_build_rows_per_gpu / passes * slack) + 1024;
  dwj_xj_config xc{};
  for (uint32_t r = 0; r < m->n; ++r) {
    dwj_config ec{};
    ec.device = cfg->devices[r];
    ec.key_bytes = ec.payload_bytes = cfg->key_bytes;
    ec.flags = cfg->flags;
    ec.max_build_rows = eng_rows;
    ec.load_factor = std::min(0.9, (cfg->load_factor > 0 ? cfg->load_factor : 0.5) * slack);   // same slots as an unpadded table
    ec.hash_seed = cfg->hash_seed;
    if (int rc = dwj_create(&ec, &m->eng[r])) return bail(rc);
    xc = dwj_xj_config{};
    xc.rank = (int32_t)r;
    xc.world = (int32_t)m->n;
    xc.max_build_rows = cfg->max_build_rows_per_gpu;
    xc.max_probe_rows = cfg->max_probe_rows_per_gpu;
    xc.chunk_rows = cfg->chunk_rows;
    xc.passes = passes;
    xc.recv_slack = slack;
    xc.force_scatter_pull = cfg->force_scatter_pull;
    uint64_t bytes = 0;
    if (int rc = dwj_xj_block_bytes(m->eng[r], &xc, &bytes)) return bail(rc);
    cudaSetDevice(cfg->devices[r]);
    if (cudaMalloc(&m->block[r], bytes) != cudaSuccess) return bail(xfail(DWJ_ERR_OOM, "cudaMalloc of the %llu-byte exchange block on GPU %d failed", (unsigned long long)bytes, cfg->devices[r]));
    if (cudaMalloc((void **)&m->d_count[r], 64) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) != cudaSuccess)
      return bail(xfail(DWJ_ERR_OOM, "scratch allocation failed on GPU %d", cfg->devices[r]));
  }
  for (uint32_t r = 0; r < m->n; ++r) {
    xc.rank = (int32_t)r;
    if (int rc = dwj_xj_create(m->eng[r], &xc, m->block, &m->xj[r])) return bail(rc);
    m->xj[r]->host_barrier = host_barrier_arrive;
    m->xj[r]->host_barrier_ctx = &m->barrier;
  }
  *out = m;
  return DWJ_OK;
}

int dwj_mg_join(dwj_mg *m, const void *const *d_build_keys, const void *const *d_build_vals, const uint64_t *n_build,
                const void *const *d_probe_keys, const void *const *d_probe_vals, const uint64_t *n_probe, void *const *d_out_key,
                void *const *d_out_build_val, void *const *d_out_probe_val, const uint64_t *capacity, uint64_t *n_out, dwj_mg_timing *timing) {
  if (!m || !d_build_keys || !d_build_vals || !n_build || !d_probe_keys || !d_probe_vals || !n_probe || !d_out_build_val || !d_out_probe_val ||
      !capacity || !n_out)
    return xfail(DWJ_ERR_INVALID, "null argument");
  int rcs[MAX_WORLD] = {0};
  char msgs[MAX_WORLD][512];
  dwj_xj_timing xt[MAX_WORLD]{};
  auto run = [&](uint32_t r) {
    cudaSetDevice(m->cfg.devices[r]);
    msgs[r][0] = 0;
    int rc = dwj_xj_join(m->xj[r], d_build_keys[r], d_build_vals[r], n_build[r], d_probe_keys[r], d_probe_vals[r], n_probe[r],
                         d_out_key ? d_out_key[r] : nullptr, d_out_build_val[r], d_out_probe_val[r], capacity[r], (uint64_t *)m->d_count[r],
                         m->stream[r]);
    if (!rc) rc = dwj_xj_sync_timings(m->xj[r], &xt[r]);
    unsigned long long cnt = 0;
    if (!rc && cudaMemcpyAsync(&cnt, m->d_count[r], 8, cudaMemcpyDeviceToHost, m->stream[r]) == cudaSuccess && cudaStreamSynchronize(m->stream[r]) == cudaSuccess)
      n_out[r] = cnt;
    else if (!rc) rc = DWJ_ERR_CUDA;
    if (rc) {
      snprintf(msgs[r], sizeof(msgs[r]), "%s", dwj_last_error());
      m->barrier.abort();                        // the other ranks must not wait for this one
    }
    rcs[r] = rc;
  };
  m->barrier.reset(m->n);
  std::vector<std::thread> th;
  for (uint32_t r = 1; r < m->n; ++r) th.emplace_back(run, r);
  run(0);
  for (auto &t : th) t.join();
  {
    char all[480] = "";
    int first = 0;
    for (uint32_t r = 0; r < m->n; ++r)
      if (rcs[r]) {
        if (!first) first = rcs[r];
        const size_t at = strlen(all);
        snprintf(all + at, sizeof(all) - at, "%sGPU %u: %.150s", at ? "; " : "", r, msgs[r]);
      }
    if (first) return xfail(first, "%s", all);
  }
  for (uint32_t r = 0; r < m->n; ++r)
    if (n_out[r] > capacity[r]) return xfail(DWJ_ERR_OVERFLOW, "GPU %u produced %llu rows, output capacity is %llu", r, (unsigned long long)n_out[r], (unsigned long long)capacity[r]);
  if (timing) {
    *timing = dwj_mg_timing{};
    for (uint32_t r = 0; r < m->n; ++r) {     // the slowest rank per mark
      timing->counts_ms = std::max(timing->counts_ms, xt[r].counts_ms);
      timing->partition_ms = std::max(timing->partition_ms, xt[r].scattered_ms);
      timing->build_ms = std::max(timing->build_ms, xt[r].built_ms);
      timing->total_ms = std::max(timing->total_ms, xt[r].total_ms);
      timing->remote_bytes += xt[r].remote_bytes;
    }
  }
  return DWJ_OK;
}

// The whole multi-GPU join with HOST columns: row i of a relation goes to GPU i / ceil(n / n_gpus) (arrival order, not by
// key), H2D, dwj_mg_join, D2H of every GPU's result rows one after the other.  For the host framework and parity tests.
int dwj_mg_join_host(dwj_mg *m, const void *build_keys, const void *build_vals, uint64_t n_build, const void *probe_keys,
                     const void *probe_vals, uint64_t n_probe, void *out_key, void *out_build_val, void *out_probe_val, uint64_t out_capacity,
                     uint64_t *n_out, dwj_mg_timing *timing) {
  if (!m || !n_out) return xfail(DWJ_ERR_INVALID, "null argument");
  if ((n_build && (!build_keys || !build_vals)) || (n_probe && (!probe_keys || !probe_vals))) return xfail(DWJ_ERR_INVALID, "null input column");
  if (out_capacity && (!out_build_val || !out_probe_val)) return xfail(DWJ_ERR_INVALID, "null output column");
  const uint32_t n = m->n;
  const uint64_t W = (uint64_t)m->cfg.key_bytes;
  const uint64_t per_b = (n_build + n - 1) / n, per_p = (n_probe + n - 1) / n;
  if (per_b > m->cfg.max_build_rows_per_gpu || per_p > m->cfg.max_probe_rows_per_gpu)
    return xfail(DWJ_ERR_CAPACITY, "%llu build / %llu probe rows per GPU exceed the %llu / %llu this join was created for", (unsigned long long)per_b,
                 (unsigned long long)per_p, (unsigned long long)m->cfg.max_build_rows_per_gpu, (unsigned long long)m->cfg.max_probe_rows_per_gpu);
  const double slack = m->cfg.recv_slack > 0 ? m->cfg.recv_slack : 1.25;
  const bool with_key = out_key != nullptr;
  const uint64_t rows_b = std::max<uint64_t>(m->cfg.max_build_rows_per_gpu, 1), rows_p = std::max<uint64_t>(m->cfg.max_probe_rows_per_gpu, 1);
  uint64_t rows_o = (uint64_t)((double)rows_p * slack) + 1024;          // unique build keys: at most one row per received probe row
  if (!(m->cfg.flags & DWJ_FLAG_UNIQUE_BUILD_KEYS)) rows_o = std::max<uint64_t>(rows_o, out_capacity + 1024);   // any multiplicity: the caller's bound, on every GPU
  if (!m->stage[0] || m->stage_rows_b < rows_b || m->stage_rows_p < rows_p || m->stage_rows_o < rows_o) {
    for (uint32_t r = 0; r < n; ++r) {
      cudaSetDevice(m->cfg.devices[r]);
      if (m->stage[r]) { cudaDeviceSynchronize(); cudaFree(m->stage[r]); m->stage[r] = nullptr; }
      if (cudaMalloc(&m->stage[r], (2 * rows_b + 2 * rows_p + 3 * rows_o) * W) != cudaSuccess)
        return xfail(DWJ_ERR_OOM, "staging buffers on GPU %d: %s", m->cfg.devices[r], cudaGetErrorString(cudaGetLastError()));
    }
    m->stage_rows_b = rows_b; m->stage_rows_p = rows_p; m->stage_rows_o = rows_o;
  }
  const void *bk[MAX_WORLD], *bv[MAX_WORLD], *pk[MAX_WORLD], *pv[MAX_WORLD];
  void *ok[MAX_WORLD], *ob[MAX_WORLD], *op[MAX_WORLD];
  uint64_t nb[MAX_WORLD], np_[MAX_WORLD], cap[MAX_WORLD], got[MAX_WORLD] = {0};
  for (uint32_t r = 0; r < n; ++r) {
    cudaSetDevice(m->cfg.devices[r]);
    char *base = (char *)m->stage[r];
    const uint64_t b0 = std::min<uint64_t>(r * per_b, n_build), b1 = std::min<uint64_t>(b0 + per_b, n_build);
    const uint64_t p0 = std::min<uint64_t>(r * per_p, n_probe), p1 = std::min<uint64_t>(p0 + per_p, n_probe);
    nb[r] = b1 - b0; np_[r] = p1 - p0; cap[r] = m->stage_rows_o;
    bk[r] = base; bv[r] = base + rows_b * W; pk[r] = base + 2 * rows_b * W; pv[r] = base + (2 * rows_b + rows_p) * W;
    ob[r] = base + (2 * rows_b + 2 * rows_p) * W; op[r] = (char *)ob[r] + m->stage_rows_o * W; ok[r] = (char *)op[r] + m->stage_rows_o * W;
    if (nb[r]) {
      XCU(cudaMemcpyAsync((void *)bk[r], (const char *)build_keys + b0 * W, nb[r] * W, cudaMemcpyHostToDevice, m->stream[r]));
      XCU(cudaMemcpyAsync((void *)bv[r], (const char *)build_vals + b0 * W, nb[r] * W, cudaMemcpyHostToDevice, m->stream[r]));
    }
    if (np_[r]) {
      XCU(cudaMemcpyAsync((void *)pk[r], (const char *)probe_keys + p0 * W, np_[r] * W, cudaMemcpyHostToDevice, m->stream[r]));
      XCU(cudaMemcpyAsync((void *)pv[r], (const char *)probe_vals + p0 * W, np_[r] * W, cudaMemcpyHostToDevice, m->stream[r]));
    }
  }
  if (int rc = dwj_mg_join(m, bk, bv, nb, pk, pv, np_, with_key ? ok : nullptr, ob, op, cap, got, timing)) return rc;
  uint64_t total = 0;
  for (uint32_t r = 0; r < n; ++r) total += got[r];
  *n_out = total;
  if (total > out_capacity) return xfail(DWJ_ERR_OVERFLOW, "join produced %llu rows, output capacity is %llu", (unsigned long long)total, (unsigned long long)out_capacity);
  uint64_t at = 0;
  for (uint32_t r = 0; r < n; ++r) {
    cudaSetDevice(m->cfg.devices[r]);
    if (got[r]) {
      if (with_key) XCU(cudaMemcpyAsync((char *)out_key + at * W, ok[r], got[r] * W, cudaMemcpyDeviceToHost, m->stream[r]));
      XCU(cudaMemcpyAsync((char *)out_build_val + at * W, ob[r], got[r] * W, cudaMemcpyDeviceToHost, m->stream[r]));
      XCU(cudaMemcpyAsync((char *)out_probe_val + at * W, op[r], got[r] * W, cudaMemcpyDeviceToHost, m->stream[r]));
    }
    at += got[r];
  }
  for (uint32_t r = 0; r < n; ++r) {
    cudaSetDevice(m->cfg.devices[r]);
    XCU(cudaStreamSynchronize(m->stream[r]));
  }
  return DWJ_OK;
}

int dwj_mg_describe(const dwj_mg *m, uint32_t rank, dwj_xj_info *info) {
  if (!m || rank >= m->n) return xfail(DWJ_ERR_INVALID, "bad argument");
  return dwj_xj_describe(m->xj[rank], info);
}

}  // extern "C"
