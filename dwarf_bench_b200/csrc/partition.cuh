// partition.cuh -- radix partition of (key, payload) columns on the key hash.
//
// No reference counterpart (the reference is single-device).  Used for the multi-GPU
// exchange: rows with equal keys land in the same partition, partitions are contiguous, and the
// per-partition offsets drive an all-to-all-v.
//
// Three launches: histogram -> offsets (one block) -> scatter.  The scatter stages each tile in
// shared memory grouped by partition and reserves one contiguous output range per (tile,
// partition) with a single atomicAdd, so global writes are contiguous runs instead of a
// row-by-row scatter.
#pragma once
#include "table.cuh"

namespace dwj {

constexpr int PART_MAX = 256;
constexpr int PART_THREADS = 256;

template <int W> struct PartitionArgs {
  using K = typename KeyT<W>::type;
  const K *keys;
  const K *vals;       // may be null
  uint64_t n;
  uint32_t log2_parts;
  uint64_t seed;
  K *out_keys;
  K *out_vals;         // may be null
  unsigned long long *hist;     // [PART_MAX] zeroed before the histogram kernel
  unsigned long long *cursor;   // [PART_MAX] running write positions
  unsigned long long *offsets;  // [parts + 1] result
};

template <int W>
__global__ void __launch_bounds__(PART_THREADS) partition_hist_kernel(PartitionArgs<W> a) {
  __shared__ unsigned int s_hist[PART_MAX];
  const uint32_t parts = 1u << a.log2_parts;
  for (uint32_t p = threadIdx.x; p < parts; p += blockDim.x) s_hist[p] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
    atomicAdd(&s_hist[partition_of(load_stream(a.keys + i), a.log2_parts, a.seed)], 1u);
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < parts; p += blockDim.x)
    if (s_hist[p]) atomicAdd(a.hist + p, (unsigned long long)s_hist[p]);
}

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

template <int W, int ITEMS>
__global__ void __launch_bounds__(PART_THREADS) partition_scatter_kernel(PartitionArgs<W> a) {
  using K = typename KeyT<W>::type;
  constexpr int TILE = PART_THREADS * ITEMS;
  __shared__ K s_keys[TILE];
  __shared__ K s_vals[TILE];
  __shared__ unsigned int s_count[PART_MAX];       // rows of this tile per partition
  __shared__ unsigned int s_start[PART_MAX + 1];   // exclusive scan of s_count (staging offsets)
  __shared__ unsigned long long s_gbase[PART_MAX]; // reserved global start per partition

  const uint32_t parts = 1u << a.log2_parts;
  const uint64_t num_tiles = (a.n + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint64_t base = tile * TILE;
    const uint32_t rows = (uint32_t)min((uint64_t)TILE, a.n - base);
    for (uint32_t p = threadIdx.x; p < parts; p += PART_THREADS) s_count[p] = 0;
    __syncthreads();

    K k[ITEMS], v[ITEMS];
    uint32_t part[ITEMS], rank[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t r = j * PART_THREADS + threadIdx.x;
      if (r < rows) {
        k[j] = load_stream(a.keys + base + r);
        v[j] = a.vals ? load_stream(a.vals + base + r) : (K)0;
        part[j] = partition_of(k[j], a.log2_parts, a.seed);
        rank[j] = atomicAdd(&s_count[part[j]], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                         // parts <= 256: a serial scan is negligible
      unsigned int run = 0;
      for (uint32_t p = 0; p < parts; ++p) { s_start[p] = run; run += s_count[p]; }
      s_start[parts] = run;
    }
    for (uint32_t p = threadIdx.x; p < parts; p += PART_THREADS)
      s_gbase[p] = s_count[p] ? atomicAdd(a.cursor + p, (unsigned long long)s_count[p]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t r = j * PART_THREADS + threadIdx.x;
      if (r < rows) {
        const uint32_t s = s_start[part[j]] + rank[j];
        s_keys[s] = k[j];
        s_vals[s] = v[j];
      }
    }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < rows; s += PART_THREADS) {
      const K key = s_keys[s];
      const uint32_t p = partition_of(key, a.log2_parts, a.seed);
      const unsigned long long dst = s_gbase[p] + (s - s_start[p]);
      store_stream(a.out_keys + dst, key);
      if (a.out_vals) store_stream(a.out_vals + dst, s_vals[s]);
    }
    __syncthreads();
  }
}

}  // namespace dwj
