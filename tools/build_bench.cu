// build_bench.cu -- development microbenchmark: what does one hash-table insert cost on B200, piece by piece?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build_bench tools/build_bench.cu
// Rows are synthetic: row i goes to bucket b(i) = region(i) * slice + random(i) % slice, where region(i) = i / (n / P)
// (P = 1: uniformly random over the whole table; P > 1: the order the engine's region partition produces).
// Table: 32-byte buckets of 4 x 8-byte slots; fill: one uint32 ticket counter per bucket; load factor 0.5.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__host__ __device__ inline uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

struct Args {
  unsigned long long *table;
  unsigned int *fill;
  const uint32_t *keys, *vals;
  uint64_t n, buckets, rows_per_region, slice_mask;
  unsigned long long *out;
};

__device__ __forceinline__ uint64_t bucket_of(const Args &a, uint64_t i) {
  const uint64_t region = i / a.rows_per_region;
  const uint64_t r = ((uint64_t)fmix32((uint32_t)i * 2654435761u + 17u) ^ ((uint64_t)fmix32((uint32_t)(i >> 7) + 99u) << 11)) & a.slice_mask;
  return region * (a.slice_mask + 1) + r;
}

// V: 0 ticket only (result used)   1 ticket + slot store   2 keys/vals loaded + ticket + slot store
//    3 slot store only (slot = i & 3)   4 RED on the counter (no result)   5 ticket + store, home bucket via LDG of keys only
//    6 V2 + one overflow hop when ticket >= 4
template <int V, int ROWS> __global__ void __launch_bounds__(256) kb(Args a) {
  const uint64_t base = (uint64_t)blockIdx.x * 256 * ROWS + threadIdx.x;
  uint32_t k[ROWS], v[ROWS], t[ROWS];
  uint64_t b[ROWS];
  unsigned long long acc = 0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    if (V == 2 || V == 6) { k[r] = i < a.n ? __ldcs(a.keys + i) : 0; v[r] = i < a.n ? __ldcs(a.vals + i) : 0; }
    else { k[r] = (uint32_t)i; v[r] = (uint32_t)i; }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    b[r] = bucket_of(a, (V == 2 || V == 6) ? (uint64_t)k[r] : i);
    t[r] = 0;
    if (i < a.n) {
      if (V == 0 || V == 1 || V == 2 || V == 6) t[r] = atomicAdd(a.fill + b[r], 1u);
      if (V == 4) atomicAdd(a.fill + b[r], 1u);
      if (V == 3) t[r] = (uint32_t)i;
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    if (i < a.n) {
      if (V == 1 || V == 2 || V == 3) a.table[(b[r] << 2) + (t[r] & 3)] = (unsigned long long)k[r] | ((unsigned long long)v[r] << 32);
      if (V == 6) {
        if (t[r] < 4) a.table[(b[r] << 2) + t[r]] = (unsigned long long)k[r] | ((unsigned long long)v[r] << 32);
      }
      if (V == 0) acc ^= t[r];
    }
  }
  if (V == 6) {
    uint32_t pending = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) if (base + (uint64_t)r * 256 < a.n && t[r] >= 4) pending |= 1u << r;
    while (pending) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) if (pending >> r & 1) { b[r] = (b[r] + 1) & (a.buckets - 1); t[r] = atomicAdd(a.fill + b[r], 1u); }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) if ((pending >> r & 1) && t[r] < 4) { a.table[(b[r] << 2) + t[r]] = (unsigned long long)k[r] | ((unsigned long long)v[r] << 32); pending &= ~(1u << r); }
    }
  }
  if (V == 0 && acc == 0x123456789ull) a.out[0] = acc;
}


// Variants with look-ahead on the NEXT region's slice (region-partitioned order only):
//   A = 1: prefetch.global.L2 of the chunk of slice r+1 that corresponds to this tile     (table + fill lines)
//   A = 2: CLEAR that chunk in-kernel with full-line stores (no global memset before the kernel; race ignored here)
// then the real thing (loaded keys + ticket + store + overflow hops).  STORE_ONLY drops the ticket.
template <int A, bool STORE_ONLY, int AHEAD> __global__ void __launch_bounds__(256) kc(Args a) {
  constexpr int ROWS = 4;
  const uint64_t tile = blockIdx.x;
  const uint64_t tiles_per_region = a.rows_per_region / (256 * ROWS);
  const uint64_t region = tile / tiles_per_region, j = tile % tiles_per_region;
  const uint64_t regions = a.buckets / (a.slice_mask + 1);
  if (region + AHEAD < regions) {
    const uint64_t slice_bytes = (a.slice_mask + 1) * 32, chunk = slice_bytes / tiles_per_region;       // table bytes per tile
    char *tb = (char *)a.table + (region + AHEAD) * slice_bytes + j * chunk;
    char *fb = (char *)a.fill + ((region + AHEAD) * slice_bytes + j * chunk) / 8;
    if (A == 1) {
      for (uint64_t o = threadIdx.x * 128ull; o < chunk; o += 256 * 128ull) asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + o));
      for (uint64_t o = threadIdx.x * 128ull; o < chunk / 8; o += 256 * 128ull) asm volatile("prefetch.global.L2 [%0];" ::"l"(fb + o));
    } else if (A == 2) {
      const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0, 0, 0, 0);
      for (uint64_t o = threadIdx.x * 16ull; o < chunk; o += 256 * 16ull) *(uint4 *)(tb + o) = ones;
      for (uint64_t o = threadIdx.x * 16ull; o < chunk / 8; o += 256 * 16ull) *(uint4 *)(fb + o) = zero;
    }
  }
  const uint64_t base = tile * 256 * ROWS + threadIdx.x;
  uint32_t k[ROWS], v[ROWS], t[ROWS];
  uint64_t b[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    k[r] = i < a.n ? __ldcs(a.keys + i) : 0; v[r] = i < a.n ? __ldcs(a.vals + i) : 0;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    b[r] = bucket_of(a, (uint64_t)k[r]);
    t[r] = 0;
    if (i < a.n) t[r] = STORE_ONLY ? (k[r] & 3) : atomicAdd(a.fill + b[r], 1u);
  }
  uint32_t pending = 0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const uint64_t i = base + (uint64_t)r * 256;
    if (i < a.n) {
      if (t[r] < 4) a.table[(b[r] << 2) + t[r]] = (unsigned long long)k[r] | ((unsigned long long)v[r] << 32);
      else pending |= 1u << r;
    }
  }
  while (pending) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) if (pending >> r & 1) { b[r] = (b[r] + 1) & (a.buckets - 1); t[r] = atomicAdd(a.fill + b[r], 1u); }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) if ((pending >> r & 1) && t[r] < 4) { a.table[(b[r] << 2) + t[r]] = (unsigned long long)k[r] | ((unsigned long long)v[r] << 32); pending &= ~(1u << r); }
  }
}

template <int A, bool STORE_ONLY, int AHEAD> void runc(Args a, const char *name, int reps) {
  const unsigned grid = (unsigned)((a.n + 1023) / 1024);
  const uint64_t slice_bytes = (a.slice_mask + 1) * 32;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9;
  for (int r = 0; r < reps; ++r) {
    if (A == 2) {   // poison everything, then clear only the first AHEAD slices: the kernel clears the rest itself
      CK(cudaMemsetAsync(a.table, 0x55, a.buckets * 32)); CK(cudaMemsetAsync(a.fill, 0x55, a.buckets * 4));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      CK(cudaMemsetAsync(a.table, 0xFF, slice_bytes * AHEAD)); CK(cudaMemsetAsync(a.fill, 0, slice_bytes * AHEAD / 8));
    } else {
      CK(cudaMemsetAsync(a.table, 0xFF, a.buckets * 32)); CK(cudaMemsetAsync(a.fill, 0, a.buckets * 4));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
    }
    kc<A, STORE_ONLY, AHEAD><<<grid, 256>>>(a); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r && ms < best) best = ms;
  }
  // sanity for the in-kernel clear: every counter must be < 2^20 and the total must be >= n (overflow hops add more)
  printf("    %-70s %8.3f ms  %6.1f G rows/s\n", name, best, a.n / best / 1e6);
}

__global__ void gen(uint32_t *keys, uint32_t *vals, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { keys[i] = (uint32_t)i; vals[i] = (uint32_t)(i * 3); }
}

template <int V, int ROWS> void run(Args a, const char *name, bool clear_each, int reps) {
  const unsigned grid = (unsigned)((a.n + 256 * ROWS - 1) / (256 * ROWS));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9, best_clear = 1e9;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    if (clear_each) { CK(cudaMemsetAsync(a.table, 0xFF, a.buckets * 32)); CK(cudaMemsetAsync(a.fill, 0, a.buckets * 4)); }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float msc; CK(cudaEventElapsedTime(&msc, e0, e1));
    CK(cudaEventRecord(e0)); kb<V, ROWS><<<grid, 256>>>(a); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r) { if (ms < best) best = ms; if (msc < best_clear) best_clear = msc; }
  }
  printf("    %-58s rows/thr %d  %8.3f ms  %6.1f G rows/s   (clear %.3f ms)\n", name, ROWS, best, a.n / best / 1e6, best_clear);
}

int main() {
  unsigned long long *out; CK(cudaMalloc(&out, 64));
  struct Cfg { uint64_t mb; uint64_t regions; };
  for (Cfg c : {Cfg{16, 1}, Cfg{32, 1}, Cfg{256, 1}, Cfg{256, 8}, Cfg{4096, 1}, Cfg{4096, 128}}) {
    const uint64_t bytes = c.mb << 20, buckets = bytes / 32, n = buckets * 2;     // 4 slots per bucket, load 0.5
    Args a{};
    CK(cudaMalloc(&a.table, bytes)); CK(cudaMalloc(&a.fill, buckets * 4));
    uint32_t *k, *v; CK(cudaMalloc(&k, n * 4)); CK(cudaMalloc(&v, n * 4));
    gen<<<1184, 256>>>(k, v, n);
    a.keys = k; a.vals = v; a.n = n; a.buckets = buckets; a.rows_per_region = n / c.regions; a.slice_mask = buckets / c.regions - 1; a.out = out;
    printf("table %llu MB, %llu rows, %llu region(s)\n", (unsigned long long)c.mb, (unsigned long long)n, (unsigned long long)c.regions);
    run<0, 4>(a, "ticket only (atomicAdd, result used)", true, 4);
    run<4, 4>(a, "RED on the counter (no result)", true, 4);
    run<3, 4>(a, "slot store only", true, 4);
    run<1, 4>(a, "ticket + slot store", true, 4);
    run<2, 4>(a, "keys/vals loaded + ticket + slot store", true, 4);
    run<2, 8>(a, "keys/vals loaded + ticket + slot store", true, 4);
    run<2, 2>(a, "keys/vals loaded + ticket + slot store", true, 4);
    run<2, 1>(a, "keys/vals loaded + ticket + slot store", true, 4);
    run<6, 4>(a, "loaded + ticket + store + overflow hops (the real thing)", true, 4);
    if (c.regions > 1) {
      runc<0, true, 1>(a, "loaded + store only, no look-ahead (kc baseline)", 4);
      runc<1, true, 1>(a, "loaded + store only, L2 prefetch of the next region's slice", 4);
      runc<0, false, 1>(a, "real thing, no look-ahead (kc baseline)", 4);
      runc<1, false, 1>(a, "real thing + L2 prefetch of the next region's slice", 4);
      runc<1, false, 2>(a, "real thing + L2 prefetch two regions ahead", 4);
      runc<2, false, 1>(a, "real thing + in-kernel CLEAR of the next region's slice (no memset; incl. first-slice clear)", 4);
      runc<2, false, 2>(a, "real thing + in-kernel CLEAR two regions ahead (no memset)", 4);
    }
    if (c.mb <= 32) {
      run<0, 4>(a, "ticket only, table NOT cleared between runs (L2-warm)", false, 4);
      run<1, 4>(a, "ticket + store, NOT cleared (L2-warm)", false, 4);
    }
    CK(cudaFree(a.table)); CK(cudaFree(a.fill)); CK(cudaFree(k)); CK(cudaFree(v));
  }
  return 0;
}
