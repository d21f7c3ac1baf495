// gather_bench.cu -- development microbenchmark: how fast can one B200 fetch random 32-byte sectors, and how many
// DRAM bytes does each load flavour really move?  (Run under `ncu --metrics dram__bytes_read.sum,...` to see the
// fetch granularity.)  Not part of the product; informs the probe kernel's load instruction.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }

template <int F> __device__ __forceinline__ unsigned long long gather(const char *p) {
  unsigned long long a = 0, b = 0, c = 0, d = 0;
  if (F == 0) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 2) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 3) { asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
                asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16)); }
  if (F == 4) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
  if (F == 5) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 6) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 7) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
  if (F == 8) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_normal.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if (F == 9) { asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(0x14F0000000000000ull));   // evict_last policy descriptor
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(c), "=l"(d) : "l"(p + 16), "l"(0x14F0000000000000ull)); }
  return a ^ b ^ c ^ d;
}

template <int F, int ITEMS> __global__ void __launch_bounds__(256) k(const char *table, uint64_t mask, uint64_t n, unsigned long long *out) {
  const uint64_t base = ((uint64_t)blockIdx.x * 256 * ITEMS) + threadIdx.x;
  unsigned long long acc = 0;
  unsigned long long v[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint64_t i = base + (uint64_t)j * 256;
    const uint64_t b = ((uint64_t)fmix32((uint32_t)i) ^ ((uint64_t)fmix32((uint32_t)(i >> 3) + 77u) << 20)) & mask;
    v[j] = i < n ? gather<F>(table + (b << 5)) : 0ull;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) acc ^= v[j];
  if (acc == 0x1234567ull) out[0] = acc;   // keep the loads alive
}

template <int F> float run(const char *table, uint64_t buckets, uint64_t n, unsigned long long *out, const char *name, double mb) {
  constexpr int ITEMS = 8;
  const unsigned grid = (unsigned)((n + 256 * ITEMS - 1) / (256 * ITEMS));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    CK(cudaEventRecord(a)); k<F, ITEMS><<<grid, 256>>>(table, buckets - 1, n, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (r && ms < best) best = ms;
  }
  printf("  table %6.0f MB  %-52s %7.3f ms  %6.1f G gathers/s  (%6.0f GB/s of 32 B sectors)\n", mb, name, best, n / best / 1e6, n * 32.0 / best / 1e6);
  return best;
}

// ---- random atomics on an L2-resident table: what does one insert cost? ------------------------------------------
template <int F, int ITEMS> __global__ void __launch_bounds__(256) ka(char *table, uint64_t mask, uint64_t n, unsigned long long *out) {
  const uint64_t base = ((uint64_t)blockIdx.x * 256 * ITEMS) + threadIdx.x;
  unsigned long long acc = 0, r[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint64_t i = base + (uint64_t)j * 256;
    const uint64_t b = ((uint64_t)fmix32((uint32_t)i) ^ ((uint64_t)fmix32((uint32_t)(i >> 3) + 77u) << 20)) & mask;
    char *p = table + (b << 5) + ((i & 3) << 3);
    r[j] = 0;
    if (i < n) {
      if (F == 0) r[j] = atomicCAS((unsigned long long *)p, ~0ull, i);                       // 64-bit CAS, result used
      if (F == 1) r[j] = atomicCAS((unsigned int *)p, ~0u, (unsigned)i);                     // 32-bit CAS, result used
      if (F == 2) r[j] = atomicExch((unsigned long long *)p, i);                             // 64-bit exchange
      if (F == 3) atomicOr((unsigned int *)p, 1u << (i & 31));                               // 32-bit RED (no result)
      if (F == 4) *(volatile unsigned long long *)p = i;                                     // plain 64-bit store
      if (F == 5) { unsigned __int128 *q = (unsigned __int128 *)(table + (b << 5) + ((i & 1) << 4));
                    r[j] = (unsigned long long)atomicCAS(q, ~(unsigned __int128)0, (unsigned __int128)i); }   // 128-bit CAS
    }
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) acc ^= r[j];
  if (acc == 0x1234567ull) out[0] = acc;
}

template <int F> void run_atomic(char *table, uint64_t buckets, uint64_t n, unsigned long long *out, const char *name, double mb) {
  constexpr int ITEMS = 4;
  const unsigned grid = (unsigned)((n + 256 * ITEMS - 1) / (256 * ITEMS));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    CK(cudaMemset(table, 0xFF, buckets * 32));
    CK(cudaEventRecord(a)); ka<F, ITEMS><<<grid, 256>>>(table, buckets - 1, n, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (r && ms < best) best = ms;
  }
  printf("  table %6.0f MB  %-52s %7.3f ms  %6.1f G ops/s\n", mb, name, best, n / best / 1e6);
}

int main(int argc, char **argv) {
  if (argc > 1 && argv[1][0] == 'a') {            // `gather_bench atomics`
    const uint64_t n = 1ull << 26;
    unsigned long long *out; CK(cudaMalloc(&out, 64));
    for (uint64_t mb : {32ull, 1024ull}) {
      const uint64_t bytes = mb << 20, buckets = bytes / 32;
      char *table; CK(cudaMalloc(&table, bytes));
      run_atomic<0>(table, buckets, n, out, "atomicCAS 64-bit (result used)", (double)mb);
      run_atomic<1>(table, buckets, n, out, "atomicCAS 32-bit (result used)", (double)mb);
      run_atomic<5>(table, buckets, n, out, "atomicCAS 128-bit (result used)", (double)mb);
      run_atomic<2>(table, buckets, n, out, "atomicExch 64-bit", (double)mb);
      run_atomic<3>(table, buckets, n, out, "atomicOr 32-bit, no result (RED)", (double)mb);
      run_atomic<4>(table, buckets, n, out, "plain 64-bit store", (double)mb);
      CK(cudaFree(table));
    }
    return 0;
  }
  const uint64_t n = 1ull << 27;
  size_t gran = 0;
  CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", gran);
  if (argc > 1) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]))); CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity)); printf("now %zu\n", gran); }
  unsigned long long *out; CK(cudaMalloc(&out, 64));
  for (uint64_t mb : {32ull, 256ull, 2048ull}) {
    const uint64_t bytes = mb << 20, buckets = bytes / 32;
    char *table; CK(cudaMalloc(&table, bytes)); CK(cudaMemset(table, 0xAB, bytes));
    run<0>(table, buckets, n, out, "256b nc L1::no_allocate L2::evict_last", (double)mb);
    run<1>(table, buckets, n, out, "256b nc L1::no_allocate", (double)mb);
    run<2>(table, buckets, n, out, "256b nc", (double)mb);
    run<8>(table, buckets, n, out, "256b nc L1::no_allocate L2::evict_normal", (double)mb);
    run<5>(table, buckets, n, out, "256b nc L1::no_allocate L2::evict_first", (double)mb);
    run<6>(table, buckets, n, out, "256b cg", (double)mb);
    run<3>(table, buckets, n, out, "2 x 128b nc L1::no_allocate", (double)mb);
    run<9>(table, buckets, n, out, "2 x 128b nc L1::no_allocate L2::cache_hint(evict_last)", (double)mb);
    run<4>(table, buckets, n, out, "1 x 128b nc L1::no_allocate (half sector)", (double)mb);
    run<7>(table, buckets, n, out, "1 x 64b nc", (double)mb);
    CK(cudaFree(table));
  }
  return 0;
}
