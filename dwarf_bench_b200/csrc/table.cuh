// table.cuh -- slot / bucket layout and hash functions of the join table.
//
// Layout (HBM): an array of 32-byte BUCKETS, one bucket == one DRAM/L2 sector.
//   key_bytes 4: bucket = 4 slots of 8 B,  slot = key | payload << 32
//   key_bytes 8: bucket = 2 slots of 16 B, slot = { key, payload }
// The bucket count is a power of two; bucket(key) = hash(key) & (buckets - 1); a chain
// continues in the next bucket (linear probing at sector granularity).  A probe therefore
// costs exactly one 32 B sector (one LDG.256) unless the home bucket is full.
// Empty slots are all-ones (the reference's `empty_element`, join/join.cpp:10).
//
// This replaces SimpleNonOwningHashTable's three separate arrays (keys[], vals[], bitmask[],
// common/dpcpp/hashtable.hpp:59-66), which cost three dependent sectors per lookup.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dwj {

#if defined(__CUDACC__)
#define DWJ_HD __host__ __device__ __forceinline__
#define DWJ_D __device__ __forceinline__
#else
#define DWJ_HD inline
#define DWJ_D inline
#endif

// ---- hashing ---------------------------------------------------------------------------
// Slot hash: the Murmur3 finalisers (fmix32 is the reference's own finaliser,
// common/dpcpp/hashfunctions.hpp:77-85).  The full MurmurHash3_x86_32 body adds nothing for a
// single 4-byte block and the hash choice cannot change the match multiset.
DWJ_HD uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
DWJ_HD uint64_t fmix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

template <int W> struct KeyT;
template <> struct KeyT<4> { using type = uint32_t; };
template <> struct KeyT<8> { using type = uint64_t; };

// Bucket hash (low bits are used) and partition hash (top bits are used).  For 4-byte keys the
// partition hash is a second, differently-keyed mix so that partition id and bucket index are
// independent; for 8-byte keys one 64-bit mix has enough bits for both ends.
DWJ_HD uint64_t slot_hash(uint32_t key, uint64_t seed) { return fmix32(key ^ (uint32_t)seed); }
DWJ_HD uint64_t slot_hash(uint64_t key, uint64_t seed) { return fmix64(key ^ seed); }
DWJ_HD uint32_t part_hash32(uint32_t key, uint64_t seed) {
  return fmix32(key * 0x9E3779B1u + (uint32_t)(seed >> 32) + 0x7F4A7C15u);
}
DWJ_HD uint32_t partition_of(uint32_t key, uint32_t log2_parts, uint64_t seed) {
  return log2_parts ? part_hash32(key, seed) >> (32 - log2_parts) : 0u;
}
DWJ_HD uint32_t partition_of(uint64_t key, uint32_t log2_parts, uint64_t seed) {
  return log2_parts ? (uint32_t)(fmix64(key ^ seed) >> (64 - log2_parts)) : 0u;
}

// Pass filter (DWJ_OPT_PASS_FILTER): a join whose working set exceeds one GPU is run as several passes over key
// CLASSES -- bits of the partition hash below the destination-rank bits.  A row whose class is not the current one is
// treated as absent by the partition, histogram and build kernels.  mask == 0: no filter.
struct PassFilter {
  uint32_t shift, mask, want;
};
DWJ_HD bool pass_ok(uint32_t key, uint64_t seed, PassFilter f) {
  return !f.mask || ((part_hash32(key, seed) >> f.shift) & f.mask) == f.want;
}
DWJ_HD bool pass_ok(uint64_t key, uint64_t seed, PassFilter f) {
  return !f.mask || ((uint32_t)(fmix64(key ^ seed) >> f.shift) & f.mask) == f.want;
}

#if defined(__CUDACC__)

// ---- bucket access -----------------------------------------------------------------------
// A bucket in registers: `K f[..]` alternating key, payload (k0 v0 k1 v1 ...), exactly the 32 bytes of
// the sector.  A slot is empty when its KEY field is all-ones; the all-ones key is therefore reserved
// (it is the reference's `empty_element`, join/join.cpp:10, which the reference cannot join either:
// its compaction loop drops such rows, join.cpp:123-129).  Rows carrying it are ignored by the build
// and never match in a probe.
template <int W> struct Bucket;

template <> struct Bucket<4> {
  using key_t = uint32_t;
  static constexpr int SLOTS = 4;
  uint32_t f[8];
  DWJ_D bool empty(int i) const { return f[2 * i] == 0xFFFFFFFFu; }
  DWJ_D bool match(int i, uint32_t k) const { return f[2 * i] == k; }
  DWJ_D uint32_t payload(int i) const { return f[2 * i + 1]; }
  // Slots fill in order and are never freed, so the occupied slots of a bucket form a prefix.
  DWJ_D bool any_empty() const { return f[6] == 0xFFFFFFFFu; }
};

template <> struct Bucket<8> {
  using key_t = uint64_t;
  static constexpr int SLOTS = 2;
  unsigned long long f[4];
  DWJ_D bool empty(int i) const { return f[2 * i] == ~0ull; }
  DWJ_D bool match(int i, uint64_t k) const { return f[2 * i] == k; }
  DWJ_D uint64_t payload(int i) const { return f[2 * i + 1]; }
  DWJ_D bool any_empty() const { return f[2] == ~0ull; }
};

// One 32-byte sector in ONE instruction (LDG.E.256, sm_100+): measured on B200, a random-sector gather
// costs one L1TEX wavefront per instruction, so two 128-bit loads would halve the probe rate
// (tools/gather_bench.cu: 285 vs 143 G gathers/s on an L2-resident table).  Read-only path for the probe
// (the table is immutable while a probe kernel runs).  The line is allowed into L1: for uniform keys that
// neither helps nor hurts (same gather rate with and without L1::no_allocate), but with skewed probe keys
// (Zipf 1.0: 6 % of all rows ask for the same sector) every SM then serves the hot buckets from its own L1
// instead of all 148 SMs queueing on one L2 slice -- bypassing L1 made that workload 4x slower.
DWJ_D Bucket<4> load_bucket_ro(const void *table, uint64_t b, Bucket<4> *) {
  Bucket<4> r;
  const char *p = (const char *)table + (b << 5);
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.f[0]), "=r"(r.f[1]), "=r"(r.f[2]), "=r"(r.f[3]), "=r"(r.f[4]), "=r"(r.f[5]), "=r"(r.f[6]), "=r"(r.f[7])
               : "l"(p));
  return r;
}
DWJ_D Bucket<8> load_bucket_ro(const void *table, uint64_t b, Bucket<8> *) {
  Bucket<8> r;
  const char *p = (const char *)table + (b << 5);
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.f[0]), "=l"(r.f[1]), "=l"(r.f[2]), "=l"(r.f[3])
               : "l"(p));
  return r;
}
template <int W> DWJ_D Bucket<W> load_bucket_ro(const void *table, uint64_t b) {
  return load_bucket_ro(table, b, (Bucket<W> *)nullptr);
}
// L1-bypassing variant for the staged PAIRS kernel (unique build keys): that kernel keeps 32-48 KB of staging per
// CTA in shared memory and measured 20-60 % slower when the table sectors also competed for the L1 carve-out.
// It is therefore the one probe kernel that is not protected against hot probe keys (open items).
DWJ_D Bucket<4> load_bucket_stream(const void *table, uint64_t b, Bucket<4> *) {
  Bucket<4> r;
  const char *p = (const char *)table + (b << 5);
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.f[0]), "=r"(r.f[1]), "=r"(r.f[2]), "=r"(r.f[3]), "=r"(r.f[4]), "=r"(r.f[5]), "=r"(r.f[6]), "=r"(r.f[7])
               : "l"(p));
  return r;
}
DWJ_D Bucket<8> load_bucket_stream(const void *table, uint64_t b, Bucket<8> *) {
  Bucket<8> r;
  const char *p = (const char *)table + (b << 5);
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.f[0]), "=l"(r.f[1]), "=l"(r.f[2]), "=l"(r.f[3])
               : "l"(p));
  return r;
}
template <int W> DWJ_D Bucket<W> load_bucket_stream(const void *table, uint64_t b) {
  return load_bucket_stream(table, b, (Bucket<W> *)nullptr);
}
// ---- segmented input ---------------------------------------------------------------------------------------------------
// A relation may be handed to the build / probe / scatter kernels as a LIST OF SEGMENTS instead of one contiguous range
// (dwj_*_segments): every segment carries its own column pointers, so a segment may live anywhere -- in particular in a
// PEER GPU's memory mapped over NVLink.  That is how the multi-GPU join moves data: the sender groups its rows by
// destination in its own memory and the receiver's kernels read ("pull") their rows straight out of the senders'
// buffers, one segment per (table region, source rank), without an intermediate copy (dwj_xj.cu).  A kernel's work unit
// (tile / chunk) never straddles two segments: segment i owns units [first_unit[i], first_unit[i+1]).
struct Seg {
  unsigned long long first_unit;   // first tile / chunk of this segment (entry n_segs holds the total)
  const void *keys;                // first key of the segment
  const void *vals;                // first payload (may be null where the kernel takes no payloads)
  unsigned long long rows;
};
DWJ_D uint32_t find_segment(const Seg *segs, uint32_t n_segs, unsigned long long unit) {
  uint32_t lo = 0, hi = n_segs;                     // segs[lo].first_unit <= unit < segs[hi].first_unit
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&segs[mid].first_unit) <= unit) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- one-to-many runs (csr.cuh) ------------------------------------------------------------------------------------
// Engines without DWJ_FLAG_UNIQUE_BUILD_KEYS keep distinct keys in the bucket table; a slot's payload field is the index of
// the 32-byte GRANULE (one sector) where the key's run [count | payload 0 | payload 1 | ...] starts in the `runs` array.
template <int W> struct CsrGeom {
  static constexpr uint32_t G = 32 / W;                                    // words per granule
  DWJ_HD static uint64_t granules(uint64_t count) { return (count + G) / G; }  // header + count payloads, rounded up
};

// Count and first payloads of a key's run with ONE 32-byte load, like a bucket: the count and the first 7 (4-byte keys)
// or 3 (8-byte keys) payloads.  pl[] receives up to four payloads; longer runs are read by the emitting code.
template <class K> DWJ_D uint32_t csr_head(const K *run, K (&pl)[4]) {
  if constexpr (sizeof(K) == 4) {
    const Bucket<4> h = load_bucket_ro<4>(run, 0);
    pl[0] = h.f[1];
    pl[1] = h.f[2];
    pl[2] = h.f[3];
    pl[3] = h.f[4];
    return h.f[0];
  } else {
    const Bucket<8> h = load_bucket_ro<8>(run, 0);
    pl[0] = h.f[1];
    pl[1] = h.f[2];
    pl[2] = h.f[3];
    if (h.f[0] > 3) pl[3] = __ldg(run + 4);
    return (uint32_t)min(h.f[0], 0xFFFFFFFFull);
  }
}

// ---- streaming column access ---------------------------------------------------------------
template <class T> DWJ_D T load_stream(const T *p) { return __ldcs(p); }   // ld.global.cs: evict-first
template <class T> DWJ_D void store_stream(T *p, T v) { __stcs(p, v); }    // st.global.cs

#endif  // __CUDACC__

}  // namespace dwj
