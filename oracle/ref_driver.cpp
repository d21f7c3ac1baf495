// ref_driver.cpp -- thin extern "C" driver over the REFERENCE'S OWN headers.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (target `ref`) into
// oracle/_ref/libref_join.so, compiling the reference sources where they lie
// under $(REF) (default /root/reference) -- nothing from the reference is
// copied into this repository:
//   join/join_helpers/join_helpers.hpp   seq_join, operator== (the result definition)
//   common/dpcpp/hashfunctions.hpp       MurmurHash3_x86_32, StaticSimpleHasher
//   common/dpcpp/hashtable.hpp           SimpleNonOwningHashTable (insert / at / has)
// against oracle/sycl_shim/CL/sycl.hpp (no SYCL compiler in this image).
//
// What it is used for:
//   * pinning oracle/join_oracle.c (tests/test_oracle_vs_reference.py);
//   * the CPU baseline of bench.py (`cpu_baseline.kind == "reference"`): the
//     two parallel_for bodies of join/join.cpp:69-75 and :93-103 are executed
//     verbatim in shape -- a SimpleNonOwningHashTable view constructed per work
//     item, ht.insert / ht.at -- with the SYCL CPU device's work-item loop
//     replaced by an OpenMP `parallel for` (the only part that is ours).
#include <climits>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <cmath>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common/dpcpp/hashtable.hpp"          // also pulls hashfunctions.hpp
#include "join/join_helpers/join_helpers.hpp"

namespace {

// Runtime-sized `v % size` functor with the same body as StaticSimpleHasher
// (hashfunctions.hpp:33-35) -- SimpleHasher<uint32_t> (hashfunctions.hpp:43-49)
// is exactly that and is the reference's own type.
using ModuloHasher = SimpleHasher<uint32_t>;

template <class Hash>
using RefTable = SimpleNonOwningHashTable<uint32_t, uint32_t, Hash>;

template <class Hash>
void insert_all(uint64_t size, uint64_t bitmask_sz, uint32_t *keys, uint32_t *vals,
                uint32_t *bitmask, Hash hasher, const uint32_t *in_k, const uint32_t *in_v,
                uint64_t n, uint32_t *slots, bool parallel) {
  if (parallel) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      RefTable<Hash> ht(size, bitmask_sz, keys, vals, bitmask, hasher);
      auto r = ht.insert(in_k[i], in_v[i]);
      if (slots) slots[i] = r.first;
    }
  } else {
    for (uint64_t i = 0; i < n; ++i) {
      RefTable<Hash> ht(size, bitmask_sz, keys, vals, bitmask, hasher);
      auto r = ht.insert(in_k[i], in_v[i]);
      if (slots) slots[i] = r.first;
    }
  }
}

template <class Hash>
void at_all(uint64_t size, uint64_t bitmask_sz, uint32_t *keys, uint32_t *vals,
            uint32_t *bitmask, Hash hasher, const uint32_t *q, uint64_t n, uint32_t *found,
            uint32_t *val, uint32_t *has) {
  for (uint64_t i = 0; i < n; ++i) {
    RefTable<Hash> ht(size, bitmask_sz, keys, vals, bitmask, hasher);
    auto r = ht.at(q[i]);
    found[i] = r.second;
    val[i] = r.second ? r.first : 0;
    has[i] = ht.has(q[i]);
  }
}

} // namespace

extern "C" {

int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

uint64_t ref_murmur_slot(uint32_t v, uint32_t seed, int len, uint64_t sz) {
  return MurmurHash3_x86_32(sz, len, seed)(v);
}

// join_helpers::seq_join + copy-out.  Returns the row count.
uint64_t ref_seq_join_u32(const uint32_t *ak, const uint32_t *av, uint64_t na,
                          const uint32_t *bk, const uint32_t *bv, uint64_t nb, uint32_t *ok,
                          uint32_t *oa, uint32_t *ob, uint64_t cap) {
  std::vector<uint32_t> a_keys(ak, ak + na), a_vals(av, av + na), b_keys(bk, bk + nb),
      b_vals(bv, bv + nb);
  auto res = join_helpers::seq_join<uint32_t, uint32_t, uint32_t>(a_keys, a_vals, b_keys, b_vals);
  uint64_t m = join_helpers::get_size(res);
  for (uint64_t i = 0; i < m && i < cap; ++i) {
    ok[i] = res.first[i];
    oa[i] = res.second.first[i];
    ob[i] = res.second.second[i];
  }
  return m;
}

uint64_t ref_seq_join_u64(const uint64_t *ak, const uint64_t *av, uint64_t na,
                          const uint64_t *bk, const uint64_t *bv, uint64_t nb, uint64_t *ok,
                          uint64_t *oa, uint64_t *ob, uint64_t cap) {
  std::vector<uint64_t> a_keys(ak, ak + na), a_vals(av, av + na), b_keys(bk, bk + nb),
      b_vals(bv, bv + nb);
  auto res = join_helpers::seq_join<uint64_t, uint64_t, uint64_t>(a_keys, a_vals, b_keys, b_vals);
  uint64_t m = join_helpers::get_size(res);
  for (uint64_t i = 0; i < m && i < cap; ++i) {
    ok[i] = res.first[i];
    oa[i] = res.second.first[i];
    ob[i] = res.second.second[i];
  }
  return m;
}

// ColJoinedTableTy operator== (sort-based, order-insensitive).
int ref_rows_equal_u32(const uint32_t *k1, const uint32_t *a1, const uint32_t *b1, uint64_t n1,
                       const uint32_t *k2, const uint32_t *a2, const uint32_t *b2, uint64_t n2) {
  using namespace join_helpers;
  auto t1 = zip<uint32_t, uint32_t, uint32_t>({k1, k1 + n1}, {a1, a1 + n1}, {b1, b1 + n1});
  auto t2 = zip<uint32_t, uint32_t, uint32_t>({k2, k2 + n2}, {a2, a2 + n2}, {b2, b2 + n2});
  return t1 == t2;
}

// Row<->column round trip (tests/join_tests.cpp:44-59).
int ref_roundtrip_equal_u32(const uint32_t *k, const uint32_t *a, const uint32_t *b, uint64_t n) {
  using namespace join_helpers;
  auto t = zip<uint32_t, uint32_t, uint32_t>({k, k + n}, {a, a + n}, {b, b + n});
  return t == to_col_store(to_row_store(t));
}

// hash_kind: 0 = MurmurHash3_x86_32(size, 4, seed), 1 = v % size.
// Inserts (in_k[i], in_v[i]) for i < n into caller-owned, caller-initialised
// arrays; slots[i] (optional) receives the slot insert() returned.
void ref_table_insert(int hash_kind, uint32_t seed, uint64_t size, uint64_t bitmask_sz,
                      uint32_t *keys, uint32_t *vals, uint32_t *bitmask, const uint32_t *in_k,
                      const uint32_t *in_v, uint64_t n, uint32_t *slots, int parallel) {
  if (hash_kind == 1)
    insert_all(size, bitmask_sz, keys, vals, bitmask, ModuloHasher(size), in_k, in_v, n, slots,
               parallel != 0);
  else
    insert_all(size, bitmask_sz, keys, vals, bitmask, MurmurHash3_x86_32(size, 4, seed), in_k,
               in_v, n, slots, parallel != 0);
}

void ref_table_at(int hash_kind, uint32_t seed, uint64_t size, uint64_t bitmask_sz,
                  uint32_t *keys, uint32_t *vals, uint32_t *bitmask, const uint32_t *q,
                  uint64_t n, uint32_t *found, uint32_t *val, uint32_t *has) {
  if (hash_kind == 1)
    at_all(size, bitmask_sz, keys, vals, bitmask, ModuloHasher(size), q, n, found, val, has);
  else
    at_all(size, bitmask_sz, keys, vals, bitmask, MurmurHash3_x86_32(size, 4, seed), q, n, found,
           val, has);
}

// One iteration of Join::_run's loop body (join/join.cpp:36-113) over the
// reference's table type.  timing_us = {build, probe, host}.
int ref_join_build_probe_u32(const uint32_t *ak, const uint32_t *av, uint64_t na,
                             const uint32_t *bk, const uint32_t *bv, uint64_t nb,
                             uint32_t murmur_seed, uint32_t *out_key, uint32_t *out_present,
                             uint32_t *out_val, double *timing_us) {
  constexpr uint32_t empty_element = std::numeric_limits<uint32_t>::max();
  const size_t ht_size = na * 2;
  const size_t bitmask_sz = std::ceil((float)ht_size / 32);
  MurmurHash3_x86_32 hasher(ht_size, sizeof(uint32_t), murmur_seed);
  std::vector<uint32_t> bitmask(bitmask_sz ? bitmask_sz : 1, 0);
  std::vector<uint32_t> data(ht_size ? ht_size : 1, 0);
  std::vector<uint32_t> keys(ht_size ? ht_size : 1, empty_element);
  std::fill(out_key, out_key + nb, empty_element);
  std::fill(out_present, out_present + nb, empty_element);
  std::fill(out_val, out_val + nb, empty_element);
  uint32_t *kp = keys.data(), *dp = data.data(), *mp = bitmask.data();

  auto host_start = std::chrono::steady_clock::now();
  if (na) {
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < (int64_t)na; ++idx) {
      SimpleNonOwningHashTable<uint32_t, uint32_t, MurmurHash3_x86_32> ht(ht_size, bitmask_sz, kp,
                                                                         dp, mp, hasher);
      ht.insert(ak[idx], av[idx]);
    }
  }
  auto build_end = std::chrono::steady_clock::now();
  if (na) {
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < (int64_t)nb; ++idx) {
      SimpleNonOwningHashTable<uint32_t, uint32_t, MurmurHash3_x86_32> ht(ht_size, bitmask_sz, kp,
                                                                         dp, mp, hasher);
      auto ans = ht.at(bk[idx]);
      if (ans.second) {
        out_key[idx] = bk[idx];
        out_present[idx] = ans.first;
        out_val[idx] = bv[idx];
      }
    }
  }
  auto host_end = std::chrono::steady_clock::now();
  using us = std::chrono::duration<double, std::micro>;
  if (timing_us) {
    timing_us[0] = us(build_end - host_start).count();
    timing_us[1] = us(host_end - build_end).count();
    timing_us[2] = us(host_end - host_start).count();
  }
  return 0;
}

// The same loop body over 64-bit keys and payloads (BASELINE configs 4-5 have no reference counterpart: the reference
// Join is uint32-only).  The reference's OWN templates instantiated at 64 bits: SimpleNonOwningHashTable<uint64_t,
// uint64_t, SimpleHasher<uint64_t>> (hashtable.hpp:5-93, hashfunctions.hpp:43-49 -- Murmur3_x86_32 takes 32-bit input,
// SimpleHasher is the reference's hasher that accepts any integral key), T = 2n slots, sentinel-filled probe-aligned
// outputs, timed as join.cpp:59-113.  Slot indices are uint32_t in the reference (hashtable.hpp:16): na < 2^31.
int ref_join_build_probe_u64(const uint64_t *ak, const uint64_t *av, uint64_t na,
                             const uint64_t *bk, const uint64_t *bv, uint64_t nb,
                             uint64_t *out_key, uint64_t *out_present, uint64_t *out_val,
                             double *timing_us) {
  constexpr uint64_t empty_element = std::numeric_limits<uint64_t>::max();
  if (na >= (1ull << 31)) return -1;
  const size_t ht_size = na * 2;
  const size_t bitmask_sz = std::ceil((float)ht_size / 32);
  SimpleHasher<uint64_t> hasher(ht_size ? ht_size : 1);
  std::vector<uint32_t> bitmask(bitmask_sz ? bitmask_sz : 1, 0);
  std::vector<uint64_t> data(ht_size ? ht_size : 1, 0);
  std::vector<uint64_t> keys(ht_size ? ht_size : 1, empty_element);
  std::fill(out_key, out_key + nb, empty_element);
  std::fill(out_present, out_present + nb, empty_element);
  std::fill(out_val, out_val + nb, empty_element);
  uint64_t *kp = keys.data(), *dp = data.data();
  uint32_t *mp = bitmask.data();
  using Table = SimpleNonOwningHashTable<uint64_t, uint64_t, SimpleHasher<uint64_t>>;
  auto host_start = std::chrono::steady_clock::now();
  if (na) {
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < (int64_t)na; ++idx) {
      Table ht(ht_size, bitmask_sz, kp, dp, mp, hasher);
      ht.insert(ak[idx], av[idx]);
    }
  }
  auto build_end = std::chrono::steady_clock::now();
  if (na) {
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < (int64_t)nb; ++idx) {
      Table ht(ht_size, bitmask_sz, kp, dp, mp, hasher);
      auto ans = ht.at(bk[idx]);
      if (ans.second) {
        out_key[idx] = bk[idx];
        out_present[idx] = ans.first;
        out_val[idx] = bv[idx];
      }
    }
  }
  auto host_end = std::chrono::steady_clock::now();
  using us = std::chrono::duration<double, std::micro>;
  if (timing_us) {
    timing_us[0] = us(build_end - host_start).count();
    timing_us[1] = us(host_end - build_end).count();
    timing_us[2] = us(host_end - host_start).count();
  }
  return 0;
}
} // extern "C"
