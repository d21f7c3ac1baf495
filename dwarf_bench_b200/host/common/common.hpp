// common.hpp -- input generators and small helpers (reference: common/common.{hpp,cpp}).
//
// Same names and distributions as the reference; two deliberate differences: the generators can be seeded
// (environment variable DWARF_BENCH_SEED, otherwise std::random_device as in the reference, whose runs are
// not reproducible -- common.cpp:10,24,32), and make_unique_random uses a membership bitmap instead of a
// std::set (1.8 s per call at n = 1 Mi in the reference).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "dwarf_framework.hpp"

template <class Collection> void dump_collection(const Collection &c, std::ostream &os = std::cout) {
  const char *sep = "";
  for (const auto &e : c) {
    os << sep << e;
    sep = " ";
  }
  os << "\n";
}

namespace helpers {

// Seed source shared by all generators: DWARF_BENCH_SEED (+ a per-call counter) or std::random_device.
uint64_t next_seed();

// `size` sorted, unique values drawn uniformly from [0, 10 * size)   (common.cpp:7-20)
std::vector<uint32_t> make_unique_random(size_t size);

// `size` values uniform in [left, right]; the reference's defaults are [1, 10000]   (common.hpp:31-40)
template <class T> std::vector<T> make_random(size_t size, size_t random_range_left = 1, size_t random_range_right = 10000) {
  std::mt19937_64 gen(next_seed());
  std::uniform_int_distribution<T> dist(static_cast<T>(random_range_left), static_cast<T>(random_range_right));
  std::vector<T> out(size);
  for (T &v : out) v = dist(gen);
  return out;
}

uint32_t make_random();                                  // the hash seed, uniform in [1, 1000]   (common.cpp:30-36)
std::vector<int> make_random_uniform_binary(size_t size);
std::string get_kernels_root_env(const char *argv0);     // DWARF_BENCH_ROOT or the executable's directory
void set_dpcpp_filter_env_no_overwrite(const char *filter);
void set_dpcpp_filter_env(const RunOptions &opts);       // kept for source compatibility; there is no SYCL runtime here

template <typename T, typename U> bool check_first(const T &v1, const U &v2, size_t sz) {
  for (size_t i = 0; i < sz; ++i)
    if (v1[i] != v2[i]) return false;
  return true;
}

}  // namespace helpers
