#include "common.hpp"

#include <atomic>
#include <climits>
#include <unistd.h>

namespace helpers {

uint64_t next_seed() {
  static std::atomic<uint64_t> calls{0};
  if (const char *fixed = std::getenv("DWARF_BENCH_SEED"))
    return std::strtoull(fixed, nullptr, 10) * 0x9E3779B97F4A7C15ull + calls.fetch_add(1);
  std::random_device rd;
  return (static_cast<uint64_t>(rd()) << 32) ^ rd();
}

std::vector<uint32_t> make_unique_random(size_t size) {
  constexpr size_t size_multiplier = 10;
  const uint64_t domain = std::min<uint64_t>(static_cast<uint64_t>(size) * size_multiplier, 0xFFFFFFFFull);
  std::vector<uint32_t> out;
  out.reserve(size);
  if (size == 0) return out;
  std::mt19937_64 gen(next_seed());
  std::uniform_int_distribution<uint64_t> dist(0, domain - 1);
  std::vector<bool> taken(domain, false);
  for (size_t have = 0; have < size;) {
    const uint64_t v = dist(gen);
    if (!taken[v]) {
      taken[v] = true;
      ++have;
    }
  }
  for (uint64_t v = 0; v < domain; ++v)     // ascending, as iterating the reference's std::set yields
    if (taken[v]) out.push_back(static_cast<uint32_t>(v));
  return out;
}

std::vector<int> make_random_uniform_binary(size_t size) {
  std::mt19937_64 gen(next_seed());
  std::vector<int> out(size);
  for (int &v : out) v = static_cast<int>(gen() & 1u);
  return out;
}

uint32_t make_random() {
  std::mt19937_64 gen(next_seed());
  return std::uniform_int_distribution<uint32_t>(1, 1000)(gen);
}

std::string get_kernels_root_env(const char *argv0) {
  if (const char *val = std::getenv("DWARF_BENCH_ROOT")) return val;
  char buf[PATH_MAX];
  const ssize_t n = ::readlink("/proc/self/exe", buf, sizeof(buf) - 1);
  std::string exe = n > 0 ? std::string(buf, static_cast<size_t>(n)) : std::string(argv0 ? argv0 : ".");
  const size_t slash = exe.find_last_of('/');
  return slash == std::string::npos ? "." : exe.substr(0, slash);
}

void set_dpcpp_filter_env_no_overwrite(const char *filter) { setenv("SYCL_DEVICE_FILTER", filter, 0); }

void set_dpcpp_filter_env(const RunOptions &opts) {
  if (opts.device_ty == RunOptions::DeviceType::GPU) set_dpcpp_filter_env_no_overwrite("cuda");
}

}  // namespace helpers
