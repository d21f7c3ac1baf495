#include "b200_dwarfs.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <unordered_map>

#include <cuda_runtime_api.h>

#include "dwj.h"
#include "join/join_helpers/sort_join.hpp"

using namespace join_helpers;

// ---- engine plumbing ----------------------------------------------------------------------------------------
namespace b200 {

void Engine::check(int rc) {
  if (rc != DWJ_OK) throw std::runtime_error(std::string("dwj: ") + dwj_last_error());
}

Engine::Engine(size_t max_build_rows, unsigned flags, int key_bytes, double load_factor) {
  dwj_config cfg{};
  cfg.device = 0;
  cfg.key_bytes = key_bytes;
  cfg.payload_bytes = key_bytes;
  cfg.flags = flags;
  cfg.max_build_rows = max_build_rows;
  cfg.load_factor = load_factor > 0 ? load_factor : 0.5;     // join/join.cpp:30: ht_size = buf_size * 2
  cfg.hash_seed = helpers::make_random();      // join/join.cpp:32: seed = make_random()
  check(dwj_create(&cfg, &e_));
}

Engine::~Engine() { dwj_destroy(e_); }

void require_gpu(const RunOptions &opts, const std::string &dwarf) {
  if (opts.device_ty == RunOptions::DeviceType::CPU)
    throw std::logic_error(dwarf + ": this build serves --device=gpu only (B200 engine, no CPU fallback)");
}

}  // namespace b200

namespace {

// This tree's HashJoinResult reports build/probe times as extra CSV columns; against the reference's own
// framework (oracle/dropin_main.cpp) the two base columns are all there is.
#ifdef DWJ_HOST_FRAMEWORK
constexpr const char *kJoinReportHeader = "host_time_ms,kernel_time_ms,build_time_ms,probe_time_ms";
#else
constexpr const char *kJoinReportHeader = "host_time_ms,kernel_time_ms";
#endif

using Clock = std::chrono::steady_clock;
constexpr uint32_t empty_element = std::numeric_limits<uint32_t>::max();

void std_init(Dwarf &d, const RunOptions &opts) {          // join/join.cpp:150-154
  d.meter().set_opts(opts);
  d.meter().set_params(DwarfParams{{"device_type", to_string(opts.device_ty)}});
}

void announce_device() {
  static bool once = false;
  if (once) return;
  once = true;
  dwj_info info{};
  (void)info;
  std::cout << "Selected device: NVIDIA B200 join engine (libdwj_b200, ABI " << dwj_abi_version() << ")\n";
}

Duration ms(float v) { return Duration(static_cast<double>(v) * 1000.0); }

// One iteration of the Join flow (join/join.cpp:36-140): table of 2n slots, build, probe into probe-aligned
// sentinel-filled arrays, host compaction, comparison with the expected rows.
std::unique_ptr<HashJoinResult> join_iteration(b200::Engine &eng, const std::vector<uint32_t> &ak, const std::vector<uint32_t> &av,
                                               const std::vector<uint32_t> &bk, const std::vector<uint32_t> &bv,
                                               const ColJoinedTableTy<uint32_t, uint32_t, uint32_t> &expected) {
  const size_t n = bk.size();
  std::vector<uint32_t> key_out(n, empty_element), key_present_out(n, empty_element), val_out(n, empty_element);
  auto result = std::make_unique<HashJoinResult>();
  dwj_timing t{};
  uint64_t n_out = 0;
  const auto host_start = Clock::now();
  b200::Engine::check(dwj_join_host(eng.get(), ak.data(), av.data(), ak.size(), bk.data(), bv.data(), n, DWJ_OUT_ALIGNED,
                                    key_out.data(), key_present_out.data(), val_out.data(), n, &n_out, &t));
  const auto host_end = Clock::now();
  result->host_time = host_end - host_start;                // wall clock incl. the H2D/D2H the reference hides in buffers
  result->build_time = ms(t.h2d_ms + t.build_ms);           // build_end - host_start (join.cpp:112)
  result->probe_time = ms(t.total_ms - t.h2d_ms - t.build_ms);
  result->kernel_time = ms(t.total_ms);                     // device-event time of the whole call

  std::vector<uint32_t> res_k, res_present, res_val;        // join.cpp:119-129
  for (size_t i = 0; i < n; ++i)
    if (key_out[i] != empty_element) {
      res_k.push_back(key_out[i]);
      res_present.push_back(key_present_out[i]);
      res_val.push_back(val_out[i]);
    }
  result->iterations = 1;                                   // result.hpp:15-17: present in the reference, never filled there
  result->bytes = result->bytes_per_iteration = (2 * ak.size() + 2 * n + 3 * res_k.size()) * sizeof(uint32_t);   // columns in, rows out
#ifdef DWJ_HOST_FRAMEWORK      // extra fields exist only in this tree's HashJoinResult
  result->matches = res_k.size();
  result->tuples_per_second = (ak.size() + n) / (result->kernel_time.count() * 1e-6);
#endif
  const ColJoinedTableTy<uint32_t, uint32_t, uint32_t> output{res_k, {res_present, res_val}};
  if (!(output == expected)) {                              // join.cpp:130-136
    std::cerr << "Incorrect results" << std::endl;
    result->valid = false;
  }
  return result;
}

// GPUs the Join dwarfs use: RunOptions::gpus in this tree (`--gpus N`); under the reference's own headers (the drop-in
// build, oracle/dropin_main.cpp) RunOptions has no such field and the environment variable DWARF_BENCH_GPUS decides.
size_t gpus_of(const RunOptions &opts) {
  size_t n = 1;
#ifdef DWJ_HOST_FRAMEWORK
  n = opts.gpus;
#else
  (void)opts;
#endif
  if (const char *v = std::getenv("DWARF_BENCH_GPUS")) n = std::max<size_t>(n, std::strtoul(v, nullptr, 10));
  return std::max<size_t>(n, 1);
}

// The Join flow over N GPUs (dwj_mg_*): the host columns are dealt to the GPUs in arrival order, hash-partitioned and
// exchanged over NVLink, joined locally; the result comes back compacted, so the reference's host compaction loop
// (join.cpp:119-129) has nothing left to do.  Same check against the expected rows.
void join_run_multi_gpu(const size_t buf_size, Meter &meter, size_t gpus, const std::vector<uint32_t> &ak, const std::vector<uint32_t> &av,
                        const std::vector<uint32_t> &bk, const std::vector<uint32_t> &bv,
                        const ColJoinedTableTy<uint32_t, uint32_t, uint32_t> &expected) {
  const RunOptions &opts = meter.opts();
  dwj_mg_config cfg{};
  cfg.n_gpus = static_cast<int32_t>(gpus);
  for (size_t i = 0; i < gpus && i < 8; ++i) cfg.devices[i] = static_cast<int32_t>(i);
  cfg.key_bytes = 4;
  cfg.flags = DWJ_FLAG_UNIQUE_BUILD_KEYS;
  cfg.max_build_rows_per_gpu = cfg.max_probe_rows_per_gpu = std::max<size_t>((buf_size + gpus - 1) / gpus, 1);
  cfg.recv_slack = 1.5;                                     // small inputs spread unevenly over the ranks
  dwj_mg *mg = nullptr;
  b200::Engine::check(dwj_mg_create(&cfg, &mg));
  struct Guard { dwj_mg *m; ~Guard() { dwj_mg_destroy(m); } } guard{mg};
  for (unsigned it = 0; it < opts.iterations; ++it) {
    const size_t n = bk.size();
    std::vector<uint32_t> res_k(n), res_present(n), res_val(n);
    auto result = std::make_unique<HashJoinResult>();
    dwj_mg_timing t{};
    uint64_t n_out = 0;
    const auto host_start = Clock::now();
    b200::Engine::check(dwj_mg_join_host(mg, ak.data(), av.data(), ak.size(), bk.data(), bv.data(), n, res_k.data(), res_present.data(),
                                         res_val.data(), n, &n_out, &t));
    const auto host_end = Clock::now();
    result->host_time = host_end - host_start;
    result->build_time = ms(t.build_ms);                    // partition + exchange + local build, slowest GPU
    result->probe_time = ms(t.total_ms - t.build_ms);
    result->kernel_time = ms(t.total_ms);
    res_k.resize(n_out);
    res_present.resize(n_out);
    res_val.resize(n_out);
    result->iterations = 1;
    result->bytes = result->bytes_per_iteration = (2 * ak.size() + 2 * n + 3 * n_out) * sizeof(uint32_t);
#ifdef DWJ_HOST_FRAMEWORK
    result->matches = n_out;
    result->tuples_per_second = (ak.size() + n) / (result->kernel_time.count() * 1e-6);
#endif
    const ColJoinedTableTy<uint32_t, uint32_t, uint32_t> output{res_k, {res_present, res_val}};
    if (!(output == expected)) {
      std::cerr << "Incorrect results" << std::endl;
      result->valid = false;
    }
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}

void join_run(const size_t buf_size, Meter &meter) {
  const RunOptions &opts = meter.opts();
  const std::vector<uint32_t> table_a_keys = helpers::make_unique_random(buf_size);          // join.cpp:13-21
  const std::vector<uint32_t> table_a_values = helpers::make_unique_random(table_a_keys.size());
  const std::vector<uint32_t> table_b_keys = helpers::make_unique_random(buf_size);
  const std::vector<uint32_t> table_b_values = helpers::make_unique_random(table_b_keys.size());
  announce_device();
  const auto expected = sort_join<uint32_t, uint32_t, uint32_t>(table_a_keys, table_a_values, table_b_keys, table_b_values);
  if (const size_t gpus = gpus_of(opts); gpus > 1) {
    join_run_multi_gpu(buf_size, meter, gpus, table_a_keys, table_a_values, table_b_keys, table_b_values, expected);
    return;
  }
  b200::Engine eng(std::max<size_t>(buf_size, 1), DWJ_FLAG_UNIQUE_BUILD_KEYS);
  for (unsigned it = 0; it < opts.iterations; ++it) {
    auto result = join_iteration(eng, table_a_keys, table_a_values, table_b_keys, table_b_values, expected);
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}

// Device copies of host columns for the build-only / probe-only dwarfs.
struct DeviceColumn {
  void *ptr = nullptr;
  explicit DeviceColumn(size_t bytes);
  ~DeviceColumn();
};

}  // namespace

namespace {
DeviceColumn::DeviceColumn(size_t bytes) {
  if (cudaMalloc(&ptr, bytes ? bytes : 4) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
}
DeviceColumn::~DeviceColumn() { cudaFree(ptr); }
constexpr cudaMemcpyKind kH2D = cudaMemcpyHostToDevice, kD2H = cudaMemcpyDeviceToHost;
}  // namespace

// ---- Join / SlabJoin ------------------------------------------------------------------------------------------
Join::Join() : Dwarf("Join") { reporting_header_ = kJoinReportHeader; }
void Join::_run(const size_t buf_size, Meter &meter) { join_run(buf_size, meter); }
void Join::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void Join::init(const RunOptions &opts) { std_init(*this, opts); }

SlabJoin::SlabJoin() : Dwarf("SlabJoin") { reporting_header_ = kJoinReportHeader; }
void SlabJoin::_run(const size_t buf_size, Meter &meter) { join_run(buf_size, meter); }
void SlabJoin::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void SlabJoin::init(const RunOptions &opts) { std_init(*this, opts); }

// ---- build-only dwarfs: timed build, untimed has() check --------------------------------------------------------
// HashBuild (hash/hash_build.cpp:8-87) and its siblings differ in the reference by table type, fill and key generator;
// here one table serves them all, created at each dwarf's own load factor:
//   HashBuild            bitmask table, T = 2n                      (hash_build.cpp:17)        keys in [1,10000]
//   HashBuildNonBitmask  CAS-on-key table, T = n                    (hash_build_non_bitmask.cpp:19; 0.9 is this engine's cap)
//   SlabHashBuild        chained slabs at 60 % utilisation = 0.625  (slab_hash_build.cpp:11, slab_hash.hpp:30-58)
//   CuckooHashBuild      cuckoo table, T = 4n                       (cuckoo_hash_build.cpp:14)  unique keys
namespace {
void hash_build_run(const size_t buf_size, Meter &meter, double load_factor, bool unique_keys) {
  const RunOptions &opts = meter.opts();
  const std::vector<uint32_t> host_src = unique_keys ? helpers::make_unique_random(buf_size)      // cuckoo_hash_build.cpp:12
                                                     : helpers::make_random<uint32_t>(buf_size);  // hash_build.cpp:10-11
  announce_device();
  b200::Engine eng(std::max<size_t>(buf_size, 1), unique_keys ? DWJ_FLAG_UNIQUE_BUILD_KEYS : 0, 4, load_factor);
  DeviceColumn src(buf_size * 4), flags(buf_size * 4);
  for (unsigned it = 0; it < opts.iterations; ++it) {
    auto result = std::make_unique<Result>();
    const auto host_start = Clock::now();                                                    // :35, H2D inside as with sycl::buffer
    if (cudaMemcpy(src.ptr, host_src.data(), buf_size * 4, kH2D) != 0) throw std::runtime_error("cudaMemcpy H2D failed");
    b200::Engine::check(dwj_build(eng.get(), src.ptr, src.ptr, buf_size, nullptr));           // ht.insert(s[idx], s[idx]) :48
    cudaDeviceSynchronize();
    result->host_time = Clock::now() - host_start;
    dwj_timing t{};
    b200::Engine::check(dwj_timings(eng.get(), &t));
    result->kernel_time = ms(t.build_ms);
    dwj_info info{};
    b200::Engine::check(dwj_get_info(eng.get(), &info));
    result->iterations = 1;                                                                  // result.hpp:15-17, never filled by the reference
    result->bytes = result->bytes_per_iteration = buf_size * 2 * sizeof(uint32_t) + info.table_bytes;   // rows in, the table written
    std::vector<uint32_t> output(buf_size, 0);                                               // :61-81
    b200::Engine::check(dwj_probe_contains(eng.get(), src.ptr, buf_size, static_cast<uint32_t *>(flags.ptr), nullptr));
    if (cudaMemcpy(output.data(), flags.ptr, buf_size * 4, kD2H) != 0) throw std::runtime_error("cudaMemcpy D2H failed");
    if (std::any_of(output.begin(), output.end(), [](uint32_t f) { return f != 1; })) {
      std::cerr << "Incorrect results" << std::endl;
      result->valid = false;
    }
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}
}  // namespace

#define B200_BUILD_ONLY_DWARF(NAME, LOAD, UNIQUE)                                           \
  NAME::NAME() : Dwarf(#NAME) {}                                                            \
  void NAME::_run(const size_t buf_size, Meter &meter) { hash_build_run(buf_size, meter, LOAD, UNIQUE); } \
  void NAME::run(const RunOptions &opts) {                                                  \
    b200::require_gpu(opts, name());                                                        \
    for (auto size : opts.input_size) _run(size, meter());                                  \
  }                                                                                         \
  void NAME::init(const RunOptions &opts) { std_init(*this, opts); }
B200_BUILD_ONLY_DWARF(HashBuild, 0.5, false)
B200_BUILD_ONLY_DWARF(HashBuildNonBitmask, 0.9, false)
B200_BUILD_ONLY_DWARF(SlabHashBuild, 0.625, false)
B200_BUILD_ONLY_DWARF(CuckooHashBuild, 0.25, true)

// ---- GroupBy / GroupByCuda: hash aggregation (groupby/groupby.cpp:24-112) -----------------------------------------------
// keys in [0, groups_count), values in [1, 10000]; output[g] = SUM(value) of group g (uint32 wrap-around).  Timed as the
// reference: both "kernels" -- the aggregation (ht.add, :60-72) and the read-back (ht.at, :84-92) -- inside host_time.
// The expected sums are always checked (the reference only does under !NDEBUG).
namespace {
void groupby_run(const size_t buf_size, Meter &meter) {
  // As the reference (:26): the caller passes a GroupByRunOptions -- main.cpp:87-90, bench.cpp:80 and the dwarf tests do.
  const auto &base = static_cast<const GroupByRunOptions &>(meter.opts());
  const size_t groups_count = std::max<size_t>(base.groups_count, 1);
  const std::vector<uint32_t> host_src_vals = helpers::make_random<uint32_t>(buf_size);                        // :30-31
  const std::vector<uint32_t> host_src_keys = helpers::make_random<uint32_t>(buf_size, 0, groups_count - 1);   // :32-33
  std::vector<uint32_t> expected(groups_count, 0);                                                             // :8-19
  for (size_t i = 0; i < buf_size; ++i) expected[host_src_keys[i]] += host_src_vals[i];
  std::vector<uint32_t> present(groups_count, 0);
  for (uint32_t k : host_src_keys) present[k] = 1;
  std::vector<uint32_t> group_ids(groups_count);
  std::iota(group_ids.begin(), group_ids.end(), 0u);
  announce_device();
  b200::Engine eng(std::max<size_t>(groups_count, 1), DWJ_FLAG_UNIQUE_BUILD_KEYS);            // the table holds one slot per group
  DeviceColumn keys(buf_size * 4), vals(buf_size * 4), ids(groups_count * 4), ok(groups_count * 4), osum(groups_count * 4), oid(groups_count * 4);
  if (cudaMemcpy(ids.ptr, group_ids.data(), groups_count * 4, kH2D) != 0) throw std::runtime_error("cudaMemcpy H2D failed");
  for (unsigned it = 0; it < base.iterations; ++it) {
    auto result = std::make_unique<Result>();
    const auto host_start = Clock::now();                                                    // :59
    if (cudaMemcpy(keys.ptr, host_src_keys.data(), buf_size * 4, kH2D) != 0 || cudaMemcpy(vals.ptr, host_src_vals.data(), buf_size * 4, kH2D) != 0)
      throw std::runtime_error("cudaMemcpy H2D failed");
    b200::Engine::check(dwj_aggregate_sum(eng.get(), keys.ptr, vals.ptr, buf_size, nullptr));                  // ht.add(sk[idx], sv[idx]) :70
    b200::Engine::check(dwj_probe_aligned(eng.get(), ids.ptr, ids.ptr, groups_count, ok.ptr, osum.ptr, oid.ptr, nullptr));   // ht.at(...) :87-90
    std::vector<uint32_t> output(groups_count, 0), found(groups_count, 0);
    if (cudaMemcpy(output.data(), osum.ptr, groups_count * 4, kD2H) != 0 || cudaMemcpy(found.data(), ok.ptr, groups_count * 4, kD2H) != 0)
      throw std::runtime_error("cudaMemcpy D2H failed");
    result->host_time = Clock::now() - host_start;                                           // :93-98
    dwj_timing t{};
    b200::Engine::check(dwj_timings(eng.get(), &t));
    result->kernel_time = ms(t.build_ms + t.probe_ms);
    result->iterations = 1;
    result->bytes = result->bytes_per_iteration = buf_size * 2 * sizeof(uint32_t);
    for (size_t g = 0; g < groups_count; ++g) {
      const bool hit = found[g] != empty_element;                                            // a group no row fell into has no slot: sum 0
      if (hit != (present[g] != 0) || (hit ? output[g] : 0u) != expected[g]) {
        std::cerr << "Incorrect results" << std::endl;                                       // :102-105
        result->valid = false;
        break;
      }
    }
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}
}  // namespace

GroupBy::GroupBy() : Dwarf("GroupBy") {}
void GroupBy::_run(const size_t buf_size, Meter &meter) { groupby_run(buf_size, meter); }
void GroupBy::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void GroupBy::init(const RunOptions &opts) { std_init(*this, opts); }

GroupByCuda::GroupByCuda() : Dwarf("GroupByCuda") {}                                         // the name bench.cpp:22-23 asks for on GPU
void GroupByCuda::_run(const size_t buf_size, Meter &meter) { groupby_run(buf_size, meter); }
void GroupByCuda::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void GroupByCuda::init(const RunOptions &opts) { std_init(*this, opts); }

// ---- SlabProbe: untimed build, timed find (probe/slab_probe.cpp:9-107) ------------------------------------------
SlabProbe::SlabProbe() : Dwarf("SlabProbe") {}
void SlabProbe::_run(const size_t buf_size, Meter &meter) {
  const RunOptions &opts = meter.opts();
  const std::vector<uint32_t> host_src = helpers::make_unique_random(buf_size);              // slab_probe.cpp:16
  announce_device();
  b200::Engine eng(std::max<size_t>(buf_size, 1), DWJ_FLAG_UNIQUE_BUILD_KEYS);
  DeviceColumn src(buf_size * 4), flags(buf_size * 4);
  for (unsigned it = 0; it < opts.iterations; ++it) {
    if (cudaMemcpy(src.ptr, host_src.data(), buf_size * 4, kH2D) != 0) throw std::runtime_error("cudaMemcpy H2D failed");
    b200::Engine::check(dwj_build(eng.get(), src.ptr, src.ptr, buf_size, nullptr));           // ht.insert(s[i], s[i]) :56, untimed
    cudaDeviceSynchronize();
    auto result = std::make_unique<Result>();
    std::vector<uint32_t> output(buf_size, 0);
    const auto host_start = Clock::now();                                                    // :64
    b200::Engine::check(dwj_probe_contains(eng.get(), src.ptr, buf_size, static_cast<uint32_t *>(flags.ptr), nullptr));
    cudaDeviceSynchronize();
    result->host_time = Clock::now() - host_start;
    dwj_timing t{};
    b200::Engine::check(dwj_timings(eng.get(), &t));
    result->kernel_time = ms(t.probe_ms);
    if (cudaMemcpy(output.data(), flags.ptr, buf_size * 4, kD2H) != 0) throw std::runtime_error("cudaMemcpy D2H failed");
    if (std::any_of(output.begin(), output.end(), [](uint32_t f) { return f != 1; })) {
      std::cerr << "Incorrect results" << std::endl;
      result->valid = false;
    }
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}
void SlabProbe::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void SlabProbe::init(const RunOptions &opts) { std_init(*this, opts); }

// ---- JoinOmnisci[Cuda]: one-to-many join on row ids (join/join_omnisci.cpp:49-108) -------------------------------
namespace {
void omnisci_run(const size_t buf_size, Meter &meter) {
  const RunOptions &opts = meter.opts();
  const std::vector<uint32_t> table_a_keys = helpers::make_random<uint32_t>(buf_size);       // :53-58, values in [1,10000]
  const std::vector<uint32_t> table_b_keys = helpers::make_random<uint32_t>(buf_size);
  announce_device();
  // Payload = row id on both sides: the result rows (key, build row, probe row) carry exactly the
  // information of the reference's per-probe-row {ids + pos[slot], cnt[slot]} lists (omnisci_hashtable.hpp:188-190).
  std::vector<uint32_t> a_ids(buf_size), b_ids(buf_size);
  std::iota(a_ids.begin(), a_ids.end(), 0u);
  std::iota(b_ids.begin(), b_ids.end(), 0u);
  // Expected number of (build row, probe row) pairs = sum over keys of count_a * count_b.
  std::unordered_map<uint32_t, uint64_t> count_a;
  for (uint32_t k : table_a_keys) ++count_a[k];
  uint64_t expected_pairs = 0;
  for (uint32_t k : table_b_keys) {
    const auto it = count_a.find(k);
    if (it != count_a.end()) expected_pairs += it->second;
  }
  b200::Engine eng(std::max<size_t>(buf_size, 1), 0);
  DeviceColumn ak(buf_size * 4), av(buf_size * 4), bk(buf_size * 4), bv(buf_size * 4);
  DeviceColumn out_b(std::max<uint64_t>(expected_pairs, 1) * 4), out_p(std::max<uint64_t>(expected_pairs, 1) * 4);
  for (unsigned it = 0; it < opts.iterations; ++it) {
    auto result = std::make_unique<HashJoinResult>();
    const auto host_start = Clock::now();                                                    // :78
    cudaMemcpy(ak.ptr, table_a_keys.data(), buf_size * 4, kH2D);
    cudaMemcpy(av.ptr, a_ids.data(), buf_size * 4, kH2D);
    b200::Engine::check(dwj_build(eng.get(), ak.ptr, av.ptr, buf_size, nullptr));             // build_table + build_id_buffer
    cudaDeviceSynchronize();
    const auto build_end = Clock::now();                                                     // :84
    cudaMemcpy(bk.ptr, table_b_keys.data(), buf_size * 4, kH2D);
    cudaMemcpy(bv.ptr, b_ids.data(), buf_size * 4, kH2D);
    uint64_t pairs = 0;
    b200::Engine::check(dwj_probe_pairs(eng.get(), bk.ptr, bv.ptr, buf_size, nullptr, out_b.ptr, out_p.ptr, expected_pairs, nullptr,
                                        &pairs, nullptr));                                   // lookup :86
    const auto host_end = Clock::now();
    result->host_time = host_end - host_start;
    result->build_time = build_end - host_start;
    result->probe_time = host_end - build_end;
    dwj_timing t{};
    b200::Engine::check(dwj_timings(eng.get(), &t));
    result->kernel_time = ms(t.build_ms + t.probe_ms);
#ifdef DWJ_HOST_FRAMEWORK
    result->matches = pairs;
#endif
    // The reference checks this only in debug builds (join_omnisci.cpp:97-102); here always: every emitted pair joins
    // equal keys, and the pair count is the exact one-to-many total.
    bool ok = pairs == expected_pairs;
    if (ok && pairs) {
      std::vector<uint32_t> hb(pairs), hp(pairs);
      cudaMemcpy(hb.data(), out_b.ptr, pairs * 4, kD2H);
      cudaMemcpy(hp.data(), out_p.ptr, pairs * 4, kD2H);
      for (uint64_t i = 0; i < pairs && ok; ++i) ok = hb[i] < buf_size && hp[i] < buf_size && table_a_keys[hb[i]] == table_b_keys[hp[i]];
    }
    if (!ok) {
      std::cerr << "Incorrect results" << std::endl;
      result->valid = false;
    }
    meter.add_result(DwarfParams{{"buf_size", std::to_string(buf_size)}}, std::move(result));
  }
}
}  // namespace

JoinOmnisci::JoinOmnisci() : Dwarf("JoinOmnisci") { reporting_header_ = kJoinReportHeader; }
void JoinOmnisci::_run(const size_t buf_size, Meter &meter) { omnisci_run(buf_size, meter); }
void JoinOmnisci::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void JoinOmnisci::init(const RunOptions &opts) { std_init(*this, opts); }

JoinOmnisciCuda::JoinOmnisciCuda() : Dwarf("JoinOmnisciCuda") { reporting_header_ = kJoinReportHeader; }
void JoinOmnisciCuda::_run(const size_t buf_size, Meter &meter) { omnisci_run(buf_size, meter); }
void JoinOmnisciCuda::run(const RunOptions &opts) {
  b200::require_gpu(opts, name());
  for (auto size : opts.input_size) _run(size, meter());
}
void JoinOmnisciCuda::init(const RunOptions &opts) { std_init(*this, opts); }
