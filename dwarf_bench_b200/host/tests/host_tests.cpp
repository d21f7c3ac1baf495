// host_tests -- C++ tests of the host framework, written to read like the reference's own:
//   `host_tests cpu`  tests/join_tests.cpp (HelpersSeqJoin / HelpersEqual / HelpersConvert), report formats,
//                     option parsing, registry, generators, error behaviour -- no GPU needed
//   `host_tests gpu`  tests/dwarf_tests/dwarf_tests.cpp: every dwarf x {128..4096} x 10 iterations, all valid
// A tiny assert harness stands in for gtest (FetchContent needs a network).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <set>
#include <sstream>

#include "bench.hpp"
#include "common/common.hpp"
#include "common/registry.hpp"
#include "join/b200_dwarfs.hpp"
#include "join/join_helpers/sort_join.hpp"
#include "register_dwarfs.hpp"

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                   \
  do {                                                                                \
    ++g_checks;                                                                       \
    if (!(cond)) {                                                                    \
      ++g_failed;                                                                     \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
    }                                                                                 \
  } while (0)

struct StdoutCapture {   // tests/dwarf_tests/utils.cpp:59-64
  std::stringstream buffer;
  std::streambuf *old;
  StdoutCapture() : old(std::cout.rdbuf(buffer.rdbuf())) {}
  ~StdoutCapture() { std::cout.rdbuf(old); }
};

// ---- tests/join_tests.cpp ---------------------------------------------------------------------------------------
static void join_helpers_tests() {
  using namespace std;
  using namespace join_helpers;
  vector<int> keys_a = {1, 2, 3, 4, 5, 5, 7}, vals_a = {5, 1, 4, 6, 6, 5, 0};
  vector<int> keys_b = {6, 2, 3, 4, 5, 5, 7}, vals_b = {3, 2, 1, 1, 3, 8, 8};
  auto res = seq_join(keys_a, vals_a, keys_b, vals_b);                 // HelpersSeqJoin
  CHECK(res.first.size() == res.second.first.size());
  CHECK(res.first.size() == res.second.second.size());
  CHECK(res.first.size() == 8);
  auto again = seq_join(keys_a, vals_a, keys_b, vals_b);               // HelpersEqual
  CHECK(res == again);
  CHECK(to_row_store(res) == to_row_store(again));
  CHECK(res == to_col_store(to_row_store(res)));                       // HelpersConvert
  auto sorted = sort_join(keys_a, vals_a, keys_b, vals_b);             // the O(n log n) check the dwarfs use
  CHECK(sorted == res);
  CHECK(get_size(sorted) == 8);
  auto rows = to_row_store(res);
  std::reverse(rows.begin(), rows.end());
  CHECK(to_col_store(rows) == res);                                    // order-insensitive
  rows[0].second.second += 1;
  CHECK(!(to_col_store(rows) == res));
  ColJoinedTableTy<int, int, int> bad{{1, 2}, {{1}, {1, 2}}};
  bool threw = false;
  try {
    get_size(bad);
  } catch (const std::invalid_argument &) {
    threw = true;
  }
  CHECK(threw);
  // random inputs with duplicates: sort_join is the same multiset as seq_join
  setenv("DWARF_BENCH_SEED", "7", 1);
  auto ka = helpers::make_random<uint32_t>(700, 1, 90), kb = helpers::make_random<uint32_t>(500, 1, 90);
  auto va = helpers::make_random<uint32_t>(700, 0, 1u << 30), vb = helpers::make_random<uint32_t>(500, 0, 1u << 30);
  CHECK((sort_join<uint32_t, uint32_t, uint32_t>(ka, va, kb, vb) == seq_join<uint32_t, uint32_t, uint32_t>(ka, va, kb, vb)));
  std::ostringstream os;
  os << zip<int, int, int>({1, 2}, {3, 4}, {5, 6});
  CHECK(os.str() == "1 3 5\n2 4 6");
}

// ---- report formats (common/result.cpp) -----------------------------------------------------------------------------
static void report_format_tests() {
  HashJoinResult r;
  r.kernel_time = Duration(2500.0);
  r.host_time = Duration(1234.5);
  r.build_time = Duration(400.25);
  r.probe_time = Duration(834.25);
  std::ostringstream os;
  os << static_cast<const Result &>(r);
  CHECK(os.str() == "Kernel duration: 2.5 us\nHost duration:   1234.5 us\nBuild time: 400.25 us\nProbe time: 834.25 us\n");
  Result base;
  base.host_time = Duration(10.0);
  std::ostringstream os2;
  os2 << base;
  CHECK(os2.str() == "Kernel duration: 0 us\nHost duration:   10 us\n");

  const std::string path = "/tmp/dwarf_bench_b200_report_test.csv";
  std::remove(path.c_str());
  MeasureResults mr("T");
  auto res = std::make_unique<Result>();
  res->host_time = Duration(1999.9);     // truncated to whole microseconds, then / 1000
  res->kernel_time = Duration(250.0);
  mr.add_result({{"device_type", "GPU"}, {"buf_size", "1024"}}, std::move(res));
  mr.write_csv(path);
  mr.write_csv(path);                    // append: header once
  std::ifstream in(path);
  std::string l1, l2, l3, l4;
  std::getline(in, l1);
  std::getline(in, l2);
  std::getline(in, l3);
  CHECK(l1 == "device_type,buf_size_bytes,host_time_ms,kernel_time_ms");
  CHECK(l2 == "GPU,4096,1.999,0.25");
  CHECK(l3 == l2);
  CHECK(!std::getline(in, l4));
  bool threw = false;
  try {
    mr.write_csv("/nonexistent_dir/x.csv");
  } catch (const std::runtime_error &e) {
    threw = std::string(e.what()) == "Could not open the file at /nonexistent_dir/x.csv";
  }
  CHECK(threw);
  MeasureResults missing("M");
  missing.add_result({{"buf_size", "1"}}, std::make_unique<Result>());
  threw = false;
  try {
    missing.write_csv(path);             // params.at("device_type") must throw, as in the reference
  } catch (const std::out_of_range &) {
    threw = true;
  }
  CHECK(threw);
  std::remove(path.c_str());
}

// ---- options, registry, generators -----------------------------------------------------------------------------------
static void framework_tests() {
  for (auto [text, want] : std::vector<std::pair<const char *, RunOptions::DeviceType>>{
           {"cpu", RunOptions::CPU}, {"GPU", RunOptions::GPU}, {"iGpu", RunOptions::iGPU}, {"whatever", RunOptions::Default}}) {
    std::istringstream in(text);
    RunOptions::DeviceType dt;
    in >> dt;
    CHECK(dt == want);
  }
  CHECK(to_string(RunOptions::CPU) == "CPU" && to_string(RunOptions::GPU) == "GPU" && to_string(RunOptions::iGPU) == "iGPU");
  CHECK(to_string(RunOptions::Default) == "GPU");                          // options.cpp:26-28
  RunOptions o;
  CHECK(o.device_ty == RunOptions::Default && o.iterations == 1 && o.input_size.empty());

  populate_registry();
  populate_registry();                                                     // idempotent
  std::set<std::string> names;
  for (const auto &dw : *Registry::instance()) names.insert(dw.first);
  CHECK((names == std::set<std::string>{"CuckooHashBuild", "GroupBy", "GroupByCuda", "HashBuild", "HashBuildNonBitmask", "Join", "JoinOmnisci", "JoinOmnisciCuda",
                                        "SlabHashBuild", "SlabJoin", "SlabProbe"}));
  CHECK(Registry::instance()->find("Join") != nullptr && Registry::instance()->find("Join")->name() == "Join");
  CHECK(Registry::instance()->find("NoSuchDwarf") == nullptr);

  for (size_t n : {size_t(1), size_t(128), size_t(4096)}) {
    auto v = helpers::make_unique_random(n);
    CHECK(v.size() == n && std::is_sorted(v.begin(), v.end()) && std::adjacent_find(v.begin(), v.end()) == v.end());
    CHECK(v.back() < 10 * n);
  }
  auto r = helpers::make_random<uint32_t>(5000);
  CHECK(*std::min_element(r.begin(), r.end()) >= 1 && *std::max_element(r.begin(), r.end()) <= 10000);
  const uint32_t seed = helpers::make_random();
  CHECK(seed >= 1 && seed <= 1000);

  // Meter: stable params merged under per-run params; Dwarf: results, clear.
  Dwarf *join = Registry::instance()->find("Join");
  RunOptions cpu;
  cpu.device_ty = RunOptions::CPU;
  cpu.input_size = {128};
  join->init(cpu);
  bool threw = false;
  try {
    join->run(cpu);                                                        // no CPU path in this build
  } catch (const std::logic_error &) {
    threw = true;
  }
  CHECK(threw);
  join->meter().add_result({{"buf_size", "128"}}, std::make_unique<HashJoinResult>());
  const DwarfRunResult &first = *join->get_results().begin();
  CHECK(first.params.at("device_type") == "CPU" && first.params.at("buf_size") == "128");
  join->clear_results();
  CHECK(join->get_results().begin() == join->get_results().end());

  DwarfBench::DwarfBench db;
  threw = false;
  try {
    db.makeMeasurements({DwarfBench::DeviceType::CPU, 128, 1, DwarfBench::Dwarf::Join});
  } catch (const DwarfBench::DwarfBenchException &) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    db.makeMeasurements({DwarfBench::DeviceType::GPU, 128, 1, DwarfBench::Dwarf::Sort});
  } catch (const DwarfBench::DwarfBenchException &) {
    threw = true;
  }
  CHECK(threw);
}

// ---- tests/dwarf_tests/dwarf_tests.cpp ----------------------------------------------------------------------------------
template <class DwarfClass> static void test_dwarf(size_t size) {
  StdoutCapture c;
  RunOptions plain;
  plain.device_ty = RunOptions::GPU;
  plain.input_size = {size};
  plain.iterations = 10;                                                   // utils.cpp:19-27
  const GroupByRunOptions opts(plain, 64, 1024);                           // utils.cpp: get_gpu_test_opts_groupby (every dwarf accepts the subclass)
  std::unique_ptr<Dwarf> dwarf = std::make_unique<DwarfClass>();
  dwarf->init(opts);
  dwarf->run(opts);
  size_t n = 0;
  for (const DwarfRunResult &res : dwarf->get_results()) {
    CHECK(res.result->valid);
    CHECK(res.params.at("buf_size") == std::to_string(size) && res.params.at("device_type") == "GPU");
    CHECK(res.result->host_time.count() > 0);
    ++n;
  }
  CHECK(n == 10);
}

template <class DwarfClass> static void test_suite() {
  for (size_t size : {128, 256, 512, 1024, 2048, 4096}) test_dwarf<DwarfClass>(size);   // dwarf_tests.cpp:44-50
}

static void gpu_tests() {
  test_suite<GroupBy>();                  // dwarf_tests.cpp:68 (GENERATE_TEST_SUITE_GROUPBY)
  test_suite<GroupByCuda>();
  test_suite<HashBuild>();
  test_suite<HashBuildNonBitmask>();      // dwarf_tests.cpp:68, :77-78 (EXPERIMENTAL in the reference)
  test_suite<SlabHashBuild>();
  test_suite<CuckooHashBuild>();
  test_suite<Join>();
  test_suite<SlabJoin>();
  test_suite<SlabProbe>();
  test_suite<JoinOmnisci>();
  test_suite<JoinOmnisciCuda>();
  test_dwarf<Join>(1);                  // the CLI's default input size (main.cpp:81-83)
  test_dwarf<Join>(100000);
  {
    StdoutCapture c;
    DwarfBench::DwarfBench db;
    auto ms = db.makeMeasurements({DwarfBench::DeviceType::GPU, 2048, 5, DwarfBench::Dwarf::Join});   // asserts in the reference
    CHECK(ms.size() == 5);
    for (const auto &m : ms) CHECK(m.dataSize == 2048 && m.microseconds > 0);
  }
}

int main(int argc, char **argv) {
  const std::string what = argc > 1 ? argv[1] : "cpu";
  join_helpers_tests();
  report_format_tests();
  framework_tests();
  if (what == "gpu") gpu_tests();
  std::printf("%s: %d checks, %d failed\n", what.c_str(), g_checks, g_failed);
  return g_failed ? 1 : 0;
}
