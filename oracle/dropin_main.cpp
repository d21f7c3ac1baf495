// dropin_main.cpp -- proof that the B200 Join dwarfs are a source-level drop-in for the reference.
//
// TEST INFRASTRUCTURE ONLY (built by `make -C oracle dropin` into oracle/_ref/, only where the reference checkout
// exists).  dwarf_bench_b200/host/join/b200_dwarfs.cpp is compiled HERE against the REFERENCE'S OWN headers
// (common/common.hpp, common/dwarf.hpp, common/meter.hpp, common/result.hpp, common/registry.hpp,
// join/join_helpers/join_helpers.hpp -- all found through -I$(REF), nothing copied) and linked with the
// reference's own common/{result,meter,options,registry}.cpp.  Only helpers:: (common/common.cpp needs
// boost::dll, absent from this image) is provided below.  The program is the reference's test_dwarf<>()
// (tests/dwarf_tests/dwarf_tests.cpp:12-23) over the replaced dwarfs plus a CSV report through the reference's
// own MeasureResults::write_csv.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <set>

#include "common/common.hpp"
#include "common/registry.hpp"
#include "join/b200_dwarfs.hpp"

namespace helpers {
std::vector<uint32_t> make_unique_random(size_t size) {   // common/common.cpp:7-20 semantics
  std::random_device rd;
  std::mt19937 gen(rd());
  std::uniform_int_distribution<int> dist(1, std::min((long)(size * 10), (long)((uint32_t)-1)));
  std::set<uint32_t> s;
  while (s.size() < size) s.insert(dist(gen) % (size * 10));
  return std::vector<uint32_t>(s.begin(), s.end());
}
uint32_t make_random() {
  std::random_device rd;
  std::mt19937 gen(rd());
  return std::uniform_int_distribution<int>(1, 1000)(gen);
}
}  // namespace helpers

template <class DwarfClass> static int test_dwarf(size_t size) {
  RunOptions opts;
  opts.device_ty = RunOptions::DeviceType::GPU;
  opts.input_size = {size};
  opts.iterations = 10;
  std::unique_ptr<Dwarf> dwarf = std::make_unique<DwarfClass>();
  dwarf->init(opts);
  dwarf->run(opts);
  int bad = 0, n = 0;
  for (const DwarfRunResult &res : dwarf->get_results()) {
    bad += !res.result->valid;
    ++n;
  }
  return bad + (n != 10);
}

int main() {
  int bad = 0;
  for (size_t size : {128, 256, 512, 1024, 2048, 4096}) {
    bad += test_dwarf<Join>(size);
    bad += test_dwarf<HashBuild>(size);
    bad += test_dwarf<JoinOmnisci>(size);
    bad += test_dwarf<SlabProbe>(size);
  }
  // The reference's registry and CSV writer over the replaced dwarf.
  Registry::instance()->registerd(new Join());
  Dwarf *join = Registry::instance()->find("Join");
  RunOptions opts;
  opts.device_ty = RunOptions::DeviceType::GPU;
  opts.input_size = {1024};
  opts.iterations = 3;
  opts.report_path = "/tmp/dropin_report.csv";
  std::remove(opts.report_path.c_str());
  join->init(opts);
  join->run(opts);
  join->report(opts);
  std::ifstream in(opts.report_path);
  std::string header, row;
  std::getline(in, header);
  std::getline(in, row);
  bad += header.rfind("device_type,buf_size_bytes,host_time_ms,kernel_time_ms", 0) != 0;
  bad += row.rfind("GPU,4096,", 0) != 0;
  std::printf("dropin: %d failures\n", bad);
  return bad != 0;
}
