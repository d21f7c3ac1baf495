#include "bench.hpp"

#include "common/common.hpp"
#include "register_dwarfs.hpp"

namespace DwarfBench {

DwarfBenchException::DwarfBenchException(const std::string &message) : message_(message) {}
const char *DwarfBenchException::what() const noexcept { return message_.c_str(); }

DwarfBench::DwarfImpl DwarfBench::dwarfToImpl(Dwarf dwarf) {
  if (dwarf == Dwarf::Join) return DwarfImpl::JoinOmnisci;                          // bench.cpp:112-113
  if (dwarf == Dwarf::GroupBy) return DwarfImpl::GroupBy;                           // bench.cpp:115-116
  return DwarfImpl::Unsupported;
}

std::string DwarfBench::dwarfToString(DwarfImpl dwarf, DeviceType device) {
  switch (dwarf) {
  case JoinOmnisci: return device == DeviceType::CPU ? "JoinOmnisci" : "JoinOmnisciCuda";   // bench.cpp:45-47
  case GroupBy: return device == DeviceType::CPU ? "GroupBy" : "GroupByCuda";               // bench.cpp:22-23
  case Join: return "Join";
  case HashBuild: return "HashBuild";
  case SlabProbe: return "SlabProbe";
  default: return "Unknown Dwarf";
  }
}

std::vector<Measurement> DwarfBench::makeMeasurements(const RunConfig &conf) {
  static Registry *reg = []() {
    populate_registry();
    return Registry::instance();
  }();

  if (conf.device == DeviceType::CPU)
    throw DwarfBenchException("this build serves DeviceType::GPU only (B200 engine, no CPU fallback)");
  RunOptions base;
  base.device_ty = RunOptions::DeviceType::GPU;
  base.input_size = {conf.inputSize};
  base.iterations = conf.iterations;
  base.report_path = "";
  GroupByRunOptions opts(base, 20, 1024);            // the reference always passes this subclass (bench.cpp:80)

  const std::string dwarfName = dwarfToString(dwarfToImpl(conf.dwarf), conf.device);
  ::Dwarf *dwarf = reg->find(dwarfName);
  if (dwarf == nullptr) throw DwarfBenchException("dwarf '" + dwarfName + "' is not part of this build (Join hot path only)");

  dwarf->clear_results();
  dwarf->init(opts);
  dwarf->run(opts);

  std::vector<Measurement> ms;
  for (const DwarfRunResult &res : dwarf->get_results())
    ms.push_back(Measurement{static_cast<size_t>(std::stoull(res.params.at("buf_size"))),
                             static_cast<size_t>(res.result->host_time.count())});
  return ms;
}

}  // namespace DwarfBench
