#include "dwarf_framework.hpp"

#include <algorithm>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

// ---- Result printing (byte-compatible with common/result.cpp:9-27) ----------------------------------------
std::ostream &operator<<(std::ostream &os, const Result &res) { return res.print_to_stream(os); }

std::ostream &Result::print_to_stream(std::ostream &os) const {
  // The reference divides kernel_time by 1000 and still labels it "us" (result.cpp:10); kept as is.
  os << "Kernel duration: " << kernel_time.count() / 1000.0 << " us\n";
  os << "Host duration:   " << host_time.count() << " us\n";
  return os;
}

std::vector<Duration> Result::get_reported_timings_list() const { return {host_time, kernel_time}; }

std::ostream &HashJoinResult::print_to_stream(std::ostream &os) const {
  Result::print_to_stream(os);
  os << "Build time: " << build_time.count() << " us\n";
  os << "Probe time: " << probe_time.count() << " us\n";
  return os;
}

// host_time, kernel_time first (the reference's two columns, result.cpp:16-18), then the engine's extras.
std::vector<Duration> HashJoinResult::get_reported_timings_list() const {
  return {host_time, kernel_time, build_time, probe_time};
}

// ---- MeasureResults ---------------------------------------------------------------------------------------
void MeasureResults::add_result(DwarfParams params, std::unique_ptr<Result> result) {
  results_.push_back(DwarfRunResult{std::move(params), std::move(result)});
}

namespace {
double whole_us_as_ms(const Duration &d) {
  return std::chrono::duration_cast<std::chrono::microseconds>(d).count() / 1000.0;
}
}  // namespace

void MeasureResults::write_csv(const std::string &filename) const {
  const bool had_file = std::ifstream(filename).good();
  std::ofstream out(filename, std::ios::app);       // append: sweeps re-run incrementally
  if (!out.is_open()) throw std::runtime_error("Could not open the file at " + filename);
  if (!had_file) out << "device_type,buf_size_bytes," << header_ << "\n";
  for (const DwarfRunResult &run : results_) {
    const size_t buf_size_bytes = std::stoll(run.params.at("buf_size")) * sizeof(int);
    out << run.params.at("device_type") << "," << buf_size_bytes << ",";
    std::string row;
    for (const Duration &d : run.result->get_reported_timings_list()) {
      std::ostringstream cell;
      cell << whole_us_as_ms(d);
      if (!row.empty()) row += ",";
      row += cell.str();
    }
    out << row << "\n";
  }
}

// ---- options ----------------------------------------------------------------------------------------------
std::istream &operator>>(std::istream &in, RunOptions::DeviceType &dt) {
  std::string word;
  in >> word;
  for (char &c : word) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  dt = word == "cpu"    ? RunOptions::DeviceType::CPU
       : word == "gpu"  ? RunOptions::DeviceType::GPU
       : word == "igpu" ? RunOptions::DeviceType::iGPU
                        : RunOptions::DeviceType::Default;
  return in;
}

std::string to_string(const RunOptions::DeviceType &dt) {
  switch (dt) {
  case RunOptions::DeviceType::CPU: return "CPU";
  case RunOptions::DeviceType::iGPU: return "iGPU";
  case RunOptions::DeviceType::GPU:
  case RunOptions::DeviceType::Default: return "GPU";   // options.cpp:26-28: Default reports as GPU
  }
  throw std::logic_error("Unsupported device type!");
}

// ---- meter / dwarf / registry --------------------------------------------------------------------------------
void Meter::add_result(DwarfParams &&params, std::unique_ptr<Result> result) {
  DwarfParams merged = params_;                      // stable params first, per-run params on top (meter.cpp:3-11)
  merged.insert(params.begin(), params.end());
  result_.add_result(std::move(merged), std::move(result));
}

void Dwarf::report(const RunOptions &opts) {
  if (opts.report_path.empty()) {
    for (const DwarfRunResult &run : results_) std::cout << *run.result;
    return;
  }
  results_.set_report_header(reporting_header_);
  results_.write_csv(opts.report_path);
}

Registry *Registry::instance() {
  static std::unique_ptr<Registry> the_registry(new Registry());
  return the_registry.get();
}

void Registry::registerd(Dwarf *dw) {
  std::unique_ptr<Dwarf> owned(dw);
  dwarfs_.emplace(owned->name(), std::move(owned));
}

Dwarf *Registry::find(const std::string &name) const {
  const auto it = dwarfs_.find(name);
  return it == dwarfs_.end() ? nullptr : it->second.get();
}
