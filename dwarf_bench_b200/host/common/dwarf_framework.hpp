// dwarf_framework.hpp -- host-side plugin surface of the B200 Join engine.
//
// Mirrors, name for name and format for format, the part of dwarf_bench's framework that the Join hot
// path touches, so that a dwarf written against the reference compiles against this tree unchanged:
//   DwarfParams, Duration, Result, HashJoinResult, DwarfRunResult, MeasureResults  common/result.{hpp,cpp}
//   RunOptions, GroupByRunOptions, operator>>, to_string                          common/options.{hpp,cpp}
//   Meter                                                                          common/meter.{hpp,cpp}
//   Dwarf                                                                          common/dwarf.hpp
//   Registry                                                                       common/registry.{hpp,cpp}
// The reference spreads these over five headers; here they live in one file and the reference's header
// names (common/dwarf.hpp, common/meter.hpp, ...) are thin forwarders to it.  Behavioural contract kept:
// stdout block "Kernel duration / Host duration / Build time / Probe time" (result.cpp:9-27), CSV append
// with header "device_type,buf_size_bytes,host_time_ms,kernel_time_ms" and buf_size_bytes = buf_size *
// sizeof(int), times as whole microseconds / 1000 (result.cpp:59-91), params "device_type" / "buf_size".
#pragma once
#define DWJ_HOST_FRAMEWORK 1

#include <chrono>
#include <iosfwd>
#include <map>
#include <memory>
#include <string>
#include <vector>

using DwarfParams = std::map<std::string, std::string>;
using Duration = std::chrono::duration<double, std::micro>;

// ---- results ----------------------------------------------------------------------------------------
struct Result {
  size_t thread_x = 1, thread_y = 1, tread_z = 1;   // (sic) the reference's field name
  size_t group_size = 1;
  size_t bytes = 0;
  size_t iterations = 0;
  size_t bytes_per_iteration = 0;
  Duration kernel_time{0};
  Duration host_time{0};
  bool valid = true;

  virtual ~Result() = default;
  virtual std::vector<Duration> get_reported_timings_list() const;

protected:
  virtual std::ostream &print_to_stream(std::ostream &os) const;
  friend std::ostream &operator<<(std::ostream &out, const Result &instance);
};

struct HashJoinResult : public Result {
  Duration probe_time{0};
  Duration build_time{0};
  // Engine extras (not in the reference): filled by the B200 dwarfs, reported as extra CSV columns.
  size_t matches = 0;
  double tuples_per_second = 0;
  std::vector<Duration> get_reported_timings_list() const override;
  std::ostream &print_to_stream(std::ostream &os) const override;
};

std::ostream &operator<<(std::ostream &os, const Result &res);

struct DwarfRunResult {
  DwarfParams params;
  std::unique_ptr<Result> result;
};

static constexpr auto default_report_header = "host_time_ms,kernel_time_ms";
using SingleRunResults = std::vector<DwarfRunResult>;

class MeasureResults {
public:
  using const_iterator = SingleRunResults::const_iterator;
  explicit MeasureResults(const std::string &name) : name_(name), header_(default_report_header) {}

  void add_result(DwarfParams params, std::unique_ptr<Result> result);
  const_iterator begin() const { return results_.begin(); }
  const_iterator end() const { return results_.end(); }
  void set_report_header(const std::string &header) { header_ = header; }
  void write_csv(const std::string &filename) const;
  void clear() { results_.clear(); }

private:
  SingleRunResults results_;
  const std::string name_;
  std::string header_;
};

// ---- options ----------------------------------------------------------------------------------------
struct RunOptions {
  enum DeviceType { CPU, GPU, iGPU, Default };
  DeviceType device_ty = DeviceType::Default;
  std::vector<size_t> input_size;
  size_t iterations = 1;
  std::string root_path;
  std::string report_path;
  size_t gpus = 1;   // not in the reference (one queue on one device, join/join.cpp:23-24): `--gpus N` runs the Join
                     // dwarfs over N GPUs of the box (dwj_mg_*); 1 = the single-GPU engine
};

struct GroupByRunOptions : public RunOptions {
  GroupByRunOptions(const RunOptions &opts, size_t groups_count, size_t executors)
      : RunOptions(opts), groups_count(groups_count), executors(executors) {}
  size_t groups_count;
  size_t executors;
};

std::istream &operator>>(std::istream &in, RunOptions::DeviceType &dt);
std::string to_string(const RunOptions::DeviceType &dt);

// ---- meter ------------------------------------------------------------------------------------------
class Meter {
public:
  Meter(const std::string &dwarf_name, MeasureResults &result) : dwarf_name_(dwarf_name), result_(result) {}
  void add_result(DwarfParams &&params, std::unique_ptr<Result> result);
  void set_params(DwarfParams params) { params_ = std::move(params); }
  void set_opts(const RunOptions &opts) { opts_ = &opts; }
  const RunOptions &opts() const { return *opts_; }

private:
  const std::string dwarf_name_;
  MeasureResults &result_;
  DwarfParams params_;
  RunOptions const *opts_ = nullptr;
};

// ---- the plugin base class ----------------------------------------------------------------------------
class Dwarf {
public:
  explicit Dwarf(const std::string &name)
      : reporting_header_(default_report_header), name_(name), results_(name), meter_(name, results_) {}
  virtual ~Dwarf() = default;

  const std::string &name() const { return name_; }

  virtual void run(const RunOptions &opts) = 0;
  virtual void init(const RunOptions &opts) = 0;
  void report(const RunOptions &opts);

  Meter &meter() { return meter_; }
  const MeasureResults &get_results() const { return results_; }
  void clear_results() { results_.clear(); }

protected:
  std::string reporting_header_;

private:
  std::string name_;
  MeasureResults results_;
  Meter meter_;
};

// ---- registry ---------------------------------------------------------------------------------------
class Registry {
public:
  using const_iterator = std::map<std::string, std::unique_ptr<Dwarf>>::const_iterator;
  static Registry *instance();
  void registerd(Dwarf *dw);   // takes ownership; a second dwarf of the same name is dropped (map::emplace)

  Dwarf *find(const std::string &name) const;
  void set_root(const std::string &root) { root_path_ = root; }

  const_iterator begin() const { return dwarfs_.begin(); }
  const_iterator end() const { return dwarfs_.end(); }

private:
  Registry() = default;
  std::map<std::string, std::unique_ptr<Dwarf>> dwarfs_;
  std::string root_path_;
};
