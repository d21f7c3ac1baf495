// bench.hpp -- library facade (reference: bench.hpp / bench.cpp, installed there as libdbench + CMake package
// dbench::dbench; consumed by HDK, docs/hdk.md:30-35).  Same namespace, enums, structs and entry point.
//
// Differences, all on the Join path: Dwarf::Join with DeviceType::GPU runs (the reference asserts,
// bench.cpp:40-44) and is served by the B200 engine; DeviceType::CPU throws DwarfBenchException (this build has
// no CPU path).  Join and GroupBy are served; Scan / Sort are outside this build's scope and throw DwarfBenchException.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

namespace DwarfBench {

enum Dwarf {
  Scan,
  Join,
  GroupBy,
  Sort,
};

enum DeviceType { CPU, GPU };

// dataSize: the column size in ELEMENTS (the reference's doc comment says bytes, its code stores elements:
// bench.cpp:96-97); microseconds: host wall-clock time of one iteration.
struct Measurement {
  size_t dataSize;
  size_t microseconds;
};

struct RunConfig {
  DeviceType device;
  size_t inputSize;
  size_t iterations;
  Dwarf dwarf;
};

class DwarfBenchException : public std::exception {
private:
  std::string message_;

public:
  explicit DwarfBenchException(const std::string &message);
  const char *what() const noexcept override;
};

class DwarfBench {
public:
  DwarfBench() = default;
  // One Measurement per iteration.  Not re-entrant (shared registry, as in the reference).
  std::vector<Measurement> makeMeasurements(const RunConfig &conf);

private:
  enum DwarfImpl { GroupBy, HashBuild, Join, JoinOmnisci, SlabProbe, Unsupported };
  DwarfImpl dwarfToImpl(Dwarf dwarf);
  std::string dwarfToString(DwarfImpl dwarf, DeviceType device);
};

}  // namespace DwarfBench
