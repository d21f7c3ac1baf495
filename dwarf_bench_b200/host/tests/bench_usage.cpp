// Library-API smoke program (the role of the reference's example/bench_usage/main.cpp): drive
// DwarfBench::makeMeasurements the way HDK does and print one line per measurement.
#include <bench.hpp>
#include <iostream>

int main() {
  DwarfBench::DwarfBench db;
  int failures = 0;
  for (DwarfBench::DeviceType device : {DwarfBench::DeviceType::GPU, DwarfBench::DeviceType::CPU}) {
    for (DwarfBench::Dwarf dwarf : {DwarfBench::Dwarf::Join, DwarfBench::Dwarf::Sort}) {
      DwarfBench::RunConfig rc{device, 1024, 10, dwarf};
      try {
        for (const DwarfBench::Measurement &m : db.makeMeasurements(rc))
          std::cout << dwarf << ' ' << device << " RESULT: " << m.dataSize << ' ' << m.microseconds << std::endl;
        if (!(dwarf == DwarfBench::Dwarf::Join && device == DwarfBench::DeviceType::GPU)) ++failures;   // must have thrown
      } catch (const DwarfBench::DwarfBenchException &e) {
        std::cout << dwarf << ' ' << device << " not served: " << e.what() << std::endl;
        if (dwarf == DwarfBench::Dwarf::Join && device == DwarfBench::DeviceType::GPU) ++failures;
      }
    }
  }
  return failures;
}
