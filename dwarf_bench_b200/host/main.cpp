// dwarf_bench -- command line front end (reference: main.cpp, there on boost::program_options).
//
//   dwarf_bench <Dwarf|list> [--input_size a b ...] [--iterations N] [--device cpu|gpu|igpu]
//               [--report_path P] [--groups_count N] [--executors N] [--gpus N] [--help]
// Same spellings: positional dwarf name, multitoken --input_size, "--opt value" and "--opt=value",
// case-insensitive device.  Deliberate difference: a caught exception makes the exit code 2 (the reference
// prints "Caught exception" and still returns 0, main.cpp:97-100).
#include <iostream>
#include <memory>
#include <sstream>

#include "common/common.hpp"
#include "common/registry.hpp"
#include "register_dwarfs.hpp"

namespace {

bool isGroupBy(const std::string &dwarfName) { return dwarfName.find("GroupBy") != std::string::npos; }

void print_help() {
  std::cout << "Dwarf bench:\n"
               "  --help                Show help message\n"
               "  --dwarf arg           Dwarf to run. List all with 'list' option.\n"
               "  --input_size arg      Data array size, ususally a column size in elements\n"
               "  --iterations arg      Number of iterations to run a bmark.\n"
               "  --device arg          Device to run on.\n"
               "  --report_path arg     Full/Relative path to a report file.\n"
               "  --groups_count arg    Number of unique keys for dwarfs with keys (groupby, hash build etc.).\n"
               "  --executors arg       Number of executors for GroupByLocal.\n"
               "  --gpus arg            GPUs of this box the Join dwarfs run on: 1, 2, 4 or 8 (not in the reference).\n";
}

size_t to_size(const std::string &opt, const std::string &text) {
  size_t pos = 0;
  unsigned long long v = 0;
  try {
    v = std::stoull(text, &pos);
  } catch (const std::exception &) {
    pos = 0;
  }
  if (pos != text.size() || text.empty()) throw std::invalid_argument("the argument ('" + text + "') for option '--" + opt + "' is invalid");
  return static_cast<size_t>(v);
}

}  // namespace

int main(int argc, char *argv[]) {
  populate_registry();
  auto registry = Registry::instance();

  std::unique_ptr<RunOptions> opts = std::make_unique<RunOptions>();
  size_t groups_count = 1, executors = 1;
  opts->root_path = helpers::get_kernels_root_env(argv[0]);
  std::cout << "DWARF_BENCH_ROOT is set to " << opts->root_path << std::endl
            << "You can change that with 'export DWARF_BENCH_ROOT=/your/path'\n";

  std::string dwarf_name;
  bool want_help = false;
  try {
    for (int i = 1; i < argc; ++i) {
      std::string arg = argv[i];
      if (arg.rfind("--", 0) != 0) {                 // positional: the dwarf name (first one wins)
        if (!dwarf_name.empty()) throw std::invalid_argument("too many positional options have been specified on the command line");
        dwarf_name = arg;
        continue;
      }
      std::string name = arg.substr(2), value;
      bool has_value = false;
      const size_t eq = name.find('=');
      if (eq != std::string::npos) {
        value = name.substr(eq + 1);
        name = name.substr(0, eq);
        has_value = true;
      }
      auto next_value = [&]() -> std::string {
        if (has_value) return value;
        if (i + 1 >= argc) throw std::invalid_argument("the required argument for option '--" + name + "' is missing");
        return argv[++i];
      };
      if (name == "help") {
        want_help = true;
      } else if (name == "dwarf") {
        dwarf_name = next_value();
      } else if (name == "input_size") {             // multitoken: consume every following non-option token
        opts->input_size.push_back(to_size(name, next_value()));
        while (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) opts->input_size.push_back(to_size(name, argv[++i]));
      } else if (name == "iterations") {
        opts->iterations = to_size(name, next_value());
      } else if (name == "device") {
        std::istringstream in(next_value());
        in >> opts->device_ty;
      } else if (name == "report_path") {
        opts->report_path = next_value();
      } else if (name == "gpus") {
        opts->gpus = to_size(name, next_value());
      } else if (name == "groups_count") {
        groups_count = to_size(name, next_value());
      } else if (name == "executors") {
        executors = to_size(name, next_value());
      } else {
        throw std::invalid_argument("unrecognised option '--" + name + "'");
      }
    }

    if (dwarf_name == "list") {
      std::cout << "Supported dwarfs:\n";
      for (const auto &dw : *registry) std::cout << "\t" << dw.first << std::endl;
      return 0;
    }
    Dwarf *dwarf = registry->find(dwarf_name);
    if (want_help) {
      print_help();
      return 0;
    }
    if (!dwarf) {
      std::cerr << "List supported dwarfs to run with '" << argv[0] << " list'" << std::endl;
      return 1;
    }
    if (opts->input_size.empty()) opts->input_size.push_back(1);       // main.cpp:81-83
    helpers::set_dpcpp_filter_env(*opts);
    if (isGroupBy(dwarf_name)) opts = std::make_unique<GroupByRunOptions>(*opts, groups_count, executors);

    dwarf->init(*opts);
    dwarf->run(*opts);
    dwarf->report(*opts);
    for (const DwarfRunResult &res : dwarf->get_results())
      if (!res.result->valid) return 3;              // a wrong answer is not a success
  } catch (std::exception &e) {
    std::cerr << "Caught exception: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
