#include "register_dwarfs.hpp"

#include "common/registry.hpp"
#include "join/b200_dwarfs.hpp"

// Reference: register_dwarfs.cpp:20-56.  Registry::registerd drops a second dwarf of the same name, so calling
// this twice is harmless (the CLI and makeMeasurements may both call it in one process).
void populate_registry() {
  auto registry = Registry::instance();
  registry->registerd(new GroupBy());
  registry->registerd(new GroupByCuda());
  registry->registerd(new HashBuild());
  registry->registerd(new HashBuildNonBitmask());
  registry->registerd(new SlabHashBuild());
  registry->registerd(new CuckooHashBuild());
  registry->registerd(new Join());
  registry->registerd(new JoinOmnisci());
  registry->registerd(new JoinOmnisciCuda());
  registry->registerd(new SlabJoin());
  registry->registerd(new SlabProbe());
}
