// b200_dwarfs.hpp -- the dwarfs of the Join hot path, each backed by libdwj_b200.so (include/dwj.h).
//
// Registry names and what they time are the reference's:
//   Join            join/join.cpp            build + probe, probe-aligned outputs, host compaction, == seq_join
//   HashBuild       hash/hash_build.cpp      build only (val = key, keys in [1,10000]); has() == 1 for every key
//   HashBuildNonBitmask, SlabHashBuild, CuckooHashBuild   hash/*.cpp   the same at their own fill (0.9 / 0.625 / 0.25)
//   SlabProbe       probe/slab_probe.cpp     probe only (build untimed); find() == 1 for every key
//   SlabJoin        join/slab_join.cpp       Join's flow (the reference's slab variant); here the same table
//   JoinOmnisci     join/join_omnisci.cpp    one-to-many join on row ids, keys in [1,10000]
//   JoinOmnisciCuda join/join_omnisci.hpp    the name bench.cpp:45-47 asks for on GPU
//   GroupBy, GroupByCuda  groupby/groupby.cpp  hash aggregation SUM(value) BY key (dwj_aggregate_sum), keys in [0, groups_count)
// device=cpu is refused (std::logic_error): the engine has no CPU path by design.
#pragma once
#include "common/common.hpp"

struct dwj_engine;

namespace b200 {
// RAII handle over a dwj_engine; throws std::runtime_error carrying dwj_last_error() on any failure.
class Engine {
public:
  Engine(size_t max_build_rows, unsigned flags, int key_bytes = 4, double load_factor = 0.0);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;
  dwj_engine *get() const { return e_; }
  static void check(int rc);
  static std::string device_name(int device = 0);

private:
  dwj_engine *e_ = nullptr;
};
void require_gpu(const RunOptions &opts, const std::string &dwarf);
}  // namespace b200

#define B200_DECLARE_DWARF(NAME)                         \
  class NAME : public Dwarf {                            \
  public:                                                \
    NAME();                                              \
    void run(const RunOptions &opts) override;           \
    void init(const RunOptions &opts) override;          \
                                                         \
  private:                                               \
    void _run(const size_t buffer_size, Meter &meter);   \
  };

B200_DECLARE_DWARF(Join)
B200_DECLARE_DWARF(SlabJoin)
B200_DECLARE_DWARF(HashBuild)
B200_DECLARE_DWARF(HashBuildNonBitmask)
B200_DECLARE_DWARF(SlabHashBuild)
B200_DECLARE_DWARF(CuckooHashBuild)
B200_DECLARE_DWARF(SlabProbe)
B200_DECLARE_DWARF(GroupBy)
B200_DECLARE_DWARF(GroupByCuda)
B200_DECLARE_DWARF(JoinOmnisci)
B200_DECLARE_DWARF(JoinOmnisciCuda)
