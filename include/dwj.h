/*
 * dwj.h -- C ABI of the B200-native hash-join engine (libdwj_b200.so).
 *
 * This is the drop-in boundary for dwarf_bench's Join hot path.  Plain C, POD
 * structs, raw device/host pointers and sizes only: no C++ types, no torch
 * types, no exceptions across the boundary.  Every call returns 0 on success
 * or a negative DWJ_ERR_* code; dwj_last_error() holds the message for the
 * calling thread.  There is NO CPU fallback: when no CUDA device is usable the
 * calls fail with DWJ_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to kurapov-peter/dwarf_bench):
 *   dwj_build            kernel `join_build`  join/join.cpp:60-77
 *                        = SimpleNonOwningHashTable::insert
 *                          common/dpcpp/hashtable.hpp:15-21,70-92;
 *                        kernel `hash_build`  hash/hash_build.cpp:36-50
 *   dwj_probe_aligned    kernel `join_probe`  join/join.cpp:80-104
 *                        = SimpleNonOwningHashTable::at  hashtable.hpp:23-40,
 *                          probe-aligned 0xFFFFFFFF-filled outputs join.cpp:41-43
 *   dwj_probe_contains   kernel `hash_build_check` hash/hash_build.cpp:61-76
 *                        = SimpleNonOwningHashTable::has hashtable.hpp:42-58;
 *                        SlabHashTable::find as used by probe/slab_probe.cpp:69-86
 *   dwj_probe_pairs      host compaction loop join/join.cpp:119-129, moved on
 *                        device; row definition = join_helpers::seq_join
 *                        join/join_helpers/join_helpers.hpp:85-104
 *   dwj_probe_count      size of that row list (join_helpers.hpp:21-25 get_size)
 *   dwj_join_host        the whole timed region join/join.cpp:45-113 with the
 *                        implicit sycl::buffer H2D/D2H copies made explicit
 *   dwj_aggregate_sum    kernel `hash_build` of the GroupBy dwarf groupby/groupby.cpp:60-72
 *                        = NonOwningHashTableNonBitmask::add hashtable.hpp:136-153
 *   dwj_timings          HashJoinResult{build_time,probe_time,host_time,
 *                        kernel_time}  common/result.hpp:11-33
 *   dwj_partition*, dwj_xpart_*, dwj_*_grouped, dwj_*_segments, dwj_xj_*, dwj_mg_*
 *                        new (no reference counterpart: the reference runs one
 *                        queue on one device, join/join.cpp:23-24): radix partition
 *                        on the key hash and the multi-GPU exchange join
 */
#ifndef DWJ_H
#define DWJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWJ_ABI_VERSION 2

#if defined(__GNUC__)
#define DWJ_API __attribute__((visibility("default")))
#else
#define DWJ_API
#endif

/* error codes */
#define DWJ_OK 0
#define DWJ_ERR_INVALID (-1)   /* bad argument                                        */
#define DWJ_ERR_CUDA (-2)      /* CUDA runtime error / no device                       */
#define DWJ_ERR_OOM (-3)       /* device or host allocation failed                     */
#define DWJ_ERR_OVERFLOW (-4)  /* output capacity too small; *n_matches holds the need */
#define DWJ_ERR_STATE (-5)     /* call order (probe before build, ...)                 */
#define DWJ_ERR_CAPACITY (-6)  /* more build rows than the table was created for       */

/* dwj_config.flags */
#define DWJ_FLAG_UNIQUE_BUILD_KEYS 0x1u /* build keys are distinct: a probe row stops at its
                                           first hit (SimpleNonOwningHashTable::at semantics).
                                           Without it every equal build row is matched
                                           (seq_join semantics) and the engine builds a
                                           ONE-TO-MANY table: distinct keys in the table, the
                                           payloads of each key in one contiguous run -- the
                                           layout of OmniSci::HashTable
                                           (common/dpcpp/omnisci_hashtable.hpp:80-192).  Such
                                           an engine holds at most max_build_rows rows.     */
#define DWJ_FLAG_L2_PERSIST 0x2u        /* pin the table in L2 with an access-policy window
                                           when it fits the device's persisting-L2 limit    */
#define DWJ_FLAG_NO_PARTITION 0x8u      /* never radix-partition inputs by table region (keeps
                                           dwj_probe_pairs in probe-row order for tables
                                           larger than L2, at the price of ~100 B of DRAM line
                                           fills per probe row)                                */
#define DWJ_FLAG_UNORDERED_OUTPUT 0x4u  /* dwj_probe_pairs may emit rows in any order (same
                                           multiset): output ranges are handed out with one
                                           atomicAdd per chunk instead of the order-preserving
                                           decoupled look-back scan                           */

typedef struct dwj_engine dwj_engine; /* opaque; owns the table and all scratch device memory */

typedef struct {
  int32_t device;          /* CUDA device ordinal                                         */
  int32_t key_bytes;       /* 4 (uint32 keys, the reference's type) or 8 (uint64)         */
  int32_t payload_bytes;   /* must equal key_bytes                                        */
  uint32_t flags;          /* DWJ_FLAG_*                                                  */
  uint64_t max_build_rows; /* table capacity in rows                                      */
  double load_factor;      /* (0, 0.9]; slots = next_pow2(max_build_rows / load_factor);
                              0 selects the reference's 0.5 (join/join.cpp:30: T = 2n)    */
  uint64_t hash_seed;      /* mixes into the slot hash (reference: MurmurHash3 seed,
                              join/join.cpp:32); does not affect results                  */
} dwj_config;

typedef struct {
  float build_ms;     /* device time of the last dwj_build (table clear + insert)          */
  float probe_ms;     /* device time of the last dwj_probe_*                               */
  float partition_ms; /* device time of the last dwj_partition / dwj_xpart_scatter          */
  float h2d_ms;       /* dwj_join_host only: host->device copies                           */
  float d2h_ms;       /* dwj_join_host only: device->host copies                           */
  float total_ms;     /* dwj_join_host: first copy to last copy; else build_ms + probe_ms  */
  float probe_kernel_ms; /* the probe kernel proper of the last dwj_probe_* (probe_ms also
                            covers the region partition of the probe relation and memsets)  */
  float build_kernel_ms; /* the insert kernel proper of the last dwj_build                  */
} dwj_timing;

typedef struct {
  uint64_t slots;            /* table slots (power of two)                                 */
  uint64_t table_bytes;      /* bytes of the slot array                                    */
  uint64_t build_rows;       /* rows inserted by the last dwj_build                        */
  uint32_t slot_bytes;       /* key_bytes + payload_bytes                                  */
  uint32_t slots_per_bucket; /* slots in one 32-byte sector                                */
  uint32_t l2_persist;       /* 1 when an access-policy window is active for the table     */
  uint32_t sm_count;
  uint64_t l2_bytes;
  uint32_t launches_build;   /* kernels (incl. memsets) one dwj_build enqueues             */
  uint32_t launches_probe;   /* kernels (incl. memsets) the last dwj_probe_* enqueued      */
  uint32_t radix_parts;      /* > 1: build and probe rows are first radix-partitioned into
                                this many table regions (L2 locality); dwj_probe_pairs then
                                emits rows region by region instead of in probe-row order   */
  uint32_t probe_passes;     /* > 1: dwj_probe_pairs (unique build keys) sweeps the probe relation
                                this many times, one table slice per pass, instead of partitioning it */
  uint32_t flags;            /* dwj_config.flags the engine was created with                */
  int32_t device;            /* its CUDA device                                             */
  uint64_t hash_seed;
  uint64_t max_build_rows;   /* dwj_config.max_build_rows                                   */
  uint32_t hot_probe_keys;   /* 1: the last dwj_probe_pairs over unique build keys found its probe keys skewed (a
                                sample of 4096 keys put >= 0.6 % into one bucket) and let table sectors into L1  */
  uint32_t reserved;
} dwj_info;

DWJ_API int dwj_abi_version(void);
DWJ_API const char *dwj_last_error(void);

DWJ_API int dwj_create(const dwj_config *cfg, dwj_engine **out);
DWJ_API int dwj_destroy(dwj_engine *e);
DWJ_API int dwj_get_info(const dwj_engine *e, dwj_info *info);

/* Insert n_rows (key, payload) pairs from DEVICE columns into a freshly cleared table.
 * DWJ_FLAG_UNIQUE_BUILD_KEYS engines give every row its own slot (as hashtable.hpp:15-21 does);
 * the others count the rows per distinct key, lay the keys' payload runs out and fill them
 * (three kernels, no host round trip; omnisci_hashtable.hpp:80-192).  `stream` is a
 * cudaStream_t (NULL = the legacy default stream).  Asynchronous. */
DWJ_API int dwj_build(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *stream);

/* Reference-shaped probe: for probe row i writes (key, build payload, probe payload) at index i
 * when the key is present, all-ones sentinels otherwise (join/join.cpp:41-43,93-103).  First hit
 * wins, as SimpleNonOwningHashTable::at.  Outputs are n_rows long.  Asynchronous. */
DWJ_API int dwj_probe_aligned(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows,
                      void *d_out_key, void *d_out_build_val, void *d_out_probe_val, void *stream);

/* d_out_flags[i] (uint32) = 1 when probe key i is in the table, else 0.  Asynchronous. */
DWJ_API int dwj_probe_contains(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t *d_out_flags,
                       void *stream);

/* Compacted join output -- in probe-row order while the table is small enough to be probed directly
 * (dwj_info.radix_parts == 1), else region by region in no particular order (same multiset; DWJ_FLAG_NO_PARTITION
 * keeps probe-row order at any size).  Row r of the result is
 * (d_out_key[r], d_out_build_val[r], d_out_probe_val[r]); d_out_key may be NULL to skip the key
 * column.  At most `capacity` rows are written; the total match count is stored to *d_n_matches
 * (device, uint64) and, when n_matches != NULL, the call synchronises the stream and returns it
 * on the host too (DWJ_ERR_OVERFLOW when it exceeds capacity).  With n_matches == NULL the call
 * is asynchronous. */
DWJ_API int dwj_probe_pairs(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows,
                    void *d_out_key, void *d_out_build_val, void *d_out_probe_val,
                    uint64_t capacity, uint64_t *d_n_matches, uint64_t *n_matches, void *stream);

/* Match count only (no materialisation).  Same n_matches convention as dwj_probe_pairs. */
DWJ_API int dwj_probe_count(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint64_t *d_n_matches,
                    uint64_t *n_matches, void *stream);

/* Device time of the last build / probe / partition on this engine (cudaEvent pairs on the
 * caller's stream).  Synchronises on those events. */
DWJ_API int dwj_timings(dwj_engine *e, dwj_timing *t);

/* output shapes for dwj_join_host */
#define DWJ_OUT_ALIGNED 0 /* probe-aligned sentinel arrays, n_probe long (reference shape) */
#define DWJ_OUT_PAIRS 1   /* compacted rows, *n_out of them                                */
#define DWJ_OUT_COUNT 2   /* only *n_out                                                   */

/* The whole join with HOST columns: H2D of both relations, build, probe, D2H of the result --
 * what the reference's sycl::buffer scopes do implicitly inside its timed window
 * (join/join.cpp:45-117).  The probe relation is streamed through in chunks so copies overlap
 * the kernels.  Host pointers may be pageable; pinned memory makes the copies asynchronous.
 * out_* must hold n_probe rows (ALIGNED) or out_capacity rows (PAIRS); out_key may be NULL. */
DWJ_API int dwj_join_host(dwj_engine *e, const void *build_keys, const void *build_vals, uint64_t n_build,
                  const void *probe_keys, const void *probe_vals, uint64_t n_probe, int out_mode,
                  void *out_key, void *out_build_val, void *out_probe_val, uint64_t out_capacity,
                  uint64_t *n_out, dwj_timing *timing);

/* Radix partition of DEVICE columns on the key hash, for the multi-GPU exchange: rows whose
 * partition id (top hash bits, independent of the slot hash) is p end up contiguous in
 * d_out_keys/d_out_vals[d_offsets[p] .. d_offsets[p+1]).  n_parts must be a power of two <= 256.
 * d_offsets: n_parts + 1 uint64 (device).  d_out_vals/d_vals may both be NULL.  Asynchronous. */
DWJ_API int dwj_partition(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows,
                  uint32_t n_parts, void *d_out_keys, void *d_out_vals, uint64_t *d_offsets,
                  void *stream);

/* Exchange planning: d_counts[p] (uint64, device) = rows of partition p, same partition function as dwj_partition.
 * Asynchronous. */
DWJ_API int dwj_partition_hist(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_parts, uint64_t *d_counts,
                       void *stream);

/* ---- exchange partition folded with the receiver's region grouping (multi-GPU) ----------------------------------
 * One pass over a relation that serves BOTH the exchange and the receiver's L2-region grouping: partition id =
 * destination rank (independent hash, as dwj_partition) x table region of the destination's table (top bits of the
 * bucket index; every rank's engine must be created with the same table size and hash seed).  Partitions are ordered
 * rank-major, region-minor.  dwj_xpart_regions() says how many regions are folded in: the engine's region count when
 * n_ranks x regions <= 512, else 1 (plain rank partition; the receiver then groups by region itself with dwj_build /
 * dwj_probe_pairs).  Usage: dwj_xpart_hist -> exchange the counts, plan the layout -> dwj_xpart_scatter into a send buffer -> copy
 * every (rank, region) run into the destination's receive buffer laid out region-major (dwj_copy_many or any
 * transport) -> dwj_build_grouped / dwj_probe_pairs_grouped on the received rows, which skip the engine's own
 * partition pass. */
DWJ_API uint32_t dwj_xpart_regions(const dwj_engine *e, uint32_t n_ranks);
/* d_counts[n_ranks * regions] (uint64, device). */
DWJ_API int dwj_xpart_hist(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_ranks, uint64_t *d_counts, void *stream);
/* Second half, after the exchange has been planned: the rows of partition p are written to d_out_keys / d_out_vals
 * [start_rows[p] ...] (HOST array of n_ranks * regions row offsets; runs must not overlap).  The runs may lie anywhere
 * in the allocation the out pointers address -- typically a send buffer for the other ranks' partitions and this
 * rank's own receive buffer for its own. */
DWJ_API int dwj_xpart_scatter(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, uint32_t n_ranks,
                      const uint64_t *start_rows, void *d_out_keys, void *d_out_vals, void *stream);
/* dwj_build / dwj_probe_pairs for rows that are ALREADY grouped by table region (region-major): no partition pass.
 * d_region_offsets (device, regions + 1 uint64 row offsets, may be NULL) enables the build's L2 look-ahead.  The
 * caller's row order carries no meaning here, so the result rows are emitted in no particular order (same multiset;
 * one atomicAdd per warp instead of the order-preserving scan) -- as for any probe of a region-partitioned table. */
DWJ_API int dwj_build_grouped(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows,
                      const uint64_t *d_region_offsets, void *stream);
DWJ_API int dwj_probe_pairs_grouped(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_key,
                            void *d_out_build_val, void *d_out_probe_val, uint64_t capacity, uint64_t *d_n_matches,
                            uint64_t *n_matches, void *stream);
/* The same for rows that arrive as a LIST OF SEGMENTS -- segment i = seg_rows[i] rows starting at seg_keys[i] /
 * seg_vals[i] (host arrays of DEVICE pointers, at most 512 segments) -- consumed in list order.  A segment may live in
 * a PEER GPU's memory mapped into this process (NVLink P2P): the kernels then read ("pull") their rows straight out of
 * the sender's buffer, which is how the multi-GPU join moves data (dwj_xj_*): one segment per (table region, source
 * rank), walked region by region -- (region 0, source 0), (region 0, source 1), ... -- so that the table is still built /
 * probed one L2-resident region at a time without any intermediate copy.  segments_per_region (build; 0 = unknown) tells
 * the look-ahead that segments [r * segments_per_region, (r+1) * segments_per_region) belong to table region r.  The
 * probe variant is implemented for DWJ_FLAG_UNIQUE_BUILD_KEYS engines. */
DWJ_API int dwj_build_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals,
                       const uint64_t *seg_rows, uint32_t segments_per_region, void *stream);
DWJ_API int dwj_probe_pairs_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals,
                             const uint64_t *seg_rows, void *d_out_key, void *d_out_build_val, void *d_out_probe_val,
                             uint64_t capacity, uint64_t *d_n_matches, uint64_t *n_matches, void *stream);
/* Pull scatter: the rows of the segments (as above; local or peer memory) are grouped by the table region of THIS
 * engine's table: the rows of region g are written to d_out_keys / d_out_vals[start_rows[g] ...] (HOST array of
 * dwj_info.radix_parts row offsets; the caller knows the region sizes from the senders' dwj_xpart_hist2 counts).  The
 * result feeds dwj_build_grouped / dwj_probe_pairs_grouped.  This is the receiving end of the exchange when ranks x
 * regions exceeds 512 and the senders therefore group by destination rank only.  seg_vals / d_out_vals may both be
 * NULL.  Asynchronous. */
DWJ_API int dwj_region_scatter_segments(dwj_engine *e, uint32_t n_segments, const void *const *seg_keys, const void *const *seg_vals,
                                const uint64_t *seg_rows, const uint64_t *start_rows, void *d_out_keys, void *d_out_vals,
                                void *stream);
/* Histogram over (destination rank, table region of the destination's table) with ALL of the engine's region bits,
 * whether or not dwj_xpart_regions() folds them into the scatter: d_counts[n_ranks * dwj_info.radix_parts] (uint64,
 * device), rank-major.  The senders count for the receivers.  n_ranks: 1, 2, 4 or 8.  Asynchronous. */
DWJ_API int dwj_xpart_hist2(dwj_engine *e, const void *d_keys, uint64_t n_rows, uint32_t n_ranks, uint64_t *d_counts, void *stream);

/* Hash aggregation -- GROUP BY key, SUM(value): the table is cleared and ends up holding one (key, sum) slot per distinct
 * key; sums wrap at the value width.  Replaces kernel `hash_build` of the GroupBy dwarf (groupby/groupby.cpp:60-72) =
 * NonOwningHashTableNonBitmask::add (common/dpcpp/hashtable.hpp:136-153).  Read the result back with dwj_probe_aligned
 * (d_out_build_val = the sum of a queried key; SimpleNonOwningHashTable::at as in groupby.cpp:84-92) or
 * dwj_probe_contains.  max_build_rows of the engine bounds the number of DISTINCT keys.  Rows are first combined per
 * CTA in shared memory, so a handful of groups does not serialise on a handful of global addresses.  Asynchronous. */
DWJ_API int dwj_aggregate_sum(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *stream);

/* The rows of the current pass filter's key class (all rows without a filter), copied to d_out_keys / d_out_vals in no
 * particular order; *d_n_out (device, uint64) = rows kept.  The out columns must hold n_rows rows in the worst case.  One
 * streaming pass, no histogram: a join that runs as passes over key classes (DWJ_OPT_PASS_FILTER) on ONE GPU compacts a
 * relation's class first and partitions the compact copy -- the many-way scatter writes full tiles again instead of the
 * class's half.  d_region_counts (optional; device, dwj_info.radix_parts uint64) receives the kept rows per table
 * region, i.e. what dwj_xpart_hist2 with one rank would count on the compact copy.  d_vals / d_out_vals may both be NULL.
 * Asynchronous. */
DWJ_API int dwj_filter_rows(dwj_engine *e, const void *d_keys, const void *d_vals, uint64_t n_rows, void *d_out_keys, void *d_out_vals,
                    uint64_t *d_n_out, uint64_t *d_region_counts, void *stream);

/* Clears the table on `stream` AHEAD of the next dwj_build* call, which then waits for this clear instead of doing its own:
 * the clear (a pure HBM write of the whole table) can overlap whatever produces the build rows -- a partition pass, an
 * exchange -- on another stream.  The table must not be probed between the two calls.  Asynchronous. */
DWJ_API int dwj_clear_table(dwj_engine *e, void *stream);

/* Engine options (dwj_set_option) */
#define DWJ_OPT_APPEND_OUTPUT 1 /* value != 0: dwj_probe_pairs* append their rows at the running count in *d_n_matches
                                   (device, required) instead of starting from zero -- several probe calls (chunks of a
                                   relation, passes over key classes) fill ONE compact result without a host sync.  Rows
                                   are then emitted in no particular order.                                          */
#define DWJ_OPT_PASS_FILTER 2   /* value = rank_bits | pass_bits << 8 | pass_id << 16.  A join whose working set exceeds
                                   one GPU runs as 2^pass_bits passes over key CLASSES (the pass_bits bits of the
                                   partition hash below the rank_bits destination-rank bits): while the filter is set,
                                   dwj_build*, dwj_partition*, dwj_xpart_* and the region partition inside dwj_probe_pairs /
                                   dwj_probe_count skip the rows of every other class.  pass_bits == 0 clears it.
                                   A one-to-many engine whose class turns out larger than max_build_rows (classes are
                                   hash bits: an even split, not an exact one) leaves the keys that no longer fit
                                   unmatched; the next probe that returns a host count reports DWJ_ERR_CAPACITY.        */
DWJ_API int dwj_set_option(dwj_engine *e, int option, uint64_t value);

/* ---- multi-GPU join -------------------------------------------------------------------------------------------------
 * No reference counterpart (one sycl::queue on one device, join/join.cpp:23-24).  Equal keys must meet on one GPU: both
 * relations are radix-partitioned on the key hash and exchanged over NVLink, then every GPU builds and probes locally.
 * The exchange is a PULL: a sender groups its rows by destination inside its own peer-mapped block and raises a flag in
 * the peers' memory; the receiver copies its rows out of every sender's block with one kernel that reads all peers at
 * once, then builds / probes the landed rows through segment lists (one GPU, or DWJ_XJ_FUSED_PULL=1: the build / probe /
 * region-scatter kernels read the senders' blocks themselves).  Counts and flags travel through the same blocks -- no
 * collective library on the data path.  (csrc/dwj_xj.cu)
 *
 * dwj_xj = one rank (one GPU).  The caller supplies `world` equally sized blocks of dwj_xj_block_bytes() bytes, one in
 * every rank's memory, ALL mapped into this process (blocks[r] = this process's pointer to rank r's block): peer
 * memory from cudaDeviceEnablePeerAccess / cudaIpcOpenMemHandle / torch symmetric memory.  Every rank creates its
 * dwj_xj over an engine created with the same key width, table size and hash seed (verified at every join), and the
 * ranks must be synchronised between the creates and the first join.  dwj_xj_join is collective: every rank calls it
 * for every step (one host thread per rank, or one process per rank); it blocks once per pass on the exchange of the
 * counts and otherwise only enqueues work.  The result of a rank -- its share of the global join, compacted, in no
 * particular order -- is on `stream` when the call returns: rows in d_out_*, their number in *d_n_matches (device). */
typedef struct dwj_xj dwj_xj;
typedef struct {
  int32_t rank, world;        /* world: 1, 2, 4 or 8                                                             */
  uint64_t max_build_rows;    /* most rows of the build relation THIS rank passes to one dwj_xj_join (same value   */
  uint64_t max_probe_rows;    /*   on every rank); sizes the send slots                                           */
  uint64_t chunk_rows;        /* the probe relation travels in pieces of this many rows: the sender partitions piece
                                 c+1 while the receivers pull piece c.  0 = 2^26 for tables up to 512 MB; for larger
                                 tables (re-read from HBM once per piece) two pieces, one piece when world == 1   */
  uint32_t passes;            /* power of two; > 1: the join runs once per key class (DWJ_OPT_PASS_FILTER) with the
                                 table, slots and landing buffers sized for one class -- for working sets larger than
                                 the GPUs' memory.  0 = 1                                                        */
  uint32_t force_scatter_pull;/* testing: take the region-scatter receive path even where the direct pull applies */
  double recv_slack;          /* head-room of the receive-side landing buffers over an even split (0 = 1.25)     */
} dwj_xj_config;
typedef struct {
  uint32_t regions;           /* table regions of every rank's table                                             */
  uint32_t fold_regions;      /* regions grouped by the SENDER's pass (== regions: the receiver pulls directly)   */
  uint32_t chunks, ring;      /* probe chunks per join; send slots they rotate through                           */
  uint32_t passes;
  uint32_t direct_pull;       /* 1: build / probe kernels consume the rows per (region, source) segment; 0: a region
                                 scatter groups them first (ranks x regions > 512, or duplicate build keys)      */
  uint32_t copy_pull;         /* 1: a copy kernel first moves the rows out of the senders' slots over NVLink (world > 1);
                                 0: the consuming kernels read the slots themselves (world == 1, DWJ_XJ_FUSED_PULL=1) */
  uint32_t compact_passes;    /* 1: one GPU, passes > 1, one probe chunk: every pass first compacts its key class out of
                                 the input (dwj_filter_rows) and partitions the compact copy                    */
  uint32_t reserved;
  uint64_t chunk_rows;
  uint64_t block_bytes;       /* = dwj_xj_block_bytes                                                            */
  uint64_t landing_bytes;     /* local buffers of the region-scatter receive path                                */
} dwj_xj_info;
typedef struct {              /* device timeline of the last join on this rank, ms from its start (first pass)   */
  float counts_ms;            /* counts of every batch computed, exchanged and read by the host                  */
  float scattered_ms;         /* last batch grouped into its send slot                                           */
  float built_ms;             /* local table built                                                               */
  float total_ms;             /* last probe chunk done (all passes)                                              */
  float build_pulled_ms;      /* copy pull: the build relation's rows have landed                                 */
  float last_pulled_ms;       /* copy pull: the last probe chunk's rows have landed                               */
  uint64_t remote_bytes;      /* bytes this rank pulled over NVLink                                              */
} dwj_xj_timing;
DWJ_API int dwj_xj_block_bytes(const dwj_engine *e, const dwj_xj_config *cfg, uint64_t *bytes);
DWJ_API int dwj_xj_create(dwj_engine *e, const dwj_xj_config *cfg, void *const *blocks, dwj_xj **out);
DWJ_API int dwj_xj_destroy(dwj_xj *x);
DWJ_API int dwj_xj_describe(const dwj_xj *x, dwj_xj_info *info);
DWJ_API int dwj_xj_join(dwj_xj *x, const void *d_build_keys, const void *d_build_vals, uint64_t n_build, const void *d_probe_keys,
                const void *d_probe_vals, uint64_t n_probe, void *d_out_key, void *d_out_build_val, void *d_out_probe_val,
                uint64_t capacity, uint64_t *d_n_matches, void *stream);
/* Waits for the last join of this rank and returns its timeline (also the place a peer time-out surfaces). */
DWJ_API int dwj_xj_sync_timings(dwj_xj *x, dwj_xj_timing *t);

/* The layout plan of one batch as pure host functions (what dwj_xj_join computes from the exchanged counts; exported so
 * the logic can be tested without a GPU).  Sender: mine[dst][region] -> start[] = first slot row of every partition of
 * its pass (world * fold_regions entries; destination-major, region-minor).  Receiver `me`: tot[src][dst], reg[src][my
 * region] -> segment lists that visit the sources in rotated order me, me+1, ... (so the ranks never all read from one
 * GPU): direct != 0: regions * world segments in walking order (region-major, source-minor); direct == 0: every
 * source's block cut into up to `pieces` pieces dealt round-robin over the sources, plus region_start[regions] of the
 * landing buffer.  seg_src[i] = source of segment i; at most max(regions, pieces) * world segments.  Rows are relative
 * to the senders' blocks. */
DWJ_API int dwj_xj_plan_send(uint32_t world, uint32_t regions, uint32_t fold_regions, uint64_t slot_base_row, const uint64_t *mine,
                     uint64_t *start);
DWJ_API int dwj_xj_plan_recv(uint32_t world, uint32_t me, uint32_t regions, uint64_t slot_base_row, const uint64_t *tot, const uint64_t *reg,
                     int direct, uint32_t pieces, uint64_t *seg_first_row, uint64_t *seg_rows, uint32_t *seg_src,
                     uint64_t *region_start, uint64_t *total, uint32_t *n_segments);
/* Table region (0 .. 2^region_bits - 1) a key falls into for a table of `buckets` 32-byte buckets -- host evaluation of
 * the kernels' function, as dwj_partition_of. */
DWJ_API uint32_t dwj_region_of(uint64_t key, int32_t key_bytes, uint64_t buckets, uint32_t region_bits, uint64_t hash_seed);

/* All GPUs of one box from ONE process (the C++ host framework's `dwarf_bench Join --device=gpu --gpus N`): one engine
 * and one dwj_xj per GPU, peer access enabled both ways, one host thread per GPU inside dwj_mg_join.  devices[] may
 * name one GPU several times (ranks sharing a GPU: testing). */
typedef struct dwj_mg dwj_mg;
typedef struct {
  int32_t n_gpus;             /* 1, 2, 4 or 8                                                                    */
  int32_t devices[8];
  int32_t key_bytes;          /* 4 or 8; payloads have the same width                                            */
  uint32_t flags;             /* DWJ_FLAG_* of the engines                                                       */
  uint64_t max_build_rows_per_gpu, max_probe_rows_per_gpu;   /* input rows a GPU holds before the exchange       */
  double load_factor;         /* of the tables at an even split (0 = 0.5)                                        */
  uint64_t hash_seed;
  uint64_t chunk_rows;        /* as dwj_xj_config                                                                */
  uint32_t passes;
  uint32_t force_scatter_pull;
  double recv_slack;
} dwj_mg_config;
typedef struct {              /* slowest GPU per mark, ms from the start of the join                             */
  float counts_ms, partition_ms, build_ms, total_ms;
  uint64_t remote_bytes;      /* bytes pulled over NVLink, all GPUs                                              */
} dwj_mg_timing;
DWJ_API int dwj_mg_create(const dwj_mg_config *cfg, dwj_mg **out);
DWJ_API int dwj_mg_destroy(dwj_mg *m);
DWJ_API int dwj_mg_describe(const dwj_mg *m, uint32_t rank, dwj_xj_info *info);
/* Per-GPU DEVICE columns in, per-GPU compacted rows out (arrays of n_gpus entries; d_out_key may be NULL); n_out[r] =
 * rows GPU r produced.  Synchronous. */
DWJ_API int dwj_mg_join(dwj_mg *m, const void *const *d_build_keys, const void *const *d_build_vals, const uint64_t *n_build,
                const void *const *d_probe_keys, const void *const *d_probe_vals, const uint64_t *n_probe, void *const *d_out_key,
                void *const *d_out_build_val, void *const *d_out_probe_val, const uint64_t *capacity, uint64_t *n_out,
                dwj_mg_timing *timing);
/* HOST columns in and out: the rows are dealt to the GPUs in arrival order (not by key), joined, and the GPUs' result
 * rows returned one GPU after the other.  Unique build keys (one output row per probe row at most). */
DWJ_API int dwj_mg_join_host(dwj_mg *m, const void *build_keys, const void *build_vals, uint64_t n_build, const void *probe_keys,
                     const void *probe_vals, uint64_t n_probe, void *out_key, void *out_build_val, void *out_probe_val,
                     uint64_t out_capacity, uint64_t *n_out, dwj_mg_timing *timing);

/* Partition id of one key on the host (same function the kernels use) -- lets callers and tests
 * reason about placement without a device. */
DWJ_API uint32_t dwj_partition_of(uint64_t key, int32_t key_bytes, uint32_t n_parts, uint64_t hash_seed);

#ifdef __cplusplus
}
#endif
#endif /* DWJ_H */
