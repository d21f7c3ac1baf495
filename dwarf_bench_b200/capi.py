"""ctypes binding of include/dwj.h (libdwj_b200.so).

Pointers cross the boundary as plain integers: `tensor.data_ptr()` for device columns,
`ndarray.ctypes.data` / pinned-tensor `data_ptr()` for host columns, `torch.cuda.Stream.cuda_stream`
for streams.  Nothing here computes anything on the CPU; a missing library is a hard error.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))

FLAG_UNIQUE_BUILD_KEYS = 0x1
FLAG_L2_PERSIST = 0x2
FLAG_UNORDERED_OUTPUT = 0x4
FLAG_NO_PARTITION = 0x8
OUT_ALIGNED, OUT_PAIRS, OUT_COUNT = 0, 1, 2

ERR_NAMES = {0: "DWJ_OK", -1: "DWJ_ERR_INVALID", -2: "DWJ_ERR_CUDA", -3: "DWJ_ERR_OOM", -4: "DWJ_ERR_OVERFLOW",
             -5: "DWJ_ERR_STATE", -6: "DWJ_ERR_CAPACITY"}

# Every symbol include/dwj.h declares (tests check the library exports exactly these).
SYMBOLS = ("dwj_abi_version", "dwj_last_error", "dwj_create", "dwj_destroy", "dwj_get_info", "dwj_build",
           "dwj_probe_aligned", "dwj_probe_contains", "dwj_probe_pairs", "dwj_probe_count", "dwj_timings",
           "dwj_join_host", "dwj_partition", "dwj_partition_hist", "dwj_partition_of",
           "dwj_xpart_regions", "dwj_xpart_hist", "dwj_xpart_hist2", "dwj_xpart_scatter", "dwj_build_grouped",
           "dwj_probe_pairs_grouped", "dwj_build_segments", "dwj_probe_pairs_segments", "dwj_region_scatter_segments",
           "dwj_set_option", "dwj_clear_table", "dwj_filter_rows", "dwj_aggregate_sum", "dwj_xj_block_bytes", "dwj_xj_create", "dwj_xj_destroy", "dwj_xj_describe", "dwj_xj_join",
           "dwj_xj_sync_timings", "dwj_xj_plan_send", "dwj_xj_plan_recv", "dwj_region_of", "dwj_mg_create", "dwj_mg_destroy", "dwj_mg_describe", "dwj_mg_join", "dwj_mg_join_host")
ABI_VERSION = 2
OPT_APPEND_OUTPUT, OPT_PASS_FILTER = 1, 2


class DwjError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("key_bytes", C.c_int32), ("payload_bytes", C.c_int32), ("flags", C.c_uint32),
                ("max_build_rows", C.c_uint64), ("load_factor", C.c_double), ("hash_seed", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("build_ms", C.c_float), ("probe_ms", C.c_float), ("partition_ms", C.c_float), ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("probe_kernel_ms", C.c_float), ("build_kernel_ms", C.c_float)]


class Info(C.Structure):
    _fields_ = [("slots", C.c_uint64), ("table_bytes", C.c_uint64), ("build_rows", C.c_uint64), ("slot_bytes", C.c_uint32),
                ("slots_per_bucket", C.c_uint32), ("l2_persist", C.c_uint32), ("sm_count", C.c_uint32),
                ("l2_bytes", C.c_uint64), ("launches_build", C.c_uint32), ("launches_probe", C.c_uint32),
                ("radix_parts", C.c_uint32), ("probe_passes", C.c_uint32), ("flags", C.c_uint32), ("device", C.c_int32),
                ("hash_seed", C.c_uint64), ("max_build_rows", C.c_uint64), ("hot_probe_keys", C.c_uint32), ("reserved", C.c_uint32)]


class XjConfig(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("max_build_rows", C.c_uint64), ("max_probe_rows", C.c_uint64),
                ("chunk_rows", C.c_uint64), ("passes", C.c_uint32), ("force_scatter_pull", C.c_uint32), ("recv_slack", C.c_double)]


class XjInfo(C.Structure):
    _fields_ = [("regions", C.c_uint32), ("fold_regions", C.c_uint32), ("chunks", C.c_uint32), ("ring", C.c_uint32),
                ("passes", C.c_uint32), ("direct_pull", C.c_uint32), ("copy_pull", C.c_uint32), ("compact_passes", C.c_uint32), ("reserved", C.c_uint32),
                ("chunk_rows", C.c_uint64), ("block_bytes", C.c_uint64),
                ("landing_bytes", C.c_uint64)]


class XjTiming(C.Structure):
    _fields_ = [("counts_ms", C.c_float), ("scattered_ms", C.c_float), ("built_ms", C.c_float), ("total_ms", C.c_float),
                ("build_pulled_ms", C.c_float), ("last_pulled_ms", C.c_float), ("remote_bytes", C.c_uint64)]


class MgConfig(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("devices", C.c_int32 * 8), ("key_bytes", C.c_int32), ("flags", C.c_uint32),
                ("max_build_rows_per_gpu", C.c_uint64), ("max_probe_rows_per_gpu", C.c_uint64), ("load_factor", C.c_double),
                ("hash_seed", C.c_uint64), ("chunk_rows", C.c_uint64), ("passes", C.c_uint32), ("force_scatter_pull", C.c_uint32),
                ("recv_slack", C.c_double)]


class MgTiming(C.Structure):
    _fields_ = [("counts_ms", C.c_float), ("partition_ms", C.c_float), ("build_ms", C.c_float), ("total_ms", C.c_float),
                ("remote_bytes", C.c_uint64)]


@dataclass
class JoinTiming:
    build_ms: float = 0.0
    probe_ms: float = 0.0
    partition_ms: float = 0.0
    h2d_ms: float = 0.0
    d2h_ms: float = 0.0
    total_ms: float = 0.0
    probe_kernel_ms: float = 0.0
    build_kernel_ms: float = 0.0


def lib_path() -> str:
    return os.environ.get("DWJ_LIBRARY", os.path.join(_HERE, "lib", "libdwj_b200.so"))


_lib = None


def load_library():
    """Load libdwj_b200.so.  No fallback: a missing or unloadable library raises."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `make -C dwarf_bench_b200/csrc` "
                                "(or __graft_entry__.build()); this engine has no CPU fallback")
    lib = C.CDLL(path)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.dwj_abi_version.restype = C.c_int
    lib.dwj_last_error.restype = C.c_char_p
    lib.dwj_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.dwj_destroy.argtypes = [vp]
    lib.dwj_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.dwj_build.argtypes = [vp, vp, vp, u64, vp]
    lib.dwj_probe_aligned.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
    lib.dwj_probe_contains.argtypes = [vp, vp, u64, vp, vp]
    lib.dwj_probe_pairs.argtypes = [vp, vp, vp, u64, vp, vp, vp, u64, vp, C.POINTER(u64), vp]
    lib.dwj_probe_count.argtypes = [vp, vp, u64, vp, C.POINTER(u64), vp]
    lib.dwj_timings.argtypes = [vp, C.POINTER(Timing)]
    lib.dwj_join_host.argtypes = [vp, vp, vp, u64, vp, vp, u64, C.c_int, vp, vp, vp, u64, C.POINTER(u64),
                                  C.POINTER(Timing)]
    lib.dwj_partition.argtypes = [vp, vp, vp, u64, u32, vp, vp, vp, vp]
    lib.dwj_partition_hist.argtypes = [vp, vp, u64, u32, vp, vp]
    lib.dwj_partition_of.argtypes = [u64, C.c_int32, u32, u64]
    lib.dwj_partition_of.restype = u32
    lib.dwj_xpart_regions.argtypes = [vp, u32]
    lib.dwj_xpart_regions.restype = u32
    lib.dwj_xpart_hist.argtypes = [vp, vp, u64, u32, vp, vp]
    lib.dwj_xpart_scatter.argtypes = [vp, vp, vp, u64, u32, C.POINTER(u64), vp, vp, vp]
    lib.dwj_build_grouped.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.dwj_probe_pairs_grouped.argtypes = [vp, vp, vp, u64, vp, vp, vp, u64, vp, C.POINTER(u64), vp]
    vpp, u64p = C.POINTER(vp), C.POINTER(u64)
    lib.dwj_xpart_hist2.argtypes = [vp, vp, u64, u32, vp, vp]
    lib.dwj_build_segments.argtypes = [vp, u32, vpp, vpp, u64p, u32, vp]
    lib.dwj_probe_pairs_segments.argtypes = [vp, u32, vpp, vpp, u64p, vp, vp, vp, u64, vp, u64p, vp]
    lib.dwj_region_scatter_segments.argtypes = [vp, u32, vpp, vpp, u64p, u64p, vp, vp, vp]
    lib.dwj_set_option.argtypes = [vp, C.c_int, u64]
    lib.dwj_clear_table.argtypes = [vp, vp]
    lib.dwj_aggregate_sum.argtypes = [vp, vp, vp, u64, vp]
    lib.dwj_filter_rows.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp]
    lib.dwj_xj_block_bytes.argtypes = [vp, C.POINTER(XjConfig), u64p]
    lib.dwj_xj_create.argtypes = [vp, C.POINTER(XjConfig), vpp, C.POINTER(vp)]
    lib.dwj_xj_destroy.argtypes = [vp]
    lib.dwj_xj_describe.argtypes = [vp, C.POINTER(XjInfo)]
    lib.dwj_xj_join.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp, vp, vp, u64, vp, vp]
    lib.dwj_xj_sync_timings.argtypes = [vp, C.POINTER(XjTiming)]
    lib.dwj_xj_plan_send.argtypes = [u32, u32, u32, u64, u64p, u64p]
    lib.dwj_xj_plan_recv.argtypes = [u32, u32, u32, u64, u64p, u64p, C.c_int, u32, u64p, u64p, C.POINTER(u32), u64p, u64p, C.POINTER(u32)]
    lib.dwj_region_of.argtypes = [u64, C.c_int32, u64, u32, u64]
    lib.dwj_region_of.restype = u32
    lib.dwj_mg_create.argtypes = [C.POINTER(MgConfig), C.POINTER(vp)]
    lib.dwj_mg_destroy.argtypes = [vp]
    lib.dwj_mg_describe.argtypes = [vp, u32, C.POINTER(XjInfo)]
    lib.dwj_mg_join.argtypes = [vp, vpp, vpp, u64p, vpp, vpp, u64p, vpp, vpp, vpp, u64p, u64p, C.POINTER(MgTiming)]
    lib.dwj_mg_join_host.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp, vp, vp, u64, u64p, C.POINTER(MgTiming)]
    for name in SYMBOLS:
        f = getattr(lib, name)
        if name not in ("dwj_last_error", "dwj_partition_of", "dwj_xpart_regions", "dwj_region_of"):
            f.restype = C.c_int
    _lib = lib
    return lib


def _ptr(x) -> int | None:
    """Device/host pointer of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take a pointer of {type(x)}")


def _stream(s) -> int | None:
    if s is None:
        return None
    return s if isinstance(s, int) else s.cuda_stream


class Engine:
    """One join table on one GPU (dwj_engine).  Mirrors include/dwj.h call for call."""

    def __init__(self, max_build_rows: int, key_bytes: int = 4, device: int = 0, load_factor: float = 0.0,
                 flags: int = 0, hash_seed: int = 42):
        self.lib = load_library()
        self.key_bytes = key_bytes
        self.device = device
        self._h = C.c_void_p()
        cfg = Config(device, key_bytes, key_bytes, flags, max_build_rows, load_factor, hash_seed)
        self._check(self.lib.dwj_create(C.byref(cfg), C.byref(self._h)))
        self.flags = int(cfg.flags)
        self.key_bytes = int(cfg.key_bytes)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise DwjError(rc, self.lib.dwj_last_error().decode())

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dwj_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def info(self) -> dict:
        i = Info()
        self._check(self.lib.dwj_get_info(self._h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in Info._fields_}

    def build(self, d_keys, d_vals, n_rows: int, stream=None) -> None:
        self._check(self.lib.dwj_build(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _stream(stream)))

    def probe_aligned(self, d_keys, d_vals, n_rows: int, d_out_key, d_out_build_val, d_out_probe_val, stream=None) -> None:
        self._check(self.lib.dwj_probe_aligned(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _ptr(d_out_key),
                                               _ptr(d_out_build_val), _ptr(d_out_probe_val), _stream(stream)))

    def probe_contains(self, d_keys, n_rows: int, d_out_flags, stream=None) -> None:
        self._check(self.lib.dwj_probe_contains(self._h, _ptr(d_keys), n_rows, _ptr(d_out_flags), _stream(stream)))

    def probe_pairs(self, d_keys, d_vals, n_rows: int, d_out_key, d_out_build_val, d_out_probe_val, capacity: int,
                    d_n_matches=None, sync: bool = True, stream=None):
        """Returns the match count when sync=True (and raises DwjError(OVERFLOW) if it exceeds capacity)."""
        n = C.c_uint64(0)
        rc = self.lib.dwj_probe_pairs(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _ptr(d_out_key), _ptr(d_out_build_val),
                                      _ptr(d_out_probe_val), capacity, _ptr(d_n_matches), C.byref(n) if sync else None,
                                      _stream(stream))
        self._check(rc)
        return int(n.value) if sync else None

    def probe_count(self, d_keys, n_rows: int, d_n_matches=None, sync: bool = True, stream=None):
        n = C.c_uint64(0)
        self._check(self.lib.dwj_probe_count(self._h, _ptr(d_keys), n_rows, _ptr(d_n_matches), C.byref(n) if sync else None,
                                             _stream(stream)))
        return int(n.value) if sync else None

    def timings(self) -> JoinTiming:
        t = Timing()
        self._check(self.lib.dwj_timings(self._h, C.byref(t)))
        return JoinTiming(*(getattr(t, f) for f, _ in Timing._fields_))

    def join_host(self, build_keys, build_vals, n_build: int, probe_keys, probe_vals, n_probe: int, out_mode: int,
                  out_key, out_build_val, out_probe_val, out_capacity: int):
        """Host-buffer join (H2D + build + probe + D2H).  Returns (n_out, JoinTiming)."""
        n = C.c_uint64(0)
        t = Timing()
        rc = self.lib.dwj_join_host(self._h, _ptr(build_keys), _ptr(build_vals), n_build, _ptr(probe_keys), _ptr(probe_vals),
                                    n_probe, out_mode, _ptr(out_key), _ptr(out_build_val), _ptr(out_probe_val), out_capacity,
                                    C.byref(n), C.byref(t))
        self._check(rc)
        return int(n.value), JoinTiming(*(getattr(t, f) for f, _ in Timing._fields_))

    def partition(self, d_keys, d_vals, n_rows: int, n_parts: int, d_out_keys, d_out_vals, d_offsets, stream=None) -> None:
        self._check(self.lib.dwj_partition(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, n_parts, _ptr(d_out_keys),
                                           _ptr(d_out_vals), _ptr(d_offsets), _stream(stream)))


    def partition_hist(self, d_keys, n_rows: int, n_parts: int, d_counts, stream=None) -> None:
        self._check(self.lib.dwj_partition_hist(self._h, _ptr(d_keys), n_rows, n_parts, _ptr(d_counts), _stream(stream)))

    # ---- exchange partition folded with the receiver's region grouping (include/dwj.h, dwj_xpart_*) -------------------
    def xpart_regions(self, n_ranks: int) -> int:
        return int(self.lib.dwj_xpart_regions(self._h, n_ranks))

    def xpart_hist(self, d_keys, n_rows: int, n_ranks: int, d_counts, stream=None) -> None:
        self._check(self.lib.dwj_xpart_hist(self._h, _ptr(d_keys), n_rows, n_ranks, _ptr(d_counts), _stream(stream)))

    def xpart_scatter(self, d_keys, d_vals, n_rows: int, n_ranks: int, start_rows, d_out_keys, d_out_vals, stream=None) -> None:
        """start_rows: numpy/sequence of n_ranks * regions row offsets relative to d_out_keys / d_out_vals."""
        import numpy as np
        st = np.ascontiguousarray(start_rows, dtype=np.uint64)
        self._check(self.lib.dwj_xpart_scatter(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, n_ranks,
                                               st.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(d_out_keys), _ptr(d_out_vals),
                                               _stream(stream)))

    def build_grouped(self, d_keys, d_vals, n_rows: int, d_region_offsets=None, stream=None) -> None:
        self._check(self.lib.dwj_build_grouped(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _ptr(d_region_offsets),
                                               _stream(stream)))

    def probe_pairs_grouped(self, d_keys, d_vals, n_rows: int, d_out_key, d_out_build_val, d_out_probe_val, capacity: int,
                            d_n_matches=None, sync: bool = True, stream=None):
        n = C.c_uint64(0)
        rc = self.lib.dwj_probe_pairs_grouped(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _ptr(d_out_key),
                                              _ptr(d_out_build_val), _ptr(d_out_probe_val), capacity, _ptr(d_n_matches),
                                              C.byref(n) if sync else None, _stream(stream))
        self._check(rc)
        return int(n.value) if sync else None

    @staticmethod
    def _seg_arrays(seg_keys, seg_vals, seg_rows):
        import numpy as np
        k, r = np.ascontiguousarray(seg_keys, dtype=np.uint64), np.ascontiguousarray(seg_rows, dtype=np.uint64)
        v = None if seg_vals is None else np.ascontiguousarray(seg_vals, dtype=np.uint64)
        vpp, u64p = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)
        return (k, v, r), len(k), k.ctypes.data_as(vpp), None if v is None else v.ctypes.data_as(vpp), r.ctypes.data_as(u64p)

    def build_segments(self, seg_keys, seg_vals, seg_rows, segments_per_region: int = 0, stream=None) -> None:
        """seg_keys / seg_vals: device pointers (ints) of each segment's first key / payload -- local or peer memory."""
        keep, n, k, v, r = self._seg_arrays(seg_keys, seg_vals, seg_rows)
        self._check(self.lib.dwj_build_segments(self._h, n, k, v, r, segments_per_region, _stream(stream)))

    def probe_pairs_segments(self, seg_keys, seg_vals, seg_rows, d_out_key, d_out_build_val, d_out_probe_val,
                             capacity: int, d_n_matches=None, sync: bool = True, stream=None):
        keep, n, k, v, r = self._seg_arrays(seg_keys, seg_vals, seg_rows)
        cnt = C.c_uint64(0)
        rc = self.lib.dwj_probe_pairs_segments(self._h, n, k, v, r, _ptr(d_out_key), _ptr(d_out_build_val), _ptr(d_out_probe_val),
                                               capacity, _ptr(d_n_matches), C.byref(cnt) if sync else None, _stream(stream))
        self._check(rc)
        return int(cnt.value) if sync else None

    def region_scatter_segments(self, seg_keys, seg_vals, seg_rows, start_rows, d_out_keys, d_out_vals, stream=None) -> None:
        import numpy as np
        keep, n, k, v, r = self._seg_arrays(seg_keys, seg_vals, seg_rows)
        st = np.ascontiguousarray(start_rows, dtype=np.uint64)
        self._check(self.lib.dwj_region_scatter_segments(self._h, n, k, v, r, st.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                         _ptr(d_out_keys), _ptr(d_out_vals), _stream(stream)))

    def xpart_hist2(self, d_keys, n_rows: int, n_ranks: int, d_counts, stream=None) -> None:
        self._check(self.lib.dwj_xpart_hist2(self._h, _ptr(d_keys), n_rows, n_ranks, _ptr(d_counts), _stream(stream)))

    def filter_rows(self, d_keys, d_vals, n_rows: int, d_out_keys, d_out_vals, d_n_out, d_region_counts=None, stream=None) -> None:
        self._check(self.lib.dwj_filter_rows(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _ptr(d_out_keys), _ptr(d_out_vals),
                                             _ptr(d_n_out), _ptr(d_region_counts), _stream(stream)))

    def aggregate_sum(self, d_keys, d_vals, n_rows: int, stream=None) -> None:
        """GROUP BY key, SUM(value) into the table; read back with probe_aligned / probe_contains."""
        self._check(self.lib.dwj_aggregate_sum(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, _stream(stream)))

    def clear_table(self, stream=None) -> None:
        self._check(self.lib.dwj_clear_table(self._h, _stream(stream)))

    def set_option(self, option: int, value: int) -> None:
        self._check(self.lib.dwj_set_option(self._h, option, value))

    def set_pass_filter(self, rank_bits: int, pass_bits: int, pass_id: int) -> None:
        self.set_option(OPT_PASS_FILTER, rank_bits | pass_bits << 8 | pass_id << 16)


class ExchangeJoinRank:
    """dwj_xj: this rank's end of the multi-GPU pull-exchange join (include/dwj.h).  `blocks`: this process's pointers to
    every rank's block (each `block_bytes(...)` bytes, peer-mapped)."""

    def __init__(self, engine: Engine, rank: int, world: int, max_build_rows: int, max_probe_rows: int, blocks,
                 chunk_rows: int = 0, passes: int = 1, recv_slack: float = 0.0, force_scatter_pull: bool = False):
        self.lib = engine.lib
        self.engine = engine
        self.cfg = XjConfig(rank, world, max_build_rows, max_probe_rows, chunk_rows, passes, 1 if force_scatter_pull else 0, recv_slack)
        self._h = C.c_void_p()
        arr = (C.c_void_p * world)(*[int(b) for b in blocks])
        engine._check(self.lib.dwj_xj_create(engine._h, C.byref(self.cfg), arr, C.byref(self._h)))

    @staticmethod
    def block_bytes(engine: Engine, rank: int, world: int, max_build_rows: int, max_probe_rows: int, chunk_rows: int = 0,
                    passes: int = 1, recv_slack: float = 0.0, force_scatter_pull: bool = False) -> int:
        cfg = XjConfig(rank, world, max_build_rows, max_probe_rows, chunk_rows, passes, 1 if force_scatter_pull else 0, recv_slack)
        n = C.c_uint64(0)
        engine._check(engine.lib.dwj_xj_block_bytes(engine._h, C.byref(cfg), C.byref(n)))
        return int(n.value)

    def describe(self) -> dict:
        i = XjInfo()
        self.engine._check(self.lib.dwj_xj_describe(self._h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in XjInfo._fields_}

    def join(self, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe, capacity,
             d_count, stream=None) -> None:
        self.engine._check(self.lib.dwj_xj_join(self._h, _ptr(build_keys), _ptr(build_vals), n_build, _ptr(probe_keys), _ptr(probe_vals),
                                                n_probe, _ptr(out_key), _ptr(out_build), _ptr(out_probe), capacity, _ptr(d_count),
                                                _stream(stream)))

    def sync_timings(self) -> dict:
        t = XjTiming()
        self.engine._check(self.lib.dwj_xj_sync_timings(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in XjTiming._fields_}

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dwj_xj_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuJoin:
    """dwj_mg: all GPUs of the box from this one process (what `dwarf_bench Join --gpus N` calls)."""

    def __init__(self, devices, key_bytes: int, max_build_rows_per_gpu: int, max_probe_rows_per_gpu: int, flags: int = FLAG_UNIQUE_BUILD_KEYS,
                 load_factor: float = 0.0, hash_seed: int = 42, chunk_rows: int = 0, passes: int = 1, recv_slack: float = 0.0,
                 force_scatter_pull: bool = False):
        self.lib = load_library()
        self.n = len(devices)
        self.key_bytes = key_bytes
        cfg = MgConfig()
        cfg.n_gpus = self.n
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        cfg.key_bytes, cfg.flags = key_bytes, flags
        cfg.max_build_rows_per_gpu, cfg.max_probe_rows_per_gpu = max_build_rows_per_gpu, max_probe_rows_per_gpu
        cfg.load_factor, cfg.hash_seed, cfg.chunk_rows, cfg.passes = load_factor, hash_seed, chunk_rows, passes
        cfg.force_scatter_pull, cfg.recv_slack = (1 if force_scatter_pull else 0), recv_slack
        self._h = C.c_void_p()
        self._check(self.lib.dwj_mg_create(C.byref(cfg), C.byref(self._h)))

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise DwjError(rc, self.lib.dwj_last_error().decode())

    def describe(self, rank: int = 0) -> dict:
        i = XjInfo()
        self._check(self.lib.dwj_mg_describe(self._h, rank, C.byref(i)))
        return {f: getattr(i, f) for f, _ in XjInfo._fields_}

    def join_host(self, build_keys, build_vals, probe_keys, probe_vals, out_key, out_build, out_probe):
        """numpy columns in, numpy columns out (out_key may be None).  Returns (rows, timing dict)."""
        n = C.c_uint64(0)
        t = MgTiming()
        self._check(self.lib.dwj_mg_join_host(self._h, _ptr(build_keys), _ptr(build_vals), len(build_keys), _ptr(probe_keys),
                                              _ptr(probe_vals), len(probe_keys), _ptr(out_key), _ptr(out_build), _ptr(out_probe),
                                              len(out_build), C.byref(n), C.byref(t)))
        return int(n.value), {f: getattr(t, f) for f, _ in MgTiming._fields_}

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dwj_mg_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition_of(key: int, key_bytes: int, n_parts: int, hash_seed: int = 42) -> int:
    """Host evaluation of the partition function (no device needed)."""
    return int(load_library().dwj_partition_of(key, key_bytes, n_parts, hash_seed))
