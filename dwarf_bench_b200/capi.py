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
           "dwj_join_host", "dwj_partition", "dwj_partition_hist", "dwj_partition_scatter_to", "dwj_partition_of",
           "dwj_xpart_regions", "dwj_xpart_hist", "dwj_xpart_scatter", "dwj_build_grouped", "dwj_probe_pairs_grouped",
           "dwj_copy_many", "dwj_push_runs", "dwj_build_segments", "dwj_probe_pairs_segments")


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
                ("radix_parts", C.c_uint32), ("probe_passes", C.c_uint32)]


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
    lib.dwj_partition_scatter_to.argtypes = [vp, vp, vp, u64, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), vp]
    lib.dwj_partition_of.argtypes = [u64, C.c_int32, u32, u64]
    lib.dwj_partition_of.restype = u32
    lib.dwj_xpart_regions.argtypes = [vp, u32]
    lib.dwj_xpart_regions.restype = u32
    lib.dwj_xpart_hist.argtypes = [vp, vp, u64, u32, vp, vp]
    lib.dwj_xpart_scatter.argtypes = [vp, vp, vp, u64, u32, C.POINTER(u64), vp, vp, vp]
    lib.dwj_build_grouped.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.dwj_probe_pairs_grouped.argtypes = [vp, vp, vp, u64, vp, vp, vp, u64, vp, C.POINTER(u64), vp]
    lib.dwj_copy_many.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(vp)]
    lib.dwj_push_runs.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), u32, vp]
    lib.dwj_build_segments.argtypes = [vp, vp, vp, u32, C.POINTER(u64), C.POINTER(u64), u32, vp]
    lib.dwj_probe_pairs_segments.argtypes = [vp, vp, vp, u32, C.POINTER(u64), C.POINTER(u64), vp, vp, vp, u64, vp, C.POINTER(u64), vp]
    for name in SYMBOLS:
        f = getattr(lib, name)
        if name not in ("dwj_last_error", "dwj_partition_of", "dwj_xpart_regions"):
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

    def partition_scatter_to(self, d_keys, d_vals, n_rows: int, n_parts: int, dst_keys, dst_vals, dst_row_offsets,
                             stream=None) -> None:
        """dst_keys / dst_vals: sequences of device pointers (ints), possibly peer memory; dst_row_offsets: ints."""
        pk = (C.c_void_p * n_parts)(*[int(x) for x in dst_keys])
        pv = (C.c_void_p * n_parts)(*[int(x) for x in dst_vals]) if d_vals is not None else None
        off = (C.c_uint64 * n_parts)(*[int(x) for x in dst_row_offsets])
        self._check(self.lib.dwj_partition_scatter_to(self._h, _ptr(d_keys), _ptr(d_vals), n_rows, n_parts, pk, pv, off,
                                                      _stream(stream)))


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

    def build_segments(self, d_keys, d_vals, seg_first_row, seg_rows, segments_per_region: int = 0, stream=None) -> None:
        import numpy as np
        f, r = (np.ascontiguousarray(a, dtype=np.uint64) for a in (seg_first_row, seg_rows))
        u64p = C.POINTER(C.c_uint64)
        self._check(self.lib.dwj_build_segments(self._h, _ptr(d_keys), _ptr(d_vals), len(f), f.ctypes.data_as(u64p),
                                                r.ctypes.data_as(u64p), segments_per_region, _stream(stream)))

    def probe_pairs_segments(self, d_keys, d_vals, seg_first_row, seg_rows, d_out_key, d_out_build_val, d_out_probe_val,
                             capacity: int, d_n_matches=None, sync: bool = True, stream=None):
        import numpy as np
        f, r = (np.ascontiguousarray(a, dtype=np.uint64) for a in (seg_first_row, seg_rows))
        u64p = C.POINTER(C.c_uint64)
        n = C.c_uint64(0)
        rc = self.lib.dwj_probe_pairs_segments(self._h, _ptr(d_keys), _ptr(d_vals), len(f), f.ctypes.data_as(u64p),
                                               r.ctypes.data_as(u64p), _ptr(d_out_key), _ptr(d_out_build_val),
                                               _ptr(d_out_probe_val), capacity, _ptr(d_n_matches), C.byref(n) if sync else None,
                                               _stream(stream))
        self._check(rc)
        return int(n.value) if sync else None

    def copy_many(self, copies) -> None:
        """copies: list of (dst pointer, src pointer, bytes, stream); device-to-device, possibly to peer memory."""
        n = len(copies)
        if not n:
            return
        d = (C.c_void_p * n)(*[int(c[0]) for c in copies])
        s = (C.c_void_p * n)(*[int(c[1]) for c in copies])
        b = (C.c_uint64 * n)(*[int(c[2]) for c in copies])
        st = (C.c_void_p * n)(*[_stream(c[3]) for c in copies])
        self._check(self.lib.dwj_copy_many(self._h, n, d, s, b, st))


    def push_runs(self, dsts, srcs, rows, n_ctas: int = 0, stream=None) -> None:
        """dwj_push_runs from three equally long numpy uint64 arrays (pointers, pointers, row counts)."""
        import numpy as np
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (dsts, srcs, rows)]
        vpp, u64p = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)
        self._check(self.lib.dwj_push_runs(self._h, len(arrs[0]), arrs[0].ctypes.data_as(vpp), arrs[1].ctypes.data_as(vpp),
                                           arrs[2].ctypes.data_as(u64p), n_ctas, _stream(stream)))

    def copy_many_arrays(self, dsts, srcs, nbytes, streams) -> None:
        """dwj_copy_many from four equally long numpy uint64 arrays (no per-copy Python work)."""
        import numpy as np
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (dsts, srcs, nbytes, streams)]
        n = len(arrs[0])
        if not n:
            return
        vpp, u64p = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)
        self._check(self.lib.dwj_copy_many(self._h, n, arrs[0].ctypes.data_as(vpp), arrs[1].ctypes.data_as(vpp),
                                           arrs[2].ctypes.data_as(u64p), arrs[3].ctypes.data_as(vpp)))


def partition_of(key: int, key_bytes: int, n_parts: int, hash_seed: int = 42) -> int:
    """Host evaluation of the partition function (no device needed)."""
    return int(load_library().dwj_partition_of(key, key_bytes, n_parts, hash_seed))
