"""ctypes bindings for the CPU checker libraries under oracle/.

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / --impl reference legs -- never by dwarf_bench_b200.

  Oracle  -> oracle/_build/libjoin_oracle.so  (our C restatement, join_oracle.c)
  Ref     -> oracle/_ref/libref_join.so       (the reference's own headers compiled
                                               unmodified against oracle/sycl_shim)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libjoin_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_join.so")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checker libraries (oracle always; _ref when the reference checkout exists)."""
    targets = ["all"] + (["ref"] if ref and os.path.isdir(os.environ.get("DWARF_REF", "/root/reference")) else [])
    subprocess.run(["make", "-C", _HERE, *targets], check=True, stdout=subprocess.DEVNULL)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class JoinTiming(C.Structure):
    _fields_ = [("build_us", C.c_double), ("probe_us", C.c_double), ("host_us", C.c_double),
                ("threads", C.c_int)]


class TableStruct(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("vals", C.c_void_p), ("bitmask", C.c_void_p),
                ("size", C.c_uint64), ("bitmask_sz", C.c_uint64), ("hash_kind", C.c_int),
                ("seed", C.c_uint32)]


HASH_MURMUR, HASH_MODULO = 0, 1


class Oracle:
    """join_oracle.c through ctypes."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = lib = C.CDLL(path)
        lib.dwo_murmur3_x86_32.restype = C.c_uint32
        lib.dwo_murmur3_x86_32.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        lib.dwo_murmur3_slot.restype = C.c_uint64
        lib.dwo_murmur3_slot.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint64]
        for suf, p in (("u32", _u32p), ("u64", _u64p)):
            for name in ("dwo_seq_join_", "dwo_sort_join_"):
                f = getattr(lib, name + suf)
                f.restype = C.c_uint64
                f.argtypes = [p, p, C.c_uint64, p, p, C.c_uint64, p, p, p, C.c_uint64]
            f = getattr(lib, "dwo_rows_equal_" + suf)
            f.restype = C.c_int
            f.argtypes = [p, p, p, C.c_uint64, p, p, p, C.c_uint64]
        lib.dwo_table_insert.restype = C.c_uint32
        lib.dwo_table_insert.argtypes = [C.POINTER(TableStruct), C.c_uint32, C.c_uint32]
        lib.dwo_table_at.restype = C.c_int
        lib.dwo_table_at.argtypes = [C.POINTER(TableStruct), C.c_uint32, C.POINTER(C.c_uint32)]
        lib.dwo_table_has.restype = C.c_int
        lib.dwo_table_has.argtypes = [C.POINTER(TableStruct), C.c_uint32]
        lib.dwo_join_build_probe_u32.restype = C.c_int
        lib.dwo_join_build_probe_u32.argtypes = [_u32p, _u32p, C.c_uint64, _u32p, _u32p, C.c_uint64,
                                                 C.c_uint32, _u32p, _u32p, _u32p, C.POINTER(JoinTiming)]
        lib.dwo_compact_u32.restype = C.c_uint64
        lib.dwo_compact_u32.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, _u32p, _u32p, _u32p]
        lib.dwo_hash_build_check_u32.restype = C.c_uint64
        lib.dwo_hash_build_check_u32.argtypes = [_u32p, C.c_uint64, C.c_uint32, C.POINTER(C.c_double),
                                                 C.POINTER(C.c_int)]
        lib.dwo_omnisci_join_u32.restype = C.c_int
        lib.dwo_omnisci_join_u32.argtypes = [_u32p, C.c_uint64, _u32p, C.c_uint64, _u64p, _u64p, _u64p]
        lib.dwo_make_unique_random.restype = None
        lib.dwo_make_unique_random.argtypes = [C.c_uint64, C.c_uint64, _u32p]
        lib.dwo_make_random_u32.restype = None
        lib.dwo_make_random_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _u32p]
        lib.dwo_max_threads.restype = C.c_int

    # -- hashing -----------------------------------------------------------
    def murmur(self, v: int, seed: int, length: int = 4) -> int:
        return int(self.lib.dwo_murmur3_x86_32(v & 0xFFFFFFFF, seed, length))

    def murmur_slot(self, v: int, seed: int, sz: int, length: int = 4) -> int:
        return int(self.lib.dwo_murmur3_slot(v & 0xFFFFFFFF, seed, length, sz))

    # -- joins ---------------------------------------------------------------
    def _join(self, fname, ak, av, bk, bv):
        wide = np.asarray(ak).dtype.itemsize == 8
        conv, dt, suf = (_u64, np.uint64, "u64") if wide else (_u32, np.uint32, "u32")
        ak, av, bk, bv = conv(ak), conv(av), conv(bk), conv(bv)
        f = getattr(self.lib, fname + suf)
        z = np.zeros(1, dt)
        m = int(f(ak, av, len(ak), bk, bv, len(bk), z, z, z, 0))      # count first
        ok, oa, ob = (np.empty(max(m, 1), dt) for _ in range(3))
        m2 = int(f(ak, av, len(ak), bk, bv, len(bk), ok, oa, ob, m))
        assert m == m2
        return ok[:m], oa[:m], ob[:m]

    def seq_join(self, ak, av, bk, bv):
        """O(n*m) restatement of join_helpers::seq_join; rows in its emission order."""
        return self._join("dwo_seq_join_", ak, av, bk, bv)

    def sort_join(self, ak, av, bk, bv):
        """Same multiset in O(n log n); rows sorted by (key, va, vb)."""
        return self._join("dwo_sort_join_", ak, av, bk, bv)

    def rows_equal(self, t1, t2) -> bool:
        wide = np.asarray(t1[0]).dtype.itemsize == 8
        conv, suf = (_u64, "u64") if wide else (_u32, "u32")
        a = [conv(x) for x in t1]
        b = [conv(x) for x in t2]
        return bool(getattr(self.lib, "dwo_rows_equal_" + suf)(*a, len(a[0]), *b, len(b[0])))

    # -- the reference table -------------------------------------------------
    def new_table(self, size: int, hash_kind: int = HASH_MURMUR, seed: int = 0, key_fill: int = 0):
        return Table(self, size, hash_kind, seed, key_fill)

    def join_build_probe(self, ak, av, bk, bv, seed: int = 42):
        ak, av, bk, bv = _u32(ak), _u32(av), _u32(bk), _u32(bv)
        nb = len(bk)
        ok, op, ov = (np.empty(max(nb, 1), np.uint32) for _ in range(3))
        t = JoinTiming()
        rc = self.lib.dwo_join_build_probe_u32(ak, av, len(ak), bk, bv, nb, seed, ok, op, ov, C.byref(t))
        if rc != 0:
            raise MemoryError("dwo_join_build_probe_u32 failed")
        return (ok[:nb], op[:nb], ov[:nb]), {"build_us": t.build_us, "probe_us": t.probe_us,
                                              "host_us": t.host_us, "threads": t.threads}

    def compact(self, ok, op, ov):
        ok, op, ov = _u32(ok), _u32(op), _u32(ov)
        n = len(ok)
        rk, rp, rv = (np.empty(max(n, 1), np.uint32) for _ in range(3))
        m = int(self.lib.dwo_compact_u32(ok, op, ov, n, rk, rp, rv))
        return rk[:m], rp[:m], rv[:m]

    def hash_build_check(self, src, seed: int = 42):
        src = _u32(src)
        us, th = C.c_double(), C.c_int()
        found = int(self.lib.dwo_hash_build_check_u32(src, len(src), seed, C.byref(us), C.byref(th)))
        return found, us.value, th.value

    def omnisci_join(self, ak, bk):
        ak, bk = _u32(ak), _u32(bk)
        ids = np.zeros(max(len(ak), 1), np.uint64)
        off = np.zeros(max(len(bk), 1), np.uint64)
        cnt = np.zeros(max(len(bk), 1), np.uint64)
        rc = self.lib.dwo_omnisci_join_u32(ak, len(ak), bk, len(bk), ids, off, cnt)
        if rc != 0:
            raise MemoryError("dwo_omnisci_join_u32 failed")
        return ids[:len(ak)], off[:len(bk)], cnt[:len(bk)]

    # -- generators ----------------------------------------------------------
    def make_unique_random(self, n: int, seed: int) -> np.ndarray:
        out = np.empty(max(n, 1), np.uint32)
        self.lib.dwo_make_unique_random(n, seed, out)
        return out[:n]

    def make_random(self, n: int, seed: int, lo: int = 1, hi: int = 10000) -> np.ndarray:
        out = np.empty(max(n, 1), np.uint32)
        self.lib.dwo_make_random_u32(n, seed, lo, hi, out)
        return out[:n]

    def max_threads(self) -> int:
        return int(self.lib.dwo_max_threads())


class Table:
    """A SimpleNonOwningHashTable restatement instance (arrays owned here, as in the reference tests)."""

    def __init__(self, oracle: Oracle, size: int, hash_kind: int, seed: int, key_fill: int):
        self.o = oracle
        self.size = size
        self.bitmask_sz = max((size + 31) // 32, 1)
        self.keys = np.full(size, key_fill, np.uint32)
        self.vals = np.zeros(size, np.uint32)
        self.bitmask = np.zeros(self.bitmask_sz, np.uint32)
        self.s = TableStruct(self.keys.ctypes.data, self.vals.ctypes.data, self.bitmask.ctypes.data,
                             size, self.bitmask_sz, hash_kind, seed)

    def insert(self, key: int, val: int) -> int:
        return int(self.o.lib.dwo_table_insert(C.byref(self.s), key, val))

    def at(self, key: int):
        v = C.c_uint32()
        hit = self.o.lib.dwo_table_at(C.byref(self.s), key, C.byref(v))
        return (int(v.value), True) if hit else (0, False)

    def has(self, key: int) -> bool:
        return bool(self.o.lib.dwo_table_has(C.byref(self.s), key))


class Ref:
    """The reference's own headers (oracle/_ref/libref_join.so).  Optional: absent => None."""

    def __init__(self, path: str = REF_SO):
        self.lib = lib = C.CDLL(path)
        lib.ref_max_threads.restype = C.c_int
        lib.ref_murmur_slot.restype = C.c_uint64
        lib.ref_murmur_slot.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint64]
        for suf, p in (("u32", _u32p), ("u64", _u64p)):
            f = getattr(lib, "ref_seq_join_" + suf)
            f.restype = C.c_uint64
            f.argtypes = [p, p, C.c_uint64, p, p, C.c_uint64, p, p, p, C.c_uint64]
        lib.ref_rows_equal_u32.restype = C.c_int
        lib.ref_rows_equal_u32.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, _u32p, _u32p, _u32p, C.c_uint64]
        lib.ref_roundtrip_equal_u32.restype = C.c_int
        lib.ref_roundtrip_equal_u32.argtypes = [_u32p, _u32p, _u32p, C.c_uint64]
        lib.ref_table_insert.restype = None
        lib.ref_table_insert.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, _u32p, _u32p, _u32p,
                                         _u32p, _u32p, C.c_uint64, C.c_void_p, C.c_int]
        lib.ref_table_at.restype = None
        lib.ref_table_at.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, _u32p, _u32p, _u32p,
                                     _u32p, C.c_uint64, _u32p, _u32p, _u32p]
        lib.ref_join_build_probe_u32.restype = C.c_int
        lib.ref_join_build_probe_u32.argtypes = [_u32p, _u32p, C.c_uint64, _u32p, _u32p, C.c_uint64,
                                                 C.c_uint32, _u32p, _u32p, _u32p, _f64p]
        if hasattr(lib, "ref_join_build_probe_u64"):
            lib.ref_join_build_probe_u64.restype = C.c_int
            lib.ref_join_build_probe_u64.argtypes = [_u64p, _u64p, C.c_uint64, _u64p, _u64p, C.c_uint64, _u64p, _u64p, _u64p, _f64p]

    @staticmethod
    def available(path: str = REF_SO) -> bool:
        return os.path.exists(path)

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def murmur_slot(self, v: int, seed: int, sz: int, length: int = 4) -> int:
        return int(self.lib.ref_murmur_slot(v & 0xFFFFFFFF, seed, length, sz))

    def seq_join(self, ak, av, bk, bv):
        wide = np.asarray(ak).dtype.itemsize == 8
        conv, dt, suf = (_u64, np.uint64, "u64") if wide else (_u32, np.uint32, "u32")
        ak, av, bk, bv = conv(ak), conv(av), conv(bk), conv(bv)
        f = getattr(self.lib, "ref_seq_join_" + suf)
        cap = max(len(ak) * 4 + len(bk) * 4, 16)
        while True:
            ok, oa, ob = (np.empty(cap, dt) for _ in range(3))
            m = int(f(ak, av, len(ak), bk, bv, len(bk), ok, oa, ob, cap))
            if m <= cap:
                return ok[:m], oa[:m], ob[:m]
            cap = m

    def rows_equal(self, t1, t2) -> bool:
        a = [_u32(x) for x in t1]
        b = [_u32(x) for x in t2]
        return bool(self.lib.ref_rows_equal_u32(*a, len(a[0]), *b, len(b[0])))

    def roundtrip_equal(self, t) -> bool:
        a = [_u32(x) for x in t]
        return bool(self.lib.ref_roundtrip_equal_u32(*a, len(a[0])))

    def table_insert(self, size, in_k, in_v, hash_kind=HASH_MURMUR, seed=0, key_fill=0, parallel=False):
        bitmask_sz = max((size + 31) // 32, 1)
        keys = np.full(size, key_fill, np.uint32)
        vals = np.zeros(size, np.uint32)
        bitmask = np.zeros(bitmask_sz, np.uint32)
        in_k, in_v = _u32(in_k), _u32(in_v)
        slots = np.zeros(max(len(in_k), 1), np.uint32)
        self.lib.ref_table_insert(hash_kind, seed, size, bitmask_sz, keys, vals, bitmask, in_k, in_v,
                                  len(in_k), slots.ctypes.data, int(parallel))
        return keys, vals, bitmask, slots[:len(in_k)]

    def table_at(self, size, keys, vals, bitmask, q, hash_kind=HASH_MURMUR, seed=0):
        q = _u32(q)
        found, val, has = (np.zeros(max(len(q), 1), np.uint32) for _ in range(3))
        self.lib.ref_table_at(hash_kind, seed, size, len(bitmask), keys, vals, bitmask, q, len(q),
                              found, val, has)
        return found[:len(q)], val[:len(q)], has[:len(q)]

    def join_build_probe(self, ak, av, bk, bv, seed: int = 42):
        nb = len(bk)
        t = np.zeros(3, np.float64)
        if np.asarray(ak).dtype.itemsize == 8:      # the reference's templates at 64 bits (ref_driver.cpp), SimpleHasher
            ak, av, bk, bv = _u64(ak), _u64(av), _u64(bk), _u64(bv)
            ok, op, ov = (np.empty(max(nb, 1), np.uint64) for _ in range(3))
            if self.lib.ref_join_build_probe_u64(ak, av, len(ak), bk, bv, nb, ok, op, ov, t) != 0:
                raise ValueError("the reference's uint32 slot index cannot address this table")
            return (ok[:nb], op[:nb], ov[:nb]), {"build_us": t[0], "probe_us": t[1], "host_us": t[2],
                                                  "threads": self.max_threads()}
        ak, av, bk, bv = _u32(ak), _u32(av), _u32(bk), _u32(bv)
        ok, op, ov = (np.empty(max(nb, 1), np.uint32) for _ in range(3))
        self.lib.ref_join_build_probe_u32(ak, av, len(ak), bk, bv, nb, seed, ok, op, ov, t)
        return (ok[:nb], op[:nb], ov[:nb]), {"build_us": t[0], "probe_us": t[1], "host_us": t[2],
                                              "threads": self.max_threads()}


def canonical_rows(*cols):
    """Rows sorted lexicographically by (col0, col1, ...) -- the form join_helpers' eq() compares in."""
    cols = [np.asarray(c) for c in cols]
    if len(cols[0]) == 0:
        return tuple(cols)
    order = np.lexsort(tuple(reversed(cols)))
    return tuple(c[order] for c in cols)


def groupby_sum(keys, vals, groups_count: int):
    """expected_GroupBy (ref:groupby/groupby.cpp:8-19) with f = x + y: result[k] += v over all rows, in the value type's
    wrap-around arithmetic.  TEST INFRASTRUCTURE ONLY."""
    keys, vals = np.asarray(keys), np.asarray(vals)
    dt = vals.dtype
    out = np.zeros(groups_count, dtype=np.uint64)
    np.add.at(out, keys.astype(np.int64), vals.astype(np.uint64))
    return (out & np.uint64(np.iinfo(dt).max)).astype(dt)
