"""Parity of the CUDA engine (through the C ABI, include/dwj.h) with the CPU oracle.

Every test here needs a B200 (`-m gpu`).  Inputs are seeded; the same arrays go to the oracle and the
engine (the reference's own generator is unseeded).  Bar: bit-exact match multiset
(sorted (key, build payload, probe payload) rows, join_helpers.hpp:106-125), and -- for unique build
keys -- bit-exact probe-aligned arrays and bit-exact compaction ORDER (join.cpp:119-129).
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dwj():
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    import dwarf_bench_b200 as d
    d.load_library()        # hard failure when the extension is missing -- no fallback exists
    return d


def dev(a: np.ndarray):
    """numpy uint32/uint64 column -> device tensor (same bits, signed torch dtype)."""
    a = np.ascontiguousarray(a)
    signed = a.view(np.int32 if a.dtype.itemsize == 4 else np.int64)
    return torch.from_numpy(signed).cuda()


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


def empty_like_dev(n, dtype):
    return torch.empty(max(n, 1), dtype=torch.int32 if np.dtype(dtype).itemsize == 4 else torch.int64, device="cuda")


def gpu_join_pairs(dwj, ak, av, bk, bv, unique=False, capacity=None, load_factor=0.0, with_key=True):
    dt = ak.dtype
    flags = dwj.FLAG_UNIQUE_BUILD_KEYS if unique else 0
    with dwj.Engine(max(len(ak), 1), key_bytes=dt.itemsize, flags=flags, load_factor=load_factor) as e:
        dak, dav, dbk, dbv = dev(ak), dev(av), dev(bk), dev(bv)
        e.build(dak, dav, len(ak))
        if capacity is None:
            capacity = e.probe_count(dbk, len(bk))
        ok, oa, ob = (empty_like_dev(capacity, dt) for _ in range(3))
        m = e.probe_pairs(dbk, dbv, len(bk), ok if with_key else None, oa, ob, capacity)
        torch.cuda.synchronize()
        return (host(ok, dt)[:m] if with_key else None), host(oa, dt)[:m], host(ob, dt)[:m]


# ---------------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------------

def test_join_tests_golden_vector(dwj, golden_dir):
    """tests/join_tests.cpp:7-23: 7x7 with a duplicated key on both sides -> exactly 8 rows."""
    g = json.load(open(os.path.join(golden_dir, "join_tests_golden.json")))
    ak, av, bk, bv = (np.array(g[k], np.uint32) for k in ("keys_a", "vals_a", "keys_b", "vals_b"))
    k, a, b = gpu_join_pairs(dwj, ak, av, bk, bv)
    assert len(k) == 8
    got = sorted(zip(k.tolist(), a.tolist(), b.tolist()))
    assert got == sorted(map(tuple, g["rows_emission_order"]))


@pytest.mark.parametrize("case", ["unique128", "unique1024", "unique4096", "dups", "empty_build", "no_match",
                                  "edge_keys", "u64"])
def test_reference_seq_join_fixtures(dwj, golden_dir, case):
    """Rows the reference's own seq_join produced (tests/golden/make_golden.py) == engine rows."""
    g = np.load(os.path.join(golden_dir, "seq_join_golden.npz"))
    ak, av, bk, bv = (g[f"{case}.{n}"] for n in ("ak", "av", "bk", "bv"))
    want = tuple(g[f"{case}.{n}"] for n in ("k", "a", "b"))
    got = pyoracle.canonical_rows(*gpu_join_pairs(dwj, ak, av, bk, bv))
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)
    if case.startswith("unique"):       # the first-hit fast path must agree when keys really are unique
        got = pyoracle.canonical_rows(*gpu_join_pairs(dwj, ak, av, bk, bv, unique=True))
        for w, x in zip(want, got):
            np.testing.assert_array_equal(w, x)


# ---------------------------------------------------------------------------------------------------
# the reference Join dwarf's shape (join/join.cpp), sizes of tests/dwarf_tests/dwarf_tests.cpp:44-50
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [128, 256, 512, 1024, 2048, 4096, 1 << 20])
def test_join_dwarf_shape(dwj, oracle, n):
    ak, av, bk, bv = (oracle.make_unique_random(n, s) for s in (1, 2, 3, 4))
    (wk, wp, wv), _ = oracle.join_build_probe(ak, av, bk, bv, seed=42)       # Join::_run restated
    with dwj.Engine(n, key_bytes=4, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        dak, dav, dbk, dbv = dev(ak), dev(av), dev(bk), dev(bv)
        e.build(dak, dav, n)
        ok, op, ov = (empty_like_dev(n, np.uint32) for _ in range(3))
        e.probe_aligned(dbk, dbv, n, ok, op, ov)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(host(ok, np.uint32), wk)               # probe-aligned, sentinel-filled
        np.testing.assert_array_equal(host(op, np.uint32), wp)
        np.testing.assert_array_equal(host(ov, np.uint32), wv)
        # device compaction == the reference's host compaction loop, same ORDER (join.cpp:119-129)
        ck, cp, cv = oracle.compact(wk, wp, wv)
        m = e.probe_pairs(dbk, dbv, n, ok, op, ov, n)
        torch.cuda.synchronize()
        assert m == len(ck)
        np.testing.assert_array_equal(host(ok, np.uint32)[:m], ck)
        np.testing.assert_array_equal(host(op, np.uint32)[:m], cp)
        np.testing.assert_array_equal(host(ov, np.uint32)[:m], cv)
        assert e.probe_count(dbk, n) == m
        t = e.timings()
        assert t.build_ms > 0 and t.probe_ms > 0
    want = oracle.sort_join(ak, av, bk, bv)
    for w, x in zip(want, pyoracle.canonical_rows(ck, cp, cv)):
        np.testing.assert_array_equal(w, x)


@pytest.mark.parametrize("n", [128, 4096, 300000])
def test_hash_build_dwarf_shape(dwj, oracle, n):
    """HashBuild (hash/hash_build.cpp): keys in [1,10000] with heavy duplicates, val = key, has() == 1 for all."""
    src = oracle.make_random(n, seed=n)
    found, _, _ = oracle.hash_build_check(src, seed=5)
    assert found == n
    with dwj.Engine(n, key_bytes=4) as e:
        d = dev(src)
        e.build(d, d, n)
        flags = torch.empty(n + 64, dtype=torch.int32, device="cuda")
        e.probe_contains(d, n, flags)
        absent = dev(np.arange(10001, 10065, dtype=np.uint32))
        e.probe_contains(absent, 64, flags[n:])
        torch.cuda.synchronize()
        f = flags.cpu().numpy()
        assert f[:n].sum() == n and f[n:].sum() == 0
        # every duplicate took its own slot: matching the source against itself counts sum(c_k^2)
        _, counts = np.unique(src, return_counts=True)
        assert e.probe_count(d, n) == int((counts.astype(np.int64) ** 2).sum())


# ---------------------------------------------------------------------------------------------------
# one-to-many (seq_join semantics), skew, 64-bit
# ---------------------------------------------------------------------------------------------------

def zipf_ranks(rng, n_distinct, size, s=1.0):
    w = 1.0 / np.arange(1, n_distinct + 1) ** s
    cdf = np.cumsum(w) / w.sum()
    return np.searchsorted(cdf, rng.random(size)).clip(0, n_distinct - 1)


@pytest.mark.parametrize("wide", [False, True])
def test_duplicates_x4_zipf_probe(dwj, oracle, wide):
    """BASELINE config 3 at test size: every distinct build key 4 times, Zipf(1.0) probe, compacted output."""
    rng = np.random.default_rng(11)
    dt = np.uint64 if wide else np.uint32
    D, S = 1 << 14, 1 << 17
    distinct = rng.choice(1 << 30, D, replace=False).astype(dt)
    ak = np.repeat(distinct, 4)
    rng.shuffle(ak)
    av = np.arange(len(ak), dtype=dt)
    bk = distinct[zipf_ranks(rng, D, S)]
    bv = np.arange(S, dtype=dt)
    k, a, b = gpu_join_pairs(dwj, ak, av, bk, bv)
    assert len(k) == 4 * S
    want = oracle.sort_join(ak, av, bk, bv)
    for w, x in zip(want, pyoracle.canonical_rows(k, a, b)):
        np.testing.assert_array_equal(w, x)


def test_heavy_multiplicity(dwj, oracle):
    """One key repeated 1000x on the build side: long chains across many buckets, all emitted."""
    rng = np.random.default_rng(3)
    ak = np.concatenate([np.full(1000, 77, np.uint32), rng.choice(100000, 3000, replace=False).astype(np.uint32) + 1000])
    av = rng.integers(0, 2**32, len(ak), dtype=np.uint64).astype(np.uint32)
    bk = np.concatenate([np.full(50, 77, np.uint32), ak[1000:1500], rng.integers(200000, 300000, 500).astype(np.uint32)])
    bv = np.arange(len(bk), dtype=np.uint32)
    got = pyoracle.canonical_rows(*gpu_join_pairs(dwj, ak, av, bk, bv))
    want = oracle.sort_join(ak, av, bk, bv)
    assert len(want[0]) == 50 * 1000 + 500
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


@pytest.mark.parametrize("wide", [False, True])
def test_one_to_many_runs_of_every_length(dwj, oracle, wide):
    """The one-to-many table (csr.cuh; ref:common/dpcpp/omnisci_hashtable.hpp:80-192): build keys with 1..6, 40 and 3000
    duplicates, so that runs end inside the first 16-byte granule, span several, and are emitted both from the staging
    buffer and by the warp-cooperative copy; PAIRS in probe-row order, COUNT and ALIGNED against the oracle."""
    rng = np.random.default_rng(23 + wide)
    dt = np.uint64 if wide else np.uint32
    distinct = rng.choice(1 << 30, 30_000, replace=False).astype(dt)
    mult = rng.integers(1, 7, len(distinct))                   # 3.3 result rows per probe row: most tiles fit the staging buffer
    mult[:3] = (40, 3000, 40)
    ak = np.repeat(distinct, mult)
    rng.shuffle(ak)
    av = rng.permutation(len(ak)).astype(dt)                      # a payload names its build row
    order = np.argsort(av)
    bk = np.concatenate([distinct[rng.integers(3, len(distinct), 40_000)], rng.integers(1 << 30, 1 << 31, 3000).astype(dt),
                         distinct[:3], distinct[1:2].repeat(5)])
    rng.shuffle(bk[:43_000])                                      # the 3000-fold key stays at the end: one tile far beyond the staging buffer
    bv = np.arange(len(bk), dtype=dt) + 7
    want = oracle.sort_join(ak, av, bk, bv)
    with dwj.Engine(len(ak), key_bytes=dt().itemsize) as e:
        dak, dav, dbk, dbv = dev(ak), dev(av), dev(bk), dev(bv)
        e.build(dak, dav, len(ak))
        assert e.probe_count(dbk, len(bk)) == len(want[0])
        cap = len(want[0])
        ok, oa, ob = (empty_like_dev(cap, dt) for _ in range(3))
        assert e.probe_pairs(dbk, dbv, len(bk), ok, oa, ob, cap) == cap
        torch.cuda.synchronize()
        k, a, b = host(ok, dt)[:cap], host(oa, dt)[:cap], host(ob, dt)[:cap]
        assert (np.diff(b.astype(np.int64)) >= 0).all()           # probe-row order (ordered output is the default)
        np.testing.assert_array_equal(ak[order[a.astype(np.int64)]], k)      # every payload belongs to a build row with that key
        for w, x in zip(want, pyoracle.canonical_rows(k, a, b)):
            np.testing.assert_array_equal(w, x)
        # a second build into the same engine (runs and cursor are reused), fewer rows
        e.build(dak, dav, 5000)
        want2 = oracle.sort_join(ak[:5000], av[:5000], bk, bv)
        assert e.probe_count(dbk, len(bk)) == len(want2[0])
        m = e.probe_pairs(dbk, dbv, len(bk), None, oa, ob, cap)
        torch.cuda.synchronize()
        got2 = pyoracle.canonical_rows(bk[host(ob, dt)[:m].astype(np.int64) - 7], host(oa, dt)[:m], host(ob, dt)[:m])
        for w, x in zip(want2, got2):
            np.testing.assert_array_equal(w, x)
        # ALIGNED: one row per probe key, its payload one of the key's build rows
        n = len(bk)
        ok, oa, ob = (empty_like_dev(n, dt) for _ in range(3))
        e.build(dak, dav, len(ak))
        e.probe_aligned(dbk, dbv, n, ok, oa, ob)
        torch.cuda.synchronize()
        k, a = host(ok, dt), host(oa, dt)
        present = np.isin(bk, distinct)
        assert ((k != np.iinfo(dt).max) == present).all()
        np.testing.assert_array_equal(ak[order[a[present].astype(np.int64)]], bk[present])
        with pytest.raises(dwj.DwjError) as ei:                    # the payload runs are sized for max_build_rows
            e.build(dev(np.concatenate([ak, ak[:8]])), dev(np.concatenate([av, av[:8]])), len(ak) + 8)
        assert ei.value.code == -6


@pytest.mark.parametrize("lf", [0.25, 0.5, 0.9])
def test_u64_random_with_load_factors(dwj, oracle, lf):
    rng = np.random.default_rng(int(lf * 100))
    n = 200003                                                  # not a multiple of any tile
    ak = rng.integers(0, 2**64 - 2, n, dtype=np.uint64)
    av = rng.integers(0, 2**64 - 1, n, dtype=np.uint64)
    bk = np.concatenate([ak[rng.integers(0, n, 150000)], rng.integers(0, 2**64 - 2, 50001, dtype=np.uint64)])
    bv = rng.integers(0, 2**64 - 1, len(bk), dtype=np.uint64)
    got = pyoracle.canonical_rows(*gpu_join_pairs(dwj, ak, av, bk, bv, load_factor=lf))
    want = oracle.sort_join(ak, av, bk, bv)
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


# ---------------------------------------------------------------------------------------------------
# edge cases and error behaviour
# ---------------------------------------------------------------------------------------------------

def test_empty_and_ragged_inputs(dwj):
    z = np.zeros(0, np.uint32)
    one = np.array([5], np.uint32)
    k, a, b = gpu_join_pairs(dwj, z, z, one, one, capacity=4)          # empty build
    assert len(k) == 0
    k, a, b = gpu_join_pairs(dwj, one, one, z, z, capacity=4)          # empty probe
    assert len(k) == 0
    k, a, b = gpu_join_pairs(dwj, one, np.array([9], np.uint32), np.array([5, 6, 5], np.uint32),
                             np.array([1, 2, 3], np.uint32), capacity=4)
    assert (k.tolist(), a.tolist(), b.tolist()) == ([5, 5], [9, 9], [1, 3])
    # extreme key values; all-ones is the reserved empty marker (join.cpp:10 empty_element)
    ak = np.array([0, 0xFFFFFFFE, 0x80000000], np.uint32)
    k, a, b = gpu_join_pairs(dwj, ak, ak + 1, ak[::-1].copy(), ak, capacity=8)
    assert sorted(zip(k.tolist(), a.tolist())) == sorted(zip(ak.tolist(), (ak + 1).tolist()))


def test_unordered_output_same_multiset(dwj, oracle):
    """DWJ_FLAG_UNORDERED_OUTPUT trades the probe-order guarantee for one atomic per chunk: same rows, any order."""
    n = 300_001
    ak, av, bk, bv = (oracle.make_unique_random(n, s) for s in (21, 22, 23, 24))
    want = oracle.sort_join(ak, av, bk, bv)
    for wide in (False, True):
        dt = np.uint64 if wide else np.uint32
        cols = [c.astype(dt) for c in (ak, av, bk, bv)]
        with dwj.Engine(n, key_bytes=dt().itemsize, flags=dwj.FLAG_UNIQUE_BUILD_KEYS | dwj.FLAG_UNORDERED_OUTPUT) as e:
            dak, dav, dbk, dbv = (dev(c) for c in cols)
            e.build(dak, dav, n)
            ok, oa, ob = (empty_like_dev(n, dt) for _ in range(3))
            m = e.probe_pairs(dbk, dbv, n, ok, oa, ob, n)
            torch.cuda.synchronize()
            got = pyoracle.canonical_rows(*(host(t, dt)[:m] for t in (ok, oa, ob)))
            assert m == len(want[0])
            for w, x in zip(want, got):
                np.testing.assert_array_equal(w.astype(dt), x)


def test_key_column_optional(dwj, oracle):
    ak, av, bk, bv = (oracle.make_unique_random(5000, s) for s in (5, 6, 7, 8))
    k, a, b = gpu_join_pairs(dwj, ak, av, bk, bv, unique=True)
    k2, a2, b2 = gpu_join_pairs(dwj, ak, av, bk, bv, unique=True, with_key=False)
    assert k2 is None
    np.testing.assert_array_equal(a, a2)
    np.testing.assert_array_equal(b, b2)


def test_error_codes(dwj):
    with pytest.raises(dwj.DwjError) as ei:
        dwj.Engine(10, key_bytes=3)
    assert ei.value.code == -1
    with pytest.raises(dwj.DwjError) as ei:
        dwj.Engine(10, load_factor=0.99)
    assert ei.value.code == -1
    with pytest.raises(dwj.DwjError) as ei:
        dwj.Engine(10, device=99)
    assert ei.value.code == -1
    k = dev(np.arange(100, dtype=np.uint32))
    with dwj.Engine(16) as e:
        with pytest.raises(dwj.DwjError) as ei:                         # probe before build
            e.probe_count(k, 100)
        assert ei.value.code == -5
        with pytest.raises(dwj.DwjError) as ei:                         # more rows than the table holds
            e.build(k, k, 100)
        assert ei.value.code == -6
        e.build(k, k, 16)
        out = empty_like_dev(4, np.uint32)
        with pytest.raises(dwj.DwjError) as ei:                         # capacity too small: count still exact
            e.probe_pairs(k, k, 100, out, out, out, 4)
        assert ei.value.code == -4 and "16" in str(ei.value)
        assert e.info()["slots"] == 32 and e.info()["slots_per_bucket"] == 4


# ---------------------------------------------------------------------------------------------------
# host-buffer entry point (the reference's sycl::buffer-style call) and the partition kernels
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("wide,n_build,n_probe", [(False, 1 << 16, 1 << 18), (True, 50000, 70001),
                                                   (False, 1 << 20, (16 << 20) * 2 + 12345)])
def test_join_host_buffers(dwj, wide, n_build, n_probe):
    """FK->PK join through dwj_join_host with pageable numpy buffers; the largest case spans three probe
    chunks so the double-buffered copy/probe pipeline and the running output offset are exercised."""
    rng = np.random.default_rng(n_probe)
    dt = np.uint64 if wide else np.uint32
    ak = rng.permutation(n_build).astype(dt) * dt(2654435761 if not wide else 0x9E3779B97F4A7C15) + dt(1)
    assert len(np.unique(ak)) == n_build
    av = np.arange(n_build, dtype=dt)
    idx = rng.integers(0, n_build, n_probe)
    hit = rng.random(n_probe) < 0.75
    bk = np.where(hit, ak[idx], ak[idx] ^ dt(0x55555555))            # ~25% keys that are (almost surely) absent
    present = np.isin(bk, ak)
    bv = np.arange(n_probe, dtype=dt)
    lookup = dict()                                                 # expected payload by key, vectorised via sort
    order = np.argsort(ak)
    pos = np.searchsorted(ak[order], bk[present])
    want_a = av[order][pos]
    want_k, want_b = bk[present], bv[present]
    with dwj.Engine(n_build, key_bytes=dt().itemsize, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        ok, oa, ob = (np.empty(n_probe, dt) for _ in range(3))
        m, t = e.join_host(ak, av, n_build, bk, bv, n_probe, dwj.OUT_PAIRS, ok, oa, ob, n_probe)
        assert m == int(present.sum()) and t.total_ms > 0
        np.testing.assert_array_equal(ok[:m], want_k)               # probe order is preserved across chunks
        np.testing.assert_array_equal(oa[:m], want_a)
        np.testing.assert_array_equal(ob[:m], want_b)
        m2, _ = e.join_host(ak, av, n_build, bk, bv, n_probe, dwj.OUT_COUNT, None, None, None, 0)
        assert m2 == m
        n_al, _ = e.join_host(ak, av, n_build, bk, bv, n_probe, dwj.OUT_ALIGNED, ok, oa, ob, n_probe)
        assert n_al == n_probe
        sent = dt(~dt(0))
        np.testing.assert_array_equal(ok, np.where(present, bk, sent))
        np.testing.assert_array_equal(ob, np.where(present, bv, sent))
        np.testing.assert_array_equal(oa[present], want_a)
        assert (oa[~present] == sent).all()
        with pytest.raises(dwj.DwjError) as ei:
            e.join_host(ak, av, n_build, bk, bv, n_probe, dwj.OUT_PAIRS, ok, oa, ob, m - 1)
        assert ei.value.code == -4
    del lookup


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("parts", [1, 2, 8, 16, 64, 128, 256, 512])
def test_partition(dwj, wide, parts):
    from dwarf_bench_b200 import capi
    rng = np.random.default_rng(parts)
    dt = np.uint64 if wide else np.uint32
    n = 1_000_003
    k = rng.integers(0, 2**31, n).astype(dt)
    v = np.arange(n, dtype=dt)
    with dwj.Engine(16, key_bytes=dt().itemsize, hash_seed=42) as e:
        ok, ov = empty_like_dev(n, dt), empty_like_dev(n, dt)
        offs = torch.zeros(parts + 1, dtype=torch.int64, device="cuda")
        e.partition(dev(k), dev(v), n, parts, ok, ov, offs)
        counts = torch.zeros(parts, dtype=torch.int64, device="cuda")
        e.partition_hist(dev(k), n, parts, counts)
        torch.cuda.synchronize()
        offs = offs.cpu().numpy()
        ok, ov = host(ok, dt)[:n], host(ov, dt)[:n]
    np.testing.assert_array_equal(counts.cpu().numpy(), np.diff(offs))   # histogram-only entry point agrees
    assert offs[0] == 0 and offs[-1] == n and (np.diff(offs) >= 0).all()
    np.testing.assert_array_equal(k[ov.astype(np.int64)], ok)        # (key, payload) pairs stay together
    np.testing.assert_array_equal(np.sort(ov), v)                    # a permutation: nothing lost or duplicated
    sample = rng.integers(0, n, 2000)
    for i in sample:                                                 # every sampled row sits in its hash partition
        p = capi.partition_of(int(ok[i]), dt().itemsize, parts, 42)
        assert offs[p] <= i < offs[p + 1]
    if parts > 1:
        sizes = np.diff(offs)
        assert sizes.max() < 1.2 * n / parts + 1000                  # hash spreads uniformly


def test_table_larger_than_l2_int64(dwj):
    """32M unique 64-bit keys (1 GiB table at load 0.5 -- far beyond the 126 MB L2), 64M FK probes, exact check
    by gathering the expected payloads with torch (independent of the engine's kernels)."""
    R, S = 1 << 25, 1 << 26
    g = torch.Generator(device="cuda").manual_seed(13)
    ak = torch.randperm(R, device="cuda", generator=g, dtype=torch.int64) * 0x9E3779B1 + 12345
    av = torch.arange(R, device="cuda", dtype=torch.int64) * 3 + 1
    idx = torch.randint(0, R, (S,), device="cuda", generator=g)
    bk = ak[idx]
    bv = torch.arange(S, device="cuda", dtype=torch.int64)
    with dwj.Engine(R, key_bytes=8, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        e.build(ak, av, R)
        oa, ob = torch.empty_like(bk), torch.empty_like(bk)
        assert e.info()["radix_parts"] > 1                       # a table this size is probed region by region
        m = e.probe_pairs(bk, bv, S, None, oa, ob, S)
        assert m == S
        # rows come out region-major: check them through the probe payload (= probe row id)
        assert torch.equal(oa, av[idx[ob]])
        assert torch.equal(torch.sort(ob).values, bv)
        assert e.probe_count(bk, S) == S
    with dwj.Engine(R, key_bytes=8, flags=dwj.FLAG_UNIQUE_BUILD_KEYS | dwj.FLAG_NO_PARTITION) as e:
        assert e.info()["radix_parts"] == 1                      # opt-out keeps probe-row order
        e.build(ak, av, R)
        m = e.probe_pairs(bk, bv, S, None, oa, ob, S)
        assert m == S and torch.equal(oa, av[idx]) and torch.equal(ob, bv)


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("unique", [True, False])
def test_region_partitioned_path_small(dwj, oracle, monkeypatch, wide, unique):
    """Force the L2-region path at test size (1 MB regions) and compare with the oracle: same multiset for unique and
    duplicate build keys, same counts; region-major output order is allowed."""
    monkeypatch.setenv("DWJ_PARTITION_MIN_MB", "0")
    monkeypatch.setenv("DWJ_REGION_MB", "0.25")
    rng = np.random.default_rng(5 + wide + 2 * unique)
    dt = np.uint64 if wide else np.uint32
    n = 150_001
    if unique:
        ak = (rng.permutation(n).astype(np.uint64) * 2654435761 % (2**32 - 5)).astype(dt)
        ak = np.unique(ak)
    else:
        ak = rng.integers(0, 40_000, n).astype(dt)
    av = rng.integers(0, 2**31, len(ak)).astype(dt)
    bk = np.concatenate([ak[rng.integers(0, len(ak), 200_000)], rng.integers(2**31, 2**32 - 2, 5003).astype(dt)])
    bv = np.arange(len(bk), dtype=dt)
    flags = dwj.FLAG_UNIQUE_BUILD_KEYS if unique else 0
    with dwj.Engine(len(ak), key_bytes=dt().itemsize, flags=flags) as e:
        parts = e.info()["radix_parts"]
        assert parts >= 8
        dak, dav, dbk, dbv = dev(ak), dev(av), dev(bk), dev(bv)
        e.build(dak, dav, len(ak))
        cnt = e.probe_count(dbk, len(bk))
        ok, oa, ob = (empty_like_dev(cnt, dt) for _ in range(3))
        m = e.probe_pairs(dbk, dbv, len(bk), ok, oa, ob, cnt)
        torch.cuda.synchronize()
        got = pyoracle.canonical_rows(*(host(t, dt)[:m] for t in (ok, oa, ob)))
        # the reference-shaped entry points never partition (they are positional): still exact
        fl = torch.empty(len(bk), dtype=torch.int32, device="cuda")
        e.probe_contains(dbk, len(bk), fl)
        assert int(fl.sum().item()) == 200_000
    want = oracle.sort_join(ak, av, bk, bv)
    assert m == cnt == len(want[0])
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


def _unique_keys(rng, n, dt):
    space = 2**32 - 7 if dt == np.uint32 else 2**62
    k = np.unique(rng.integers(0, space, n + n // 8).astype(dt))[:n]
    rng.shuffle(k)
    return k


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("regions", [False, True])
def test_pass_filter_and_append(dwj, oracle, monkeypatch, wide, regions):
    """DWJ_OPT_PASS_FILTER + DWJ_OPT_APPEND_OUTPUT: the join run as 4 passes over key classes, the probe relation in three
    pieces per pass, every piece appending to ONE compact result through the device counter -- same multiset as the
    oracle's single join, with and without the engine's region partition."""
    if regions:
        monkeypatch.setenv("DWJ_PARTITION_MIN_MB", "0")
        monkeypatch.setenv("DWJ_REGION_MB", "0.125")
    from dwarf_bench_b200 import capi
    rng = np.random.default_rng(31 + wide)
    dt = np.uint64 if wide else np.uint32
    ak = _unique_keys(rng, 120_000, dt)
    av = rng.integers(0, 2**31, len(ak)).astype(dt)
    bk = np.concatenate([ak[rng.integers(0, len(ak), 260_000)], rng.integers(0, 2**31, 9_001).astype(dt)])
    bv = np.arange(len(bk), dtype=dt)
    want = oracle.sort_join(ak, av, bk, bv)
    P = 4
    with dwj.Engine(len(ak) // P * 2, key_bytes=dt().itemsize, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        assert (e.info()["radix_parts"] > 1) == regions
        dak, dav, dbk, dbv = dev(ak), dev(av), dev(bk), dev(bv)
        cap = len(bk)
        ok, oa, ob = (empty_like_dev(cap, dt) for _ in range(3))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        e.set_option(capi.OPT_APPEND_OUTPUT, 1)
        bounds = [0, 100_000, 100_001, len(bk)]
        for p in range(P):
            e.set_pass_filter(0, 2, p)
            e.build(dak, dav, len(ak))
            for c in range(3):
                lo, hi = bounds[c], bounds[c + 1]
                e.probe_pairs(dbk[lo:], dbv[lo:], hi - lo, ok, oa, ob, cap, d_n_matches=cnt, sync=False)
        e.set_pass_filter(0, 0, 0)
        e.set_option(capi.OPT_APPEND_OUTPUT, 0)
        torch.cuda.synchronize()
        m = int(cnt.item())
        got = pyoracle.canonical_rows(*(host(t, dt)[:m] for t in (ok, oa, ob)))
    assert m == len(want[0])
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


@pytest.mark.parametrize("wide", [False, True])
def test_segments_pull_and_region_scatter(dwj, oracle, monkeypatch, wide):
    """Pointer-based segment lists (the receiving end of the multi-GPU exchange, here all in local memory): the build and
    the probe read their rows from scattered pieces of several allocations; dwj_xpart_hist2 counts per (rank, region);
    dwj_region_scatter_segments gathers pieces into a region-grouped buffer that dwj_*_grouped consume."""
    monkeypatch.setenv("DWJ_PARTITION_MIN_MB", "0")
    monkeypatch.setenv("DWJ_REGION_MB", "0.125")
    from dwarf_bench_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(77 + wide)
    dt = np.uint64 if wide else np.uint32
    W = dt().itemsize
    ak = _unique_keys(rng, 90_000, dt)
    av = rng.integers(0, 2**31, len(ak)).astype(dt)
    bk = ak[rng.integers(0, len(ak), 200_003)]
    bv = np.arange(len(bk), dtype=dt)
    want = oracle.sort_join(ak, av, bk, bv)
    with dwj.Engine(len(ak), key_bytes=W, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        info = e.info()
        G = info["radix_parts"]
        assert G >= 8
        buckets = info["slots"] // info["slots_per_bucket"]
        region_bits = G.bit_length() - 1
        # (rank, region) histogram against the host functions
        dbk = dev(bk)
        counts = torch.zeros(4 * G, dtype=torch.int64, device="cuda")
        e.xpart_hist2(dbk, len(bk), 4, counts)
        sample = bk[:3000]
        exp = np.zeros(4 * G, dtype=np.int64)
        for k in sample:
            exp[lib.dwj_partition_of(int(k), W, 4, 42) * G + lib.dwj_region_of(int(k), W, buckets, region_bits, 42)] += 1
        c2 = torch.zeros(4 * G, dtype=torch.int64, device="cuda")
        e.xpart_hist2(dbk, len(sample), 4, c2)
        assert int(counts.sum().item()) == len(bk) and np.array_equal(c2.cpu().numpy(), exp)

        # pieces of the relations in separate allocations, odd sizes, one empty
        def pieces(k, v, cuts):
            out = []
            for lo, hi in zip(cuts, cuts[1:]):
                out.append((dev(k[lo:hi]) if hi > lo else None, dev(v[lo:hi]) if hi > lo else None, hi - lo))
            return out
        bp = pieces(ak, av, [0, 1, 30_011, 30_011, 70_000, len(ak)])
        pp = pieces(bk, bv, [0, 4097, 100_000, len(bk)])
        ptr = lambda t: 0 if t is None else t.data_ptr()          # noqa: E731

        # 1. region scatter of the pieces -> grouped build; probe pieces -> region scatter -> grouped probe (append)
        reg_of = lambda ks: np.array([lib.dwj_region_of(int(k), W, buckets, region_bits, 42) for k in ks], dtype=np.int64)   # noqa: E731
        rb = np.bincount(reg_of(ak), minlength=G)
        start = np.cumsum(rb) - rb
        lk, lv = empty_like_dev(len(ak), dt), empty_like_dev(len(ak), dt)
        e.region_scatter_segments([ptr(p[0]) for p in bp], [ptr(p[1]) for p in bp], [p[2] for p in bp], start, lk, lv)
        roff = torch.from_numpy(np.concatenate([start, [len(ak)]]).astype(np.int64)).cuda()
        e.build_grouped(lk, lv, len(ak), roff)
        torch.cuda.synchronize()
        got_reg = reg_of(host(lk, dt)[:len(ak)][::97])
        assert (np.diff(got_reg) >= 0).all()                        # region-grouped
        cap = len(bk)
        ok, oa, ob = (empty_like_dev(cap, dt) for _ in range(3))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        e.set_option(capi.OPT_APPEND_OUTPUT, 1)
        for k, v, n in pp:
            rp = np.bincount(reg_of(host(k, dt)), minlength=G)
            pk_, pv_ = empty_like_dev(n, dt), empty_like_dev(n, dt)
            e.region_scatter_segments([ptr(k)], [ptr(v)], [n], np.cumsum(rp) - rp, pk_, pv_)
            e.probe_pairs_grouped(pk_, pv_, n, ok, oa, ob, cap, d_n_matches=cnt, sync=False)
        torch.cuda.synchronize()
        m = int(cnt.item())
        got = pyoracle.canonical_rows(*(host(t, dt)[:m] for t in (ok, oa, ob)))
        assert m == len(want[0])
        for w, x in zip(want, got):
            np.testing.assert_array_equal(w, x)

        # 2. the kernels pull from the pieces directly
        e.set_option(capi.OPT_APPEND_OUTPUT, 0)
        e.build_segments([ptr(p[0]) for p in bp], [ptr(p[1]) for p in bp], [p[2] for p in bp])
        m2 = e.probe_pairs_segments([ptr(p[0]) for p in pp], [ptr(p[1]) for p in pp], [p[2] for p in pp], ok, oa, ob, cap)
        torch.cuda.synchronize()
        got = pyoracle.canonical_rows(*(host(t, dt)[:m2] for t in (ok, oa, ob)))
        assert m2 == len(want[0])
        for w, x in zip(want, got):
            np.testing.assert_array_equal(w, x)


@pytest.mark.parametrize("wide", [False, True])
def test_build_segments_with_duplicate_keys(dwj, oracle, monkeypatch, wide):
    """dwj_build_segments on an engine without the unique-keys flag: the three kernels of the one-to-many build read the
    build rows through the segment list (pieces of several allocations, one empty), table regions forced at test size."""
    monkeypatch.setenv("DWJ_PARTITION_MIN_MB", "0")
    monkeypatch.setenv("DWJ_REGION_MB", "0.125")
    rng = np.random.default_rng(91 + wide)
    dt = np.uint64 if wide else np.uint32
    distinct = _unique_keys(rng, 20_000, dt)
    ak = np.repeat(distinct, rng.integers(1, 6, len(distinct)))
    rng.shuffle(ak)
    av = np.arange(len(ak), dtype=dt)
    bk = np.concatenate([distinct[rng.integers(0, len(distinct), 50_000)], rng.integers(2**62, 2**63, 500).astype(dt) if wide
                         else rng.integers(2**31, 2**32 - 8, 500).astype(dt)])
    bv = np.arange(len(bk), dtype=dt)
    want = oracle.sort_join(ak, av, bk, bv)
    cuts = [0, 7, 20_001, 20_001, 45_000, len(ak)]
    parts = [(dev(ak[lo:hi]) if hi > lo else None, dev(av[lo:hi]) if hi > lo else None, hi - lo) for lo, hi in zip(cuts, cuts[1:])]
    ptr = lambda t: 0 if t is None else t.data_ptr()          # noqa: E731
    with dwj.Engine(len(ak), key_bytes=dt().itemsize) as e:
        e.build_segments([ptr(p[0]) for p in parts], [ptr(p[1]) for p in parts], [p[2] for p in parts])
        dbk, dbv = dev(bk), dev(bv)
        cap = len(want[0])
        assert e.probe_count(dbk, len(bk)) == cap
        ok, oa, ob = (empty_like_dev(cap, dt) for _ in range(3))
        assert e.probe_pairs(dbk, dbv, len(bk), ok, oa, ob, cap) == cap
        torch.cuda.synchronize()
        for w, x in zip(want, pyoracle.canonical_rows(*(host(t, dt)[:cap] for t in (ok, oa, ob)))):
            np.testing.assert_array_equal(w, x)


def test_hot_probe_keys_with_unique_build_keys(dwj, oracle):
    """Zipf(1.0) probe keys over UNIQUE build keys (the staged PAIRS kernel): same rows as the oracle, and the engine
    notices the skew on the device (a sample of the probe keys) and lets the table sectors into L1 -- while uniform
    probe keys keep the streaming loads."""
    rng = np.random.default_rng(321)
    n = 1 << 18
    ak = _unique_keys(rng, n, np.uint32)
    av = np.arange(n, dtype=np.uint32)
    bk_zipf = ak[zipf_ranks(rng, n, 1 << 20)]
    bk_unif = ak[rng.integers(0, n, 1 << 20)]
    bv = np.arange(1 << 20, dtype=np.uint32)
    with dwj.Engine(n, key_bytes=4, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
        dak, dav, dbv = dev(ak), dev(av), dev(bv)
        e.build(dak, dav, n)
        ok, oa, ob = (empty_like_dev(len(bv), np.uint32) for _ in range(3))
        for bk, hot in ((bk_zipf, 1), (bk_unif, 0), (bk_zipf, 1)):
            m = e.probe_pairs(dev(bk), dbv, len(bk), ok, oa, ob, len(bk))
            torch.cuda.synchronize()
            assert e.info()["hot_probe_keys"] == hot
            want = oracle.sort_join(ak, av, bk, bv)
            got = pyoracle.canonical_rows(*(host(t, np.uint32)[:m] for t in (ok, oa, ob)))
            assert m == len(want[0])
            for w, x in zip(want, got):
                np.testing.assert_array_equal(w, x)


def test_slab_golden_squares(dwj):
    """ref:tests/slab_tests.cpp:213-283 -- 1000 keys i*i inserted with value == key, then every key is found with its
    value.  SlabHash's insert / find are served by dwj_build / dwj_probe_* over the sector-bucket table."""
    keys = (np.arange(1000, dtype=np.uint64) ** 2).astype(np.uint32)
    keys = keys[keys != 0xFFFFFFFF]
    with dwj.Engine(len(keys), key_bytes=4, load_factor=0.625, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:      # slab_hash.hpp:57: 0.625 fill
        dk = dev(keys)
        e.build(dk, dk, len(keys))
        flags = torch.zeros(len(keys), dtype=torch.int32, device="cuda")
        e.probe_contains(dk, len(keys), flags)
        ok, oa, ob = (empty_like_dev(len(keys), np.uint32) for _ in range(3))
        e.probe_aligned(dk, dk, len(keys), ok, oa, ob)
        torch.cuda.synchronize()
        assert int(flags.sum().item()) == len(keys)
        np.testing.assert_array_equal(host(oa, np.uint32), keys)          # value == key for every key
        absent = dev(np.array([2, 3, 5, 999 * 999 + 1], dtype=np.uint32))
        f2 = torch.ones(4, dtype=torch.int32, device="cuda")
        e.probe_contains(absent, 4, f2)
        assert int(f2.sum().item()) == 0


def test_aligned_probe_under_duplicate_build_keys(dwj):
    """SimpleNonOwningHashTable::at returns ONE row per probe key even when the build side holds duplicates
    (ref:hashtable.hpp:23-40; ref:tests/hash_table_tests.cpp:112-113 pins "first inserted" for a serial insert).  With
    concurrent inserts the reference itself is racy about which duplicate comes first; the contract tested here is the
    one both share: exactly one row per probe key, and its payload is the payload of SOME build row with that key."""
    rng = np.random.default_rng(99)
    distinct = rng.choice(1 << 30, 5000, replace=False).astype(np.uint32)
    ak = np.repeat(distinct, 3)
    rng.shuffle(ak)
    av = np.arange(len(ak), dtype=np.uint32)
    bk = np.concatenate([distinct, rng.integers(1 << 30, (1 << 31), 1000).astype(np.uint32)])
    bv = np.arange(len(bk), dtype=np.uint32)
    with dwj.Engine(len(ak), key_bytes=4) as e:
        e.build(dev(ak), dev(av), len(ak))
        ok, oa, ob = (empty_like_dev(len(bk), np.uint32) for _ in range(3))
        e.probe_aligned(dev(bk), dev(bv), len(bk), ok, oa, ob)
        torch.cuda.synchronize()
    k, a, b = host(ok, np.uint32), host(oa, np.uint32), host(ob, np.uint32)
    hit = k != 0xFFFFFFFF
    assert hit[:5000].all() and not hit[5000:].any()
    np.testing.assert_array_equal(k[:5000], distinct)
    np.testing.assert_array_equal(ak[a[:5000]], distinct)                # the payload names a build row holding that key
    np.testing.assert_array_equal(b[:5000], bv[:5000])


def test_join_host_with_duplicate_build_keys(dwj, oracle):
    """dwj_join_host(PAIRS) with duplicate build keys: a piece of the probe relation may produce more rows than one
    staging slot holds (here 3 Mi probe rows x 4 matches = 12 Mi rows against 8 Mi per slot); the piece is probed again
    in smaller pieces and the result is the oracle's multiset."""
    rng = np.random.default_rng(8)
    distinct = rng.choice(1 << 30, 50_000, replace=False).astype(np.uint32)
    ak = np.repeat(distinct, 4)
    rng.shuffle(ak)
    av = np.arange(len(ak), dtype=np.uint32)
    bk = np.concatenate([distinct[rng.integers(0, len(distinct), 3 << 20)], rng.integers(1 << 30, 1 << 31, 1000).astype(np.uint32)])
    bv = np.arange(len(bk), dtype=np.uint32)
    want = oracle.sort_join(ak, av, bk, bv)
    cap = len(want[0])
    assert cap == 4 * (3 << 20)
    ok, oa, ob = (np.zeros(cap, dtype=np.uint32) for _ in range(3))
    with dwj.Engine(len(ak), key_bytes=4) as e:
        m, t = e.join_host(ak, av, len(ak), bk, bv, len(bk), dwj.OUT_PAIRS, ok, oa, ob, cap)
        assert m == cap
        m2, _ = e.join_host(ak, av, len(ak), bk, bv, len(bk), dwj.OUT_COUNT, None, None, None, 0)
        assert m2 == cap
    got = pyoracle.canonical_rows(ok, oa, ob)
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


@pytest.mark.parametrize("wide", [False, True])
def test_filter_rows_compaction(dwj, monkeypatch, wide):
    """dwj_filter_rows: the rows of one key class leave in one streaming pass, with their per-region counts; the four
    classes together are exactly the input rows, and every kept row belongs to the class asked for."""
    monkeypatch.setenv("DWJ_PARTITION_MIN_MB", "0")
    monkeypatch.setenv("DWJ_REGION_MB", "0.125")
    from dwarf_bench_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(17 + wide)
    dt = np.uint64 if wide else np.uint32
    W = dt().itemsize
    n = 300_007
    k = rng.integers(0, 2**31, n).astype(dt)
    v = np.arange(n, dtype=dt)
    with dwj.Engine(100_000, key_bytes=W) as e:
        info = e.info()
        G = info["radix_parts"]
        buckets = info["slots"] // info["slots_per_bucket"]
        dk, dv = dev(k), dev(v)
        ok, ov = empty_like_dev(n, dt), empty_like_dev(n, dt)
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        rc = torch.zeros(G, dtype=torch.int64, device="cuda")
        seen = np.zeros(n, dtype=bool)
        for p in range(4):
            e.set_pass_filter(0, 2, p)
            e.filter_rows(dk, dv, n, ok, ov, cnt, rc)
            torch.cuda.synchronize()
            m = int(cnt.item())
            kk, vv = host(ok, dt)[:m], host(ov, dt)[:m]
            assert int(rc.sum().item()) == m
            np.testing.assert_array_equal(k[vv.astype(np.int64)], kk)            # payloads still sit beside their keys
            assert not seen[vv.astype(np.int64)].any()
            seen[vv.astype(np.int64)] = True
            cls = [lib.dwj_partition_of(int(x), W, 4, 42) for x in kk[:300]]      # class = the two top bits of the partition hash
            assert set(cls) == {p}
            reg = np.bincount([lib.dwj_region_of(int(x), W, buckets, G.bit_length() - 1, 42) for x in kk[:2000]], minlength=G)
            assert (rc.cpu().numpy() >= reg).all()
        e.set_pass_filter(0, 0, 0)
        assert seen.all()


@pytest.mark.parametrize("wide", [False, True])
def test_groupby_sum(dwj, wide):
    """dwj_aggregate_sum (the GroupBy dwarf's kernel, ref:groupby/groupby.cpp:60-72): the reference's own 50-row vector
    (ref:tests/hash_table_tests.cpp:236-247), then 20 groups x 3 Mi rows (the library facade's shape, bench.cpp:80), then
    200 000 groups (beyond the per-CTA shared-memory tables), wrap-around sums, an absent group -- against the oracle."""
    from test_oracle_golden import GROUPBY_KEYS, GROUPBY_VALS
    dt = np.uint64 if wide else np.uint32
    rng = np.random.default_rng(4 + wide)
    cases = [(np.array(GROUPBY_KEYS, dtype=dt), np.array(GROUPBY_VALS, dtype=dt), 9),
             (rng.integers(0, 20, 3 << 20).astype(dt), rng.integers(1, 10001, 3 << 20).astype(dt), 20),
             (rng.integers(0, 200_000, 1 << 20).astype(dt), rng.integers(0, np.iinfo(dt).max, 1 << 20, dtype=dt), 200_001)]
    for keys, vals, groups in cases:
        want = pyoracle.groupby_sum(keys, vals, groups)
        present = np.zeros(groups, dtype=bool)
        present[keys.astype(np.int64)] = True
        with dwj.Engine(groups, key_bytes=dt().itemsize, flags=dwj.FLAG_UNIQUE_BUILD_KEYS) as e:
            e.aggregate_sum(dev(keys), dev(vals), len(keys))
            ids = np.arange(groups, dtype=dt)
            dids = dev(ids)
            ok, osum, oid = (empty_like_dev(groups, dt) for _ in range(3))
            e.probe_aligned(dids, dids, groups, ok, osum, oid)
            flags = torch.zeros(groups, dtype=torch.int32, device="cuda")
            e.probe_contains(dids, groups, flags)
            torch.cuda.synchronize()
        found = host(ok, dt) != np.iinfo(dt).max
        np.testing.assert_array_equal(found, present)
        np.testing.assert_array_equal(flags.cpu().numpy().astype(bool), present)
        np.testing.assert_array_equal(host(osum, dt)[present], want[present])
        assert (want[~present] == 0).all()
