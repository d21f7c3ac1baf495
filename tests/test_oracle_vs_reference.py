"""Live cross-check: oracle/join_oracle.c against the reference's own code (oracle/_ref).

Runs wherever oracle/_ref/libref_join.so exists (built in the dev container from /root/reference;
the .so travels to the GPU box).  Skipped otherwise -- the committed fixtures cover that case.
"""
import numpy as np
import pytest

from oracle import pyoracle

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("seed", range(4))
def test_seq_join_random_with_duplicates(oracle, ref, seed):
    rng = np.random.default_rng(seed)
    na, nb = rng.integers(1, 600, 2)
    ak = rng.integers(0, 200, na).astype(np.uint32)
    bk = rng.integers(0, 200, nb).astype(np.uint32)
    av = rng.integers(0, 2**32, na, dtype=np.uint64).astype(np.uint32)
    bv = rng.integers(0, 2**32, nb, dtype=np.uint64).astype(np.uint32)
    want = ref.seq_join(ak, av, bk, bv)
    got = oracle.seq_join(ak, av, bk, bv)
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)                       # identical emission order
    for w, g in zip(pyoracle.canonical_rows(*want), oracle.sort_join(ak, av, bk, bv)):
        np.testing.assert_array_equal(w, g)
    assert ref.rows_equal(want, oracle.sort_join(ak, av, bk, bv))  # the reference's own operator==
    assert ref.roundtrip_equal(got)                                # tests/join_tests.cpp:44-59


def test_seq_join_reference_shape_16k(oracle, ref):
    """the O(n log n) join proven equal to seq_join at n = 16 384."""
    n = 16384
    ak, av, bk, bv = (oracle.make_unique_random(n, s) for s in (1, 2, 3, 4))
    want = pyoracle.canonical_rows(*ref.seq_join(ak, av, bk, bv))
    got = oracle.sort_join(ak, av, bk, bv)
    assert 0.08 * n < len(got[0]) < 0.12 * n
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)


def test_murmur_random(oracle, ref):
    rng = np.random.default_rng(0)
    for v, s, z in zip(rng.integers(0, 2**32, 2000), rng.integers(0, 1001, 2000), rng.integers(1, 2**33, 2000)):
        assert oracle.murmur_slot(int(v), int(s), int(z)) == ref.murmur_slot(int(v), int(s), int(z))


@pytest.mark.parametrize("size,kind", [(64, pyoracle.HASH_MODULO), (100, pyoracle.HASH_MODULO),
                                        (4096, pyoracle.HASH_MURMUR), (1000, pyoracle.HASH_MURMUR)])
def test_table_sequential_layout(oracle, ref, size, kind):
    """Sequential inserts land in identical slots; at()/has() agree for hits and misses."""
    rng = np.random.default_rng(size)
    n = size // 2
    ik = rng.integers(0, 5 * size, n).astype(np.uint32)            # with duplicates
    iv = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    rk, rv, rb, rslots = ref.table_insert(size, ik, iv, hash_kind=kind, seed=17)
    t = oracle.new_table(size, kind, 17)
    slots = [t.insert(int(k), int(v)) for k, v in zip(ik, iv)]
    assert slots == rslots.tolist()
    np.testing.assert_array_equal(t.keys, rk)
    np.testing.assert_array_equal(t.vals, rv)
    np.testing.assert_array_equal(t.bitmask, rb)
    q = np.concatenate([ik[:50], rng.integers(0, 5 * size, 50).astype(np.uint32)])
    found, val, has = ref.table_at(size, rk, rv, rb, q, hash_kind=kind, seed=17)
    for qq, f, v, h in zip(q, found, val, has):
        gv, gh = t.at(int(qq))
        assert gh == bool(f) == bool(h) == t.has(int(qq))
        if gh:
            assert gv == int(v)


def test_join_build_probe_parallel(oracle, ref):
    """The timed region restated vs the reference's table code, both multi-threaded: same probe-aligned arrays
    (unique build keys, so the racy slot assignment cannot change any result)."""
    n = 1 << 16
    ak, av, bk, bv = (oracle.make_unique_random(n, s) for s in (11, 12, 13, 14))
    (rk, rp, rv), rt = ref.join_build_probe(ak, av, bk, bv, seed=42)
    (ok, op, ov), ot = oracle.join_build_probe(ak, av, bk, bv, seed=42)
    np.testing.assert_array_equal(rk, ok)
    np.testing.assert_array_equal(rp, op)
    np.testing.assert_array_equal(rv, ov)
    assert rt["host_us"] > 0 and ot["host_us"] > 0
    want = oracle.sort_join(ak, av, bk, bv)
    got = pyoracle.canonical_rows(*oracle.compact(rk, rp, rv))
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)
