"""The CPU oracle (oracle/join_oracle.c) against the reference's golden vectors.

Fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the reference's own
headers (join_helpers.hpp, hashfunctions.hpp, hashtable.hpp) compiled unmodified -- see that script.
These tests need no GPU and no reference checkout.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle


def test_join_tests_golden_vector(oracle, golden_dir):
    """tests/join_tests.cpp:7-23 -- HelpersSeqJoin: 7x7 input, exactly 8 rows (dup key 5 gives 2x2)."""
    g = json.load(open(os.path.join(golden_dir, "join_tests_golden.json")))
    k, a, b = oracle.seq_join(g["keys_a"], g["vals_a"], g["keys_b"], g["vals_b"])
    assert len(k) == len(a) == len(b) == 8
    rows = [[int(x), int(y), int(z)] for x, y, z in zip(k, a, b)]
    assert rows == g["rows_emission_order"]          # same emission order as the reference loop
    ks, as_, bs = oracle.sort_join(g["keys_a"], g["vals_a"], g["keys_b"], g["vals_b"])
    assert sorted(map(tuple, rows)) == list(zip(ks.tolist(), as_.tolist(), bs.tolist()))


def test_helpers_equal_is_order_insensitive(oracle, golden_dir):
    """tests/join_tests.cpp:25-42 -- HelpersEqual, plus the property eq() is built on (sorting)."""
    g = json.load(open(os.path.join(golden_dir, "join_tests_golden.json")))
    t = oracle.seq_join(g["keys_a"], g["vals_a"], g["keys_b"], g["vals_b"])
    assert oracle.rows_equal(t, t)
    perm = np.random.default_rng(1).permutation(len(t[0]))
    assert oracle.rows_equal(t, tuple(c[perm] for c in t))
    bad = (t[0].copy(), t[1].copy(), t[2].copy())
    bad[2][0] ^= 1
    assert not oracle.rows_equal(t, bad)
    assert not oracle.rows_equal(t, tuple(c[:-1] for c in t))


def test_murmur_matches_reference(oracle, golden_dir):
    """hashfunctions.hpp:64-137 on edge values x seeds x table sizes."""
    g = np.load(os.path.join(golden_dir, "murmur_golden.npz"))
    for i, s in enumerate(g["seeds"]):
        for j, z in enumerate(g["sizes"]):
            got = [oracle.murmur_slot(int(v), int(s), int(z)) for v in g["vals"]]
            assert got == g["slots"][i, j].tolist()


@pytest.mark.parametrize("case", ["build", "probe", "has", "wrap", "murmur512", "dups256"])
def test_table_layout_matches_reference(oracle, golden_dir, case):
    """SimpleNonOwningHashTable: slot layout after collisions, first-duplicate-wins at(), wrap-around has()."""
    g = np.load(os.path.join(golden_dir, "table_golden.npz"))
    size, kind, seed = int(g[f"{case}.size"][0]), int(g[f"{case}.hash_kind"][0]), int(g[f"{case}.seed"][0])
    t = oracle.new_table(size, kind, seed)
    slots = [t.insert(int(k), int(v)) for k, v in zip(g[f"{case}.ins_k"], g[f"{case}.ins_v"])]
    assert slots == g[f"{case}.slots"].tolist()
    np.testing.assert_array_equal(t.keys, g[f"{case}.keys"])
    np.testing.assert_array_equal(t.vals, g[f"{case}.vals"])
    np.testing.assert_array_equal(t.bitmask, g[f"{case}.bitmask"])
    for q, f, v, h in zip(g[f"{case}.q"], g[f"{case}.found"], g[f"{case}.val"], g[f"{case}.has"]):
        val, hit = t.at(int(q))
        assert hit == bool(f) and t.has(int(q)) == bool(h)
        if hit:
            assert val == int(v)


def test_hash_table_tests_known_answers(oracle):
    """The literal assertions of tests/hash_table_tests.cpp (StaticSimpleHasher<64>)."""
    t = oracle.new_table(64, pyoracle.HASH_MODULO)
    for k, v in ((2, 2), (65, 3), (66, 8), (1, 9)):
        t.insert(k, v)
    t.insert(10, 1)
    t.insert(10, 2)
    r = t.vals
    assert (r[1], r[2], r[3], r[4]) == (3, 2, 8, 9)          # :50-53
    assert r[10] + r[11] == 3                                # :54
    t = oracle.new_table(64, pyoracle.HASH_MODULO)
    for k, v in ((1, 1), (1, 5), (4, 55)):
        t.insert(k, v)
    assert t.at(1) == (1, True) and t.at(4) == (55, True)    # :112-113 first inserted duplicate wins
    t = oracle.new_table(64, pyoracle.HASH_MODULO)
    for k, v in ((1, 1), (65, 5), (129, 6), (193, 7), (4, 55)):
        t.insert(k, v)
    assert [t.has(k) for k in (1, 65, 64, 4, 129, 193)] == [True, True, False, True, True, True]  # :175-180
    # BigBuild (:183-228): 500 distinct keys into a 500-slot table all survive.
    t = oracle.new_table(500, pyoracle.HASH_MODULO)
    for k in range(500):
        t.insert(k, k)
    assert len(set(t.vals.tolist())) == 500


@pytest.mark.parametrize("case", ["unique128", "unique1024", "unique4096", "dups", "empty_build", "no_match",
                                  "edge_keys", "u64"])
def test_joins_match_reference_seq_join(oracle, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "seq_join_golden.npz"))
    ak, av, bk, bv = (g[f"{case}.{n}"] for n in ("ak", "av", "bk", "bv"))
    want = tuple(g[f"{case}.{n}"] for n in ("k", "a", "b"))
    got_sort = oracle.sort_join(ak, av, bk, bv)
    for w, s in zip(want, got_sort):
        np.testing.assert_array_equal(w, s)                  # sort_join emits the canonical order
    if len(ak) * len(bk) <= 4096 * 4096:
        got_seq = pyoracle.canonical_rows(*oracle.seq_join(ak, av, bk, bv))
        for w, s in zip(want, got_seq):
            np.testing.assert_array_equal(w, s)


@pytest.mark.parametrize("n", [128, 1024, 4096])
def test_reference_join_shape_end_to_end(oracle, golden_dir, n):
    """Join::_run restated (join.cpp:30-131): table build+probe, host compaction, == seq_join."""
    g = np.load(os.path.join(golden_dir, "seq_join_golden.npz"))
    ak, av, bk, bv = (g[f"unique{n}.{x}"] for x in ("ak", "av", "bk", "bv"))
    (ok, op, ov), timing = oracle.join_build_probe(ak, av, bk, bv, seed=42)
    assert len(ok) == n and timing["threads"] >= 1
    miss = ok == 0xFFFFFFFF
    assert (op[miss] == 0xFFFFFFFF).all() and (ov[miss] == 0xFFFFFFFF).all()   # sentinel-filled, probe-aligned
    np.testing.assert_array_equal(ok[~miss], bk[~miss])
    got = pyoracle.canonical_rows(*oracle.compact(ok, op, ov))
    for w, s in zip((g[f"unique{n}.k"], g[f"unique{n}.a"], g[f"unique{n}.b"]), got):
        np.testing.assert_array_equal(w, s)


def test_hash_build_dwarf_check(oracle):
    """HashBuild::_run (hash_build.cpp:19-81): duplicates each take a slot, has() is 1 for every key."""
    src = oracle.make_random(4096, seed=3)
    assert src.min() >= 1 and src.max() <= 10000
    found, us, th = oracle.hash_build_check(src, seed=11)
    assert found == 4096 and us >= 0 and th >= 1


def test_generators_follow_common_cpp(oracle):
    """helpers::make_unique_random (common.cpp:7-20): n sorted unique values in [0, 10n)."""
    for n in (1, 128, 4096):
        v = oracle.make_unique_random(n, seed=5)
        assert len(v) == n and (np.diff(v.astype(np.int64)) > 0).all() and v.max() < 10 * n
    a, b = oracle.make_unique_random(65536, 1), oracle.make_unique_random(65536, 2)
    rate = len(np.intersect1d(a, b)) / 65536
    assert 0.08 < rate < 0.12                                 # match rate ~0.10 n


def test_omnisci_one_to_many(oracle):
    """JoinOmnisci (join_omnisci.cpp:15-45 are_equal): per probe row the SET of matching build rows."""
    rng = np.random.default_rng(9)
    ak = rng.integers(1, 200, 1000).astype(np.uint32)
    bk = rng.integers(1, 260, 700).astype(np.uint32)
    ids, off, cnt = oracle.omnisci_join(ak, bk)
    for j in range(len(bk)):
        expect = set(np.nonzero(ak == bk[j])[0].tolist())
        got = set(ids[int(off[j]):int(off[j]) + int(cnt[j])].tolist())
        assert got == expect and len(got) == int(cnt[j])


# ref:tests/hash_table_tests.cpp:236-247 (GroupByHashTable.GroupByFunctions): 50 rows, 9 groups
GROUPBY_KEYS = [0, 1, 2, 0, 3, 4, 0, 5, 0, 1, 8, 7, 2, 4, 5, 7, 1, 2, 4, 6, 2, 4, 1, 4, 6, 2, 4, 6, 8, 1, 8, 8, 8, 8, 8, 8, 8, 8, 8,
                1, 1, 2, 3, 4, 5, 6, 7, 8, 0, 0]
GROUPBY_VALS = [12, 19, 1, 4, 30, 21, 3, 8, 6, 19, 1, 1, 2, 0, 4, 4, 0, 5, 0, 1, 0, 1, 2, 0, 3, 4, 0, 5, 0, 1, 1, 3, 1, 3, 1, 0, 23, 11, 0,
                1, 33, 91, 12, 321, 12, 9, 99, 65, 7, 4]


def test_groupby_oracle_on_the_reference_vector():
    """The group sums of the reference's own GroupBy table test, computed as its loop does (:248-250), equal the oracle's."""
    from oracle import pyoracle
    answers = [0] * 9
    for k, v in zip(GROUPBY_KEYS, GROUPBY_VALS):
        answers[k] += v
    got = pyoracle.groupby_sum(np.array(GROUPBY_KEYS, dtype=np.uint32), np.array(GROUPBY_VALS, dtype=np.uint32), 9)
    assert got.tolist() == answers == [36, 75, 103, 42, 343, 24, 18, 104, 109]
    wrap = pyoracle.groupby_sum(np.array([0, 0, 1], dtype=np.uint32), np.array([0xFFFFFFFF, 2, 5], dtype=np.uint32), 2)
    assert wrap.tolist() == [1, 5]                     # uint32 wrap-around, as the reference's uint32_t sums
