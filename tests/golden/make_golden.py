#!/usr/bin/env python
"""Generate the committed golden fixtures FROM THE REFERENCE ITSELF.

Run in the dev container (needs /root/reference):   python tests/golden/make_golden.py
It builds oracle/_ref/libref_join.so -- the reference's join_helpers.hpp, hashfunctions.hpp and
hashtable.hpp compiled unmodified -- and records its outputs on fixed inputs.  The fixtures are
what pins oracle/join_oracle.c (and, through it, the CUDA engine) on machines where the reference
checkout does not exist (the GPU box).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle  # noqa: E402

pyoracle.build(ref=True)
ref = pyoracle.Ref()
rng = np.random.default_rng(20261018)

# 1. tests/join_tests.cpp:10-19 -- the 7x7 vector, must give 8 rows.
ka, va = [1, 2, 3, 4, 5, 5, 7], [5, 1, 4, 6, 6, 5, 0]
kb, vb = [6, 2, 3, 4, 5, 5, 7], [3, 2, 1, 1, 3, 8, 8]
k, a, b = ref.seq_join(ka, va, kb, vb)
assert len(k) == 8
json.dump({"source": "tests/join_tests.cpp:10-19", "keys_a": ka, "vals_a": va, "keys_b": kb, "vals_b": vb,
           "rows_emission_order": [[int(x), int(y), int(z)] for x, y, z in zip(k, a, b)]},
          open(os.path.join(HERE, "join_tests_golden.json"), "w"), indent=1)

# 2. MurmurHash3_x86_32 (hashfunctions.hpp:64-137) on edge + random values.
vals = np.concatenate([np.array([0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF], np.uint64),
                       rng.integers(0, 2**32, 249, dtype=np.uint64)]).astype(np.uint32)
seeds = np.array([0, 1, 42, 999, 1000], np.uint32)
sizes = np.array([2, 64, 2048, 2097152, 4294967295], np.uint64)
table = np.zeros((len(seeds), len(sizes), len(vals)), np.uint64)
for i, s in enumerate(seeds):
    for j, z in enumerate(sizes):
        for l, v in enumerate(vals):
            table[i, j, l] = ref.murmur_slot(int(v), int(s), int(z))
np.savez_compressed(os.path.join(HERE, "murmur_golden.npz"), vals=vals, seeds=seeds, sizes=sizes, slots=table)

# 3. SimpleNonOwningHashTable layouts (tests/hash_table_tests.cpp scenarios + random fills).
cases = {}
def table_case(name, size, ins_k, ins_v, queries, hash_kind, seed=0):
    keys, vals_, bitmask, slots = ref.table_insert(size, ins_k, ins_v, hash_kind=hash_kind, seed=seed)
    found, val, has = ref.table_at(size, keys, vals_, bitmask, queries, hash_kind=hash_kind, seed=seed)
    for nm, arr in (("size", np.array([size], np.uint64)), ("hash_kind", np.array([hash_kind])),
                    ("seed", np.array([seed], np.uint32)), ("ins_k", np.asarray(ins_k, np.uint32)),
                    ("ins_v", np.asarray(ins_v, np.uint32)), ("q", np.asarray(queries, np.uint32)),
                    ("keys", keys), ("vals", vals_), ("bitmask", bitmask), ("slots", slots),
                    ("found", found), ("val", val), ("has", has)):
        cases[f"{name}.{nm}"] = arr
# hash_table_tests.cpp:34-41 (work-item 0's inserts, then key 10 twice -> slots 10 and 11)
table_case("build", 64, [2, 65, 66, 1, 10, 10], [2, 3, 8, 9, 1, 2], [1, 2, 10, 65, 66, 3], pyoracle.HASH_MODULO)
# hash_table_tests.cpp:92-99
table_case("probe", 64, [1, 1, 4], [1, 5, 55], [1, 4, 5], pyoracle.HASH_MODULO)
# hash_table_tests.cpp:148-160
table_case("has", 64, [1, 65, 129, 193, 4], [1, 5, 6, 7, 55], [1, 65, 64, 4, 129, 193], pyoracle.HASH_MODULO)
# a non-multiple-of-32 table that wraps, and murmur-hashed fills at load 0.5 (the Join shape)
table_case("wrap", 96, list(range(90, 96)) + [95, 95, 0], list(range(9)), [0, 95, 94, 7], pyoracle.HASH_MODULO)
ik = rng.choice(100000, 512, replace=False).astype(np.uint32)
table_case("murmur512", 1024, ik, ik * 3 + 1, np.concatenate([ik[:64], rng.integers(0, 100000, 64).astype(np.uint32)]),
           pyoracle.HASH_MURMUR, seed=42)
dk = rng.integers(1, 50, 256).astype(np.uint32)   # heavy duplicates (HashBuild's make_random shape)
table_case("dups256", 512, dk, np.arange(256), np.arange(0, 60), pyoracle.HASH_MURMUR, seed=7)
np.savez_compressed(os.path.join(HERE, "table_golden.npz"), **cases)

# 4. seq_join on seeded inputs: reference-shaped (sorted unique in [0,10n)) and duplicate-heavy.
out = {}
def join_case(name, ak, av, bk, bv):
    k, a, b = ref.seq_join(ak, av, bk, bv)
    k, a, b = pyoracle.canonical_rows(k, a, b)
    for nm, arr in (("ak", ak), ("av", av), ("bk", bk), ("bv", bv), ("k", k), ("a", a), ("b", b)):
        out[f"{name}.{nm}"] = np.asarray(arr)
for n in (128, 1024, 4096):
    gen = lambda: np.sort(rng.choice(10 * n, n, replace=False)).astype(np.uint32)
    join_case(f"unique{n}", gen(), gen(), gen(), gen())
n = 2048
join_case("dups", rng.integers(1, 300, n).astype(np.uint32), rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32),
          rng.integers(1, 300, n).astype(np.uint32), rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32))
join_case("empty_build", np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.arange(5, dtype=np.uint32), np.arange(5, dtype=np.uint32))
join_case("no_match", np.arange(0, 100, 2, dtype=np.uint32), np.arange(50, dtype=np.uint32),
          np.arange(1, 100, 2, dtype=np.uint32), np.arange(50, dtype=np.uint32))
join_case("edge_keys", np.array([0, 0, 0xFFFFFFFE, 7], np.uint32), np.array([1, 2, 3, 4], np.uint32),
          np.array([0xFFFFFFFE, 0, 9], np.uint32), np.array([10, 20, 30], np.uint32))
k64 = rng.integers(0, 2**63, 600, dtype=np.uint64)
join_case("u64", np.concatenate([k64[:400], k64[:50]]), rng.integers(0, 2**63, 450, dtype=np.uint64),
          np.concatenate([k64[300:], k64[300:350]]), rng.integers(0, 2**63, 350, dtype=np.uint64))
np.savez_compressed(os.path.join(HERE, "seq_join_golden.npz"), **out)
print("golden fixtures written to", HERE)
