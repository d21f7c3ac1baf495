"""Host-side logic of the multi-GPU join (dwarf_bench_b200/distributed.py) on CPU: world_size-2 gloo.

The device work is delegated to a JoinOps object; here a numpy stand-in DEFINED IN THIS TEST plays the device
(partition by the library's own host-side partition function, join by the oracle), so what is exercised is the
product's exchange plan: split sizes from partition offsets, the count exchange, the variable-size all-to-all,
buffer sizing, and that every key meets its partner on exactly one rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyJoinOps:
    """Test double for CudaJoinOps: same call signatures, numpy + oracle underneath."""

    def __init__(self, seed=42):
        from dwarf_bench_b200 import capi
        self.capi = capi
        self.seed = seed
        self.oracle = pyoracle.Oracle()
        self.built = None

    def partition(self, keys, vals, n, parts, out_keys, out_vals, offsets):
        k = keys[:n].numpy().view(np.uint32)
        v = vals[:n].numpy().view(np.uint32)
        pid = np.array([self.capi.partition_of(int(x), 4, parts, self.seed) for x in k], dtype=np.int64)
        order = np.argsort(pid, kind="stable")
        out_keys[:n] = torch.from_numpy(k[order].view(np.int32))
        out_vals[:n] = torch.from_numpy(v[order].view(np.int32))
        counts = np.bincount(pid, minlength=parts)
        offsets[:parts + 1] = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))

    def build(self, keys, vals, n):
        self.built = (keys[:n].numpy().view(np.uint32).copy(), vals[:n].numpy().view(np.uint32).copy())

    def probe_pairs(self, keys, vals, n, out_key, out_build, out_probe, capacity, d_count):
        k, a, b = self.oracle.sort_join(self.built[0], self.built[1], keys[:n].numpy().view(np.uint32), vals[:n].numpy().view(np.uint32))
        m = len(k)
        assert m <= capacity
        out_key[:m] = torch.from_numpy(k.view(np.int32))
        out_build[:m] = torch.from_numpy(a.view(np.int32))
        out_probe[:m] = torch.from_numpy(b.view(np.int32))
        d_count[0] = m


def _worker(rank, world, port, n_build, n_probe, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dwarf_bench_b200.distributed import ExchangeJoin
        rng = np.random.default_rng(100 + rank)
        # arrival-order slices of one global relation: unique build keys overall, probe keys drawn from all ranks' keys
        all_keys = np.random.default_rng(7).permutation(world * n_build * 3).astype(np.uint32)[: world * n_build]
        ak = all_keys[rank * n_build:(rank + 1) * n_build].copy()
        av = (np.arange(n_build, dtype=np.uint32) + rank * n_build)
        bk = np.concatenate([all_keys[rng.integers(0, world * n_build, n_probe - 50)], rng.integers(2**31, 2**32 - 2, 50).astype(np.uint32)])
        bv = (np.arange(n_probe, dtype=np.uint32) + rank * n_probe)
        t = lambda a: torch.from_numpy(a.view(np.int32).copy())
        xj = ExchangeJoin(NumpyJoinOps(), torch.device("cpu"), torch.int32)
        cap = 2 * n_probe
        ok, ob, op = (torch.empty(cap, dtype=torch.int32) for _ in range(3))
        cnt = torch.zeros(1, dtype=torch.int64)
        nb, np_ = xj.join(t(ak), t(av), n_build, t(bk), t(bv), n_probe, ok, ob, op, cap, cnt)
        m = int(cnt.item())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ak=ak, av=av, bk=bk, bv=bv, nb=nb, np_=np_,
                 k=ok[:m].numpy().view(np.uint32), a=ob[:m].numpy().view(np.uint32), b=op[:m].numpy().view(np.uint32),
                 sent=xj.stats.sent_rows, recv=xj.stats.recv_rows)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_exchange_join_world2(tmp_path, oracle, world):
    n_build, n_probe = 3000, 5000
    mp.spawn(_worker, args=(world, _free_port(), n_build, n_probe, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    cat = lambda name: np.concatenate([p[name] for p in parts])
    # the union of the per-rank results is the join of the union of the per-rank inputs
    want = oracle.sort_join(cat("ak"), cat("av"), cat("bk"), cat("bv"))
    got = pyoracle.canonical_rows(cat("k"), cat("a"), cat("b"))
    assert len(want[0]) == world * (n_probe - 50)
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)
    # every row went somewhere exactly once, and the hash spread is roughly even
    assert sum(int(p["nb"]) for p in parts) == world * n_build
    assert sum(int(p["np_"]) for p in parts) == world * n_probe
    assert sum(int(p["sent"]) for p in parts) == sum(int(p["recv"]) for p in parts) == world * (n_build + n_probe)
    for p in parts:
        assert 0.8 * n_build < int(p["nb"]) < 1.2 * n_build
    # a key and its matches live on one rank only
    from dwarf_bench_b200 import capi
    for r, p in enumerate(parts):
        assert all(capi.partition_of(int(k), 4, world, 42) == r for k in p["k"][:200])


def _sim_exchange(world, regions, fold, direct, n_rows, key_bytes, rng, lib):
    """The exchange of ONE batch, device memory played by numpy arrays, laid out by the exported plan functions
    (include/dwj.h: dwj_xj_plan_send / dwj_xj_plan_recv) exactly as dwj_xj_join uses them."""
    import ctypes as C
    u64p = C.POINTER(C.c_uint64)
    buckets = 1 << 12
    region_bits = regions.bit_length() - 1
    base = 64                                                   # the slot does not start at row 0
    dt = np.uint32 if key_bytes == 4 else np.uint64
    keys = [rng.integers(0, 2**31, n_rows[s]).astype(dt) for s in range(world)]
    dst = [np.array([lib.dwj_partition_of(int(k), key_bytes, world, 42) for k in ks], dtype=np.int64) for ks in keys]
    reg = [np.array([lib.dwj_region_of(int(k), key_bytes, buckets, region_bits, 42) for k in ks], dtype=np.int64) for ks in keys]
    counts = np.zeros((world, world, regions), dtype=np.uint64)               # [src][dst][region]
    for s in range(world):
        np.add.at(counts[s], (dst[s], reg[s]), 1)
    slots = []
    for s in range(world):                                      # sender: partition pass into its slot
        start = np.zeros(world * fold, dtype=np.uint64)
        mine = np.ascontiguousarray(counts[s])
        assert lib.dwj_xj_plan_send(world, regions, fold, base, mine.ctypes.data_as(u64p), start.ctypes.data_as(u64p)) == 0
        slot = np.full(base + n_rows[s], -1, dtype=np.int64)
        part = dst[s] * fold + (reg[s] if fold == regions else 0)
        cursor = start.astype(np.int64).copy()
        for i, p in enumerate(part):                            # any order inside a partition is fine
            slot[cursor[p]] = i
            cursor[p] += 1
        assert (slot[base:] >= 0).all() and len(set(slot[base:])) == n_rows[s]
        slots.append(slot)
    seen = [np.zeros(n_rows[s], dtype=bool) for s in range(world)]
    for me in range(world):                                     # receiver: segments into the senders' slots
        tot = np.ascontiguousarray(counts.sum(axis=2))          # [src][dst]
        rg = np.ascontiguousarray(counts[:, me, :])             # [src][region]
        pieces = 4
        cap = max(regions, pieces) * world
        first, rows, src, rstart, total, nseg = (np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint32),
                                                 np.zeros(regions, dtype=np.uint64), C.c_uint64(0), C.c_uint32(0))
        assert lib.dwj_xj_plan_recv(world, me, regions, base, tot.ctypes.data_as(u64p), rg.ctypes.data_as(u64p), 1 if direct else 0, pieces,
                                    first.ctypes.data_as(u64p), rows.ctypes.data_as(u64p), src.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    rstart.ctypes.data_as(u64p), C.byref(total), C.byref(nseg)) == 0
        n = nseg.value
        assert n == (regions * world if direct else world * max(1, min(pieces, int(tot[:, me].max()) >> 16)))
        assert total.value == tot[:, me].sum()
        assert (rstart == np.cumsum(rg.sum(axis=0)) - rg.sum(axis=0)).all()
        assert [int(x) for x in src[:world]] == [(me + i) % world for i in range(world)]      # rotated: no two ranks start at one source
        for i in range(n):
            s = int(src[i])
            idx = slots[s][int(first[i]):int(first[i] + rows[i])]
            assert (idx >= 0).all() and not seen[s][idx].any()
            seen[s][idx] = True
            assert (dst[s][idx] == me).all()                    # equal keys meet on one rank ...
            if direct:                                          # ... and arrive grouped by the receiver's table region
                assert (reg[s][idx] == i // world).all()
    for s in range(world):
        assert seen[s].all()                                    # nothing lost, nothing delivered twice


@pytest.mark.parametrize("world,regions", [(1, 4), (2, 8), (4, 1), (8, 16), (8, 64)])
@pytest.mark.parametrize("key_bytes", [4, 8])
def test_xj_plan_layout(world, regions, key_bytes):
    """Host logic of the multi-GPU pull exchange without a GPU: sender layout and receiver segment lists tile every slot
    exactly, every row reaches the rank (and region) its key hashes to, for the direct pull (regions folded into the
    sender's pass) and for the region-scatter pull (sender groups by destination only)."""
    from dwarf_bench_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(world * 100 + regions + key_bytes)
    n_rows = [int(x) for x in rng.integers(200, 900, world)]
    n_rows[-1] = 0 if world > 2 else n_rows[-1]                # a rank with nothing to send
    _sim_exchange(world, regions, regions, True, n_rows, key_bytes, rng, lib)
    _sim_exchange(world, regions, 1, False, n_rows, key_bytes, rng, lib)
    _sim_exchange(world, regions, regions, False, n_rows, key_bytes, rng, lib)      # non-unique engines: folded sender, scatter pull
