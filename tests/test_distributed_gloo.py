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


def test_plan_exchange_layout():
    """Host logic of the fused P2P exchange: source-major receive layout, no overlaps, nothing lost."""
    from dwarf_bench_b200.distributed import plan_exchange
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        counts = rng.integers(0, 1000, (world, world)).tolist()
        plans = [plan_exchange(counts, r) for r in range(world)]
        for d in range(world):
            # the ranges [offset_s, offset_s + counts[s][d]) written by every source s tile the receive buffer of d exactly
            spans = sorted((plans[s][0][d], plans[s][0][d] + counts[s][d]) for s in range(world))
            assert spans[0][0] == 0
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            assert spans[-1][1] == plans[d][1] == sum(counts[s][d] for s in range(world))


def test_plan_folded_exchange_layout():
    """Folded exchange: at every destination each relation's receive buffer is tiled exactly by the (batch, region,
    source) runs in that order; a batch is one contiguous segment; region offsets are the per-region totals; the
    sender's runs tile its send buffer chunk by chunk."""
    from dwarf_bench_b200.distributed import plan_folded_exchange
    rng = np.random.default_rng(9)
    for world, regions, chunks in ((2, 4, 1), (8, 8, 2), (4, 1, 3)):
        parts = world * regions
        counts = rng.integers(0, 300, (world, 1 + chunks, parts))
        # rows of every probe chunk at every source (the sender scatters chunk c in place of its input rows)
        bounds = [[0] + list(np.cumsum(counts[s, 1:].sum(axis=1))) for s in range(world)]      # chunks + 1 entries
        plans = [plan_folded_exchange(counts, r, regions, bounds[r]) for r in range(world)]
        for d in range(world):
            for relation in (range(0, 1), range(1, 1 + chunks)):
                spans = []
                for b in relation:
                    for g in range(regions):
                        for s in range(world):
                            p = d * regions + g
                            spans.append((int(plans[s]["dst_row"][b, p]), int(counts[s, b, p]), b, g, s))
                pos = 0
                for start, rows, b, g, s in spans:                    # already in (batch, region, source) order
                    assert start == pos, (world, regions, chunks, d, b, g, s)
                    pos += rows
                for b in relation:
                    first, rows = plans[d]["seg"][b]
                    mine = [sp for sp in spans if sp[2] == b]
                    assert first == mine[0][0] and rows == sum(sp[1] for sp in mine)
                    roff = plans[d]["region_off"][b]
                    assert roff[0] == 0 and roff[-1] == rows
                    for g in range(regions):
                        assert roff[g + 1] - roff[g] == sum(sp[1] for sp in mine if sp[3] == g)
        for s in range(world):
            src = plans[s]["src_row"]
            assert (src[0] == np.cumsum(counts[s, 0]) - counts[s, 0]).all()
            for c in range(chunks):
                assert src[1 + c, 0] == bounds[s][c]
                assert (np.diff(src[1 + c]) == counts[s, 1 + c, :-1]).all()


def test_plan_blocked_exchange_layout():
    """Blocked (source-major) folded exchange: one block per (source, destination, batch); the blocks and the rows a
    rank scatters in place tile every receive area exactly; the segment list walks a batch region by region, source by
    source, and names exactly the rows of that (source, region) run."""
    from dwarf_bench_b200.distributed import plan_blocked_exchange
    rng = np.random.default_rng(1)
    for w, G, C in ((2, 4, 1), (8, 8, 2), (4, 2, 3)):
        cnt = rng.integers(0, 50, (w, 1 + C, w * G))
        bounds = [[0] + list(np.cumsum(cnt[s, 1:].sum(axis=1))) for s in range(w)]
        plans = [plan_blocked_exchange(cnt, r, G, bounds[r]) for r in range(w)]
        tag = lambda s, b, p: (s * 100 + b) * 1000 + p                  # noqa: E731
        for d in range(w):
            for rel in (range(0, 1), range(1, 1 + C)):
                area = np.full(sum(plans[d]["seg"][b][1] for b in rel), -1)
                for s in range(w):
                    for b in rel:
                        if s == d:                                        # scattered in place, run by run
                            for g in range(G):
                                r0, n = plans[s]["own_row"][b, g], cnt[s, b, d * G + g]
                                assert (area[r0:r0 + n] == -1).all()
                                area[r0:r0 + n] = tag(s, b, d * G + g)
                        else:                                             # one transfer: the sender's (d, *) stretch
                            r0, n = plans[s]["block_dst"][b, d], plans[s]["block_rows"][b, d]
                            assert n == cnt[s, b, d * G:(d + 1) * G].sum()
                            assert plans[s]["block_src"][b, d] == plans[s]["src_row"][b, d * G]
                            ids = [np.full(cnt[s, b, d * G + g], tag(s, b, d * G + g)) for g in range(G)]
                            assert (area[r0:r0 + n] == -1).all()
                            area[r0:r0 + n] = np.concatenate(ids)
                assert (area != -1).all()
                for b in rel:
                    f, r = plans[d]["seg_first"][b], plans[d]["seg_rows"][b]
                    i = 0
                    for g in range(G):
                        for s in range(w):
                            assert r[i] == cnt[s, b, d * G + g] and (area[f[i]:f[i] + r[i]] == tag(s, b, d * G + g)).all()
                            i += 1
                    assert r.sum() == plans[d]["seg"][b][1]
