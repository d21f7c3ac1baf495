"""Worker of the multi-GPU parity tests (run as a subprocess / under torchrun by tests/test_gpu_multi.py).

    mg_worker.py virtual <world> <key_bytes> <direct|scatter> <passes> <out.json>
        one process, dwj_mg_* with `world` ranks that all live on cuda:0 (the full C++ orchestration -- streams, flags in
        "peer" memory, counts exchange, one host thread per rank -- on a one-GPU box)
    mg_worker.py ranks <key_bytes> <direct|scatter> <passes> <out_dir>         (under torchrun, one process per GPU)
        PullExchangeJoin over torch symmetric memory; every rank writes its result rows to <out_dir>/rank<r>_step<s>.npz

Several consecutive joins with fresh data per step catch slot-reuse races.  The comparison with the oracle happens in
the test process for `ranks`, here for `virtual`.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEPS = 3


def make_inputs(step, key_bytes, n_build, n_probe):
    rng = np.random.default_rng(100 + step)
    dt = np.uint32 if key_bytes == 4 else np.uint64
    space = 2**32 - 7 if key_bytes == 4 else 2**62
    ak = np.unique(rng.integers(0, space, n_build + n_build // 8).astype(dt))[:n_build]
    rng.shuffle(ak)
    av = rng.integers(0, 2**31, len(ak)).astype(dt)
    hit = ak[rng.integers(0, len(ak), n_probe - n_probe // 10)]
    miss = rng.integers(0, space, n_probe // 10).astype(dt)          # mostly absent keys
    bk = np.concatenate([hit, miss])
    rng.shuffle(bk)
    bv = np.arange(len(bk), dtype=dt) + dt(step * 1000)
    return ak, av, bk, bv


def main():
    mode = sys.argv[1]
    if os.environ.get("DWJ_TEST_WATCHDOG"):          # a hang becomes a Python traceback instead of a silent time-out
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DWJ_TEST_WATCHDOG"]), exit=True)
    import torch
    import dwarf_bench_b200 as dwj
    from oracle import pyoracle

    if mode == "virtual":
        world, key_bytes, pull, passes, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), sys.argv[6]
        dt = np.uint32 if key_bytes == 4 else np.uint64
        n_build, n_probe = 60_000, 151_003
        per_b, per_p = -(-n_build // world), -(-n_probe // world)
        o = pyoracle.Oracle()
        res = {"steps": []}
        dup = int(os.environ.get("DWJ_TEST_DUP", "1"))        # > 1: every build key `dup` times, engines without the unique-keys flag
        with dwj.MultiGpuJoin([0] * world, key_bytes, per_b * dup, per_p, chunk_rows=int(os.environ.get('DWJ_TEST_CHUNK_ROWS', 9_000)), passes=passes,
                              force_scatter_pull=(pull == "scatter"), recv_slack=1.5,
                              flags=dwj.FLAG_UNIQUE_BUILD_KEYS if dup == 1 else 0) as mg:
            info = mg.describe(0)
            res["info"] = info
            for step in range(STEPS):
                ak, av, bk, bv = make_inputs(step, key_bytes, n_build, n_probe)
                if dup > 1:
                    ak = np.repeat(ak, dup)
                    av = np.arange(len(ak), dtype=dt)[::-1].copy()
                cap = len(bk) * dup
                ok, ob, op = (np.zeros(cap, dtype=dt) for _ in range(3))
                m, t = mg.join_host(ak, av, bk, bv, ok, ob, op)
                want = o.sort_join(ak, av, bk, bv)
                got = pyoracle.canonical_rows(ok[:m], ob[:m], op[:m])
                same = m == len(want[0]) and all(np.array_equal(a, b) for a, b in zip(want, got))
                res["steps"].append({"rows": m, "want": int(len(want[0])), "same": bool(same), "timing": t})
        json.dump(res, open(out, "w"))
        return 0 if all(s["same"] for s in res["steps"]) else 1

    # ---- one process per GPU ---------------------------------------------------------------------------------------
    import torch.distributed as dist
    from dwarf_bench_b200.distributed import PullExchangeJoin
    key_bytes, pull, passes, out_dir = int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    dt = np.uint32 if key_bytes == 4 else np.uint64
    tdt = torch.int32 if key_bytes == 4 else torch.int64
    n_build, n_probe = 400_000, 1_000_003
    per_b, per_p = -(-n_build // world), -(-n_probe // world)
    eng = dwj.Engine(int(per_b / passes * 1.5) + 1024, key_bytes=key_bytes, device=local, load_factor=0.7,
                     flags=dwj.FLAG_UNIQUE_BUILD_KEYS)
    xj = PullExchangeJoin(eng, device, per_b, per_p, chunk_rows=60_000, passes=passes, recv_slack=1.5,
                          force_scatter_pull=(pull == "scatter"))
    cap = int(per_p * 1.5) + 1024
    out_k, out_b, out_p = (torch.empty(cap, dtype=tdt, device=device) for _ in range(3))
    d_count = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        json.dump(xj.info, open(os.path.join(out_dir, "info.json"), "w"))
    for step in range(STEPS):
        ak, av, bk, bv = make_inputs(step, key_bytes, n_build, n_probe)
        sl_b = slice(rank * per_b, min((rank + 1) * per_b, len(ak)))
        sl_p = slice(rank * per_p, min((rank + 1) * per_p, len(bk)))
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if key_bytes == 4 else np.int64)).to(device)  # noqa: E731
        dak, dav, dbk, dbv = dev(ak[sl_b]), dev(av[sl_b]), dev(bk[sl_p]), dev(bv[sl_p])
        xj.join(dak, dav, dak.numel(), dbk, dbv, dbk.numel(), out_k, out_b, out_p, cap, d_count)
        torch.cuda.synchronize()
        m = int(d_count.item())
        assert m <= cap
        np.savez(os.path.join(out_dir, f"rank{rank}_step{step}.npz"), k=out_k[:m].cpu().numpy().view(dt),
                 b=out_b[:m].cpu().numpy().view(dt), p=out_p[:m].cpu().numpy().view(dt))
        # the keys a rank joined all hash to that rank: equal keys met on one GPU
        from dwarf_bench_b200 import capi
        kk = out_k[:min(m, 64)].cpu().numpy().view(dt)
        assert all(capi.partition_of(int(k), key_bytes, world, 42) == rank for k in kk)
    t = xj.timings()
    print(f"[rank {rank}] last step timeline {t}", file=sys.stderr)
    xj.close()
    eng.close()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
