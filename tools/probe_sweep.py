#!/usr/bin/env python
"""Development sweep: device time of each probe mode vs table size (which part of the probe costs what).
Not a bench line -- bench.py is the contract; this only prints a table for DESIGN.md / profiles/."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dwarf_bench_b200 as dwj  # noqa: E402
from dwarf_bench_b200 import workloads  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe-rows", type=int, default=1 << 27)
    ap.add_argument("--key-bytes", type=int, default=4)
    ap.add_argument("--build-log2", type=int, nargs="+", default=[20, 22, 23, 24, 26])
    ap.add_argument("--load-factor", type=float, default=0.5)
    ap.add_argument("--flags", type=int, default=dwj.FLAG_UNIQUE_BUILD_KEYS)
    ap.add_argument("--unordered", action="store_true")
    ap.add_argument("--dup", type=int, default=0, help="dup_zipf workload with this many duplicates per build key")
    args = ap.parse_args()
    S = args.probe_rows
    if args.unordered:
        args.flags |= dwj.FLAG_UNORDERED_OUTPUT
    tdt = torch.int32 if args.key_bytes == 4 else torch.int64
    print(f"probe rows {S}, key bytes {args.key_bytes}, load factor {args.load_factor}, flags {args.flags}")
    print("DWJ_PARTITION_MIN_MB =", os.environ.get("DWJ_PARTITION_MIN_MB"), " DWJ_REGION_MB =", os.environ.get("DWJ_REGION_MB"))
    print(f"{'build':>10} {'table MB':>9} | {'build ms':>9} {'Gins/s':>7} | {'count':>8} {'contains':>8} {'aligned':>8} {'pairs':>8} {'pairs+key':>9}  (ms; G probes/s in brackets)")
    for lg in args.build_log2:
        R = 1 << lg
        inp = workloads.dup_zipf(R, S, dup=args.dup, key_bytes=args.key_bytes) if args.dup else workloads.fk_pk(R, S, args.key_bytes, keep_map=False)
        M = inp.expected_matches
        e = dwj.Engine(R, key_bytes=args.key_bytes, load_factor=args.load_factor, flags=args.flags)
        ok, ob, op = (torch.empty(M, dtype=tdt, device="cuda") for _ in range(3))
        fl = torch.empty(S, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        tb = timed(lambda: e.build(inp.build_keys, inp.build_vals, R))
        res = [
            timed(lambda: e.probe_count(inp.probe_keys, S, d_n_matches=cnt, sync=False)),
            timed(lambda: e.probe_contains(inp.probe_keys, S, fl)),
            timed(lambda: e.probe_aligned(inp.probe_keys, inp.probe_vals, S, ok, ob, op)),
            timed(lambda: e.probe_pairs(inp.probe_keys, inp.probe_vals, S, None, ob, op, M, d_n_matches=cnt, sync=False)),
            timed(lambda: e.probe_pairs(inp.probe_keys, inp.probe_vals, S, ok, ob, op, M, d_n_matches=cnt, sync=False)),
        ]
        mb = e.info()["table_bytes"] / 2**20
        print(f"parts {e.info()['radix_parts']:>4} passes {e.info()['probe_passes']:>2}", end=" ")
        print(f"{R:>10} {mb:>9.0f} | {tb:>9.3f} {R / tb / 1e6:>7.2f} | " + " ".join(f"{t:>5.2f}[{S / t / 1e6:>4.0f}]" for t in res), flush=True)
        e.close()
        del inp, ok, ob, op, fl


if __name__ == "__main__":
    main()
