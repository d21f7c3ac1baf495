#!/usr/bin/env python
"""Development sweep: device time of dwj_partition_hist and dwj_partition vs partition count (not a bench line)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dwarf_bench_b200 as dwj  # noqa: E402


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
    ap.add_argument("--rows", type=int, default=1 << 28)
    ap.add_argument("--key-bytes", type=int, default=4)
    ap.add_argument("--parts", type=int, nargs="+", default=[8, 16, 64, 128, 256, 512])
    args = ap.parse_args()
    n, W = args.rows, args.key_bytes
    tdt = torch.int32 if W == 4 else torch.int64
    g = torch.Generator(device="cuda").manual_seed(3)
    k = torch.randint(0, 2**31 - 1, (n,), device="cuda", generator=g, dtype=tdt)
    v = torch.arange(n, device="cuda", dtype=tdt)
    ok, ov = torch.empty_like(k), torch.empty_like(v)
    print(f"rows {n}, key bytes {W}, DWJ_PART_OLD={os.environ.get('DWJ_PART_OLD')}")
    print(f"{'parts':>6} | {'hist ms':>8} {'GB/s':>6} | {'hist+scatter ms':>15} {'scatter GB/s':>12}")
    with dwj.Engine(16, key_bytes=W) as e:
        for parts in args.parts:
            offs = torch.zeros(parts + 1, dtype=torch.int64, device="cuda")
            counts = torch.zeros(parts, dtype=torch.int64, device="cuda")
            th = timed(lambda: e.partition_hist(k, n, parts, counts))
            tp = timed(lambda: e.partition(k, v, n, parts, ok, ov, offs))
            assert int(offs[-1].item()) == n and int(counts.sum().item()) == n
            print(f"{parts:>6} | {th:>8.3f} {n * W / th / 1e6:>6.0f} | {tp:>15.3f} {n * W * 4 / max(tp - th, 1e-3) / 1e6:>12.0f}", flush=True)


if __name__ == "__main__":
    main()
