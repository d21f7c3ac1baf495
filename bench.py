#!/usr/bin/env python
"""bench.py -- join tuples/sec (build + probe, device-timed) on N B200s, plus roofline / e2e / CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full pass of the hot path over one batch of synthetic input: dwj_build of the build relation
followed by dwj_probe_pairs of the probe relation with compacted (build payload, probe payload) output.

N == 1 (default): BASELINE.json configs[1] -- build 16 Mi / probe 256 Mi uint32 rows, unique build keys, every
probe row matches once.  N > 1 (under torchrun): every rank holds that same amount of both relations of an
N-times larger global join (weak scaling); rows are partitioned by (destination rank, table region) in one pass,
pushed into the peers' memory by the copy engines over NVLink while the SMs scatter and probe other chunks, and
joined locally without a second partition pass (dwarf_bench_b200/distributed.py: FoldedExchangeJoin; --exchange
p2p / nccl select the fused SM-store exchange or the NCCL all-to-all-v baseline).

One JSON line on stdout (rank 0).  `--impl reference` times the reference's own table code (oracle/_ref, or the
oracle port when that was not built) on the host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "join_tuples_per_sec"
UNIT = "tuples/s"

WORKLOADS = {
    # name: (kind, build rows, probe rows, key bytes)
    "join_16Mx256M_u32_unique": ("fk_pk", 1 << 24, 1 << 28, 4),          # BASELINE configs[1]  (default)
    "join_1Mx1M_u32_reference_shape": ("reference", 1 << 20, 1 << 20, 4),  # configs[0]
    "join_16Mx256M_u32_dup4_zipf": ("dup_zipf", 1 << 24, 1 << 28, 4),    # configs[2]
    "join_512Mx1G_u64_unique": ("fk_pk", 1 << 29, 1 << 30, 8),           # configs[3]
    "join_256Mx256M_u32_unique": ("fk_pk", 1 << 28, 1 << 28, 4),         # north_star target
}
DEFAULT_WORKLOAD = "join_16Mx256M_u32_unique"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--build-rows", type=int, default=0, help="override (development only; the line says so)")
    ap.add_argument("--probe-rows", type=int, default=0)
    ap.add_argument("--load-factor", type=float, default=0.5)
    ap.add_argument("--l2-persist", action="store_true", help="access-policy window on the table (measured slower; off)")
    ap.add_argument("--no-partition", action="store_true", help="probe the table directly (probe-row output order)")
    ap.add_argument("--unordered", action="store_true", help="DWJ_FLAG_UNORDERED_OUTPUT")
    ap.add_argument("--emit-key", action="store_true", help="also materialise the key column (reference row shape)")
    ap.add_argument("--exchange", default="fold", choices=["fold", "p2p", "nccl"],
                    help="multi-GPU exchange: fold = one (rank x table region) partition pass + copy-engine pushes into peer "
                         "memory, overlapped with the local probes (default); p2p = fused partition + SM stores into peer memory; "
                         "nccl = local partition + NCCL all-to-all-v")
    ap.add_argument("--exchange-chunks", type=int, default=0,
                    help="fold exchange: the probe relation travels in this many pieces (0 = 2 on two GPUs, 4 on more: measured)")
    ap.add_argument("--exchange-transport", default="ce1", choices=["sm", "ce", "ce1"],
                    help="fold exchange: runs pushed into peer memory by the copy engines -- ce1: one copy stream, peers in "
                         "rotated order (default; 4 GPUs: 4.9 ms/step against 7.0 ms with one stream per peer), ce: one stream per "
                         "peer -- or by a small kernel (sm)")
    ap.add_argument("--exchange-layout", default="blocked", choices=["blocked", "region"],
                    help="fold exchange: receive area source-major (one large transfer per peer, segmented build/probe) or region-major")
    ap.add_argument("--push-ctas", type=int, default=64, help="fold exchange, transport sm: CTAs of the push kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-probe-rows", type=int, default=1 << 26)
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread DURING the timed region (nvidia-smi -lms is too coarse for a ~100 ms region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's table code on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_join_sample(kind, n_build, n_probe_sample, seed=7):
    """Host copies of a bounded sample of the workload: the full build relation and the first rows of the probe."""
    import numpy as np
    rng = np.random.default_rng(seed)
    if kind == "reference":
        from oracle import pyoracle
        o = pyoracle.Oracle()
        return tuple(o.make_unique_random(n, s) for n, s in ((n_build, 1), (n_build, 2), (n_probe_sample, 3), (n_probe_sample, 4)))
    if kind == "dup_zipf":
        distinct = n_build // 4
        keys = (np.arange(distinct, dtype=np.uint64) * 2654435761 + 12345).astype(np.uint32)
        ak = np.repeat(keys, 4)
        rng.shuffle(ak)
        w = 1.0 / np.arange(1, distinct + 1)
        cdf = np.cumsum(w) / w.sum()
        bk = keys[np.searchsorted(cdf, rng.random(n_probe_sample)).clip(0, distinct - 1)]
    else:
        ak = (rng.permutation(n_build).astype(np.uint64) * 2654435761 + 12345).astype(np.uint32)
        bk = ak[rng.integers(0, n_build, n_probe_sample)]
    return ak, np.arange(n_build, dtype=np.uint32), bk, np.arange(n_probe_sample, dtype=np.uint32)


def cpu_baseline_runner():
    """(callable(ak,av,bk,bv) -> timing dict, kind, threads).  oracle/_ref when built, else the oracle port."""
    from oracle import pyoracle
    if pyoracle.Ref.available():
        r = pyoracle.Ref()
        return (lambda *a: r.join_build_probe(*a, seed=42)[1]), "reference", r.max_threads()
    o = pyoracle.Oracle()
    return (lambda *a: o.join_build_probe(*a, seed=42)[1]), "port", o.max_threads()


def run_cpu_baseline(kind, n_build, n_probe_sample, repeats=1):
    run, which, threads = cpu_baseline_runner()
    ak, av, bk, bv = cpu_join_sample(kind, n_build, n_probe_sample)
    best = None
    for _ in range(repeats):
        t = run(ak, av, bk, bv)
        best = t if best is None or t["host_us"] < best["host_us"] else best
    tuples = len(ak) + len(bk)
    return {"value": tuples / (best["host_us"] * 1e-6), "unit": UNIT, "cores": threads, "kind": which,
            "sample": f"full build relation ({len(ak)} rows) + first {len(bk)} probe rows; timed like join.cpp:59-113 "
                      f"(build+probe, host steady_clock); build {best['build_us'] / 1e3:.1f} ms, probe {best['probe_us'] / 1e3:.1f} ms",
            "host_ms": best["host_us"] / 1e3}


def main_reference(args, kind, n_build, n_probe, key_bytes, workload_name, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins every worker to OMP_NUM_THREADS=1; this arm is the reference's CPU path on ALL host threads and
        # only rank 0 runs it, so undo that before the OpenMP runtime of the checker library starts.
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    if key_bytes != 4:
        out.write(json.dumps({"impl": "reference", "unavailable": "the reference Join path is uint32-only (SURVEY fact 5)"}) + "\n")
        out.flush()
        return 0
    run, which, threads = cpu_baseline_runner()
    n_sample = min(n_probe, args.cpu_sample_probe_rows)
    ak, av, bk, bv = cpu_join_sample(kind, n_build, n_sample)
    for _ in range(args.warmup):
        run(ak, av, bk, bv)
    times = [run(ak, av, bk, bv)["host_us"] for _ in range(args.steps)]
    ms = sum(times) / len(times) / 1e3
    value = (len(ak) + len(bk)) / (ms * 1e-3)
    sample = (f"each step joins the full build relation ({len(ak)} rows) with the first {len(bk)} of {n_probe} probe rows "
              f"on {threads} host threads; timed as join.cpp:59-113")
    out.write(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": {"workload": workload_name, "build_rows": n_build, "probe_rows": n_probe,
                                                        "sampled_probe_rows": len(bk)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": which, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}) + "\n")
    out.flush()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(n_build, n_probe, matches, key_bytes, slots, emit_key, l2_resident):
    """SURVEY §8(d).  Returns (probe-kernel bytes, whole-step bytes)."""
    K = P = key_bytes
    slot = K + P
    stream_probe = n_probe * (K + P) + matches * (2 * P + (K if emit_key else 0))
    stream_build = n_build * (K + P)
    if l2_resident:
        table_build, table_probe = slots * slot, 0
    else:
        table_build, table_probe = slots * slot + n_build * 64, n_probe * 32
    return stream_probe + table_probe, stream_build + stream_probe + table_build + table_probe


def make_input(kind, n_build, n_probe, key_bytes, device, **kw):
    from dwarf_bench_b200 import workloads
    if kind == "fk_pk":
        return workloads.fk_pk(n_build, n_probe, key_bytes, device=device, **kw)
    if kind == "dup_zipf":
        return workloads.dup_zipf(n_build, n_probe, key_bytes=key_bytes, device=device)
    return workloads.reference_shape(n_build, device=device)


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints to fd 1 (e.g. NCCL's version banner)
    # is sent to stderr; the JSON is written to the saved descriptor.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    kind, n_build, n_probe, key_bytes = WORKLOADS[args.workload]
    overridden = bool(args.build_rows or args.probe_rows)
    n_build = args.build_rows or n_build
    n_probe = args.probe_rows or n_probe
    if args.impl == "reference":
        return main_reference(args, kind, n_build, n_probe, key_bytes, args.workload, real_stdout)

    import torch
    import torch.distributed as dist
    import dwarf_bench_b200 as dwj

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus} ...")
        sys.exit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    if not torch.cuda.is_available():
        sys.exit("no CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        if args.exchange_chunks <= 0:
            args.exchange_chunks = 2 if world <= 2 else 4
    dwj.load_library()
    tdt = torch.int32 if key_bytes == 4 else torch.int64
    dtype_name = "u32" if key_bytes == 4 else "u64"

    # ---- inputs (resident in HBM before the timed region) ---------------------------------------------------
    if world == 1:
        inp = make_input(kind, n_build, n_probe, key_bytes, device)
    else:
        if kind != "fk_pk":
            sys.exit("multi-GPU bench supports the fk_pk workloads")
        inp = make_input(kind, n_build, n_probe, key_bytes, device, seed=7 + rank, key_base=rank * n_build,
                         key_space=world * n_build, keep_map=False)
    matches = inp.expected_matches
    flags = ((dwj.FLAG_UNIQUE_BUILD_KEYS if inp.unique_build else 0) | (dwj.FLAG_L2_PERSIST if args.l2_persist else 0)
             | (dwj.FLAG_NO_PARTITION if args.no_partition else 0) | (dwj.FLAG_UNORDERED_OUTPUT if args.unordered else 0))
    cap_rows = n_build if world == 1 else int(n_build * 1.25) + 1024
    # Multi-GPU: a rank receives ~n_build rows +- hash imbalance; 1.25x head-room at load 0.65 keeps the same table
    # size (2 slots per expected row) as the single-GPU run.
    load_factor = args.load_factor if world == 1 else min(0.9, args.load_factor * 1.3)
    eng = dwj.Engine(cap_rows, key_bytes=key_bytes, device=local_rank, load_factor=load_factor, flags=flags)
    info = eng.info()
    out_cap = matches if world == 1 else int(n_probe * 1.25) + 1024
    out_key = torch.empty(out_cap, dtype=tdt, device=device) if args.emit_key else None
    out_b = torch.empty(out_cap, dtype=tdt, device=device)
    out_p = torch.empty(out_cap, dtype=tdt, device=device)
    d_count = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream()

    if world == 1:
        def step():
            eng.build(inp.build_keys, inp.build_vals, n_build, stream=stream)
            eng.probe_pairs(inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p, out_cap, d_n_matches=d_count,
                            sync=False, stream=stream)
        launches_per_step = None
    else:
        from dwarf_bench_b200.distributed import CudaJoinOps, ExchangeJoin, FoldedExchangeJoin, P2PExchangeJoin
        exchange_used = "nccl all-to-all-v"
        xj = None
        if args.exchange in ("fold", "p2p"):
            try:
                if args.exchange == "fold":
                    xj = FoldedExchangeJoin(eng, device, tdt, cap_rows, out_cap, n_build, n_probe, chunks=args.exchange_chunks,
                                            stream=stream, transport=args.exchange_transport, push_ctas=args.push_ctas,
                                            layout=args.exchange_layout)
                    how = (f"a {args.push_ctas}-CTA push kernel" if args.exchange_transport == "sm" else
                           "copy-engine pushes on one stream, peers in rotated order," if args.exchange_transport == "ce1" else "copy-engine pushes")
                    how += " (one block per peer, receiver walks the blocks region by region)" if args.exchange_layout == "blocked" else " (one run per peer and region)"
                    exchange_used = (f"one (rank x {xj.regions} table regions) partition pass, {how} into peer memory "
                                     f"(NVLink), probe relation in {xj.chunks} chunks overlapped with the local probes, counts by one all-gather")
                else:
                    xj = P2PExchangeJoin(eng, device, tdt, cap_rows, out_cap, stream=stream)
                    exchange_used = "fused partition + P2P stores into peer memory (NVLink), counts by all-gather"
            except Exception as ex:      # peer mapping unavailable on this box: NCCL path (still GPU-only)
                print(f"[rank {rank}] symmetric memory unavailable ({ex!r}); using NCCL all-to-all-v", file=sys.stderr)
                xj = None
            # all ranks must take the same path
            ok = torch.tensor([1 if xj is not None else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                xj = None
                exchange_used = "nccl all-to-all-v"
        if xj is None:
            xj = ExchangeJoin(CudaJoinOps(eng, stream), device, tdt)

        def step():
            xj.join(inp.build_keys, inp.build_vals, n_build, inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p,
                    out_cap, d_count)

    # ---- correctness guard before any timing (property checks; the oracle parity lives in tests/) -----------
    step()
    torch.cuda.synchronize()
    got = int(d_count.item())
    total = got
    if world > 1:
        t = torch.tensor([got], device=device, dtype=torch.int64)
        dist.all_reduce(t)
        total = int(t.item())
    want_total = matches * world
    if total != want_total:
        sys.exit(f"join produced {total} rows, expected {want_total}")
    if world == 1 and inp.probe_build_row is not None:
        # Order-independent exact check (the engine may emit rows region by region): the probe payload is the probe
        # row id, so every output row names the probe row it came from.
        rows = out_p[:got].long()
        if not (torch.equal(out_b[:got], inp.build_vals[inp.probe_build_row[rows]])
                and torch.equal(torch.sort(out_p[:got]).values, inp.probe_vals)):
            sys.exit("join output differs from the expected (build payload, probe payload) rows")
        del rows
    info = eng.info()
    launches_per_step = info["launches_build"] + info["launches_probe"] + (0 if world == 1 else 2 * 4)   # + 2 x dwj_partition
    if world > 1 and isinstance(xj, FoldedExchangeJoin):     # per batch: histogram + memset, offsets + scatter; one probe per chunk
        launches_per_step = (1 + xj.chunks) * 4 + info["launches_build"] + xj.chunks * info["launches_probe"]

    # ---- timed region ----------------------------------------------------------------------------------------
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup - 1, 0)):
        step()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        for i in range(args.steps):
            if world == 1:
                ev[i][0].record(stream)
                eng.build(inp.build_keys, inp.build_vals, n_build, stream=stream)
                ev[i][1].record(stream)
                eng.probe_pairs(inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p, out_cap, d_n_matches=d_count,
                                sync=False, stream=stream)
                ev[i][2].record(stream)
            else:
                ev[i][0].record(stream)
                step()
                ev[i][2].record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = (n_build + n_probe) * world / (ms_per_step * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name,
        "data": "synthetic",
        "config": {"workload": args.workload + ("" if not overridden else "[rows overridden -- development run]"),
                   "build_rows_per_gpu": n_build, "probe_rows_per_gpu": n_probe, "matches_per_gpu": matches,
                   "key_bytes": key_bytes, "payload_bytes": key_bytes, "unique_build_keys": inp.unique_build,
                   "output": "compacted (build payload, probe payload" + (", key)" if args.emit_key else ")"),
                   "table_slots": info["slots"], "table_bytes": info["table_bytes"], "load_factor": args.load_factor,
                   "l2_persist_window": bool(info["l2_persist"]), "table_regions": info["radix_parts"], "probe_passes": info["probe_passes"],
                   "output_order": "probe-row order" if info["radix_parts"] == 1 and not args.unordered else "region-major / unordered",
                   "l2_between_iterations": "inputs and outputs (%.1f GB per step) far exceed the 126 MB L2; no explicit flush"
                                            % (((n_build + n_probe) * 2 + matches * 2) * key_bytes / 1e9),
                   "parallelism": "single GPU" if world == 1 else f"hash-partitioned x{world}, {exchange_used}"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
        "wall_ms_per_step": wall_ms / args.steps,
    }

    # ---- roofline of the dominant kernel (probe), N == 1 -----------------------------------------------------
    if world == 1:
        build_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
        probe_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
        peak, peak_src = measured_peak_hbm()
        # whichever algorithm runs, report against the NON-partitioned sector-granular model of this table
        l2_res = info["table_bytes"] <= 100e6
        probe_bytes, step_bytes = algorithmic_bytes(n_build, n_probe, matches, key_bytes, info["slots"], args.emit_key, l2_res)
        # The dominant kernel is the probe kernel proper; its duration comes from the engine's own CUDA-event pair
        # around that launch (last timed step), the phase times from the events recorded above.
        tm = eng.timings()
        kernel_ms = tm.probe_kernel_ms if tm.probe_kernel_ms > 0 else probe_ms
        kname = "probe_pairs_staged_kernel" if inp.unique_build else "probe_pairs_multi_kernel"
        slot = 2 * key_bytes
        out_row = (2 + (1 if args.emit_key else 0)) * key_bytes
        if info["probe_passes"] > 1 and inp.unique_build:
            # Multi-pass region probe: the kernel sweeps the probe KEYS once per table slice, reads each payload once,
            # writes the result rows and pulls every slice through L2 once -- all of it inside this one launch.
            kernel_bytes = n_probe * key_bytes * info["probe_passes"] + n_probe * key_bytes + matches * out_row + info["table_bytes"]
            kmodel = "passes*S*K + S*P + M*out_row + T*slot (one sweep of the probe keys per L2-resident table slice)"
        elif info["radix_parts"] > 1:
            # The kernel runs on region-partitioned input: it must read both probe columns, write the result rows and
            # pull every table region through L2 once.  (The partition pass that makes this possible is accounted for
            # in `survey_model.probe_phase`.)
            kernel_bytes = n_probe * slot + matches * out_row + info["table_bytes"]
            kmodel = "S*(K+P) + M*out_row + T*slot (input pre-partitioned into L2-resident table regions)"
        else:
            kernel_bytes = probe_bytes
            kmodel = "SURVEY 8(d): S*(K+P) + M*out_row" + ("" if l2_res else " + S*32 (one sector per probe)")
        achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:        # dram__bytes_read+write of that kernel from the committed ncu --set full capture of this command
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload][kname]
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kernel_bytes, "algorithmic_model": kmodel, "kernel_ms": kernel_ms,
                            "frac_of_nominal_8000": achieved / 8000.0,
                            # whichever algorithm runs, also report against the NON-partitioned
                            # sector-granular model of this table (32 B of table traffic per probe row).
                            "survey_model": {
                                "probe_phase": {"ms": probe_ms, "includes": "region partition of the probe relation + probe kernel",
                                                "algorithmic_bytes": probe_bytes,
                                                "achieved": probe_bytes / (probe_ms * 1e-3) / 1e9,
                                                "frac": probe_bytes / (probe_ms * 1e-3) / 1e9 / peak},
                                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                                "table_model": "L2-resident" if l2_res else "HBM-resident (sector-granular)"}}
        line["phases_ms"] = {"build": build_ms, "probe": probe_ms}
        line["rates"] = {"build_tuples_per_s": n_build / (build_ms * 1e-3), "probe_tuples_per_s": n_probe / (probe_ms * 1e-3)}
    else:
        item = key_bytes * 2
        sent = (n_build + n_probe) * item * (world - 1) / world
        line["exchange"] = {"bytes_sent_per_gpu_per_step": sent, "nvlink_peak_gbs": 770.0,
                            "min_exchange_ms_at_peak": sent / 770e9 * 1e3}

    # ---- end to end through the host-buffer entry point -------------------------------------------------------
    if not args.no_e2e and world == 1 and inp.unique_build:
        hb_k, hb_v = inp.build_keys.cpu().pin_memory(), inp.build_vals.cpu().pin_memory()
        hp_k, hp_v = inp.probe_keys.cpu().pin_memory(), inp.probe_vals.cpu().pin_memory()
        ho_k = torch.empty(matches, dtype=tdt).pin_memory() if args.emit_key else None
        ho_b, ho_p = torch.empty(matches, dtype=tdt).pin_memory(), torch.empty(matches, dtype=tdt).pin_memory()
        e2e_steps = max(1, min(args.steps, 5))
        times = []
        for i in range(1 + e2e_steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m, tm = eng.join_host(hb_k, hb_v, n_build, hp_k, hp_v, n_probe, dwj.OUT_PAIRS, ho_k, ho_b, ho_p, matches)
            dt_ms = (time.perf_counter() - t0) * 1e3
            if m != matches:
                sys.exit(f"e2e join produced {m} rows, expected {matches}")
            if i:
                times.append(max(dt_ms, tm.total_ms))
        e2e_ms = sum(times) / len(times)
        h2d = (n_build + n_probe) * 2 * key_bytes
        d2h = matches * (2 + (1 if args.emit_key else 0)) * key_bytes
        line["e2e"] = {"value": (n_build + n_probe) / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                       "api": "dwj_join_host (pinned host columns in, compacted rows out)",
                       "pcie_gbs": (h2d + d2h) / (e2e_ms * 1e-3) / 1e9}
        del hb_k, hb_v, hp_k, hp_v, ho_b, ho_p
    elif world > 1 and not args.no_e2e:
        # Multi-GPU e2e: per-rank host staging around the same step (inputs H2D, local result D2H).
        hb_k, hb_v = inp.build_keys.cpu().pin_memory(), inp.build_vals.cpu().pin_memory()
        hp_k, hp_v = inp.probe_keys.cpu().pin_memory(), inp.probe_vals.cpu().pin_memory()
        ho_b, ho_p = torch.empty(out_cap, dtype=tdt).pin_memory(), torch.empty(out_cap, dtype=tdt).pin_memory()
        times = []
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            inp.build_keys.copy_(hb_k, non_blocking=True); inp.build_vals.copy_(hb_v, non_blocking=True)
            inp.probe_keys.copy_(hp_k, non_blocking=True); inp.probe_vals.copy_(hp_v, non_blocking=True)
            step()
            m = int(d_count.item())
            ho_b[:m].copy_(out_b[:m], non_blocking=True); ho_p[:m].copy_(out_p[:m], non_blocking=True)
            barrier()
            if i:
                times.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([sum(times) / len(times)], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        line["e2e"] = {"value": (n_build + n_probe) * world / (e2e_ms * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": (n_build + n_probe) * 2 * key_bytes * world,
                       "d2h_bytes_per_step": matches * 2 * key_bytes * world, "ms_per_step": e2e_ms,
                       "api": "ExchangeJoin.join with pinned host staging per rank"}

    # ---- CPU baseline beside it (rank 0, N == 1) ---------------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline and key_bytes == 4:
        try:
            line["cpu_baseline"] = run_cpu_baseline(kind, n_build, min(n_probe, args.cpu_sample_probe_rows))
        except Exception as ex:            # the checker libraries are optional at run time; say so rather than fail
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if world > 1 and getattr(xj, "trace", False):
        print(f"[rank {rank}] exchange timeline (ms from step start): {json.dumps(xj.last_trace)}", file=sys.stderr)
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
