#!/usr/bin/env python
"""bench.py -- join tuples/sec (build + probe, device-timed) on N B200s, plus roofline / e2e / CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full pass of the hot path over one batch of synthetic input: both relations are resident in HBM when
the timed region starts; the step builds the table from the build relation, probes it with the probe relation and
materialises the compacted (build payload, probe payload) result rows.

Default workload at every N: BASELINE.json configs[4] -- the radix-partitioned join of 2^31 x 2^31 int64 rows (the
configuration the metric and the 8-GPU scaling target are quoted on), STRONG scaling: every rank holds 2^31 / N rows of
both relations in arrival order.  It runs through the C-ABI exchange join (dwj_xj_*, csrc/dwj_xj.cu): the senders group
their rows by destination in their own memory, the receivers' kernels pull them over NVLink.  On ONE GPU the working
set (69 GB of input, a 69 GB table, 34 GB of result rows) exceeds HBM, so the join runs as `passes` passes over key
classes (DWJ_OPT_PASS_FILTER) with a table for one class at a time -- stated in `config`.  With N == 1 the line also
carries a `configs` block: BASELINE configs 2, 3, 4 and the 256M x 256M target, each run through dwj_build +
dwj_probe_pairs with its own value / ms / roofline / traffic / parity.

Parity before any timing, at full size: the row count and an order-independent 128-bit checksum of all result rows
(summed over the ranks) against the generator's own expectation (dwarf_bench_b200/workloads.py).

One JSON line on stdout (rank 0).  `--impl reference` times the reference's own table code (oracle/_ref: the reference's
headers compiled unmodified; the oracle port when that was not built) on the host cores over a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import gc
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
NVLINK_PEAK_GBS = 770.0       # measured peer copy per direction per GPU (B200_PROFILING.md); nominal 900

WORKLOADS = {
    # name: (kind, build rows, probe rows, key bytes, scaling over N)
    "join_2Bx2B_u64": ("fk_pk", 1 << 31, 1 << 31, 8, "strong"),                   # BASELINE configs[4]  (default)
    "join_512Mx1G_u64_unique": ("fk_pk", 1 << 29, 1 << 30, 8, "weak"),            # configs[3]
    "join_16Mx256M_u32_unique": ("fk_pk", 1 << 24, 1 << 28, 4, "weak"),           # configs[1]
    "join_16Mx256M_u32_dup4_zipf": ("dup_zipf", 1 << 24, 1 << 28, 4, "weak"),     # configs[2]
    "join_256Mx256M_u32_unique": ("fk_pk", 1 << 28, 1 << 28, 4, "weak"),          # north_star target
    "join_1Mx1M_u32_reference_shape": ("reference", 1 << 20, 1 << 20, 4, "weak"),  # configs[0]
}
DEFAULT_WORKLOAD = "join_2Bx2B_u64"
SUB_CONFIGS = ["join_16Mx256M_u32_unique", "join_16Mx256M_u32_dup4_zipf", "join_512Mx1G_u64_unique", "join_256Mx256M_u32_unique"]


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
    ap.add_argument("--exchange", default="pull", choices=["pull", "nccl"],
                    help="multi-GPU exchange: pull = the C-ABI exchange join (kernels read their rows out of the senders' memory "
                         "over NVLink; default); nccl = local partition + NCCL all-to-all-v (baseline)")
    ap.add_argument("--chunk-rows", type=int, default=0, help="exchange join: rows per probe chunk (0 = 2^26)")
    ap.add_argument("--passes", type=int, default=0, help="exchange join: passes over key classes (0 = the fewest that fit in memory)")
    ap.add_argument("--scatter-pull", action="store_true", help="exchange join: force the region-scatter receive path")
    ap.add_argument("--engine-path", action="store_true", help="N == 1: run the workload through dwj_build + dwj_probe_pairs even if it is the exchange default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-configs", action="store_true", help="N == 1 default workload: skip the `configs` block")
    ap.add_argument("--sub-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-build-rows", type=int, default=1 << 27)
    ap.add_argument("--cpu-sample-probe-rows", type=int, default=1 << 28)
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
def cpu_join_sample(kind, n_build, n_probe, key_bytes, seed=7):
    """Host copies of a bounded sample of the workload (numpy; same key construction as workloads.py)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    dt = np.uint32 if key_bytes == 4 else np.uint64
    odd = 2654435761 if key_bytes == 4 else 0x9E3779B97F4A7C15
    if kind == "reference":
        from oracle import pyoracle
        o = pyoracle.Oracle()
        return tuple(o.make_unique_random(n, s) for n, s in ((n_build, 1), (n_build, 2), (n_probe, 3), (n_probe, 4)))
    if kind == "dup_zipf":
        distinct = n_build // 4
        keys = (np.arange(distinct, dtype=np.uint64) * np.uint64(odd) + np.uint64(12345)).astype(dt)
        ak = np.repeat(keys, 4)
        rng.shuffle(ak)
        w = 1.0 / np.arange(1, distinct + 1)
        cdf = np.cumsum(w) / w.sum()
        bk = keys[np.searchsorted(cdf, rng.random(n_probe)).clip(0, distinct - 1)]
    else:
        with np.errstate(over="ignore"):
            ak = (rng.permutation(n_build).astype(np.uint64) * np.uint64(odd) + np.uint64(12345)).astype(dt)
        bk = ak[rng.integers(0, n_build, n_probe)]
    return ak, np.arange(n_build, dtype=dt), bk, np.arange(n_probe, dtype=dt)


def cpu_baseline_runner():
    """(callable(ak,av,bk,bv) -> timing dict, kind, threads).  oracle/_ref when built, else the oracle port."""
    from oracle import pyoracle
    if pyoracle.Ref.available():
        r = pyoracle.Ref()
        return (lambda *a: r.join_build_probe(*a, seed=42)[1]), "reference", r.max_threads()
    o = pyoracle.Oracle()
    return (lambda *a: o.join_build_probe(*a, seed=42)[1]), "port", o.max_threads()


def cpu_sample_rows(args, n_build, n_probe, key_bytes):
    """The reference arm's bounded sample of a workload.  uint32 workloads up to BASELINE config 2's size run in full; the
    64-bit ones are cut down: the reference's table indexes its slots with uint32_t (hashtable.hpp:16, T = 2n < 2^32)
    and a 2^31-row relation with its table and outputs does not fit the host's memory."""
    if key_bytes == 4:
        return n_build, n_probe
    return min(n_build, args.cpu_sample_build_rows), min(n_probe, args.cpu_sample_probe_rows)


def cpu_sample_text(which, threads, key_bytes, n_build, n_probe, sb, sp):
    table = ("SimpleNonOwningHashTable<uint32_t, uint32_t, MurmurHash3_x86_32>" if key_bytes == 4 else
             "SimpleNonOwningHashTable<uint64_t, uint64_t, SimpleHasher<uint64_t>> (the reference's templates at 64 bits)")
    full = sb == n_build and sp == n_probe
    return (f"{'the full workload' if full else f'a sample of the workload: {sb} of {n_build} build rows, {sp} of {n_probe} probe rows'}; "
            f"{'the reference headers compiled unmodified (oracle/_ref)' if which == 'reference' else 'the oracle port (oracle/join_oracle.c)'}, "
            f"{table}, T = 2n slots, OpenMP over {threads} host threads, timed like join.cpp:59-113 (build + probe)")


def run_cpu_baseline(args, kind, n_build, n_probe, key_bytes, repeats=1):
    run, which, threads = cpu_baseline_runner()
    sb, sp = cpu_sample_rows(args, n_build, n_probe, key_bytes)
    if key_bytes == 8 and which != "reference":
        raise RuntimeError("the 64-bit CPU arm needs oracle/_ref (the reference's own templates)")
    ak, av, bk, bv = cpu_join_sample(kind, sb, sp, key_bytes)
    best = None
    for _ in range(repeats):
        t = run(ak, av, bk, bv)
        best = t if best is None or t["host_us"] < best["host_us"] else best
    tuples = len(ak) + len(bk)
    return {"value": tuples / (best["host_us"] * 1e-6), "unit": UNIT, "cores": threads, "kind": which,
            "sample": cpu_sample_text(which, threads, key_bytes, n_build, n_probe, sb, sp)
                      + f"; build {best['build_us'] / 1e3:.1f} ms, probe {best['probe_us'] / 1e3:.1f} ms",
            "host_ms": best["host_us"] / 1e3}


def main_reference(args, kind, n_build, n_probe, key_bytes, scaling, workload_name, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins every worker to OMP_NUM_THREADS=1; this arm is the reference's CPU path on ALL host threads and
        # only rank 0 runs it, so undo that before the OpenMP runtime of the checker library starts.
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    run, which, threads = cpu_baseline_runner()
    if key_bytes == 8 and which != "reference":
        out.write(json.dumps({"impl": "reference", "unavailable": "the 64-bit CPU arm needs oracle/_ref, which was not built"}) + "\n")
        out.flush()
        return 0
    sb, sp = cpu_sample_rows(args, n_build, n_probe, key_bytes)
    ak, av, bk, bv = cpu_join_sample(kind, sb, sp, key_bytes)
    for _ in range(args.warmup):
        run(ak, av, bk, bv)
    times = [run(ak, av, bk, bv)["host_us"] for _ in range(args.steps)]
    ms = sum(times) / len(times) / 1e3
    value = (len(ak) + len(bk)) / (ms * 1e-3)
    sample = "each step: " + cpu_sample_text(which, threads, key_bytes, n_build, n_probe, sb, sp)
    out.write(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "u32" if key_bytes == 4 else "u64", "data": "synthetic",
        "config": {"workload": workload_name, "build_rows": n_build, "probe_rows": n_probe, "sampled_build_rows": len(ak),
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


def traffic_of(workload, kernel):
    """dram__bytes_read+write per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
    except Exception:
        return None


def check_parity(torch, dist, world, device, inp, out_b, out_p, got, what):
    """Exact row count and the order-independent checksum of all result rows, summed over the ranks, against the
    generator's expectation.  Exits non-zero on any difference."""
    from dwarf_bench_b200 import workloads
    have = workloads.checksum_rows(out_b, out_p, got)
    vec = [got, inp.expected_matches, have[0], have[1], inp.expected_checksum[0], inp.expected_checksum[1]]
    if world > 1:
        mine = torch.tensor(vec, dtype=torch.int64, device=device)
        allv = torch.empty(world * len(vec), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allv, mine)
        rows = allv.cpu().view(world, len(vec)).tolist()
    else:
        rows = [vec]
    mask = (1 << 64) - 1
    tot = [sum(r[i] for r in rows) & mask for i in range(len(vec))]
    if tot[0] != tot[1]:
        sys.exit(f"{what}: join produced {tot[0]} rows, expected {tot[1]}")
    if (tot[2], tot[3]) != (tot[4], tot[5]):
        sys.exit(f"{what}: the result rows differ from the expected (build payload, probe payload) multiset "
                 f"(checksum {tot[2]:016x}{tot[3]:016x}, expected {tot[4]:016x}{tot[5]:016x})")
    return {"rows": tot[0], "checksum": f"{tot[2]:016x}{tot[3]:016x}",
            "method": "exact row count + order-independent 128-bit checksum of all (build payload, probe payload) result rows over all "
                      "ranks, against the generator's expectation (workloads.py)"}


def free_cuda(torch):
    gc.collect()
    torch.cuda.empty_cache()


def engine_run(args, wname, steps, warmup, device, local_rank, want_e2e, rows_override=None):
    """One workload on one GPU through dwj_build + dwj_probe_pairs.  Returns the fields of a bench line."""
    import torch
    import dwarf_bench_b200 as dwj
    kind, n_build, n_probe, key_bytes, _ = WORKLOADS[wname]
    if rows_override:
        n_build, n_probe = rows_override[0] or n_build, rows_override[1] or n_probe
    tdt = torch.int32 if key_bytes == 4 else torch.int64
    inp = make_input(kind, n_build, n_probe, key_bytes, device)
    free_cuda(torch)
    matches = inp.expected_matches
    flags = ((dwj.FLAG_UNIQUE_BUILD_KEYS if inp.unique_build else 0) | (dwj.FLAG_L2_PERSIST if args.l2_persist else 0)
             | (dwj.FLAG_NO_PARTITION if args.no_partition else 0) | (dwj.FLAG_UNORDERED_OUTPUT if args.unordered else 0))
    eng = dwj.Engine(n_build, key_bytes=key_bytes, device=local_rank, load_factor=args.load_factor, flags=flags)
    out_cap = matches
    out_key = torch.empty(out_cap, dtype=tdt, device=device) if args.emit_key else None
    out_b = torch.empty(out_cap, dtype=tdt, device=device)
    out_p = torch.empty(out_cap, dtype=tdt, device=device)
    d_count = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream()

    def step():
        eng.build(inp.build_keys, inp.build_vals, n_build, stream=stream)
        eng.probe_pairs(inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p, out_cap, d_n_matches=d_count,
                        sync=False, stream=stream)

    step()
    torch.cuda.synchronize()
    parity = check_parity(torch, None, 1, device, inp, out_b, out_p, int(d_count.item()), wname)
    info = eng.info()
    launches_per_step = info["launches_build"] + info["launches_probe"]
    for _ in range(max(warmup - 1, 0)):
        step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        for i in range(steps):
            ev[i][0].record(stream)
            eng.build(inp.build_keys, inp.build_vals, n_build, stream=stream)
            ev[i][1].record(stream)
            eng.probe_pairs(inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p, out_cap, d_n_matches=d_count,
                            sync=False, stream=stream)
            ev[i][2].record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
    ms_per_step = ev[0][0].elapsed_time(ev[-1][2]) / steps
    build_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    probe_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    res = {
        "value": (n_build + n_probe) / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "dtype": "u32" if key_bytes == 4 else "u64",
        "config": {"workload": wname, "build_rows_per_gpu": n_build, "probe_rows_per_gpu": n_probe, "matches_per_gpu": matches,
                   "key_bytes": key_bytes, "payload_bytes": key_bytes, "unique_build_keys": inp.unique_build,
                   "output": "compacted (build payload, probe payload" + (", key)" if args.emit_key else ")"),
                   "table_slots": info["slots"], "table_bytes": info["table_bytes"], "load_factor": args.load_factor,
                   "l2_persist_window": bool(info["l2_persist"]), "table_regions": info["radix_parts"],
                   "output_order": "probe-row order" if info["radix_parts"] == 1 and not args.unordered else "region-major / unordered",
                   "l2_between_iterations": "inputs and outputs (%.1f GB per step) far exceed the 126 MB L2; no explicit flush"
                                            % (((n_build + n_probe) * 2 + matches * 2) * key_bytes / 1e9),
                   "parallelism": "single GPU, dwj_build + dwj_probe_pairs"},
        "gpu_launches": launches_per_step * steps, "clocks": clocks.summary(), "wall_ms_per_step": wall_ms / steps,
        "parity": parity, "phases_ms": {"build": build_ms, "probe": probe_ms},
        "rates": {"build_tuples_per_s": n_build / (build_ms * 1e-3), "probe_tuples_per_s": n_probe / (probe_ms * 1e-3)},
    }
    # ---- roofline of the dominant kernel (the probe kernel proper) ------------------------------------------------------
    peak, peak_src = measured_peak_hbm()
    l2_res = info["table_bytes"] <= 100e6
    probe_bytes, step_bytes = algorithmic_bytes(n_build, n_probe, matches, key_bytes, info["slots"], args.emit_key, l2_res)
    tm = eng.timings()
    kernel_ms = tm.probe_kernel_ms if tm.probe_kernel_ms > 0 else probe_ms
    kname = "probe_pairs_staged_kernel" if inp.unique_build else "probe_pairs_multi_kernel"
    slot = 2 * key_bytes
    out_row = (2 + (1 if args.emit_key else 0)) * key_bytes
    if info["radix_parts"] > 1:
        # The kernel runs on region-partitioned input: it reads both probe columns, writes the result rows and pulls
        # every table region through L2 once.  (The partition pass is accounted for in `survey_model.probe_phase`.)
        kernel_bytes = n_probe * slot + matches * out_row + info["table_bytes"]
        kmodel = "S*(K+P) + M*out_row + T*slot (input pre-partitioned into L2-resident table regions)"
    else:
        kernel_bytes = probe_bytes
        kmodel = "SURVEY 8(d): S*(K+P) + M*out_row" + ("" if l2_res else " + S*32 (one sector per probe)")
    achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
    res["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                       "frac": achieved / peak, "traffic": traffic_of(wname, kname), "peak_source": peak_src,
                       "algorithmic_bytes_per_launch": kernel_bytes, "algorithmic_model": kmodel, "kernel_ms": kernel_ms,
                       "frac_of_nominal_8000": achieved / 8000.0,
                       # whichever algorithm runs, also report against the NON-partitioned sector-granular
                       # model of this table (32 B of table traffic per probe row) and against the compulsory bytes.
                       "survey_model": {
                           "probe_phase": {"ms": probe_ms, "includes": "region partition of the probe relation + probe kernel",
                                           "algorithmic_bytes": probe_bytes, "achieved": probe_bytes / (probe_ms * 1e-3) / 1e9,
                                           "frac": probe_bytes / (probe_ms * 1e-3) / 1e9 / peak},
                           "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                                          "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                          "frac_of_nominal_8000": step_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0},
                           "table_model": "L2-resident" if l2_res else "HBM-resident (sector-granular)"}}
    compulsory = (n_build + n_probe) * slot + matches * out_row + info["table_bytes"]
    res["roofline"]["compulsory_step"] = {"bytes": compulsory, "achieved": compulsory / (ms_per_step * 1e-3) / 1e9,
                                          "frac": compulsory / (ms_per_step * 1e-3) / 1e9 / peak,
                                          "model": "(R+S)*(K+P) + M*out_row + T*slot: every input row read once, every result row written once, the table written once"}
    # ---- end to end through the host-buffer entry point -------------------------------------------------------------------
    if want_e2e:
        def pinned(t):
            h = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        hb_k, hb_v, hp_k, hp_v = pinned(inp.build_keys), pinned(inp.build_vals), pinned(inp.probe_keys), pinned(inp.probe_vals)
        ho_k = torch.empty(matches, dtype=tdt, pin_memory=True) if args.emit_key else None
        ho_b, ho_p = torch.empty(matches, dtype=tdt, pin_memory=True), torch.empty(matches, dtype=tdt, pin_memory=True)
        e2e_steps = max(1, min(steps, 5))
        times = []
        for i in range(1 + e2e_steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m, tmh = eng.join_host(hb_k, hb_v, n_build, hp_k, hp_v, n_probe, dwj.OUT_PAIRS, ho_k, ho_b, ho_p, matches)
            dt_ms = (time.perf_counter() - t0) * 1e3
            if m != matches:
                sys.exit(f"e2e join produced {m} rows, expected {matches}")
            if i:
                times.append(max(dt_ms, tmh.total_ms))
        e2e_ms = sum(times) / len(times)
        h2d = (n_build + n_probe) * 2 * key_bytes
        d2h = matches * (2 + (1 if args.emit_key else 0)) * key_bytes
        res["e2e"] = {"value": (n_build + n_probe) / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                      "api": "dwj_join_host (pinned host columns in, compacted rows out; 3-stream chunked pipeline)",
                      "pcie_gbs": (h2d + d2h) / (e2e_ms * 1e-3) / 1e9}
        del hb_k, hb_v, hp_k, hp_v, ho_b, ho_p
    eng.close()
    del inp, out_b, out_p, out_key, eng
    free_cuda(torch)
    return res


def exchange_run(args, wname, steps, warmup, device, local_rank, world, rank, want_e2e, rows_override=None):
    """One workload through the C-ABI exchange join (dwj_xj_*), any N (N == 1: no peers, passes over key classes)."""
    import torch
    import torch.distributed as dist
    import dwarf_bench_b200 as dwj
    from dwarf_bench_b200.distributed import PullExchangeJoin
    kind, n_build_w, n_probe_w, key_bytes, scaling = WORKLOADS[wname]
    if rows_override:
        n_build_w, n_probe_w = rows_override[0] or n_build_w, rows_override[1] or n_probe_w
    if kind != "fk_pk":
        sys.exit("the exchange join bench supports the fk_pk workloads")
    # strong scaling: the workload's rows are dealt evenly to the ranks; weak: every rank holds the workload's rows
    n_build = n_build_w // world if scaling == "strong" else n_build_w
    n_probe = n_probe_w // world if scaling == "strong" else n_probe_w
    tdt = torch.int32 if key_bytes == 4 else torch.int64
    inp = make_input(kind, n_build, n_probe, key_bytes, device, seed=7 + rank, key_base=rank * n_build,
                     key_space=world * n_build, keep_map=False, probe_val_base=rank * n_probe)
    free_cuda(torch)
    matches = inp.expected_matches            # of this rank's probe rows; the rank's OUTPUT holds what it received
    flags = dwj.FLAG_UNIQUE_BUILD_KEYS
    slack = 1.03                              # hash-partitioned uniform keys: a rank receives its even share +- 0.01 %
    out_cap = n_probe if world == 1 else int(n_probe * slack) + 1024
    out_b = torch.empty(out_cap, dtype=tdt, device=device)
    out_p = torch.empty(out_cap, dtype=tdt, device=device)
    out_key = torch.empty(out_cap, dtype=tdt, device=device) if args.emit_key else None
    d_count = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream()

    def create(passes):
        # one table per key class: an even share plus head-room, at the load factor that keeps 2 slots per expected row
        rows = int(n_build / passes * slack) + 1024
        eng_ = dwj.Engine(rows, key_bytes=key_bytes, device=local_rank, load_factor=min(0.9, args.load_factor * slack * 1.02), flags=flags)
        try:
            xj_ = PullExchangeJoin(eng_, device, n_build, n_probe, chunk_rows=args.chunk_rows, passes=passes, recv_slack=slack,
                                   force_scatter_pull=args.scatter_pull, stream=stream)
        except Exception:
            eng_.close()
            raise
        return eng_, xj_

    eng = xj = None
    passes = args.passes or 1
    while True:                               # the fewest passes whose table, slots and landing buffers fit beside the data
        ok = 1
        try:
            eng, xj = create(passes)
        except (dwj.DwjError, torch.OutOfMemoryError, RuntimeError) as ex:
            if args.passes or passes >= 16 or not ("OOM" in str(ex) or "out of memory" in str(ex).lower()):
                raise
            ok = 0
            free_cuda(torch)
        if world > 1:
            t = torch.tensor([ok], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0 and ok:
                xj.close(); eng.close(); ok = 0
                free_cuda(torch)
        if ok:
            break
        passes *= 2
    info, xinfo = eng.info(), xj.info

    def step():
        xj.join(inp.build_keys, inp.build_vals, n_build, inp.probe_keys, inp.probe_vals, n_probe, out_key, out_b, out_p, out_cap, d_count)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    parity = check_parity(torch, dist, world, device, inp, out_b, out_p, int(d_count.item()), wname)
    for _ in range(max(warmup - 1, 0)):
        step()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        for i in range(steps):
            ev[i][0].record(stream)
            step()
            ev[i][1].record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    tl = xj.timings()                        # device timeline of the last timed step on this rank
    per_rank = [total_ms, tl["counts_ms"], tl["scattered_ms"], tl["built_ms"], tl["total_ms"], float(tl["remote_bytes"]),
                tl["build_pulled_ms"], tl["last_pulled_ms"]]
    if world > 1:
        t = torch.tensor(per_rank, device=device, dtype=torch.float64)
        allt = torch.empty(world * len(per_rank), device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allt, t)
        allt = allt.cpu().view(world, len(per_rank))
        total_ms = float(allt[:, 0].max())
        tl_max = [float(allt[:, i].max()) for i in range(1, 5)] + [float(allt[:, 6].max()), float(allt[:, 7].max())]
        remote = float(allt[:, 5].max())
    else:
        tl_max, remote = per_rank[1:5] + per_rank[6:8], per_rank[5]
    ms_per_step = total_ms / steps
    tuples = (n_build + n_probe) * world
    launches = xinfo["passes"] * (2 * (1 + xinfo["chunks"]) + 2            # histograms (+ memsets), publish, count wait
                                  + 3 * (1 + xinfo["chunks"])                # scatter + flags per batch
                                  + (info["launches_build"] + 2) + xinfo["chunks"] * (info["launches_probe"] + 2 + (0 if xinfo["direct_pull"] else 1)))
    res = {
        "value": tuples / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "dtype": "u32" if key_bytes == 4 else "u64",
        "scaling": scaling,
        "config": {"workload": wname, "build_rows_total": n_build * world, "probe_rows_total": n_probe * world,
                   "build_rows_per_gpu": n_build, "probe_rows_per_gpu": n_probe, "key_bytes": key_bytes, "payload_bytes": key_bytes,
                   "unique_build_keys": True, "output": "compacted (build payload, probe payload" + (", key)" if args.emit_key else ")")
                   + ", sharded by key hash, fully materialised on every GPU",
                   "table_slots_per_gpu": info["slots"], "table_bytes_per_gpu": info["table_bytes"], "table_regions": info["radix_parts"],
                   "passes_over_key_classes": xinfo["passes"],
                   "pass_structure": ("per pass and relation: compact the pass's key class out of the input (dwj_filter_rows, one streaming pass), "
                                      "histogram + 512-way region scatter of the compact copy, then build / probe; the two send slots serve as each "
                                      "other's scratch") if xinfo["compact_passes"] else None,
                   "one_gpu_note": (f"the 2^31 x 2^31 int64 working set (69 GB input + 69 GB table + 34 GB result) exceeds one GPU's HBM: "
                                    f"the join runs as {xinfo['passes']} passes over key classes, each re-reading both relations and "
                                    f"building a table for one class; all {tuples} input rows are joined and all result rows are resident at the end")
                   if world == 1 and xinfo["passes"] > 1 else None,
                   "probe_chunks": xinfo["chunks"], "chunk_rows": xinfo["chunk_rows"], "send_slots_in_rotation": xinfo["ring"],
                   "transfer": ("a copy kernel pulls every source's rows out of its slot over NVLink (blocks dealt round-robin over the "
                                "sources), " if xinfo["copy_pull"] else "the consuming kernels read the senders' slots themselves, ")
                               + ("build / probe walk them per (table region, source) segment" if xinfo["direct_pull"] else
                                  "a region scatter groups them by table region for build / probe"),
                   "exchange_block_bytes_per_gpu": xinfo["block_bytes"], "landing_bytes_per_gpu": xinfo["landing_bytes"],
                   "l2_between_iterations": "inputs and outputs (%.1f GB per GPU per step) far exceed the 126 MB L2; no explicit flush"
                                            % ((n_build + n_probe) * 2 * key_bytes * 2 / 1e9),
                   "parallelism": ("single GPU" if world == 1 else f"hash-partitioned x{world}") + ", dwj_xj_join: sender-side partition into "
                                  "peer-mapped slots, flags + counts in peer memory, receiver-side pull over NVLink, no collective on the data path"},
        "gpu_launches": launches * steps, "clocks": clocks.summary(), "wall_ms_per_step": wall_ms / steps, "parity": parity,
        "timeline_ms_last_step_max_over_ranks": {"counts_exchanged": tl_max[0], "last_batch_scattered": tl_max[1], "table_built": tl_max[2],
                                                 "last_probe_done": tl_max[3], "build_rows_pulled": tl_max[4], "last_chunk_pulled": tl_max[5],
                                                 "note": "from the step's start, first pass for the inner marks; the pulled marks are 0 on one GPU"},
    }
    if world > 1:
        window_ms = max(tl_max[3] - tl_max[0], 1e-6)
        res["exchange"] = {"remote_bytes_pulled_per_gpu_per_step": remote, "nvlink_peak_gbs": NVLINK_PEAK_GBS,
                           "nvlink_peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md; nominal 900)",
                           "min_exchange_ms_at_peak": remote / NVLINK_PEAK_GBS / 1e6,
                           "pull_window_ms": window_ms, "nvlink_gbs_in_window": remote / window_ms / 1e6,
                           "frac_of_peak": remote / window_ms / 1e6 / NVLINK_PEAK_GBS,
                           "window": "from the counts exchange to the end of the last probe on the slowest GPU (the pulls are fused into "
                                     "the build / probe / region-scatter kernels, so this is a lower bound on the link rate while pulling)"}
    # ---- roofline of the dominant kernel: the sender-side partition scatter (last launch = the last probe chunk) ----------
    peak, peak_src = measured_peak_hbm()
    tm = eng.timings()
    last_chunk_rows = n_probe - (xinfo["chunks"] - 1) * xinfo["chunk_rows"] if xinfo["chunks"] > 1 else n_probe
    kept = last_chunk_rows / xinfo["passes"]
    if xinfo["compact_passes"]:       # the scatter runs on the compact copy of the pass's key class: rows in, rows out
        scatter_bytes, scatter_model = kept * 4 * key_bytes, "kept rows * (K+P) read + kept rows * (K+P) written (the pass's key class was compacted first)"
    else:
        scatter_bytes = last_chunk_rows * key_bytes + kept * key_bytes + kept * 2 * key_bytes
        scatter_model = "chunk rows * K (every key is read) + kept rows * P (payloads of the pass's class) + kept rows * (K+P) written"
    kernels = {}
    if tm.partition_ms > 0:
        kernels["partition_scatter_many_kernel"] = {
            "ms": tm.partition_ms, "launches_per_step": xinfo["passes"] * (1 + xinfo["chunks"]),
            "algorithmic_bytes": scatter_bytes, "gbs": scatter_bytes / tm.partition_ms / 1e6,
            "model": scatter_model}
    if tm.probe_kernel_ms > 0:
        pk = kept * 2 * key_bytes + kept * 2 * key_bytes       # rows in (about one chunk's share lands here), result rows out
        kernels["probe_pairs_staged_kernel"] = {"ms": tm.probe_kernel_ms, "launches_per_step": xinfo["passes"] * xinfo["chunks"],
                                                "algorithmic_bytes": pk, "gbs": pk / tm.probe_kernel_ms / 1e6,
                                                "model": "received chunk rows * (K+P) + result rows * 2P (table regions stay in L2 across chunks)"}
    if tm.build_kernel_ms > 0:
        bb = n_build / xinfo["passes"] * 2 * key_bytes + info["table_bytes"] * 2
        kernels["build_kernel"] = {"ms": tm.build_kernel_ms, "launches_per_step": xinfo["passes"], "algorithmic_bytes": bb,
                                   "gbs": bb / tm.build_kernel_ms / 1e6,
                                   "model": "received build rows * (K+P) + the table read (look-ahead of the cleared slice) and written once"}
    res["kernels_last_launch"] = kernels
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["ms"] * kernels[k]["launches_per_step"])
        kd = kernels[dom]
        res["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kd["gbs"], "peak": peak, "unit": "GB/s", "frac": kd["gbs"] / peak,
                           "traffic": traffic_of(wname + f"@{world}", dom), "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": kd["algorithmic_bytes"], "algorithmic_model": kd["model"], "kernel_ms": kd["ms"],
                           "kernel_ms_source": "the engine's CUDA-event pair around the last launch of that kernel in the timed region "
                                               "(other streams of the join run beside it)",
                           "launches_per_step": kd["launches_per_step"], "share_of_step": kd["ms"] * kd["launches_per_step"] / ms_per_step,
                           "frac_of_nominal_8000": kd["gbs"] / 8000.0}
        slot = 2 * key_bytes
        compulsory = (n_build + n_probe) * slot + n_probe * slot + info["table_bytes"] * xinfo["passes"]
        res["roofline"]["compulsory_step_per_gpu"] = {
            "bytes": compulsory, "achieved": compulsory / (ms_per_step * 1e-3) / 1e9, "frac": compulsory / (ms_per_step * 1e-3) / 1e9 / peak,
            "model": "per GPU: (R+S)/N*(K+P) read once + M/N*2P written once + the table(s) written once"}
    # ---- end to end: this rank's inputs from pinned host memory, its result rows back -----------------------------------------
    if want_e2e:
        def pinned(t):
            h = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        hb_k, hb_v, hp_k, hp_v = pinned(inp.build_keys), pinned(inp.build_vals), pinned(inp.probe_keys), pinned(inp.probe_vals)
        ho_b, ho_p = torch.empty(out_cap, dtype=tdt, pin_memory=True), torch.empty(out_cap, dtype=tdt, pin_memory=True)
        torch.cuda.synchronize()
        times, d2h_rows = [], 0
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            inp.build_keys.copy_(hb_k, non_blocking=True); inp.build_vals.copy_(hb_v, non_blocking=True)
            inp.probe_keys.copy_(hp_k, non_blocking=True); inp.probe_vals.copy_(hp_v, non_blocking=True)
            step()
            m = int(d_count.item())
            ho_b[:m].copy_(out_b[:m], non_blocking=True); ho_p[:m].copy_(out_p[:m], non_blocking=True)
            barrier()
            d2h_rows = m
            if i:
                times.append((time.perf_counter() - t0) * 1e3)
        e2e_ms = sum(times) / len(times)
        vals = [e2e_ms, float(d2h_rows)]
        if world > 1:
            t = torch.tensor(vals, device=device, dtype=torch.float64)
            allt = torch.empty(world * 2, device=device, dtype=torch.float64)
            dist.all_gather_into_tensor(allt, t)
            allt = allt.cpu().view(world, 2)
            e2e_ms, d2h_total = float(allt[:, 0].max()), float(allt[:, 1].sum())
        else:
            d2h_total = float(d2h_rows)
        res["e2e"] = {"value": tuples / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": tuples * 2 * key_bytes,
                      "d2h_bytes_per_step": int(d2h_total) * 2 * key_bytes, "ms_per_step": e2e_ms, "steps": len(times),
                      "api": "dwj_xj_join behind pinned host columns: every rank copies its slice of both relations host->device, joins, "
                             "reads its row count and copies its result rows device->host (wall clock, max over ranks)",
                      "pcie_gbs_per_gpu": (tuples * 2 * key_bytes + d2h_total * 2 * key_bytes) / world / (e2e_ms * 1e-3) / 1e9}
        del hb_k, hb_v, hp_k, hp_v, ho_b, ho_p
    xj.close()
    eng.close()
    del inp, out_b, out_p, out_key
    free_cuda(torch)
    return res


def nccl_run(args, wname, steps, warmup, device, local_rank, world, rank):
    """The NCCL all-to-all-v baseline of the exchange (local partition, all_to_all_single, local join)."""
    import torch
    import torch.distributed as dist
    import dwarf_bench_b200 as dwj
    from dwarf_bench_b200.distributed import CudaJoinOps, ExchangeJoin
    kind, n_build_w, n_probe_w, key_bytes, scaling = WORKLOADS[wname]
    n_build = n_build_w // world if scaling == "strong" else n_build_w
    n_probe = n_probe_w // world if scaling == "strong" else n_probe_w
    tdt = torch.int32 if key_bytes == 4 else torch.int64
    inp = make_input(kind, n_build, n_probe, key_bytes, device, seed=7 + rank, key_base=rank * n_build, key_space=world * n_build,
                     keep_map=False, probe_val_base=rank * n_probe)
    eng = dwj.Engine(int(n_build * 1.25) + 1024, key_bytes=key_bytes, device=local_rank, load_factor=min(0.9, args.load_factor * 1.3),
                     flags=dwj.FLAG_UNIQUE_BUILD_KEYS)
    out_cap = int(n_probe * 1.25) + 1024
    out_b, out_p = torch.empty(out_cap, dtype=tdt, device=device), torch.empty(out_cap, dtype=tdt, device=device)
    d_count = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream()
    xj = ExchangeJoin(CudaJoinOps(eng, stream), device, tdt)

    def step():
        xj.join(inp.build_keys, inp.build_vals, n_build, inp.probe_keys, inp.probe_vals, n_probe, None, out_b, out_p, out_cap, d_count)

    step()
    dist.barrier(); torch.cuda.synchronize()
    parity = check_parity(torch, dist, world, device, inp, out_b, out_p, int(d_count.item()), wname)
    for _ in range(max(warmup - 1, 0)):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"value": (n_build + n_probe) * world / (ms * 1e-3), "ms_per_step": ms, "dtype": "u32" if key_bytes == 4 else "u64",
            "scaling": scaling, "parity": parity, "clocks": clocks.summary(), "gpu_launches": steps * 14,
            "config": {"workload": wname, "build_rows_per_gpu": n_build, "probe_rows_per_gpu": n_probe, "key_bytes": key_bytes,
                       "parallelism": f"hash-partitioned x{world}, local partition + NCCL all-to-all-v (baseline)"}}


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints to fd 1 (e.g. NCCL's version banner)
    # is sent to stderr; the JSON is written to the saved descriptor.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    kind, n_build, n_probe, key_bytes, scaling = WORKLOADS[args.workload]
    overridden = bool(args.build_rows or args.probe_rows)
    rows_override = (args.build_rows, args.probe_rows) if overridden else None
    if args.impl == "reference":
        return main_reference(args, kind, args.build_rows or n_build, args.probe_rows or n_probe, key_bytes, scaling, args.workload, real_stdout)

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
    dwj.load_library()

    use_exchange = world > 1 or (args.workload == DEFAULT_WORKLOAD and not args.engine_path)
    if world > 1 and args.exchange == "nccl":
        res = nccl_run(args, args.workload, args.steps, args.warmup, device, local_rank, world, rank)
    elif use_exchange:
        res = exchange_run(args, args.workload, args.steps, args.warmup, device, local_rank, world, rank, not args.no_e2e, rows_override)
    else:
        res = engine_run(args, args.workload, args.steps, args.warmup, device, local_rank, not args.no_e2e, rows_override)
        res["scaling"] = "weak"
    if overridden:
        res["config"]["workload"] += "[rows overridden -- development run]"
    line = {"metric": METRIC, "value": res.pop("value"), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True, "scaling": res.pop("scaling"), "vs_baseline": None,
            "dtype": res.pop("dtype"), "data": "synthetic"}
    line.update(res)

    # ---- the other BASELINE configs, in the same line (N == 1, default workload) ---------------------------------------------
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_sub_configs and not overridden:
        line["configs"] = {}
        for w in SUB_CONFIGS:
            sub = engine_run(args, w, args.sub_steps, 3, device, local_rank, not args.no_e2e and WORKLOADS[w][1] * WORKLOADS[w][3] <= (1 << 30))
            sub["unit"] = UNIT
            line["configs"][w] = sub

    # ---- CPU baseline beside it (rank 0, N == 1) ---------------------------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = run_cpu_baseline(args, kind, args.build_rows or n_build, args.probe_rows or n_probe, key_bytes)
        except Exception as ex:            # the checker libraries are optional at run time; say so rather than fail
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
