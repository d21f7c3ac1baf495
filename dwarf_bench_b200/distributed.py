"""Multi-GPU join: radix partition on the key hash -> exchange over NVLink -> local build + probe.

Three exchange implementations (profiles/r1_exchange.md):
  FoldedExchangeJoin  (default) ONE partition pass groups the rows by (destination rank, table region of the
                   destination's table); the copy engines push every rank's block into the peers' receive areas (peer
                   memory mapped through torch symmetric memory) on one stream in rotated peer order while the SMs
                   scatter and probe other chunks; the receiver builds / probes the blocks region by region through a
                   segment list, without a partition pass of its own.
  P2PExchangeJoin  the partition kernel stores every row straight into the destination rank's receive buffer: partition
                   and transfer are ONE kernel (SM stores over NVLink), then the normal local join.
  ExchangeJoin     partition locally, then NCCL all-to-all-v (the baseline, and the fallback when peer mapping is
                   unavailable).

No reference counterpart (the reference is single-device).  One process per GPU; the plumbing
is torch.distributed.  Every rank holds an arbitrary (arrival-order) slice of both relations; equal keys must meet on
one GPU, so there is one real exchange step per relation.  Only tiny count matrices travel through collectives.

The device work of ExchangeJoin is delegated to a `JoinOps` object so that the host-side logic (split sizes, buffer
sizing, ordering of collectives) can be exercised on CPU with gloo in tests, where a numpy stand-in supplied BY THE
TEST plays the device; the layout planners (plan_*) are pure numpy / Python and tested on CPU as well.  The product
only ever calls the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


class CudaJoinOps:
    """JoinOps over libdwj_b200.so on the current rank's GPU."""

    def __init__(self, engine, stream=None):
        self.e = engine
        self.stream = stream

    def partition(self, keys, vals, n, parts, out_keys, out_vals, offsets):
        self.e.partition(keys, vals, n, parts, out_keys, out_vals, offsets, stream=self.stream)

    def build(self, keys, vals, n):
        self.e.build(keys, vals, n, stream=self.stream)

    def probe_pairs(self, keys, vals, n, out_key, out_build, out_probe, capacity, d_count):
        self.e.probe_pairs(keys, vals, n, out_key, out_build, out_probe, capacity, d_n_matches=d_count, sync=False,
                           stream=self.stream)


@dataclass
class ExchangeStats:
    sent_rows: int = 0
    recv_rows: int = 0
    sent_bytes_remote: int = 0      # bytes that actually cross NVLink (everything not kept local)


class ExchangeJoin:
    """Hash-partitioned join across the ranks of a process group."""

    def __init__(self, ops, device, dtype=torch.int32, group=None, recv_slack: float = 1.25):
        self.ops = ops
        self.device = device
        self.dtype = dtype
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world & (self.world - 1):
            raise ValueError(f"world size must be a power of two, got {self.world}")
        self.recv_slack = recv_slack
        self._bufs = {}
        self.stats = ExchangeStats()

    # -- buffers ----------------------------------------------------------------------------------------------
    def _buf(self, name, n, dtype=None):
        dtype = dtype or self.dtype
        b = self._bufs.get(name)
        if b is None or b.numel() < n or b.dtype != dtype:
            b = torch.empty(max(int(n), 1), dtype=dtype, device=self.device)
            self._bufs[name] = b
        return b

    # -- one relation: partition + exchange ---------------------------------------------------------------------
    def partition_local(self, tag, keys, vals, n):
        """Group this rank's rows by destination.  Returns (keys, vals, offsets[world+1] on device)."""
        pk = self._buf(tag + ".pk", n)
        pv = self._buf(tag + ".pv", n)
        offs = self._buf(tag + ".offs", self.world + 1, torch.int64)
        self.ops.partition(keys, vals, n, self.world, pk, pv, offs)
        return pk, pv, offs

    def exchange_counts(self, offsets_list):
        """One collective for all relations: returns per relation (send_counts, recv_counts) as python lists."""
        k = len(offsets_list)
        send = torch.stack([(o[1:self.world + 1] - o[:self.world]) for o in offsets_list], dim=1).contiguous()  # [world, k]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        both = torch.stack([send, recv]).cpu()          # the one host sync of the step
        return [(both[0, :, i].tolist(), both[1, :, i].tolist()) for i in range(k)]

    def exchange_rows(self, tag, pk, pv, send_counts, recv_counts):
        n_recv = int(sum(recv_counts))
        cap = max(n_recv, int(self.recv_slack * sum(send_counts)) + 1)
        rk = self._buf(tag + ".rk", cap)
        rv = self._buf(tag + ".rv", cap)
        n_send = int(sum(send_counts))
        dist.all_to_all_single(rk[:n_recv], pk[:n_send], recv_counts, send_counts, group=self.group)
        dist.all_to_all_single(rv[:n_recv], pv[:n_send], recv_counts, send_counts, group=self.group)
        item = pk.element_size()
        self.stats.sent_rows += n_send
        self.stats.recv_rows += n_recv
        self.stats.sent_bytes_remote += 2 * item * (n_send - int(send_counts[self.rank]))
        return rk, rv, n_recv

    # -- the whole join -------------------------------------------------------------------------------------------
    def join(self, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe,
             capacity, d_count):
        """Local result (this rank's share of the global join) is written to out_*; the match count to d_count
        (device int64[1]).  Returns (n_build_local, n_probe_local) after the exchange."""
        bpk, bpv, boffs = self.partition_local("b", build_keys, build_vals, n_build)
        ppk, ppv, poffs = self.partition_local("p", probe_keys, probe_vals, n_probe)
        (bs, br), (ps, pr) = self.exchange_counts([boffs, poffs])
        rbk, rbv, nb = self.exchange_rows("b", bpk, bpv, bs, br)
        self.ops.build(rbk, rbv, nb)
        rpk, rpv, np_ = self.exchange_rows("p", ppk, ppv, ps, pr)
        self.ops.probe_pairs(rpk, rpv, np_, out_key, out_build, out_probe, capacity, d_count)
        return nb, np_


def plan_exchange(counts, rank: int):
    """counts[src][dst] = rows src sends to dst.  Receive layout on every rank: source-major (rows of rank 0, then
    rank 1, ...).  Returns (row offset of THIS rank's rows inside each destination's receive buffer, rows this rank
    receives)."""
    world = len(counts)
    offsets = [sum(int(counts[s][d]) for s in range(rank)) for d in range(world)]
    n_recv = sum(int(counts[s][rank]) for s in range(world))
    return offsets, n_recv


class P2PExchangeJoin:
    """Fused partition + exchange over peer memory (dwj_partition_scatter_to), then the local join."""

    def __init__(self, engine, device, dtype, cap_build: int, cap_probe: int, group=None, stream=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.e = engine
        self.device = device
        self.dtype = dtype
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world not in (1, 2, 4, 8):
            raise ValueError(f"P2P exchange supports 1, 2, 4 or 8 ranks, got {self.world}")
        self.stream = stream
        self.cap_build, self.cap_probe = int(cap_build), int(cap_probe)
        item = torch.empty(0, dtype=dtype).element_size()
        # one symmetric buffer per rank: [build keys | build payloads | probe keys | probe payloads]
        self.buf = symm_mem.empty(2 * (self.cap_build + self.cap_probe), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        bases = [int(p) for p in self.hdl.buffer_ptrs]
        off = [0, self.cap_build, 2 * self.cap_build, 2 * self.cap_build + self.cap_probe]
        self.dst = [[b + o * item for b in bases] for o in off]          # [column][rank] -> device pointer
        self.cols = [self.buf[o:o + n] for o, n in zip(off, (self.cap_build, self.cap_build, self.cap_probe, self.cap_probe))]
        self.counts = torch.zeros(2, self.world, dtype=torch.int64, device=device)
        self.all_counts = torch.zeros(self.world, 2, self.world, dtype=torch.int64, device=device)
        self.stats = ExchangeStats()

    def join(self, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe,
             capacity, d_count):
        e, w = self.e, self.world
        e.partition_hist(build_keys, n_build, w, self.counts[0], stream=self.stream)
        e.partition_hist(probe_keys, n_probe, w, self.counts[1], stream=self.stream)
        # The all-gather is also the point after which every rank has finished its previous local join (it is ordered
        # behind that join on every rank's stream), so the receive buffers may be overwritten.
        dist.all_gather_into_tensor(self.all_counts.view(-1), self.counts.view(-1), group=self.group)
        m = self.all_counts.cpu().tolist()                              # the one host sync of the step
        boff, nb = plan_exchange([[m[s][0][d] for d in range(w)] for s in range(w)], self.rank)
        poff, np_ = plan_exchange([[m[s][1][d] for d in range(w)] for s in range(w)], self.rank)
        if nb > self.cap_build or np_ > self.cap_probe:
            raise RuntimeError(f"receive buffers too small: {nb}/{self.cap_build} build rows, {np_}/{self.cap_probe} probe rows")
        e.partition_scatter_to(build_keys, build_vals, n_build, w, self.dst[0], self.dst[1], boff, stream=self.stream)
        e.partition_scatter_to(probe_keys, probe_vals, n_probe, w, self.dst[2], self.dst[3], poff, stream=self.stream)
        self.hdl.barrier(channel=0)                                     # every peer's stores have landed
        e.build(self.cols[0], self.cols[1], nb, stream=self.stream)
        e.probe_pairs(self.cols[2], self.cols[3], np_, out_key, out_build, out_probe, capacity, d_n_matches=d_count, sync=False,
                      stream=self.stream)
        item = self.buf.element_size()
        self.stats.sent_rows += n_build + n_probe
        self.stats.recv_rows += nb + np_
        self.stats.sent_bytes_remote += 2 * item * (n_build + n_probe - m[self.rank][0][self.rank] - m[self.rank][1][self.rank])
        return nb, np_


def plan_folded_exchange(counts, rank: int, regions: int, bounds):
    """Layout of the folded exchange.  counts[src][batch][dst * regions + region] = rows of `batch` (0 = the build relation,
    1.. = the probe chunks) that `src` sends to `dst` for table region `region`; bounds[c] = first row of probe chunk c
    inside the sender's probe relation (len = chunks + 1).

    Receive layout on every rank, per relation: batch-major (probe chunks one after the other), region-major inside a
    batch, source-minor inside a region -- so a batch is one contiguous, region-grouped segment the local build / probe
    can take as it is.  Returns a dict of numpy arrays:
      src_row[b, p]    first row of run (batch b, partition p) inside THIS rank's send buffer of that relation
      dst_row[b, p]    first row of that run inside the destination's receive buffer of that relation
      rows[b, p]       its length
      seg[b]           (first row, rows) of batch b inside THIS rank's receive buffer
      region_off[b]    regions + 1 row offsets of batch b's regions, relative to seg[b][0]  (build look-ahead)
    """
    import numpy as np
    c = np.asarray(counts, dtype=np.int64)                       # [world, batches, world * regions]
    world, batches, parts = c.shape
    assert parts == world * regions and len(bounds) == batches
    c4 = c.reshape(world, batches, world, regions)               # [src, batch, dst, region]
    per_region = c4.sum(axis=0)                                  # [batch, dst, region] rows arriving at dst for region
    total = per_region.sum(axis=2)                               # [batch, dst]
    seg_start = np.zeros_like(total)                             # batch 0 is its own relation; batches 1.. share one
    if batches > 2:
        seg_start[2:] = np.cumsum(total[1:-1], axis=0)
    region_base = np.cumsum(per_region, axis=2) - per_region     # exclusive over regions
    src_before = c4[:rank].sum(axis=0)                           # rows of lower-ranked sources, [batch, dst, region]
    dst_row = (seg_start[:, :, None] + region_base + src_before).reshape(batches, parts)
    mine = c[rank]                                               # [batch, parts]
    src_row = np.cumsum(mine, axis=1) - mine
    src_row[1:] += np.asarray(bounds[:-1], dtype=np.int64)[:, None]     # chunk c is scattered in place of its input rows
    seg = [(int(seg_start[b, rank]), int(total[b, rank])) for b in range(batches)]
    roff = np.zeros((batches, regions + 1), dtype=np.int64)
    roff[:, 1:] = np.cumsum(per_region[:, rank, :], axis=1)
    return {"src_row": src_row, "dst_row": dst_row, "rows": mine, "seg": seg, "region_off": roff}


def plan_blocked_exchange(counts, rank: int, regions: int, bounds):
    """Layout of the folded exchange with ONE BLOCK PER SOURCE at the receiver (few, large transfers).  Arguments as
    plan_folded_exchange.  Receive layout per relation: batch-major, then source-major; inside a source's block the
    runs are in the sender's own order, i.e. region-major.  The receiver walks a batch region by region through the
    segment list (region 0 of source 0, region 0 of source 1, ..., region 1 of source 0, ...).  Returns numpy arrays:
      src_row[b, p]      first row of run (batch b, partition p) inside THIS rank's send area (as plan_folded_exchange)
      own_row[b, g]      first row, inside this rank's receive area, of its OWN rows of region g (scattered in place)
      block_src[b, d], block_dst[b, d], block_rows[b, d]   the one transfer to destination d
      seg[b]             (first row, rows) of batch b inside this rank's receive area
      seg_first[b], seg_rows[b]   the batch's segment list in walking order (regions * world entries)
    """
    import numpy as np
    c = np.asarray(counts, dtype=np.int64)
    world, batches, parts = c.shape
    assert parts == world * regions and len(bounds) == batches
    c4 = c.reshape(world, batches, world, regions)               # [src, batch, dst, region]
    blocks = c4.sum(axis=3)                                      # [src, batch, dst] rows of the block src -> dst
    total = blocks.sum(axis=0)                                   # [batch, dst]
    seg_start = np.zeros_like(total)
    if batches > 2:
        seg_start[2:] = np.cumsum(total[1:-1], axis=0)
    before = np.cumsum(blocks, axis=0) - blocks                  # [src, batch, dst] rows of lower-ranked sources
    block_start = seg_start[None, :, :] + before                 # [src, batch, dst] where src's block starts at dst
    mine = c[rank]
    src_row = np.cumsum(mine, axis=1) - mine
    src_row[1:] += np.asarray(bounds[:-1], dtype=np.int64)[:, None]
    in_block = np.cumsum(c4, axis=3) - c4                        # [src, batch, dst, region] offset of a run in its block
    own_row = block_start[rank, :, rank][:, None] + in_block[rank, :, rank, :]
    seg_first = (block_start[:, :, rank][:, :, None] + in_block[:, :, rank, :]).transpose(1, 2, 0).reshape(batches, -1)
    seg_rows = c4[:, :, rank, :].transpose(1, 2, 0).reshape(batches, -1)       # [batch, region * world + src]
    return {"src_row": src_row, "own_row": own_row, "block_src": src_row[:, ::regions], "block_dst": block_start[rank],
            "block_rows": blocks[rank], "rows": mine, "seg": [(int(seg_start[b, rank]), int(total[b, rank])) for b in range(batches)],
            "seg_first": seg_first, "seg_rows": seg_rows}


class FoldedExchangeJoin:
    """Multi-GPU join whose exchange rides on the copy engines while the SMs partition and join.

    One pass per relation (dwj_xpart_*) groups the rows by (destination rank, table region of the destination's table).
    The runs of the OTHER ranks' partitions go to a local send area and are then pushed into the destination's receive
    area -- peer memory mapped through torch symmetric memory -- by plain device-to-device copies (dwj_copy_many: the
    copy engines over NVLink, no SM time); this rank's own partitions are scattered straight into its receive area.
    The receive layout is region-major, so the receiver gets its rows already grouped by table region and runs
    dwj_build_grouped / dwj_probe_pairs_grouped without the engine's own partition pass: the one partition pass of
    the single-GPU join is the only one here too.  The probe relation travels in `chunks` pieces: while the copy
    engines move piece c+1, the SMs scatter piece c+2 and probe piece c.  One all-gather of the count matrix plans
    everything (the step's one host sync).

    The result stays sharded and comes out as one segment per probe chunk: `self.segments[c]` = (first output row,
    capacity) and `self.chunk_counts[c]` = rows written there (device).
    """

    def __init__(self, engine, device, dtype, cap_build: int, cap_probe: int, max_build: int, max_probe: int, chunks: int = 2,
                 group=None, stream=None, transport: str = "ce1", push_ctas: int = 64, layout: str = "blocked"):
        import os
        # transport "ce1": copy engines (dwj_copy_many), ONE copy stream, peers in the order rank+1, rank+2, ... so that
        #   every receiver is one sender's target at a time (the classic all-to-all schedule): ~600 GB/s per GPU.
        #   "ce": one copy stream per peer -- concurrent copies to several peers share ~400 GB/s and finish together.
        #   "sm": dwj_push_runs (a few CTAs store into peer memory), ~430 GB/s.
        # layout "blocked": one block per source at the receiver -> one large transfer per peer and relation, the
        #   receiver walks the blocks region by region (dwj_*_segments).  The copy engines serialise copies at ~27 us
        #   each, so this is the layout for them.  Needs an engine with DWJ_FLAG_UNIQUE_BUILD_KEYS.
        # layout "region": the receive area itself is region-major -> ranks x regions runs per relation.
        self.transport = transport
        if layout == "blocked" and not getattr(engine, "flags", 1) & 1:     # DWJ_FLAG_UNIQUE_BUILD_KEYS
            layout = "region"                   # the segmented probe exists for unique build keys only
        self.layout = layout
        self.push_ctas = int(push_ctas)
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        self.np = np
        self.e = engine
        self.device = device
        self.dtype = dtype
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.regions = engine.xpart_regions(self.world)
        if self.regions == 0:
            raise ValueError(f"folded exchange needs a power-of-two world size, got {self.world}")
        self.parts = self.world * self.regions
        self.folded = self.regions > 1 and self.regions == engine.info()["radix_parts"]
        self.chunks = max(1, int(chunks))
        self.cap_build, self.cap_probe = int(cap_build), int(cap_probe)
        self.max_build, self.max_probe = int(max_build), int(max_probe)
        self.item = torch.empty(0, dtype=dtype).element_size()
        # One symmetric (peer-mapped) allocation: a key block and a payload block of identical layout
        #   [receive build | receive probe | send build | send probe]
        # so that one scatter can write a partition's keys and payloads at the same row offset of either block: this
        # rank's own partitions go straight into its receive area, the others into the send area.
        self.row_off = [0, self.cap_build, self.cap_build + self.cap_probe, self.cap_build + self.cap_probe + self.max_build]
        self.block_rows = self.row_off[3] + self.max_probe
        self.buf = symm_mem.empty(2 * self.block_rows, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        bases = np.array([int(p) for p in self.hdl.buffer_ptrs], dtype=np.uint64)
        self.block_ptr = [bases, bases + np.uint64(self.block_rows * self.item)]          # [keys | payloads][rank]
        self.blocks = [self.buf[:self.block_rows], self.buf[self.block_rows:]]
        self.recv_build = [blk[self.row_off[0]:self.row_off[1]] for blk in self.blocks]   # [keys, payloads]
        self.recv_probe = [blk[self.row_off[1]:self.row_off[2]] for blk in self.blocks]
        B = 1 + self.chunks
        self.counts = torch.zeros(B, self.parts, dtype=torch.int64, device=device)
        self.all_counts = torch.zeros(self.world, B, self.parts, dtype=torch.int64, device=device)
        self.region_off = torch.zeros(self.regions + 1, dtype=torch.int64, device=device)
        self.region_off_host = torch.zeros(self.regions + 1, dtype=torch.int64).pin_memory()
        self.chunk_counts = torch.zeros(self.chunks, dtype=torch.int64, device=device)
        self.stream = stream
        # Priorities: the barrier stream's one-CTA kernels must never queue behind a scatter or probe grid; the scatters
        # go ahead of the local build / probes, because every later transfer and probe waits for them (measured: with
        # the probes in front the last chunk left 1.5 ms later and the step grew from 5.3 to 6.1 ms on 2 GPUs).
        self.ps = torch.cuda.Stream(device=device, priority=-1)           # partition (scatter) stream
        self.js = torch.cuda.Stream(device=device)                        # local join stream, default priority
        self.xs = torch.cuda.Stream(device=device, priority=-2)           # barrier stream
        self.cps = torch.cuda.Stream(device=device, priority=-2)          # push kernel stream (transport "sm")
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(self.world)]
        self.copy_stream_ids = np.array([s.cuda_stream for s in self.copy_streams], dtype=np.uint64)
        self.trace = bool(int(os.environ.get("DWJ_XCHG_TRACE", "0")))     # development: device timeline of every step
        T = self.trace
        self.ev_plan = torch.cuda.Event(enable_timing=T)
        self.ev_scattered = [torch.cuda.Event(enable_timing=T) for _ in range(B)]
        self.ev_copied = [[torch.cuda.Event(enable_timing=T) for _ in range(self.world)] for _ in range(B)]
        self.ev_arrived = [torch.cuda.Event(enable_timing=T) for _ in range(B)]
        self.ev_t0, self.ev_hist, self.ev_built = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        self.ev_probed = [torch.cuda.Event(enable_timing=True) for _ in range(self.chunks)]
        self.last_trace = None
        self.segments = []
        self.stats = ExchangeStats()

    def join(self, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe,
             capacity, d_count):
        np, e, w, C, B = self.np, self.e, self.world, self.chunks, 1 + self.chunks
        cs = self.stream if self.stream is not None else torch.cuda.current_stream()
        bounds = [n_probe * c // C for c in range(C + 1)]
        rel = [(build_keys, build_vals, 0, n_build)] + [(probe_keys, probe_vals, bounds[c], bounds[c + 1] - bounds[c])
                                                        for c in range(C)]              # (keys, vals, first row, rows)
        if self.trace:
            self.ev_t0.record(cs)
        for b, (k, _, r0, n) in enumerate(rel):
            e.xpart_hist(k[r0:], n, w, self.counts[b], stream=cs)
        if self.trace:
            self.ev_hist.record(cs)
        with torch.cuda.stream(cs):
            # Ordered behind this rank's previous local join: once every rank's counts are in, every receive area is free.
            dist.all_gather_into_tensor(self.all_counts.view(-1), self.counts.view(-1), group=self.group)
            self.ev_plan.record(cs)
            m = self.all_counts.cpu().numpy()                             # the one host sync of the step
        blocked = self.layout == "blocked"
        plan = (plan_blocked_exchange if blocked else plan_folded_exchange)(m, self.rank, self.regions, bounds)
        nb, np_ = plan["seg"][0][1], sum(n for _, n in plan["seg"][1:])
        if nb > self.cap_build or np_ > self.cap_probe:
            raise RuntimeError(f"receive buffers too small: {nb}/{self.cap_build} build rows, {np_}/{self.cap_probe} probe rows")
        if np_ > capacity:
            raise RuntimeError(f"output capacity {capacity} below the {np_} probe rows this rank receives")
        if not blocked:
            self.region_off_host.copy_(torch.from_numpy(plan["region_off"][0]))
        dest_of_part = np.repeat(np.arange(w), self.regions)
        own = dest_of_part == self.rank
        item = np.uint64(self.item)
        remote_streams = [(d, st) for d, st in enumerate(self.copy_streams) if d != self.rank]
        # ---- partition stream: one scatter per batch; copy streams (one per destination): its runs ---------------------
        self.ps.wait_event(self.ev_plan)
        for b, (k, v, r0, n) in enumerate(rel):
            recv_off, send_off = (self.row_off[0], self.row_off[2]) if b == 0 else (self.row_off[1], self.row_off[3])
            # own partitions land in this rank's receive area at their final position, the others in the send area
            if blocked:
                own_dst = np.zeros(self.parts, dtype=np.int64)
                own_dst[own] = plan["own_row"][b]
            else:
                own_dst = plan["dst_row"][b]
            start = np.where(own, recv_off + own_dst, send_off + plan["src_row"][b])
            e.xpart_scatter(k[r0:], v[r0:], n, w, start, self.blocks[0], self.blocks[1], stream=self.ps)
            self.ev_scattered[b].record(self.ps)
            if blocked:         # one transfer per destination: the whole (destination, *) stretch of the send area
                dests = np.arange(w)
                nbytes = np.where(dests == self.rank, 0, plan["block_rows"][b]).astype(np.uint64) * item
                src = (send_off + plan["block_src"][b]).astype(np.uint64) * item
                dst = (recv_off + plan["block_dst"][b]).astype(np.uint64) * item
            else:
                dests = dest_of_part
                nbytes = np.where(own, 0, plan["rows"][b]).astype(np.uint64) * item
                src = (send_off + plan["src_row"][b]).astype(np.uint64) * item
                dst = (recv_off + plan["dst_row"][b]).astype(np.uint64) * item
            dsts = np.concatenate([self.block_ptr[blk][dests] + dst for blk in (0, 1)])
            srcs = np.concatenate([self.block_ptr[blk][self.rank] + src for blk in (0, 1)])
            self.xs.wait_event(self.ev_scattered[b])
            if self.transport == "sm":
                # every rank starts with a different peer (rank+1, rank+2, ...), so no receiver is everybody's target at once
                order = np.argsort((np.concatenate([dests, dests]) - self.rank - 1) % w, kind="stable")
                self.cps.wait_event(self.ev_scattered[b])
                e.push_runs(dsts[order], srcs[order], np.concatenate([nbytes, nbytes])[order] // item, self.push_ctas, stream=self.cps)
                self.ev_copied[b][0].record(self.cps)
                self.xs.wait_event(self.ev_copied[b][0])
            elif self.transport == "ce1":
                # one copy stream, peers in the order rank+1, rank+2, ...: at any time every receiver is the target of one sender
                order = np.argsort((np.concatenate([dests, dests]) - self.rank - 1) % w, kind="stable")
                st = self.copy_streams[0]
                st.wait_event(self.ev_scattered[b])
                e.copy_many_arrays(dsts[order], srcs[order], np.concatenate([nbytes, nbytes])[order],
                                   np.full(len(order), self.copy_stream_ids[0], dtype=np.uint64))
                self.ev_copied[b][0].record(st)
                self.xs.wait_event(self.ev_copied[b][0])
            else:
                for _, st in remote_streams:
                    st.wait_event(self.ev_scattered[b])
                e.copy_many_arrays(dsts, srcs, np.concatenate([nbytes, nbytes]), np.concatenate([self.copy_stream_ids[dests]] * 2))
                for d, st in remote_streams:
                    self.ev_copied[b][d].record(st)
                    self.xs.wait_event(self.ev_copied[b][d])
            with torch.cuda.stream(self.xs):
                self.hdl.barrier(channel=0)                               # every rank's runs of this batch have landed everywhere
            self.ev_arrived[b].record(self.xs)
        # ---- join stream: local build, then one probe per received chunk; the caller's stream rejoins at the end ---------
        caller = cs
        cs = self.js
        cs.wait_event(self.ev_plan)
        cs.wait_event(self.ev_arrived[0])
        if self.folded and blocked:
            e.build_segments(self.recv_build[0], self.recv_build[1], plan["seg_first"][0], plan["seg_rows"][0], w, stream=cs)
        elif self.folded:
            with torch.cuda.stream(cs):
                self.region_off.copy_(self.region_off_host, non_blocking=True)
            e.build_grouped(self.recv_build[0], self.recv_build[1], nb, self.region_off, stream=cs)
        else:                   # regions not folded into the exchange: the engine groups the received rows itself
            e.build(self.recv_build[0], self.recv_build[1], nb, stream=cs)
        if self.trace:
            self.ev_built.record(cs)
        probe = e.probe_pairs_grouped if self.folded else e.probe_pairs
        for c, (row0, rows) in enumerate(plan["seg"][1:]):
            cs.wait_event(self.ev_arrived[1 + c])
            if self.folded and blocked:
                e.probe_pairs_segments(self.recv_probe[0], self.recv_probe[1], plan["seg_first"][1 + c], plan["seg_rows"][1 + c],
                                       None if out_key is None else out_key[row0:], out_build[row0:], out_probe[row0:], rows,
                                       d_n_matches=self.chunk_counts[c:], sync=False, stream=cs)
            else:
                probe(self.recv_probe[0][row0:], self.recv_probe[1][row0:], rows, None if out_key is None else out_key[row0:],
                      out_build[row0:], out_probe[row0:], rows, d_n_matches=self.chunk_counts[c:], sync=False, stream=cs)
            if self.trace:
                self.ev_probed[c].record(cs)
        with torch.cuda.stream(cs):
            torch.sum(self.chunk_counts, dim=0, keepdim=True, out=d_count)
        caller.wait_stream(cs)
        # the send area may be rewritten once this step's copies are done: the next scatter waits for them
        for st in {"sm": [self.cps], "ce1": [self.copy_streams[0]]}.get(self.transport, [st for _, st in remote_streams]):
            self.ps.wait_stream(st)
        self.segments = plan["seg"][1:]
        if self.trace:
            torch.cuda.synchronize()
            t = lambda ev: round(self.ev_t0.elapsed_time(ev), 3)    # noqa: E731
            self.last_trace = {"hist": t(self.ev_hist), "plan": t(self.ev_plan), "scattered": [t(x) for x in self.ev_scattered],
                               "copied": [[t(self.ev_copied[b][d]) for d in ([0] if self.transport in ("sm", "ce1") else [d for d, _ in remote_streams])]
                                          for b in range(B)],
                               "arrived": [t(x) for x in self.ev_arrived], "built": t(self.ev_built),
                               "probed": [t(x) for x in self.ev_probed]}
        sent_local = int(plan["rows"][:, own].sum())
        self.stats.sent_rows += n_build + n_probe
        self.stats.recv_rows += nb + np_
        self.stats.sent_bytes_remote += 2 * self.item * (n_build + n_probe - sent_local)
        return nb, np_
