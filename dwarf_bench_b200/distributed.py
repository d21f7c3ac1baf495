"""Multi-GPU join, one process per GPU: torch.distributed plumbing around the C-ABI exchange join (dwj_xj_*).

No reference counterpart (the reference is single-device).  Every rank holds an arbitrary
(arrival-order) slice of both relations; equal keys must meet on one GPU, so each relation needs one exchange step.

  PullExchangeJoin  (default) a thin caller of csrc/dwj_xj.cu: torch symmetric memory supplies one peer-mapped block per
                    rank, everything else -- counting, planning, the senders' partition passes, flags in peer memory,
                    the receivers' kernels pulling their rows out of the senders' blocks over NVLink -- happens behind
                    dwj_xj_join.  No collective on the data path.
  ExchangeJoin      partition locally, then NCCL all-to-all-v: the baseline the pull exchange is measured against and
                    the fallback when peer mapping is unavailable.

The device work of ExchangeJoin is delegated to a `JoinOps` object so that its host-side logic (split sizes, buffer
sizing, ordering of collectives) can be exercised on CPU with gloo in tests, where a numpy stand-in supplied BY THE
TEST plays the device.  The product only ever calls the C ABI.
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


class PullExchangeJoin:
    """This rank's dwj_xj over a torch symmetric-memory block (world > 1) or a plain device buffer (world == 1)."""

    def __init__(self, engine, device, max_build_rows: int, max_probe_rows: int, chunk_rows: int = 0, passes: int = 1,
                 recv_slack: float = 0.0, force_scatter_pull: bool = False, group=None, stream=None):
        from .capi import ExchangeJoinRank
        self.e = engine
        self.device = device
        self.stream = stream
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        args = (self.rank, self.world, int(max_build_rows), int(max_probe_rows), int(chunk_rows), int(passes), float(recv_slack),
                bool(force_scatter_pull))
        nbytes = ExchangeJoinRank.block_bytes(engine, *args)
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self.group = group or dist.group.WORLD
            # every rank must ask for the same size (it is a function of the shared configuration; make sure)
            sizes = torch.tensor([nbytes], dtype=torch.int64, device=device)
            dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=self.group)
            nbytes = int(sizes.item())
            self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            self.hdl = symm_mem.rendezvous(self.buf, self.group)
            blocks = [int(p) for p in self.hdl.buffer_ptrs]
        else:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            blocks = [self.buf.data_ptr()]
        self.block_bytes = nbytes
        self.xj = ExchangeJoinRank(engine, self.rank, self.world, int(max_build_rows), int(max_probe_rows), blocks, int(chunk_rows),
                                   int(passes), float(recv_slack), bool(force_scatter_pull))
        self.info = self.xj.describe()
        if self.world > 1:
            dist.barrier(group=self.group)          # every control block is cleared before anyone raises a flag in it

    def join(self, build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe,
             capacity, d_count):
        """The rank's share of the global join lands in out_* (compacted, any order), its row count in d_count (device
        int64[1]).  Asynchronous apart from one count exchange per pass."""
        self.xj.join(build_keys, build_vals, n_build, probe_keys, probe_vals, n_probe, out_key, out_build, out_probe, capacity,
                     d_count, stream=self.stream if self.stream is not None else torch.cuda.current_stream())

    def timings(self) -> dict:
        return self.xj.sync_timings()

    def close(self):
        self.xj.close()
