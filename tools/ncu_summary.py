#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): headline counters + stall reasons + hottest SASS lines."""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_sectors.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(rep, top=25):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for vals in raw[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"### {name[:110]}")
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:78s} {units[i]:14s} {vals[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    blocks, cur = [], None
    for row in src:                                  # one block per kernel: "Kernel Name" line, header line, data
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1] if len(row) > 1 else "?", "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    for blk in blocks:
        if len(blk["rows"]) < 2:
            continue
        h = blk["rows"][0]
        idx = {k: i for i, k in enumerate(h)}
        data = [r for r in blk["rows"][1:] if len(r) == len(h)]
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
        agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
        print(f"## source: {blk['name'][:100]}")
        print(f"  SASS instructions: {len(data)}, stall samples: {tot}")
        print("  stalls: " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 > tot))
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:top]:
            st = sorted(((s[6:], int(r[idx[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
            print(f"   {r[idx['Address']][-5:]} {int(r[idx['# Samples']] or 0):7d} {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']][:72]:72s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
