#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics and the executed-instruction opcode mix per kernel.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [sites_per_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def main():
    rep = sys.argv[1]
    sites = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print('==', r[hdr.index('Kernel Name')][:100])
        for k in KEYS:
            if k in hdr:
                print(f'   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
        stall = [(h, r[i]) for i, h in enumerate(hdr) if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
        for h, v in sorted(stall, key=lambda t: -float(t[1] or 0))[:8]:
            print(f'   {h:75s} {v:>16s} %')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    for bi, st in enumerate(starts[:1]):
        hdr = rows[st]
        end = starts[bi + 1] - 1 if bi + 1 < len(starts) else len(rows)
        iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
        ops, samp, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[st + 1:end]:
            try:
                e = int(r[iE])
            except (ValueError, IndexError):
                continue
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
            op = m.group(2) if m else r[iS][:10]
            ops[op] += e
            tot += e
            try:
                samp[op] += int(r[iN])
            except ValueError:
                pass
        print('total warp instructions', tot, '' if not sites else f'= {32 * tot / sites:.1f} thread-instr per site-update')
        for op, c in ops.most_common(32):
            per = f'{32 * c / sites:7.1f}/site' if sites else ''
            print(f'   {op:12s} {c:11d} {100 * c / tot:5.1f}% {per}  samples {samp[op]}')


if __name__ == '__main__':
    main()
