#!/usr/bin/env python
"""Summarise ncu outputs into text files for profiles/:
   scripts/ncu_summary.py launches <launches.csv>          -> per-kernel device-time shares
   scripts/ncu_summary.py kernel <raw.csv>                 -> key metrics of one captured kernel (ncu -i rep --page raw --csv)
"""
import collections
import csv
import sys


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for row in r:
        v = float(row[vi].replace(',', ''))
        u = row[ui]
        v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
        a = agg.setdefault(row[ki].split('(')[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total device time in captured launches: {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:11.3f} ms {v[0]:5d}x {100 * v[1] / tot:6.2f}%  avg {v[1] / v[0]:9.4f} ms  {k}")


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'launch__func_cache_config', 'smsp__cycles_active.avg']


def kernel(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('kernel:', name[:150])
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:95s} {units[i]:18s} {vals[i]}")


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
