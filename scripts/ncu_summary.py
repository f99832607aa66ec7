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
    mi = hdr.index('Metric Name')
    agg = collections.OrderedDict()
    for row in r:
        if row[mi] != 'gpu__time_duration.sum':
            continue
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


def stage(path, bakes, out_prefix=None):
    """Per-kernel totals of a multi-metric launch list (time, DRAM bytes, warp instructions, issue-slot utilisation, active lanes), per bake,
    and the two JSON files bench.py reads: <out_prefix>_stage_traffic.json and <out_prefix>_issue.json."""
    import json
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, mi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Metric Name'), hdr.index('ID')
    per = collections.OrderedDict()   # (id) -> dict
    for row in r:
        d = per.setdefault(row[ii], {'name': row[ki].split('(')[0].replace('void ', '').replace('ommb200::', '')[:60]})
        v = float(row[vi].replace(',', ''))
        u = row[ui]
        if row[mi] == 'gpu__time_duration.sum':
            v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
        if row[mi].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d[row[mi]] = v
    # 'auto': a bake is counted when its LAST kernel (WriteIndexBuffer) was captured; launches after the last complete bake are dropped
    ids = sorted(per, key=int)
    if bakes == 'auto':
        ends = [i for i in ids if per[i]['name'].startswith('WriteIndexBuffer')]
        starts = [i for i in ids if per[i]['name'].startswith('SetupTriangles')]
        first = int(starts[0]) if starts else 0
        last = int(ends[-1]) if ends else int(ids[-1])
        ids = [i for i in ids if first <= int(i) <= last]
        bakes = max(1, len([i for i in ends if int(i) >= first]))
    else:
        bakes = int(bakes)
    agg = collections.OrderedDict()
    for d in (per[i] for i in ids):
        a = agg.setdefault(d['name'], {'launches': 0, 'ms': 0.0, 'rd': 0.0, 'wr': 0.0, 'inst': 0.0, 'issue_w': 0.0, 'lanes_w': 0.0})
        ms = d.get('gpu__time_duration.sum', 0.0)
        inst = d.get('smsp__inst_executed.sum', 0.0)
        a['launches'] += 1; a['ms'] += ms; a['rd'] += d.get('dram__bytes_read.sum', 0.0); a['wr'] += d.get('dram__bytes_write.sum', 0.0); a['inst'] += inst
        a['issue_w'] += ms * d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.0)
        a['lanes_w'] += inst * d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0.0)
    tot = sum(a['ms'] for a in agg.values())
    print(f"{bakes} bakes captured; per bake: {tot / bakes:.3f} ms of kernel time (serialised, cold caches: compare SHARES)")
    print(f"{'kernel':50s} {'launches':>8s} {'ms/bake':>9s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'warp inst/bake':>15s} {'issue %':>8s} {'lanes':>6s}")
    hier = {'ms': 0.0, 'rd': 0.0, 'wr': 0.0, 'inst': 0.0, 'issue_w': 0.0, 'lanes_w': 0.0}
    rows = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
        issue = a['issue_w'] / a['ms'] if a['ms'] else 0.0
        lanes = a['lanes_w'] / a['inst'] if a['inst'] else 0.0
        print(f"{k:50s} {a['launches'] // bakes:8d} {a['ms'] / bakes:9.3f} {100 * a['ms'] / tot:6.2f}% {a['rd'] / bakes / 1e6:11.1f} {a['wr'] / bakes / 1e6:11.1f} {a['inst'] / bakes:15.3e} {issue:8.1f} {lanes:6.2f}")
        rows[k] = {'ms_per_bake': a['ms'] / bakes, 'warp_inst_per_bake': a['inst'] / bakes, 'issue_active_pct': issue, 'lanes_active': lanes,
                   'dram_read_bytes_per_bake': a['rd'] / bakes, 'dram_write_bytes_per_bake': a['wr'] / bakes}
        if k.startswith('Hier'):
            for f in hier:
                hier[f] += a[f]
    if out_prefix:
        allinst = sum(a['inst'] for a in agg.values())
        alllanes = sum(a['lanes_w'] for a in agg.values())
        issue = {'source': f'{path}: ncu --metrics smsp__inst_executed.sum, smsp__issue_active.avg.pct_of_peak_sustained_active, smsp__thread_inst_executed_per_inst_executed.ratio over every launch of {bakes} bakes (bench.py under ncu)',
                 'warp_inst_per_bake': allinst / bakes, 'classify_stage_warp_inst_per_bake': hier['inst'] / bakes,
                 'issue_active_pct': sum(a['issue_w'] for a in agg.values()) / tot, 'classify_stage_issue_active_pct': hier['issue_w'] / hier['ms'] if hier['ms'] else None,
                 'lanes_active': alllanes / allinst if allinst else None, 'classify_stage_lanes_active': hier['lanes_w'] / hier['inst'] if hier['inst'] else None,
                 'issue_peak_warp_inst_per_s': 148 * 4 * 1.965e9,
                 'ms_at_issue_peak': allinst / bakes / (148 * 4 * 1.965e9) * 1e3,
                 'note': 'frac_of_issue_peak = ms_at_issue_peak / measured ms_per_step (148 SMs x 4 schedulers x 1 warp instruction per cycle at 1965 MHz)',
                 'kernels': rows}
        with open(out_prefix + '_issue.json', 'w') as f:
            json.dump(issue, f, indent=1)
        traffic = {'stage': 'classification (the Hier* kernels)', 'workload': 'C3 full size, 1 GPU', 'dram_bytes_read_per_bake': hier['rd'] / bakes,
                   'dram_bytes_written_per_bake': hier['wr'] / bakes, 'whole_bake_dram_bytes_read': sum(a['rd'] for a in agg.values()) / bakes,
                   'whole_bake_dram_bytes_written': sum(a['wr'] for a in agg.values()) / bakes,
                   'source': f'{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the launches, {bakes} bakes captured, per bake)'}
        with open(out_prefix + '_stage_traffic.json', 'w') as f:
            json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel, 'stage': stage}[sys.argv[1]](*sys.argv[2:])
