#!/usr/bin/env python
"""Regenerate the tables of profiles/r01_ncu_summary.md from ncu output.

    python profiles/make_summary.py <launches.csv> <full.ncu-rep>   > tables.md

launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ...
full.ncu-rep : ncu --set full --import-source on --clock-control none -k regex:... -o ...
Also rewrites profiles/raster_traffic.json from the rasterizer's DRAM counters."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__grid_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def launch_table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        v = v / 1e3 if r[iu] in ('ns', 'nsecond') else (v * 1e3 if r[iu] in ('ms', 'msecond') else v)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if 'pgdvs' in k}
    tot = sum(v[1] for v in ours.values())
    out = ["| kernel | launches | avg µs | share of our kernels |", "|---|---|---|---|"]
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:70]}` | {n} | {t / n:.1f} | {t / tot * 100:.1f} % |")
    return "\n".join(out)


def metric_tables(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    res = collections.OrderedDict()
    for r in rows[2:]:
        res[r[h.index('Kernel Name')]] = {w: (r[h.index(w)], u[h.index(w)]) for w in WANT if w in h}
    return res


def fmt(v):
    try:
        return f"{float(v):.6g}"
    except ValueError:
        return v


if __name__ == "__main__":
    print(launch_table(sys.argv[1]))
    res = metric_tables(sys.argv[2])
    for name, m in res.items():
        print(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---|---|")
        for w, (val, un) in m.items():
            print(f"| `{w}` | {fmt(val)} | {un} |")
        if 'k_raster' in name:
            rd, wr = float(m['dram__bytes_read.sum'][0]), float(m['dram__bytes_write.sum'][0])
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
            rd *= scale[m['dram__bytes_read.sum'][1]]
            wr *= scale[m['dram__bytes_write.sum'][1]]
            p = Path(__file__).resolve().parent / "raster_traffic.json"
            old = json.loads(p.read_text()) if p.exists() else {}
            old.update({"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr})
            p.write_text(json.dumps(old, indent=1) + "\n")
