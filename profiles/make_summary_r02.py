#!/usr/bin/env python
"""Regenerate the tables of profiles/r02_ncu_summary.md from the round-2 ncu captures.

    python profiles/make_summary_r02.py gpurun_out/r02  > /tmp/tables.md

gpurun_out/r02 holds: launches_c2.csv (ncu --metrics gpu__time_duration.sum ... bench.py) and the
`ncu --set full --import-source on` reports c2_full / c3_tile / c4_tile / c5_k8 / c5_k32 (.ncu-rep).
Also rewrites profiles/raster_traffic.json from the C2 rasterizer's DRAM counters."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_summary import WANT, launch_table  # noqa: E402

EXTRA = ['launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
         'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
         'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
         'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
         'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
         'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
         'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
         'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_requests_srcunit_tex_op_red.sum',
         'lts__t_requests_srcunit_tex_op_atom_dot_alu.sum', 'lts__t_sectors_srcunit_tex_op_atom.sum']


def tables(rep):
    txt = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    out = collections.OrderedDict()
    for r in rows[2:]:
        out[r[h.index('Kernel Name')]] = {w: (r[h.index(w)], u[h.index(w)]) for w in WANT + EXTRA if w in h}
    return out


def main():
    d = Path(sys.argv[1])
    print("### launch list (C2 bench)\n")
    print(launch_table(d / "launches_c2.csv"))
    for label, rep in (("C2 (bench.py, 144 views)", "c2_full"), ("C4 (16 views)", "c4_tile"), ("C3 (8 views)", "c3_tile"),
                       ("C5 K=8 r=0.01 (2 views)", "c5_k8"), ("C5 K=32 r=0.02 (2 views)", "c5_k32"),
                       ("KNN outlier statistic (K = 51, tools/knn_time.py; the launch with -s 12)", "knn_warp")):
        p = d / (rep + ".ncu-rep")
        if not p.exists():
            continue
        for kern, m in tables(p).items():
            print(f"\n### {label}: `{kern}`\n")
            print("| metric | value | unit |\n|---|---|---|")
            for k, (v, unit) in m.items():
                try:
                    v = f"{float(v.replace(',', '')):.6g}"
                except ValueError:
                    pass
                print(f"| `{k}` | {v} | {unit} |")
            if "k_raster_pair" in kern:
                def gb(key):
                    v, unit = m[key]
                    return float(v.replace(',', '')) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
                rd, wr = gb('dram__bytes_read.sum'), gb('dram__bytes_write.sum')
                Path(__file__).resolve().parent.joinpath("raster_traffic.json").write_text(json.dumps({
                    "kernel": kern, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                    "source": "ncu --set full, one launch of the 144-view C2 step (profiles/r02_ncu_summary.md)"}, indent=1))


if __name__ == "__main__":
    main()
