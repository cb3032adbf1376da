#!/usr/bin/env python
"""Run bench.py over the five BASELINE.json configs (1 GPU) and print a markdown table.

    python tools/sweep.py [--steps 10] > gpurun_out/sweep.md
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CASES = [
    ("C1 NVIDIA 1 view", ["--workload", "c1_nvidia_1view"]),
    ("C2 NVIDIA sequence, 144 views", ["--workload", "c2_nvidia_seq"]),
    ("C2, i.i.d. flow (stress)", ["--workload", "c2_nvidia_seq", "--flow", "iid"]),
    ("C2, fragments not materialised", ["--workload", "c2_nvidia_seq", "--no-fragments"]),
    ("C3 iPhone 360x480, S=6, K=16, 16 views", ["--workload", "c3_iphone"]),
    ("C4 DAVIS 480x854, 80 frames", ["--workload", "c4_davis"]),
    ("C5 1080p, S=8, K=8, r=0.01, 8 views", ["--workload", "c5_stress", "--K", "8", "--radius", "0.01"]),
    ("C5 1080p, K=16, r=0.005", ["--workload", "c5_stress", "--K", "16", "--radius", "0.005"]),
    ("C5 1080p, K=32, r=0.02", ["--workload", "c5_stress", "--K", "32", "--radius", "0.02", "--views", "4"]),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    print("| config | views/step | points/view | K | step ms | views/s | Mpoints/s | k_raster ms | roofline frac | e2e views/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for name, extra in CASES:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--steps", str(args.steps), "--warmup", "3",
               "--no-cpu-baseline"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(f"| {name} | failed: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else '?'} |")
            continue
        d = json.loads(line[-1])
        c, rf = d["config"], d["roofline"]
        print(f"| {name} | {c['views_per_gpu']} | {c['points_per_view']} | {c['points_per_pixel']} | "
              f"{d['ms_per_step']:.3f} | {d['value']:.0f} | {d['mpoints_per_s']:.0f} | {rf['avg_launch_ms']:.3f} | "
              f"{100 * rf['frac']:.1f} % | {d['e2e']['value']:.0f} |", flush=True)


if __name__ == "__main__":
    main()
