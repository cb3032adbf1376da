#!/usr/bin/env python
"""Developer tool: where does the host time of the end-to-end step go?  (run on the GPU box)"""
import sys, time, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import pgdvs_b200
from pgdvs_b200 import synthetic
from pgdvs_b200.dyn_renderer import prepare_views, render_prepared

dev = torch.device("cuda:0")
wl = synthetic.make_workload("c2_nvidia_seq", dev)
cp, cc = wl.jobs(range(36))
for _ in range(3):
    p = prepare_views(cp, cc, wl.H, wl.W, dev)
    o = render_prepared(p, radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb[:36])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    p = prepare_views(cp, cc, wl.H, wl.W, dev)
t1 = time.perf_counter()
for _ in range(20):
    o = render_prepared(p, radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb[:36])
t2 = time.perf_counter()
torch.cuda.synchronize()
t3 = time.perf_counter()
print(f"prepare_views(36 views): {(t1 - t0) / 20 * 1e3:.3f} ms host   render_prepared launch: {(t2 - t1) / 20 * 1e3:.3f} ms host   (+sync {(t3 - t2) * 1e3:.1f} ms)")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    p = prepare_views(cp, cc, wl.H, wl.W, dev)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
