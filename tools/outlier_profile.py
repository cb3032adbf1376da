#!/usr/bin/env python
"""Developer tool: device time per kernel and wall time of the batched render with the statistical
outlier filter (C2, 144 views, 24 source-pair clouds).  Run on the GPU box."""
import sys
import time
from pathlib import Path
from types import SimpleNamespace

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from pgdvs_b200 import synthetic  # noqa: E402
from pgdvs_b200.dyn_renderer import render_views_filtered  # noqa: E402

dev = torch.device("cuda:0")
wl = synthetic.make_workload("c2_nvidia_seq", dev)
pairs, cams = wl.jobs(range(wl.n_views))
cfg = SimpleNamespace(dyn_pcl_outlier_knn=50, dyn_pcl_outlier_std_thres=0.1)


def step():
    return render_views_filtered(pairs, cams, wl.H, wl.W, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                 static_rgb=wl.static_rgb, render_cfg=cfg, return_fragments=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("wall ms per step", 1e3 * (time.perf_counter() - t0) / 5)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
for evt in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    if evt.device_time_total > 0:
        print(f"{evt.key[:80]:80s} {evt.device_time_total / 1e3:9.3f} ms x{evt.count}")
print("host-side self time (top):")
for evt in sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:8]:
    print(f"{evt.key[:80]:80s} {evt.self_cpu_time_total / 1e3:9.3f} ms x{evt.count}")
