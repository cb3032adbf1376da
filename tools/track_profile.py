#!/usr/bin/env python
"""Developer tool: where the time of PGDVSDynamicTrackRenderer.forward goes on the C3 + tracks workload of
bench.py's matrix (device time per kernel, wall time of one call).  Run on the GPU box."""
import sys
import time
from pathlib import Path
from types import SimpleNamespace

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from pgdvs_b200 import synthetic, track  # noqa: E402

dev = torch.device("cuda:0")
data = synthetic.make_data_dict("c3_iphone", dev, n_views=16, n_track_one_side=3, closest_mask_mode="ellipse")
c3 = synthetic.CONFIGS["c3_iphone"]
rcfg = SimpleNamespace(dyn_render_type="pcl", dyn_render_pcl_pt_radius=c3["radius"], dyn_render_pcl_pts_per_pixel=c3["K"],
                       dyn_render_use_flow_consistency=False, dyn_pcl_remove_outlier=False, dyn_pcl_outlier_knn=50,
                       dyn_pcl_outlier_std_thres=0.1, dyn_pcl_track_track2base_thres_mult=50)
rend = track.PGDVSDynamicTrackRenderer(tracker=synthetic.SyntheticTracker(seed=1234, p_visible=0.5))
for _ in range(2):
    rend(data, None, rcfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
rend(data, None, rcfg)
torch.cuda.synchronize()
print("forward wall ms", 1e3 * (time.perf_counter() - t0))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    rend(data, None, rcfg)
    torch.cuda.synchronize()
tot = 0.0
for evt in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:18]:
    if evt.device_time_total > 0:
        print(f"{evt.key[:80]:80s} {evt.device_time_total / 1e3:9.3f} ms x{evt.count}")
print("sum of device time ms", sum(e.device_time_total for e in prof.key_averages() if e.device_type.name == "CUDA") / 1e3)
