#!/usr/bin/env python
"""Developer tool: time the KNN outlier statistic at a realistic cloud size (run on the GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import pgdvs_b200
from pgdvs_b200 import ops, synthetic
from pgdvs_b200.dyn_renderer import unproject_warp_project, opencv_to_p3d_camera

dev = torch.device("cuda:0")
wl = synthetic.make_workload("c1_nvidia_1view", dev)
pairs, cams = wl.jobs(range(1))
p3d = [opencv_to_p3d_camera(K, c, wl.H, wl.W) for (K, c) in cams]
cloud = unproject_warp_project(pairs[:1], p3d, wl.H, wl.W, dev, want_world=True)
P = int(cloud["total"].item())
pw = cloud["xyz_world"][:P].contiguous()
for n in (20000, 60000, P):
    q = pw[:n]
    ops.knn_mean_dist(q, q, 51, skip_first=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.knn_mean_dist(q, q, 51, skip_first=1)
    e1.record()
    torch.cuda.synchronize()
    print(f"P={n}: knn_mean_dist(K=51) {e0.elapsed_time(e1):.2f} ms")
