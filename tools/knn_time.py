#!/usr/bin/env python
"""Developer tool: time the KNN outlier statistic at realistic cloud sizes, for the in-tree library and
for the builds named in LIBS (exp_libs/lib_<name>.so), and compare their values (run on the GPU box)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from pgdvs_b200 import _cabi, ops, synthetic
from pgdvs_b200.dyn_renderer import unproject_warp_project, opencv_to_p3d_camera

ROOT = Path(__file__).resolve().parent.parent
dev = torch.device("cuda:0")
wl = synthetic.make_workload("c1_nvidia_1view", dev)
pairs, cams = wl.jobs(range(1))
p3d = [opencv_to_p3d_camera(K, c, wl.H, wl.W) for (K, c) in cams]
cloud = unproject_warp_project(pairs[:1], p3d, wl.H, wl.W, dev, want_world=True)
P = int(cloud["total"].item())
pw = cloud["xyz_world"][:P].contiguous()
libs = [("in-tree", _cabi.LIB_PATH)] + [(n, ROOT / "exp_libs" / f"lib_{n}.so") for n in os.environ.get("LIBS", "").split(",") if n]
ref = {}
for name, path in libs:
    _cabi.LIB_PATH, _cabi._lib = path, None
    for n in (20000, 60000, P):
        q = pw[:n]
        ops.knn_mean_dist(q, q, 51, skip_first=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ops.knn_mean_dist(q, q, 51, skip_first=1)
        e1.record()
        torch.cuda.synchronize()
        msg = ""
        if n in ref:
            d = (out - ref[n]).abs() / ref[n].abs().clamp_min(1e-30)
            msg = f" max rel diff vs first build {float(d.max()):.2e}"
        else:
            ref[n] = out.clone()
        print(f"{name}: P={n}: knn_mean_dist(K=51, self) {e0.elapsed_time(e1) / 5:.3f} ms{msg}")
