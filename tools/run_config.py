#!/usr/bin/env python
"""Developer tool: run one synthetic config through the batched hot path a few times (for ncu).
    python tools/run_config.py c4_davis [n_views] [K] [radius]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from pgdvs_b200 import synthetic
from pgdvs_b200.dyn_renderer import prepare_views, render_prepared

name = sys.argv[1]
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else None
K = int(sys.argv[3]) if len(sys.argv) > 3 else None
radius = float(sys.argv[4]) if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
wl = synthetic.make_workload(name, dev, n_views=n_views, K=K, radius=radius)
pairs, cams = wl.jobs(range(wl.n_views))
prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
for _ in range(3):
    out = render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb,
                          return_fragments=True)
torch.cuda.synchronize()
print("ok", name, wl.n_views, int(out["cloud"]["total"].item()))
