"""Work counters of the rasterize-and-composite stage on the synthetic BASELINE configs
(candidate tests per pixel, in-radius hits, kept hits, list-full rate), computed on
the CPU from the oracle's NDC cloud of one target view.  They explain why the kernel is
instruction-bound rather than HBM-bound and are cross-checked against what the
geometry predicts.

    python -m pytest tests/test_work_counters.py -s            # prints the table rows
    WORK_COUNTERS_ALL=1 python -m pytest tests/test_work_counters.py -s   # + C4, C5 (minutes)
"""
import math
import os

import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

from oracle import pgdvs_ref as ref
from pgdvs_b200 import synthetic

CASES = [("c1_nvidia_1view", {}), ("c3_iphone", {})]
if os.environ.get("WORK_COUNTERS_ALL"):
    CASES += [("c4_davis", {}), ("c5_stress", {}), ("c5_stress", dict(K=16, radius=0.005)), ("c5_stress", dict(K=32, radius=0.02))]


def _ndc_cloud(wl, view):
    sc, H, W = wl.scene, wl.H, wl.W
    Kc, c2w_t = wl.view_cams[view]
    flat_tgt = torch.cat([torch.tensor([float(H), float(W)]), torch.from_numpy(Kc).reshape(-1),
                          torch.from_numpy(c2w_t).reshape(-1)])
    pcl = []
    for p in wl.view_pairs[view]:
        a = p._src_frames
        o = ref.compute_dyn_pcl(
            dyn_mask_1=sc.mask[a[0]], rgb_1=sc.rgb[a[0]], depth_1=sc.depth[a[0]], flow_12=p.flow_12.reshape(H, W, 2),
            flow_12_occ_mask=torch.zeros(H, W, 1), rgb_2=sc.rgb[a[1]], depth_2=sc.depth[a[1]],
            K_1=torch.from_numpy(sc.K), c2w_1=torch.from_numpy(sc.c2w[a[0]]), K_2=torch.from_numpy(sc.K),
            c2w_2=torch.from_numpy(sc.c2w[a[1]]), time_1=torch.tensor(sc.times[a[0]]),
            time_2=torch.tensor(sc.times[a[1]]), time_tgt=torch.tensor(p._t_tgt))
        pcl.append(o["pcl"])
    return ref.world_to_ndc(torch.cat(pcl), ref.camera_from_flat_cam(flat_tgt)).numpy()


@pytest.mark.parametrize("name,kw", CASES)
def test_work_counters(name, kw):
    wl = synthetic.make_workload(name, torch.device("cpu"), n_views=1, **kw)
    H, W, K, r = wl.H, wl.W, wl.K, wl.radius
    ndc = _ndc_cloud(wl, 0)
    s = min(H, W) / 2.0
    r_px = r * s
    halo = int(math.floor(r_px + 0.5 + 1.0 / 64.0))  # csrc/common.cuh: halo_cells
    u = W / 2.0 - ndc[:, 0].astype(np.float64) * s      # OpenCV pixel coordinates
    v = H / 2.0 - ndc[:, 1].astype(np.float64) * s
    ok = ndc[:, 2] >= 0
    u, v = u[ok], v[ok]
    # cell of the nearest pixel centre, extended grid
    gx = np.floor(u).astype(np.int64) + halo
    gy = np.floor(v).astype(np.int64) + halo
    GW, GH = W + 2 * halo, H + 2 * halo
    inside = (gx >= 0) & (gx < GW) & (gy >= 0) & (gy < GH)
    counts = np.zeros((GH, GW), np.int64)
    np.add.at(counts, (gy[inside], gx[inside]), 1)
    # candidates per pixel = records filed under the (2 halo + 1)^2 cells around it
    c = np.pad(counts, ((1, 0), (1, 0))).cumsum(0).cumsum(1)
    span = 2 * halo + 1
    cand = c[span:span + H, span:span + W] - c[:H, span:span + W] - c[span:span + H, :W] + c[:H, :W]
    # exact in-radius hits per pixel
    tree = cKDTree(np.stack([u[inside], v[inside]], 1))
    yy, xx = np.mgrid[0:H, 0:W]
    centres = np.stack([xx.ravel() + 0.5, yy.ravel() + 0.5], 1)
    hits = tree.query_ball_point(centres, r_px, return_length=True).reshape(H, W)
    kept = np.minimum(hits, K)
    P = int(inside.sum())
    density = P / (H * W)
    row = (f"| {name} K={K} r={r} | {P} | {halo} | {span * span} | {cand.mean():.1f} | {cand.max()} | "
           f"{hits.mean():.1f} | {kept.mean():.2f} | {100 * (hits >= K).mean():.1f} % | {100 * (hits == 0).mean():.2f} % | "
           f"{100 * hits.sum() / max(cand.sum(), 1):.0f} % |")
    print("\n| config | points filed | halo | window cells | candidates / pixel (mean) | (max) | in-radius hits / pixel | "
          "kept / pixel | pixels with a full list | empty pixels | hit rate of the candidate test |")
    print(row)
    # every hit of a pixel is one of its candidates (that is what the halo guarantees) ...
    assert np.all(hits <= cand)
    # ... and the means are what the geometry predicts (interior density x window / disc area)
    assert abs(cand.mean() / (span * span * density) - 1) < 0.1
    assert abs(hits.mean() / (math.pi * r_px * r_px * density) - 1) < 0.1
