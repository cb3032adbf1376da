"""Synthetic workloads of the five BASELINE.json configs.

Shapes, dtypes and layouts follow the reference's data dict
(/root/reference/pgdvs/datasets/nvidia_eval.py:545-604): channels-last fp32 maps, flow in
pixels (+u right / +v down), OpenCV K / c2w, frame ids as float times.  Data generation is
not part of the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .dyn_renderer import SourcePair

CONFIGS: Dict[str, dict] = {
    # name: H, W, frames, sources per view (S), K, radius, default #views
    "c1_nvidia_1view": dict(H=288, W=544, frames=2, S=2, K=8, radius=0.01, views=1),
    "c2_nvidia_seq": dict(H=288, W=544, frames=12, S=2, K=8, radius=0.01, views=144),
    "c3_iphone": dict(H=360, W=480, frames=8, S=6, K=16, radius=0.01, views=16),
    "c4_davis": dict(H=480, W=854, frames=80, S=2, K=8, radius=0.01, views=80),
    "c5_stress": dict(H=1080, W=1920, frames=9, S=8, K=8, radius=0.01, views=8),
    "tiny": dict(H=24, W=40, frames=3, S=2, K=4, radius=0.06, views=3),
    "tiny_track": dict(H=32, W=48, frames=8, S=2, K=4, radius=0.05, views=4),
}


@dataclass
class Scene:
    H: int
    W: int
    rgb: torch.Tensor        # [F,H,W,3]
    depth: torch.Tensor      # [F,H,W,1]
    mask: torch.Tensor       # [F,H,W,1]
    flow_next: torch.Tensor  # [F,H,W,2] flow frame f -> f+1 (last: -> f-1)
    flow_prev: torch.Tensor  # [F,H,W,2] flow frame f -> f-1 (first: -> f+1)
    K: np.ndarray            # [4,4]
    c2w: np.ndarray          # [F,4,4]
    times: np.ndarray        # [F]


@dataclass
class Workload:
    name: str
    H: int
    W: int
    K: int
    radius: float
    scene: Scene
    view_pairs: List[List[SourcePair]]          # per view: S source pairs (view index = 0 in each)
    view_cams: List[Tuple[np.ndarray, np.ndarray]]  # per view (K44, c2w44)
    static_rgb: Optional[torch.Tensor] = None   # [V,H,W,3] stand-in for the GNT static render
    meta: dict = field(default_factory=dict)

    @property
    def n_views(self):
        return len(self.view_pairs)

    def jobs(self, views: Sequence[int]):
        """Flatten the selected views into (pairs sorted by local view index, cams)."""
        pairs, cams = [], []
        for local, v in enumerate(views):
            for p in self.view_pairs[v]:
                q = SourcePair.__new__(SourcePair)
                q.__dict__.update(p.__dict__)
                q.view = local
                pairs.append(q)
            cams.append(self.view_cams[v])
        return pairs, cams

    def points_per_view(self):
        return len(self.view_pairs[0]) * self.H * self.W


def _box_smooth(x: torch.Tensor, k: int = 9) -> torch.Tensor:
    # x [F,H,W,1] -> 9x9 box filter (plausible surfaces keep z-ties / KNN realistic)
    y = x.permute(0, 3, 1, 2)
    y = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(y, (k // 2,) * 4, mode="replicate"), k, stride=1)
    return y.permute(0, 2, 3, 1).contiguous()


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)


def make_scene(H, W, frames, device, seed=1234, mask_mode="full", flow_std=3.0, smooth=True,
               flow_mode="smooth") -> Scene:
    g = torch.Generator(device=device).manual_seed(seed)
    rgb = torch.rand((frames, H, W, 3), generator=g, device=device)
    depth = 1.0 + 9.0 * torch.rand((frames, H, W, 1), generator=g, device=device)
    if smooth:
        depth = _box_smooth(depth)
    if mask_mode == "full":
        mask = torch.ones((frames, H, W, 1), device=device)
    elif mask_mode == "ellipse":  # centred ellipse covering ~15 % of the pixels
        yy, xx = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
        a, b = 0.31 * W / 2 * 1.4, 0.31 * H / 2 * 1.4
        m = (((xx - W / 2) / a) ** 2 + ((yy - H / 2) / b) ** 2) < 1.0
        mask = m.float()[None, :, :, None].repeat(frames, 1, 1, 1)
    else:
        raise ValueError(mask_mode)

    def flow():
        f = torch.randn((frames, H, W, 2), generator=g, device=device)
        if flow_mode == "smooth":
            # piecewise-smooth motion like real optical flow: low-pass the noise with the same 9x9
            # box used for depth, rescale to the requested std, add a little sub-pixel jitter
            f = _box_smooth(f.reshape(frames, H, W, 2).permute(0, 3, 1, 2).reshape(frames * 2, H, W, 1))
            f = f.reshape(frames, 2, H, W).permute(0, 2, 3, 1)
            f = f / f.std().clamp_min(1e-6)
            f = flow_std * f + 0.1 * torch.randn((frames, H, W, 2), generator=g, device=device)
        elif flow_mode == "iid":
            f = flow_std * f  # stress case: every pixel moves independently (incoherent gathers)
        else:
            raise ValueError(flow_mode)
        uu = torch.arange(W, device=device, dtype=torch.float32)[None, None, :]
        vv = torch.arange(H, device=device, dtype=torch.float32)[None, :, None]
        # clip so that uv + flow stays inside the image (most points survive the validity test)
        f[..., 0] = torch.minimum(torch.maximum(f[..., 0], -uu), (W - 1) - uu)
        f[..., 1] = torch.minimum(torch.maximum(f[..., 1], -vv), (H - 1) - vv)
        return f.contiguous()

    Kmat = np.eye(4, dtype=np.float32)
    Kmat[0, 0] = Kmat[1, 1] = 0.9 * W
    Kmat[0, 2], Kmat[1, 2] = W / 2.0, H / 2.0
    c2w = np.tile(np.eye(4, dtype=np.float32), (frames, 1, 1))
    for f in range(frames):
        c2w[f, :3, :3] = _rot_y(0.004 * (f - frames / 2))
        c2w[f, :3, 3] = np.array([0.05 * (f - frames / 2), 0.01 * math.sin(f), 0.0], np.float32)
    times = np.arange(frames, dtype=np.float32)
    return Scene(H, W, rgb, depth, mask, flow(), flow(), Kmat, c2w, times)


def _target_cam(scene: Scene, t: float, k: int, n_cams: int):
    """novel camera: small rotation (<= 3 deg) + translation on an ellipse around the
    interpolated source pose (cf. create_bt_poses, datasets/nvidia_vis.py:692-722)."""
    f0 = int(min(max(math.floor(t), 0), scene.c2w.shape[0] - 1))
    c2w = scene.c2w[f0].copy()
    ang = 2 * math.pi * k / max(n_cams, 1)
    c2w[:3, :3] = c2w[:3, :3] @ _rot_y(math.radians(2.0) * math.cos(ang)) @ _rot_x(math.radians(1.5) * math.sin(ang))
    c2w[:3, 3] += np.array([0.15 * math.cos(ang), 0.08 * math.sin(ang), 0.05 * math.sin(2 * ang)], np.float32)
    return scene.K.copy(), c2w


def _pair(scene: Scene, a: int, b: int, t_tgt: float) -> SourcePair:
    flow = scene.flow_next[a] if b == a + 1 else scene.flow_prev[a]
    sp = SourcePair(depth_1=scene.depth[a], rgb_1=scene.rgb[a], mask_1=scene.mask[a], flow_12=flow,
                    depth_2=scene.depth[b], rgb_2=scene.rgb[b], K_1=scene.K, c2w_1=scene.c2w[a],
                    K_2=scene.K, c2w_2=scene.c2w[b], time_1=scene.times[a], time_2=scene.times[b],
                    time_tgt=t_tgt, view=0)
    sp._src_frames, sp._t_tgt = (a, b), float(t_tgt)  # provenance, used by the CPU reference arm
    return sp


def make_workload(name: str, device, n_views: Optional[int] = None, seed: int = 1234,
                  K: Optional[int] = None, radius: Optional[float] = None, mask_mode: str = "full",
                  with_static: bool = True, flow_mode: str = "smooth") -> Workload:
    cfg = dict(CONFIGS[name])
    H, W, F, S = cfg["H"], cfg["W"], cfg["frames"], cfg["S"]
    V = n_views if n_views is not None else cfg["views"]
    K = K if K is not None else cfg["K"]
    radius = radius if radius is not None else cfg["radius"]
    scene = make_scene(H, W, F, device, seed=seed, mask_mode=mask_mode, flow_mode=flow_mode)
    view_pairs, view_cams = [], []
    n_cams = 12 if name.startswith("c2") else max(V, 1)
    for v in range(V):
        if name.startswith("c2"):
            step, cam_k = v // 12, v % 12           # 12 time steps x 12 cameras (N_CAMS, nvidia_eval.py:53)
        else:
            step, cam_k = v, v
        a = step % max(F - 1, 1)                     # temporally closest source frames a, a+1
        b = a + 1
        t_tgt = float(scene.times[a]) + 0.5 if not name.startswith("c4") else float(scene.times[a]) + (v % 7 + 1) / 8.0
        pairs = []
        # S source segments: the two closest frames in both directions first, then +-2, +-3 ...
        order = []
        d = 0
        while len(order) < S:
            lo, hi = a - d, b + d
            if lo >= 0:
                order.append((lo, lo + 1))           # forward pair  lo -> lo+1
            if len(order) < S and hi < F:
                order.append((hi, hi - 1))           # backward pair hi -> hi-1
            if lo < 0 and hi >= F:
                order.append((a, b))                 # short scene: repeat (keeps P = S*H*W)
            d += 1
        for (i, j) in order[:S]:
            pairs.append(_pair(scene, i, j, t_tgt))
        view_pairs.append(pairs)
        view_cams.append(_target_cam(scene, t_tgt, cam_k, n_cams))
    static = None
    if with_static:
        g = torch.Generator(device=device).manual_seed(seed + 1)
        static = torch.rand((V, H, W, 3), generator=g, device=device)
    return Workload(name, H, W, K, radius, scene, view_pairs, view_cams, static,
                    meta=dict(S=S, frames=F, mask_mode=mask_mode, seed=seed, flow_mode=flow_mode))


# ------------------------------------------------------------------ reference-shaped `data` dicts
def _flat_cam(H, W, K44, c2w44) -> torch.Tensor:
    """[h, w, K(4x4 row-major), c2w(4x4 row-major)] (datasets/nvidia_eval.py:827-832)."""
    return torch.from_numpy(np.concatenate([np.array([H, W], np.float32), np.asarray(K44, np.float32).reshape(-1),
                                            np.asarray(c2w44, np.float32).reshape(-1)]))


def make_data_dict(name: str, device, n_views: Optional[int] = None, seed: int = 1234, n_track_one_side: int = 3,
                   track_mask_mode: str = "ellipse", scene: Optional[Scene] = None,
                   closest_mask_mode: str = "full") -> Dict[str, torch.Tensor]:
    """The `data` dict PGDVSDynamic(Track)Renderer.forward consumes, with the reference's keys,
    shapes and dtypes (datasets/nvidia_eval.py:545-604): per batch item the two temporally closest
    source frames (+ forward / backward flow between them) and, for the track branch, up to
    `n_track_one_side` older (`*_track_fwd2tgt`) and newer (`*_track_bwd2tgt`) frames in ascending
    time, padded with the nearest closest frame (nvidia_eval.py:281-317).  The closest frames carry
    the config's full dynamic mask (worst case P = H*W, like the batched workloads); the track
    frames carry `track_mask_mode` (a centred ellipse, ~15 % of the pixels, by default) because
    every dynamic pixel of a track frame becomes a query point (pgdvs_renderer_dyn_track.py:482-489)."""
    cfg = dict(CONFIGS[name])
    H, W, F = cfg["H"], cfg["W"], cfg["frames"]
    B = n_views if n_views is not None else cfg["views"]
    n = n_track_one_side
    sc = scene if scene is not None else make_scene(H, W, F, device, seed=seed)
    def mask_of(mode):
        return sc.mask if mode == "full" else make_scene(H, W, 1, device, seed=seed, mask_mode=mode).mask.expand(F, -1, -1, -1)
    tmask, cmask = mask_of(track_mask_mode), mask_of(closest_mask_mode)
    keys = ("rgb", "depth", "dyn_mask", "flat_cam", "time")
    out = {f"{k}_src_temporal{s}": [] for k in keys for s in ("", "_track_fwd2tgt", "_track_bwd2tgt")}
    extra = {k: [] for k in ("flow_fwd", "flow_bwd", "flow_fwd_occ_mask", "flow_bwd_occ_mask", "flat_cam_tgt", "time_tgt",
                             "n_actual_temporal", "n_actual_temporal_track_fwd2tgt", "n_actual_temporal_track_bwd2tgt")}
    fcs = [_flat_cam(H, W, sc.K, sc.c2w[f]) for f in range(F)]

    def frames(ids, suffix, mask):
        out["rgb_src_temporal" + suffix].append(torch.stack([sc.rgb[i] for i in ids]))
        out["depth_src_temporal" + suffix].append(torch.stack([sc.depth[i] for i in ids]))
        out["dyn_mask_src_temporal" + suffix].append(torch.stack([mask[i] for i in ids]))
        out["flat_cam_src_temporal" + suffix].append(torch.stack([fcs[i] for i in ids]))
        out["time_src_temporal" + suffix].append(torch.tensor([float(sc.times[i]) for i in ids]))

    for v in range(B):
        a = v % max(F - 1, 1)
        b = a + 1
        t_tgt = float(sc.times[a]) + 0.5
        frames([a, b], "", cmask)
        older = list(range(max(0, a - n), a))
        newer = list(range(b + 1, min(F, b + 1 + n)))
        frames(older + [a] * (n - len(older)), "_track_fwd2tgt", tmask)
        frames(newer + [b] * (n - len(newer)), "_track_bwd2tgt", tmask)
        extra["n_actual_temporal"].append(torch.tensor([2]))
        extra["n_actual_temporal_track_fwd2tgt"].append(torch.tensor([len(older)]))
        extra["n_actual_temporal_track_bwd2tgt"].append(torch.tensor([len(newer)]))
        extra["flow_fwd"].append(sc.flow_next[a])
        extra["flow_bwd"].append(sc.flow_prev[b])
        extra["flow_fwd_occ_mask"].append(torch.zeros((H, W, 1), device=sc.rgb.device))
        extra["flow_bwd_occ_mask"].append(torch.zeros((H, W, 1), device=sc.rgb.device))
        Kt, c2wt = _target_cam(sc, t_tgt, v, max(B, 1))
        extra["flat_cam_tgt"].append(_flat_cam(H, W, Kt, c2wt))
        extra["time_tgt"].append(torch.tensor([t_tgt]))
    data = {k: torch.stack(v).to(device) for k, v in {**out, **extra}.items()}
    return data


class SyntheticTracker:
    """Stand-in for TAPIR / CoTracker inference (outside the hot-path scope): maps
    prepare_data()'s frame window to (query_pts [Q,3] (t,row,col), tracks [Q,F,2] (col,row),
    visibles [Q,F] bool) exactly as run_track_func does for its queries — every dynamic pixel of
    every real-track frame is a query (pgdvs_renderer_dyn_track.py:482-489) — with
    tracks = query position + a cumulative N(0, step_px) walk away from the query frame and
    visibles ~ Bernoulli(p_visible) (always visible at the query frame), SURVEY.md §8(d)."""

    def __init__(self, seed: int = 1234, step_px: float = 1.0, p_visible: float = 0.8):
        self.seed, self.step_px, self.p_visible = seed, step_px, p_visible
        self.calls = 0

    def __call__(self, data_for_track):
        masks = data_for_track["dyn_masks_for_track"]
        dev = masks.device
        n_f = int(data_for_track["n_actual_frames"])
        H, W = masks.shape[1], masks.shape[2]
        qs = []
        for idx in data_for_track["idx_real_track"]:
            rows, cols = torch.nonzero(masks[idx, ..., 0] > 0.0, as_tuple=True)
            qs.append(torch.stack((torch.full_like(rows, idx), rows, cols), dim=1).float())
        query = torch.cat(qs, dim=0) if qs else torch.zeros((0, 3), device=dev)
        Q = query.shape[0]
        g = torch.Generator(device=dev).manual_seed(self.seed + self.calls)
        self.calls += 1
        steps = self.step_px * torch.randn((Q, n_f, 2), generator=g, device=dev)
        walk = torch.cumsum(steps, dim=1)
        qf = query[:, 0].long()
        walk = walk - walk[torch.arange(Q, device=dev), qf][:, None, :]  # zero displacement at the query frame
        tracks = torch.stack((query[:, 2], query[:, 1]), dim=1)[:, None, :] + walk  # (col, row)
        tracks[..., 0].clamp_(0, W - 1)
        tracks[..., 1].clamp_(0, H - 1)
        vis = torch.rand((Q, n_f), generator=g, device=dev) < self.p_visible
        vis[torch.arange(Q, device=dev), qf] = True
        return query, tracks, vis
