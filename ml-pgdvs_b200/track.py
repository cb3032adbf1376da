"""L2, the other callers of the point-splat kernels:

* the track branch of /root/reference/pgdvs/renderers/pgdvs_renderer_dyn_track.py
  (`compute_pcl_for_tgt` :98-396 and the render step of `render_with_track` :27-96).  Tracker
  inference (TAPIR / CoTracker, `run_track` :398-558) is out of scope: tracks and visibilities
  are inputs;
* `StaticGeoPointRenderer.forward` of /root/reference/pgdvs/renderers/st_geo_renderer.py:26-122.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _cabi, ops
from .dyn_renderer import PGDVSDynamicRenderer, _cfg, _fill, _np44, _plane, _upload_structs


def _mask_of(indices: Sequence[int]) -> int:
    m = 0
    for i in indices:
        m |= 1 << int(i)
    return m


def track_points(*, tracks, visibles, rgbs, depths, flat_cams, times, time_tgt,
                 idx_temporal_closest, idx_real_track, return_track_id: bool = False):
    """compute_pcl_for_tgt up to its KNN filters (pgdvs_renderer_dyn_track.py:98-284).

    tracks [Q,F,2] (col,row) f32, visibles [Q,F] bool, rgbs [F,H,W,3], depths [F,H,W,1],
    flat_cams [F,34], times [F], time_tgt scalar.  Returns (pcl [P,3], rgb [P,3][, track_id])."""
    ops._require_cuda(tracks, "tracks")
    dev = tracks.device
    Q, F, _ = tracks.shape
    _, H, W, _ = rgbs.shape
    if F > 32:
        raise ValueError("at most 32 frames per track window")
    tr = tracks.to(torch.float32).contiguous()
    vis = visibles.to(device=dev, dtype=torch.uint8).contiguous()
    fc = flat_cams.detach().cpu().numpy().astype(np.float32)
    tt = times.detach().cpu().numpy().astype(np.float32)
    keep, frames = [], []
    for f in range(F):
        Kf, c2w = _np44(fc[f, 2:18]), _np44(fc[f, 18:34])
        fr = _cabi.PgdvsTrackFrame()
        rgb_f, dep_f = _plane(rgbs[f]), _plane(depths[f])
        keep += [rgb_f, dep_f]
        fr.rgb, fr.depth = rgb_f.data_ptr(), dep_f.data_ptr()
        _fill(fr.M, c2w[:3, :3] @ np.linalg.inv(Kf[:3, :3]).astype(np.float32))
        _fill(fr.o, c2w[:3, 3])
        fr.time = float(tt[f])
        frames.append(fr)
    frames_dev = _upload_structs(frames, _cabi.PgdvsTrackFrame, dev)
    pcl = torch.empty((max(Q, 1), 3), dtype=torch.float32, device=dev)
    rgb = torch.empty((max(Q, 1), 3), dtype=torch.float32, device=dev)
    tid = torch.empty((max(Q, 1),), dtype=torch.int32, device=dev) if return_track_id else None
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    L = _cabi.lib()
    nbytes = ctypes.c_size_t(0)
    _cabi.check(L.pgdvs_track_workspace_bytes(Q, ctypes.byref(nbytes)), "pgdvs_track_workspace_bytes")
    ws = ops._WS.get(dev, nbytes.value, tag="track")
    with torch.cuda.device(dev):
        _cabi.check(L.pgdvs_track_points(
            tr.data_ptr(), vis.data_ptr(), Q, F, frames_dev.data_ptr(), _mask_of(idx_temporal_closest),
            _mask_of(idx_real_track), float(time_tgt), H, W, pcl.data_ptr(), rgb.data_ptr(),
            tid.data_ptr() if tid is not None else None, count.data_ptr(), ops._aligned_ptr(ws),
            nbytes.value, ops._stream_ptr(dev)), "pgdvs_track_points")
    ops.LAUNCHES["count"] += 3
    n = int(count.item())  # the reference syncs here as well (boolean-mask indexing, :129-137)
    del keep
    if return_track_id:
        return pcl[:n], rgb[:n], tid[:n].long()
    return pcl[:n], rgb[:n]


def _stat_outlier_threshold(avg, std_thres):
    return torch.median(avg) + torch.std(avg) * std_thres


def compute_pcl_for_tgt(*, tracks, visibles, rgbs, depths, flat_cams, times, time_tgt,
                        idx_temporal_closest, idx_real_track, render_cfg, base_pcl_info):
    """Full pgdvs_renderer_dyn_track.py:98-396: track cloud, track-to-base filter (:296-331),
    statistical self filter (:333-380), concatenation with the base cloud (:390-394)."""
    pcl, rgb = track_points(tracks=tracks, visibles=visibles, rgbs=rgbs, depths=depths, flat_cams=flat_cams,
                            times=times, time_tgt=time_tgt, idx_temporal_closest=idx_temporal_closest,
                            idx_real_track=idx_real_track)
    knn = int(_cfg(render_cfg, "dyn_pcl_outlier_knn"))
    base_pcl = base_pcl_info.get("pcl")
    base_thres = base_pcl_info.get("pcl_nn_dist_thres")
    if pcl.shape[0] > 0:
        if base_pcl is not None and base_pcl.shape[0] > 0:
            avg = ops.knn_mean_dist(pcl, base_pcl, knn + 1, skip_first=0)
            mult = float(getattr(render_cfg, "dyn_pcl_track_track2base_thres_mult", 50))
            flag = avg < base_thres * mult
            pcl, rgb = pcl[flag], rgb[flag]
        if pcl.shape[0] > 0:
            avg = ops.knn_mean_dist(pcl, pcl, knn + 1, skip_first=1)
            thres = base_thres if base_thres is not None else _stat_outlier_threshold(
                avg, float(_cfg(render_cfg, "dyn_pcl_outlier_std_thres")))
            flag = avg < thres
            pcl, rgb = pcl[flag], rgb[flag]
        if base_pcl is not None and pcl.shape[0] > 0:
            pcl = torch.cat((pcl, base_pcl), dim=0)
            rgb = torch.cat((rgb, base_pcl_info["pcl_rgbs"]), dim=0)
    return pcl, rgb


def render_with_track(*, tracks, visibles, rgbs, depths, flat_cams, times, time_tgt, idx_temporal_closest,
                      idx_real_track, flat_cam_tgt, render_cfg, base_pcl_info, H: int, W: int):
    """Steps 2-3 of render_with_track (pgdvs_renderer_dyn_track.py:54-81) for one target view:
    -> (track_rgb [H,W,3], track_mask [H,W,1])."""
    pcl, rgb = compute_pcl_for_tgt(tracks=tracks, visibles=visibles, rgbs=rgbs, depths=depths,
                                   flat_cams=flat_cams, times=times, time_tgt=time_tgt,
                                   idx_temporal_closest=idx_temporal_closest, idx_real_track=idx_real_track,
                                   render_cfg=render_cfg, base_pcl_info=base_pcl_info)
    r = PGDVSDynamicRenderer()
    return r.render_dyn_pcl(dyn_mask=torch.zeros(H, W, 1, device=tracks.device), dyn_pcl=pcl, rgbs=rgb,
                            flat_cam=flat_cam_tgt, render_cfg=render_cfg)


class StaticGeoPointRenderer(torch.nn.Module):
    """st_geo_renderer.py:20-122: the static scene as a point cloud, splatted with the same kernels
    (`st_render_pcl_pt_radius`, `st_render_pcl_pts_per_pixel`, optional statistical outlier removal)."""

    def __init__(self, model_cfg=None):
        super().__init__()

    def forward(self, *, tgt_h, tgt_w, flat_tgt_cam, st_pcl_rgb, render_cfg):
        assert st_pcl_rgb.ndim == 2, f"{st_pcl_rgb.shape}"
        dev = st_pcl_rgb.device
        st_pcl, st_rgb = st_pcl_rgb[:, :3].contiguous(), st_pcl_rgb[:, 3:].contiguous()
        if getattr(render_cfg, "st_pcl_remove_outlier", False) and st_pcl.shape[0] > 0:
            knn = int(getattr(render_cfg, "st_pcl_outlier_knn", 50))
            avg = ops.knn_mean_dist(st_pcl, st_pcl, knn + 1, skip_first=1)
            flag = avg < _stat_outlier_threshold(avg, float(getattr(render_cfg, "st_pcl_outlier_std_thres", 0.1)))
            st_pcl, st_rgb = st_pcl[flag], st_rgb[flag]
        if st_pcl.shape[0] == 0:
            return torch.zeros((tgt_h, tgt_w, 3), device=dev), torch.zeros((tgt_h, tgt_w, 1), device=dev)
        from types import SimpleNamespace
        cfg = SimpleNamespace(dyn_render_pcl_pt_radius=float(render_cfg.st_render_pcl_pt_radius),
                              dyn_render_pcl_pts_per_pixel=int(render_cfg.st_render_pcl_pts_per_pixel))
        r = PGDVSDynamicRenderer()
        return r.render_dyn_pcl(dyn_mask=torch.zeros(tgt_h, tgt_w, 1, device=dev), dyn_pcl=st_pcl, rgbs=st_rgb,
                                flat_cam=flat_tgt_cam, render_cfg=cfg)
