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
from .dyn_renderer import PGDVSDynamicRenderer, _cfg, _fill, _np44, _plane, _upload_structs  # noqa: F401


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


def track_knn_filters(pcl, rgb, base_pcl_info, render_cfg):
    """pgdvs_renderer_dyn_track.py:286-396: track-to-base distance filter (:296-331), statistical
    self filter (:333-380), concatenation with the base cloud (:390-394)."""
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


def compute_pcl_for_tgt(*, tracks, visibles, rgbs, depths, flat_cams, times, time_tgt,
                        idx_temporal_closest, idx_real_track, render_cfg, base_pcl_info):
    """Full pgdvs_renderer_dyn_track.py:98-396: track cloud, then the KNN filters."""
    pcl, rgb = track_points(tracks=tracks, visibles=visibles, rgbs=rgbs, depths=depths, flat_cams=flat_cams,
                            times=times, time_tgt=time_tgt, idx_temporal_closest=idx_temporal_closest,
                            idx_real_track=idx_real_track)
    return track_knn_filters(pcl, rgb, base_pcl_info, render_cfg)


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


def prepare_data(i_b, data, n_views, device=None):
    """PGDVSDynamicTrackRenderer.prepare_data (pgdvs_renderer_dyn_track.py:599-764): the frame
    window of batch item i_b in the order [fwd2tgt frames, temporally closest, bwd2tgt frames],
    times shifted to start at 0, images repeated to the fixed tracker window of `n_views` frames.
    Same keys as upstream.  (`n_actual_*` are read on the host, as upstream's slicing does.)"""
    rgbs, masks, depths, cams, times = [], [], [], [], []
    idx_real_track, idx_fwd, idx_bwd = [], [], []
    n_frames = 0

    def take(suffix, n):
        rgbs.append(data["rgb_src_temporal" + suffix][i_b, :n])
        masks.append(data["dyn_mask_src_temporal" + suffix][i_b, :n])
        depths.append(data["depth_src_temporal" + suffix][i_b, :n])
        cams.append(data["flat_cam_src_temporal" + suffix][i_b, :n])
        times.append(data["time_src_temporal" + suffix][i_b, :n])

    n_fwd = int(data["n_actual_temporal_track_fwd2tgt"][i_b, 0])
    if n_fwd > 0:
        take("_track_fwd2tgt", n_fwd)
        idx_fwd = list(range(n_fwd))
        idx_real_track.extend(idx_fwd)
        n_frames += n_fwd
    n_tmp = int(data["n_actual_temporal"][i_b, 0])
    idx_closest = [n_frames + i for i in range(n_tmp)]
    take("", n_tmp)
    n_frames += n_tmp
    n_bwd = int(data["n_actual_temporal_track_bwd2tgt"][i_b, 0])
    if n_bwd > 0:
        take("_track_bwd2tgt", n_bwd)
        idx_bwd = [n_frames + i for i in range(n_bwd)]
        idx_real_track.extend(idx_bwd)
    idx_full = idx_closest + idx_real_track
    assert len(idx_full) == len(set(idx_full)) and set(idx_full) == set(range(len(idx_full))), f"{idx_full}"
    rgbs, masks, depths = torch.cat(rgbs, 0), torch.cat(masks, 0), torch.cat(depths, 0)
    cams, times = torch.cat(cams, 0), torch.cat(times, 0)
    min_time = torch.min(times)
    times = times - min_time
    n_actual = rgbs.shape[0]
    n_rep = int(np.ceil(n_views / n_actual))
    return {
        "n_actual_frames": n_actual,
        "rgbs_for_track": rgbs.repeat(n_rep, 1, 1, 1)[:n_views],
        "dyn_masks_for_track": masks.repeat(n_rep, 1, 1, 1)[:n_views],
        "depths_for_track": depths,
        "flat_cams_for_track": cams,
        "time_for_track": times,
        "time_tgt": data["time_tgt"][i_b, :] - min_time,
        "idx_temporal_closest": idx_closest,
        "idx_real_track": idx_real_track,
        "idx_real_track_fwd": idx_fwd,
        "idx_real_track_bwd": idx_bwd,
        "time_real_track": times[idx_real_track],
    }


class PGDVSDynamicTrackRenderer(PGDVSDynamicRenderer):
    """Drop-in for pgdvs.renderers.pgdvs_renderer_dyn_track.PGDVSDynamicTrackRenderer
    (`dyn_render_track_temporal == "no_tgt"`, pgdvs_renderer.py:66-80).

    Tracker INFERENCE (TAPIR / CoTracker, `run_track` :398-558) is outside the hot-path scope
    (SURVEY.md §8): the tracker is injected — `tracker(data_for_track) -> (query_pts [Q,3] (t,row,col),
    tracks [Q,F,2] (col,row), visibles [Q,F] bool)` over the `n_actual_frames` frames of
    `prepare_data`'s window; `synthetic.SyntheticTracker` is the stand-in used by tests and bench.
    Everything downstream of the tracker runs here: frame selection, sampling, unprojection,
    time interpolation (csrc/track.cu), the KNN filters and one BATCHED splat of all views' track
    clouds (the reference renders them one by one)."""

    def __init__(self, *, cfg=None, softsplat_metric_abs_alpha=100.0, proj_func=None, local_rank=0,
                 use_tracker=True, tracker=None):
        super().__init__(cfg=cfg, softsplat_metric_abs_alpha=softsplat_metric_abs_alpha, proj_func=proj_func,
                         local_rank=local_rank, use_tracker=use_tracker, tracker=tracker)

    prepare_data = staticmethod(prepare_data)

    def run_track(self, data_for_track, for_debug=False, disable_tqdm=True):
        if self.tracker is None:
            raise NotImplementedError("no tracker was injected (tracker inference is outside the hot-path scope)")
        return self.tracker(data_for_track)

    def compute_pcl_for_tgt(self, *, data_for_track, query_pts, tracks, track_visibles, render_cfg,
                            base_pcl_info, device=None, for_debug=False):
        """pgdvs_renderer_dyn_track.py:98-396 with upstream's argument names."""
        n = data_for_track["n_actual_frames"]
        return compute_pcl_for_tgt(
            tracks=tracks[:, :n], visibles=track_visibles[:, :n], rgbs=data_for_track["rgbs_for_track"][:n],
            depths=data_for_track["depths_for_track"], flat_cams=data_for_track["flat_cams_for_track"],
            times=data_for_track["time_for_track"], time_tgt=float(data_for_track["time_tgt"].reshape(-1)[0]),
            idx_temporal_closest=data_for_track["idx_temporal_closest"],
            idx_real_track=data_for_track["idx_real_track"], render_cfg=render_cfg, base_pcl_info=base_pcl_info)

    def track_clouds(self, data, render_cfg, base_pcl_info, for_debug=False, disable_tqdm=True):
        """Steps 1-2 of render_with_track (:41-66) for every batch item: per view the world-space
        track cloud (+ base cloud) and its colours, or an empty cloud (:82-86)."""
        dev = data["rgb_src_temporal_track_fwd2tgt"].device
        n_b, n_one_side = data["rgb_src_temporal_track_fwd2tgt"].shape[:2]
        n_views = n_one_side * 2 + 2  # the last 2: the two temporally closest source views (:36)
        clouds = []
        for i_b in range(n_b):
            dft = self.prepare_data(i_b, data, n_views, dev)
            dyn = dft["dyn_masks_for_track"][dft["idx_real_track"]]
            if dyn.numel() > 0 and float(dyn.sum()) > 0:
                query_pts, tracks, vis = self.run_track(dft, for_debug=for_debug, disable_tqdm=disable_tqdm)
                cur = {"pcl": base_pcl_info["pcl"][i_b], "pcl_rgbs": base_pcl_info["pcl_rgbs"][i_b],
                       "pcl_nn_dist_thres": base_pcl_info["pcl_nn_dist_thres"][i_b]}
                clouds.append(self.compute_pcl_for_tgt(data_for_track=dft, query_pts=query_pts, tracks=tracks,
                                                       track_visibles=vis, render_cfg=render_cfg, base_pcl_info=cur,
                                                       device=dev, for_debug=for_debug))
            else:
                z = torch.zeros((0, 3), device=dev)
                clouds.append((z, z))
        return clouds

    def render_with_track(self, data, render_cfg, base_pcl_info, for_debug=False, disable_tqdm=True):
        """pgdvs_renderer_dyn_track.py:27-96 -> (track_rgbs [B,3,H,W], track_masks [B,1,H,W])."""
        n_b, _, H, W, _ = data["rgb_src_temporal"].shape
        dev = data["rgb_src_temporal"].device
        clouds = self.track_clouds(data, render_cfg, base_pcl_info, for_debug, disable_tqdm)
        img, mask = self.render_clouds_batched(clouds, data["flat_cam_tgt"], H, W, render_cfg, dev)
        return img.permute(0, 3, 1, 2).contiguous(), mask.permute(0, 3, 1, 2).contiguous()


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
