"""ORACLE — TEST INFRASTRUCTURE ONLY (torch-CPU restatement of the in-tree part of the path).

Restates, with torch CPU ops (so that `grid_sample`, `inverse`, `bmm` quirks are inherited
rather than re-derived), the PGDVS functions on the hot path:

  get_batched_rays          /root/reference/pgdvs/renderers/pgdvs_renderer_base.py:17-57
  compute_dyn_pcl           /root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:275-457
  render_dyn_pcl            /root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:671-724
  dyn/track merge           /root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:229-235
  static/dynamic blend      /root/reference/pgdvs/renderers/pgdvs_renderer.py:169-172
  camera conversion         /root/reference/pgdvs/utils/pytorch3d_utils.py:5-47
  compute_projections       /root/reference/pgdvs/models/gnt/projector.py:41-73
  track -> point cloud      /root/reference/pgdvs/renderers/pgdvs_renderer_dyn_track.py:98-284
  track KNN filters         /root/reference/pgdvs/renderers/pgdvs_renderer_dyn_track.py:286-396
  prepare_data              /root/reference/pgdvs/renderers/pgdvs_renderer_dyn_track.py:599-764
  resize_rgb_mask           /root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:259-270

and the pytorch3d 0.7.4 glue they call (PointsRasterizer.transform, PointsRenderer.forward,
knn_points) from the published algorithm (parity unpinned for those; see raster_cpu.cpp).

The in-tree functions ARE pinned: tests/golden/make_golden.py imports the real reference
modules (with stub packages for the absent third-party imports) and stores their outputs;
tests/test_oracle_golden.py checks this file against those fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import raster as _raster


# ----------------------------------------------------------------------------- cameras
def split_flat_cam(flat_cam):
    """flat_cam[34] = [h, w, K(4x4 row-major), c2w(4x4 row-major)] (datasets/nvidia_eval.py:827-832)."""
    flat_cam = flat_cam.reshape(-1)
    h, w = int(flat_cam[0]), int(flat_cam[1])
    K = flat_cam[2:18].reshape(4, 4)
    c2w = flat_cam[18:34].reshape(4, 4)
    return h, w, K, c2w


def get_batched_rays(H, W, K44, c2w44):
    """pgdvs_renderer_base.py:17-57 for batch_size=1, render_stride=1.

    Pixel grid u=col, v=row with NO half-pixel offset; rays_d = (c2w[:3,:3] @ K^-1) @ [u,v,1].
    """
    u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    u = u.reshape(-1).float()
    v = v.reshape(-1).float()
    pix = torch.stack((u, v, torch.ones_like(u)), dim=0)[None]  # [1,3,HW]
    K44 = K44.reshape(1, 4, 4)
    c2w44 = c2w44.reshape(1, 4, 4)
    rays_d = c2w44[:, :3, :3].bmm(torch.inverse(K44[:, :3, :3])).bmm(pix).transpose(1, 2)
    rays_o = c2w44[:, :3, 3].unsqueeze(1).repeat(1, rays_d.shape[1], 1)
    uvs = pix[:, :2, :].permute(0, 2, 1).reshape(-1, 2)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), uvs


def cameras_from_opencv_projection(R, tvec, camera_matrix, image_size_hw):
    """pytorch3d.utils.cameras_from_opencv_projection as restated in-tree at
    pgdvs/utils/pytorch3d_utils.py:5-47.  Returns a dict instead of PerspectiveCameras:
    R_p3d [3,3] (row-vector convention), T_p3d [3], focal [2], p0 [2] (NDC)."""
    R = R.reshape(1, 3, 3).float()
    tvec = tvec.reshape(1, 3).float()
    camera_matrix = camera_matrix.reshape(1, 3, 3).float()
    image_size = torch.as_tensor(image_size_hw).reshape(1, 2)
    focal_length = torch.stack([camera_matrix[:, 0, 0], camera_matrix[:, 1, 1]], dim=-1)
    principal_point = camera_matrix[:, :2, 2]
    image_size_wh = image_size.to(R).flip(dims=(1,))
    scale = image_size_wh.min(dim=1, keepdim=True)[0] / 2.0
    scale = scale.expand(-1, 2)
    c0 = image_size_wh / 2.0
    focal_p3d = focal_length / scale
    p0_p3d = -(principal_point - c0) / scale
    R_p3d = R.clone().permute(0, 2, 1).contiguous()
    T_p3d = tvec.clone()
    R_p3d[:, :, :2] *= -1
    T_p3d[:, :2] *= -1
    return {"R": R_p3d[0], "T": T_p3d[0], "focal": focal_p3d[0], "p0": p0_p3d[0]}


def world_to_ndc(points_world, cam):
    """PointsRasterizer.transform for PerspectiveCameras(in_ndc=True):
    view = [p,1] @ [[R,0],[T,1]]; proj = [view,1] @ K^T with
    K = [[fx,0,px,0],[0,fy,py,0],[0,0,0,1],[0,0,1,0]]; xy / w; z := view z."""
    P = points_world.shape[0]
    M = torch.eye(4)
    M[:3, :3] = cam["R"]
    M[3, :3] = cam["T"]
    ph = torch.cat([points_world.float(), torch.ones(P, 1)], dim=1)
    view = ph @ M
    view = view[:, :3] / view[:, 3:]
    Kp = torch.zeros(4, 4)
    Kp[0, 0], Kp[1, 1] = cam["focal"][0], cam["focal"][1]
    Kp[0, 2], Kp[1, 2] = cam["p0"][0], cam["p0"][1]
    Kp[3, 2] = 1.0
    Kp[2, 3] = 1.0
    vh = torch.cat([view, torch.ones(P, 1)], dim=1)
    proj = vh @ Kp.t().contiguous()
    ndc = proj[:, :3] / proj[:, 3:]
    ndc[:, 2] = view[:, 2]
    return ndc


def camera_from_flat_cam(flat_cam):
    """render_dyn_pcl's camera set-up, pgdvs_renderer_dyn.py:674-687."""
    h, w, K, c2w = split_flat_cam(flat_cam)
    w2c = torch.inverse(c2w)
    return cameras_from_opencv_projection(w2c[:3, :3], w2c[:3, 3], K[:3, :3], (h, w))


def compute_projections(xyz, flat_cam):
    """Projector.compute_projections (models/gnt/projector.py:41-73) for one camera.
    xyz [P,3] -> (uv [P,2] clamped to +-1e6, mask [P] = z>0)."""
    _, _, K, c2w = split_flat_cam(flat_cam)
    xyz_h = torch.cat([xyz, torch.ones_like(xyz[:, :1])], dim=-1)
    proj = K[None].bmm(torch.inverse(c2w[None])).bmm(xyz_h.t()[None])
    proj = proj.permute(0, 2, 1)[0]
    uv = proj[:, :2] / torch.clamp(proj[:, 2:3], min=1e-8)
    uv = torch.clamp(uv, min=-1e6, max=1e6)
    return uv, proj[:, 2] > 0


# ------------------------------------------------------------------ unproject/warp/lerp
def knn_outlier_flags(pcl, knn=50, std_thres=0.1, chunk=2048):
    """pgdvs_renderer_dyn.py:405-427: pytorch3d knn_points(K=knn+1) brute force (squared L2,
    ascending), drop the self match, avg; keep iff avg < median + std*thres (unbiased std)."""
    P = pcl.shape[0]
    k = min(knn + 1, P)
    avg = torch.empty(P)
    for s in range(0, P, chunk):
        q = pcl[s:s + chunk]
        d2 = ((q[:, None, :] - pcl[None, :, :]) ** 2).sum(-1)
        nn = torch.topk(d2, k, dim=1, largest=False, sorted=True).values
        avg[s:s + chunk] = nn[:, 1:].mean(dim=1)
    thres = torch.median(avg) + torch.std(avg) * std_thres
    return avg < thres, thres, avg


def compute_dyn_pcl(*, dyn_mask_1, rgb_1, depth_1, flow_12, flow_12_occ_mask, rgb_2, depth_2,
                    K_1, c2w_1, K_2, c2w_2, time_1, time_2, time_tgt,
                    use_flow_consistency=False, flag_not_outlier=None):
    """pgdvs_renderer_dyn.py:275-457 up to (and including) the outlier compaction.

    Inputs are the [H,W,C] channels-last maps of the reference's data dict.
    Returns dict(pcl [P,3] world, rgb [P,3], src_pix [P] (flat row-major source pixel index),
    n_masked, n_valid).  `flag_not_outlier` (bool [P_valid]) replaces the KNN statistics when
    given; None keeps every point (dyn_pcl_remove_outlier=False, :453-457).
    """
    H, W, _ = dyn_mask_1.shape
    raw_shape = torch.FloatTensor((W, H)).reshape(1, 2)
    rays_o, rays_d, uvs = get_batched_rays(H, W, K_1, c2w_1)

    flat_mask = dyn_mask_1.reshape(-1).bool()
    if use_flow_consistency:
        flat_mask = ~(flow_12_occ_mask > 0).reshape(-1) & flat_mask
    uv1 = uvs[flat_mask, :2]
    uv2 = uv1 + flow_12.reshape(-1, 2)[flat_mask, :]
    valid = torch.all((uv2 >= 0) & (uv2 <= raw_shape - 1), dim=1)

    pcl_1 = (rays_o + rays_d * depth_1.reshape(-1, 1))[flat_mask, :][valid, :]
    src_pix = torch.arange(H * W)[flat_mask][valid]

    if float(time_1) == float(time_2):
        pcl = pcl_1
        rgb = rgb_1.reshape(-1, 3)[flat_mask, :][valid, :]
    else:
        uv2 = uv2[valid, :]
        grid = 2 * uv2 / raw_shape - 1.0
        depth_w = torch.nn.functional.grid_sample(
            depth_2[None].permute(0, 3, 1, 2), grid[None, None], mode="nearest",
            align_corners=False)[0, 0, 0, :]
        rgb = torch.nn.functional.grid_sample(
            rgb_2[None].permute(0, 3, 1, 2), grid[None, None], mode="bilinear",
            align_corners=False)[0, :, 0, :].T
        uv2_h = torch.cat((uv2, torch.ones_like(uv2[:, :1])), dim=1)
        rays_d2 = torch.matmul(c2w_2[:3, :3], torch.matmul(torch.inverse(K_2[:3, :3]), uv2_h.T)).T
        rays_o2 = c2w_2[:3, 3][None, :].expand(rays_d2.shape[0], -1)
        pcl_2 = rays_o2 + rays_d2 * depth_w[:, None]
        w1 = (time_2 - time_tgt) / (time_2 - time_1)
        w2 = (time_tgt - time_1) / (time_2 - time_1)
        pcl = w1 * pcl_1 + w2 * pcl_2

    n_valid = int(pcl.shape[0])
    if flag_not_outlier is not None:
        pcl = pcl[flag_not_outlier, :]
        rgb = rgb[flag_not_outlier, :]
        src_pix = src_pix[flag_not_outlier]
    return {"pcl": pcl, "rgb": rgb, "src_pix": src_pix, "n_masked": int(flat_mask.sum()),
            "n_valid": n_valid}


# ------------------------------------------------------------------------------ render
def render_dyn_pcl(*, H, W, dyn_pcl, rgbs, flat_cam, radius, points_per_pixel,
                   compositor="norm", n_threads=1, banded=False, return_fragments=False):
    """pgdvs_renderer_dyn.py:671-724: camera conversion -> PointsRasterizer(bin_size=0) ->
    PointsRenderer(NormWeightedCompositor(bg 0)) twice (rgb features, then ones for the mask)."""
    if dyn_pcl.shape[0] == 0:
        img = torch.zeros(H, W, 3)
        mask = torch.zeros(H, W, 1)
        return (img, mask, None) if return_fragments else (img, mask)
    cam = camera_from_flat_cam(flat_cam)
    ndc = world_to_ndc(dyn_pcl, cam).numpy()
    P = ndc.shape[0]
    fi = np.zeros(1, dtype=np.int64)
    npc = np.full(1, P, dtype=np.int64)
    img, frags = _raster.render_points(ndc, fi, npc, rgbs.numpy(), (H, W), radius,
                                       points_per_pixel, compositor, background=(0, 0, 0),
                                       n_threads=n_threads, banded=banded)
    ones, _ = _raster.render_points(ndc, fi, npc, np.ones((P, 3), np.float32), (H, W), radius,
                                    points_per_pixel, compositor, background=(0, 0, 0),
                                    n_threads=n_threads, banded=banded)
    img_t = torch.from_numpy(img[0, :, :, :3])
    mask_t = torch.from_numpy((ones[0, :, :, :1] > 0.0).astype(np.float32))
    if return_fragments:
        return img_t, mask_t, (ndc,) + frags
    return img_t, mask_t


def merge_dyn_track(dyn_rgb, dyn_mask, track_rgb, track_mask):
    """pgdvs_renderer_dyn.py:229-235."""
    m = ((~(dyn_mask > 0)) & (track_mask > 0)).float()
    rgb = (1 - m) * dyn_rgb + m * track_rgb
    mask = ((dyn_mask > 0) | (track_mask > 0)).float()
    return rgb, mask


def blend_static_dynamic(static_rgb, dyn_rgb, dyn_mask):
    """pgdvs_renderer.py:169-172."""
    return (1 - dyn_mask) * static_rgb + dyn_mask * dyn_rgb


# ------------------------------------------------------------------------------ tracks
def compute_pcl_for_tgt(*, tracks, visibles, rgbs, depths, flat_cams, times, time_tgt,
                        idx_temporal_closest, idx_real_track):
    """pgdvs_renderer_dyn_track.py:98-284 (before the KNN filters).

    tracks [Q,F,2] (col,row), visibles [Q,F] bool, rgbs [F,H,W,3], depths [F,H,W,1],
    flat_cams [F,34], times [F], time_tgt scalar tensor.
    Returns (pcl [P,3], rgb [P,3], track_id [P])."""
    vis_closest = visibles[:, idx_temporal_closest]
    flag_invis = torch.all(~vis_closest, dim=1)
    flag_enough = torch.sum(visibles[:, idx_real_track].float(), dim=1) >= 2
    flag_valid = flag_invis & flag_enough
    n_valid = int(flag_valid.sum())
    if n_valid == 0:
        return torch.zeros(0, 3), torch.zeros(0, 3), torch.zeros(0, dtype=torch.long)
    v_tracks = tracks[flag_valid]
    v_vis = visibles[flag_valid]
    t_all = times[None, :].expand(n_valid, -1)
    t_diff = (t_all - time_tgt).masked_fill_(~v_vis, float("inf"))
    order = torch.argsort(t_diff.abs(), dim=1, descending=False)[:, :2]
    ar = torch.arange(n_valid)[:, None].expand(-1, 2)
    t_use = t_all[ar, order]
    uv_use = v_tracks[ar, order, :]  # [n_valid,2,2]
    _, H, W, _ = rgbs.shape
    wh = torch.FloatTensor((W, H))[None, :]
    pts = torch.zeros(n_valid, 2, 3)
    cols = torch.zeros(n_valid, 2, 3)
    for f in torch.unique(order).tolist():
        sel = order == f  # [n_valid,2]
        uv = uv_use[sel]  # [#,2]
        grid = 2 * uv / wh - 1
        c = torch.nn.functional.grid_sample(rgbs[f:f + 1].permute(0, 3, 1, 2), grid[None, None],
                                            mode="bilinear", align_corners=True)[0, :, 0, :].permute(1, 0)
        d = torch.nn.functional.grid_sample(depths[f:f + 1].permute(0, 3, 1, 2), grid[None, None],
                                            mode="nearest", align_corners=False)[0, 0, 0, :]
        Kf = flat_cams[f, 2:18].reshape(4, 4)
        c2w = flat_cams[f, 18:34].reshape(4, 4)
        uvh = torch.nn.functional.pad(uv, (0, 1), value=1).permute(1, 0)
        rd = (c2w[None, :3, :3].bmm(torch.inverse(Kf[None, :3, :3])).bmm(uvh[None])).transpose(1, 2)[0]
        ro = c2w[None, :3, 3].repeat(rd.shape[0], 1)
        pts[sel] = ro + rd * d[:, None]
        cols[sel] = c
    rgb = torch.mean(cols, dim=1)
    ratio = (time_tgt - t_use[:, :1]) / (t_use[:, 1:2] - t_use[:, :1] + 1e-8)
    pcl = pts[:, 0, :] + (pts[:, 1, :] - pts[:, 0, :]) * ratio
    return pcl, rgb, torch.nonzero(flag_valid)[:, 0]


def knn_mean_sq_dist(query, ref_pts, K, chunk=2048):
    """pytorch3d.ops.knn_points(query, ref, K) -> mean over all K squared distances (brute force).
    When the reference cloud has fewer than K points pytorch3d pads the missing columns with 0 and
    torch.mean still divides by K."""
    Q, R = query.shape[0], ref_pts.shape[0]
    k = min(K, R)
    out = torch.empty(Q)
    for s in range(0, Q, chunk):
        q = query[s:s + chunk]
        d2 = ((q[:, None, :] - ref_pts[None, :, :]) ** 2).sum(-1)
        nn = torch.topk(d2, k, dim=1, largest=False, sorted=True).values
        out[s:s + chunk] = nn.sum(dim=1) / K
    return out


def track_knn_filters(pcl_track, pcl_track_rgbs, base_pcl_info, *, knn=50, std_thres=0.1, track2base_mult=50.0):
    """pgdvs_renderer_dyn_track.py:286-396: drop track points far from the base cloud
    (mean of the knn+1 squared distances to it < base_thres * mult), then the statistical self
    filter (mean of the knn nearest OTHER points < base_thres, or median + std*thres when the base
    threshold is None), then append the base cloud."""
    base_pcl, base_thres = base_pcl_info.get("pcl"), base_pcl_info.get("pcl_nn_dist_thres")
    if pcl_track.shape[0] > 0:
        if base_pcl is not None and base_pcl.shape[0] > 0:
            avg = knn_mean_sq_dist(pcl_track, base_pcl, knn + 1)
            flag = avg < base_thres * track2base_mult
            pcl_track, pcl_track_rgbs = pcl_track[flag], pcl_track_rgbs[flag]
        if pcl_track.shape[0] > 0:
            P = pcl_track.shape[0]
            k = min(knn + 1, P)
            d2 = ((pcl_track[:, None, :] - pcl_track[None, :, :]) ** 2).sum(-1)
            nn = torch.topk(d2, k, dim=1, largest=False, sorted=True).values
            avg = nn[:, 1:].sum(dim=1) / knn  # torch.mean over the knn columns (zero-padded if P <= knn)
            thres = base_thres if base_thres is not None else torch.median(avg) + torch.std(avg) * std_thres
            flag = avg < thres
            pcl_track, pcl_track_rgbs = pcl_track[flag], pcl_track_rgbs[flag]
        if base_pcl is not None and pcl_track.shape[0] > 0:
            pcl_track = torch.cat((pcl_track, base_pcl), dim=0)
            pcl_track_rgbs = torch.cat((pcl_track_rgbs, base_pcl_info["pcl_rgbs"]), dim=0)
    return pcl_track, pcl_track_rgbs


def prepare_data(i_b, data, n_views):
    """PGDVSDynamicTrackRenderer.prepare_data (pgdvs_renderer_dyn_track.py:599-764): assemble the
    frame window of batch item i_b in the order [fwd2tgt frames, temporally closest, bwd2tgt frames]."""
    rgbs, masks, depths, cams, times = [], [], [], [], []
    idx_real_track, idx_fwd, idx_bwd = [], [], []
    n_frames = 0
    n_fwd = int(data["n_actual_temporal_track_fwd2tgt"][i_b, 0])
    if n_fwd > 0:
        rgbs.append(data["rgb_src_temporal_track_fwd2tgt"][i_b, :n_fwd])
        masks.append(data["dyn_mask_src_temporal_track_fwd2tgt"][i_b, :n_fwd])
        depths.append(data["depth_src_temporal_track_fwd2tgt"][i_b, :n_fwd])
        cams.append(data["flat_cam_src_temporal_track_fwd2tgt"][i_b, :n_fwd])
        times.append(data["time_src_temporal_track_fwd2tgt"][i_b, :n_fwd])
        idx_fwd = list(range(n_fwd))
        idx_real_track.extend(idx_fwd)
        n_frames += n_fwd
    n_tmp = int(data["n_actual_temporal"][i_b, 0])
    idx_closest = [n_frames + i for i in range(n_tmp)]
    rgbs.append(data["rgb_src_temporal"][i_b, :n_tmp])
    depths.append(data["depth_src_temporal"][i_b, :n_tmp])
    cams.append(data["flat_cam_src_temporal"][i_b, :n_tmp])
    masks.append(data["dyn_mask_src_temporal"][i_b, :n_tmp])
    times.append(data["time_src_temporal"][i_b, :n_tmp])
    n_frames += n_tmp
    n_bwd = int(data["n_actual_temporal_track_bwd2tgt"][i_b, 0])
    if n_bwd > 0:
        rgbs.append(data["rgb_src_temporal_track_bwd2tgt"][i_b, :n_bwd])
        masks.append(data["dyn_mask_src_temporal_track_bwd2tgt"][i_b, :n_bwd])
        depths.append(data["depth_src_temporal_track_bwd2tgt"][i_b, :n_bwd])
        cams.append(data["flat_cam_src_temporal_track_bwd2tgt"][i_b, :n_bwd])
        times.append(data["time_src_temporal_track_bwd2tgt"][i_b, :n_bwd])
        idx_bwd = [n_frames + i for i in range(n_bwd)]
        idx_real_track.extend(idx_bwd)
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
        "depths_for_track": depths, "flat_cams_for_track": cams, "time_for_track": times,
        "time_tgt": data["time_tgt"][i_b, :] - min_time,
        "idx_temporal_closest": idx_closest, "idx_real_track": idx_real_track,
        "idx_real_track_fwd": idx_fwd, "idx_real_track_bwd": idx_bwd,
        "time_real_track": times[idx_real_track],
    }


def resize_rgb_mask(rgb, mask, render_h, render_w):
    """pgdvs_renderer_dyn.py:259-270."""
    rgb = torch.nn.functional.interpolate(rgb, size=(render_h, render_w), mode="bicubic", align_corners=True,
                                          antialias=True)
    mask = torch.nn.functional.interpolate(mask, size=(render_h, render_w), mode="nearest")
    return rgb, mask


# ----------------------------------------------------------------------------- softmax splatting
# /root/reference/pgdvs/utils/softsplat.py:280-427 (the forward kernel is a cupy-compiled CUDA
# string upstream and asserts on CPU tensors, so it cannot be executed here: PARITY UNPINNED for
# softsplat_forward, restated from the in-tree kernel text) and
# /root/reference/pgdvs/renderers/pgdvs_renderer_base.py:59-138 (pure torch: pinned by
# tests/golden/softsplat_metric.npz, generated from the real module).
def softsplat_forward(ten_in, ten_flow):
    """softsplat_func.forward (softsplat.py:342-420): tenIn [N,C,H,W], tenFlow [N,2,H,W]."""
    ten_in = np.asarray(ten_in, np.float32)
    ten_flow = np.asarray(ten_flow, np.float32)
    N, C, H, W = ten_in.shape
    out = np.zeros_like(ten_in)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    for n in range(N):
        fx = (xs + ten_flow[n, 0]).astype(np.float32)
        fy = (ys + ten_flow[n, 1]).astype(np.float32)
        ok = np.isfinite(fx) & np.isfinite(fy)
        fx0, fy0 = np.where(ok, fx, 0).astype(np.float32), np.where(ok, fy, 0).astype(np.float32)
        nwx, nwy = np.floor(fx0), np.floor(fy0)
        sex, sey = nwx + 1, nwy + 1
        taps = [
            (nwx, nwy, (sex - fx0) * (sey - fy0)),   # north-west
            (sex, nwy, (fx0 - nwx) * (sey - fy0)),   # north-east
            (nwx, sey, (sex - fx0) * (fy0 - nwy)),   # south-west
            (sex, sey, (fx0 - nwx) * (fy0 - nwy)),   # south-east
        ]
        for tx, ty, w in taps:
            m = ok & (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
            txi, tyi = tx[m].astype(np.int64), ty[m].astype(np.int64)
            wv = w[m].astype(np.float32)
            for c in range(C):
                np.add.at(out[n, c], (tyi, txi), (ten_in[n, c][m] * wv).astype(np.float32))
    return out


def softsplat(ten_in, ten_flow, ten_metric, mode):
    """softsplat.softsplat (softsplat.py:280-334) on torch CPU tensors."""
    base = mode.split("-")[0]
    assert base in ("sum", "avg", "linear", "soft")
    if base == "avg":
        ten_in = torch.cat([ten_in, ten_in.new_ones([ten_in.shape[0], 1, ten_in.shape[2], ten_in.shape[3]])], 1)
    elif base == "linear":
        ten_in = torch.cat([ten_in * ten_metric, ten_metric], 1)
    elif base == "soft":
        ten_in = torch.cat([ten_in * ten_metric.exp(), ten_metric.exp()], 1)
    out = torch.from_numpy(softsplat_forward(ten_in.numpy(), ten_flow.numpy()))
    if base in ("avg", "linear", "soft"):
        norm = out[:, -1:, :, :]
        parts = mode.split("-")
        if len(parts) == 1 or parts[1] == "addeps":
            norm = norm + 0.0000001
        elif parts[1] == "zeroeps":
            norm = norm.clone()
            norm[norm == 0.0] = 1.0
        elif parts[1] == "clipeps":
            norm = norm.clip(0.0000001, None)
        out = out[:, :-1, :, :] / norm
    return out


def backwarp_for_softsplat_metric(ten_in, ten_flow):
    """pgdvs_renderer_base.py:100-138."""
    H, W = ten_flow.shape[2], ten_flow.shape[3]
    hor = torch.linspace(-1.0, 1.0, W, dtype=ten_flow.dtype).view(1, 1, 1, -1).repeat(1, 1, H, 1)
    ver = torch.linspace(-1.0, 1.0, H, dtype=ten_flow.dtype).view(1, 1, -1, 1).repeat(1, 1, 1, W)
    grid = torch.cat([hor, ver], 1)
    flow = torch.cat([ten_flow[:, 0:1] / ((ten_in.shape[3] - 1.0) / 2.0),
                      ten_flow[:, 1:2] / ((ten_in.shape[2] - 1.0) / 2.0)], 1)
    return torch.nn.functional.grid_sample(input=ten_in, grid=(grid + flow).permute(0, 2, 3, 1), mode="bilinear",
                                           padding_mode="zeros", align_corners=True)


def softsplat_img(*, rgb_src1, flow_src1_to_tgt, rgb_src2=None, flow_src1_to_src2=None, metric=None, alpha=100.0):
    """PGDVSBaseRenderer.softsplat_img (pgdvs_renderer_base.py:59-98)."""
    if metric is None:
        warp = backwarp_for_softsplat_metric(rgb_src2, flow_src1_to_src2)
        metric = torch.nn.functional.l1_loss(input=rgb_src1, target=warp, reduction="none").mean(dim=1, keepdim=True)
    out = softsplat(rgb_src1, flow_src1_to_tgt, (-alpha * metric).clip(-alpha, alpha), "soft")
    return out, metric


def softsplat_dyn_render(*, rgb_1, dyn_mask_1, rgb_2, flow_1_to_tgt, flow_12, noise=None, alpha=100.0):
    """The softsplat branch of PGDVSDynamicRenderer.forward (pgdvs_renderer_dyn.py:157-209).
    Channels-last inputs [B,H,W,C]; `noise` stands for clamp(randn_like(rgb), 0, 1) (zeros if None).
    Returns (render_dyn_rgb [B,3,H,W], render_dyn_mask [B,1,H,W], metric [B,1,H,W])."""
    m = dyn_mask_1.permute(0, 3, 1, 2)
    f_t = flow_1_to_tgt.permute(0, 3, 1, 2)
    r2 = rgb_2.permute(0, 3, 1, 2)
    f12 = flow_12.permute(0, 3, 1, 2)
    r1 = rgb_1.permute(0, 3, 1, 2)
    nz = noise.permute(0, 3, 1, 2) if noise is not None else torch.zeros_like(r1)
    r1 = r1 * m + nz * (1 - m)
    splat, metric = softsplat_img(rgb_src1=r1, flow_src1_to_tgt=f_t, rgb_src2=r2, flow_src1_to_src2=f12, alpha=alpha)
    mask, _ = softsplat_img(rgb_src1=m, flow_src1_to_tgt=f_t, rgb_src2=r2, flow_src1_to_src2=f12, metric=metric,
                            alpha=alpha)
    mask = (mask > 1e-3).float()
    return splat * mask, mask, metric


# ----------------------------------------------------------------------------- mesh mode
def mesh_faces_from_mask(rows, cols, H, W):
    """Grid-topology faces of render_dyn_mesh (pgdvs_renderer_dyn.py:549-604): every valid pixel
    (row, col) spawns (r,c),(r+1,c),(r+1,c+1) and (r,c),(r+1,c+1),(r,c+1); a face survives if its
    three corners are in bounds and carry a vertex index > 0 (sic: vertex 0 never gets a face)."""
    vert_idx = -torch.ones((H, W), dtype=torch.long)
    vert_idx[rows, cols] = torch.arange(rows.shape[0])
    c1 = torch.stack([torch.stack((rows, cols), 1), torch.stack((rows + 1, cols), 1),
                      torch.stack((rows + 1, cols + 1), 1)], 1)
    c2 = torch.stack([torch.stack((rows, cols), 1), torch.stack((rows + 1, cols + 1), 1),
                      torch.stack((rows, cols + 1), 1)], 1)
    cand = torch.cat((c1, c2), 0)
    inb = torch.all((cand[..., 0] >= 0) & (cand[..., 0] < H) & (cand[..., 1] >= 0) & (cand[..., 1] < W), dim=1)
    cand = cand[inb]
    fv = vert_idx[cand[..., 0], cand[..., 1]]
    return fv[torch.all(fv > 0, dim=1)]


def render_dyn_mesh(*, rows, cols, dyn_mask, dyn_pcl, rgbs, flat_cam):
    """PGDVSDynamicRenderer.render_dyn_mesh (pgdvs_renderer_dyn.py:542-669): MeshRasterizer
    (blur 0, 1 face per pixel, naive) + SimpleShader (vertex colours interpolated with the
    perspective-correct barycentrics, hard blend on a black background); the mask is the same
    render with all-ones vertex colours, > 0.  Returns (img [H,W,3], mask [H,W,1], fragments)."""
    H, W, _ = dyn_mask.shape
    faces = mesh_faces_from_mask(rows, cols, H, W)
    if faces.shape[0] == 0:
        return torch.zeros(H, W, 3), torch.zeros(H, W, 1), None
    ndc = world_to_ndc(dyn_pcl, camera_from_flat_cam(flat_cam))
    fv = ndc[faces]  # [F,3,3]
    p2f, zbuf, bary, _ = _raster.rasterize_meshes(fv.numpy(), (H, W), 1, 0.0, True)
    p2f_t = torch.from_numpy(p2f[..., 0]).long()
    bary_t = torch.from_numpy(bary[..., 0, :])
    hit = p2f_t >= 0
    fcol = rgbs[faces]  # [F,3,3] colours of the face corners
    idx = p2f_t.clamp(min=0)
    # interpolate_face_attributes: sum_i bary_i * attr_i ; background pixels -> 0
    img = (bary_t[..., :, None] * fcol[idx]).sum(dim=-2)
    ones = bary_t.sum(dim=-1, keepdim=True)
    img = torch.where(hit[..., None], img, torch.zeros_like(img))
    mask = (torch.where(hit[..., None], ones, torch.zeros_like(ones)) > 0.0).float()
    return img, mask, (p2f[..., 0], zbuf[..., 0], bary[..., 0, :], faces)
