#!/usr/bin/env python
"""Generate golden fixtures by running the REAL reference code (read from /root/reference).

Runs only in the build container (where /root/reference exists); the resulting small
`.npz` fixtures are committed so that the tests never need the reference at run time.

What is real and what is stubbed
--------------------------------
* Real (imported unmodified from /root/reference/pgdvs): PGDVSBaseRenderer.get_batched_rays,
  PGDVSDynamicRenderer.compute_dyn_pcl / render_dyn_pcl / forward,
  PGDVSDynamicTrackRenderer.compute_pcl_for_tgt, Projector.compute_projections,
  utils.pytorch3d_utils.cameras_from_opencv_to_pytorch3d.
* Stubbed because the package is absent from this image and un-vendored by the reference:
  hydra, trimesh, cupy, skimage, jax/haiku (tracker NNs) and **pytorch3d**.  The pytorch3d stub
    - implements `ops.knn_points` by brute force in torch (squared L2, ascending) and
    - RECORDS what the reference passes across the pytorch3d boundary (camera R/t/K/image size,
      raster settings, point cloud, features), returning zeros as the rendered image.
  So the fixtures pin everything the reference computes in-tree up to that boundary; the
  pytorch3d arithmetic itself stays "parity unpinned" (see oracle/raster_cpu.cpp).

Usage:  python tests/golden/make_golden.py     (writes tests/golden/*.npz)
"""
import importlib
import pathlib
import sys
import types

import numpy as np
import torch

REF_ROOT = pathlib.Path("/root/reference")
OUT_DIR = pathlib.Path(__file__).resolve().parent


# ------------------------------------------------------------------------------ stubs
class _StubObj:
    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use
        return self

    def __getattr__(self, item):
        return _StubObj(self._name + "." + item)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _StubObj(self.__name__ + "." + item)


def _install_stub(name):
    mod = _StubModule(name)
    sys.modules[name] = mod
    return mod


BOUNDARY_LOG = []  # what the reference hands to pytorch3d


def _knn_points(p1, p2, K, return_nn=False, **kw):
    d2 = ((p1[0][:, None, :] - p2[0][None, :, :]) ** 2).sum(-1)
    k = min(K, p2.shape[1])
    val, idx = torch.topk(d2, k, dim=1, largest=False, sorted=True)
    nn = p2[0][idx] if return_nn else None
    return val[None], idx[None], (nn[None] if nn is not None else None)


def _install_pytorch3d_stub():
    p3d = _install_stub("pytorch3d")
    p3d.utils = _install_stub("pytorch3d.utils")
    p3d.ops = _install_stub("pytorch3d.ops")
    p3d.renderer = _install_stub("pytorch3d.renderer")
    p3d.structures = _install_stub("pytorch3d.structures")
    p3d.ops.knn_points = _knn_points

    def cameras_from_opencv_projection(R, tvec, camera_matrix, image_size):
        rec = {"R": R.clone(), "tvec": tvec.clone(), "camera_matrix": camera_matrix.clone(),
               "image_size": image_size.clone()}
        BOUNDARY_LOG.append(("camera", rec))
        return rec

    class PointsRasterizationSettings:
        def __init__(self, image_size=256, radius=0.01, points_per_pixel=8, bin_size=None,
                     max_points_per_bin=None):
            self.image_size, self.radius = image_size, radius
            self.points_per_pixel, self.bin_size = points_per_pixel, bin_size
            self.max_points_per_bin = max_points_per_bin

    class PointsRasterizer:
        def __init__(self, cameras=None, raster_settings=None):
            self.cameras, self.raster_settings = cameras, raster_settings

    class NormWeightedCompositor:
        def __init__(self, background_color=None):
            self.background_color = background_color

    class AlphaCompositor(NormWeightedCompositor):
        pass

    class Pointclouds:
        def __init__(self, points, features=None):
            self.points, self.features = points, features

    class PointsRenderer:
        def __init__(self, rasterizer, compositor):
            self.rasterizer, self.compositor = rasterizer, compositor

        def __call__(self, pcl):
            s = self.rasterizer.raster_settings
            BOUNDARY_LOG.append(("render", {
                "points": pcl.points.clone(), "features": pcl.features.clone(),
                "image_size": tuple(s.image_size), "radius": float(s.radius),
                "points_per_pixel": int(s.points_per_pixel), "bin_size": s.bin_size,
                "compositor": type(self.compositor).__name__,
                "background_color": tuple(self.compositor.background_color)}))
            h, w = s.image_size
            return torch.zeros(pcl.points.shape[0], h, w, pcl.features.shape[-1])

    class PerspectiveCameras:
        def __init__(self, **kw):
            self.kw = kw

    p3d.utils.cameras_from_opencv_projection = cameras_from_opencv_projection
    r = p3d.renderer
    r.PointsRasterizationSettings = PointsRasterizationSettings
    r.PointsRasterizer = PointsRasterizer
    r.PointsRenderer = PointsRenderer
    r.NormWeightedCompositor = NormWeightedCompositor
    r.AlphaCompositor = AlphaCompositor
    r.PerspectiveCameras = PerspectiveCameras
    p3d.structures.Pointclouds = Pointclouds
    return p3d


def import_reference():
    """Import the reference renderers with stubs for the absent third-party packages."""
    for name in ("hydra", "hydra.utils", "trimesh", "cupy", "skimage", "skimage.metrics", "omegaconf",
                 "pgdvs.models.tapnet.interface", "pgdvs.models.cotracker.interface"):
        _install_stub(name)
    _install_pytorch3d_stub()
    # the reference creates debug/ directories next to its sources at import time; the
    # reference tree is read-only, so make that a no-op for paths below it.
    real_mkdir = pathlib.Path.mkdir

    def mkdir(self, *a, **k):
        if str(self).startswith(str(REF_ROOT)):
            return None
        return real_mkdir(self, *a, **k)

    pathlib.Path.mkdir = mkdir
    sys.path.insert(0, str(REF_ROOT))
    mods = {
        "base": importlib.import_module("pgdvs.renderers.pgdvs_renderer_base"),
        "dyn": importlib.import_module("pgdvs.renderers.pgdvs_renderer_dyn"),
        "track": importlib.import_module("pgdvs.renderers.pgdvs_renderer_dyn_track"),
        "projector": importlib.import_module("pgdvs.models.gnt.projector"),
        "p3d_utils": importlib.import_module("pgdvs.utils.pytorch3d_utils"),
    }
    pathlib.Path.mkdir = real_mkdir
    return mods


# ----------------------------------------------------------------------- synthetic data
def make_scene(seed, H, W, n_frames=2, mask_frac=0.6, flow_std=2.0, zero_flow_frac=0.15):
    g = torch.Generator().manual_seed(seed)
    f = 0.9 * W
    K = torch.eye(4)
    K[0, 0] = f
    K[1, 1] = f * 1.05
    K[0, 2] = W / 2 - 0.7
    K[1, 2] = H / 2 + 0.4

    def pose(i):
        ang = 0.03 * (i - 1)
        c, s = np.cos(ang), np.sin(ang)
        c2w = torch.eye(4)
        c2w[:3, :3] = torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float32)
        c2w[:3, 3] = torch.tensor([0.1 * i, -0.05 * i, 0.02 * i])
        return c2w

    flat = []
    for i in range(n_frames + 1):  # last one is the target camera
        flat.append(torch.cat([torch.tensor([float(H), float(W)]), K.reshape(-1), pose(i).reshape(-1)]))
    flat = torch.stack(flat)
    rgb = torch.rand(n_frames, H, W, 3, generator=g)
    depth = 2.0 + 3.0 * torch.rand(n_frames, H, W, 1, generator=g)
    mask = (torch.rand(n_frames, H, W, 1, generator=g) < mask_frac).float()
    flow = flow_std * torch.randn(H, W, 2, generator=g)
    # exact-zero and exact-integer flows exercise grid_sample's half-pixel rounding
    z = torch.rand(H, W, 1, generator=g) < zero_flow_frac
    flow = torch.where(z, torch.round(flow), flow)
    occ = (torch.rand(H, W, 1, generator=g) < 0.2).float()
    return {"flat_cam_src": flat[:n_frames], "flat_cam_tgt": flat[n_frames], "rgb": rgb,
            "depth": depth, "mask": mask, "flow": flow, "occ": occ}


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def main():
    mods = import_reference()
    ns = types.SimpleNamespace
    out = {}

    # ---- 1. get_batched_rays (pgdvs_renderer_base.py:17-57)
    base = mods["base"].PGDVSBaseRenderer()
    sc = make_scene(1234, 12, 20)
    K1 = sc["flat_cam_src"][0, 2:18].reshape(1, 4, 4)
    c2w1 = sc["flat_cam_src"][0, 18:34].reshape(1, 4, 4)
    ro, rd, uvs, _, _ = base.get_batched_rays(device="cpu", batch_size=1, H=12, W=20,
                                              render_stride=1, intrinsics=K1, c2w=c2w1)
    np.savez_compressed(OUT_DIR / "rays.npz", flat_cam=_np(sc["flat_cam_src"][0]),
                        rays_o=_np(ro), rays_d=_np(rd), uvs=_np(uvs))

    # ---- 2. compute_projections (projector.py:41-73)
    proj = mods["projector"].Projector()
    g = torch.Generator().manual_seed(7)
    xyz = torch.randn(64, 3, generator=g) * 2 + torch.tensor([0.0, 0.0, 4.0])
    uv, m = proj.compute_projections(xyz[:, None, :], sc["flat_cam_tgt"][None])
    np.savez_compressed(OUT_DIR / "projections.npz", xyz=_np(xyz), flat_cam=_np(sc["flat_cam_tgt"]),
                        uv=_np(uv[0, :, 0, :]), mask=_np(m[0, :, 0]))

    # ---- 3. camera conversion (utils/pytorch3d_utils.py:5-47)
    c2w = sc["flat_cam_tgt"][18:34].reshape(4, 4)
    Kt = sc["flat_cam_tgt"][2:18].reshape(4, 4)
    w2c = torch.inverse(c2w)
    cams = mods["p3d_utils"].cameras_from_opencv_to_pytorch3d(
        w2c[None, :3, :3], w2c[None, :3, 3], Kt[None, :3, :3], torch.LongTensor([[12, 20]]))
    kw = cams.kw
    np.savez_compressed(OUT_DIR / "camera.npz", flat_cam=_np(sc["flat_cam_tgt"]),
                        R=_np(kw["R"][0]), T=_np(kw["T"][0]), focal=_np(kw["focal_length"][0]),
                        p0=_np(kw["principal_point"][0]))

    # ---- 4. compute_dyn_pcl + render_dyn_pcl boundary (pgdvs_renderer_dyn.py:275-540, 671-724)
    dyn = mods["dyn"].PGDVSDynamicRenderer(cfg=ns(rgb_range="0_1"),
                                           proj_func=proj.compute_projections)
    cases = []
    for ci, (H, W, t1, t2, tt, consist, rm, seed) in enumerate([
        (12, 20, 0.0, 1.0, 0.5, False, False, 11),
        (12, 20, 3.0, 4.0, 3.25, True, True, 12),
        (16, 16, 2.0, 2.0, 2.0, False, True, 13),   # time_1 == time_2 branch
        (20, 12, 5.0, 4.0, 4.5, True, False, 14),   # portrait, backward pair
    ]):
        sc = make_scene(seed, H, W)
        cfg = ns(dyn_render_type="pcl", dyn_render_use_flow_consistency=consist,
                 dyn_pcl_outlier_knn=8, dyn_pcl_outlier_std_thres=0.1, dyn_pcl_remove_outlier=rm,
                 dyn_render_pcl_pt_radius=0.05, dyn_render_pcl_pts_per_pixel=4)
        K1 = sc["flat_cam_src"][0, 2:18].reshape(1, 4, 4)
        c2w1 = sc["flat_cam_src"][0, 18:34].reshape(1, 4, 4)
        ro, rd, uvs, _, _ = dyn.get_batched_rays(device="cpu", batch_size=1, H=H, W=W,
                                                 render_stride=1, intrinsics=K1, c2w=c2w1)
        BOUNDARY_LOG.clear()
        flow_1_to_tgt, valid_mask, info = dyn.compute_dyn_pcl(
            dyn_mask_1=sc["mask"][0], rgb_1=sc["rgb"][0], uvs_1=uvs, ray_o_1=ro, ray_d_1=rd,
            depth_1=sc["depth"][0], flow_12=sc["flow"], flow_12_occ_mask=sc["occ"],
            rgb_2=sc["rgb"][1], depth_2=sc["depth"][1],
            K_2=sc["flat_cam_src"][1, 2:18].reshape(4, 4),
            c2w_2=sc["flat_cam_src"][1, 18:34].reshape(4, 4),
            flat_cam_tgt=sc["flat_cam_tgt"], time_1=torch.tensor(t1), time_2=torch.tensor(t2),
            time_tgt=torch.tensor(tt), render_cfg=cfg)
        cam_rec = [r for k, r in BOUNDARY_LOG if k == "camera"][0]
        rend = [r for k, r in BOUNDARY_LOG if k == "render"]
        assert len(rend) == 2 and rend[0]["bin_size"] == 0
        assert rend[0]["compositor"] == "NormWeightedCompositor"
        assert torch.all(rend[1]["features"] == 1)
        case = {
            "H": H, "W": W, "t1": t1, "t2": t2, "tt": tt, "consist": consist, "rm": rm,
            "flat_cam_src": _np(sc["flat_cam_src"]), "flat_cam_tgt": _np(sc["flat_cam_tgt"]),
            "rgb": _np(sc["rgb"]), "depth": _np(sc["depth"]), "mask": _np(sc["mask"]),
            "flow": _np(sc["flow"]), "occ": _np(sc["occ"]),
            "out_pcl": _np(info["pcl"]), "out_rgb": _np(info["pcl_rgbs"]),
            "out_thres": _np(info["pcl_nn_dist_thres"]), "out_valid_mask": _np(valid_mask),
            "out_flow_1_to_tgt": _np(flow_1_to_tgt),
            "b_R": _np(cam_rec["R"]), "b_tvec": _np(cam_rec["tvec"]),
            "b_K": _np(cam_rec["camera_matrix"]), "b_image_size": _np(cam_rec["image_size"]),
            "b_points": _np(rend[0]["points"]), "b_features": _np(rend[0]["features"]),
            "b_radius": rend[0]["radius"], "b_ppp": rend[0]["points_per_pixel"],
        }
        cases.append(case)
        np.savez_compressed(OUT_DIR / f"dyn_pcl_case{ci}.npz", **case)

    # ---- 5. track branch: compute_pcl_for_tgt (pgdvs_renderer_dyn_track.py:98-284), KNN filters
    # disabled by passing no base cloud and a huge threshold so only the geometry is pinned.
    trk = mods["track"].PGDVSDynamicTrackRenderer.__new__(mods["track"].PGDVSDynamicTrackRenderer)
    torch.nn.Module.__init__(trk)
    H, W, F, Q = 12, 20, 6, 80
    sc = make_scene(21, H, W, n_frames=F)
    g = torch.Generator().manual_seed(22)
    q_uv = torch.stack([torch.rand(Q, generator=g) * (W - 1), torch.rand(Q, generator=g) * (H - 1)], 1)
    tracks = q_uv[:, None, :] + torch.cumsum(torch.randn(Q, F, 2, generator=g), dim=1)
    tracks[:, :, 0].clamp_(0, W - 1)
    tracks[:, :, 1].clamp_(0, H - 1)
    visibles = torch.rand(Q, F, generator=g) < 0.7
    idx_closest = [2, 3]
    visibles[: Q // 2, 2] = False
    visibles[: Q // 2, 3] = False
    times = torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0, 5.0])
    time_tgt = torch.tensor(2.4)
    data_for_track = {
        "idx_temporal_closest": idx_closest, "idx_real_track": [0, 1, 4, 5],
        "time_for_track": times, "time_tgt": time_tgt, "rgbs_for_track": sc["rgb"],
        "depths_for_track": sc["depth"], "flat_cams_for_track": sc["flat_cam_src"],
    }
    cfg = ns(dyn_pcl_outlier_knn=4, dyn_pcl_outlier_std_thres=0.1,
             dyn_pcl_track_track2base_thres_mult=50)
    query_pts = torch.cat([torch.zeros(Q, 1), q_uv[:, 1:2], q_uv[:, 0:1]], dim=1)
    pcl_t, rgb_t = trk.compute_pcl_for_tgt(
        data_for_track=data_for_track, query_pts=query_pts, tracks=tracks,
        track_visibles=visibles, render_cfg=cfg,
        base_pcl_info={"pcl": None, "pcl_rgbs": None, "pcl_nn_dist_thres": torch.tensor(1e30)},
        device="cpu")
    np.savez_compressed(OUT_DIR / "track_pcl.npz", H=H, W=W, tracks=_np(tracks),
                        visibles=_np(visibles), rgbs=_np(sc["rgb"]), depths=_np(sc["depth"]),
                        flat_cams=_np(sc["flat_cam_src"]), times=_np(times), time_tgt=_np(time_tgt),
                        idx_temporal_closest=np.array(idx_closest), idx_real_track=np.array([0, 1, 4, 5]),
                        out_pcl=_np(pcl_t), out_rgb=_np(rgb_t))
    # ---- 6. softsplat importance metric (pgdvs_renderer_base.py:59-138).  The splat kernel itself
    #         is cupy-compiled CUDA and asserts on CPU tensors, so only the pure-torch part
    #         (back-warp + L1 metric) can be recorded from the real module.
    g = torch.Generator().manual_seed(11)
    Bs, Hs, Ws = 2, 14, 22
    rgb1 = torch.rand(Bs, 3, Hs, Ws, generator=g)
    rgb2 = torch.rand(Bs, 3, Hs, Ws, generator=g)
    flow12 = torch.randn(Bs, 2, Hs, Ws, generator=g) * 3.0
    warp = base.backwarp_for_softsplat_metric(tenIn=rgb2, tenFlow=flow12)
    metric = torch.nn.functional.l1_loss(input=rgb1, target=warp, reduction="none").mean(dim=1, keepdim=True)
    np.savez_compressed(OUT_DIR / "softsplat_metric.npz", rgb1=_np(rgb1), rgb2=_np(rgb2), flow12=_np(flow12),
                        warp=_np(warp), metric=_np(metric))
    print("wrote", sorted(p.name for p in OUT_DIR.glob("*.npz")))


if __name__ == "__main__":
    main()
