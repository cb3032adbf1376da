"""GPU parity of the track branch:

* `track.compute_pcl_for_tgt` — the track kernel followed by BOTH KNN filters and the
  concatenation with the base cloud (pgdvs_renderer_dyn_track.py:98-396) — against
  oracle/pgdvs_ref.compute_pcl_for_tgt + track_knn_filters;
* `PGDVSDynamicTrackRenderer.forward` on a reference-shaped data dict (prepare_data :599-764,
  render_with_track :27-96, dyn/track merge pgdvs_renderer_dyn.py:229-235) against the oracle
  pipeline, every differing pixel attributed to a boundary flip or a z tie (tests/_attrib.py).

Tolerances: survivors of the filters identical (count and order), positions rtol/atol 2e-4 (the two
frames nearest in time can be picked in either order on |dt| ties), colours 5e-6; images 1e-4 on
every pixel that is not attributed."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import pgdvs_ref as ref
from _attrib import attribute_mismatches

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _flat(H, W, tx, ty=0.0):
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([tx, ty, 0.0])
    return torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), c2w.reshape(-1)])


def _track_case(seed, H=36, W=52, F=8, Q=6000):
    gen = torch.Generator().manual_seed(seed)
    rgbs = torch.rand(F, H, W, 3, generator=gen)
    # smooth depth: neighbouring track points land near each other and near the base cloud
    depths = 3 + 0.5 * torch.rand(F, 1, 1, 1, generator=gen) + 0.05 * torch.rand(F, H, W, 1, generator=gen)
    flat = torch.stack([_flat(H, W, 0.03 * f, -0.01 * f) for f in range(F)])
    uv0 = torch.stack([torch.rand(Q, generator=gen) * (W - 1), torch.rand(Q, generator=gen) * (H - 1)], 1)
    tracks = uv0[:, None, :] + torch.cumsum(0.7 * torch.randn(Q, F, 2, generator=gen), dim=1)
    visibles = torch.rand(Q, F, generator=gen) < 0.6
    times = torch.arange(F, dtype=torch.float32)
    kw = dict(tracks=tracks, visibles=visibles, rgbs=rgbs, depths=depths, flat_cams=flat, times=times,
              time_tgt=torch.tensor(3.3), idx_temporal_closest=[3, 4], idx_real_track=[0, 1, 2, 5, 6, 7])
    # base cloud: frame 3 warped towards frame 4 (pgdvs_renderer_dyn.py:275-457)
    flow = 0.8 * torch.randn(H, W, 2, generator=gen)
    base = ref.compute_dyn_pcl(
        dyn_mask_1=(torch.rand(H, W, 1, generator=gen) < 0.7).float(), rgb_1=rgbs[3], depth_1=depths[3], flow_12=flow,
        flow_12_occ_mask=torch.zeros(H, W, 1), rgb_2=rgbs[4], depth_2=depths[4], K_1=flat[3, 2:18].reshape(4, 4),
        c2w_1=flat[3, 18:34].reshape(4, 4), K_2=flat[4, 2:18].reshape(4, 4), c2w_2=flat[4, 18:34].reshape(4, 4),
        time_1=torch.tensor(3.0), time_2=torch.tensor(4.0), time_tgt=torch.tensor(3.3))
    return kw, base


@pytest.mark.parametrize("with_base_thres", [True, False])
def test_compute_pcl_for_tgt_with_knn_filters_vs_oracle(with_base_thres):
    from pgdvs_b200 import track
    d = _dev()
    knn = 8
    kw, base = _track_case(11)
    _, thres, _ = ref.knn_outlier_flags(base["pcl"], knn=knn)
    cfg = SimpleNamespace(dyn_pcl_outlier_knn=knn, dyn_pcl_outlier_std_thres=0.1, dyn_pcl_track_track2base_thres_mult=50)
    if with_base_thres:
        info_cpu = {"pcl": base["pcl"], "pcl_rgbs": base["rgb"], "pcl_nn_dist_thres": thres}
        info_gpu = {k: v.to(d) for k, v in info_cpu.items()}
    else:  # no base cloud: only the statistical self filter with its own median + std threshold (:363-366)
        info_cpu = {"pcl": None, "pcl_rgbs": None, "pcl_nn_dist_thres": None}
        info_gpu = dict(info_cpu)
    e_pcl, e_rgb, _ = ref.compute_pcl_for_tgt(**kw)
    n_track = e_pcl.shape[0]
    e_pcl, e_rgb = ref.track_knn_filters(e_pcl, e_rgb, info_cpu, knn=knn, std_thres=0.1, track2base_mult=50.0)
    pcl, rgb = track.compute_pcl_for_tgt(**{k: (v.to(d) if torch.is_tensor(v) and k != "time_tgt" else v)
                                            for k, v in kw.items()}, render_cfg=cfg, base_pcl_info=info_gpu)
    n_base = base["pcl"].shape[0] if with_base_thres else 0
    assert n_track > 300 and 0 < e_pcl.shape[0] - n_base < n_track  # both filters removed something, not everything
    assert pcl.shape == e_pcl.shape and rgb.shape == e_rgb.shape
    np.testing.assert_allclose(pcl.cpu().numpy(), e_pcl.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(rgb.cpu().numpy(), e_rgb.numpy(), atol=5e-6, rtol=0)


@pytest.mark.parametrize("remove_outlier", [False, True])
def test_forward_with_tracker_vs_oracle(remove_outlier):
    from pgdvs_b200 import synthetic, track
    d = _dev()
    name = "tiny_track"
    cfgd = synthetic.CONFIGS[name]
    H, W, K, radius, B = cfgd["H"], cfgd["W"], cfgd["K"], cfgd["radius"], cfgd["views"]
    knn = 8
    data = synthetic.make_data_dict(name, torch.device("cpu"), n_views=B, n_track_one_side=3)
    cfg = SimpleNamespace(dyn_render_type="pcl", dyn_render_pcl_pt_radius=radius, dyn_render_pcl_pts_per_pixel=K,
                          dyn_render_use_flow_consistency=False, dyn_pcl_remove_outlier=remove_outlier,
                          dyn_pcl_outlier_knn=knn, dyn_pcl_outlier_std_thres=0.1, dyn_pcl_track_track2base_thres_mult=50)
    recorded = []
    inner = synthetic.SyntheticTracker(seed=7, p_visible=0.5)  # (0.5: enough tracks unseen by both closest frames)

    def tracker(dft):
        out = inner(dft)
        recorded.append(tuple(t.cpu() for t in out))
        return out

    r = track.PGDVSDynamicTrackRenderer(tracker=tracker)
    rgb, mask, info = r({k: v.to(d) for k, v in data.items()}, None, cfg)
    assert rgb.shape == (B, 3, H, W) and mask.shape == (B, 1, H, W) and len(recorded) == B
    n_views = 3 * 2 + 2
    fs = data["flat_cam_src_temporal"]
    n_track_px = 0
    for b in range(B):
        o = ref.compute_dyn_pcl(
            dyn_mask_1=data["dyn_mask_src_temporal"][b, 0], rgb_1=data["rgb_src_temporal"][b, 0],
            depth_1=data["depth_src_temporal"][b, 0], flow_12=data["flow_fwd"][b],
            flow_12_occ_mask=data["flow_fwd_occ_mask"][b], rgb_2=data["rgb_src_temporal"][b, 1],
            depth_2=data["depth_src_temporal"][b, 1], K_1=fs[b, 0, 2:18].reshape(4, 4), c2w_1=fs[b, 0, 18:34].reshape(4, 4),
            K_2=fs[b, 1, 2:18].reshape(4, 4), c2w_2=fs[b, 1, 18:34].reshape(4, 4), time_1=data["time_src_temporal"][b, 0],
            time_2=data["time_src_temporal"][b, 1], time_tgt=data["time_tgt"][b, 0])
        flags, thres, _ = ref.knn_outlier_flags(o["pcl"], knn=knn)
        base_pcl, base_rgb = (o["pcl"][flags], o["rgb"][flags]) if remove_outlier else (o["pcl"], o["rgb"])
        d_img, d_mask = ref.render_dyn_pcl(H=H, W=W, dyn_pcl=base_pcl, rgbs=base_rgb, flat_cam=data["flat_cam_tgt"][b],
                                           radius=radius, points_per_pixel=K)
        dft = ref.prepare_data(b, data, n_views)
        n = dft["n_actual_frames"]
        q, tr, vis = recorded[b]
        t_pcl, t_rgb, _ = ref.compute_pcl_for_tgt(
            tracks=tr[:, :n], visibles=vis[:, :n], rgbs=dft["rgbs_for_track"][:n], depths=dft["depths_for_track"],
            flat_cams=dft["flat_cams_for_track"], times=dft["time_for_track"], time_tgt=dft["time_tgt"][0],
            idx_temporal_closest=dft["idx_temporal_closest"], idx_real_track=dft["idx_real_track"])
        t_pcl, t_rgb = ref.track_knn_filters(t_pcl, t_rgb, {"pcl": base_pcl, "pcl_rgbs": base_rgb, "pcl_nn_dist_thres": thres},
                                             knn=knn, std_thres=0.1, track2base_mult=50.0)
        t_img, t_mask = ref.render_dyn_pcl(H=H, W=W, dyn_pcl=t_pcl, rgbs=t_rgb, flat_cam=data["flat_cam_tgt"][b],
                                           radius=radius, points_per_pixel=K)
        e_rgb, e_mask = ref.merge_dyn_track(d_img, d_mask, t_img, t_mask)
        n_track_px += int(((d_mask == 0) & (t_mask > 0)).sum())
        # every differing pixel must be a boundary flip or a z tie of the clouds that feed it
        cam = ref.camera_from_flat_cam(data["flat_cam_tgt"][b])
        clouds = [c for c in (base_pcl, t_pcl) if c.shape[0] > 0]
        ndc = ref.world_to_ndc(torch.cat(clouds, 0), cam).numpy() if clouds else np.zeros((0, 3), np.float32)
        g_rgb, g_mask = rgb[b].cpu().permute(1, 2, 0), mask[b, 0].cpu()
        bad = ((g_rgb - e_rgb).abs().max(dim=-1).values > 1e-4) | (g_mask != e_mask[..., 0])
        for name_i, g_i, e_i in (("temporal_closest", info["temporal_closest_mask"][b, 0].cpu(), d_mask[..., 0]),
                                 ("temporal_track", info["temporal_track_mask"][b, 0].cpu(), t_mask[..., 0])):
            bad |= g_i != e_i
        counts = attribute_mismatches(bad.numpy(), ndc, H, W, radius, K)
        assert bad.float().mean() < 0.02, (b, counts)
    assert n_track_px > 0  # the track branch really contributed pixels the closest frames do not cover
