"""GPU parity for the stages around the rasterizer: fused unproject->warp->project, the
stand-alone compositors, merge/blend, KNN statistics and the PGDVS-shaped L2 API, against the
torch-CPU oracle (oracle/pgdvs_ref.py) and the golden fixtures generated from the real
reference code."""
import numpy as np
import pytest
import torch

from oracle import pgdvs_ref as ref
from _attrib import attribute_mismatches
from oracle import raster as oracle

pytestmark = pytest.mark.gpu
T = torch.from_numpy

PTS_RTOL, PTS_ATOL = 1e-5, 2e-5   # world / NDC coordinates (fp32 matmul association differs)
RGB_ATOL = 2e-6                   # bilinear colour
IMG_ATOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _golden_pair(g, dev, use_occ):
    from pgdvs_b200.dyn_renderer import SourcePair
    fs = g["flat_cam_src"]
    return SourcePair(
        depth_1=T(g["depth"][0]).to(dev), rgb_1=T(g["rgb"][0]).to(dev), mask_1=T(g["mask"][0]).to(dev),
        flow_12=T(g["flow"]).to(dev), depth_2=T(g["depth"][1]).to(dev), rgb_2=T(g["rgb"][1]).to(dev),
        K_1=fs[0, 2:18], c2w_1=fs[0, 18:34], K_2=fs[1, 2:18], c2w_2=fs[1, 18:34],
        time_1=float(g["t1"]), time_2=float(g["t2"]), time_tgt=float(g["tt"]), view=0,
        occ_12=T(g["occ"]).to(dev) if use_occ else None)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_uwp_matches_reference_golden(golden_dir, case):
    """The fused kernel vs what the REAL reference computed (fixtures) for compute_dyn_pcl's
    geometry: same survivors in the same order, world points / colours within tolerance."""
    from pgdvs_b200.dyn_renderer import opencv_to_p3d_camera, unproject_warp_project
    g = np.load(golden_dir / f"dyn_pcl_case{case}.npz")
    H, W = int(g["H"]), int(g["W"])
    dev = _dev()
    pair = _golden_pair(g, dev, bool(g["consist"]))
    ft = g["flat_cam_tgt"]
    cam = opencv_to_p3d_camera(ft[2:18], ft[18:34], H, W)
    cloud = unproject_warp_project([pair], [cam], H, W, dev, want_world=True, want_src_pix=True)
    torch.cuda.synchronize()
    P = int(cloud["total"].item())
    assert int(cloud["num_points"][0]) == P and int(cloud["first_idx"][0]) == 0
    # oracle (pinned to the fixtures by tests/test_oracle_golden.py) without outlier removal
    fs = T(g["flat_cam_src"])
    o = ref.compute_dyn_pcl(
        dyn_mask_1=T(g["mask"][0]), rgb_1=T(g["rgb"][0]), depth_1=T(g["depth"][0]), flow_12=T(g["flow"]),
        flow_12_occ_mask=T(g["occ"]), rgb_2=T(g["rgb"][1]), depth_2=T(g["depth"][1]),
        K_1=fs[0, 2:18].reshape(4, 4), c2w_1=fs[0, 18:34].reshape(4, 4), K_2=fs[1, 2:18].reshape(4, 4),
        c2w_2=fs[1, 18:34].reshape(4, 4), time_1=torch.tensor(float(g["t1"])),
        time_2=torch.tensor(float(g["t2"])), time_tgt=torch.tensor(float(g["tt"])),
        use_flow_consistency=bool(g["consist"]))
    assert P == o["pcl"].shape[0]
    assert np.array_equal(cloud["src_pix"][:P].cpu().numpy(), o["src_pix"].numpy())  # exact order
    np.testing.assert_allclose(cloud["xyz_world"][:P].cpu().numpy(), o["pcl"].numpy(), rtol=PTS_RTOL, atol=PTS_ATOL)
    np.testing.assert_allclose(cloud["rgb"][:P].cpu().numpy(), o["rgb"].numpy(), atol=RGB_ATOL, rtol=0)
    ndc = ref.world_to_ndc(o["pcl"], ref.camera_from_flat_cam(T(ft)))
    np.testing.assert_allclose(cloud["xyz_ndc"][:P].cpu().numpy(), ndc.numpy(), rtol=PTS_RTOL, atol=PTS_ATOL)
    if not bool(g["rm"]):
        # no outlier removal in this fixture: the reference's own cloud must match directly
        np.testing.assert_allclose(cloud["xyz_world"][:P].cpu().numpy(), g["out_pcl"], rtol=PTS_RTOL, atol=PTS_ATOL)
        np.testing.assert_allclose(cloud["rgb"][:P].cpu().numpy(), g["out_rgb"], atol=RGB_ATOL, rtol=0)


def test_uwp_batched_views_order_and_nearest_sampling():
    """Several jobs per view, several views, odd image size (scalar-load path), integer flows
    (grid_sample's half-pixel / round-half-even behaviour must match torch exactly)."""
    from pgdvs_b200.dyn_renderer import SourcePair, opencv_to_p3d_camera, unproject_warp_project
    dev = _dev()
    H, W = 13, 19  # H*W odd -> no float4 path
    gen = torch.Generator().manual_seed(5)
    F = 3
    rgb = torch.rand(F, H, W, 3, generator=gen)
    depth = 1 + 4 * torch.rand(F, H, W, 1, generator=gen)
    mask = (torch.rand(F, H, W, 1, generator=gen) < 0.7).float()
    flow = torch.round(2 * torch.randn(F, H, W, 2, generator=gen))  # integers: exact .5 sampling points
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2
    c2w = [torch.eye(4) for _ in range(F + 1)]
    for i in range(F + 1):
        c2w[i][:3, 3] = torch.tensor([0.05 * i, 0.02 * i, 0.0])
    jobs, exp = [], []
    spec = [(0, 1, 0.5, 0), (1, 0, 0.5, 0), (1, 2, 1.25, 1), (2, 2, 2.0, 2)]  # (a, b, t, view); last: same time
    for (a, b, tt, v) in spec:
        t1, t2 = float(a), float(b)
        jobs.append(SourcePair(depth_1=depth[a].to(dev), rgb_1=rgb[a].to(dev), mask_1=mask[a].to(dev),
                               flow_12=flow[a].to(dev), depth_2=depth[b].to(dev), rgb_2=rgb[b].to(dev),
                               K_1=Kc, c2w_1=c2w[a], K_2=Kc, c2w_2=c2w[b], time_1=t1, time_2=t2,
                               time_tgt=tt, view=v))
        exp.append(ref.compute_dyn_pcl(dyn_mask_1=mask[a], rgb_1=rgb[a], depth_1=depth[a], flow_12=flow[a],
                                       flow_12_occ_mask=torch.zeros(H, W, 1), rgb_2=rgb[b], depth_2=depth[b],
                                       K_1=Kc, c2w_1=c2w[a], K_2=Kc, c2w_2=c2w[b], time_1=torch.tensor(t1),
                                       time_2=torch.tensor(t2), time_tgt=torch.tensor(tt)))
    cams = [opencv_to_p3d_camera(Kc, c2w[F], H, W)] * 3
    cloud = unproject_warp_project(jobs, cams, H, W, dev, want_world=True, want_src_pix=True)
    torch.cuda.synchronize()
    n = [e["pcl"].shape[0] for e in exp]
    assert cloud["num_points"].tolist() == [n[0] + n[1], n[2], n[3]]
    assert cloud["first_idx"].tolist() == [0, n[0] + n[1], n[0] + n[1] + n[2]]
    P = sum(n)
    assert int(cloud["total"]) == P
    pcl = torch.cat([e["pcl"] for e in exp]).numpy()
    col = torch.cat([e["rgb"] for e in exp]).numpy()
    sp = torch.cat([e["src_pix"] for e in exp]).numpy()
    assert np.array_equal(cloud["src_pix"][:P].cpu().numpy(), sp)
    np.testing.assert_allclose(cloud["xyz_world"][:P].cpu().numpy(), pcl, rtol=PTS_RTOL, atol=PTS_ATOL)
    np.testing.assert_allclose(cloud["rgb"][:P].cpu().numpy(), col, atol=RGB_ATOL, rtol=0)


def test_uwp_large_matches_oracle_counts():
    """Config-1-shaped image with the float4 path and >1 scan tile per job: survivor count and
    order vs the oracle, geometry within tolerance."""
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import opencv_to_p3d_camera, unproject_warp_project
    dev = _dev()
    wl = synthetic.make_workload("c1_nvidia_1view", dev, mask_mode="ellipse")
    pairs, cams = wl.jobs([0])
    p3d = [opencv_to_p3d_camera(K, c, wl.H, wl.W) for (K, c) in cams]
    cloud = unproject_warp_project(pairs, p3d, wl.H, wl.W, dev, want_world=True, want_src_pix=True)
    torch.cuda.synchronize()
    sc = wl.scene
    tot = 0
    for j, (a, b) in enumerate([(0, 1), (1, 0)]):
        fl = sc.flow_next[a] if b == a + 1 else sc.flow_prev[a]
        e = ref.compute_dyn_pcl(dyn_mask_1=sc.mask[a].cpu(), rgb_1=sc.rgb[a].cpu(), depth_1=sc.depth[a].cpu(),
                                flow_12=fl.cpu(), flow_12_occ_mask=torch.zeros(wl.H, wl.W, 1),
                                rgb_2=sc.rgb[b].cpu(), depth_2=sc.depth[b].cpu(), K_1=T(sc.K), c2w_1=T(sc.c2w[a]),
                                K_2=T(sc.K), c2w_2=T(sc.c2w[b]), time_1=torch.tensor(sc.times[a]),
                                time_2=torch.tensor(sc.times[b]), time_tgt=torch.tensor(0.5))
        n = e["pcl"].shape[0]
        assert n > 1000
        assert np.array_equal(cloud["src_pix"][tot:tot + n].cpu().numpy(), e["src_pix"].numpy())
        np.testing.assert_allclose(cloud["xyz_world"][tot:tot + n].cpu().numpy(), e["pcl"].numpy(),
                                   rtol=PTS_RTOL, atol=PTS_ATOL)
        np.testing.assert_allclose(cloud["rgb"][tot:tot + n].cpu().numpy(), e["rgb"].numpy(), atol=RGB_ATOL, rtol=0)
        tot += n
    assert int(cloud["total"]) == tot


@pytest.mark.parametrize("name,views", [("tiny", 3), ("c1_nvidia_1view", 1)])
def test_fused_pipeline_equals_staged(name, views):
    """pgdvs_uwp_bin (uwp kernel files points under raster cells itself, packed rgbd frames)
    vs the stage-by-stage path (uwp -> [P,3] cloud -> pgdvs_bin_points): identical clouds and
    bit-identical fragments / images / masks; and the fragments match the oracle on that cloud."""
    import pgdvs_b200
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    dev = _dev()
    wl = synthetic.make_workload(name, dev, n_views=views, mask_mode="ellipse" if name != "tiny" else "full")
    pairs, cams = wl.jobs(range(views))
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    kw = dict(radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb,
              return_fragments=True)
    a = render_prepared(prep, fused=True, return_cloud=True, **kw)
    b = render_prepared(prep, fused=False, **kw)
    torch.cuda.synchronize()
    P = int(a["cloud"]["total"])
    assert P == int(b["cloud"]["total"]) and P > 0
    assert torch.equal(a["first_idx"], b["first_idx"]) and torch.equal(a["num_points"], b["num_points"])
    assert torch.equal(a["cloud"]["xyz_ndc"][:P], b["cloud"]["xyz_ndc"][:P])
    assert torch.equal(a["cloud"]["rgb"][:P], b["cloud"]["rgb"][:P])
    for k in ("idx", "zbuf", "dists", "image", "mask"):
        assert torch.equal(a[k], b[k]), k
    ndc = a["cloud"]["xyz_ndc"][:P].cpu().numpy()
    ref_frags = oracle.rasterize_points(ndc, a["first_idx"].cpu().numpy(), a["num_points"].cpu().numpy(),
                                        (wl.H, wl.W), wl.radius, wl.K, n_threads=8, banded=True)
    assert np.array_equal(a["idx"].cpu().numpy(), ref_frags[0])
    assert np.array_equal(a["zbuf"].cpu().numpy(), ref_frags[1])
    assert np.array_equal(a["dists"].cpu().numpy(), ref_frags[2])


def test_job_groups_do_not_change_results():
    """24 views = 2 time steps x 12 cameras: the jobs of a time step form a group (shared source
    pair, different target camera).  Grouped and ungrouped runs must be bit-identical."""
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    dev = _dev()
    wl = synthetic.make_workload("c2_nvidia_seq", dev, n_views=24, mask_mode="ellipse")
    pairs, cams = wl.jobs(range(24))
    a_prep = prepare_views(pairs, cams, wl.H, wl.W, dev, group_jobs=True)
    b_prep = prepare_views(pairs, cams, wl.H, wl.W, dev, group_jobs=False)
    assert a_prep.n_groups == 4 and b_prep.n_groups == 0  # 2 steps x (fwd, bwd) pairs
    kw = dict(radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb,
              return_fragments=True, return_cloud=True)
    a = render_prepared(a_prep, **kw)
    b = render_prepared(b_prep, **kw)
    c = render_prepared(a_prep, fused=False, **{k: v for k, v in kw.items() if k != "return_cloud"})
    torch.cuda.synchronize()
    P = int(a["cloud"]["total"])
    assert P == int(b["cloud"]["total"]) == int(c["cloud"]["total"]) and P > 0
    for x in (b, c):
        assert torch.equal(a["first_idx"], x["first_idx"]) and torch.equal(a["num_points"], x["num_points"])
        assert torch.equal(a["cloud"]["xyz_ndc"][:P], x["cloud"]["xyz_ndc"][:P])
        assert torch.equal(a["cloud"]["rgb"][:P], x["cloud"]["rgb"][:P])
        for k in ("idx", "zbuf", "dists", "image", "mask"):
            assert torch.equal(a[k], x[k]), k


def test_captured_render_replays_the_eager_step():
    """CapturedRender: the step as one CUDA graph; replays give the eager bits and follow in-place
    updates of the inputs."""
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import CapturedRender, prepare_views, render_prepared
    dev = _dev()
    wl = synthetic.make_workload("tiny", dev, n_views=2)
    pairs, cams = wl.jobs(range(2))
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    kw = dict(radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb,
              return_fragments=True)
    eager = {k: v.clone() for k, v in render_prepared(prep, **kw).items() if torch.is_tensor(v)}
    cap = CapturedRender(prep, **kw)
    for _ in range(2):
        out = cap.replay()
        torch.cuda.synchronize()
        for k in ("idx", "zbuf", "dists", "image", "mask"):
            assert torch.equal(out[k], eager[k]), k
    # new content behind the captured pointers: the static frame is blended in the epilogue
    wl.static_rgb.mul_(0.5)
    out = cap.replay()
    again = render_prepared(prep, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out["image"], again["image"]) and not torch.equal(out["image"], eager["image"])
    assert torch.equal(out["idx"], eager["idx"])


def test_job_groups_larger_than_one_member_chunk():
    """k_uwp stages the members of a group in shared memory 32 at a time: a group of 48 target views
    (the 12 cameras of a time step, four times over) takes two chunks and must still equal the
    ungrouped run bit for bit."""
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    dev = _dev()
    wl = synthetic.make_workload("c2_nvidia_seq", dev, n_views=12, mask_mode="ellipse")
    views = list(range(12)) * 4
    pairs, cams = wl.jobs(views)
    a_prep = prepare_views(pairs, cams, wl.H, wl.W, dev, group_jobs=True)
    b_prep = prepare_views(pairs, cams, wl.H, wl.W, dev, group_jobs=False)
    assert a_prep.n_groups == 2 and b_prep.n_groups == 0  # (fwd, bwd) pair of the one time step, 48 members each
    kw = dict(radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb[views],
              return_fragments=True, return_cloud=True)
    a = render_prepared(a_prep, **kw)
    b = render_prepared(b_prep, **kw)
    torch.cuda.synchronize()
    P = int(a["cloud"]["total"])
    assert P == int(b["cloud"]["total"]) and P > 0
    assert torch.equal(a["first_idx"], b["first_idx"]) and torch.equal(a["num_points"], b["num_points"])
    assert torch.equal(a["cloud"]["xyz_ndc"][:P], b["cloud"]["xyz_ndc"][:P])
    assert torch.equal(a["cloud"]["rgb"][:P], b["cloud"]["rgb"][:P])
    for k in ("idx", "zbuf", "dists", "image", "mask"):
        assert torch.equal(a[k], b[k]), k
    # the four replicas of a camera render the same frame
    assert torch.equal(a["image"][:12], a["image"][36:48])


@pytest.mark.parametrize("mode,fn", [("alpha", "alpha_composite"), ("norm", "norm_weighted_sum"), ("wsum", "weighted_sum")])
def test_standalone_compositors(mode, fn):
    import pgdvs_b200
    rng = np.random.default_rng(2)
    N, K, H, W, C, P = 2, 5, 9, 11, 6, 300
    idx = rng.integers(-1, P, (N, K, H, W)).astype(np.int64)
    al = rng.uniform(0, 1, (N, K, H, W)).astype(np.float32)
    ft = rng.uniform(-1, 1, (C, P)).astype(np.float32)
    d = _dev()
    out = getattr(pgdvs_b200, fn)(T(idx).to(d), T(al).to(d), T(ft).to(d)).cpu().numpy()
    exp = oracle.composite(idx, al, ft, mode)
    assert np.array_equal(out, exp)  # same op order, explicitly rounded -> bit-exact


def test_merge_blend():
    import pgdvs_b200
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 7, 9
    dr, tr, st = (torch.rand(B, 3, H, W, generator=g) for _ in range(3))
    dm = (torch.rand(B, 1, H, W, generator=g) < 0.5).float()
    tm = (torch.rand(B, 1, H, W, generator=g) < 0.5).float()
    d = _dev()
    rgb, m, comb = pgdvs_b200.ops.merge_blend(dr.to(d), dm.to(d), tr.to(d), tm.to(d), st.to(d))
    e_rgb, e_m = ref.merge_dyn_track(dr, dm, tr, tm)
    e_c = ref.blend_static_dynamic(st, e_rgb, e_m)
    assert torch.equal(rgb.cpu(), e_rgb) and torch.equal(m.cpu(), e_m) and torch.equal(comb.cpu(), e_c)


@pytest.mark.parametrize("shape", [(2, 5, 7, 3), (1, 288, 544, 3), (3,), (1, 4)])
def test_quantize_u8_matches_evaluator_rule(shape):
    """engines/evaluator_pgdvs.py:51-77: nan_to_num, clamp(0,1), (x*255).byte()."""
    import pgdvs_b200
    g = torch.Generator().manual_seed(3)
    x = torch.rand(shape, generator=g) * 1.4 - 0.2
    flat = x.view(-1)
    flat[0] = float("nan")
    flat[-1] = 1.0
    if flat.numel() > 4:
        flat[1], flat[2], flat[3] = 0.0, 254.999 / 255.0, 1.0 / 255.0
    exp = (torch.nan_to_num(x, nan=0.0).clamp(0.0, 1.0) * 255).to(torch.uint8)
    out = pgdvs_b200.ops.quantize_u8(x.to(_dev()))
    assert out.dtype == torch.uint8 and out.shape == x.shape
    assert torch.equal(out.cpu(), exp)


def test_knn_mean_dist():
    import pgdvs_b200
    g = torch.Generator().manual_seed(1)
    pcl = torch.randn(3000, 3, generator=g)
    flags, thres, avg = ref.knn_outlier_flags(pcl, knn=50, std_thres=0.1)
    out = pgdvs_b200.ops.knn_mean_dist(pcl.to(_dev()), pcl.to(_dev()), 51, skip_first=1).cpu()
    np.testing.assert_allclose(out.numpy(), avg.numpy(), rtol=2e-5, atol=1e-7)
    q = torch.randn(500, 3, generator=g)
    d2 = ((q[:, None] - pcl[None]) ** 2).sum(-1)
    exp = torch.topk(d2, 7, dim=1, largest=False).values.mean(1)
    out = pgdvs_b200.ops.knn_mean_dist(q.to(_dev()), pcl.to(_dev()), 7).cpu()
    np.testing.assert_allclose(out.numpy(), exp.numpy(), rtol=2e-5, atol=1e-7)


def test_render_dyn_pcl_l2_matches_oracle(golden_dir):
    """PGDVSDynamicRenderer.render_dyn_pcl on the cloud the REAL reference handed to pytorch3d."""
    import pgdvs_b200
    from types import SimpleNamespace
    g = np.load(golden_dir / "dyn_pcl_case0.npz")
    H, W = int(g["H"]), int(g["W"])
    d = _dev()
    cfg = SimpleNamespace(dyn_render_pcl_pt_radius=float(g["b_radius"]), dyn_render_pcl_pts_per_pixel=int(g["b_ppp"]))
    r = pgdvs_b200.PGDVSDynamicRenderer()
    img, mask, frags = r.render_dyn_pcl(dyn_mask=torch.zeros(H, W, 1, device=d), dyn_pcl=T(g["b_points"][0]).to(d),
                                        rgbs=T(g["b_features"][0]).to(d), flat_cam=T(g["flat_cam_tgt"]).to(d),
                                        render_cfg=cfg, return_fragments=True)
    e_img, e_mask, (ndc, e_idx, e_z, e_d) = ref.render_dyn_pcl(
        H=H, W=W, dyn_pcl=T(g["b_points"][0]), rgbs=T(g["b_features"][0]), flat_cam=T(g["flat_cam_tgt"]),
        radius=cfg.dyn_render_pcl_pt_radius, points_per_pixel=cfg.dyn_render_pcl_pts_per_pixel,
        return_fragments=True)
    # NDC inputs differ by fp32 rounding of the camera transform, so compare idx as a fraction
    agree = (frags["idx"].cpu().numpy() == e_idx).mean()
    assert agree > 0.995, agree
    same = np.all(frags["idx"].cpu().numpy() == e_idx, axis=-1)[0]
    np.testing.assert_allclose(img.cpu().numpy()[same], e_img.numpy()[same], atol=1e-4, rtol=0)
    assert (mask.cpu().numpy() == e_mask.numpy()).mean() > 0.995
    # empty cloud -> zeros (pgdvs_renderer_dyn.py:680-682)
    img0, mask0 = r.render_dyn_pcl(dyn_mask=torch.zeros(H, W, 1, device=d), dyn_pcl=torch.zeros(0, 3, device=d),
                                   rgbs=torch.zeros(0, 3, device=d), flat_cam=T(g["flat_cam_tgt"]).to(d), render_cfg=cfg)
    assert img0.shape == (H, W, 3) and float(img0.abs().sum()) == 0 and float(mask0.sum()) == 0


def test_forward_end_to_end_vs_oracle():
    """PGDVSDynamicRenderer.forward on a reference-shaped data dict: the fused GPU path vs the
    oracle pipeline run on the GPU's own NDC cloud (bit-exact raster) and vs the oracle's own
    cloud (tolerance + idx-agreement fraction)."""
    import pgdvs_b200
    from types import SimpleNamespace
    d = _dev()
    B, H, W = 2, 24, 40
    g = torch.Generator().manual_seed(3)
    data = {
        "rgb_src_temporal": torch.rand(B, 2, H, W, 3, generator=g),
        "depth_src_temporal": 2 + 3 * torch.rand(B, 2, H, W, 1, generator=g),
        "dyn_mask_src_temporal": (torch.rand(B, 2, H, W, 1, generator=g) < 0.8).float(),
        "flow_fwd": 1.5 * torch.randn(B, H, W, 2, generator=g),
        "flow_fwd_occ_mask": (torch.rand(B, H, W, 1, generator=g) < 0.1).float(),
        "time_src_temporal": torch.tensor([[0.0, 1.0], [4.0, 5.0]]),
        "time_tgt": torch.tensor([[0.25], [4.5]]),
    }
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2

    def flat(tx):
        c2w = torch.eye(4)
        c2w[:3, 3] = torch.tensor([tx, 0.01, 0.0])
        return torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), c2w.reshape(-1)])

    data["flat_cam_src_temporal"] = torch.stack([torch.stack([flat(0.0), flat(0.05)]), torch.stack([flat(0.1), flat(0.15)])])
    data["flat_cam_tgt"] = torch.stack([flat(0.02), flat(0.13)])
    data["dyn_mask_src_temporal"][1, 0] = 0  # second batch item: empty mask branch (:104, :133-152)
    cfg = SimpleNamespace(dyn_render_type="pcl", dyn_render_pcl_pt_radius=0.05, dyn_render_pcl_pts_per_pixel=4,
                          dyn_render_use_flow_consistency=True, dyn_pcl_remove_outlier=False)
    static = torch.rand(B, 3, H, W, generator=g)
    r = pgdvs_b200.PGDVSDynamicRenderer()
    rgb, mask, info = r(({k: v.to(d) for k, v in data.items()}), None, cfg, static_rgb=static.to(d))
    assert rgb.shape == (B, 3, H, W) and mask.shape == (B, 1, H, W)
    assert float(mask[1].sum()) == 0 and float(rgb[1].abs().sum()) == 0
    fs = data["flat_cam_src_temporal"]
    o = ref.compute_dyn_pcl(
        dyn_mask_1=data["dyn_mask_src_temporal"][0, 0], rgb_1=data["rgb_src_temporal"][0, 0],
        depth_1=data["depth_src_temporal"][0, 0], flow_12=data["flow_fwd"][0],
        flow_12_occ_mask=data["flow_fwd_occ_mask"][0], rgb_2=data["rgb_src_temporal"][0, 1],
        depth_2=data["depth_src_temporal"][0, 1], K_1=fs[0, 0, 2:18].reshape(4, 4), c2w_1=fs[0, 0, 18:34].reshape(4, 4),
        K_2=fs[0, 1, 2:18].reshape(4, 4), c2w_2=fs[0, 1, 18:34].reshape(4, 4), time_1=torch.tensor(0.0),
        time_2=torch.tensor(1.0), time_tgt=torch.tensor(0.25), use_flow_consistency=True)
    e_img, e_mask = ref.render_dyn_pcl(H=H, W=W, dyn_pcl=o["pcl"], rgbs=o["rgb"], flat_cam=data["flat_cam_tgt"][0],
                                       radius=0.05, points_per_pixel=4)
    # the two sides project their own clouds (fused kernel vs torch CPU ops): a differing pixel must
    # be a boundary flip or a z tie of the oracle's cloud, anything else fails
    diff = (rgb[0].cpu().permute(1, 2, 0) - e_img).abs().max(dim=-1).values
    bad = (diff > 1e-4) | (mask[0, 0].cpu() != e_mask[..., 0])
    ndc = ref.world_to_ndc(o["pcl"], ref.camera_from_flat_cam(data["flat_cam_tgt"][0])).numpy()
    attribute_mismatches(bad.numpy(), ndc, H, W, 0.05, 4)
    assert bad.float().mean() < 0.01
    comb = info["combined_rgb"].cpu()
    e_comb = ref.blend_static_dynamic(static, rgb.cpu(), mask.cpu())
    assert torch.equal(comb, e_comb)


def test_compat_namespace_runs_the_reference_call_sequence(golden_dir):
    """`import pgdvs_b200.compat as pytorch3d`: the statement sequence of
    pgdvs_renderer_dyn.py:676-722 (render_dyn_pcl) and :405-419 (knn statistics), written against
    the pytorch3d names, runs unchanged and matches the oracle."""
    import pgdvs_b200.compat as pytorch3d
    import pgdvs_b200.compat.ops as p3d_ops
    g = np.load(golden_dir / "dyn_pcl_case0.npz")
    d = _dev()
    h, w = int(g["H"]), int(g["W"])
    flat_cam = T(g["flat_cam_tgt"]).to(d)
    dyn_pcl, rgbs = T(g["b_points"][0]).to(d), T(g["b_features"][0]).to(d)
    radius, ppp = float(g["b_radius"]), int(g["b_ppp"])
    # --- reference statements (names as in the reference) ---
    K = flat_cam[2:18].reshape((4, 4))
    c2w = flat_cam[18:34].reshape((4, 4))
    w2c = torch.inverse(c2w)
    img_size = torch.LongTensor([h, w]).reshape((1, 2))
    cameras_pytorch3d = pytorch3d.utils.cameras_from_opencv_projection(
        w2c[None, :3, :3], w2c[None, :3, 3], K[None, :3, :3], img_size)
    raster_settings = pytorch3d.renderer.PointsRasterizationSettings(
        image_size=(h, w), radius=radius, points_per_pixel=ppp, bin_size=0)
    rasterizer = pytorch3d.renderer.PointsRasterizer(cameras=cameras_pytorch3d, raster_settings=raster_settings)
    point_renderer = pytorch3d.renderer.PointsRenderer(
        rasterizer=rasterizer, compositor=pytorch3d.renderer.NormWeightedCompositor(background_color=(0, 0, 0)))
    dy_mesh = pytorch3d.structures.Pointclouds(points=dyn_pcl[None, ...], features=rgbs[None, ...])
    mesh_img = point_renderer(dy_mesh)[0, :, :, :3]
    dy_mesh.features = torch.ones_like(rgbs)[None, ...]
    mesh_mask = (point_renderer(dy_mesh)[0, :, :, :1] > 0.0).float()
    nn_dists, nn_idxs, nn_pts = p3d_ops.knn_points(dyn_pcl[None, ...], dyn_pcl[None, ...], K=9, return_nn=True)
    avg_nn_dist = torch.mean(nn_dists[0, :, 1:], dim=1)
    # --- checks ---
    e_img, e_mask = ref.render_dyn_pcl(H=h, W=w, dyn_pcl=dyn_pcl.cpu(), rgbs=rgbs.cpu(), flat_cam=flat_cam.cpu(),
                                       radius=radius, points_per_pixel=ppp)
    diff = (mesh_img.cpu() - e_img).abs().max(dim=-1).values
    bad = (diff > 1e-4) | (mesh_mask.cpu() != e_mask)[..., 0]
    ndc = ref.world_to_ndc(dyn_pcl.cpu(), ref.camera_from_flat_cam(flat_cam.cpu())).numpy()
    attribute_mismatches(bad.numpy(), ndc, h, w, radius, ppp)  # boundary flip or z tie, else fail
    assert bad.float().mean() < 0.01
    _, _, e_avg = ref.knn_outlier_flags(dyn_pcl.cpu(), knn=8)
    np.testing.assert_allclose(avg_nn_dist.cpu().numpy(), e_avg.numpy(), rtol=2e-5, atol=1e-7)
    assert nn_idxs.shape == (1, dyn_pcl.shape[0], 9) and nn_pts.shape == (1, dyn_pcl.shape[0], 9, 3)
    assert torch.equal(nn_idxs[0, :, 0].cpu(), torch.arange(dyn_pcl.shape[0]))  # nearest neighbour is the point itself
    d2 = ((dyn_pcl[:, None] - nn_pts[0]) ** 2).sum(-1)
    np.testing.assert_allclose(d2.cpu().numpy(), nn_dists[0].cpu().numpy(), rtol=1e-4, atol=1e-6)


def test_track_points_match_reference_golden(golden_dir):
    """Track branch kernel vs the cloud the REAL reference computed (fixture) and vs the oracle on
    a larger random case (order of survivors exact, positions / colours within tolerance)."""
    from pgdvs_b200 import track
    d = _dev()
    g = np.load(golden_dir / "track_pcl.npz")
    kw = dict(tracks=T(g["tracks"]), visibles=T(g["visibles"]), rgbs=T(g["rgbs"]), depths=T(g["depths"]),
              flat_cams=T(g["flat_cams"]), times=T(g["times"]), time_tgt=T(g["time_tgt"]),
              idx_temporal_closest=g["idx_temporal_closest"].tolist(), idx_real_track=g["idx_real_track"].tolist())
    pcl, rgb, tid = track.track_points(**{k: (v.to(d) if torch.is_tensor(v) and k not in ("time_tgt",) else v)
                                          for k, v in kw.items()}, return_track_id=True)
    assert pcl.shape[0] == g["out_pcl"].shape[0] > 0
    np.testing.assert_allclose(pcl.cpu().numpy(), g["out_pcl"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(rgb.cpu().numpy(), g["out_rgb"], atol=5e-6, rtol=0)
    _, _, e_tid = ref.compute_pcl_for_tgt(**kw)
    assert torch.equal(tid.cpu(), e_tid)
    # larger random case against the oracle
    gen = torch.Generator().manual_seed(4)
    H, W, F, Q = 36, 52, 8, 5000
    rgbs = torch.rand(F, H, W, 3, generator=gen)
    depths = 1 + 4 * torch.rand(F, H, W, 1, generator=gen)
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2
    flat = []
    for f in range(F):
        c2w = torch.eye(4)
        c2w[:3, 3] = torch.tensor([0.03 * f, -0.01 * f, 0.0])
        flat.append(torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), c2w.reshape(-1)]))
    flat = torch.stack(flat)
    uv0 = torch.stack([torch.rand(Q, generator=gen) * (W - 1), torch.rand(Q, generator=gen) * (H - 1)], 1)
    tracks = uv0[:, None, :] + torch.cumsum(0.7 * torch.randn(Q, F, 2, generator=gen), dim=1)
    visibles = torch.rand(Q, F, generator=gen) < 0.6
    times = torch.arange(F, dtype=torch.float32)
    kw = dict(tracks=tracks, visibles=visibles, rgbs=rgbs, depths=depths, flat_cams=flat, times=times,
              time_tgt=torch.tensor(3.3), idx_temporal_closest=[3, 4], idx_real_track=[0, 1, 2, 5, 6, 7])
    e_pcl, e_rgb, e_tid = ref.compute_pcl_for_tgt(**kw)
    pcl, rgb, tid = track.track_points(**{k: (v.to(d) if torch.is_tensor(v) and k != "time_tgt" else v)
                                          for k, v in kw.items()}, return_track_id=True)
    assert torch.equal(tid.cpu(), e_tid) and e_tid.numel() > 100
    # |dt| ties between two frames may be ordered differently by torch.argsort: the interpolated
    # point is the same up to rounding, so compare with a tolerance
    np.testing.assert_allclose(pcl.cpu().numpy(), e_pcl.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(rgb.cpu().numpy(), e_rgb.numpy(), atol=5e-6, rtol=0)


def test_static_geo_point_renderer():
    """StaticGeoPointRenderer.forward (st_geo_renderer.py:26-122) vs the oracle render."""
    import pgdvs_b200
    from types import SimpleNamespace
    d = _dev()
    gen = torch.Generator().manual_seed(8)
    H, W, P = 30, 44, 6000
    pts = torch.randn(P, 3, generator=gen) * torch.tensor([1.2, 0.8, 0.6]) + torch.tensor([0.0, 0.0, 4.0])
    cols = torch.rand(P, 3, generator=gen)
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2
    flat = torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), torch.eye(4).reshape(-1)])
    cfg = SimpleNamespace(st_pcl_remove_outlier=False, st_render_pcl_pt_radius=0.03, st_render_pcl_pts_per_pixel=3)
    r = pgdvs_b200.StaticGeoPointRenderer()
    img, mask = r(tgt_h=H, tgt_w=W, flat_tgt_cam=flat.to(d), st_pcl_rgb=torch.cat([pts, cols], 1).to(d), render_cfg=cfg)
    e_img, e_mask = ref.render_dyn_pcl(H=H, W=W, dyn_pcl=pts, rgbs=cols, flat_cam=flat, radius=0.03, points_per_pixel=3)
    assert (mask.cpu() == e_mask).float().mean() > 0.995
    diff = (img.cpu() - e_img).abs().max(dim=-1).values
    assert (diff < 1e-4).float().mean() > 0.99
    cfg.st_pcl_remove_outlier, cfg.st_pcl_outlier_knn, cfg.st_pcl_outlier_std_thres = True, 20, 0.1
    img2, mask2 = r(tgt_h=H, tgt_w=W, flat_tgt_cam=flat.to(d), st_pcl_rgb=torch.cat([pts, cols], 1).to(d), render_cfg=cfg)
    flags, _, _ = ref.knn_outlier_flags(pts, knn=20, std_thres=0.1)
    e_img2, e_mask2 = ref.render_dyn_pcl(H=H, W=W, dyn_pcl=pts[flags], rgbs=cols[flags], flat_cam=flat, radius=0.03,
                                         points_per_pixel=3)
    assert (mask2.cpu() == e_mask2).float().mean() > 0.99
    assert float(mask2.sum()) < float(mask.sum())  # outliers removed -> fewer covered pixels


def test_pytorch3d_facade_generic_equals_fused():
    """The pytorch3d-shaped classes: the generic two-step path (rasterize -> torch weights ->
    stand-alone compositor, any C) and the fused path give the same image."""
    import pgdvs_b200 as p3
    d = _dev()
    g = torch.Generator().manual_seed(9)
    H, W, P = 20, 30, 1500
    pts = torch.randn(1, P, 3, generator=g) * torch.tensor([1.0, 0.6, 0.5]) + torch.tensor([0.0, 0.0, 4.0])
    feats = torch.rand(1, P, 3, generator=g)
    Kc = torch.eye(3)[None].clone()
    Kc[0, 0, 0] = Kc[0, 1, 1] = 0.9 * W
    Kc[0, 0, 2], Kc[0, 1, 2] = W / 2, H / 2
    cams = p3.cameras_from_opencv_projection(torch.eye(3)[None].to(d), torch.zeros(1, 3).to(d), Kc.to(d),
                                             torch.LongTensor([[H, W]]))
    s = p3.PointsRasterizationSettings(image_size=(H, W), radius=0.08, points_per_pixel=5, bin_size=0)
    rast = p3.PointsRasterizer(cameras=cams, raster_settings=s)
    for comp in (p3.NormWeightedCompositor(background_color=(0, 0, 0)), p3.AlphaCompositor(background_color=(0.1, 0.2, 0.3))):
        rend = p3.PointsRenderer(rasterizer=rast, compositor=comp)
        cloud = p3.Pointclouds(points=pts.to(d), features=feats.to(d))
        fused = rend(cloud)
        generic = rend(cloud, dummy_kwarg=True) if False else None
        frags = rast(cloud)
        w = 1 - frags.dists.permute(0, 3, 1, 2) / (0.08 * 0.08)
        two_step = comp(frags.idx.long().permute(0, 3, 1, 2), w, feats[0].to(d).permute(1, 0)).permute(0, 2, 3, 1)
        assert fused.shape == (1, H, W, 3)
        np.testing.assert_allclose(fused.cpu().numpy(), two_step.cpu().numpy(), atol=IMG_ATOL, rtol=0)
        assert float((frags.idx >= 0).float().mean()) > 0.05


def _knn_brute(q, r, K, skip):
    """the brute-force kernel (no workspace -> never the grid path)"""
    import pgdvs_b200
    from pgdvs_b200 import _cabi, ops
    out = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
    _cabi.check(_cabi.lib().pgdvs_knn_mean_dist(q.data_ptr(), q.shape[0], r.data_ptr(), r.shape[0], K, skip,
                                                out.data_ptr(), None, 0, ops._stream_ptr(q.device)), "knn")
    return out


@pytest.mark.parametrize("kind", ["surface", "volume", "clustered", "planar_dups"])
def test_knn_grid_equals_brute_force(kind):
    """The uniform-grid search (large clouds) is exact: same K-nearest statistics as brute force."""
    import pgdvs_b200
    g = torch.Generator().manual_seed(17)
    P = 30000
    if kind == "surface":  # a depth-map-like height field
        uv = torch.rand(P, 2, generator=g) * torch.tensor([8.0, 5.0])
        z = 4 + torch.sin(uv[:, :1] * 1.3) + 0.02 * torch.randn(P, 1, generator=g)
        pts = torch.cat([uv, z], 1)
    elif kind == "volume":
        pts = torch.rand(P, 3, generator=g) * torch.tensor([3.0, 2.0, 4.0])
    elif kind == "clustered":
        centres = torch.randn(12, 3, generator=g) * 3
        pts = centres[torch.randint(0, 12, (P,), generator=g)] + 0.05 * torch.randn(P, 3, generator=g)
        pts[:200] = torch.randn(200, 3, generator=g) * 30  # far outliers stretch the box
    else:  # a plane with many exact duplicates
        pts = torch.cat([torch.rand(P, 2, generator=g), torch.zeros(P, 1)], 1)
        pts[1000:2000] = pts[:1000]
    d = _dev()
    pts = pts.float().contiguous().to(d)
    for K, skip in ((51, 1), (7, 0)):
        fast = pgdvs_b200.ops.knn_mean_dist(pts, pts, K, skip_first=skip)
        slow = _knn_brute(pts, pts, K, skip)
        np.testing.assert_allclose(fast.cpu().numpy(), slow.cpu().numpy(), rtol=1e-5, atol=1e-9)
    # cross search: queries that lie (far) outside the reference cloud's box
    q = (pts[:5000] * 1.7 + torch.tensor([0.5, -2.0, 9.0], device=d)).contiguous()
    fast = pgdvs_b200.ops.knn_mean_dist(q, pts, 51, skip_first=0)
    slow = _knn_brute(q, pts, 51, 0)
    np.testing.assert_allclose(fast.cpu().numpy(), slow.cpu().numpy(), rtol=1e-5, atol=1e-9)


def test_forward_outlier_filter_is_applied():
    """dyn_pcl_remove_outlier=True in the batched forward: the keep mask must reach the kernels
    (descriptor caches are invalidated when it is attached)."""
    import pgdvs_b200
    from types import SimpleNamespace
    d = _dev()
    B, H, W = 3, 24, 40
    g = torch.Generator().manual_seed(31)
    one = {
        "rgb_src_temporal": torch.rand(1, 2, H, W, 3, generator=g),
        "depth_src_temporal": 2 + 0.2 * torch.rand(1, 2, H, W, 1, generator=g),
        "dyn_mask_src_temporal": (torch.rand(1, 2, H, W, 1, generator=g) < 0.9).float(),
        "flow_fwd": 0.5 * torch.randn(1, H, W, 2, generator=g),
        "flow_fwd_occ_mask": torch.zeros(1, H, W, 1),
    }
    one["depth_src_temporal"][0, :, 4:10, 4:10] = 9.0  # a floating blob in both frames: statistical outliers
    data = {k: v.expand(B, *v.shape[1:]).contiguous() for k, v in one.items()}
    data["time_src_temporal"] = torch.tensor([[0.0, 1.0]] * B)
    data["time_tgt"] = torch.tensor([[0.5]] * B)
    Kc = torch.eye(4)
    Kc[0, 0] = Kc[1, 1] = 0.9 * W
    Kc[0, 2], Kc[1, 2] = W / 2, H / 2

    def flat(tx):
        c2w = torch.eye(4)
        c2w[:3, 3] = torch.tensor([tx, 0.0, 0.0])
        return torch.cat([torch.tensor([float(H), float(W)]), Kc.reshape(-1), c2w.reshape(-1)])

    data["flat_cam_src_temporal"] = torch.stack([torch.stack([flat(0.0), flat(0.05)])] * B)
    data["flat_cam_tgt"] = torch.stack([flat(0.01), flat(0.02), flat(0.03)])
    dd = {k: v.to(d) for k, v in data.items()}
    r = pgdvs_b200.PGDVSDynamicRenderer()
    base = dict(dyn_render_type="pcl", dyn_render_pcl_pt_radius=0.1, dyn_render_pcl_pts_per_pixel=4,
                dyn_render_use_flow_consistency=False, dyn_pcl_outlier_knn=20, dyn_pcl_outlier_std_thres=0.1)
    calls = {"n": 0}
    orig = pgdvs_b200.ops.knn_mean_dist

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    pgdvs_b200.ops.knn_mean_dist = counting
    try:
        rgb_on, m_on, _ = r(dd, None, SimpleNamespace(dyn_pcl_remove_outlier=True, **base))
    finally:
        pgdvs_b200.ops.knn_mean_dist = orig
    rgb_off, m_off, _ = r(dd, None, SimpleNamespace(dyn_pcl_remove_outlier=False, **base))
    # (views only share a KNN run when their pairs reference the SAME tensors, as render_views
    #  callers can arrange; a reference-style data dict carries one copy per batch item)
    assert 1 <= calls["n"] <= B
    # the filter removed the blob's points: the renders differ where they used to splat
    assert not torch.equal(rgb_on, rgb_off) and float(m_on.sum()) <= float(m_off.sum())


def test_fused_outlier_filter_matches_oracle():
    """render_views_filtered (statistical outlier filter fused into the batched path, no host sync:
    world points by pixel slot -> grid KNN -> device-side median/std -> keep mask) against the
    oracle's per-pair knn_outlier_flags (pgdvs_renderer_dyn.py:401-457): same thresholds (rtol 1e-4),
    same survivors per view, and the render of the surviving cloud bit-exact in idx."""
    from types import SimpleNamespace
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import render_views_filtered
    d = _dev()
    knn = 8
    wl = synthetic.make_workload("tiny", d)
    sc = wl.scene  # (CPU and CUDA generators differ: the oracle gets the GPU scene's tensors)
    H, W, K, r = wl.H, wl.W, wl.K, wl.radius
    pairs, cams = wl.jobs(range(wl.n_views))
    cfg = SimpleNamespace(dyn_pcl_outlier_knn=knn, dyn_pcl_outlier_std_thres=0.1)
    out = render_views_filtered(pairs, cams, H, W, radius=r, points_per_pixel=K, render_cfg=cfg, compositor="norm",
                                return_fragments=True, return_cloud=True)
    num = out["num_points"].cpu().numpy()
    first = out["first_idx"].cpu().numpy()
    thres_gpu = sorted(out["outlier_thres"].cpu().tolist())
    thres_ref = {}
    for v in range(wl.n_views):
        pcl, rgb = [], []
        for p in wl.view_pairs[v]:
            a = p._src_frames
            o = ref.compute_dyn_pcl(
                dyn_mask_1=sc.mask[a[0]].cpu(), rgb_1=sc.rgb[a[0]].cpu(), depth_1=sc.depth[a[0]].cpu(),
                flow_12=p.flow_12.reshape(H, W, 2).cpu(), flow_12_occ_mask=torch.zeros(H, W, 1), rgb_2=sc.rgb[a[1]].cpu(),
                depth_2=sc.depth[a[1]].cpu(), K_1=T(sc.K), c2w_1=T(sc.c2w[a[0]]), K_2=T(sc.K), c2w_2=T(sc.c2w[a[1]]),
                time_1=torch.tensor(sc.times[a[0]]), time_2=torch.tensor(sc.times[a[1]]), time_tgt=torch.tensor(p._t_tgt))
            flags, th, _ = ref.knn_outlier_flags(o["pcl"], knn=knn)
            thres_ref[(a, p._t_tgt)] = float(th)
            pcl.append(o["pcl"][flags])
            rgb.append(o["rgb"][flags])
        pcl, rgb = torch.cat(pcl), torch.cat(rgb)
        assert 0 < pcl.shape[0] == int(num[v]), (v, pcl.shape[0], int(num[v]))
        Kc, c2w = wl.view_cams[v]
        flat = torch.cat([torch.tensor([float(H), float(W)]), T(Kc).reshape(-1), T(c2w).reshape(-1)])
        ndc = ref.world_to_ndc(pcl, ref.camera_from_flat_cam(flat)).numpy()
        g_ndc = out["cloud"]["xyz_ndc"][first[v]:first[v] + num[v]].cpu().numpy()
        np.testing.assert_allclose(g_ndc, ndc, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(thres_gpu, sorted(thres_ref.values()), rtol=1e-4)
    assert out["outlier_thres"].numel() == len(thres_ref) < len(pairs)  # views sharing a source pair share the statistics
    # the splat of the filtered cloud equals the oracle's on the GPU's own NDC points
    P = int(out["cloud"]["total"].item())
    img, (idx, zbuf, dists) = oracle.render_points(out["cloud"]["xyz_ndc"][:P].cpu().numpy(), first, num,
                                                   out["cloud"]["rgb"][:P].cpu().numpy(), (H, W), r, K, "norm",
                                                   background=(0, 0, 0))
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    np.testing.assert_allclose(out["image"].cpu().numpy(), img, atol=1e-5, rtol=0)
