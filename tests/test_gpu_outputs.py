"""GPU parity for the extra products of the rasterize-and-composite pass and for the `proj_func`
kernel:

* composited depth (named in the north star): the compositor applied to the
  hits' view-space z as one more feature channel — oracle = the same compositor restatement at
  C = 4 (features = rgb + z);
* 8-bit frames / masks written by the epilogue (engines/evaluator_pgdvs.py:51-77) — must equal
  the stand-alone quantiser bit for bit;
* Projector.compute_projections (models/gnt/projector.py:41-73) against the fixture recorded from
  the real reference (tests/golden/projections.npz) and the oracle restatement.

Tolerances: depth |delta| <= 1e-5 * max(1, z) (fp32 sum order, like the images); u8 exact;
projections rtol 1e-5 / atol 1e-4 pixels (fp32 4x4 product association)."""
import numpy as np
import pytest
import torch

from oracle import pgdvs_ref as ref
from oracle import raster as oracle

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _dev():
    return torch.device("cuda:0")


def _cloud(rng, H, W, P, zq=None):
    s = min(H, W) / 2
    pts = np.stack([rng.uniform(-W / 2 / s - 0.2, W / 2 / s + 0.2, P), rng.uniform(-H / 2 / s - 0.2, H / 2 / s + 0.2, P),
                    rng.uniform(0.5, 6.0, P)], 1).astype(np.float32)
    if zq:
        pts[:, 2] = np.round(pts[:, 2] * zq) / zq
    return pts


@pytest.mark.parametrize("compositor", ["norm", "alpha"])
@pytest.mark.parametrize("H,W,K,r,P", [(26, 38, 8, 0.12, 4000),     # generic kernel (halo 2)
                                       (96, 128, 8, 0.029, 30000),  # pair kernel (halo 1, K = 8)
                                       (96, 128, 5, 0.029, 30000),  # tile kernel, K not a template size
                                       (40, 56, 16, 0.1, 9000)])
def test_composited_depth_matches_oracle(compositor, H, W, K, r, P):
    from pgdvs_b200 import _cabi
    _cabi.debug_switch("no_pair", 0)  # small launch: take the pair kernel wherever it applies
    try:
        _depth_case(compositor, H, W, K, r, P)
    finally:
        _cabi.debug_switch("no_pair", -1)


def _depth_case(compositor, H, W, K, r, P):
    import pgdvs_b200
    rng = np.random.default_rng(K * 100 + H)
    d = _dev()
    pts = _cloud(rng, H, W, P, zq=32)
    rgb = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    fi = np.array([0, P // 3], np.int64)
    npc = np.array([P // 3, P - P // 3], np.int64)
    out = pgdvs_b200.render_packed(T(pts).to(d), T(rgb).to(d), T(fi).to(d), T(npc).to(d), (H, W), r, K,
                                   compositor=compositor, background=(0, 0, 0), return_depth=True)
    # oracle: the compositor restatement with view z appended as a 4th feature channel
    feats4 = np.concatenate([rgb, pts[:, 2:3]], 1)
    img4, (idx, zbuf, dists) = oracle.render_points(pts, fi, npc, feats4, (H, W), r, K, compositor,
                                                    background=(0, 0, 0, 0), n_threads=4)
    depth = out["depth"].cpu().numpy()
    assert depth.shape == (2, H, W, 1)
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    np.testing.assert_allclose(depth, img4[..., 3:4], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out["image"].cpu().numpy(), img4[..., :3], rtol=0, atol=1e-5)
    # empty pixels: depth 0 (zero background); nearest-hit depth is zbuf[..., 0]
    empty = idx[..., 0] < 0
    assert np.all(depth[empty] == 0)
    assert np.array_equal(out["zbuf"].cpu().numpy()[..., 0], zbuf[..., 0])
    # ... and the same bits as running the fused compositor over C = 4 features
    out4 = pgdvs_b200.render_packed(T(pts).to(d), T(feats4).to(d), T(fi).to(d), T(npc).to(d), (H, W), r, K,
                                    compositor=compositor, background=(0, 0, 0, 0))
    assert torch.equal(out4["image"][..., 3:4], out["depth"])
    if compositor == "norm":
        # a convex combination of the kept depths lies between the nearest and the farthest of them
        zb = out["zbuf"].cpu().numpy()
        hit = ~empty
        zmax = np.where(zb >= 0, zb, -np.inf).max(-1)
        assert np.all(depth[..., 0][hit] >= zb[..., 0][hit] * (1 - 1e-5))
        assert np.all(depth[..., 0][hit] <= zmax[hit] * (1 + 1e-5))


@pytest.mark.parametrize("name,views", [("tiny", 3), ("c1_nvidia_1view", 1)])
def test_fused_path_depth_and_u8_outputs(name, views):
    """render_views: depth from the fused (uwp -> bin -> raster) path equals the staged path's, and
    the 8-bit frame / mask written by the epilogue equal pgdvs_quantize_u8 of the fp32 ones."""
    import pgdvs_b200
    from pgdvs_b200 import ops, synthetic
    d = _dev()
    wl = synthetic.make_workload(name, d, n_views=views)
    pairs, cams = wl.jobs(range(wl.n_views))
    kw = dict(radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=wl.static_rgb)
    a = pgdvs_b200.render_views(pairs, cams, wl.H, wl.W, return_depth=True, return_u8=True, return_fragments=True, **kw)
    b = pgdvs_b200.render_views(pairs, cams, wl.H, wl.W, return_depth=True, fused=False, return_fragments=True, **kw)
    assert torch.equal(a["depth"], b["depth"]) and torch.equal(a["image"], b["image"])
    assert torch.equal(a["image_u8"], ops.quantize_u8(a["image"]))
    assert torch.equal(a["mask_u8"], ops.quantize_u8(a["mask"]))
    assert a["image_u8"].dtype == torch.uint8 and tuple(a["image_u8"].shape) == (views, wl.H, wl.W, 3)
    # depth is bounded by the fragments it was composited from
    zb = a["zbuf"]
    hit = a["idx"][..., 0] >= 0
    assert bool(hit.any())
    zmax = torch.where(zb >= 0, zb, torch.full_like(zb, -1e30)).amax(-1)
    dd = a["depth"][..., 0]
    assert bool((dd[hit] >= zb[..., 0][hit] * (1 - 1e-5)).all()) and bool((dd[hit] <= zmax[hit] * (1 + 1e-5)).all())
    assert bool((dd[~hit] == 0).all())
    # 8-bit only (no fp32 frame written at all): same bytes
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    prep = prepare_views(pairs, cams, wl.H, wl.W, d)
    c = render_prepared(prep, return_u8=True, return_f32=False, **kw)
    assert c["image"] is None and torch.equal(c["image_u8"], a["image_u8"]) and torch.equal(c["mask_u8"], a["mask_u8"])
    # caller-owned 8-bit buffers (what a per-step frame gather double-buffers)
    img8 = torch.full(tuple(a["image_u8"].shape), 7, dtype=torch.uint8, device=d)
    msk8 = torch.full(tuple(a["mask_u8"].shape), 7, dtype=torch.uint8, device=d)
    e = render_prepared(prep, return_u8=True, u8_out=(img8, msk8), **kw)
    assert e["image_u8"] is img8 and torch.equal(img8, a["image_u8"]) and torch.equal(msk8, a["mask_u8"])
    with pytest.raises(ValueError):
        render_prepared(prep, return_u8=True, u8_out=(img8.float(), msk8), **kw)


def test_compute_projections_matches_reference_golden(golden_dir):
    """`proj_func` (Projector.compute_projections): fixture recorded from the real reference."""
    from pgdvs_b200 import ops
    g = np.load(golden_dir / "projections.npz")
    d = _dev()
    xyz = T(g["xyz"]).to(d)[:, None, :]          # [P, 1, 3] like the call at pgdvs_renderer_dyn.py:470-473
    cams = T(g["flat_cam"])[None, :]
    uv, mask = ops.compute_projections(xyz, cams)
    assert tuple(uv.shape) == (1, xyz.shape[0], 1, 2) and tuple(mask.shape) == (1, xyz.shape[0], 1)
    assert mask.dtype == torch.bool
    np.testing.assert_allclose(uv[0, :, 0].cpu().numpy(), g["uv"], rtol=1e-5, atol=1e-4)
    assert np.array_equal(mask[0, :, 0].cpu().numpy(), g["mask"])


def test_compute_projections_clamp_semantics_and_many_cameras():
    """Points behind / on the camera plane: clamp(z, min=1e-8), clamp(uv, +-1e6), mask = z > 0 —
    against the oracle restatement, for several cameras at once."""
    from pgdvs_b200 import ops
    rng = np.random.default_rng(7)
    d = _dev()
    n_cam, R, S = 3, 50, 4
    xyz = rng.uniform(-2, 2, (R, S, 3)).astype(np.float32)
    xyz[0, 0] = (0.3, -0.2, 0.0)        # on the camera plane of the identity camera
    xyz[1, 0] = (1.0, 1.0, -1.5)        # behind it
    xyz[2, 0] = (5.0, 0.0, 1e-9)        # huge u: clamped to 1e6
    cams = np.zeros((n_cam, 34), np.float32)
    for c in range(n_cam):
        Kc = np.eye(4, dtype=np.float32)
        Kc[0, 0] = Kc[1, 1] = 100 + 30 * c
        Kc[0, 2], Kc[1, 2] = 64, 48
        c2w = np.eye(4, dtype=np.float32)
        if c:
            a = 0.2 * c
            c2w[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
            c2w[:3, 3] = (0.1 * c, -0.05, 0.2)
        cams[c] = np.concatenate([[96, 128], Kc.reshape(-1), c2w.reshape(-1)])
    uv, mask = ops.compute_projections(T(xyz).to(d), T(cams))
    assert tuple(uv.shape) == (n_cam, R, S, 2)
    for c in range(n_cam):
        uv_o, m_o = ref.compute_projections(T(xyz).reshape(-1, 3), T(cams[c]))
        got = uv[c].reshape(-1, 2).cpu().numpy()
        want = uv_o.reshape(-1, 2).numpy()
        big = np.abs(want) >= 1e5  # near the plane the quotient is ill-conditioned: compare the clamp only
        np.testing.assert_allclose(got[~big], want[~big], rtol=2e-5, atol=2e-4)
        assert np.all(np.abs(got) <= 1e6)
        assert np.array_equal(mask[c].reshape(-1).cpu().numpy(), m_o.reshape(-1).numpy())
    with pytest.raises(AttributeError):
        ops.compute_projections(T(xyz).reshape(-1, 3).to(d), T(cams))
