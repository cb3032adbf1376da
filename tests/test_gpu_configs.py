"""Parity on all five BASELINE.json configs (one target view of each at full image size):
the whole fused CUDA path (unproject -> warp -> project -> bin -> rasterize -> composite ->
blend) against the CPU oracle run on the GPU path's own NDC cloud.

Bar: idx / zbuf / dists bit-exact, mask exact, blended image |delta| <= 1e-5.  For the 1080p
stress config the oracle is evaluated on a band of rows (its cost per row is what makes the full
image impractical on the CPU), the GPU result is checked there bit for bit and by
size-independent properties everywhere else."""
import numpy as np
import pytest
import torch

from oracle import raster as oracle

pytestmark = pytest.mark.gpu
IMG_ATOL = 1e-5


def _render(name, K=None, radius=None, views=1, **kw):
    import pgdvs_b200
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    dev = torch.device("cuda:0")
    wl = synthetic.make_workload(name, dev, n_views=views, K=K, radius=radius, **kw)
    pairs, cams = wl.jobs(range(views))
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    out = render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                          static_rgb=wl.static_rgb, return_fragments=True, return_cloud=True)
    torch.cuda.synchronize()
    return wl, out


def _check_full(wl, out, threads=16):
    P = int(out["cloud"]["total"])
    assert P > 0.5 * wl.points_per_view() * wl.n_views  # most source pixels survive
    ndc = out["cloud"]["xyz_ndc"][:P].cpu().numpy()
    rgb = out["cloud"]["rgb"][:P].cpu().numpy()
    fi, npc = out["first_idx"].cpu().numpy(), out["num_points"].cpu().numpy()
    img, (idx, zbuf, dists) = oracle.render_points(ndc, fi, npc, rgb, (wl.H, wl.W), wl.radius, wl.K, "norm",
                                                   background=(0, 0, 0), n_threads=threads, banded=True)
    ones, _ = oracle.render_points(ndc, fi, npc, np.ones_like(rgb), (wl.H, wl.W), wl.radius, wl.K, "norm",
                                   background=(0, 0, 0), n_threads=threads, banded=True)
    mask = (ones[..., :1] > 0).astype(np.float32)
    g_idx = out["idx"].cpu().numpy()
    bad = np.argwhere(g_idx != idx)
    assert bad.shape[0] == 0, f"{bad.shape[0]} idx mismatches, first {bad[:3].tolist()}"
    assert np.array_equal(out["zbuf"].cpu().numpy(), zbuf)
    assert np.array_equal(out["dists"].cpu().numpy(), dists)
    assert np.array_equal(out["mask"].cpu().numpy(), mask)
    blend = (1 - mask) * wl.static_rgb.cpu().numpy() + mask * img
    np.testing.assert_allclose(out["image"].cpu().numpy(), blend, atol=IMG_ATOL, rtol=0)
    return (g_idx[..., wl.K - 1] >= 0).mean()


def test_config1_nvidia_single_view():
    wl, out = _render("c1_nvidia_1view")
    assert (wl.H, wl.W, wl.K) == (288, 544, 8) and wl.points_per_view() == 313344
    assert _check_full(wl, out) > 0.5


def test_config2_nvidia_sequence_batch():
    """several views of the 144-view sequence in ONE launch (pytorch3d's N dimension)"""
    wl, out = _render("c2_nvidia_seq", views=5)
    assert out["idx"].shape == (5, 288, 544, 8)
    _check_full(wl, out)


def test_config3_iphone_k16_six_sources():
    wl, out = _render("c3_iphone")
    assert (wl.H, wl.W, wl.K) == (360, 480, 16) and wl.points_per_view() == 1036800
    _check_full(wl, out)


def test_config4_davis():
    wl, out = _render("c4_davis", views=2)
    assert (wl.H, wl.W, wl.K) == (480, 854, 8) and wl.points_per_view() == 819840
    _check_full(wl, out)


@pytest.mark.parametrize("K,radius", [(8, 0.01), (16, 0.005), (32, 0.02)])
def test_config5_stress_1080p(K, radius):
    wl, out = _render("c5_stress", K=K, radius=radius)
    H, W = wl.H, wl.W
    assert (H, W) == (1080, 1920) and wl.points_per_view() == 16588800
    P = int(out["cloud"]["total"])
    ndc = out["cloud"]["xyz_ndc"][:P].cpu().numpy()
    idx, zbuf, dists = (out[k].cpu().numpy() for k in ("idx", "zbuf", "dists"))
    filled = idx >= 0
    # properties that hold at any size
    assert np.all(filled[..., :-1] >= filled[..., 1:])                       # filled slots form a prefix
    zb = np.where(filled, zbuf, np.float32(3e38))
    assert np.all(zb[..., 1:] >= zb[..., :-1])                               # ascending depth
    assert np.all(dists[filled] < np.float32(radius) * np.float32(radius)) and np.all(dists[filled] >= 0)
    assert np.all(zbuf[filled] == ndc[idx[filled], 2])                       # zbuf is the point's own z
    assert np.all((zbuf == -1) == ~filled) and np.all((dists == -1) == ~filled)
    # bit-exact against the oracle on a band of rows
    rows = np.arange(500, 500 + (12 if radius <= 0.01 else 6))
    _, yf = oracle.pixel_center_ndc(H, W)
    near = (ndc[:, 1] < yf[rows[0]] + 1.5 * radius) & (ndc[:, 1] > yf[rows[-1]] - 1.5 * radius)
    sub = np.nonzero(near)[0]
    one = np.zeros(1, np.int64)
    ri, rz, rd = oracle.rasterize_points(ndc[sub], one, np.array([sub.size], np.int64), (H, W), radius, K,
                                         n_threads=16, banded=True)
    ref_idx = np.where(ri[0, rows] >= 0, sub[np.clip(ri[0, rows], 0, None)], -1)
    assert np.array_equal(idx[0, rows], ref_idx)
    assert np.array_equal(zbuf[0, rows], rz[0, rows]) and np.array_equal(dists[0, rows], rd[0, rows])
