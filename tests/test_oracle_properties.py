"""Internal-consistency properties of the CPU oracle's point rasterizer and compositors.

The oracle restates pytorch3d 0.7.4 from its published algorithm (no golden vectors exist for it:
"parity unpinned", oracle/raster.py).  What CAN be pinned without the library is that the
restatement behaves like the function it claims to be — against an independent brute-force
formulation in numpy, and through properties that hold for any correct K-nearest rasterizer."""
import numpy as np
import pytest

from oracle import raster as oracle


def _cloud(rng, H, W, P, zq=None):
    s = min(H, W) / 2
    pts = np.stack([rng.uniform(-W / 2 / s - 0.2, W / 2 / s + 0.2, P),
                    rng.uniform(-H / 2 / s - 0.2, H / 2 / s + 0.2, P),
                    rng.uniform(-0.3, 5.0, P)], 1).astype(np.float32)
    if zq:
        pts[:, 2] = np.round(pts[:, 2] * zq) / zq
    return pts


def _brute(pts, H, W, r, K):
    """Independent formulation: per pixel, every point with z >= 0 and dx*dx + dy*dy < r*r (fp32,
    two roundings), sorted by (z, idx), first K.  Pixel centres from the closed form of
    PixToNonSquareNdc: centre(i) = range*((S-1-i) + 0.5)/S - range/2."""
    xf, yf = oracle.pixel_center_ndc(H, W)
    r2 = np.float32(r) * np.float32(r)
    idx = -np.ones((H, W, K), np.int32)
    zb = -np.ones((H, W, K), np.float32)
    d2 = -np.ones((H, W, K), np.float32)
    px, py, pz = pts[:, 0], pts[:, 1], pts[:, 2]
    for y in range(H):
        dy = (py - yf[y]).astype(np.float32)
        for x in range(W):
            dx = (px - xf[x]).astype(np.float32)
            dist = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
            ok = np.flatnonzero((dist < r2) & ~(pz < 0))
            order = ok[np.lexsort((ok, pz[ok]))][:K]
            n = len(order)
            idx[y, x, :n], zb[y, x, :n], d2[y, x, :n] = order, pz[order], dist[order]
    return idx, zb, d2


@pytest.mark.parametrize("H,W,K,r,P,zq", [(9, 13, 4, 0.3, 400, None), (12, 8, 3, 0.25, 300, 4), (7, 7, 8, 0.6, 200, 2)])
def test_naive_rasterizer_equals_brute_force_formulation(H, W, K, r, P, zq):
    rng = np.random.default_rng(H * W + K)
    pts = _cloud(rng, H, W, P, zq)
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    idx, zb, d2 = oracle.rasterize_points(pts, fi, npc, (H, W), r, K)
    bi, bz, bd = _brute(pts, H, W, r, K)
    assert np.array_equal(idx[0], bi)
    assert np.array_equal(zb[0].view(np.int32), bz.view(np.int32))
    assert np.array_equal(d2[0].view(np.int32), bd.view(np.int32))


def test_prefix_radius_and_batch_properties():
    rng = np.random.default_rng(17)
    H, W, P = 14, 18, 1500
    pts = _cloud(rng, H, W, P, zq=8)
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    a8 = oracle.rasterize_points(pts, fi, npc, (H, W), 0.2, 8)
    a3 = oracle.rasterize_points(pts, fi, npc, (H, W), 0.2, 3)
    # the K' nearest are a prefix of the K nearest
    for big, small in zip(a8, a3):
        assert np.array_equal(big[..., :3], small)
    # sorted by (z, idx) among the filled slots
    idx, zb, _ = a8
    filled = idx >= 0
    zz = np.where(filled, zb, np.float32(np.inf))
    assert np.all(zz[..., 1:] >= zz[..., :-1])
    tie = filled[..., 1:] & (zz[..., 1:] == zz[..., :-1])
    assert tie.any() and np.all(idx[..., 1:][tie] > idx[..., :-1][tie])
    assert np.all(filled[..., :-1] | ~filled[..., 1:])  # no hole before a filled slot
    # a larger radius can only add hits
    b = oracle.rasterize_points(pts, fi, npc, (H, W), 0.3, 8)
    assert np.all((b[0] >= 0).sum(-1) >= (idx >= 0).sum(-1))
    # two clouds in a batch == two calls, with packed indices
    fi2, npc2 = np.array([0, 600], np.int64), np.array([600, 900], np.int64)
    c = oracle.rasterize_points(pts, fi2, npc2, (H, W), 0.2, 8)
    c0 = oracle.rasterize_points(pts[:600], fi, np.array([600], np.int64), (H, W), 0.2, 8)
    c1 = oracle.rasterize_points(pts[600:], fi, np.array([900], np.int64), (H, W), 0.2, 8)
    assert np.array_equal(c[0][0], c0[0][0])
    assert np.array_equal(c[0][1], np.where(c1[0][0] >= 0, c1[0][0] + 600, -1))
    assert np.array_equal(c[1][1], c1[1][0]) and np.array_equal(c[2][1], c1[2][0])


def test_compositors_against_numpy_formulas():
    rng = np.random.default_rng(23)
    N, K, H, W, C, P = 2, 5, 6, 7, 3, 50
    idx = rng.integers(-1, P, (N, K, H, W)).astype(np.int64)
    alphas = rng.uniform(0, 1, (N, K, H, W)).astype(np.float32)
    feats = rng.uniform(0, 1, (C, P)).astype(np.float32)
    valid = idx >= 0
    f = np.where(valid[:, :, None], feats[:, np.clip(idx, 0, None)].transpose(1, 2, 0, 3, 4), 0.0)  # [N,K,C,H,W]
    w = np.where(valid, alphas, 0.0).astype(np.float64)
    wsum = (w[:, :, None] * f).sum(1)
    np.testing.assert_allclose(oracle.composite(idx, alphas, feats, "wsum"), wsum, atol=1e-5)
    norm = wsum / np.maximum(w.sum(1), 1e-4)[:, None]
    np.testing.assert_allclose(oracle.composite(idx, alphas, feats, "norm"), norm, atol=1e-5)
    out = np.zeros((N, C, H, W))
    T = np.ones((N, H, W))
    for k in range(K):  # front to back; invalid slots are skipped and do not attenuate
        out += (T * w[:, k])[:, None] * f[:, k]
        T = T * np.where(valid[:, k], 1.0 - alphas[:, k], 1.0)
    np.testing.assert_allclose(oracle.composite(idx, alphas, feats, "alpha"), out, atol=1e-5)
