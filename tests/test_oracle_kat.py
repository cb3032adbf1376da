"""Hand-derived known-answer tests for the rasterizer/compositor oracle.

The reference has no tests for this path and pytorch3d is absent, so these KATs are derived
from the published pytorch3d 0.7.4 semantics restated in SURVEY.md §8a rows 7-11."""
import numpy as np
import pytest

from oracle import raster


def _one_cloud(P):
    return np.zeros(1, np.int64), np.full(1, P, np.int64)


def test_pixel_centres_square_and_non_square():
    xf, yf = raster.pixel_center_ndc(4, 4)
    # square 4x4: centres at +-0.75, +-0.25; +X is left, +Y is up
    np.testing.assert_allclose(xf, [0.75, 0.25, -0.25, -0.75], atol=1e-7)
    np.testing.assert_allclose(yf, [0.75, 0.25, -0.25, -0.75], atol=1e-7)
    xf, yf = raster.pixel_center_ndc(2, 4)  # H=2, W=4: x spans [-2,2], y spans [-1,1]
    np.testing.assert_allclose(xf, [1.5, 0.5, -0.5, -1.5], atol=1e-7)
    np.testing.assert_allclose(yf, [0.5, -0.5], atol=1e-7)
    xf, yf = raster.pixel_center_ndc(4, 2)  # portrait
    np.testing.assert_allclose(xf, [0.5, -0.5], atol=1e-7)
    np.testing.assert_allclose(yf, [1.5, 0.5, -0.5, -1.5], atol=1e-7)


def test_single_point_at_pixel_centre():
    H = W = 4
    pts = np.array([[0.25, -0.25, 2.0]], np.float32)  # centre of column 1, row 2
    idx, zbuf, dists = raster.rasterize_points(pts, *_one_cloud(1), (H, W), 0.1, 2)
    assert idx.shape == (1, H, W, 2)
    assert idx[0, 2, 1, 0] == 0 and zbuf[0, 2, 1, 0] == 2.0 and dists[0, 2, 1, 0] == 0.0
    hit = np.zeros((H, W), bool)
    hit[2, 1] = True
    assert np.all(idx[0, ~hit] == -1) and np.all(zbuf[0, ~hit] == -1) and np.all(dists[0, ~hit] == -1)
    assert idx[0, 2, 1, 1] == -1 and zbuf[0, 2, 1, 1] == -1 and dists[0, 2, 1, 1] == -1


def test_radius_test_is_strict():
    H = W = 4
    # dx = 0.25 exactly, r = 0.25 -> dist2 == r*r -> NOT a hit; slightly larger radius hits
    pts = np.array([[0.5, 0.25, 1.0]], np.float32)  # between column 0 (0.75) and 1 (0.25) on row 1
    idx, _, _ = raster.rasterize_points(pts, *_one_cloud(1), (H, W), 0.25, 1)
    assert np.all(idx == -1)
    idx, _, dists = raster.rasterize_points(pts, *_one_cloud(1), (H, W), np.nextafter(np.float32(0.25), np.float32(1)), 1)
    assert idx[0, 1, 0, 0] == 0 and idx[0, 1, 1, 0] == 0 and dists[0, 1, 0, 0] == np.float32(0.0625)


def test_behind_camera_culled_zero_depth_kept():
    H = W = 2
    pts = np.array([[0.5, 0.5, -1e-6], [0.5, 0.5, 0.0]], np.float32)
    idx, zbuf, _ = raster.rasterize_points(pts, *_one_cloud(2), (H, W), 0.1, 2)
    assert idx[0, 0, 0].tolist() == [1, -1] and zbuf[0, 0, 0, 0] == 0.0


def test_z_tie_smaller_index_first_and_k_nearest_kept():
    H = W = 2
    z = [3.0, 1.0, 2.0, 1.0, 5.0, 2.0, 0.5]
    pts = np.array([[0.5, 0.5, zz] for zz in z], np.float32)
    idx, zbuf, _ = raster.rasterize_points(pts, *_one_cloud(len(z)), (H, W), 0.2, 4)
    assert idx[0, 0, 0].tolist() == [6, 1, 3, 2]  # 0.5, 1.0(idx1), 1.0(idx3), 2.0(idx2 before idx5)
    assert zbuf[0, 0, 0].tolist() == [0.5, 1.0, 1.0, 2.0]
    assert np.all(idx[0, 1, 1] == -1)


def test_batch_uses_packed_indices_and_empty_cloud():
    H = W = 2
    pts = np.array([[0.5, 0.5, 1.0], [-0.5, -0.5, 2.0]], np.float32)
    fi = np.array([0, 1, 2], np.int64)
    npc = np.array([1, 1, 0], np.int64)
    idx, _, _ = raster.rasterize_points(pts, fi, npc, (H, W), 0.1, 1)
    assert idx[0, 0, 0, 0] == 0 and idx[1, 1, 1, 0] == 1  # packed (global) index
    assert np.all(idx[2] == -1) and (idx >= 0).sum() == 2


def test_radius_shape_error():
    with pytest.raises(ValueError):
        raster.rasterize_points(np.zeros((3, 3), np.float32), *_one_cloud(3), (2, 2),
                                np.ones(2, np.float32), 1)


@pytest.mark.parametrize("H,W,K,r", [(24, 40, 4, 0.15), (40, 24, 3, 0.08), (17, 17, 8, 0.3)])
def test_banded_checker_is_bitwise_naive(H, W, K, r):
    rng = np.random.default_rng(H * 100 + W)
    P = 4000
    s = min(H, W) / 2
    pts = np.stack([rng.uniform(-W / 2 / s - 0.2, W / 2 / s + 0.2, P),
                    rng.uniform(-H / 2 / s - 0.2, H / 2 / s + 0.2, P),
                    rng.uniform(-0.5, 5, P)], 1).astype(np.float32)
    pts[:, 2] = np.round(pts[:, 2] * 8) / 8  # many exact z ties
    fi = np.array([0, 1500], np.int64)
    npc = np.array([1500, 2500], np.int64)
    rad = rng.uniform(0.5 * r, r, P).astype(np.float32)
    a = raster.rasterize_points(pts, fi, npc, (H, W), rad, K)
    b = raster.rasterize_points(pts, fi, npc, (H, W), rad, K, n_threads=3, banded=True)
    c = raster.rasterize_points(pts, fi, npc, (H, W), rad, K, n_threads=4)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    # zbuf ascending along K where filled
    zb = np.where(a[0] >= 0, a[1], np.float32(3e38))
    assert np.all(zb[..., 1:] >= zb[..., :-1])


def test_compositors_known_answers():
    # one pixel, K=3, second slot empty
    idx = np.array([0, -1, 1], np.int64).reshape(1, 3, 1, 1)
    al = np.array([0.5, 0.9, 0.25], np.float32).reshape(1, 3, 1, 1)
    feat = np.array([[1.0, 3.0], [0.0, 8.0]], np.float32)  # [C=2, P=2]
    a = raster.composite(idx, al, feat, "alpha")
    # alpha: 0.5*1 + (1-0.5)*0.25*3 = 0.875 ; 0 + 0.5*0.25*8 = 1.0
    np.testing.assert_allclose(a.reshape(-1), [0.875, 1.0], rtol=1e-7)
    n = raster.composite(idx, al, feat, "norm")
    np.testing.assert_allclose(n.reshape(-1), [(0.5 * 1 + 0.25 * 3) / 0.75, (0.25 * 8) / 0.75], rtol=1e-6)
    w = raster.composite(idx, al, feat, "wsum")
    np.testing.assert_allclose(w.reshape(-1), [1.25, 2.0], rtol=1e-7)
    # norm: denominator clamps at 1e-4
    al2 = np.array([1e-6, 0, 0], np.float32).reshape(1, 3, 1, 1)
    idx2 = np.array([0, -1, -1], np.int64).reshape(1, 3, 1, 1)
    n2 = raster.composite(idx2, al2, feat, "norm")
    np.testing.assert_allclose(n2.reshape(-1), [1e-6 / 1e-4, 0.0], rtol=1e-6)


def test_render_points_mask_semantics():
    """mask = render(ones)[..., :1] > 0 (pgdvs_renderer_dyn.py:719-722): true exactly where some
    slot is filled with positive weight; rgb of an isolated splat equals the point colour."""
    H = W = 4
    pts = np.array([[0.25, -0.25, 2.0]], np.float32)
    col = np.array([[0.2, 0.4, 0.6]], np.float32)
    img, (idx, _, _) = raster.render_points(pts, *_one_cloud(1), col, (H, W), 0.1, 2, "norm",
                                            background=(0, 0, 0))
    ones, _ = raster.render_points(pts, *_one_cloud(1), np.ones_like(col), (H, W), 0.1, 2, "norm",
                                   background=(0, 0, 0))
    mask = ones[0, :, :, 0] > 0
    assert mask.sum() == 1 and mask[2, 1]
    np.testing.assert_allclose(img[0, 2, 1], col[0], rtol=1e-6)
    assert np.all(img[0][~mask] == 0)


def test_softsplat_forward_known_answers():
    """softsplat.py:355-393 by hand: one pixel moved by (0.25, 0.5) lands on four pixels with the
    bilinear weights; non-finite or out-of-image targets contribute nothing."""
    from oracle import pgdvs_ref as ref
    x = np.zeros((1, 1, 4, 4), np.float32)
    x[0, 0, 1, 1] = 2.0
    flow = np.zeros((1, 2, 4, 4), np.float32)
    flow[0, 0, 1, 1], flow[0, 1, 1, 1] = 0.25, 0.5
    out = ref.softsplat_forward(x, flow)
    assert out[0, 0, 1, 1] == np.float32(2.0 * 0.75 * 0.5) and out[0, 0, 1, 2] == np.float32(2.0 * 0.25 * 0.5)
    assert out[0, 0, 2, 1] == np.float32(2.0 * 0.75 * 0.5) and out[0, 0, 2, 2] == np.float32(2.0 * 0.25 * 0.5)
    assert np.isclose(out.sum(), 2.0)
    flow[0, 0, 1, 1] = np.nan
    assert ref.softsplat_forward(x, flow).sum() == 0
    flow[0, 0, 1, 1] = -10.0
    assert ref.softsplat_forward(x, flow).sum() == 0
    # target exactly on the last column: the east taps fall outside and are dropped
    flow[0, 0, 1, 1], flow[0, 1, 1, 1] = 2.0, 0.0
    out = ref.softsplat_forward(x, flow)
    assert out[0, 0, 1, 3] == 2.0 and out.sum() == 2.0


def test_mesh_rasterizer_known_answers():
    """RasterizeMeshesNaiveCpu restatement by hand: coverage is 'all barycentrics > 0', depth is the
    perspective-correct interpolation, nearer face wins, z ties go to the smaller face index,
    degenerate / behind-the-camera faces are skipped."""
    # a big triangle (z 2..4) and a small nearer one (z = 1) over the centre of a 4x4 image
    big = [[0.9, 0.9, 2.0], [-0.9, 0.9, 2.0], [0.0, -0.9, 4.0]]
    small = [[0.4, 0.4, 1.0], [-0.4, 0.4, 1.0], [0.0, -0.4, 1.0]]
    fv = np.array([big, small], np.float32)
    p2f, z, b, d = raster.rasterize_meshes(fv, (4, 4))
    assert p2f[0, 1, 0] == 0 and p2f[3, 0, 0] == -1
    assert p2f[1, 1, 0] == 1 and np.isclose(z[1, 1, 0], 1.0)  # pixel centre (0.25, 0.25) is inside the small face
    assert np.isclose(b[1, 1, 0].sum(), 1.0) and np.all(b[1, 1, 0] > 0) and d[1, 1, 0] < 0
    # perspective-correct depth: 1/z is what interpolates linearly
    w = raster.rasterize_meshes(fv[:1], (4, 4), perspective_correct=False)[2][1, 1, 0]
    zc = 1.0 / (w[0] / 2.0 + w[1] / 2.0 + w[2] / 4.0)
    assert np.isclose(raster.rasterize_meshes(fv[:1], (4, 4))[1][1, 1, 0], zc, rtol=1e-6)
    # exact tie: two copies of the same face -> the smaller index
    assert raster.rasterize_meshes(np.stack([fv[1], fv[1]]), (4, 4))[0][1, 1, 0] == 0
    # degenerate and behind-the-camera faces never cover anything
    deg = np.array([[[0.5, 0.5, 1.0], [0.5, 0.5, 1.0], [-0.5, -0.5, 1.0]]], np.float32)
    assert (raster.rasterize_meshes(deg, (4, 4))[0] >= 0).sum() == 0
    behind = fv[1:].copy()
    behind[..., 2] = -1.0
    assert (raster.rasterize_meshes(behind, (4, 4))[0] >= 0).sum() == 0
