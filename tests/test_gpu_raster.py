"""GPU parity: bin + rasterize + composite (through the C ABI) vs the CPU oracle.

Bar: idx / zbuf / dists bit-exact given identical NDC input, mask exact,
composited images |delta| <= 1e-5."""
import numpy as np
import pytest
import torch

from oracle import raster as oracle

pytestmark = pytest.mark.gpu

IMG_ATOL = 1e-5  # stated fp32 tolerance for composited images


def _dev():
    return torch.device("cuda:0")


def _cloud(rng, H, W, P, zq=None, spill=0.25):
    s = min(H, W) / 2
    pts = np.stack([rng.uniform(-W / 2 / s - spill, W / 2 / s + spill, P),
                    rng.uniform(-H / 2 / s - spill, H / 2 / s + spill, P),
                    rng.uniform(-0.3, 6.0, P)], 1).astype(np.float32)
    if zq:
        pts[:, 2] = np.round(pts[:, 2] * zq) / zq  # exact z ties
    return pts


def _run(pts, feats, fi, npc, H, W, radius, K, compositor, static=None, bg=(0, 0, 0)):
    import pgdvs_b200
    d = _dev()
    rad = torch.from_numpy(radius).to(d) if isinstance(radius, np.ndarray) else radius
    out = pgdvs_b200.render_packed(
        torch.from_numpy(pts).to(d), torch.from_numpy(feats).to(d) if feats is not None else None,
        torch.from_numpy(fi).to(d), torch.from_numpy(npc).to(d), (H, W), rad, K, compositor=compositor,
        background=bg if compositor else None, static_rgb=static,
        rr_weight=(float(radius.max()) ** 2 if isinstance(radius, np.ndarray) and compositor else None))
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}


def _assert_frags(out, ref):
    idx, zbuf, dists = ref
    bad = np.argwhere(out["idx"] != idx)
    assert bad.shape[0] == 0, f"{bad.shape[0]} idx mismatches, first at {bad[:3].tolist()}"
    assert np.array_equal(out["zbuf"].view(np.int32), zbuf.view(np.int32)), "zbuf not bit-exact"
    assert np.array_equal(out["dists"].view(np.int32), dists.view(np.int32)), "dists not bit-exact"


@pytest.mark.parametrize("H,W,K,r,P,zq", [
    (24, 40, 4, 0.15, 3000, 8),     # landscape, many z ties, big splats
    (40, 24, 3, 0.08, 3000, None),  # portrait
    (17, 17, 8, 0.30, 2000, 4),     # square, odd size, huge splats
    (32, 64, 1, 0.05, 5000, None),  # K=1 (reference default)
    (31, 33, 16, 0.20, 4000, 16),
    (16, 16, 32, 0.50, 3000, None),
    (12, 20, 5, 0.12, 800, 2),      # K not a template size
    (9, 14, 40, 0.9, 1500, None),   # K in (32, 64]
    (8, 8, 150, 1.5, 600, 3),       # kMaxPointsPerPixel, local-memory list
])
def test_fragments_bit_exact_random(H, W, K, r, P, zq):
    rng = np.random.default_rng(H * 1000 + W * 10 + K)
    pts = _cloud(rng, H, W, P, zq)
    fi = np.array([0], np.int64)
    npc = np.array([P], np.int64)
    out = _run(pts, None, fi, npc, H, W, r, K, None)
    _assert_frags(out, oracle.rasterize_points(pts, fi, npc, (H, W), r, K, n_threads=4))


def test_batch_with_gaps_empty_and_per_point_radius():
    rng = np.random.default_rng(7)
    H, W, K = 20, 28, 6
    P = 5000
    pts = _cloud(rng, H, W, P, zq=8)
    # three clouds; the middle one is empty, and packed points 3000..3499 belong to no cloud
    fi = np.array([0, 2000, 3500], np.int64)
    npc = np.array([2000, 0, 1500], np.int64)
    rad = rng.uniform(0.03, 0.2, P).astype(np.float32)
    out = _run(pts, None, fi, npc, H, W, rad, K, None)
    _assert_frags(out, oracle.rasterize_points(pts, fi, npc, (H, W), rad, K, n_threads=4))
    assert np.all(out["idx"][1] == -1) and np.all(out["zbuf"][1] == -1) and np.all(out["dists"][1] == -1)


@pytest.mark.parametrize("compositor", ["norm", "alpha", "wsum"])
@pytest.mark.parametrize("C", [3, 4, 1])
def test_fused_composite_matches_oracle(compositor, C):
    rng = np.random.default_rng(11 + C)
    H, W, K, r, P = 26, 38, 8, 0.12, 4000
    pts = _cloud(rng, H, W, P, zq=16)
    feats = rng.uniform(0, 1, (P, C)).astype(np.float32)
    fi = np.array([0, 1500], np.int64)
    npc = np.array([1500, 2500], np.int64)
    bg = tuple([0.25, 0.5, 0.75, 1.0][:C])
    out = _run(pts, feats, fi, npc, H, W, r, K, compositor, bg=bg)
    img, frags = oracle.render_points(pts, fi, npc, feats, (H, W), r, K, compositor, background=bg, n_threads=4)
    ones, _ = oracle.render_points(pts, fi, npc, np.ones_like(feats), (H, W), r, K, compositor,
                                   background=(0,) * C, n_threads=4)
    _assert_frags(out, frags)
    np.testing.assert_allclose(out["image"], img, atol=IMG_ATOL, rtol=0)
    assert np.array_equal(out["mask"], (ones[..., :1] > 0).astype(np.float32))


def test_fused_static_blend():
    """combined = (1-mask)*static + mask*dyn (pgdvs_renderer.py:169-172) fused in the epilogue."""
    rng = np.random.default_rng(3)
    H, W, K, r, P = 22, 30, 4, 0.1, 600  # sparse: many background pixels
    pts = _cloud(rng, H, W, P)
    feats = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    static = rng.uniform(0, 1, (1, H, W, 3)).astype(np.float32)
    out = _run(pts, feats, fi, npc, H, W, r, K, "norm", static=torch.from_numpy(static).to(_dev()))
    img, _ = oracle.render_points(pts, fi, npc, feats, (H, W), r, K, "norm", background=(0, 0, 0))
    ones, _ = oracle.render_points(pts, fi, npc, np.ones_like(feats), (H, W), r, K, "norm", background=(0, 0, 0))
    m = (ones[..., :1] > 0).astype(np.float32)
    assert 0.05 < m.mean() < 0.95
    np.testing.assert_allclose(out["image"], (1 - m) * static + m * img, atol=IMG_ATOL, rtol=0)


def test_known_answers_on_gpu():
    H = W = 4
    fi, one = np.array([0], np.int64), np.array([1], np.int64)
    # single point at a pixel centre
    out = _run(np.array([[0.25, -0.25, 2.0]], np.float32), None, fi, one, H, W, 0.1, 2, None)
    assert out["idx"][0, 2, 1].tolist() == [0, -1] and out["dists"][0, 2, 1, 0] == 0.0
    assert (out["idx"] >= 0).sum() == 1
    # strict radius test: dist2 == r*r is not a hit
    p = np.array([[0.5, 0.25, 1.0]], np.float32)
    assert np.all(_run(p, None, fi, one, H, W, 0.25, 1, None)["idx"] == -1)
    r_up = float(np.nextafter(np.float32(0.25), np.float32(1)))
    o = _run(p, None, fi, one, H, W, r_up, 1, None)
    assert o["idx"][0, 1, 0, 0] == 0 and o["idx"][0, 1, 1, 0] == 0
    # pz < 0 culled, pz == 0 kept
    p = np.array([[0.5, 0.5, -1e-6], [0.5, 0.5, 0.0]], np.float32)
    o = _run(p, None, fi, np.array([2], np.int64), 2, 2, 0.1, 2, None)
    assert o["idx"][0, 0, 0].tolist() == [1, -1]
    # z ties -> smaller index first; more than K hits keeps the K nearest
    z = [3.0, 1.0, 2.0, 1.0, 5.0, 2.0, 0.5]
    p = np.array([[0.5, 0.5, zz] for zz in z], np.float32)
    o = _run(p, None, fi, np.array([len(z)], np.int64), 2, 2, 0.2, 4, None)
    assert o["idx"][0, 0, 0].tolist() == [6, 1, 3, 2]
    # no points at all
    o = _run(np.zeros((0, 3), np.float32), None, fi, np.array([0], np.int64), 3, 5, 0.1, 2, None)
    assert np.all(o["idx"] == -1) and np.all(o["zbuf"] == -1) and np.all(o["dists"] == -1)


def test_permutation_invariance_and_batch_equivalence():
    """Properties that hold at any size: the result does not depend on the packed order except
    through tie-breaking by index; a batch equals separate calls."""
    rng = np.random.default_rng(5)
    H, W, K, r, P = 36, 52, 8, 0.07, 20000
    pts = _cloud(rng, H, W, P)  # continuous z: ties have probability ~0
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    a = _run(pts, None, fi, npc, H, W, r, K, None)
    perm = rng.permutation(P)
    b = _run(pts[perm], None, fi, npc, H, W, r, K, None)
    remap = np.where(b["idx"] >= 0, perm[np.clip(b["idx"], 0, None)], -1)
    assert np.array_equal(a["idx"], remap)
    assert np.array_equal(a["zbuf"], b["zbuf"]) and np.array_equal(a["dists"], b["dists"])
    zb = np.where(a["idx"] >= 0, a["zbuf"], np.float32(3e38))
    assert np.all(zb[..., 1:] >= zb[..., :-1])
    # batch of two halves == two separate calls (indices are packed/global)
    fi2 = np.array([0, P // 2], np.int64)
    npc2 = np.array([P // 2, P - P // 2], np.int64)
    c = _run(pts, None, fi2, npc2, H, W, r, K, None)
    c0 = _run(pts[: P // 2], None, fi, np.array([P // 2], np.int64), H, W, r, K, None)
    c1 = _run(pts[P // 2:], None, fi, np.array([P - P // 2], np.int64), H, W, r, K, None)
    assert np.array_equal(c["idx"][0], c0["idx"][0])
    assert np.array_equal(c["idx"][1], np.where(c1["idx"][0] >= 0, c1["idx"][0] + P // 2, -1))


def test_config1_full_size_against_banded_oracle():
    """BASELINE config 1 at full size (288x544, P=313k, K=8, r=0.01): bit-exact fragments."""
    rng = np.random.default_rng(1234)
    H, W, K, r = 288, 544, 8, 0.01
    P = 2 * H * W
    pts = _cloud(rng, H, W, P, spill=0.02)
    feats = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    out = _run(pts, feats, fi, npc, H, W, r, K, "norm")
    img, frags = oracle.render_points(pts, fi, npc, feats, (H, W), r, K, "norm", background=(0, 0, 0),
                                      n_threads=8, banded=True)
    _assert_frags(out, frags)
    np.testing.assert_allclose(out["image"], img, atol=IMG_ATOL, rtol=0)
    assert (out["idx"][..., K - 1] >= 0).mean() > 0.5  # most pixels have all K slots filled


def test_stress_shape_properties():
    """1080x1920 (config 5 shape), 2M points, K=16: size-independent properties."""
    rng = np.random.default_rng(99)
    H, W, K, r = 1080, 1920, 16, 0.01
    P = H * W
    pts = _cloud(rng, H, W, P, spill=0.01)
    fi, npc = np.array([0], np.int64), np.array([P], np.int64)
    out = _run(pts, None, fi, npc, H, W, r, K, None)
    idx, zbuf, dists = out["idx"], out["zbuf"], out["dists"]
    filled = idx >= 0
    assert np.all(dists[filled] < np.float32(r) * np.float32(r)) and np.all(dists[filled] >= 0)
    assert np.all((zbuf == -1) == ~filled) and np.all(zbuf[filled] == pts[idx[filled], 2])
    zb = np.where(filled, zbuf, np.float32(3e38))
    assert np.all(zb[..., 1:] >= zb[..., :-1])
    # filled slots are a prefix of the K list
    assert np.all(filled[..., :-1] >= filled[..., 1:])
    # spot-check 64 random rows against the naive oracle restricted to nearby points
    rows = rng.choice(H, 6, replace=False)
    _, yf = oracle.pixel_center_ndc(H, W)
    for y in rows:
        near = np.abs(pts[:, 1] - yf[y]) < 2 * r
        sub = np.nonzero(near)[0]
        ri, rz, rd = oracle.rasterize_points(pts[sub], fi, np.array([sub.size], np.int64), (H, W), r, K, n_threads=8, banded=True)
        ref_idx = np.where(ri[0, y] >= 0, sub[np.clip(ri[0, y], 0, None)], -1)
        assert np.array_equal(idx[0, y], ref_idx)
        assert np.array_equal(zbuf[0, y], rz[0, y]) and np.array_equal(dists[0, y], rd[0, y])


def test_argument_errors():
    import pgdvs_b200
    d = _dev()
    pts = torch.zeros(4, 3, device=d)
    fi = torch.zeros(1, dtype=torch.int64, device=d)
    npc = torch.full((1,), 4, dtype=torch.int64, device=d)
    with pytest.raises(ValueError):
        pgdvs_b200.rasterize_points_packed(pts, fi, npc, (8, 8), 0.1, 151)
    with pytest.raises(ValueError):
        pgdvs_b200.rasterize_points_packed(pts, fi, npc, (8, 8), torch.ones(3, device=d), 2)
    with pytest.raises(ValueError):
        pgdvs_b200.rasterize_points_packed(pts, fi, npc, (8, 0), 0.1, 2)
    with pytest.raises(RuntimeError):
        pgdvs_b200.rasterize_points_packed(pts.cpu(), fi, npc, (8, 8), 0.1, 2)  # no CPU fallback


# ---------------------------------------------------------------------------------------------
# The TMA-staged tile kernel (halo 1..3, scalar radius) and its rare paths.  r_px ~ 1.4 at 96x128.
# ---------------------------------------------------------------------------------------------
def _tile_case(rng, H, W, P, z_mode, cluster=0.0):
    s = min(H, W) / 2
    x = rng.uniform(-W / 2 / s, W / 2 / s, P)
    y = rng.uniform(-H / 2 / s, H / 2 / s, P)
    if cluster > 0:  # a dense blob: its tiles overflow the staging buffer (sized from the MEAN density)
        n = int(P * cluster)
        x[:n] = rng.uniform(0.10, 0.10 + 16 / s, n)
        y[:n] = rng.uniform(-0.20, -0.20 + 16 / s, n)
    if z_mode == "wide":      # depth ratio 1e9 inside every tile: no room for exact 32-bit keys
        z = 10.0 ** rng.uniform(-3, 6, P)
    elif z_mode == "ties":    # four depth planes: almost every pixel has exact z ties
        z = rng.integers(1, 5, P).astype(np.float64)
    elif z_mode == "zero":    # +0.0 / -0.0 / denormals next to ordinary depths
        z = rng.uniform(0.5, 4.0, P)
        z[::7] = 0.0
        z[3::7] = -0.0
        z[5::7] = 1e-42
    else:
        z = rng.uniform(1.0, 10.0, P)
    return np.stack([x, y, z], 1).astype(np.float32)


@pytest.mark.parametrize("z_mode,cluster,K,P", [
    ("smooth", 0.0, 8, 30000),   # the fast path: exact 32-bit keys
    ("wide", 0.0, 8, 30000),     # truncated keys + ambiguity rescan
    ("ties", 0.0, 8, 30000),     # exact ties everywhere -> rescan_exact on most pixels
    ("zero", 0.0, 4, 20000),     # signed zeros and denormal depths
    ("smooth", 0.6, 8, 40000),   # overflowing tiles: unstaged walk in the same launch
    ("ties", 0.5, 5, 30000),     # both at once, K not a template size
    ("smooth", 0.0, 16, 30000),  # long lists inside the tile kernel (pair list)
    ("wide", 0.3, 32, 30000),
    ("ties", 0.0, 2, 20000),     # K = 2 and K = 3 (KP = 2 / 4 of the pair kernel, K < KP)
    ("smooth", 0.3, 3, 20000),
])
@pytest.mark.parametrize("kernel", ["tile", "pair"])
def test_tile_kernel_paths_bit_exact(z_mode, cluster, K, P, kernel):
    """Both staged kernels (k_raster_tile: 32x8 tiles, one pixel per thread; k_raster_pair: 32x8
    tiles (32x16 before session 3), a vertical pixel pair per thread, K in {2, 4, 8}) on every special path."""
    from pgdvs_b200 import _cabi
    _cabi.debug_switch("no_pair", 1 if kernel == "tile" else 0)
    try:
        _tile_kernel_case(z_mode, cluster, K, P)
    finally:
        _cabi.debug_switch("no_pair", -1)


@pytest.mark.parametrize("H,W", [(93, 121), (50, 70), (17, 33)])
@pytest.mark.parametrize("kernel", ["tile", "pair"])
def test_staged_kernels_on_ragged_images(H, W, kernel):
    """Image sizes that are no multiple of the tile (32 x 8) or of the pixel pair: the last tile
    row / column is partial and the last pixel row may have no partner."""
    from pgdvs_b200 import _cabi
    _cabi.debug_switch("no_pair", 1 if kernel == "tile" else 0)
    try:
        _tile_kernel_case("smooth", 0.0, 8, 12000, H=H, W=W, r=1.45 / (min(H, W) / 2))
        _tile_kernel_case("ties", 0.2, 4, 9000, H=H, W=W, r=1.45 / (min(H, W) / 2))
    finally:
        _cabi.debug_switch("no_pair", -1)


def _tile_kernel_case(z_mode, cluster, K, P, H=96, W=128, r=0.029):
    rng = np.random.default_rng(sum(map(ord, z_mode)) * 1000 + K * 10 + int(cluster * 10))
    pts = _tile_case(rng, H, W, P, z_mode, cluster)
    feats = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    fi = np.array([0, P // 3], np.int64)
    npc = np.array([P // 3, P - P // 3], np.int64)
    out = _run(pts, feats, fi, npc, H, W, r, K, "norm")
    img, frags = oracle.render_points(pts, fi, npc, feats, (H, W), r, K, "norm", background=(0, 0, 0), n_threads=4)
    _assert_frags(out, frags)
    np.testing.assert_allclose(out["image"], img, atol=IMG_ATOL, rtol=0)


@pytest.mark.parametrize("sort,K,r", [("1", 8, 0.25), ("0", 8, 0.25), ("1", 3, 0.4), ("1", 40, 0.6)])
def test_generic_kernel_with_z_sorted_cells(sort, K, r):
    """The generic kernel forced onto clouds the tile kernel would take, with and without the
    in-place z-sort of the small cells (k_sort_cells): exact ties, -0.0 / huge z, cells larger
    than the sort cap, per-point radii and a fused compositor all match the oracle bit for bit."""
    from pgdvs_b200 import _cabi
    _cabi.debug_switch("force_generic", 1)
    _cabi.debug_switch("sort_cells", int(sort))
    try:
        _generic_kernel_case(K, r)
    finally:
        _cabi.debug_switch("force_generic", -1)
        _cabi.debug_switch("sort_cells", -1)


def _generic_kernel_case(K, r):
    rng = np.random.default_rng(41 + K)
    H, W, P = 18, 26, 6000  # ~13 points per pixel: cells of 2..30 records
    pts = _cloud(rng, H, W, P, zq=4)
    pts[:40, :2] = pts[0, :2] + rng.uniform(-0.01, 0.01, (40, 2)).astype(np.float32)  # one cell > cap
    pts[40:50, 2] = np.float32(-0.0)
    pts[50:55, 2] = np.float32(3e38)
    feats = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    fi = np.array([0, 2500], np.int64)
    npc = np.array([2500, 3500], np.int64)
    out = _run(pts, feats, fi, npc, H, W, r, K, "norm")
    img, frags = oracle.render_points(pts, fi, npc, feats, (H, W), r, K, "norm", background=(0, 0, 0), n_threads=4)
    _assert_frags(out, frags)
    np.testing.assert_allclose(out["image"], img, atol=IMG_ATOL, rtol=0)
    rad = rng.uniform(0.05, r, P).astype(np.float32)
    out = _run(pts, None, fi, npc, H, W, rad, K, None)
    _assert_frags(out, oracle.rasterize_points(pts, fi, npc, (H, W), rad, K, n_threads=4))
