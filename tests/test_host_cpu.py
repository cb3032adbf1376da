"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol
the header declares, argument validation, host camera algebra vs the reference fixtures,
synthetic workload shapes, view sharding."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from pgdvs_b200 import _cabi
    if not _cabi.LIB_PATH.exists():
        g.build()
    return _cabi.lib()


def test_library_exports_every_declared_symbol(lib):
    from pgdvs_b200 import _cabi
    header = (ROOT / "include" / "pgdvs_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pgdvs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pgdvs_b200.h but not exported"
    assert lib.pgdvs_abi_version() == 1


def _ctype_class(t):
    """pointer / i32 / i64 / u32 / size / f32 class of a ctypes argtype."""
    if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int))) and \
            issubclass(t, ctypes._Pointer):
        return "ptr"
    return {ctypes.c_int: "i32", ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_uint32: "u32",
            ctypes.c_size_t: "size", ctypes.c_float: "f32"}[t]


def _c_class(decl):
    if "*" in decl or "[" in decl:
        return "ptr"
    for key, cls in (("size_t", "size"), ("int64_t", "i64"), ("uint32_t", "u32"), ("float", "f32"),
                     ("int32_t", "i32"), ("int", "i32")):
        if re.search(r"\b%s\b" % key, decl):
            return cls
    raise AssertionError(f"unparsed parameter {decl!r}")


def test_ctypes_prototypes_match_header(lib):
    """Every prototype in include/pgdvs_b200.h against the argtypes _cabi.py binds it with:
    same arity, and per argument the same class (pointer, int32, int64, uint32, size_t, float)."""
    header = (ROOT / "include" / "pgdvs_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(?:int|const char\s*\*)\s*(pgdvs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 20
    checked = 0
    for name, params in protos:
        params = " ".join(params.split())
        want = [] if params in ("", "void") else [_c_class(x.strip()) for x in params.split(",")]
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name}: no argtypes bound in _cabi.py"
        got = [_ctype_class(t) for t in fn.argtypes]
        assert got == want, f"{name}: header {want} vs _cabi.py {got}"
        checked += 1
    assert checked == len(protos)


def test_error_strings_and_host_side_validation(lib):
    assert lib.pgdvs_error_string(0) == b"ok"
    assert b"150" in lib.pgdvs_error_string(-2)
    n = ctypes.c_size_t(0)
    assert lib.pgdvs_bin_workspace_bytes(1, 288, 544, 313344, 0.01, ctypes.byref(n)) == 0
    # counters for (288+2)*(544+2) cells + cell id (4 B) + two float4 records per point
    assert n.value >= 290 * 546 * 4 + 313344 * (4 + 16 + 16)
    assert lib.pgdvs_bin_workspace_bytes(1, 0, 544, 10, 0.01, ctypes.byref(n)) == -1
    assert lib.pgdvs_bin_workspace_bytes(1, 8, 8, 10, -1.0, ctypes.byref(n)) == -1
    assert lib.pgdvs_uwp_workspace_bytes(4, 288, 544, ctypes.byref(n)) == 0 and n.value > 0
    # argument errors are detected before any CUDA call (safe without a GPU)
    assert lib.pgdvs_rasterize_composite(None, 0, 1, 0, 8, 8, 4, 0.1, 0, 0, 0, 1.0, None, None, None,
                                         None, None, None, None, None) == -1
    dummy = ctypes.c_void_p(256)
    assert lib.pgdvs_rasterize_composite(dummy, 1 << 40, 1, 0, 8, 8, 151, 0.1, 0, 0, 0, 1.0, None, None,
                                         None, None, None, None, None, None) == -2
    assert lib.pgdvs_rasterize_composite(dummy, 16, 1, 0, 8, 8, 4, 0.1, 0, 0, 0, 1.0, None, None, None,
                                         None, None, None, None, None) == -3
    assert lib.pgdvs_knn_mean_dist(None, 5, None, 5, 65, 0, None, None, 0, None) == -6  # PGDVS_E_KNN_K
    assert b"64" in lib.pgdvs_error_string(-6)
    # more records than int32 float4 indices can address are refused up front
    assert lib.pgdvs_bin_workspace_bytes(1, 8, 8, 1 << 30, 0.1, ctypes.byref(n)) == -1
    assert lib.pgdvs_rasterize_composite(dummy, 1 << 40, 1, 1 << 30, 8, 8, 4, 0.1, 0, 0, 0, 1.0, None, None,
                                         None, None, None, None, None, None) == -1
    # softsplat / mesh / knn-grid entry points validate before touching the device
    assert lib.pgdvs_softsplat_workspace_bytes(2, 8, 8, ctypes.byref(n)) == 0 and n.value == 2 * 8 * 8 * 8 * 4
    assert lib.pgdvs_softsplat_dyn(None, None, None, None, None, None, 100.0, 1, 8, 8, None, None, None,
                                   None, 0, None) == -1
    assert lib.pgdvs_mesh_workspace_bytes(8, 12, ctypes.byref(n)) == 0 and n.value == 8 * 12 * 8
    assert lib.pgdvs_rasterize_mesh(None, 3, None, 1, 8, 8, 1, None, None, None, None, None, None, None, 0,
                                    None) == -1
    assert lib.pgdvs_knn_workspace_bytes(100, 100, ctypes.byref(n)) == 0 and n.value == 256
    assert lib.pgdvs_knn_workspace_bytes(100000, 100000, ctypes.byref(n)) == 0 and n.value > 16 << 20


def test_struct_layout_matches_library(lib):
    from pgdvs_b200 import _cabi
    lay = (ctypes.c_int32 * 4)()
    assert lib.pgdvs_struct_layout(lay) == 0
    assert list(lay) == [ctypes.sizeof(_cabi.PgdvsCamera), ctypes.sizeof(_cabi.PgdvsUwpJob),
                         _cabi.PgdvsUwpJob.M1.offset, _cabi.PgdvsUwpJob.view.offset]
    assert lay[0] == 64 and lay[2] == 72  # 9 pointers precede M1


def test_host_camera_matches_reference_fixture(golden_dir):
    from pgdvs_b200.dyn_renderer import opencv_to_p3d_camera
    g = np.load(golden_dir / "camera.npz")
    fc = g["flat_cam"]
    R, T, f, p0 = opencv_to_p3d_camera(fc[2:18], fc[18:34], int(fc[0]), int(fc[1]))
    np.testing.assert_allclose(R.reshape(3, 3), g["R"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(T, g["T"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(f, g["focal"], rtol=1e-6)
    np.testing.assert_allclose(p0, g["p0"], rtol=1e-6, atol=1e-7)


def test_facade_camera_conversion_matches_fixture(golden_dir):
    import pgdvs_b200 as p3
    g = np.load(golden_dir / "camera.npz")
    fc = torch.from_numpy(g["flat_cam"])
    w2c = torch.inverse(fc[18:34].reshape(4, 4))
    cams = p3.cameras_from_opencv_projection(w2c[None, :3, :3], w2c[None, :3, 3], fc[2:18].reshape(4, 4)[None, :3, :3],
                                             torch.LongTensor([[int(fc[0]), int(fc[1])]]))
    np.testing.assert_allclose(cams.R[0].numpy(), g["R"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cams.T[0].numpy(), g["T"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cams.focal_length[0].numpy(), g["focal"], rtol=1e-6)
    np.testing.assert_allclose(cams.principal_point[0].numpy(), g["p0"], rtol=1e-6, atol=1e-7)


def test_no_cpu_fallback():
    import pgdvs_b200
    pts = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        pgdvs_b200.rasterize_points_packed(pts, torch.zeros(1, dtype=torch.int64), torch.full((1,), 4), (8, 8), 0.1, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        pgdvs_b200.norm_weighted_sum(torch.zeros(1, 1, 2, 2, dtype=torch.int64), torch.zeros(1, 1, 2, 2), torch.zeros(3, 4))


def test_product_package_never_imports_oracle():
    for p in (ROOT / "ml-pgdvs_b200").rglob("*.py"):
        src = p.read_text()
        assert "oracle" not in re.sub(r"#.*", "", src).replace("oracle/", ""), f"{p} references the oracle"


def test_parse_image_size_and_settings():
    from pgdvs_b200.ops import parse_image_size
    assert parse_image_size(64) == (64, 64) and parse_image_size((3, 5)) == (3, 5)
    for bad in [(0, 5), (3, 5, 7), (3.5, 4)]:
        with pytest.raises(ValueError):
            parse_image_size(bad)


def test_pointclouds_structure():
    import pgdvs_b200 as p3
    a, b = torch.rand(5, 3), torch.rand(7, 3)
    pc = p3.Pointclouds([a, b], [torch.rand(5, 3), torch.rand(7, 3)])
    assert pc.points_packed().shape == (12, 3)
    assert pc.num_points_per_cloud().tolist() == [5, 7] and pc.cloud_to_packed_first_idx().tolist() == [0, 5]
    pc.features = torch.ones(2, 5, 3)[:, :5]
    assert pc.features_packed().shape[1] == 3


@pytest.mark.parametrize("name,P", [("c1_nvidia_1view", 313344), ("c3_iphone", 1036800), ("c4_davis", 819840)])
def test_synthetic_workload_shapes(name, P):
    from pgdvs_b200 import synthetic
    cfg = synthetic.CONFIGS[name]
    assert cfg["S"] * cfg["H"] * cfg["W"] == P  # BASELINE.md point counts
    wl = synthetic.make_workload("tiny", torch.device("cpu"), n_views=5)
    assert wl.n_views == 5 and len(wl.view_pairs[0]) == 2 and wl.static_rgb.shape == (5, 24, 40, 3)
    pairs, cams = wl.jobs([4, 1])
    assert [p.view for p in pairs] == [0, 0, 1, 1] and len(cams) == 2


def test_batched_camera_conversion_matches_single():
    """prepare_views converts all target cameras with one batched inverse: same bits as one by one."""
    from pgdvs_b200.dyn_renderer import opencv_to_p3d_camera, opencv_to_p3d_cameras
    rng = np.random.default_rng(2)
    n, H, W = 7, 288, 544
    Ks = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    Ks[:, 0, 0] = 480 + rng.random(n).astype(np.float32) * 40
    Ks[:, 1, 1] = 470
    Ks[:, 0, 2], Ks[:, 1, 2] = W / 2 + 3, H / 2 - 2
    c2w = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    ang = rng.random(n).astype(np.float32) * 0.3
    c2w[:, 0, 0], c2w[:, 0, 2], c2w[:, 2, 0], c2w[:, 2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
    c2w[:, :3, 3] = rng.random((n, 3)).astype(np.float32)
    batch = opencv_to_p3d_cameras(Ks, c2w, H, W)
    assert batch.shape == (n, 16) and batch.dtype == np.float32
    for i in range(n):
        one = np.concatenate(opencv_to_p3d_camera(Ks[i], c2w[i], H, W))
        assert np.array_equal(one.view(np.int32), batch[i].view(np.int32))


def test_job_descriptor_cache_is_invalidated_by_keep():
    """SourcePair caches its C descriptor; attaching an outlier mask must rebuild it."""
    import ctypes
    from pgdvs_b200 import _cabi
    from pgdvs_b200.dyn_renderer import SourcePair
    H, W = 4, 6
    z = torch.zeros
    p = SourcePair(depth_1=z(H, W, 1), rgb_1=z(H, W, 3), mask_1=z(H, W, 1), flow_12=z(H, W, 2), depth_2=z(H, W, 1),
                   rgb_2=z(H, W, 3), K_1=torch.eye(4), c2w_1=torch.eye(4), K_2=torch.eye(4), c2w_2=torch.eye(4),
                   time_1=0.0, time_2=1.0, time_tgt=0.5, view=3)
    rec0, key0 = p.record().copy(), p.group_key()
    assert rec0.shape == (ctypes.sizeof(_cabi.PgdvsUwpJob),)
    off = _cabi.PgdvsUwpJob.keep.offset
    assert int(rec0[off:off + 8].view(np.uint64)[0]) == 0
    keep = torch.ones(H * W, dtype=torch.uint8)
    p.keep = keep
    rec1 = p.record()
    assert int(rec1[off:off + 8].view(np.uint64)[0]) == keep.data_ptr()
    assert p.group_key() != key0
    voff = _cabi.PgdvsUwpJob.view.offset
    assert int(rec1[voff:voff + 4].view(np.int32)[0]) == 3


def test_mesh_faces_match_oracle_construction():
    """Grid-topology faces of mesh mode: same list, same order, same vertex-0 quirk as the oracle's
    restatement of pgdvs_renderer_dyn.py:549-604 (index plumbing, runs on any device)."""
    from oracle import pgdvs_ref as ref
    from pgdvs_b200.mesh import mesh_faces_from_mask
    g = torch.Generator().manual_seed(6)
    H, W = 9, 13
    mask = torch.rand(H, W, generator=g) < 0.7
    rows, cols = torch.nonzero(mask, as_tuple=True)
    a = mesh_faces_from_mask(rows, cols, H, W)
    b = ref.mesh_faces_from_mask(rows, cols, H, W)
    assert a.dtype == torch.int32 and torch.equal(a.long(), b) and b.shape[0] > 10
    assert int((a == 0).sum()) == 0  # vertex 0 never appears in a face (upstream uses `> 0`)


def test_prepare_data_and_resize_match_the_restatement():
    """PGDVSDynamicTrackRenderer.prepare_data (pgdvs_renderer_dyn_track.py:599-764) and
    resize_rgb_mask (pgdvs_renderer_dyn.py:259-270) are host glue: same keys, values and index
    lists as the oracle restatement on a reference-shaped synthetic data dict (window truncated at
    both ends of the video, padded frames ignored through n_actual_*)."""
    import torch
    from oracle import pgdvs_ref as ref
    from pgdvs_b200 import synthetic, track
    data = synthetic.make_data_dict("tiny_track", torch.device("cpu"), n_views=7, n_track_one_side=3)
    n_views = 3 * 2 + 2
    seen = set()
    for b in range(7):
        a, e = track.prepare_data(b, data, n_views), ref.prepare_data(b, data, n_views)
        assert a.keys() == e.keys()
        for k in a:
            if torch.is_tensor(a[k]):
                assert torch.equal(a[k], e[k]), k
            else:
                assert a[k] == e[k], k
        assert a["rgbs_for_track"].shape[0] == n_views and a["depths_for_track"].shape[0] == a["n_actual_frames"]
        assert float(a["time_for_track"].min()) == 0.0
        seen.add((len(a["idx_real_track_fwd"]), len(a["idx_real_track_bwd"])))
    assert (0, 3) in seen and (3, 3) in seen and (3, 0) in seen  # start, middle and end of the video
    # the synthetic tracker keeps the reference's conventions: (t,row,col) queries on the real-track
    # frames' dynamic pixels, (col,row) tracks that pass through the query, visible at the query frame
    dft = track.prepare_data(3, data, n_views)
    q, tr, vis = synthetic.SyntheticTracker(seed=1)(dft)
    n = dft["n_actual_frames"]
    assert tr.shape == (q.shape[0], n, 2) and vis.shape == (q.shape[0], n) and vis.dtype == torch.bool
    assert set(q[:, 0].long().tolist()) == set(dft["idx_real_track"])
    ar = torch.arange(q.shape[0])
    assert torch.equal(tr[ar, q[:, 0].long()], torch.stack((q[:, 2], q[:, 1]), 1)) and bool(vis[ar, q[:, 0].long()].all())
    # resize: the same two interpolate calls as upstream
    g = torch.Generator().manual_seed(0)
    rgb, mask = torch.rand(2, 3, 24, 40, generator=g), (torch.rand(2, 1, 24, 40, generator=g) > 0.5).float()
    from pgdvs_b200.dyn_renderer import PGDVSDynamicRenderer
    a_rgb, a_mask = PGDVSDynamicRenderer.resize_rgb_mask(rgb, mask, 36, 60)
    e_rgb, e_mask = ref.resize_rgb_mask(rgb, mask, 36, 60)
    assert torch.equal(a_rgb, e_rgb) and torch.equal(a_mask, e_mask) and a_rgb.shape == (2, 3, 36, 60)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port on the host cores) prints
    ONE JSON line with the keys the driver reads; no GPU, no CUDA library involved."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-step-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rendered_novel_views_per_s" and d["unit"] == "views/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == "c2_nvidia_seq"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
