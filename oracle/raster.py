"""ORACLE — TEST INFRASTRUCTURE ONLY (numpy/ctypes front-end of oracle/raster_cpu.cpp).

CPU restatement of pytorch3d 0.7.4's naive point rasterizer and compositors as the PGDVS
dynamic renderer calls them (/root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:684-722).
PARITY UNPINNED: the reference holds no tests or golden vectors for this path and pytorch3d
is neither vendored nor installed; see the header of raster_cpu.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (pgdvs_b200) never does.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle_raster.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/raster_cpu.cpp with the committed Makefile (g++, no FMA contraction)."""
    src = _HERE / "raster_cpu.cpp"
    if force or (not _LIB_PATH.exists()) or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        for name in ("oracle_rasterize_points_naive", "oracle_rasterize_points_banded"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [fp, lp, lp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int,
                           ip, fp, fp, ctypes.c_int]
        lib.oracle_rasterize_points_naive_rows.restype = ctypes.c_int
        lib.oracle_rasterize_points_naive_rows.argtypes = [
            fp, lp, lp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ip, fp, fp, ctypes.c_int]
        lib.oracle_composite.restype = ctypes.c_int
        lib.oracle_composite.argtypes = [lp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, fp]
        lib.oracle_rasterize_meshes_naive.restype = ctypes.c_int
        lib.oracle_rasterize_meshes_naive.argtypes = [fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_float, ctypes.c_int, ip, fp, fp, fp]
        for name in ("oracle_pixel_center_x", "oracle_pixel_center_y"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_float
            fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def rasterize_points(points, first_idx, num_pts, image_size, radius, points_per_pixel,
                     n_threads: int = 1, banded: bool = False):
    """pytorch3d `_C.rasterize_points(..., bin_size=0)` on CPU tensors.

    points [P,3] f32 NDC; first_idx/num_pts [N] i64; radius float or [P] f32.
    Returns (idx i32 [N,H,W,K], zbuf f32, dists f32), -1 filled.
    `banded=True` uses the accelerated-but-bitwise-identical checker.
    """
    lib = _load()
    points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    first_idx = np.ascontiguousarray(first_idx, dtype=np.int64).reshape(-1)
    num_pts = np.ascontiguousarray(num_pts, dtype=np.int64).reshape(-1)
    P = points.shape[0]
    N = first_idx.shape[0]
    H, W = int(image_size[0]), int(image_size[1])
    K = int(points_per_pixel)
    if np.isscalar(radius):
        radius = np.full((P,), radius, dtype=np.float32)  # _format_radius: torch.full(float32)
    radius = np.ascontiguousarray(radius, dtype=np.float32).reshape(-1)
    if radius.shape[0] != P:
        raise ValueError("Radius must be of shape (P,): got %s" % (radius.shape,))
    idx = np.empty((N, H, W, K), dtype=np.int32)
    zbuf = np.empty((N, H, W, K), dtype=np.float32)
    dists = np.empty((N, H, W, K), dtype=np.float32)
    fn = lib.oracle_rasterize_points_banded if banded else lib.oracle_rasterize_points_naive
    rc = fn(_ptr(points, ctypes.c_float), _ptr(first_idx, ctypes.c_int64),
            _ptr(num_pts, ctypes.c_int64), N, H, W, _ptr(radius, ctypes.c_float), K,
            _ptr(idx, ctypes.c_int32), _ptr(zbuf, ctypes.c_float), _ptr(dists, ctypes.c_float),
            int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle rasterize failed rc={rc}")
    return idx, zbuf, dists


def rasterize_points_rows(points, first_idx, num_pts, image_size, radius, points_per_pixel, y0, y1,
                          n_threads: int = 1):
    """Naive rasterizer restricted to image rows [y0, y1): outputs [N, y1-y0, W, K].
    bench.py times this as a bounded sample of the reference's CPU rasterization path."""
    lib = _load()
    points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    first_idx = np.ascontiguousarray(first_idx, dtype=np.int64).reshape(-1)
    num_pts = np.ascontiguousarray(num_pts, dtype=np.int64).reshape(-1)
    P, N = points.shape[0], first_idx.shape[0]
    H, W = int(image_size[0]), int(image_size[1])
    K = int(points_per_pixel)
    rad = np.full((P,), radius, dtype=np.float32)
    R = int(y1) - int(y0)
    idx = np.empty((N, R, W, K), dtype=np.int32)
    zbuf = np.empty((N, R, W, K), dtype=np.float32)
    dists = np.empty((N, R, W, K), dtype=np.float32)
    rc = lib.oracle_rasterize_points_naive_rows(
        _ptr(points, ctypes.c_float), _ptr(first_idx, ctypes.c_int64), _ptr(num_pts, ctypes.c_int64),
        N, H, W, _ptr(rad, ctypes.c_float), K, int(y0), int(y1), _ptr(idx, ctypes.c_int32),
        _ptr(zbuf, ctypes.c_float), _ptr(dists, ctypes.c_float), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle rasterize rows failed rc={rc}")
    return idx, zbuf, dists


_MODES = {"alpha": 1, "alpha_composite": 1, "norm": 2, "norm_weighted_sum": 2,
          "wsum": 3, "weighted_sum": 3}


def composite(idx_nkhw, alphas_nkhw, features_cp, mode):
    """pytorch3d alpha_composite / norm_weighted_sum / weighted_sum forward (CPU loops)."""
    lib = _load()
    idx = np.ascontiguousarray(idx_nkhw, dtype=np.int64)
    alphas = np.ascontiguousarray(alphas_nkhw, dtype=np.float32)
    feats = np.ascontiguousarray(features_cp, dtype=np.float32)
    N, K, H, W = idx.shape
    C, P = feats.shape
    out = np.zeros((N, C, H, W), dtype=np.float32)
    rc = lib.oracle_composite(_ptr(idx, ctypes.c_int64), _ptr(alphas, ctypes.c_float),
                              _ptr(feats, ctypes.c_float), N, K, H, W, C, P, _MODES[mode],
                              _ptr(out, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"oracle composite failed rc={rc}")
    return out


def pixel_center_ndc(H, W):
    """(xf[W], yf[H]) NDC pixel centres, PixToNonSquareNdc with the axis flips."""
    lib = _load()
    xf = np.array([lib.oracle_pixel_center_x(x, H, W) for x in range(W)], dtype=np.float32)
    yf = np.array([lib.oracle_pixel_center_y(y, H, W) for y in range(H)], dtype=np.float32)
    return xf, yf


def render_points(points, first_idx, num_pts, features_pc, image_size, radius, K,
                  compositor="norm", background=None, n_threads=1, banded=False):
    """PointsRenderer.forward restatement (pytorch3d renderer/points/renderer.py):
    fragments -> weights = 1 - dists/(r*r) -> compositor -> background fill -> [N,H,W,C].

    `r*r` is evaluated in Python float (double) and the tensor division happens in fp32,
    exactly as `1 - dists2 / (r * r)` does with a Python-float radius.
    """
    idx, zbuf, dists = rasterize_points(points, first_idx, num_pts, image_size, radius, K,
                                        n_threads=n_threads, banded=banded)
    r = float(radius)
    dists2 = np.transpose(dists, (0, 3, 1, 2))
    weights = (np.float32(1.0) - dists2 / np.float32(r * r)).astype(np.float32)
    idx_nkhw = np.transpose(idx, (0, 3, 1, 2)).astype(np.int64)
    feats_cp = np.ascontiguousarray(np.asarray(features_pc, dtype=np.float32).T)
    img = composite(idx_nkhw, weights, feats_cp, compositor)  # [N,C,H,W]
    if background is not None:
        bg = np.asarray(background, dtype=np.float32).reshape(-1)
        mask = idx_nkhw[:, 0] < 0  # [N,H,W]
        img = np.transpose(img, (0, 2, 3, 1)).copy()
        img[mask] = bg[: img.shape[-1]]
        return img, (idx, zbuf, dists)
    return np.transpose(img, (0, 2, 3, 1)).copy(), (idx, zbuf, dists)


def rasterize_meshes(face_verts, image_size, faces_per_pixel=1, blur_radius=0.0, perspective_correct=True):
    """pytorch3d `rasterize_meshes(..., bin_size=0)` for ONE mesh on CPU (naive algorithm).
    face_verts [F,3,3] f32 (x_ndc, y_ndc, z_view).  Returns (pix_to_face i32 [H,W,K], zbuf f32 [H,W,K],
    bary f32 [H,W,K,3], dists f32 [H,W,K]), -1 filled."""
    lib = _load()
    fv = np.ascontiguousarray(face_verts, dtype=np.float32).reshape(-1, 3, 3)
    H, W = int(image_size[0]), int(image_size[1])
    K = int(faces_per_pixel)
    p2f = np.empty((H, W, K), np.int32)
    zbuf = np.empty((H, W, K), np.float32)
    bary = np.empty((H, W, K, 3), np.float32)
    dists = np.empty((H, W, K), np.float32)
    rc = lib.oracle_rasterize_meshes_naive(_ptr(fv, ctypes.c_float), fv.shape[0], H, W, K, float(blur_radius),
                                           1 if perspective_correct else 0, _ptr(p2f, ctypes.c_int32),
                                           _ptr(zbuf, ctypes.c_float), _ptr(bary, ctypes.c_float),
                                           _ptr(dists, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"oracle rasterize_meshes failed rc={rc}")
    return p2f, zbuf, bary, dists
