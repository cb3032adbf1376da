"""ctypes binding of include/pgdvs_b200.h (the drop-in C ABI).  No torch types cross this
boundary: only raw device pointers, sizes and a CUDA stream handle.

There is deliberately no CPU fallback: if the shared library is missing, importing any op
raises immediately with the build command."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libpgdvs_b200.so"

COMPOSITE_NONE, COMPOSITE_ALPHA, COMPOSITE_NORM_WEIGHTED, COMPOSITE_WEIGHTED_SUM = 0, 1, 2, 3
MAX_POINTS_PER_PIXEL = 150
MAX_FUSED_CHANNELS = 4
IPC_HANDLE_BYTES = 64

# every symbol include/pgdvs_b200.h declares (tests check the library exports each one)
EXPORTED_SYMBOLS = (
    "pgdvs_abi_version", "pgdvs_error_string", "pgdvs_struct_layout", "pgdvs_bin_workspace_bytes", "pgdvs_bin_points",
    "pgdvs_rasterize_composite", "pgdvs_rasterize_composite_ex", "pgdvs_compute_projections", "pgdvs_debug_switch", "pgdvs_composite", "pgdvs_uwp_workspace_bytes",
    "pgdvs_unproject_warp_project", "pgdvs_project_points", "pgdvs_merge_blend",
    "pgdvs_uwp_bin_workspace_bytes", "pgdvs_uwp_bin", "pgdvs_pack_rgbd",
    "pgdvs_knn_workspace_bytes", "pgdvs_knn_mean_dist", "pgdvs_knn_points",
    "pgdvs_track_workspace_bytes", "pgdvs_track_points", "pgdvs_quantize_u8",
    "pgdvs_softsplat_forward", "pgdvs_softsplat_workspace_bytes", "pgdvs_softsplat_dyn",
    "pgdvs_mesh_workspace_bytes", "pgdvs_rasterize_mesh",
    "pgdvs_uwp_world_by_pixel", "pgdvs_outlier_keep",
    "pgdvs_ipc_alloc", "pgdvs_ipc_open", "pgdvs_ipc_close", "pgdvs_ipc_free", "pgdvs_copy_async",
)


class PgdvsCamera(ctypes.Structure):
    _fields_ = [("R", c_float * 9), ("T", c_float * 3), ("focal", c_float * 2), ("p0", c_float * 2)]


class PgdvsUwpJob(ctypes.Structure):
    _fields_ = [
        ("depth1", c_void_p), ("rgb1", c_void_p), ("mask1", c_void_p), ("flow12", c_void_p),
        ("occ12", c_void_p), ("depth2", c_void_p), ("rgb2", c_void_p), ("keep", c_void_p),
        ("rgbd2", c_void_p),
        ("M1", c_float * 9), ("o1", c_float * 3), ("K2inv", c_float * 9), ("R2", c_float * 9),
        ("o2", c_float * 3), ("w1", c_float), ("w2", c_float), ("same_time", c_int32),
        ("view", c_int32),
    ]


class PgdvsRasterExtra(ctypes.Structure):
    _fields_ = [("depth", c_void_p), ("image_u8", c_void_p), ("mask_u8", c_void_p)]


class PgdvsFramePack(ctypes.Structure):
    _fields_ = [("rgb", c_void_p), ("depth", c_void_p), ("rgbd", c_void_p)]


class PgdvsTrackFrame(ctypes.Structure):
    _fields_ = [("rgb", c_void_p), ("depth", c_void_p), ("M", c_float * 9), ("o", c_float * 3),
                ("time", c_float), ("_pad", c_float)]


class PgdvsError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing. pgdvs_b200 has no CPU fallback: build the sm_100a library first "
            "with `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
    L = ctypes.CDLL(str(LIB_PATH))
    L.pgdvs_abi_version.restype = c_int
    L.pgdvs_abi_version.argtypes = []
    L.pgdvs_error_string.restype = c_char_p
    L.pgdvs_error_string.argtypes = [c_int]
    L.pgdvs_struct_layout.restype = c_int
    L.pgdvs_struct_layout.argtypes = [POINTER(c_int32)]
    L.pgdvs_bin_workspace_bytes.restype = c_int
    L.pgdvs_bin_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int64, c_float, POINTER(c_size_t)]
    L.pgdvs_bin_points.restype = c_int
    L.pgdvs_bin_points.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64,
                                   c_void_p, c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]
    L.pgdvs_rasterize_composite.restype = c_int
    L.pgdvs_rasterize_composite.argtypes = [
        c_void_p, c_size_t, c_int, c_int64, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
        c_float, POINTER(c_float), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p]
    L.pgdvs_rasterize_composite_ex.restype = c_int
    L.pgdvs_rasterize_composite_ex.argtypes = [
        c_void_p, c_size_t, c_int, c_int64, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
        c_float, POINTER(c_float), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        POINTER(PgdvsRasterExtra), c_void_p]
    L.pgdvs_debug_switch.restype = c_int
    L.pgdvs_debug_switch.argtypes = [c_int, c_int]
    L.pgdvs_compute_projections.restype = c_int
    L.pgdvs_compute_projections.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.pgdvs_composite.restype = c_int
    L.pgdvs_composite.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int64, c_int, c_void_p, c_void_p]
    L.pgdvs_uwp_workspace_bytes.restype = c_int
    L.pgdvs_uwp_workspace_bytes.argtypes = [c_int, c_int, c_int, POINTER(c_size_t)]
    L.pgdvs_unproject_warp_project.restype = c_int
    L.pgdvs_unproject_warp_project.argtypes = [
        c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    L.pgdvs_uwp_bin_workspace_bytes.restype = c_int
    L.pgdvs_uwp_bin_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_float, POINTER(c_size_t)]
    L.pgdvs_uwp_bin.restype = c_int
    L.pgdvs_uwp_bin.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_void_p, c_size_t, c_void_p]
    L.pgdvs_pack_rgbd.restype = c_int
    L.pgdvs_pack_rgbd.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p]
    L.pgdvs_project_points.restype = c_int
    L.pgdvs_project_points.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    L.pgdvs_merge_blend.restype = c_int
    L.pgdvs_merge_blend.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.pgdvs_knn_workspace_bytes.restype = c_int
    L.pgdvs_knn_workspace_bytes.argtypes = [c_int64, c_int64, POINTER(c_size_t)]
    L.pgdvs_knn_mean_dist.restype = c_int
    L.pgdvs_knn_mean_dist.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p,
                                      c_void_p, c_size_t, c_void_p]
    L.pgdvs_track_workspace_bytes.restype = c_int
    L.pgdvs_track_workspace_bytes.argtypes = [c_int64, POINTER(c_size_t)]
    L.pgdvs_track_points.restype = c_int
    L.pgdvs_track_points.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, ctypes.c_uint32,
                                     ctypes.c_uint32, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]
    L.pgdvs_quantize_u8.restype = c_int
    L.pgdvs_quantize_u8.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    L.pgdvs_softsplat_forward.restype = c_int
    L.pgdvs_softsplat_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    L.pgdvs_softsplat_workspace_bytes.restype = c_int
    L.pgdvs_softsplat_workspace_bytes.argtypes = [c_int, c_int, c_int, POINTER(c_size_t)]
    L.pgdvs_softsplat_dyn.restype = c_int
    L.pgdvs_softsplat_dyn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]
    L.pgdvs_mesh_workspace_bytes.restype = c_int
    L.pgdvs_mesh_workspace_bytes.argtypes = [c_int, c_int, POINTER(c_size_t)]
    L.pgdvs_rasterize_mesh.restype = c_int
    L.pgdvs_rasterize_mesh.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]
    L.pgdvs_knn_points.restype = c_int
    L.pgdvs_knn_points.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    L.pgdvs_uwp_world_by_pixel.restype = c_int
    L.pgdvs_uwp_world_by_pixel.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_size_t, c_void_p]
    L.pgdvs_outlier_keep.restype = c_int
    L.pgdvs_outlier_keep.argtypes = [c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p, c_void_p]
    L.pgdvs_ipc_alloc.restype = c_int
    L.pgdvs_ipc_alloc.argtypes = [c_size_t, POINTER(c_void_p), c_void_p]
    L.pgdvs_ipc_open.restype = c_int
    L.pgdvs_ipc_open.argtypes = [c_void_p, POINTER(c_void_p)]
    L.pgdvs_ipc_close.restype = c_int
    L.pgdvs_ipc_close.argtypes = [c_void_p]
    L.pgdvs_ipc_free.restype = c_int
    L.pgdvs_ipc_free.argtypes = [c_void_p]
    L.pgdvs_copy_async.restype = c_int
    L.pgdvs_copy_async.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    if L.pgdvs_abi_version() != 1:
        raise ImportError("libpgdvs_b200.so ABI version mismatch; rebuild it")
    lay = (c_int32 * 4)()
    L.pgdvs_struct_layout(lay)
    mine = [ctypes.sizeof(PgdvsCamera), ctypes.sizeof(PgdvsUwpJob), PgdvsUwpJob.M1.offset, PgdvsUwpJob.view.offset]
    if list(lay) != mine:
        raise ImportError(f"struct layout mismatch between _cabi.py {mine} and libpgdvs_b200.so {list(lay)}")
    _lib = L
    return L


DEBUG_SWITCHES = {"sort_cells": 0, "force_generic": 1, "no_pair": 2}


def debug_switch(name: str, value: int):
    """Kernel-selection switches of the rasterizer (-1 automatic, 0 off, 1 on); results are the
    same bits in every setting.  For tests and A/B measurements."""
    check(lib().pgdvs_debug_switch(DEBUG_SWITCHES[name], int(value)), "pgdvs_debug_switch")


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().pgdvs_error_string(rc).decode()
        raise PgdvsError(f"{what} failed: {msg} (code {rc})")
