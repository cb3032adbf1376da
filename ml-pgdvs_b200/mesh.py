"""Mesh mode of the dynamic renderer (`dyn_render_type = mesh`) on the B200 path.

`render_dyn_mesh` mirrors `PGDVSDynamicRenderer.render_dyn_mesh`
(pgdvs/renderers/pgdvs_renderer_dyn.py:542-669): grid-topology triangles from the dynamic mask,
vertices at the target time, pytorch3d MeshRasterizer (blur 0, one face per pixel) + the
reference's SimpleShader, mask from an all-ones render — here one scatter + one resolve kernel."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _cabi, ops


def mesh_faces_from_mask(rows: torch.Tensor, cols: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Faces of render_dyn_mesh (:549-604): every valid pixel (row, col) spawns
    (r,c),(r+1,c),(r+1,c+1) and (r,c),(r+1,c+1),(r,c+1); a face survives if its corners are in
    bounds and carry a vertex index > 0 (as upstream: vertex 0 never gets a face).  All first
    triangles come before all second triangles, each in pixel order.  Index plumbing only."""
    dev = rows.device
    vert_idx = -torch.ones((H, W), dtype=torch.long, device=dev)
    vert_idx[rows, cols] = torch.arange(rows.shape[0], device=dev)
    c1 = torch.stack([torch.stack((rows, cols), 1), torch.stack((rows + 1, cols), 1),
                      torch.stack((rows + 1, cols + 1), 1)], 1)
    c2 = torch.stack([torch.stack((rows, cols), 1), torch.stack((rows + 1, cols + 1), 1),
                      torch.stack((rows, cols + 1), 1)], 1)
    cand = torch.cat((c1, c2), 0)
    inb = torch.all((cand[..., 0] >= 0) & (cand[..., 0] < H) & (cand[..., 1] >= 0) & (cand[..., 1] < W), dim=1)
    cand = cand[inb]
    fv = vert_idx[cand[..., 0], cand[..., 1]]
    return fv[torch.all(fv > 0, dim=1)].to(torch.int32).contiguous()


def rasterize_mesh(verts_ndc: torch.Tensor, faces: torch.Tensor, image_size, vert_rgb: Optional[torch.Tensor] = None,
                   perspective_correct: bool = True, return_fragments: bool = True):
    """One mesh, one face per pixel.  verts_ndc [V,3] (x_ndc, y_ndc, z_view), faces int [F,3].
    Returns dict(pix_to_face [H,W] i32, zbuf [H,W], bary [H,W,3], image [H,W,3]?, mask [H,W,1])."""
    ops._require_cuda(verts_ndc, "verts_ndc")
    dev = verts_ndc.device
    H, W = int(image_size[0]), int(image_size[1])
    v = ops._f32c(verts_ndc).reshape(-1, 3)
    f = faces.to(device=dev, dtype=torch.int32).contiguous().reshape(-1, 3)
    rgb = ops._f32c(vert_rgb).reshape(-1, 3) if vert_rgb is not None else None
    out = {}
    if return_fragments:
        out["pix_to_face"] = torch.empty((H, W), dtype=torch.int32, device=dev)
        out["zbuf"] = torch.empty((H, W), dtype=torch.float32, device=dev)
        out["bary"] = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    if rgb is not None:
        out["image"] = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    out["mask"] = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    L = _cabi.lib()
    nbytes = ctypes.c_size_t(0)
    _cabi.check(L.pgdvs_mesh_workspace_bytes(H, W, ctypes.byref(nbytes)), "pgdvs_mesh_workspace_bytes")
    ws = ops._WS.get(dev, nbytes.value, tag="mesh")
    ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
    with torch.cuda.device(dev):
        _cabi.check(L.pgdvs_rasterize_mesh(
            v.data_ptr(), v.shape[0], f.data_ptr(), f.shape[0], H, W, 1 if perspective_correct else 0,
            rgb.data_ptr() if rgb is not None else None, ptr("pix_to_face"), ptr("zbuf"), ptr("bary"), ptr("image"),
            ptr("mask"), ops._aligned_ptr(ws), nbytes.value, ops._stream_ptr(dev)), "pgdvs_rasterize_mesh")
    ops.LAUNCHES["count"] += 2
    return out


def render_dyn_mesh(*, rows, cols, dyn_mask, dyn_pcl, rgbs, flat_cam, for_debug: bool = False):
    """Same keyword arguments and return value as the reference method (:542-669):
    (mesh_img [H,W,3], mesh_mask [H,W,1])."""
    from .dyn_renderer import opencv_to_p3d_camera
    H, W, _ = dyn_mask.shape
    dev = dyn_pcl.device
    faces = mesh_faces_from_mask(rows, cols, H, W)
    if faces.shape[0] == 0:
        return torch.zeros(H, W, 3, device=dev), torch.zeros(H, W, 1, device=dev)
    fc = flat_cam.detach().cpu()
    cam = ops.camera_struct_tensor(*opencv_to_p3d_camera(fc[2:18], fc[18:34], H, W), dev)
    ndc = ops.project_points(dyn_pcl, cam)
    out = rasterize_mesh(ndc, faces, (H, W), vert_rgb=rgbs, return_fragments=False)
    return out["image"], out["mask"]
