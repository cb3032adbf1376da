"""Softmax splatting on the B200 path — the reference's default `dyn_render_type`.

Mirrors the call surface of `pgdvs/utils/softsplat.py` (`softsplat(tenIn, tenFlow, tenMetric,
strMode)`, :280-334) and of `PGDVSBaseRenderer.softsplat_img` /
`backwarp_for_softsplat_metric` (`pgdvs/renderers/pgdvs_renderer_base.py:59-138`); the forward
kernel replaces the cupy-compiled string upstream (which cannot target sm_100).  Forward only,
like everything on this path (the reference renders under `torch.no_grad`).

`softsplat_dyn` is the fused form the dynamic renderer uses: noise fill of the static regions,
back-warp, importance metric, and ONE splat for rgb + mask (`pgdvs_renderer_dyn.py:157-209`).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _cabi, ops


def softsplat_forward(tenIn: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    """`softsplat_func.forward` (softsplat.py:342-427): tenIn [N,C,H,W], tenFlow [N,2,H,W]."""
    ops._require_cuda(tenIn, "tenIn")
    ops._require_cuda(tenFlow, "tenFlow")
    x, f = ops._f32c(tenIn), ops._f32c(tenFlow)
    if x.dim() != 4 or f.dim() != 4 or f.shape[1] != 2 or f.shape[0] != x.shape[0] or f.shape[2:] != x.shape[2:]:
        raise ValueError("softsplat expects tenIn [N,C,H,W] and tenFlow [N,2,H,W]")
    N, C, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().pgdvs_softsplat_forward(x.data_ptr(), f.data_ptr(), N, C, H, W, out.data_ptr(),
                                                        ops._stream_ptr(x.device)), "pgdvs_softsplat_forward")
    ops.LAUNCHES["count"] += 1
    return out


def softsplat(tenIn: torch.Tensor, tenFlow: torch.Tensor, tenMetric: Optional[torch.Tensor], strMode: str):
    """Same semantics (and assertions) as the reference function, softsplat.py:280-334."""
    base = strMode.split("-")[0]
    assert base in ["sum", "avg", "linear", "soft"]
    if strMode == "sum" or strMode == "avg":
        assert tenMetric is None
    if base in ("linear", "soft"):
        assert tenMetric is not None
    if strMode == "avg":
        tenIn = torch.cat([tenIn, tenIn.new_ones([tenIn.shape[0], 1, tenIn.shape[2], tenIn.shape[3]])], 1)
    elif base == "linear":
        tenIn = torch.cat([tenIn * tenMetric, tenMetric], 1)
    elif base == "soft":
        tenIn = torch.cat([tenIn * tenMetric.exp(), tenMetric.exp()], 1)
    tenOut = softsplat_forward(tenIn, tenFlow)
    if base in ["avg", "linear", "soft"]:
        tenNormalize = tenOut[:, -1:, :, :]
        parts = strMode.split("-")
        if len(parts) == 1 or parts[1] == "addeps":
            tenNormalize = tenNormalize + 0.0000001
        elif parts[1] == "zeroeps":
            tenNormalize = tenNormalize.clone()
            tenNormalize[tenNormalize == 0.0] = 1.0
        elif parts[1] == "clipeps":
            tenNormalize = tenNormalize.clip(0.0000001, None)
        tenOut = tenOut[:, :-1, :, :] / tenNormalize
    return tenOut


def softsplat_dyn(*, rgb_1: torch.Tensor, dyn_mask_1: torch.Tensor, rgb_2: torch.Tensor,
                  flow_1_to_tgt: torch.Tensor, flow_12: torch.Tensor, alpha: float = 100.0,
                  noise: Optional[torch.Tensor] = None, return_metric: bool = False):
    """The softsplat branch of `PGDVSDynamicRenderer.forward` (pgdvs_renderer_dyn.py:157-209) in two
    launches.  Channels-last inputs: rgb_1 / rgb_2 / noise [B,H,W,3], dyn_mask_1 [B,H,W,1] (the
    VALID dynamic mask returned by compute_dyn_pcl), flows [B,H,W,2] in pixels.  `noise` plays
    the role of `clamp(randn_like(rgb), 0, 1)` upstream (None = black static regions).
    Returns (render_dyn_rgb [B,3,H,W], render_dyn_mask [B,1,H,W][, metric [B,1,H,W]])."""
    for name, t in (("rgb_1", rgb_1), ("dyn_mask_1", dyn_mask_1), ("rgb_2", rgb_2),
                    ("flow_1_to_tgt", flow_1_to_tgt), ("flow_12", flow_12)):
        ops._require_cuda(t, name)
    dev = rgb_1.device
    r1, m1, r2 = ops._f32c(rgb_1), ops._f32c(dyn_mask_1), ops._f32c(rgb_2)
    ft, f12 = ops._f32c(flow_1_to_tgt), ops._f32c(flow_12)
    B, H, W, _ = r1.shape
    if m1.shape != (B, H, W, 1) or r2.shape != r1.shape or ft.shape != (B, H, W, 2) or f12.shape != (B, H, W, 2):
        raise ValueError("softsplat_dyn: inconsistent input shapes")
    nz = ops._f32c(noise) if noise is not None else None
    if nz is not None and nz.shape != r1.shape:
        raise ValueError("softsplat_dyn: noise must have the shape of rgb_1")
    out_rgb = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    out_mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    metric = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if return_metric else None
    L = _cabi.lib()
    nbytes = ctypes.c_size_t(0)
    _cabi.check(L.pgdvs_softsplat_workspace_bytes(B, H, W, ctypes.byref(nbytes)), "pgdvs_softsplat_workspace_bytes")
    ws = ops._WS.get(dev, nbytes.value, tag="softsplat")
    with torch.cuda.device(dev):
        _cabi.check(L.pgdvs_softsplat_dyn(
            r1.data_ptr(), m1.data_ptr(), nz.data_ptr() if nz is not None else None, r2.data_ptr(), ft.data_ptr(),
            f12.data_ptr(), float(alpha), B, H, W, out_rgb.data_ptr(), out_mask.data_ptr(),
            metric.data_ptr() if metric is not None else None, ops._aligned_ptr(ws), nbytes.value,
            ops._stream_ptr(dev)), "pgdvs_softsplat_dyn")
    ops.LAUNCHES["count"] += 2
    return (out_rgb, out_mask, metric) if return_metric else (out_rgb, out_mask)
