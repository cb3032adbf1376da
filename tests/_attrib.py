"""Mismatch attribution for end-to-end comparisons that cross the projection.

When the GPU path and the oracle each build their own NDC cloud (fused kernels vs torch CPU ops),
the coordinates agree only to the stated point tolerance, so a point that sits on the rim of a
pixel's splat disc, or two points at (almost) the same depth, can legitimately end up on different
sides.  Every differing pixel must be explained by one of those two causes, computed from the
oracle's own cloud; anything else fails the test.

    boundary flip : some point's distance to the pixel centre is within `d_tol` of the radius
    z tie         : two of the nearest in-radius points (first K+1 by depth) are closer than z_rtol
"""
import numpy as np

from oracle import raster as oracle


def explain_pixel(ndc, xf, yf, radius, K, d_tol, z_rtol):
    dx = ndc[:, 0] - xf
    dy = ndc[:, 1] - yf
    dist = np.sqrt(dx.astype(np.float64) ** 2 + dy.astype(np.float64) ** 2)
    ok = ndc[:, 2] >= 0
    if np.any(ok & (np.abs(dist - radius) <= d_tol)):
        return "boundary"
    z = np.sort(ndc[ok & (dist < radius), 2].astype(np.float64))[: K + 1]
    if z.size >= 2 and np.any(np.diff(z) <= z_rtol * np.maximum(np.abs(z[1:]), 1.0)):
        return "z-tie"
    return None


def attribute_mismatches(mismatch_hw, ndc, H, W, radius, K, d_tol=4e-5, z_rtol=2e-5):
    """mismatch_hw: bool [H,W] pixels where the two results differ; ndc [P,3] the oracle's cloud of
    that view.  Returns dict(counts per cause); raises AssertionError for an unexplained pixel."""
    xf, yf = oracle.pixel_center_ndc(H, W)
    counts = {"boundary": 0, "z-tie": 0}
    ys, xs = np.nonzero(mismatch_hw)
    for y, x in zip(ys.tolist(), xs.tolist()):
        why = explain_pixel(ndc, float(xf[x]), float(yf[y]), float(radius), K, d_tol, z_rtol)
        assert why is not None, (f"pixel ({y},{x}) differs from the oracle and is neither a boundary flip "
                                 f"(|dist - r| <= {d_tol}) nor a z tie (rel {z_rtol})")
        counts[why] += 1
    return counts
