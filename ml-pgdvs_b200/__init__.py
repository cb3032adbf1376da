"""pgdvs_b200 — B200-native (sm_100a) implementation of the PGDVS dynamic-content point-splat
hot path: unproject -> flow-warp -> project -> K-nearest z-buffer splat -> composite -> blend.

The directory is named `ml-pgdvs_b200/` after the reference repository; since a hyphen cannot
appear in a Python module name, `pgdvs_b200` at the repository root is a symbolic link to it:
`import pgdvs_b200` imports this package directly (no exec shim).

CUDA-only: importing the operator modules requires the built C-ABI library
(ml-pgdvs_b200/lib/libpgdvs_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _cabi  # noqa: F401
from . import ops, renderer, dyn_renderer, synthetic  # noqa: F401
from .ops import (alpha_composite, norm_weighted_sum, weighted_sum, rasterize_points_packed,  # noqa: F401
                  render_packed)
from .renderer import (AlphaCompositor, NormWeightedCompositor, PerspectiveCameras, Pointclouds,  # noqa: F401
                       PointFragments, PointsRasterizationSettings, PointsRasterizer, PointsRenderer,
                       cameras_from_opencv_projection, rasterize_points)
from .dyn_renderer import PGDVSDynamicRenderer, SourcePair, render_views  # noqa: F401
from . import track, softsplat, mesh  # noqa: F401
from .track import StaticGeoPointRenderer  # noqa: F401
