"""pytorch3d.renderer subset used by PGDVS."""
from ..renderer import (AlphaCompositor, NormWeightedCompositor, PerspectiveCameras, PointFragments,  # noqa: F401
                        PointsRasterizationSettings, PointsRasterizer, PointsRenderer, rasterize_points)
