"""pytorch3d.ops subset: knn_points."""
from ..ops import knn_points  # noqa: F401
