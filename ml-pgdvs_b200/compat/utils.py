"""pytorch3d.utils subset: cameras_from_opencv_projection."""
from ..renderer import cameras_from_opencv_projection  # noqa: F401
