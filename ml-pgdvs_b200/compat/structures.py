"""pytorch3d.structures subset: Pointclouds."""
from ..renderer import Pointclouds  # noqa: F401
