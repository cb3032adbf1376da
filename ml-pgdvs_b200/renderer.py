"""pytorch3d-shaped facade: exactly the classes/functions PGDVS constructs at
/root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:684-722 and st_geo_renderer.py:85-120,
so `import pgdvs_b200.renderer as pytorch3d_renderer` style substitution works:

    cameras_from_opencv_projection, PerspectiveCameras, Pointclouds,
    PointsRasterizationSettings, PointsRasterizer, PointFragments, rasterize_points,
    AlphaCompositor, NormWeightedCompositor, PointsRenderer

Forward only (the reference runs under torch.no_grad: engines/evaluator_pgdvs.py:27).
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple, Union

import torch

from . import ops


# ----------------------------------------------------------------------------- structures
class Pointclouds:
    """Minimal pytorch3d.structures.Pointclouds: list of [P_i,3] tensors or a padded [N,P,3]
    tensor, with optional per-point features of the same leading shape."""

    def __init__(self, points, features=None):
        if torch.is_tensor(points):
            if points.ndim != 3 or points.shape[2] != 3:
                raise ValueError("Points tensor has incorrect dimensions.")
            pts = [points[i] for i in range(points.shape[0])]
        else:
            pts = list(points)
        if features is None:
            fts = None
        elif torch.is_tensor(features):
            fts = [features[i] for i in range(features.shape[0])]
        else:
            fts = list(features)
        if fts is not None and len(fts) != len(pts):
            raise ValueError("points and features must have the same batch size")
        self._points_list = pts
        self._features_list = fts
        self.device = pts[0].device if pts else torch.device("cpu")
        self._packed = None

    def __len__(self):
        return len(self._points_list)

    @property
    def features(self):
        return self._features_list

    @features.setter
    def features(self, value):
        # PGDVS re-assigns `.features` for the all-ones mask render (pgdvs_renderer_dyn.py:719)
        if value is None:
            self._features_list = None
        elif torch.is_tensor(value):
            self._features_list = [value[i] for i in range(value.shape[0])]
        else:
            self._features_list = list(value)

    def points_list(self):
        return self._points_list

    def points_packed(self):
        return torch.cat(self._points_list, dim=0) if self._points_list else torch.zeros(0, 3)

    def features_packed(self):
        if self._features_list is None:
            return None
        return torch.cat(self._features_list, dim=0)

    def num_points_per_cloud(self):
        return torch.tensor([p.shape[0] for p in self._points_list], dtype=torch.int64, device=self.device)

    def cloud_to_packed_first_idx(self):
        n = self.num_points_per_cloud()
        return torch.cumsum(n, dim=0) - n

    def update_packed(self, new_points_packed):
        sizes = [p.shape[0] for p in self._points_list]
        out = Pointclouds(list(torch.split(new_points_packed, sizes, dim=0)), self._features_list)
        return out


class PerspectiveCameras:
    """pytorch3d PerspectiveCameras(in_ndc=True) as produced by cameras_from_opencv_projection."""

    def __init__(self, R, T, focal_length, principal_point, image_size=None, in_ndc=True, device=None):
        if not in_ndc:
            raise NotImplementedError("only NDC-space cameras are on the PGDVS path")
        self.R = R
        self.T = T
        self.focal_length = focal_length
        self.principal_point = principal_point
        self.image_size = image_size
        self.device = device if device is not None else R.device

    def __len__(self):
        return self.R.shape[0]

    def struct_tensor(self, device):
        return ops.camera_struct_tensor(self.R, self.T, self.focal_length, self.principal_point, device)


def cameras_from_opencv_projection(R, tvec, camera_matrix, image_size) -> PerspectiveCameras:
    """pytorch3d.utils.cameras_from_opencv_projection (restated in-tree at
    /root/reference/pgdvs/utils/pytorch3d_utils.py:5-47): OpenCV (R, t, K, (h,w)) ->
    NDC PerspectiveCameras with R transposed and the x/y axes flipped."""
    focal_length = torch.stack([camera_matrix[:, 0, 0], camera_matrix[:, 1, 1]], dim=-1)
    principal_point = camera_matrix[:, :2, 2]
    image_size_wh = image_size.to(R).flip(dims=(1,))
    scale = image_size_wh.min(dim=1, keepdim=True)[0] / 2.0
    scale = scale.expand(-1, 2)
    c0 = image_size_wh / 2.0
    focal_p3d = focal_length / scale
    p0_p3d = -(principal_point - c0) / scale
    R_p3d = R.clone().permute(0, 2, 1)
    T_p3d = tvec.clone()
    R_p3d[:, :, :2] *= -1
    T_p3d[:, :2] *= -1
    return PerspectiveCameras(R=R_p3d, T=T_p3d, focal_length=focal_p3d, principal_point=p0_p3d,
                              image_size=image_size, device=R.device)


# ----------------------------------------------------------------------------- rasterizer
class PointFragments(NamedTuple):
    idx: torch.Tensor
    zbuf: torch.Tensor
    dists: torch.Tensor


class PointsRasterizationSettings:
    def __init__(self, image_size: Union[int, Tuple[int, int]] = 256, radius: Union[float, torch.Tensor] = 0.01,
                 points_per_pixel: int = 8, bin_size: Optional[int] = None,
                 max_points_per_bin: Optional[int] = None):
        self.image_size = image_size
        self.radius = radius
        self.points_per_pixel = points_per_pixel
        self.bin_size = bin_size
        self.max_points_per_bin = max_points_per_bin


def rasterize_points(pointclouds: Pointclouds, image_size=256, radius=0.01, points_per_pixel: int = 8,
                     bin_size: Optional[int] = None, max_points_per_bin: Optional[int] = None):
    """pytorch3d.renderer.points.rasterize_points (forward)."""
    return ops.rasterize_points_packed(
        pointclouds.points_packed(), pointclouds.cloud_to_packed_first_idx(),
        pointclouds.num_points_per_cloud(), image_size, radius, points_per_pixel, bin_size,
        max_points_per_bin)


class PointsRasterizer(torch.nn.Module):
    def __init__(self, cameras=None, raster_settings=None):
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else PointsRasterizationSettings()

    def transform(self, point_clouds: Pointclouds, **kwargs) -> Pointclouds:
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass")
        pts = point_clouds.points_list()
        cams = cameras.struct_tensor(pts[0].device)
        if cams.shape[0] not in (1, len(pts)):
            raise ValueError("number of cameras must be 1 or match the batch of clouds")
        out = [ops.project_points(p, cams[i if cams.shape[0] > 1 else 0]) for i, p in enumerate(pts)]
        return Pointclouds(out, point_clouds.features)

    def forward(self, point_clouds: Pointclouds, **kwargs) -> PointFragments:
        ndc = self.transform(point_clouds, **kwargs)
        s = kwargs.get("raster_settings", self.raster_settings)
        idx, zbuf, dists = rasterize_points(ndc, image_size=s.image_size, radius=s.radius,
                                            points_per_pixel=s.points_per_pixel, bin_size=s.bin_size,
                                            max_points_per_bin=s.max_points_per_bin)
        return PointFragments(idx=idx, zbuf=zbuf, dists=dists)


# ----------------------------------------------------------------------------- compositors
def _add_background_color_to_images(pix_idxs, images, background_color):
    """pytorch3d compositor._add_background_color_to_images: pixels whose nearest slot is empty
    (idx[:,0] < 0) take the background colour.  images [N,C,H,W]."""
    background_mask = pix_idxs[:, 0] < 0
    if not torch.is_tensor(background_color):
        background_color = images.new_tensor(background_color)
    if background_color.ndim == 0:
        background_color = background_color.expand(images.shape[1])
    if background_color.ndim > 1:
        raise ValueError("Wrong shape of background_color")
    background_color = background_color.to(images)
    if background_color.shape[0] + 1 == images.shape[1]:
        background_color = torch.cat([background_color, images.new_ones(1)])
    if images.shape[1] != background_color.shape[0]:
        raise ValueError("Background color has %s channels not %s" % (background_color.shape[0], images.shape[1]))
    out = images.permute(0, 2, 3, 1).clone()
    out[background_mask] = background_color
    return out.permute(0, 3, 1, 2)


class _Compositor(torch.nn.Module):
    mode = None
    fn = None

    def __init__(self, background_color=None):
        super().__init__()
        self.background_color = background_color

    def forward(self, fragments, alphas, ptclds, **kwargs):
        background_color = kwargs.get("background_color", self.background_color)
        images = type(self).fn(fragments, alphas, ptclds)
        if background_color is not None:
            return _add_background_color_to_images(fragments, images, background_color)
        return images


class AlphaCompositor(_Compositor):
    mode = "alpha"
    fn = staticmethod(ops.alpha_composite)


class NormWeightedCompositor(_Compositor):
    mode = "norm"
    fn = staticmethod(ops.norm_weighted_sum)


class PointsRenderer(torch.nn.Module):
    """pytorch3d PointsRenderer.forward: rasterize -> weights = 1 - dists/(r*r) -> compositor ->
    [N,H,W,C].  When the features have <= 4 channels and the radius is a float, everything is
    ONE fused kernel pass (bin + rasterize + composite); otherwise it falls back to the
    pytorch3d-shaped two-step (still CUDA) path."""

    def __init__(self, rasterizer, compositor):
        super().__init__()
        self.rasterizer = rasterizer
        self.compositor = compositor

    def forward(self, point_clouds: Pointclouds, **kwargs) -> torch.Tensor:
        s = self.rasterizer.raster_settings
        feats = point_clouds.features_packed()
        fused = (isinstance(self.compositor, _Compositor) and not torch.is_tensor(s.radius)
                 and feats is not None and feats.shape[1] <= 4 and not kwargs)
        if fused:
            ndc = self.rasterizer.transform(point_clouds)
            bg = self.compositor.background_color
            if bg is not None:
                bg = [float(b) for b in (bg.tolist() if torch.is_tensor(bg) else bg)]
                if len(bg) + 1 == feats.shape[1]:
                    bg = bg + [1.0]
            out = ops.render_packed(ndc.points_packed(), feats, ndc.cloud_to_packed_first_idx(),
                                    ndc.num_points_per_cloud(), s.image_size, s.radius,
                                    s.points_per_pixel, compositor=self.compositor.mode, background=bg,
                                    return_fragments=False, return_mask=False)
            img = out["image"]
            if bg is None:
                return img
            return img
        fragments = self.rasterizer(point_clouds, **kwargs)
        r = s.radius
        dists2 = fragments.dists.permute(0, 3, 1, 2)
        weights = 1 - dists2 / (r * r)
        images = self.compositor(fragments.idx.long().permute(0, 3, 1, 2), weights,
                                 feats.permute(1, 0), **kwargs)
        return images.permute(0, 2, 3, 1)
