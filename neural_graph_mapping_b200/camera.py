"""Pinhole ``Camera`` with the reference's constructor and the two methods on the render path
(ngm/camera.py:22-79, 98-116, 186-203, 215-292), evaluated by ``ngm_sample_rays``."""
from __future__ import annotations

import ctypes as C
import numbers
from typing import Optional, Tuple, Union

import torch

from . import _lib


def camera_struct(camera) -> _lib.NgmCamera:
    """NgmCamera from this Camera or the reference's own Camera object (duck-typed)."""
    fx, fy, cx0, cy0, s = camera.get_pinhole_camera_parameters(0.0)
    if s != 0:
        raise NotImplementedError("Skew != 0 not supported.")
    return _lib.NgmCamera(float(fx), float(fy), float(cx0), float(cy0), int(camera.width), int(camera.height))


class Camera:
    """Pinhole camera parameters (ngm/camera.py:15-79)."""

    def __init__(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float,
                 s: float = 0.0, pixel_center: float = 0.0) -> None:
        self.fx = fx
        self.fy = fy
        self.cx = cx - pixel_center + 0.5  # stored as pixel center 0.5 (camera.py:69-70)
        self.cy = cy - pixel_center + 0.5
        self.s = s
        if self.s != 0:
            raise NotImplementedError("Skew != 0 not supported.")
        self.width = width
        self.height = height

    def get_pinhole_camera_parameters(self, pixel_center: float) -> Tuple:
        """ngm/camera.py:98-116."""
        return self.fx, self.fy, self.cx - 0.5 + pixel_center, self.cy - 0.5 + pixel_center, self.s

    def scaled_camera(self, scale_factor: float) -> "Camera":
        """ngm/camera.py:205-213 (note: like the reference, passes the stored cx/cy on)."""
        return Camera(int(self.width * scale_factor), int(self.height * scale_factor), self.fx * scale_factor,
                      self.fy * scale_factor, self.cx * scale_factor, self.cy * scale_factor)

    def ijs_to_directions(self, ijs: torch.Tensor, convention: str = "opengl") -> torch.Tensor:
        """Row/column indices -> unit directions (ngm/camera.py:186-203)."""
        if convention != "opengl":
            raise NotImplementedError("the CUDA path implements the OpenGL convention (the renderer's)")
        pts, _ = sample_rays(self, ijs, 1, 1.0, 1.0, jitter=torch.zeros(*ijs.shape[:-1], 1, device=ijs.device))
        return pts.squeeze(-2)

    def sample_ijs_uniform(
        self,
        ijs: torch.Tensor,
        num_samples: int,
        near_distances: Optional[Union[float, torch.Tensor]] = None,
        far_distances: Optional[Union[float, torch.Tensor]] = None,
        weights: Optional[torch.Tensor] = None,
        boundaries: Optional[torch.Tensor] = None,
        convention: str = "opengl",
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Stratified samples along camera rays (ngm/camera.py:215-292, uniform branch).
        Jitter is drawn with ``torch.rand`` exactly like the reference (camera.py:274)."""
        if (weights is None) != (boundaries is None):
            raise ValueError("Either both or none of weights and boundaries must be None.")
        if boundaries is not None:
            raise NotImplementedError("weighted-bin sampling is never used by the renderer (camera.py:277-289)")
        if convention != "opengl":
            raise NotImplementedError("the CUDA path implements the OpenGL convention (the renderer's)")
        jitter = torch.rand(*ijs.shape[:-1], num_samples, device=ijs.device)
        return sample_rays(self, ijs, num_samples, near_distances, far_distances, jitter=jitter)


def _per_ray(x, leading, device, name):
    """float -> (None, scalar); tensor -> (contiguous fp32 tensor broadcast to leading, 0.0)."""
    if isinstance(x, numbers.Number):
        return None, float(x)
    t = _lib.dev_f32(x.to(device), name).expand(leading).contiguous()
    return t, 0.0


def sample_rays_args(camera, ijs, num_samples, near, far, gt=None, num_samples_guided=0, range_guided=0.0,
                     c2ws=None, jitter=None, jitter_guided=None, seed=0, offset=0, want_world=False, want_depth=False,
                     want_cam=True):
    """The ``NgmSampleArgs`` of one ``ngm_sample_rays`` call: (args, outputs, tensors the args point into)."""
    if not ijs.is_cuda:
        raise RuntimeError("ijs must be a CUDA tensor: neural_graph_mapping_b200 has no CPU path")
    dev = ijs.device
    leading = tuple(ijs.shape[:-1])
    n = 1
    for d in leading:
        n *= d
    ij = ijs.to(torch.int64).contiguous()
    G = num_samples_guided if gt is not None else 0
    St = num_samples + G
    a = _lib.NgmSampleArgs()
    a.cam = camera_struct(camera)
    a.num_rays = n
    a.ijs = ij.data_ptr()
    near_t, a.near_scalar = _per_ray(near, leading, dev, "near_distances")
    far_t, a.far_scalar = _per_ray(far, leading, dev, "far_distances")
    a.near, a.far = _lib.ptr(near_t), _lib.ptr(far_t)
    gt_t = None if gt is None else _lib.dev_f32(gt, "gt_distances").expand(leading).contiguous()
    a.gt = _lib.ptr(gt_t)
    jt = None if jitter is None else _lib.dev_f32(jitter, "jitter")
    jg = None if jitter_guided is None else _lib.dev_f32(jitter_guided, "jitter_guided")
    a.jitter, a.jitter_guided = _lib.ptr(jt), _lib.ptr(jg)
    a.seed, a.offset = seed, offset
    a.range_guided = float(range_guided)
    a.num_samples, a.num_samples_guided = num_samples, num_samples_guided
    if c2ws is None:
        c2w_t = torch.eye(4, device=dev)
        a.c2w_per_ray = 0
    else:
        c2w_t = _lib.dev_f32(c2ws, "c2ws")
        if c2w_t.numel() == 16:
            a.c2w_per_ray = 0
        else:
            c2w_t = c2w_t.expand(*leading, 4, 4).contiguous()
            a.c2w_per_ray = 1
    a.c2ws = c2w_t.data_ptr()
    # points_cam is what Camera.sample_ijs_uniform returns; the renderer itself only needs world points, distances
    # and depths (24 + 20 St bytes per ray), so its callers pass want_cam=False and get None here
    pts = torch.empty(*leading, St, 3, device=dev) if want_cam else None
    dist = torch.empty(*leading, St, device=dev)
    a.points_cam, a.distances = _lib.ptr(pts), dist.data_ptr()
    outs = [pts, dist]
    if want_world:
        pw = torch.empty(*leading, St, 3, device=dev)
        a.points_world = pw.data_ptr()
        outs.append(pw)
    if want_depth:
        dz = torch.empty(*leading, St, device=dev)
        a.depths = dz.data_ptr()
        outs.append(dz)
    return a, tuple(outs), (ij, near_t, far_t, gt_t, jt, jg, c2w_t)


def sample_rays(camera, ijs, *args, **kwargs):
    """``ngm_sample_rays``: returns (points_cam, distances[, points_world][, depths]); arguments as
    :func:`sample_rays_args`."""
    a, outs, _keep = sample_rays_args(camera, ijs, *args, **kwargs)
    dev = ijs.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ngm_sample_rays(C.byref(a), _lib.stream_ptr(dev)))
    return outs
