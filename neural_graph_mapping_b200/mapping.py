"""BASELINE config 3: the mapping loop around the render path, driven by a synthetic RGB-D stream with gt poses.

The reference's SLAM driver (``NeuralGraphMap.fit``) is out of scope -- it stays an untouched caller of the render
path.  What this module provides is the smallest loop that exercises that path the way the driver does,
so the training call can be run, timed and checked on a GPU box that has neither the reference nor its datasets:

* ``SyntheticStream``  -- an analytic scene (a box room with a sphere in it) observed by a camera on a circle, with the
  NRGBD intrinsics, ground-truth poses and a keyframe every ``keyframe_every`` frames (the reference's
  ``pose_source: gt`` / ``pg_source: fixed_kf_freq``, ngm/slam_dataset.py:164-166, 407-422);
* ``MappingLoop``      -- ``RenderState`` plus the callers either side of the render, each mirroring the reference method of
  the same name: ``_extend_global_map_dict`` (ngm/run_mapping.py:267-345: cover the depth image with fields on a shifted
  grid), ``_add_fields`` (:365-389), ``_update_slam_state`` (:1599-1641, multi_view mode), ``_update_mv_training_data``
  (:1686-1713), ``_optimization_iteration`` (:1124-1182: ``_sample_target_mv`` -> ``_render_ijs`` under autograd ->
  ``_compute_losses`` -> ``_update_step``), ``_compute_losses`` (:1769-1871 with ngm/losses.py:10-78) and
  ``_current_frame_optimization`` (:1224-1251: ``num_iterations_per_frame`` iterations per input frame, wall-clock fps).

Everything data-parallel inside an iteration runs in libngm_b200 (target sampling, render forward/backward, Adam); the
losses are a dozen masked reductions in torch, as in the reference.
"""
from __future__ import annotations

import math
import time
from typing import Dict, Optional

import torch

from . import optim
from .renderer import RenderState
from .targets import get_observed_fields, sample_target_mv


class SyntheticStream:
    """Analytic RGB-D stream: camera inside the box [-hx, hx] x [-hy, hy] x [-hz, hz] that contains a sphere; the
    camera moves on a circle of radius ``orbit`` around the y axis looking at the sphere.  Colours are smooth functions
    of the surface point, depth is z-depth in metres (what the reference's datasets deliver, ngm/slam_dataset.py:95-107)."""

    def __init__(self, camera, device, num_frames: int = 200, keyframe_every: int = 5, half_extent=(2.5, 1.5, 2.5),
                 sphere=(0.2, -0.5, -0.1, 0.8), orbit: float = 1.1) -> None:
        self.camera, self.device = camera, device
        self.num_frames, self.keyframe_every = num_frames, keyframe_every
        self.half = torch.tensor(half_extent, device=device)
        self.sphere_c = torch.tensor(sphere[:3], device=device)
        self.sphere_r = float(sphere[3])
        self.orbit = orbit
        h, w = camera.height, camera.width
        ij = torch.cartesian_prod(torch.arange(h, device=device), torch.arange(w, device=device))
        fx, fy, cx, cy, _ = camera.get_pinhole_camera_parameters(0.0)
        d = torch.stack(((ij[:, 1].float() - cx) / fx, -(ij[:, 0].float() - cy) / fy, -torch.ones(len(ij), device=device)), -1)
        self._dirs_cam = torch.nn.functional.normalize(d, dim=-1)  # ngm/camera.py:186-203, OpenGL

    def __len__(self) -> int:
        return self.num_frames

    def is_keyframe(self, frame_id: int) -> bool:
        return frame_id % self.keyframe_every == 0

    def c2w(self, frame_id: int) -> torch.Tensor:
        """OpenGL camera-to-world: x right, y up, camera looks down -z."""
        a = 2.0 * math.pi * frame_id / max(self.num_frames, 1) * 0.5
        eye = torch.tensor([self.orbit * math.sin(a), 0.15 * math.sin(3 * a), self.orbit * math.cos(a) + 0.9], device=self.device)
        fwd = torch.nn.functional.normalize(self.sphere_c - eye, dim=0)
        up = torch.tensor([0.0, 1.0, 0.0], device=self.device)
        right = torch.nn.functional.normalize(torch.linalg.cross(fwd, up), dim=0)
        up = torch.linalg.cross(right, fwd)
        m = torch.eye(4, device=self.device)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
        return m

    def _shade(self, p: torch.Tensor, on_sphere: torch.Tensor) -> torch.Tensor:
        room = 0.5 + 0.35 * torch.stack((torch.sin(1.3 * p[:, 0] + 0.5 * p[:, 1]), torch.sin(1.1 * p[:, 1] - 0.7 * p[:, 2] + 1.0),
                                         torch.sin(0.9 * p[:, 2] + 0.6 * p[:, 0] + 2.0)), -1)
        ball = 0.5 + 0.4 * torch.stack((torch.sin(4.0 * p[:, 1]), torch.cos(3.0 * p[:, 0] + 1.0), torch.sin(3.5 * p[:, 2] + 0.5)), -1)
        return torch.where(on_sphere[:, None], ball, room)

    @torch.no_grad()
    def frame(self, frame_id: int) -> Dict[str, torch.Tensor]:
        """{"rgbd": (H, W, 4), "c2w": (4, 4)} -- the reference's dataset item (ngm/slam_dataset.py:95-107)."""
        cam = self.camera
        c2w = self.c2w(frame_id)
        o = c2w[:3, 3]
        d = self._dirs_cam @ c2w[:3, :3].T
        # box interior: first exit
        t_hi = torch.where(d > 0, (self.half - o) / d, (-self.half - o) / d)
        t_hi = torch.where(d.abs() < 1e-9, torch.full_like(t_hi, 1e9), t_hi)
        t_box = t_hi.min(dim=-1)[0]
        # sphere
        oc = o - self.sphere_c
        b = (d * oc).sum(-1)
        disc = b * b - ((oc * oc).sum() - self.sphere_r ** 2)
        t_s = -b - torch.sqrt(disc.clamp_min(0.0))
        hit = (disc > 0) & (t_s > 0) & (t_s < t_box)
        t = torch.where(hit, t_s, t_box)
        p = o + d * t[:, None]
        rgb = self._shade(p, hit)
        depth = t * (-self._dirs_cam[:, 2])
        return {"rgbd": torch.cat((rgb, depth[:, None]), -1).view(cam.height, cam.width, 4).contiguous(), "c2w": c2w}


def psnr(prediction: torch.Tensor, target: torch.Tensor) -> float:
    """ngm/evaluation.py:46-56: clamp to [0, 1], 10 log10(1 / MSE)."""
    mse = ((prediction.clamp(0, 1) - target.clamp(0, 1)) ** 2).mean().item()
    return 10.0 * math.log10(1.0 / max(mse, 1e-12))


class MappingLoop(RenderState):
    """The mapping iteration of the reference driver (multi_view update mode) on a synthetic stream."""

    _reference_flow = True  # every _render_ijs first calls _set_vmap_fields, as the driver does (run_mapping.py:500)

    def __init__(self, config: dict, camera, stream: SyntheticStream) -> None:
        super().__init__(config)
        self._camera, self._dataset = camera, stream
        self._num_train_fields = config.get("num_train_fields", 32)
        self._num_rays_per_field = config.get("num_rays_per_field", 512)
        self._num_iterations_per_frame = config.get("num_iterations_per_frame", 5)
        self._termination_weight = config.get("termination_weight", 0.0)
        self._photometric_weight = config.get("photometric_weight", 1.0)
        self._photometric_loss = config.get("photometric_loss", "l1")
        self._depth_weight = config.get("depth_weight", 1.0)
        self._depth_loss = config.get("depth_loss", "huber")
        self._keyframes_only = False
        self._max_depth = config.get("max_depth", None)
        self._global_map_dict["kf_ids"] = torch.zeros(32, device=self._device, dtype=torch.long)  # run_mapping.py:231-246
        cap = config.get("max_keyframes", 64)
        h, w = camera.height, camera.width
        # run_mapping.py:1674-1684 (1000 slots there; slot 0 is the current frame)
        self._free_rgbd_tensor_indices = list(range(1, cap))
        self._nc_rgbd_tensor = torch.empty((cap, h, w, 4), device=self._device)
        self._nc_frame_id_tensor = torch.full((cap,), -1, device=self._device, dtype=torch.long)
        self._frame_c2ws: Dict[int, torch.Tensor] = {}
        self._current_frame_id = -1
        self._current_iteration = 0
        self._current_field_ids = torch.zeros(0, dtype=torch.long, device=self._device)
        self._total_optimization_time = 0.0
        self._fps_estimate = 0.0
        self._target = None

    @property
    def _num_fields(self) -> int:
        return self._global_map_dict["num"]

    # ---- field growth (run_mapping.py:248-345, 365-389) ----
    def _extend_map_dict(self, required_size: int) -> None:
        g = self._global_map_dict
        r = math.ceil(required_size / g["positions"].shape[0])
        g["positions"], g["orientations"] = g["positions"].repeat((r, 1)), g["orientations"].repeat((r, 1))
        g["kf_ids"], g["training_iterations"] = g["kf_ids"].repeat(r), g["training_iterations"].repeat((r,))

    def _add_fields(self, num_new: int) -> None:
        self._model.add_fields(num_new)
        self._optim_state = optim.new_optim_state(self._model.all_fields_params, self._optim_state)

    @torch.no_grad()
    def _extend_global_map_dict(self, depth_image: torch.Tensor, frame_id: int, c2w: torch.Tensor,
                                shift: Optional[torch.Tensor] = None) -> int:
        """Ensure fields cover the depth image, adding new ones on a randomly shifted grid of cell 2 r / sqrt(3)
        (run_mapping.py:267-345).  Points already inside an existing field are skipped (``ball_query`` with K = 1 there;
        a chunked nearest-centre test here).  Returns the number of fields added."""
        cam, dev, g = self._camera, self._device, self._global_map_dict
        ij = torch.nonzero(depth_image)  # camera.py:350-380 depth_to_pointcloud: pixels with depth
        fx, fy, cx, cy, _ = cam.get_pinhole_camera_parameters(0.0)
        z = depth_image[ij[:, 0], ij[:, 1]]
        xyz_cam = torch.stack(((ij[:, 1].float() - cx) / fx * z, -(ij[:, 0].float() - cy) / fy * z, -z), -1)
        xyz_world = xyz_cam @ c2w[:3, :3].T + c2w[:3, 3]
        num_prev = g["num"]
        if num_prev > 0:
            centres = g["positions"][:num_prev]
            keep = torch.ones(len(xyz_world), dtype=torch.bool, device=dev)
            for s0 in range(0, len(xyz_world), 1 << 16):
                d2 = torch.cdist(xyz_world[s0:s0 + (1 << 16)], centres).min(dim=-1)[0]
                keep[s0:s0 + (1 << 16)] = d2 >= self._field_radius
            xyz_world = xyz_world[keep]
        cell = 2 * self._field_radius / math.sqrt(3)
        if shift is None:
            shift = torch.empty((3,), device=dev).uniform_(0.0, cell)
        to_cover = ((xyz_world + shift) / cell).floor().unique(dim=0)
        covered = (((g["positions"][:num_prev] + shift) / cell).floor().unique(dim=0) if num_prev
                   else torch.empty(0, 3, device=dev))
        combined = torch.cat((to_cover, covered))
        _, inv, counts = combined.unique(dim=0, return_inverse=True, return_counts=True)
        new_ijk = to_cover[counts[inv[:len(to_cover)]] == 1]
        num_new = len(new_ijk)
        if num_new == 0:
            return 0
        if g["positions"].shape[0] <= num_prev + num_new:
            self._extend_map_dict(num_prev + num_new)
        start, end = num_prev, num_prev + num_new
        g["positions"][start:end] = (new_ijk - shift + 0.5) * cell
        g["orientations"][start:end] = 0.0
        g["orientations"][start:end, 0] = 1.0
        g["kf_ids"][start:end] = frame_id
        g["training_iterations"][start:end] = 0
        g["num"] = end
        self._add_fields(num_new)
        return num_new

    # ---- keyframe store (run_mapping.py:1686-1713) ----
    def _update_mv_training_data(self) -> None:
        self._nc_rgbd_tensor[0] = self._current_rgbd
        self._nc_frame_id_tensor[0] = self._current_frame_id
        if self._current_is_keyframe:
            if not self._free_rgbd_tensor_indices:
                raise ValueError("Maximum number of keyframes reached.")
            idx = self._free_rgbd_tensor_indices.pop(0)
            self._nc_rgbd_tensor[idx] = self._current_rgbd
            self._nc_frame_id_tensor[idx] = self._current_frame_id
        mask = self._nc_frame_id_tensor != -1
        self._frame_cid_to_ncid = torch.arange(len(mask), device=self._device)[mask]
        ids = self._nc_frame_id_tensor[mask].tolist()
        self._c_c2w_tensor = torch.stack([self._frame_c2ws[i] for i in ids])

    # ---- per-frame state (run_mapping.py:1599-1641) ----
    @torch.no_grad()
    def _update_slam_state(self, frame_id: int) -> None:
        item = self._dataset.frame(frame_id)
        self._current_frame_id = frame_id
        self._current_rgbd, self._current_c2w = item["rgbd"], item["c2w"]
        if self._max_depth is not None:
            self._current_rgbd[..., 3][self._current_rgbd[..., 3] > self._max_depth] = 0.0
        self._frame_c2ws[frame_id] = self._current_c2w
        self._current_is_keyframe = self._dataset.is_keyframe(frame_id)
        if self._current_is_keyframe:
            self._extend_global_map_dict(self._current_rgbd[:, :, 3], frame_id, self._current_c2w)
        self._current_field_ids = get_observed_fields(self, self._current_rgbd, self._current_c2w)
        self._update_mv_training_data()

    # ---- losses (run_mapping.py:1769-1871; ngm/losses.py:10-78) ----
    def _compute_losses(self, target, prediction) -> dict:
        """The reference's loss terms.  The reference compacts every masked operand first (`x[mask]`, a nonzero +
        gather + a sorting index_put in the backward: ~60 small launches per iteration) and then takes means; the
        same means are computed here as masked sums over the dense tensors divided by the mask count -- equal up to
        the order of the floating-point sums, NaN for an empty mask exactly like `mean()` of an empty tensor."""
        depth_mask = target.depth_mask * (prediction.term_probs > 0.8)  # :1787
        rgb_mask = depth_mask                                           # :1788
        out = {}

        def masked_mean(x, m, per_element=1):
            return (x * m).sum() / (m.sum() * per_element)

        tm = target.term_mask.float()
        term = masked_mean((prediction.term_probs - target.term_probs) ** 2, tm)  # :1803-1806
        combined = 0 + self._termination_weight * term
        out["termination"] = term
        m = rgb_mask.float()
        p_rgb, t_rgb = prediction.rgbds[..., :3], target.rgbds[..., :3]
        if self._photometric_loss == "l1":      # losses.py:22-23
            photo = masked_mean(torch.abs(p_rgb - t_rgb), m[..., None], 3)
        elif self._photometric_loss == "l2":    # losses.py:24-25
            photo = masked_mean((p_rgb - t_rgb) ** 2, m[..., None], 3)
        else:
            raise NotImplementedError(f"photometric_loss={self._photometric_loss}")
        combined = combined + self._photometric_weight * photo
        out[f"photometric_{self._photometric_loss}"] = photo
        t_d, p_d = target.rgbds[..., 3], prediction.rgbds[..., 3]
        if self._depth_loss == "huber":         # losses.py:60-61: F.huber_loss(rendered, measured, delta=0.05)
            dl = masked_mean(torch.nn.functional.huber_loss(p_d, t_d, delta=0.05, reduction="none"), m)
        elif self._depth_loss == "gaussian_nll":  # losses.py:62-67
            v = prediction.depth_vars + 1e-15
            dl = masked_mean(0.5 * (p_d - t_d) ** 2 / v + torch.log(torch.sqrt(v)), m)
        else:
            raise NotImplementedError(f"depth_loss={self._depth_loss}")
        combined = combined + self._depth_weight * dl
        out[f"depth_{self._depth_loss}"] = dl
        if prediction.freespace_geometry is not None:  # :1842-1846
            fs = ((prediction.freespace_geometry - self._truncation_distance) ** 2).mean()
            combined = combined + self._freespace_weight * fs
            out["freespace"] = fs
        if prediction.tsdf_residuals is not None:      # :1848-1851
            ts = (prediction.tsdf_residuals ** 2).mean()
            combined = combined + self._tsdf_weight * ts
            out["tsdf"] = ts
        out["combined"] = combined
        return out

    # ---- one iteration (run_mapping.py:1124-1182) ----
    def _optimization_iteration(self, draws: Optional[dict] = None, jitter=None, jitter_guided=None) -> dict:
        target = sample_target_mv(self, self._current_field_ids, draws)
        self._target = target
        if len(target.field_ids) == 0:
            return {}
        prediction = self._render_ijs(target.ijs, target.c2ws, self._camera, near_distances=target.near_distances,
                                      far_distances=target.far_distances, gt_distances=target.gt_distances,
                                      field_ids=target.field_ids, use_vmap=True, jitter=jitter, jitter_guided=jitter_guided)
        loss_dict = self._compute_losses(target, prediction)
        self._update_step(loss_dict, target.field_ids)
        self._current_iteration += 1
        return loss_dict

    # ---- one input frame (run_mapping.py:1224-1251) ----
    def _current_frame_optimization(self, frame_id: int) -> dict:
        torch.cuda.synchronize()
        start = time.time()
        self._update_slam_state(frame_id)
        loss_dict = {}
        for _ in range(self._num_iterations_per_frame):
            loss_dict = self._optimization_iteration()
        torch.cuda.synchronize()
        self._total_optimization_time += time.time() - start
        self._fps_estimate = (frame_id + 1) / self._total_optimization_time
        return loss_dict

    def fit(self, num_frames: Optional[int] = None, log_every: int = 0) -> dict:
        """Run the loop over the stream (ngm/run_mapping.py:1092-1095).  Returns fps and the last losses."""
        n = len(self._dataset) if num_frames is None else num_frames
        self.train()
        last = {}
        for frame_id in range(n):
            last = self._current_frame_optimization(frame_id)
            if log_every and frame_id % log_every == 0 and last:
                print(f"[mapping] frame {frame_id}: fields {self._num_fields}, combined loss {float(last['combined']):.5f}, "
                      f"fps {self._fps_estimate:.2f}", flush=True)
        return {"frames": n, "fps": self._fps_estimate, "seconds": self._total_optimization_time,
                "fields": self._num_fields, "iterations": self._current_iteration,
                "losses": {k: float(v) for k, v in last.items()}}

    @torch.no_grad()
    def evaluate_frame(self, frame_id: int) -> dict:
        """Render the frame through ``render_image`` (kNN path over all fields) and compare with the stream
        (ngm/run_mapping.py:1977-2010: PSNR per ngm/evaluation.py:46-56, depth L1 per :59-62)."""
        item = self._dataset.frame(frame_id)
        self.eval()
        rgbd, _ = self.render_image(item["c2w"], self._camera)
        self.train()
        gt = item["rgbd"]
        valid = gt[..., 3] > 0
        return {"psnr": psnr(rgbd[..., :3], gt[..., :3]),
                "depth_l1": (rgbd[..., 3][valid] - gt[..., 3][valid]).abs().mean().item()}
