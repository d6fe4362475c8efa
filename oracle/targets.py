"""Plain-PyTorch (CPU) restatement of the reference's multi-view target sampling.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``ngm/`` abbreviates
``/root/reference/src/neural_graph_mapping/``.  Pinned by ``tests/golden/target_mv.npz``, which
``oracle/make_target_fixture.py`` produced by running the unmodified ``NeuralGraphMap._sample_target_mv`` on CPU
while recording its random draws.
"""
from __future__ import annotations

from collections import namedtuple

import torch

from . import restatement as R

# ngm/run_mapping.py:43-58
Target = namedtuple("Target", ["ijs", "c2ws", "near_distances", "far_distances", "gt_distances", "field_ids", "rgbds",
                               "rgb_mask", "depth_mask", "term_probs", "term_mask"])


def transform_points_inv(points, transforms):
    """utils.transform_points(..., inv=True) (ngm/utils.py:279-282): R^T (p - t)."""
    return torch.einsum("...kd,...k -> ...d", transforms[..., :3, :3], points - transforms[..., :3, 3])


def project_points_opengl(points, cam: R.CameraSpec, pixel_center: float = 0.5):
    """Camera.project_points(points, "opengl") (ngm/camera.py:119-154, 173-177)."""
    cx = cam.cx - cam.pixel_center + pixel_center  # get_pinhole_camera_parameters(pixel_center), :98-116
    cy = cam.cy - cam.pixel_center + pixel_center
    m = torch.tensor([[cam.fx, 0, -cx], [0, -cam.fy, -cy], [0, 0, -1]], dtype=points.dtype, device=points.device)
    h = torch.einsum("oi,...i->...o", m, points)
    return h[..., :2] / h[..., 2].unsqueeze(-1)


def directions(ijs, cam: R.CameraSpec, convention: str, dtype=torch.float32):
    """Camera.ijs_to_directions (ngm/camera.py:186-203)."""
    cx0, cy0 = cam.cx - cam.pixel_center, cam.cy - cam.pixel_center
    dx = (ijs[..., 1].to(dtype) - cx0) / cam.fx
    dy = (ijs[..., 0].to(dtype) - cy0) / cam.fy
    dz = torch.ones_like(dx)
    if convention == "opengl":
        dy, dz = -dy, -dz
    return torch.nn.functional.normalize(torch.stack([dx, dy, dz], -1), dim=-1)


def choose_fields(current_field_ids, num_train_fields, num_fields, subset_observed, subset_random):
    """ngm/run_mapping.py:1295-1317 with the two multinomial draws supplied."""
    observed = current_field_ids[subset_observed]
    num_random = min(num_train_fields - len(observed), num_fields - len(observed))
    if num_random > 0:
        return torch.unique(torch.cat((torch.arange(num_fields, device=observed.device)[subset_random], observed)))
    return observed


def visibility(cam: R.CameraSpec, c2ws, rgbds, frame_to_store, positions, field_ids, probe_offsets, train_radius):
    """ngm/run_mapping.py:1319-1362 and 1386-1389: (field_kf_mask (F,K), min_xys (F,K,2), max_xys (F,K,2),
    probes' image coordinates (F,S,K,2))."""
    F, S, K = len(field_ids), probe_offsets.shape[0], c2ws.shape[0]
    pos_w = positions[field_ids]
    samples_w = pos_w.unsqueeze(1) + probe_offsets * train_radius * 1.0
    samples_c = transform_points_inv(samples_w.unsqueeze(-2), c2ws)  # (F, S, K, 3)
    depths = -samples_c[..., 2]
    xy = project_points_opengl(samples_c, cam)
    frame_cids = torch.arange(K, device=c2ws.device).expand(F, S, -1).reshape(-1)
    ij = xy.int().view(-1, 2)
    valid = (ij[:, 0] >= 0) * (ij[:, 0] < cam.width) * (ij[:, 1] >= 0) * (ij[:, 1] < cam.height)
    kf_depths = torch.zeros_like(depths)
    kf_depths[valid.view(F, S, K)] = rgbds[frame_to_store[frame_cids[valid]], ij[valid, 1], ij[valid, 0], 3].to(depths.dtype)
    in_front = (depths > 0).any(dim=-2)
    closer = (depths < kf_depths).any(dim=-2)
    in_frustum = valid.view(F, S, K).any(dim=-2)
    mask = in_front * closer * in_frustum
    min_xys = xy.min(dim=1)[0].clamp_min(0.0)
    max_xys = torch.minimum(xy.max(dim=1)[0], torch.tensor((cam.width, cam.height), dtype=xy.dtype, device=xy.device))
    return mask, min_xys, max_xys, xy


def rays(cam: R.CameraSpec, c2ws, rgbds, frame_to_store, positions, field_ids, frame_cids, uv, min_xys, max_xys,
         train_radius) -> Target:
    """ngm/run_mapping.py:1391-1459 for already-filtered fields (rows of min/max_xys = field_ids)."""
    lo = torch.gather(min_xys, 1, frame_cids[..., None].expand(-1, -1, 2))
    hi = torch.gather(max_xys, 1, frame_cids[..., None].expand(-1, -1, 2))
    xys = (hi - lo) * uv + lo
    jis = torch.minimum(xys.int(), torch.tensor((cam.width - 1, cam.height - 1), dtype=torch.int32, device=xys.device))
    ijs = torch.stack((jis[..., 1], jis[..., 0]), dim=-1)
    t_c2ws = c2ws[frame_cids]
    pos_c = transform_points_inv(positions[field_ids].unsqueeze(1), t_c2ws)
    dirs = directions(ijs, cam, "opengl", pos_c.dtype)
    center = (pos_c * dirs).sum(-1)
    near = (center - train_radius).clamp_min(0.0)
    far = (center + train_radius).clamp_min(0.0)
    t_rgbds = rgbds[frame_to_store[frame_cids], ijs[..., 0].long(), ijs[..., 1].long()]
    gt = t_rgbds[..., 3] / directions(ijs, cam, "opencv", t_rgbds.dtype)[..., 2]  # camera.py:339-340
    valid_depth = gt != 0.0
    return Target(ijs=ijs, c2ws=t_c2ws, near_distances=near, far_distances=far, gt_distances=gt, field_ids=field_ids,
                  rgbds=t_rgbds, rgb_mask=(t_rgbds[..., :2] != 0.0).any(dim=-1),
                  depth_mask=(gt > near) * (gt < far) * valid_depth, term_probs=(gt < far).float(),
                  term_mask=(gt > near) * valid_depth)


def sample_target_mv(cam: R.CameraSpec, c2ws, rgbds, frame_to_store, positions, num_fields, current_field_ids,
                     num_train_fields, train_radius, draws) -> Target:
    """``NeuralGraphMap._sample_target_mv`` (ngm/run_mapping.py:1261-1459) with its five random draws supplied."""
    ids = choose_fields(current_field_ids, num_train_fields, num_fields, draws["subset_observed"], draws.get("subset_random"))
    off = draws["probe_offsets"] / torch.linalg.norm(draws["probe_offsets"], dim=-1, keepdim=True)
    mask, lo, hi, _ = visibility(cam, c2ws, rgbds, frame_to_store, positions, ids, off, train_radius)
    fm = mask.any(dim=-1)
    return rays(cam, c2ws, rgbds, frame_to_store, positions, ids[fm], draws["frame_cids"], draws["uv"], lo[fm], hi[fm],
                train_radius)


def observed_fields(cam: R.CameraSpec, depth, c2w, positions, field_radius, subset):
    """``NeuralGraphMap._get_observed_fields`` (ngm/run_mapping.py:1643-1670) with its multinomial draw supplied:
    ids of the fields (rows of ``positions``) the frame observes."""
    pos_c = transform_points_inv(positions, c2w)
    ids = torch.arange(positions.shape[0], device=positions.device)
    cx0, cy0 = cam.cx - cam.pixel_center, cam.cy - cam.pixel_center
    ijs = torch.nonzero(depth)  # Camera.depth_to_pointcloud, camera.py:372-385 (OpenGL)
    d = depth[ijs[:, 0], ijs[:, 1]]
    pts = torch.stack(((ijs[:, 1].to(d.dtype) - cx0) * d / cam.fx, -(ijs[:, 0].to(d.dtype) - cy0) * d / cam.fy, -d), -1)
    pts = pts[subset]
    lo, hi = pts.min(dim=0)[0], pts.max(dim=0)[0]
    in_box = ((pos_c - field_radius <= hi).all(-1)) * ((pos_c + field_radius >= lo).all(-1))  # geometry.py:25-42
    c, ids = pos_c[in_box], ids[in_box]
    sq = (pts * pts).sum(-1, keepdim=True)  # segments from the origin: geometry.py:86-103
    sq[sq == 0] = 1.0
    t = ((c[:, None] * pts).sum(-1, keepdim=True) / sq).clamp(0.0, 1.0)
    d2 = ((c[:, None] - pts * t) ** 2).sum(-1)
    return ids[(d2 <= field_radius ** 2).any(-1)]
