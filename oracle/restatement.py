"""Plain-PyTorch (CPU, fp32) restatement of the reference ray-render hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA path
and the CPU baseline of ``bench.py``.  Never imported by ``neural_graph_mapping_b200``.

Every function cites the reference lines it follows; ``ngm/`` abbreviates
``/root/reference/src/neural_graph_mapping/``.  The restatement is *pinned*: the
``not gpu`` tests compare it with the committed golden vectors in ``tests/golden``
that ``oracle/make_golden.py`` produced by running the unmodified reference on CPU
(and, where ``/root/reference`` exists, with the live reference as well).

Exception: the permutohedral hash encoding lives in ``oracle/permuto.py`` and is
PARITY UNPINNED (third-party source absent from the reference tree).
"""
from __future__ import annotations

import math
from collections import namedtuple
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import permuto as _permuto

# ngm/run_mapping.py:59-69
Prediction = namedtuple(
    "Prediction",
    ["rgbds", "color_vars", "depth_vars", "term_probs", "freespace_geometry", "tsdf_residuals"],
)


# --------------------------------------------------------------------------------------
# specs (plain data; mirror the reference's YAML keys, ngm/config/neural_graph_map.yaml)
# --------------------------------------------------------------------------------------
@dataclass
class CameraSpec:
    """ngm/camera.py:22-79 ctor arguments."""

    width: int = 640
    height: int = 480
    fx: float = 554.2562584220408
    fy: float = 554.2562584220408
    cx: float = 319.5
    cy: float = 239.5
    pixel_center: float = 0.0

    def principal_point_pc0(self):
        """cx/cy as used by ``ijs_to_directions``: stored +0.5 (camera.py:69-70), read back
        with pixel_center 0 -> -0.5 (camera.py:114-115,188)."""
        cx_s = self.cx - self.pixel_center + 0.5
        cy_s = self.cy - self.pixel_center + 0.5
        return cx_s - 0.5 + 0.0, cy_s - 0.5 + 0.0


@dataclass
class FieldSpec:
    """ngm/models.py:69-79 ctor arguments (encoding given by kind + kwargs)."""

    encoding: str = "nerf"  # nerf | fourier | triplane | permuto
    encoding_kwargs: dict = field(default_factory=lambda: {"dim_in": 3, "num_octaves": 4})
    num_layers: int = 2
    dim_out: int = 4
    dim_mlp_out: Optional[int] = 32
    skip_mode: str = "no"

    def dim_encoding(self) -> int:
        k = self.encoding_kwargs
        if self.encoding == "nerf":  # positional_encodings.py:274-276
            return k.get("dim_in", 3) * k.get("num_octaves", 8) * 2
        if self.encoding == "fourier":  # :214-216
            return k["dim_out"]
        if self.encoding == "triplane":  # :123-130
            n = k.get("num_components", 64)
            return 3 * n if k.get("mode", "sum") == "concat" else n
        if self.encoding == "permuto":  # :64-66 -> output_dims()
            return k["nr_levels"] * k["nr_feat_per_level"] + (
                k.get("pos_dim", 3) if k.get("concat_points", False) else 0
            )
        raise ValueError(self.encoding)

    def dims(self):
        """ngm/models.py:96-126."""
        e = self.dim_encoding()
        w = self.dim_mlp_out if self.dim_mlp_out is not None else e
        w_in = w + e if self.skip_mode == "concat" else w
        dims_in = [e] + [w_in] * self.num_layers
        dims_out = [w] * self.num_layers + [self.dim_out]
        return dims_in, dims_out


@dataclass
class RenderSpec:
    """Driver state read by ``_render_ijs`` (ngm/run_mapping.py:116-215)."""

    num_samples: int = 32
    num_samples_depth_guided: int = 0
    range_depth_guided: float = 0.1
    near_distance: float = 0.0
    far_distance: float = 8.0
    truncation_distance: float = 0.1
    freespace_weight: float = 40.0
    tsdf_weight: float = 50.0
    geometry_mode: str = "nrgbd"
    geometry_factor: float = 20.0
    color_factor: float = 1.0
    field_radius: float = 1.0
    scale_mode: str = "unit_cube"
    num_knn: int = 2
    distance_factor: float = 10.0
    outside_value: float = 1.0
    block_size: int = 3_000_000
    pixel_block_size: int = 8192


# --------------------------------------------------------------------------------------
# a16: quaternion helpers (pytorch3d @47d5dc88 semantics; used ngm/models.py:240-241)
# --------------------------------------------------------------------------------------
def quaternion_invert(q: torch.Tensor) -> torch.Tensor:
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)


def _quat_mul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        (
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ),
        -1,
    )


def quaternion_apply(q: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    pq = torch.cat((p.new_zeros(p.shape[:-1] + (1,)), p), -1)
    return _quat_mul(_quat_mul(q, pq), quaternion_invert(q))[..., 1:]


# --------------------------------------------------------------------------------------
# a1/a2: ray sampler (ngm/camera.py:186-203, 215-292)
# --------------------------------------------------------------------------------------
def ijs_to_directions(ijs: torch.Tensor, cam: CameraSpec) -> torch.Tensor:
    cx0, cy0 = cam.principal_point_pc0()
    d_x = (ijs[..., 1] - cx0) / cam.fx  # camera.py:189
    d_y = -((ijs[..., 0] - cy0) / cam.fy)  # :190,194
    d_z = -torch.ones_like(d_x)  # :195
    dirs = torch.stack([d_x, d_y, d_z], dim=-1)
    return torch.nn.functional.normalize(dirs, dim=-1)  # :202


def sample_ijs_uniform(ijs, cam: CameraSpec, num_samples: int, near, far, jitter: torch.Tensor):
    """camera.py:261-292, uniform branch; ``jitter`` replaces ``torch.rand(*leading, S)`` (:274)."""
    leading = ijs.shape[:-1]
    dirs = ijs_to_directions(ijs, cam)
    if not torch.is_tensor(far):
        far = torch.tensor(float(far)).expand(leading)
    if not torch.is_tensor(near):
        near = torch.tensor(float(near)).expand(leading)
    deltas = (far - near) / num_samples
    boundaries = torch.linspace(0.0, 1.0, steps=num_samples + 1)
    boundaries = boundaries[None] * (far - near)[..., None]
    distances = (deltas[..., None] * jitter + boundaries[..., :-1]) + near[..., None]
    points = dirs.unsqueeze(-2) * distances.unsqueeze(-1)
    return points, distances


def transform_points(points: torch.Tensor, transforms: torch.Tensor) -> torch.Tensor:
    """ngm/utils.py:276-286 (forward direction)."""
    return (
        torch.einsum("...dk,...k -> ...d", transforms[..., :3, :3], points)
        + transforms[..., :3, 3]
    )


# --------------------------------------------------------------------------------------
# a8-a10: encodings (ngm/positional_encodings.py)
# --------------------------------------------------------------------------------------
def encode_nerf(points, num_octaves=8, start_octave=0, dim_in=3):
    """positional_encodings.py:245-272."""
    leading = points.shape[:-1]
    octaves = torch.arange(start_octave, start_octave + num_octaves, dtype=torch.float)
    multipliers = 2**octaves * torch.pi
    scaled = points.unsqueeze(-1) * multipliers
    sines = torch.sin(scaled).reshape(*leading, -1)
    cosines = torch.cos(scaled).reshape(*leading, -1)
    return torch.cat((sines, cosines), -1)


def encode_fourier(points, weight, raw_coords: bool):
    """positional_encodings.py:197-212; ``weight`` = ``_encoding._linear.weight`` (out,in)."""
    ff = torch.sin(points @ weight.transpose(-1, -2))
    return torch.cat((points, ff), dim=-1) if raw_coords else ff


def encode_triplane(points, plane_coef, mode="sum"):
    """positional_encodings.py:132-161; ``plane_coef`` (3, C, res, res)."""
    original_shape = points.shape
    x = points.reshape(-1, 3)
    plane_coord = torch.stack([x[..., [0, 1]], x[..., [0, 2]], x[..., [1, 2]]], dim=0)
    plane_coord = plane_coord.view(3, -1, 1, 2)
    feats = torch.nn.functional.grid_sample(
        plane_coef, plane_coord, align_corners=True, padding_mode="border"
    )
    c = plane_coef.shape[1]
    if mode == "product":
        feats = feats.prod(0).squeeze(-1).T
    elif mode == "sum":
        feats = feats.sum(0).squeeze(-1).T
    elif mode == "concat":
        feats = feats.squeeze(-1).reshape(3 * c, -1).T
    else:
        raise ValueError(mode)
    return feats.reshape(*original_shape[:-1], -1)


def encode(points, spec: FieldSpec, params: Dict[str, torch.Tensor]):
    k = spec.encoding_kwargs
    if spec.encoding == "nerf":
        return encode_nerf(points, k.get("num_octaves", 8), k.get("start_octave", 0))
    if spec.encoding == "fourier":
        return encode_fourier(points, params["_encoding._linear.weight"], k["raw_coords"])
    if spec.encoding == "triplane":
        return encode_triplane(points, params["_encoding.plane_coef"], k.get("mode", "sum"))
    if spec.encoding == "permuto":
        return _permuto.encode(
            points,
            params["_encoding.lattice_values"],
            params["_encoding.random_shift_per_level"],
            _permuto.scale_factors(k),
            concat_points=k.get("concat_points", False),
            concat_points_scaling=k.get("concat_points_scaling", 1.0),
        )
    raise ValueError(spec.encoding)


# --------------------------------------------------------------------------------------
# a11: NeuralField.forward (ngm/models.py:143-182)
# --------------------------------------------------------------------------------------
def field_forward(points, spec: FieldSpec, params: Dict[str, torch.Tensor]):
    """``params`` uses the reference's state_dict names for ONE field
    (``_linears.{i}.weight`` (out,in), ``_linears.{i}.bias``, ``_rezero``, ``_encoding.*``)."""
    enc = outs = encode(points, spec, params)
    e = enc.shape[-1]
    for i in range(spec.num_layers + 1):
        prev = outs
        w, b = params[f"_linears.{i}.weight"], params[f"_linears.{i}.bias"]
        outs = torch.nn.functional.linear(outs, w, b)  # models.py:151
        if i == spec.num_layers:
            break  # :153-154
        outs = torch.relu(outs)  # :157
        if spec.skip_mode == "concat":  # :160-161
            outs = torch.cat((outs, enc), dim=-1)
        elif spec.skip_mode == "add":  # :162-169
            outs = torch.cat((outs[..., :e] + enc, outs[..., e:]), dim=-1)
        elif spec.skip_mode == "rezero":  # :170-180
            a = params["_rezero"][i]
            if i == 0:
                outs = torch.cat((a * outs[..., :e] + prev, a * outs[..., e:]), dim=-1)
            else:
                outs = a * outs + prev
    return outs


def init_field_params(spec: FieldSpec, generator: torch.Generator, geometry_bias: float = 0.0):
    """Seeded ``nn.Linear``-style init (uniform +-1/sqrt(fan_in)) for one field.
    Not bit-identical to ``torch.nn.Linear.reset_parameters`` -- parity runs copy the
    reference's own state_dict instead; this is for synthetic benchmarks/tests."""
    dims_in, dims_out = spec.dims()
    p = {}
    for i, (di, do) in enumerate(zip(dims_in, dims_out)):
        bound = 1.0 / math.sqrt(di)
        p[f"_linears.{i}.weight"] = (torch.rand(do, di, generator=generator) * 2 - 1) * bound
        p[f"_linears.{i}.bias"] = (torch.rand(do, generator=generator) * 2 - 1) * bound
    p[f"_linears.{spec.num_layers}.bias"][-1] += geometry_bias  # models.py:135
    k = spec.encoding_kwargs
    if spec.encoding == "fourier":
        n = k["dim_out"] - k["dim_in"] if k["raw_coords"] else k["dim_out"]
        p["_encoding._linear.weight"] = k.get("mu", 0.0) + k.get("sigma", 1.0) * torch.randn(
            n, k["dim_in"], generator=generator
        )
    elif spec.encoding == "triplane":
        r, c = k.get("resolution", 32), k.get("num_components", 64)
        p["_encoding.plane_coef"] = k.get("init_scale", 0.1) * torch.randn(
            3, c, r, r, generator=generator
        )
    elif spec.encoding == "permuto":
        p.update(_permuto.init_params(k, generator))
    if spec.skip_mode == "rezero":
        p["_rezero"] = torch.zeros(spec.num_layers)
    return p


def stack_params(per_field: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """``all_fields_params`` layout: name -> (F, *shape) (ngm/models.py:254-264)."""
    return {k: torch.stack([p[k] for p in per_field]) for k in per_field[0]}


# --------------------------------------------------------------------------------------
# a6/a7: NeuralFieldSet.forward (ngm/models.py:278-405)
# --------------------------------------------------------------------------------------
def scale_local_points(x, scale_mode: str, field_radius: float):
    if scale_mode == "unit_cube":
        return x / (2 * field_radius) + 0.5  # models.py:280
    if scale_mode == "unit_ball":
        return x / field_radius  # :282
    if scale_mode == "no":
        return x
    raise NotImplementedError(scale_mode)


def fieldset_forward_vmap(query, positions, orientations, fspec, stacked_params, rspec):
    """vmap branch, models.py:329-345.  query (F,N,3); positions (F,3)|None; orientations (F,4)."""
    if positions is not None:
        local = query - positions.unsqueeze(-2)
        local = quaternion_apply(quaternion_invert(orientations).unsqueeze(-2), local)
    else:
        local = query
    local = scale_local_points(local, rspec.scale_mode, rspec.field_radius)
    outs = []
    for f in range(local.shape[0]):
        outs.append(field_forward(local[f], fspec, {k: v[f] for k, v in stacked_params.items()}))
    return torch.stack(outs)


def fieldset_forward_knn(query, positions, orientations, field_ids, fspec, all_params, rspec,
                         field_radius: Optional[float] = None):
    """kNN branch, models.py:347-405.  query (...,3); positions (F,3); orientations (F,4)."""
    if field_radius is None:
        field_radius = rspec.field_radius
    if field_ids is None:
        field_ids = torch.arange(len(positions))
    leading = query.shape[:-1]
    q = query.reshape(-1, 3)
    n = len(q)
    k = min(rspec.num_knn, len(positions))
    d2 = ((q[:, None, :] - positions[None, :, :]) ** 2).sum(-1)  # pytorch3d knn_points: squared L2
    knn_d2, knn_idx = torch.topk(d2, k, dim=-1, largest=False, sorted=True)
    knn_d = torch.sqrt(knn_d2)  # :367
    radius_mask = knn_d[:, 0] < field_radius  # :369
    knn_d, knn_idx, qm = knn_d[radius_mask], knn_idx[radius_mask], q[radius_mask]
    local = qm.unsqueeze(-2) - positions[knn_idx]  # :377
    local = quaternion_apply(quaternion_invert(orientations[knn_idx]), local)
    local = scale_local_points(local, rspec.scale_mode, rspec.field_radius)  # uses the class radius (:381)
    w = torch.softmax(-rspec.distance_factor * knn_d, dim=-1)  # :384
    knn_outs = torch.empty(*local.shape[:-1], 4)
    for fi in knn_idx.unique().tolist():  # :386-396
        mask = knn_idx == fi
        params = {kk: v[field_ids[fi]] for kk, v in all_params.items()}
        knn_outs[mask] = field_forward(local[mask], fspec, params)
    in_radius = (w.unsqueeze(-1) * knn_outs).sum(-2)  # :399
    outs = torch.full((n, 4), rspec.outside_value)  # :401
    outs[radius_mask] = in_radius
    return outs.reshape(*leading, -1)


# --------------------------------------------------------------------------------------
# a13: quadrature (ngm/run_mapping.py:709-799)
# --------------------------------------------------------------------------------------
def quadrature(sample_colors, sample_geometries, sample_distances, sample_depths, neus_isds,
               geometry_mode: str, geometry_factor: float):
    leading = sample_geometries.shape[:-1]
    if geometry_mode == "density":  # :746-749
        deltas = sample_distances[..., 1:] - sample_distances[..., :-1]
        occ = 1 - torch.exp(-deltas * torch.relu(sample_geometries[..., :-1]))
        last = -1
    elif geometry_mode == "occupancy":  # :750-752
        occ = torch.sigmoid(geometry_factor * sample_geometries)
        last = None
    elif geometry_mode == "neus":  # :753-758
        tno = torch.sigmoid(neus_isds * geometry_factor * sample_geometries)
        occ = torch.clamp_min((tno[..., :-1] - tno[..., 1:]) / (tno[..., :-1] + 1e-5), 0)
        last = -1
    elif geometry_mode == "nrgbd":  # :759-762
        temp = geometry_factor * sample_geometries
        occ = 4 * torch.sigmoid(temp) * torch.sigmoid(-temp)
        last = None
    else:
        raise ValueError(geometry_mode)
    non_term = torch.cat(
        [occ.new_ones(*leading, 1), torch.cumprod(1 - occ[..., :-1], dim=-1)], dim=-1
    )  # :764-770
    weights = occ * non_term  # :771
    bg = 1 - torch.sum(weights, dim=-1)  # :774
    colors = torch.sum(sample_colors[..., :last, :] * weights[..., None], dim=-2)  # :776-778
    depths = torch.sum(sample_depths[..., :last] * weights, dim=-1)  # :779
    color_vars = torch.sum(
        weights[..., None] * (colors.unsqueeze(-2) - sample_colors[..., :last, :]) ** 2, dim=-2
    )  # :781-785
    depth_vars = torch.sum(weights * (depths[..., None] - sample_depths[..., :last]) ** 2, dim=-1)
    return colors, depths, color_vars, depth_vars, 1.0 - bg, weights


# --------------------------------------------------------------------------------------
# a3/a12/a14: _render_ijs (ngm/run_mapping.py:440-666)
# --------------------------------------------------------------------------------------
def render_rays(ijs, c2ws, cam: CameraSpec, rspec: RenderSpec, fspec: FieldSpec,
                all_params, positions, orientations, field_ids=None, use_vmap=False,
                near_distances=None, far_distances=None, gt_distances=None,
                overwrite_samples_behind_camera=True, jitter=None, jitter_guided=None,
                neus_sd=None):
    """``positions``/``orientations`` are the *global* (num_fields, .) tables
    (``_global_map_dict``); ``all_params`` the stacked ``all_fields_params``."""
    if near_distances is None or bool((near_distances >= 0).all()):  # :494-495
        overwrite_samples_behind_camera = False
    if use_vmap and field_ids is None:
        raise ValueError("field_ids=None only supported for use_vmap=False")  # :497-498
    if field_ids is not None:  # :502-508
        f_pos, f_ori = positions[field_ids], orientations[field_ids]
    else:
        f_pos, f_ori = positions, orientations
    if c2ws.dim() == 2:
        c2ws = c2ws[None]
    near = rspec.near_distance if near_distances is None else near_distances
    far = rspec.far_distance if far_distances is None else far_distances
    points_cam, dists = sample_ijs_uniform(ijs, cam, rspec.num_samples, near, far, jitter)

    if gt_distances is not None and rspec.num_samples_depth_guided > 0:  # :521-545
        mask = (gt_distances == 0.0) + (near_distances > gt_distances) + (far_distances < gt_distances)
        g_near = gt_distances - rspec.range_depth_guided
        g_far = gt_distances + rspec.range_depth_guided
        g_near[mask] = near_distances[mask]
        g_far[mask] = far_distances[mask]
        g_points, g_dists = sample_ijs_uniform(
            ijs, cam, rspec.num_samples_depth_guided, g_near, g_far, jitter_guided
        )
        points_cam = torch.cat([points_cam, g_points], dim=-2)
        dists = torch.cat([dists, g_dists], dim=-1)
        dists, sort_idx = torch.sort(dists, dim=-1)
        points_cam = torch.gather(points_cam, -2, sort_idx.unsqueeze(-1).expand(*sort_idx.shape, 3))

    points_world = transform_points(points_cam, c2ws.unsqueeze(-3))  # :547

    if use_vmap:  # :577-585
        vmap_params = {k: v[field_ids] for k, v in all_params.items()}  # models.py:274-276
        outs = fieldset_forward_vmap(
            points_world.reshape(len(field_ids), -1, 3), f_pos, f_ori, fspec, vmap_params, rspec
        ).view(len(field_ids), points_world.shape[1], points_world.shape[2], -1)
    else:  # :586-595 (batched_evaluation is a pure chunking of the same math)
        outs = fieldset_forward_knn(points_world.reshape(-1, 3), f_pos, f_ori, field_ids, fspec,
                                    all_params, rspec).view(*points_world.shape[:-1], 4)

    colors = rspec.color_factor * outs[..., :3]  # :610
    geom = outs[..., 3].clone()  # :611
    depths = -points_cam[..., 2]  # :612

    if overwrite_samples_behind_camera:  # :614-622
        fill = -100.0 if rspec.geometry_mode in ("occupancy", "density") else 1.0
        geom[points_cam[..., 2] > 0] = fill

    tau = rspec.truncation_distance
    if rspec.freespace_weight != 0.0 and gt_distances is not None:  # :624-630
        m = dists < (gt_distances[..., None] - tau) * (gt_distances[..., None] != 0.0)
        freespace = geom[m] * tau
    else:
        freespace = None
    if rspec.tsdf_weight != 0.0 and gt_distances is not None:  # :632-639
        deltas = gt_distances[..., None] - dists
        m = (torch.abs(deltas) < tau) * (gt_distances[..., None] != 0.0)
        tsdf = geom[m] * tau - deltas[m]
    else:
        tsdf = None

    neus_isds = None
    if rspec.geometry_mode == "neus" and use_vmap:  # :641-644
        neus_isds = 1.0 / torch.abs(all_params["_neus_sd"][field_ids].view(-1, 1, 1))

    c, d, cv, dv, tp, _ = quadrature(colors, geom, dists, depths, neus_isds,
                                     rspec.geometry_mode, rspec.geometry_factor)
    return Prediction(torch.cat([c, d[..., None]], dim=-1), cv, dv, tp, freespace, tsdf)


def render_image(c2w, cam: CameraSpec, rspec: RenderSpec, fspec, all_params, positions,
                 orientations, jitter_full: torch.Tensor):
    """ngm/run_mapping.py:402-437 + utils.py:220-251; ``jitter_full`` (H*W, S)."""
    h, w = cam.height, cam.width
    ijs = torch.cartesian_prod(torch.arange(h), torch.arange(w))
    rgbds, dvars = [], []
    for s in range(0, ijs.shape[0], rspec.pixel_block_size):
        e = min(s + rspec.pixel_block_size, ijs.shape[0])
        p = render_rays(ijs[s:e], c2w, cam, rspec, fspec, all_params, positions, orientations,
                        jitter=jitter_full[s:e])
        rgbds.append(p.rgbds)
        dvars.append(p.depth_vars)
    return torch.cat(rgbds).reshape(h, w, 4), torch.cat(dvars).reshape(h, w)


# --------------------------------------------------------------------------------------
# parity metrics (ngm/evaluation.py:46-62)
# --------------------------------------------------------------------------------------
def psnr(est: torch.Tensor, gt: torch.Tensor, crop: int = 0) -> float:
    est, gt = est.clamp(0, 1), gt.clamp(0, 1)
    if crop:
        est, gt = est[crop:-crop, crop:-crop], gt[crop:-crop, crop:-crop]
    mse = torch.mean((est - gt) ** 2).item()
    return 10.0 * math.log10(1.0 / max(mse, 1e-20))


def depth_l1(est: torch.Tensor, gt: torch.Tensor) -> float:
    mask = gt != 0
    return torch.mean(torch.abs(est[mask] - gt[mask])).item()
