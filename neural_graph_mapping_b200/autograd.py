"""Differentiable render path: what the reference's training step needs from ``_render_ijs``
(ngm/run_mapping.py:1164-1186 -- the prediction feeds ``_compute_losses`` and
``loss_dict["combined"].backward()`` must reach ``model.vmap_fields_params``).

Used by ``renderer.render_rays`` whenever autograd is recording and a field parameter requires
grad; evaluation (``torch.no_grad``) keeps the fused tcgen05 kernel.  Stage structure:

* sampler            -- ``ngm_sample_rays`` (CUDA; sample positions carry no gradient)
* world -> local     -- elementwise torch ops on constants (ngm/models.py:331-339)
* encoding           -- NeRF / Fourier / Triplane: their PyTorch expressions
                        (ngm/positional_encodings.py:245-272, 197-212, 132-161);
                        permutohedral: ``ngm_encode_fwd`` / ``ngm_encode_bwd`` (CUDA, table
                        gradient by atomics)
* MLP                -- precision "fp16" (skip mode "no", 1-4 hidden layers of width <= 128):
                        ``ngm_field_fwd`` / ``ngm_field_bwd`` -- the tcgen05 forward kernel and the
                        tcgen05 backward kernel (csrc/field_tc_bwd.cu: forward recomputed per tile,
                        weight gradients accumulated in TMEM, NeRF encoding inside the kernel);
                        otherwise batched GEMMs over the stacked per-field parameters
                        (ngm/models.py:143-182, all four skip modes; autograd derives the gradients)
* compositor         -- ``ngm_composite`` forward, ``ngm_composite_bwd`` backward (CUDA)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


# ------------------------------------------------------------------------------------------
# world -> field-local (pytorch3d quaternion_invert + quaternion_apply, ngm/models.py:331-335)
# ------------------------------------------------------------------------------------------
def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def world_to_local(points_world: torch.Tensor, positions: torch.Tensor, orientations: torch.Tensor,
                   scale_mode: str, field_radius) -> torch.Tensor:
    """(F, N, 3) world points -> scaled field-local coordinates (ngm/models.py:278-285, 331-339)."""
    conj = orientations.new_tensor([1.0, -1.0, -1.0, -1.0])
    q_inv = (orientations * conj)[:, None]                       # quaternion_invert
    local = points_world - positions[:, None]
    p = torch.cat((torch.zeros_like(local[..., :1]), local), -1)
    out = _quat_mul(_quat_mul(q_inv, p), q_inv * conj)           # quaternion_apply(q_inv, local)
    local = out[..., 1:]
    if scale_mode == "unit_cube":
        return local / (2.0 * field_radius) + 0.5
    if scale_mode == "unit_ball":
        return local / field_radius
    if scale_mode == "no":
        return local
    raise ValueError(f"scale_mode={scale_mode} is not available.")  # models.py:285


# ------------------------------------------------------------------------------------------
# encodings
# ------------------------------------------------------------------------------------------
class _PermutoEncodeFn(torch.autograd.Function):
    """lattice_values (F, L, capacity, feats) -> features (F, N, E); gradient to the table only."""

    @staticmethod
    def forward(ctx, table, proto, params, local):
        F, N = local.shape[0], local.shape[1]
        a = _lib.NgmEncodeArgs()
        p = dict(params)
        p["_encoding.lattice_values"] = table.detach()
        a.field, keep = proto.field_desc(p, True)
        pts = _lib.dev_f32(local, "points")
        out = torch.empty(F, N, proto._dim_encoding, device=pts.device)
        a.points_per_field, a.num_fields = N, F
        a.points, a.out = pts.data_ptr(), out.data_ptr()
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib.ngm_encode_fwd(C.byref(a), _lib.stream_ptr(pts.device)))
        ctx.proto, ctx.params = proto, {k: v.detach() for k, v in p.items()}
        ctx.save_for_backward(pts)
        ctx.table_shape = table.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        (pts,) = ctx.saved_tensors
        F, N = pts.shape[0], pts.shape[1]
        a = _lib.NgmEncodeArgs()
        a.field, keep = ctx.proto.field_desc(ctx.params, True)
        g = _lib.dev_f32(g_out, "d_out")
        d_table = torch.zeros(ctx.table_shape, device=pts.device)
        a.points_per_field, a.num_fields = N, F
        a.points, a.d_out, a.d_param0 = pts.data_ptr(), g.data_ptr(), d_table.data_ptr()
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib.ngm_encode_bwd(C.byref(a), _lib.stream_ptr(pts.device)))
        return d_table, None, None, None


def encode(proto, params: Dict[str, torch.Tensor], local: torch.Tensor) -> torch.Tensor:
    """(F, N, 3) field-local points -> (F, N, E) with the stacked encoding parameters of F fields."""
    enc = proto._encoding
    if enc.KIND == "nerf":  # positional_encodings.py:245-272
        octaves = torch.arange(enc.start_octave, enc.start_octave + enc.num_octaves, device=local.device,
                               dtype=torch.float)
        scaled = local.unsqueeze(-1) * (2**octaves * torch.pi)
        lead = local.shape[:-1]
        return torch.cat((torch.sin(scaled).reshape(*lead, -1), torch.cos(scaled).reshape(*lead, -1)), -1)
    if enc.KIND == "fourier":  # :197-212 (bias-free linear, then sin)
        w = params["_encoding._linear.weight"]  # (F, n, 3)
        feats = torch.sin(torch.bmm(local, w.transpose(1, 2)))
        return torch.cat((local, feats), -1) if enc._raw_coords else feats
    if enc.KIND == "triplane":  # :132-161, batched over fields
        coef = params["_encoding.plane_coef"]  # (F, 3, C, res, res)
        F, N = local.shape[0], local.shape[1]
        coord = torch.stack((local[..., [0, 1]], local[..., [0, 2]], local[..., [1, 2]]), 1)  # (F, 3, N, 2)
        feats = torch.nn.functional.grid_sample(
            coef.reshape(F * 3, *coef.shape[2:]), coord.reshape(F * 3, N, 1, 2), align_corners=True,
            padding_mode="border").reshape(F, 3, coef.shape[2], N)
        if enc.mode == "product":
            return feats.prod(1).transpose(1, 2)
        if enc.mode == "sum":
            return feats.sum(1).transpose(1, 2)
        return feats.reshape(F, 3 * coef.shape[2], N).transpose(1, 2)  # concat
    if enc.KIND == "permuto":
        return _PermutoEncodeFn.apply(params["_encoding.lattice_values"], proto, params, local)
    raise NotImplementedError(enc.KIND)


# ------------------------------------------------------------------------------------------
# MLP over stacked per-field parameters (ngm/models.py:143-182 under torch.vmap)
# ------------------------------------------------------------------------------------------
def field_forward(proto, params: Dict[str, torch.Tensor], local: torch.Tensor) -> torch.Tensor:
    """(F, N, 3) field-local points -> (F, N, dim_out); differentiable in ``params``.  Always fp32: the
    upstream gradients of mean-reduced losses sit far below the fp16 range, so fp16-operand backward GEMMs would
    need loss scaling the reference's training loop does not have (measured: 1.5x faster, gradients off)."""
    enc = outs = encode(proto, params, local)
    E = proto._dim_encoding
    for i in range(proto._num_layers + 1):
        prev = outs
        w, b = params[f"_linears.{i}.weight"], params[f"_linears.{i}.bias"]
        outs = torch.baddbmm(b[:, None], outs, w.transpose(1, 2))
        if i == proto._num_layers:
            break
        outs = torch.relu(outs)
        if proto._skip_mode == "concat":
            outs = torch.cat((outs, enc), -1)
        elif proto._skip_mode == "add":
            outs = torch.cat((outs[..., :E] + enc, outs[..., E:]), -1)
        elif proto._skip_mode == "rezero":
            alpha = params["_rezero"][:, i, None, None]
            if i == 0:
                outs = torch.cat((alpha * outs[..., :E] + prev, alpha * outs[..., E:]), -1)
            else:
                outs = alpha * outs + prev
    return outs


# ------------------------------------------------------------------------------------------
# MLP on tcgen05: ngm_field_fwd / ngm_field_bwd
# ------------------------------------------------------------------------------------------
def _rows_half(enc: torch.Tensor, ep: int) -> torch.Tensor:
    """(F, N, E) fp32 features -> (F * N, EP) fp16 rows, zero padded to the K multiple of 16."""
    F, N, E = enc.shape
    rows = torch.zeros(F * N, ep, device=enc.device, dtype=torch.float16)
    rows[:, :E] = enc.reshape(F * N, E)
    return rows


class _FieldTcFn(torch.autograd.Function):
    """``NeuralField.forward`` of F stacked fields on the tensor cores, differentiable in the linears' weights and
    biases (and in the encoding output when it is given as rows).  args: the L + 1 weights, then the L + 1 biases."""

    @staticmethod
    def forward(ctx, proto, geom, points, enc, *wb):
        nl = proto._num_layers + 1
        params = {f"_linears.{i}.weight": wb[i].detach() for i in range(nl)}
        params.update({f"_linears.{i}.bias": wb[nl + i].detach() for i in range(nl)})
        params.update(geom["enc_params"])
        dev = points.device
        F, N = points.shape[0], points.shape[1]
        a = _lib.NgmFieldFwdArgs()
        with torch.cuda.device(dev):
            a.field, keep = proto.field_desc(params, True)
            out = torch.empty(F, N, proto._dim_out, device=dev)
            a.points_per_field, a.num_fields = N, F
            pts = _lib.dev_f32(points, "points")
            a.points = pts.data_ptr()
            if geom["positions"] is not None:
                a.positions, a.orientations = geom["positions"].data_ptr(), geom["orientations"].data_ptr()
            a.out = out.data_ptr()
            a.scale_mode = _lib.SCALE[geom["scale_mode"]]
            a.field_radius = float(geom["field_radius"] or 0.0)
            a.precision = _lib.PREC["fp16"]
            rows = None
            if enc is not None:
                rows = _rows_half(enc.detach(), (proto._dim_encoding + 15) // 16 * 16)
                a.rows_half = rows.data_ptr()
            need = C.c_size_t(0)
            _lib.check(_lib.lib.ngm_field_fwd_workspace_bytes(C.byref(a), C.byref(need)))
            ws = torch.empty(max(need.value, 16), device=dev, dtype=torch.uint8)
            a.workspace, a.workspace_bytes = ws.data_ptr(), need.value
            _lib.check(_lib.lib.ngm_field_fwd(C.byref(a), _lib.stream_ptr(dev)))
        ctx.proto, ctx.geom, ctx.nl = proto, geom, nl
        ctx.enc_grad = enc is not None and enc.requires_grad
        ctx.enc_shape = None if enc is None else tuple(enc.shape)
        ctx.save_for_backward(pts, rows, *[t.detach() for t in wb])
        return out

    @staticmethod
    def backward(ctx, g_out):
        pts, rows, *wb = ctx.saved_tensors
        proto, geom, nl = ctx.proto, ctx.geom, ctx.nl
        params = {f"_linears.{i}.weight": wb[i] for i in range(nl)}
        params.update({f"_linears.{i}.bias": wb[nl + i] for i in range(nl)})
        params.update(geom["enc_params"])
        dev = pts.device
        F, N = pts.shape[0], pts.shape[1]
        b = _lib.NgmFieldBwdArgs()
        with torch.cuda.device(dev):
            b.fwd.field, keep = proto.field_desc(params, True)
            b.fwd.points_per_field, b.fwd.num_fields = N, F
            b.fwd.points = pts.data_ptr()
            if geom["positions"] is not None:
                b.fwd.positions, b.fwd.orientations = geom["positions"].data_ptr(), geom["orientations"].data_ptr()
            b.fwd.scale_mode = _lib.SCALE[geom["scale_mode"]]
            b.fwd.field_radius = float(geom["field_radius"] or 0.0)
            b.fwd.precision = _lib.PREC["fp16"]
            if rows is not None:
                b.fwd.rows_half = rows.data_ptr()
            g = _lib.dev_f32(g_out, "d_out")
            b.d_out = g.data_ptr()
            grads = [torch.empty_like(t) for t in wb]
            for i in range(nl):
                b.d_weights[i], b.d_biases[i] = grads[i].data_ptr(), grads[nl + i].data_ptr()
            d_enc = None
            if ctx.enc_grad:
                d_enc = torch.empty(ctx.enc_shape, device=dev)
                b.d_encoding = d_enc.data_ptr()
            need = C.c_size_t(0)
            _lib.check(_lib.lib.ngm_field_bwd_workspace_bytes(C.byref(b), C.byref(need)))
            ws = torch.empty(max(need.value, 16), device=dev, dtype=torch.uint8)
            b.fwd.workspace, b.fwd.workspace_bytes = ws.data_ptr(), need.value
            _lib.check(_lib.lib.ngm_field_bwd(C.byref(b), _lib.stream_ptr(dev)))
        return (None, None, None, d_enc, *grads)


def tc_training_supported(proto) -> bool:
    """Shapes the tcgen05 backward handles (csrc/field_tc_bwd.cu: field_bwd_tc_supported)."""
    w = proto._dim_mlp_out
    return (proto._skip_mode == "no" and 1 <= proto._num_layers <= 4 and w % 32 == 0 and 32 <= w <= 128
            and proto._dim_out <= 8 and proto._dim_encoding <= 64)


def field_forward_tc(proto, params: Dict[str, torch.Tensor], points_world: torch.Tensor, positions, orientations,
                     scale_mode: str, field_radius) -> torch.Tensor:
    """(F, N, 3) WORLD points -> (F, N, dim_out) through the tcgen05 kernels; differentiable in ``params``.
    NeRF encodings are evaluated inside the kernels; the permutohedral encoding through ``ngm_encode_fwd`` /
    ``ngm_encode_bwd`` and Fourier / Triplane through their PyTorch expressions, handed over as fp16 rows."""
    enc = None
    kind = proto._encoding.KIND
    enc_params = {k: v.detach() for k, v in params.items() if k.startswith("_encoding.")}
    pos = None if positions is None else _lib.dev_f32(positions, "positions")
    ori = None if orientations is None else _lib.dev_f32(orientations, "orientations")
    if kind != "nerf":
        with torch.no_grad():
            local = world_to_local(points_world, pos, ori, scale_mode, field_radius) if pos is not None else points_world
        enc = encode(proto, params, local)
    geom = dict(positions=pos, orientations=ori, scale_mode=scale_mode, field_radius=field_radius,
                enc_params=enc_params)
    nl = proto._num_layers + 1
    wb = [params[f"_linears.{i}.weight"] for i in range(nl)] + [params[f"_linears.{i}.bias"] for i in range(nl)]
    return _FieldTcFn.apply(proto, geom, points_world, enc, *wb)


# ------------------------------------------------------------------------------------------
# compositor
# ------------------------------------------------------------------------------------------
class _CompositeFn(torch.autograd.Function):
    """``ngm_composite`` / ``ngm_composite_bwd`` on the packed (N, S, 4) MLP output."""

    @staticmethod
    def forward(ctx, packed, isd, dist, depth, gt, cfg):
        dev = packed.device
        N, S = dist.shape
        pk = _lib.dev_f32(packed.detach(), "sample_outs")
        a = _lib.NgmCompositeArgs()
        _fill_composite_args(a, pk, isd, dist, depth, gt, cfg)
        rgbd = torch.empty(N, 4, device=dev)
        cvar = torch.empty(N, 3, device=dev)
        dvar = torch.empty(N, device=dev)
        term = torch.empty(N, device=dev)
        a.rgbd, a.color_var, a.depth_var, a.term_prob = rgbd.data_ptr(), cvar.data_ptr(), dvar.data_ptr(), term.data_ptr()
        fs = fs_m = ts = ts_m = None
        if cfg["want_freespace"]:
            fs, fs_m = torch.empty(N, S, device=dev), torch.empty(N, S, device=dev, dtype=torch.bool)
            a.freespace, a.freespace_mask = fs.data_ptr(), fs_m.data_ptr()
        if cfg["want_tsdf"]:
            ts, ts_m = torch.empty(N, S, device=dev), torch.empty(N, S, device=dev, dtype=torch.bool)
            a.tsdf, a.tsdf_mask = ts.data_ptr(), ts_m.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.ngm_composite(C.byref(a), _lib.stream_ptr(dev)))
        ctx.cfg = cfg
        ctx.has_isd = isd is not None
        ctx.save_for_backward(pk, isd, dist, depth, gt)
        empty = torch.empty(0, device=dev)
        outs = (rgbd, cvar, dvar, term, fs if fs is not None else empty, ts if ts is not None else empty,
                fs_m if fs_m is not None else empty.bool(), ts_m if ts_m is not None else empty.bool())
        ctx.mark_non_differentiable(outs[6], outs[7])
        return outs

    @staticmethod
    def backward(ctx, g_rgbd, g_cvar, g_dvar, g_term, g_fs, g_ts, _gm0, _gm1):
        pk, isd, dist, depth, gt = ctx.saved_tensors
        cfg = ctx.cfg
        dev = pk.device
        N, S = dist.shape
        b = _lib.NgmCompositeBwdArgs()
        _fill_composite_args(b.fwd, pk, isd, dist, depth, gt, cfg)
        keep = []

        def up(g, shape):
            if g is None:
                return None
            g = _lib.dev_f32(g, "upstream gradient").expand(shape).contiguous()
            keep.append(g)
            return g.data_ptr()

        b.g_rgbd, b.g_color_var = up(g_rgbd, (N, 4)), up(g_cvar, (N, 3))
        b.g_depth_var, b.g_term_prob = up(g_dvar, (N,)), up(g_term, (N,))
        if cfg["want_freespace"]:
            b.g_freespace = up(g_fs, (N, S))
        if cfg["want_tsdf"]:
            b.g_tsdf = up(g_ts, (N, S))
        d_packed = torch.empty(N, S, 4, device=dev)
        b.d_colors, b.d_geometries = d_packed.data_ptr(), d_packed.data_ptr() + 12
        d_isd_ray = None
        if ctx.has_isd:
            d_isd_ray = torch.empty(N, device=dev)
            b.d_neus_isd = d_isd_ray.data_ptr()
        ws = torch.empty(2 * N * S, device=dev)
        b.workspace = ws.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.ngm_composite_bwd(C.byref(b), _lib.stream_ptr(dev)))
        d_isd = d_isd_ray.view(-1, cfg["rays_per_isd"]).sum(1) if ctx.has_isd else None
        return d_packed, d_isd, None, None, None, None


def _fill_composite_args(a, pk, isd, dist, depth, gt, cfg) -> None:
    N, S = dist.shape
    a.num_rays, a.num_samples = N, S
    a.colors, a.geometries = pk.data_ptr(), pk.data_ptr() + 12
    a.color_stride = a.geometry_stride = 4
    a.distances, a.depths = dist.data_ptr(), depth.data_ptr()
    a.geometry_mode = _lib.GEOM[cfg["geometry_mode"]]
    a.geometry_factor, a.color_factor = cfg["geometry_factor"], cfg["color_factor"]
    a.truncation = cfg["truncation"]
    a.overwrite_behind_camera = int(cfg["overwrite"])
    a.overwrite_gate = _lib.ptr(cfg.get("overwrite_gate"))  # device flag of run_mapping.py:494-495
    if isd is not None:
        a.neus_isd, a.rays_per_isd = isd.data_ptr(), cfg["rays_per_isd"]
    a.gt = _lib.ptr(gt)


# ------------------------------------------------------------------------------------------
# the training-time _render_ijs (use_vmap=True)
# ------------------------------------------------------------------------------------------
def render_rays_vmap(driver, camera, ijs, c2ws, params, positions, orientations, near, far, gt, overwrite,
                     jitter, jitter_guided, seed, precision="fp32", sample_offset=0):
    """Differentiable twin of the fused renderer; returns the six ``Prediction`` members."""
    from .camera import sample_rays

    model = driver._model
    proto = model._prototype_field
    F, R = ijs.shape[0], ijs.shape[1]
    S = int(driver._num_samples)
    G = int(driver._num_samples_depth_guided) if gt is not None else 0
    St = S + G
    with torch.no_grad():
        _, dist, world, depth = sample_rays(
            camera, ijs, S, near if near is not None else float(driver._near_distance),
            far if far is not None else float(driver._far_distance), gt=gt, num_samples_guided=G,
            range_guided=float(driver._range_depth_guided or 0.0), c2ws=c2ws, jitter=jitter, jitter_guided=jitter_guided,
            seed=seed, offset=sample_offset, want_world=True, want_depth=True, want_cam=False)
    if precision == "fp16" and tc_training_supported(proto):
        # tensor-core training path: tcgen05 forward + tcgen05 backward (weight gradients accumulated in TMEM)
        outs = field_forward_tc(proto, params, world.view(F, R * St, 3), positions, orientations, model._scale_mode,
                                model._field_radius)
    else:
        with torch.no_grad():
            local = world_to_local(world.view(F, R * St, 3), _lib.dev_f32(positions, "positions"),
                                   _lib.dev_f32(orientations, "orientations"), model._scale_mode, model._field_radius)
        outs = field_forward(proto, params, local)  # (F, R*St, 4), differentiable in params, fp32 arithmetic
    mode = driver._geometry_mode
    isd = None
    if mode == "neus":
        isd = 1.0 / torch.abs(params["_neus_sd"].reshape(-1))  # run_mapping.py:641-644
    want_fs = driver._freespace_weight != 0.0 and gt is not None  # :624
    want_ts = driver._tsdf_weight != 0.0 and gt is not None       # :632
    cfg = dict(geometry_mode=mode, geometry_factor=float(driver._geometry_factor), color_factor=float(driver._color_factor),
               truncation=float(driver._truncation_distance or 0.0), overwrite=bool(overwrite), rays_per_isd=R,
               want_freespace=want_fs, want_tsdf=want_ts,
               overwrite_gate=((_lib.dev_f32(near, "near") < 0).any().to(torch.int32).reshape(1)
                               if overwrite and torch.is_tensor(near) else None))
    gt_flat = None if gt is None else _lib.dev_f32(gt, "gt_distances").expand(F, R).reshape(-1).contiguous()
    rgbd, cvar, dvar, term, fs, ts, fs_m, ts_m = _CompositeFn.apply(
        outs.reshape(F * R, St, 4), isd, dist.view(F * R, St), depth.view(F * R, St), gt_flat, cfg)
    # 1-D tensors of data-dependent length (:628, :637).  masked_select = the reference's boolean-mask indexing, but
    # its backward is one masked_scatter instead of index_put's sort + accumulate (18 radix-sort launches per step)
    freespace = torch.masked_select(fs, fs_m) if want_fs else None
    tsdf = torch.masked_select(ts, ts_m) if want_ts else None
    return rgbd.view(F, R, 4), cvar.view(F, R, 3), dvar.view(F, R), term.view(F, R), freespace, tsdf
