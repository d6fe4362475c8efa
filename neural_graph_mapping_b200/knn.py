"""kNN (use_vmap=False) branch: ``NeuralFieldSet.forward`` over ALL fields with the K=2 blend
(ngm/models.py:347-405) and the eval-shape ``_render_ijs`` that uses it (ngm/run_mapping.py:586-595).
All arithmetic is in libngm_b200 (csrc/knn.cu + the gather mode of the field kernel)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, models
from .camera import sample_rays


POINT_BLOCK_BYTES = 6 << 30
POINT_BYTES = 160  # pair tables, bucketed entries, per-pair outputs (K = 2) + fp16 rows of the widest row encoding


def fieldset_forward_knn(model, query_points, field_positions, field_orientations, field_ids, field_radius,
                         precision=None):
    """``ngm_fieldset_knn_fwd``: (..., 3) world points -> (..., 4).  ``precision`` "fp16" evaluates the
    bucketed (point, neighbour) entries with the tcgen05 field kernel in gather mode."""
    precision = precision or getattr(model, "precision", None) or models.get_default_precision()
    if field_positions is None or field_orientations is None:
        raise ValueError("the kNN path needs field_positions and field_orientations")
    params = model.all_fields_params
    models._no_autograd(query_points, *params.values())
    leading = tuple(query_points.shape[:-1])
    with torch.no_grad():
        pts = _lib.dev_f32(query_points, "query_points").reshape(-1, 3)
        dev = pts.device
        pos = _lib.dev_f32(field_positions, "field_positions")
        ori = _lib.dev_f32(field_orientations, "field_orientations")
        F = pos.shape[0]
        n = pts.shape[0]
        out = torch.empty(n, 4, device=dev)
        a = _lib.NgmKnnFwdArgs()
        with torch.cuda.device(dev):
            a.field, keep = model._prototype_field.field_desc(
                params, True, model.packed_images(params) if precision == "fp16" else None)
            a.num_points, a.num_fields = n, F
            a.points, a.positions, a.orientations = pts.data_ptr(), pos.data_ptr(), ori.data_ptr()
            if field_ids is not None:
                slots = field_ids.to(device=dev, dtype=torch.int64).contiguous()
                keep.append(slots)
                a.field_slots = slots.data_ptr()
            a.out = out.data_ptr()
            a.field_radius = float(field_radius if field_radius is not None else model._field_radius)
            a.scale_radius = float(model._field_radius or 0.0)
            a.distance_factor = float(model._distance_factor)
            a.outside_value = float(model._outside_value)
            a.num_knn = int(model._num_knn)
            a.scale_mode = _lib.SCALE[model._scale_mode]
            a.precision = _lib.PREC[precision]
            need = C.c_size_t(0)
            _lib.check(_lib.lib.ngm_fieldset_knn_workspace_bytes(C.byref(a), C.byref(need)))
            ws = torch.empty(max(need.value, 16), device=dev, dtype=torch.uint8)
            a.workspace, a.workspace_bytes = ws.data_ptr(), need.value
            _lib.check(_lib.lib.ngm_fieldset_knn_fwd(C.byref(a), _lib.stream_ptr(dev)))
    return out.reshape(*leading, 4)


def render_rays_knn(driver, ijs, c2ws, camera, field_ids, near, far, gt, overwrite, jitter, seed=None, sample_offset=0):
    """``_render_ijs`` with use_vmap=False (ngm/run_mapping.py:586-595): sampler stage -> kNN
    field set (in blocks of ``_block_size`` points, like utils.batched_evaluation) -> compositor."""
    from .renderer import Prediction, _next_seed, _overwrite_gate, _precision, composite

    model = driver._model
    dev = ijs.device
    if field_ids is not None:
        positions = driver._global_map_dict["positions"][field_ids]
        orientations = driver._global_map_dict["orientations"][field_ids]
    else:
        num = driver._global_map_dict["num"]
        positions = driver._global_map_dict["positions"][:num]
        orientations = driver._global_map_dict["orientations"][:num]
    leading = tuple(ijs.shape[:-1])
    S = int(driver._num_samples)
    G = int(driver._num_samples_depth_guided) if gt is not None else 0
    with torch.no_grad():
        _, dist, world, depth = sample_rays(
            camera, ijs, S, driver._near_distance if near is None else near,
            driver._far_distance if far is None else far, gt=gt, num_samples_guided=G,
            range_guided=float(driver._range_depth_guided or 0.0), c2ws=c2ws, jitter=jitter,
            seed=0 if jitter is not None else (_next_seed() if seed is None else int(seed)), offset=int(sample_offset),
            want_world=True, want_depth=True, want_cam=False)
        St = dist.shape[-1]
        pts = world.reshape(-1, 3)
        # The reference evaluates the field set in blocks of `block_size` (3 M) points to bound the memory of its
        # PyTorch intermediates (run_mapping.py:588).  Points are independent; the CUDA path holds ~150 B per point in
        # flight, so its blocks are as large as POINT_BLOCK_BYTES allows and never smaller than the configured size.
        block = max(int(driver._block_size), POINT_BLOCK_BYTES // POINT_BYTES)
        if pts.shape[0] <= block:
            o = fieldset_forward_knn(model, pts, positions, orientations, field_ids, None, _precision(driver))
        else:
            o = torch.empty(pts.shape[0], 4, device=pts.device)
            for s0 in range(0, pts.shape[0], block):
                o[s0:s0 + block] = fieldset_forward_knn(model, pts[s0:s0 + block], positions, orientations, field_ids,
                                                        None, _precision(driver))
        n = dist.numel() // St
        gt_t = None if gt is None else _lib.dev_f32(gt, "gt").expand(leading).reshape(-1).contiguous()
        want_fs = driver._freespace_weight != 0.0 and gt is not None
        want_ts = driver._tsdf_weight != 0.0 and gt is not None
        gate = _overwrite_gate(_lib.dev_f32(near, "near")) if overwrite and torch.is_tensor(near) else None
        rgbd, cvar, dvar, term, _, aux = composite(
            o, o[:, 3], dist.reshape(n, St), depth.reshape(n, St), driver._geometry_mode, driver._geometry_factor,
            driver._color_factor, gt=gt_t, truncation=float(driver._truncation_distance or 0.0),
            overwrite_behind_camera=overwrite, want_aux=(want_fs, want_ts), color_stride=4, geometry_stride=4,
            overwrite_gate=gate)
        fs, fs_m, ts, ts_m = aux
        freespace = fs[fs_m] if fs is not None else None
        tsdf = ts[ts_m] if ts is not None else None
    return Prediction(rgbd.reshape(*leading, 4), cvar.reshape(*leading, 3), dvar.reshape(leading),
                      term.reshape(leading), freespace, tsdf)
