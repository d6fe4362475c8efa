"""Multi-view target sampling: drop-in for ``NeuralGraphMap._sample_target_mv`` (ngm/run_mapping.py:1261-1459).

The choice of fields (a few multinomial draws over at most ``num_fields`` entries, :1295-1315) and the keyframe
draw (:1381-1383) stay torch calls on the device -- same functions, same generator, so seeded runs draw what the
reference draws; the two data-parallel halves around them are one CUDA launch each (``ngm_target_visibility``,
``ngm_target_rays``) instead of ~60 small kernels over (fields x 20 probes x keyframes) intermediates.
"""
import ctypes as C
from collections import namedtuple
from typing import Optional

import torch

from . import _lib
from .camera import camera_struct

# ngm/run_mapping.py:43-58
Target = namedtuple("Target", ["ijs", "c2ws", "near_distances", "far_distances", "gt_distances", "field_ids", "rgbds",
                               "rgb_mask", "depth_mask", "term_probs", "term_mask"])

NUM_FIELD_SAMPLES = 20  # :1291


def _aligned(t: torch.Tensor, nbytes: int) -> torch.Tensor:
    """The kernels read poses / pixels / boxes as 16- or 8-byte vectors; a view that starts off that grid is copied."""
    return t if t.data_ptr() % nbytes == 0 else t.clone()


def _store(c2ws, rgbds, frame_to_store, positions, camera):
    c2ws = _aligned(_lib.dev_f32(c2ws, "_c_c2w_tensor"), 16)
    rgbds = _aligned(_lib.dev_f32(rgbds, "_nc_rgbd_tensor"), 16)
    positions = _lib.dev_f32(positions, "positions")
    if c2ws.dim() != 3 or c2ws.shape[1:] != (4, 4):
        raise ValueError(f"keyframe poses must be (num_frames, 4, 4), got {tuple(c2ws.shape)}")
    if rgbds.dim() != 4 or rgbds.shape[1:] != (camera.height, camera.width, 4):
        raise ValueError(f"keyframe buffer must be (num_stored, {camera.height}, {camera.width}, 4), got {tuple(rgbds.shape)}")
    f2s = None
    if frame_to_store is not None:
        f2s = frame_to_store.to(device=c2ws.device, dtype=torch.int64).contiguous()
        if f2s.numel() != c2ws.shape[0]:
            raise ValueError("_frame_cid_to_ncid and _c_c2w_tensor disagree on the number of frames")
    return c2ws, rgbds, f2s, positions


def target_visibility(camera, c2ws, rgbds, frame_to_store, positions, field_ids, probe_offsets, train_radius):
    """``ngm_target_visibility``: (field_kf_mask (F, K) bool, min_xys (F, K, 2), max_xys (F, K, 2))."""
    c2ws, rgbds, f2s, positions = _store(c2ws, rgbds, frame_to_store, positions, camera)
    dev = c2ws.device
    ids = field_ids.to(device=dev, dtype=torch.int64).contiguous()
    probes = _lib.dev_f32(probe_offsets, "probe_offsets")
    F, K = ids.numel(), c2ws.shape[0]
    mask = torch.empty(F, K, dtype=torch.bool, device=dev)
    lo, hi = torch.empty(F, K, 2, device=dev), torch.empty(F, K, 2, device=dev)
    a = _lib.NgmTargetVisArgs()
    a.cam = camera_struct(camera)
    a.c2ws, a.rgbds, a.frame_to_store, a.positions = c2ws.data_ptr(), rgbds.data_ptr(), _lib.ptr(f2s), positions.data_ptr()
    a.field_ids, a.probe_offsets = ids.data_ptr(), probes.data_ptr()
    a.num_frames, a.num_fields, a.num_probes, a.train_radius = K, F, probes.shape[0], float(train_radius)
    a.field_kf_mask, a.min_xys, a.max_xys = mask.data_ptr(), lo.data_ptr(), hi.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ngm_target_visibility(C.byref(a), _lib.stream_ptr(dev)))
    return mask, lo, hi


def target_rays(camera, c2ws, rgbds, frame_to_store, positions, field_ids, frame_cids, uv, min_xys, max_xys,
                train_radius) -> Target:
    """``ngm_target_rays``: the reference's ``Target`` for the drawn (field, keyframe, pixel) triples."""
    c2ws, rgbds, f2s, positions = _store(c2ws, rgbds, frame_to_store, positions, camera)
    dev = c2ws.device
    ids = field_ids.to(device=dev, dtype=torch.int64).contiguous()
    cids = frame_cids.to(device=dev, dtype=torch.int64).contiguous()
    uv = _aligned(_lib.dev_f32(uv, "uv"), 8)
    lo, hi = _aligned(_lib.dev_f32(min_xys, "min_xys"), 8), _aligned(_lib.dev_f32(max_xys, "max_xys"), 8)
    F, R, K = ids.numel(), cids.shape[1] if cids.dim() == 2 else 0, c2ws.shape[0]
    if cids.shape != (F, R) or uv.shape != (F, R, 2) or lo.shape != (F, K, 2) or hi.shape != (F, K, 2):
        raise ValueError("frame_cids (F,R), uv (F,R,2) and min/max_xys (F,K,2) disagree")
    f32 = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
    b8 = lambda *s: torch.empty(*s, device=dev, dtype=torch.bool)  # noqa: E731
    t = Target(ijs=torch.empty(F, R, 2, dtype=torch.int64, device=dev), c2ws=f32(F, R, 4, 4), near_distances=f32(F, R),
               far_distances=f32(F, R), gt_distances=f32(F, R), field_ids=ids, rgbds=f32(F, R, 4), rgb_mask=b8(F, R),
               depth_mask=b8(F, R), term_probs=f32(F, R), term_mask=b8(F, R))
    a = _lib.NgmTargetRaysArgs()
    a.cam = camera_struct(camera)
    a.c2ws, a.rgbds, a.frame_to_store, a.positions = c2ws.data_ptr(), rgbds.data_ptr(), _lib.ptr(f2s), positions.data_ptr()
    a.field_ids, a.frame_cids, a.uv, a.min_xys, a.max_xys = ids.data_ptr(), cids.data_ptr(), uv.data_ptr(), lo.data_ptr(), hi.data_ptr()
    a.num_frames, a.rays_per_field, a.num_fields, a.train_radius = K, R, F, float(train_radius)
    a.ijs, a.out_c2ws, a.near, a.far, a.gt = (t.ijs.data_ptr(), t.c2ws.data_ptr(), t.near_distances.data_ptr(),
                                              t.far_distances.data_ptr(), t.gt_distances.data_ptr())
    a.out_rgbds, a.rgb_mask, a.depth_mask = t.rgbds.data_ptr(), t.rgb_mask.data_ptr(), t.depth_mask.data_ptr()
    a.term_probs, a.term_mask = t.term_probs.data_ptr(), t.term_mask.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ngm_target_rays(C.byref(a), _lib.stream_ptr(dev)))
    return t


def sample_target_mv(driver, current_field_ids: torch.Tensor, draws: Optional[dict] = None) -> Target:
    """Drop-in for ``NeuralGraphMap._sample_target_mv``.  ``draws`` (parity runs) may inject the random numbers the
    reference draws, in its order: ``subset_observed``, ``subset_random`` (torch.multinomial over fields),
    ``probe_offsets`` (the torch.randn (20, 3) before normalisation), ``frame_cids`` (torch.multinomial over the
    visibility mask) and ``uv`` (torch.rand (F, R, 2))."""
    draws = draws or {}
    dev = driver._device
    train_radius = driver._field_radius + 0.0  # MARGIN = 0.0 (:1289-1290)
    num_fields = getattr(driver, "_num_fields", None)
    if num_fields is None:
        num_fields = driver._global_map_dict["num"]
    with torch.no_grad():
        # ---- which fields (:1295-1317)
        num_observed = min(driver._num_train_fields // 2, len(current_field_ids))
        subset_observed = draws.get("subset_observed")
        if subset_observed is None:
            subset_observed = torch.multinomial(torch.ones(len(current_field_ids), device=dev), num_observed)
        observed = current_field_ids[subset_observed.to(dev)]
        num_random = min(driver._num_train_fields - len(observed), num_fields - len(observed))
        if num_random > 0:
            subset_random = draws.get("subset_random")
            if subset_random is None:
                dist = torch.ones(num_fields, device=dev)
                dist[observed] = 0.0
                subset_random = torch.multinomial(dist, num_random)
            field_ids = torch.unique(torch.cat((subset_random.to(dev), observed)))  # get_field_ids() = arange (:2180)
        else:
            field_ids = observed
        # ---- where each field can be supervised (:1321-1362, 1386-1389)
        offsets = draws.get("probe_offsets")
        if offsets is None:
            offsets = torch.randn((NUM_FIELD_SAMPLES, 3), device=dev)
        offsets = offsets.to(dev) / torch.linalg.norm(offsets.to(dev), dim=-1, keepdim=True)
        store = (driver._c_c2w_tensor, driver._nc_rgbd_tensor, driver._frame_cid_to_ncid,
                 driver._global_map_dict["positions"])
        mask, lo, hi = target_visibility(driver._camera, *store, field_ids, offsets, train_radius)
        # ---- only fields some keyframe sees (:1364-1378); data-dependent size, as in the reference
        field_mask = mask.any(dim=-1)
        mask, field_ids, lo, hi = mask[field_mask], field_ids[field_mask], lo[field_mask], hi[field_mask]
        # ---- keyframe and pixel of every ray (:1381-1408)
        R = driver._num_rays_per_field
        frame_cids = draws.get("frame_cids")
        if frame_cids is None:
            frame_cids = torch.multinomial(mask.float(), R, replacement=True)
        uv = draws.get("uv")
        if uv is None:
            uv = torch.rand(len(field_ids), R, 2, device=dev)
        return target_rays(driver._camera, *store, field_ids, frame_cids, uv, lo, hi, train_radius)


NUM_OBSERVED_POINTS = 500  # :1651


def observed_fields(camera, depth_image: torch.Tensor, pixel_ids: torch.Tensor, c2w: torch.Tensor,
                    positions: torch.Tensor, field_radius: float) -> torch.Tensor:
    """``ngm_observed_fields``: bool (num_fields,) -- field observed by the back-projected pixels ``pixel_ids``
    (row-major indices into ``depth_image``, which may be a strided channel view of an (H, W, 4) RGB-D image)."""
    if not depth_image.is_cuda:
        raise RuntimeError("depth image must be a CUDA tensor: neural_graph_mapping_b200 has no CPU path")
    dev = depth_image.device
    if depth_image.dtype != torch.float32 or depth_image.shape != (camera.height, camera.width):
        raise ValueError(f"depth image must be fp32 ({camera.height}, {camera.width}), got {depth_image.dtype} {tuple(depth_image.shape)}")
    sh, sw = depth_image.stride()
    if sh != sw * camera.width:  # not a plain channel view: make it one
        depth_image = depth_image.contiguous()
        sh, sw = depth_image.stride()
    pix = pixel_ids.to(device=dev, dtype=torch.int64).contiguous()
    c2w = _aligned(_lib.dev_f32(c2w, "c2w"), 16)
    positions = _lib.dev_f32(positions, "positions")
    out = torch.empty(positions.shape[0], dtype=torch.bool, device=dev)
    a = _lib.NgmObservedArgs()
    a.cam = camera_struct(camera)
    a.depth, a.pixel_ids, a.c2w, a.positions = depth_image.data_ptr(), pix.data_ptr(), c2w.data_ptr(), positions.data_ptr()
    a.pixel_stride, a.num_points, a.num_fields, a.field_radius = sw, pix.numel(), positions.shape[0], float(field_radius)
    a.observed = out.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ngm_observed_fields(C.byref(a), _lib.stream_ptr(dev)))
    return out


def get_observed_fields(driver, rgbd_image: torch.Tensor, c2w: torch.Tensor, draws: Optional[dict] = None) -> torch.Tensor:
    """Drop-in for ``NeuralGraphMap._get_observed_fields`` (ngm/run_mapping.py:1643-1670).  The choice of pixels is
    the reference's (``torch.nonzero`` of the depth channel, ``torch.multinomial`` of 500 of them -- same functions,
    same generator); ``draws["subset"]`` injects it for parity runs."""
    dev = driver._device
    num_fields = getattr(driver, "_num_fields", None)
    if num_fields is None:
        num_fields = driver._global_map_dict["num"]
    with torch.no_grad():
        depth = rgbd_image[..., 3]
        ijs = torch.nonzero(depth)  # camera.py:373
        subset = (draws or {}).get("subset")
        if subset is None:
            subset = torch.multinomial(torch.ones(len(ijs), device=dev), NUM_OBSERVED_POINTS)
        ijs = ijs[subset.to(dev)]
        pix = ijs[:, 0] * driver._camera.width + ijs[:, 1]
        observed = observed_fields(driver._camera, depth, pix, c2w, driver._global_map_dict["positions"][:num_fields],
                                   driver._field_radius)
        return torch.nonzero(observed)[:, 0]
