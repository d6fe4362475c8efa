"""The ray renderer behind the reference's Python surface.

``render_rays`` is signature- and semantics-compatible with
``NeuralGraphMap._render_ijs`` (ngm/run_mapping.py:440-666), ``quadrature`` with
``_quadrature`` (:709-799) and ``render_image`` with ``render_image`` (:402-437).  They read
the same driver attributes the reference methods read (``_num_samples``, ``_near_distance``,
``_geometry_mode``, ``_model``, ``_global_map_dict`` ...), so they can be

* installed on the reference's own driver class:  ``install(NeuralGraphMap)``  (INTEGRATION.md), or
* used stand-alone through ``RenderState`` (same attribute names, built from the same YAML keys).
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Optional, Sequence, Tuple

import torch

from . import _lib, models
from .camera import camera_struct
from .utils import batched_evaluation, str_to_object

# ngm/run_mapping.py:59-69
Prediction = namedtuple(
    "Prediction",
    ["rgbds", "color_vars", "depth_vars", "term_probs", "freespace_geometry", "tsdf_residuals"],
)


_SEED_BOX = torch.empty((), dtype=torch.int64)


def _next_seed() -> int:
    """Seed of the in-kernel jitter stream, drawn from torch's CPU generator (so torch.manual_seed makes renders
    reproducible, like the reference's torch.rand): one in-place draw into a preallocated CPU scalar, no device work."""
    return int(_SEED_BOX.random_(0, 2**62))


_WORKSPACES: dict = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand and reused by every render call on that stream
    (stream order makes the reuse safe; calls on different streams get different buffers)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes * 1.25), 1 << 16), device=dev, dtype=torch.uint8)
        _WORKSPACES[key] = ws
    return ws


def _overwrite_gate(near_t: torch.Tensor) -> torch.Tensor:
    """Device flag of ngm/run_mapping.py:494-495: the behind-camera overwrite stays on only if SOME near distance
    is negative.  Evaluated on the device (the reference synchronises the host with ``.all()`` here)."""
    return (near_t < 0).any().to(torch.int32).reshape(1)


def _precision(driver) -> str:
    return getattr(driver, "_precision", None) or getattr(driver._model, "precision", None) \
        or models.get_default_precision()


def render_rays(
    driver,
    ijs: torch.Tensor,
    c2ws: torch.Tensor,
    camera,
    field_ids: Optional[torch.Tensor] = None,
    use_vmap: bool = False,
    near_distances: Optional[torch.Tensor] = None,
    far_distances: Optional[torch.Tensor] = None,
    gt_distances: Optional[torch.Tensor] = None,
    overwrite_samples_behind_camera: bool = True,
    jitter: Optional[torch.Tensor] = None,
    jitter_guided: Optional[torch.Tensor] = None,
    out: Optional[tuple] = None,
    seed: Optional[int] = None,
    sample_offset: int = 0,
    mirrors: Optional[Sequence[int]] = None,
) -> Prediction:
    """Drop-in for ``NeuralGraphMap._render_ijs`` (ngm/run_mapping.py:440-666).

    Extra keyword-only-in-practice arguments ``jitter`` / ``jitter_guided`` inject the
    stratified-sampling noise (parity runs); by default it is generated in-kernel (Philox).
    ``out`` = preallocated (rgbds, color_vars, depth_vars, term_probs) to render into (used by the
    multi-GPU all-gather so the tiles land directly in the send buffer).  ``seed`` / ``sample_offset`` pin the
    in-kernel jitter stream: sample k of ray r draws element ``sample_offset + r * St + k`` of stream ``seed``, so
    a shard of a batch rendered with the batch's seed and its first sample's global index as offset draws
    exactly what the whole batch would (multi-GPU field sharding).  ``mirrors`` = byte offsets from the ``out``
    tensors to peer / multicast mappings of the same tile (``distributed.TileExchange``): the fused kernel repeats
    every Prediction store there, which replaces the all-gather.
    """
    if use_vmap and field_ids is None:
        raise ValueError("field_ids=None only supported for use_vmap=False")  # run_mapping.py:497-498
    if getattr(driver, "_single_field_id", None) is not None:
        raise NotImplementedError("single_field_id debugging mode is not implemented by the CUDA path")
    if not ijs.is_cuda:
        raise RuntimeError("render_rays needs CUDA tensors: neural_graph_mapping_b200 has no CPU path")
    # run_mapping.py:494-495: no overwrite without near distances, or when all of them are >= 0 -- the second half
    # is a device flag (`_overwrite_gate`) the kernels read, so the decision costs no host synchronisation.  (It is
    # not implied by the per-sample test: depth-guided samples can lie behind the camera although near >= 0.)
    overwrite = bool(overwrite_samples_behind_camera) and near_distances is not None
    if not use_vmap:
        from .knn import render_rays_knn

        if mirrors:
            raise NotImplementedError("mirrors= exists on the use_vmap=True path only")

        return render_rays_knn(driver, ijs, c2ws, camera, field_ids, near_distances, far_distances,
                               gt_distances, overwrite, jitter, seed, sample_offset)

    model = driver._model
    dev = ijs.device
    if ijs.dim() != 3:
        raise ValueError("use_vmap=True expects ijs of shape (num_fields, num_rays, 2)")
    F, R = ijs.shape[0], ijs.shape[1]
    if len(field_ids) != F:
        raise ValueError("field_ids must have one entry per leading row of ijs")
    keep = []
    a = _lib.NgmRenderArgs()
    proto = model._prototype_field
    if getattr(driver, "_reference_flow", True) and hasattr(driver, "_set_vmap_fields"):
        # the reference driver: keep its side effects (gather copy + optimizer remap, :500,668-707)
        driver._set_vmap_fields(field_ids)
        params = model.vmap_fields_params
        positions = driver._global_map_dict["positions"][field_ids]
        orientations = driver._global_map_dict["orientations"][field_ids]
        slots = None
    else:
        # stand-alone: read the stacked tables in place through field_slots (no gather copy)
        params = model.all_fields_params
        positions = driver._global_map_dict["positions"]
        orientations = driver._global_map_dict["orientations"]
        slots = field_ids.to(device=dev, dtype=torch.int64).contiguous()
    if torch.is_grad_enabled() and any(t.requires_grad for t in params.values()):
        # training (run_mapping.py:1164-1186): autograd must reach the field parameters
        from . import autograd as ag

        if slots is not None:  # gather the active fields (differentiable; what set_vmap_fields does, models.py:274-276)
            params = {k: v[slots] for k, v in params.items()}
            positions, orientations = positions[slots], orientations[slots]
        if out is not None or mirrors:
            raise ValueError("out= / mirrors= are not supported on the differentiable path")
        return Prediction(*ag.render_rays_vmap(
            driver, camera, ijs, c2ws, params, positions, orientations, near_distances, far_distances, gt_distances,
            overwrite, jitter, jitter_guided,
            0 if jitter is not None else (_next_seed() if seed is None else int(seed)), _precision(driver),
            sample_offset=int(sample_offset)))
    with torch.no_grad(), torch.cuda.device(dev):
        packed = model.packed_images(params) if (slots is not None and _precision(driver) == "fp16") else None
        a.field, k2 = proto.field_desc(params, True, packed)
        keep += k2
        a.cam = camera_struct(camera)
        a.num_fields, a.rays_per_field = F, R
        ij = ijs.to(torch.int64).contiguous()
        a.ijs = ij.data_ptr()
        c2w_t = _lib.dev_f32(c2ws, "c2ws")
        if c2w_t.numel() == 16:
            a.c2w_per_ray = 0
        else:
            c2w_t = c2w_t.expand(F, R, 4, 4).contiguous()
            a.c2w_per_ray = 1
        a.c2ws = c2w_t.data_ptr()

        def per_ray(x, name):
            return None if x is None else _lib.dev_f32(x, name).expand(F, R).contiguous()

        near_t, far_t, gt_t = per_ray(near_distances, "near"), per_ray(far_distances, "far"), per_ray(gt_distances, "gt")
        a.near, a.far, a.gt = _lib.ptr(near_t), _lib.ptr(far_t), _lib.ptr(gt_t)
        a.near_scalar, a.far_scalar = float(driver._near_distance), float(driver._far_distance)
        S = int(driver._num_samples)
        G = int(driver._num_samples_depth_guided) if gt_t is not None else 0
        if G > 0 and (near_t is None or far_t is None):
            raise ValueError("depth-guided sampling needs per-ray near/far distances")  # :522-525 would fail
        St = S + G
        a.num_samples, a.num_samples_guided = S, G
        a.range_guided = float(driver._range_depth_guided or 0.0)
        jt = None if jitter is None else _lib.dev_f32(jitter, "jitter")
        jg = None if jitter_guided is None else _lib.dev_f32(jitter_guided, "jitter_guided")
        a.jitter, a.jitter_guided = _lib.ptr(jt), _lib.ptr(jg)
        a.seed = 0 if jt is not None else (_next_seed() if seed is None else int(seed))
        a.offset = int(sample_offset)
        pos = _lib.dev_f32(positions, "positions")
        ori = _lib.dev_f32(orientations, "orientations")
        a.positions, a.orientations, a.field_slots = pos.data_ptr(), ori.data_ptr(), _lib.ptr(slots)
        a.scale_mode = _lib.SCALE[model._scale_mode]
        a.field_radius = float(model._field_radius or 0.0)
        a.geometry_mode = _lib.GEOM[driver._geometry_mode]
        a.geometry_factor, a.color_factor = float(driver._geometry_factor), float(driver._color_factor)
        a.truncation = float(driver._truncation_distance or 0.0)
        a.overwrite_behind_camera = int(overwrite)
        if overwrite:
            gate = _overwrite_gate(near_t)
            keep.append(gate)
            a.overwrite_gate = gate.data_ptr()
        a.precision = _lib.PREC[_precision(driver)]
        if driver._geometry_mode == "neus":
            sd = _lib.dev_f32(params["_neus_sd"], "_neus_sd")
            keep.append(sd)
            a.neus_sd = sd.data_ptr()
        if out is not None:
            rgbd, cvar, dvar, term = out
            for t_, shp in ((rgbd, (F, R, 4)), (cvar, (F, R, 3)), (dvar, (F, R)), (term, (F, R))):
                if tuple(t_.shape) != shp or not t_.is_contiguous() or t_.dtype != torch.float32:
                    raise ValueError("out tensors must be contiguous fp32 of the Prediction shapes")
            if rgbd.data_ptr() % 16 != 0:  # the kernels store a ray's rgbd as one 16-byte vector
                raise ValueError("out[0] (rgbds) must be 16-byte aligned")
        else:
            rgbd = torch.empty(F, R, 4, device=dev)
            cvar = torch.empty(F, R, 3, device=dev)
            dvar = torch.empty(F, R, device=dev)
            term = torch.empty(F, R, device=dev)
        a.rgbd, a.color_var, a.depth_var, a.term_prob = rgbd.data_ptr(), cvar.data_ptr(), dvar.data_ptr(), term.data_ptr()
        if mirrors:
            if out is None:
                raise ValueError("mirrors= needs out= (views of the symmetric tile buffer)")
            if len(mirrors) > len(a.mirror_delta):
                raise ValueError(f"at most {len(a.mirror_delta)} mirrors")
            a.num_mirrors = len(mirrors)
            for i, d in enumerate(mirrors):
                a.mirror_delta[i] = int(d)
        fs = fs_m = ts = ts_m = None
        if driver._freespace_weight != 0.0 and gt_t is not None:  # :624
            fs = torch.empty(F, R, St, device=dev)
            fs_m = torch.empty(F, R, St, device=dev, dtype=torch.bool)
            a.freespace, a.freespace_mask = fs.data_ptr(), fs_m.data_ptr()
        if driver._tsdf_weight != 0.0 and gt_t is not None:  # :632
            ts = torch.empty(F, R, St, device=dev)
            ts_m = torch.empty(F, R, St, device=dev, dtype=torch.bool)
            a.tsdf, a.tsdf_mask = ts.data_ptr(), ts_m.data_ptr()
        need = C.c_size_t(0)
        _lib.check(_lib.lib.ngm_render_workspace_bytes(C.byref(a), C.byref(need)))
        ws = _workspace(dev, need.value)
        a.workspace, a.workspace_bytes = ws.data_ptr(), need.value
        _lib.check(_lib.lib.ngm_render_rays_fwd(C.byref(a), _lib.stream_ptr(dev)))
        # the reference returns 1-D masked tensors (data-dependent length, :628,637)
        freespace = fs[fs_m] if fs is not None else None
        tsdf = ts[ts_m] if ts is not None else None
    return Prediction(rgbd, cvar, dvar, term, freespace, tsdf)


def composite_args(colors, geometries, distances, depths, geometry_mode, geometry_factor, color_factor=1.0,
                   neus_isd=None, rays_per_isd=1, gt=None, truncation=0.0, overwrite_behind_camera=False,
                   want_weights=False, want_aux=(False, False), color_stride=None, geometry_stride=None,
                   overwrite_gate=None):
    """The ``NgmCompositeArgs`` of one ``ngm_composite`` call and its freshly allocated outputs."""
    dev = distances.device
    N, S = distances.shape
    a = _lib.NgmCompositeArgs()
    a.num_rays, a.num_samples = N, S
    a.colors, a.geometries = colors.data_ptr(), geometries.data_ptr()
    a.color_stride = color_stride if color_stride is not None else 3
    a.geometry_stride = geometry_stride if geometry_stride is not None else 1
    a.distances, a.depths = distances.data_ptr(), depths.data_ptr()
    a.geometry_mode = _lib.GEOM[geometry_mode]
    a.geometry_factor, a.color_factor, a.truncation = float(geometry_factor), float(color_factor), float(truncation)
    a.overwrite_behind_camera = int(overwrite_behind_camera)
    a.overwrite_gate = _lib.ptr(overwrite_gate)  # device int32 flag (run_mapping.py:494-495) or None = unconditional
    if neus_isd is not None:
        a.neus_isd, a.rays_per_isd = neus_isd.data_ptr(), rays_per_isd
    a.gt = _lib.ptr(gt)
    rgbd = torch.empty(N, 4, device=dev)
    cvar = torch.empty(N, 3, device=dev)
    dvar = torch.empty(N, device=dev)
    term = torch.empty(N, device=dev)
    a.rgbd, a.color_var, a.depth_var, a.term_prob = rgbd.data_ptr(), cvar.data_ptr(), dvar.data_ptr(), term.data_ptr()
    weights = None
    if want_weights:
        weights = torch.empty(N, S, device=dev)
        a.weights = weights.data_ptr()
    fs = fs_m = ts = ts_m = None
    if want_aux[0]:
        fs, fs_m = torch.empty(N, S, device=dev), torch.empty(N, S, device=dev, dtype=torch.bool)
        a.freespace, a.freespace_mask = fs.data_ptr(), fs_m.data_ptr()
    if want_aux[1]:
        ts, ts_m = torch.empty(N, S, device=dev), torch.empty(N, S, device=dev, dtype=torch.bool)
        a.tsdf, a.tsdf_mask = ts.data_ptr(), ts_m.data_ptr()
    return a, (rgbd, cvar, dvar, term, weights, (fs, fs_m, ts, ts_m))


def composite(colors, geometries, distances, depths, *args, **kwargs):
    """``ngm_composite`` on flattened rays.  colors/geometries may alias one packed (N,S,4) tensor; arguments as
    :func:`composite_args`."""
    a, outs = composite_args(colors, geometries, distances, depths, *args, **kwargs)
    dev = distances.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ngm_composite(C.byref(a), _lib.stream_ptr(dev)))
    return outs


def quadrature(driver, sample_colors, sample_geometries, sample_distances, sample_depths, neus_isds) -> Tuple:
    """Drop-in for ``NeuralGraphMap._quadrature`` (ngm/run_mapping.py:709-799)."""
    models._no_autograd(sample_colors, sample_geometries)
    leading = tuple(sample_geometries.shape[:-1])
    S = sample_geometries.shape[-1]
    with torch.no_grad():
        col = _lib.dev_f32(sample_colors, "sample_colors").reshape(-1, S, 3)
        geo = _lib.dev_f32(sample_geometries, "sample_geometries").reshape(-1, S)
        dist = _lib.dev_f32(sample_distances, "sample_distances").expand(*leading, S).reshape(-1, S).contiguous()
        dep = _lib.dev_f32(sample_depths, "sample_depths").expand(*leading, S).reshape(-1, S).contiguous()
        N = geo.shape[0]
        isd, rays_per = None, 1
        mode = driver._geometry_mode
        if mode == "neus":
            # neus_isds broadcasts against (..., S); supported: one value per leading index 0 (:641-644)
            isd = _lib.dev_f32(neus_isds, "neus_isds").reshape(-1)
            if isd.numel() == 1:
                rays_per = max(N, 1)
            elif len(leading) >= 1 and isd.numel() == leading[0]:
                rays_per = N // leading[0]
            elif isd.numel() == N:
                rays_per = 1
            else:
                raise NotImplementedError("neus_isds must be scalar, per leading row, or per ray")
        rgbd, cvar, dvar, term, weights, _ = composite(
            col, geo, dist, dep, mode, driver._geometry_factor, 1.0, neus_isd=isd, rays_per_isd=rays_per,
            want_weights=True)
        if mode in ("density", "neus"):
            weights = weights[:, :-1]  # last sample dropped (last_index=-1, :749,757)
    return (rgbd[:, :3].reshape(*leading, 3), rgbd[:, 3].reshape(leading), cvar.reshape(*leading, 3),
            dvar.reshape(leading), term.reshape(leading), weights.reshape(*leading, -1))


# render_image pixel blocks: the reference cuts the image into `pixel_block_size` (8,192) pixel blocks to bound the
# memory of its PyTorch intermediates (ngm/run_mapping.py:424-435).  Pixels are independent, so the block size only
# changes memory and launch count (a 640x480x64 frame: 11.1 ms as one block, 12.9 ms as 38 blocks, tools/bench_knn.py).
# The CUDA path therefore renders blocks as large as a transient-memory budget allows -- never smaller than the
# configured `pixel_block_size` -- counting samples, not pixels: the kNN path holds ~320 B per SAMPLE in flight
# (world point, distance, depth, K = 2 fp16 feature rows, bucketing tables, raw outputs), so with the reference's
# default eval setting of 640 samples per ray a block is ~40k pixels, with 64 samples a whole 640x480 frame.
IMAGE_BLOCK_BYTES = 8 << 30
IMAGE_BYTES_PER_SAMPLE = 320


def image_block_pixels(driver) -> int:
    samples = max(int(driver._num_samples), 1)
    return max(int(driver._pixel_block_size), IMAGE_BLOCK_BYTES // (samples * IMAGE_BYTES_PER_SAMPLE))


@torch.no_grad()
def render_image(driver, c2w: torch.Tensor, camera, progressbar: bool = False):
    """Drop-in for ``NeuralGraphMap.render_image`` (ngm/run_mapping.py:402-437)."""
    h, w = camera.height, camera.width
    dev = driver._device
    ijs = torch.cartesian_prod(torch.arange(h, device=dev), torch.arange(w, device=dev))
    rgbds, _, d_vars, _, _, _ = batched_evaluation(
        lambda x: driver._render_ijs(x, c2w, camera), ijs, block_size=image_block_pixels(driver),
        progressbar=progressbar)
    return rgbds.reshape(h, w, 4), d_vars.reshape(h, w)


def install(driver_cls, optimizer: bool = True, targets: bool = True) -> None:
    """Patch the reference's ``NeuralGraphMap`` so its renderer runs on libngm_b200 while the
    SLAM driver, keyframe selection and loop-closure code stay untouched callers.  With ``optimizer`` the Adam
    update of the active fields (``_set_vmap_fields`` / ``_update_step``) runs as one in-place launch too;
    ``_init_optimizer``, ``_add_fields`` and ``_optim_state`` stay the driver's own.  With ``targets`` the
    multi-view target sampling right before the render (``_sample_target_mv``) runs as two launches and
    ``_get_observed_fields`` as one."""
    driver_cls._render_ijs = render_rays
    driver_cls._quadrature = quadrature
    driver_cls.render_image = render_image
    if optimizer:
        from . import optim

        driver_cls._set_vmap_fields = optim.set_vmap_fields
        driver_cls._update_step = optim.update_step
    if targets:
        from .targets import get_observed_fields, sample_target_mv

        driver_cls._sample_target_mv = sample_target_mv
        driver_cls._get_observed_fields = get_observed_fields


class RenderState:
    """Stand-alone holder of exactly the driver state the renderer reads, built from the
    reference's own config keys (ngm/run_mapping.py:116-215, ngm/config/neural_graph_map.yaml)."""

    # False: ``_render_ijs`` reads ``all_fields_params`` in place through field slots (no gather copy per call).
    # True: the reference driver's flow -- every ``_render_ijs`` first calls ``_set_vmap_fields`` (:500) and runs
    # on the gathered ``vmap_fields_params``, which ``_update_step`` then optimises.
    _reference_flow = False

    def __init__(self, config: dict) -> None:
        self._config = config
        self._device = config.get("device", "cuda")
        mk = dict(config["model_kwargs"])
        model_type = config.get("model_type", "neural_graph_mapping_b200.models.NeuralFieldSet")
        model_cls = str_to_object(model_type) if isinstance(model_type, str) else model_type
        self._model = model_cls(**mk).to(self._device)
        self._freespace_weight = config.get("freespace_weight", 0.0)
        self._tsdf_weight = config.get("tsdf_weight", 0.0)
        self._geometry_mode = config["geometry_mode"]
        self._geometry_factor = config.get("geometry_factor", 1.0)
        self._color_factor = config.get("color_factor", 1.0)
        self._truncation_distance = config.get("truncation_distance", None)
        self._field_radius = config.get("field_radius", mk.get("field_radius"))
        self._block_size = config.get("block_size", 3_000_000)
        self._pixel_block_size = config.get("pixel_block_size", 8192)
        self._num_samples_depth_guided = config.get("num_samples_depth_guided", 0)
        self._range_depth_guided = config.get("range_depth_guided", None)
        if self._range_depth_guided is None:
            self._range_depth_guided = self._truncation_distance
        self._single_field_id = config.get("single_field_id", None)
        self._precision = config.get("precision", None)
        self._train_near_distance = config.get("near_distance", 0.0)
        self._train_far_distance = config.get("far_distance", 8.0)
        self._train_num_samples = config.get("num_samples_coarse", 8)
        self._eval_near_distance = config.get("eval_near_distance", 0.0)
        self._eval_far_distance = config.get("eval_far_distance", 8.0)
        self._eval_num_samples = config.get("eval_num_samples", None)
        if self._eval_num_samples is None:  # run_mapping.py:199-207
            if self._num_samples_depth_guided > 0:
                spacing = 2 * self._range_depth_guided / self._num_samples_depth_guided
            else:
                spacing = 2 * self._field_radius / self._train_num_samples
            self._eval_num_samples = int((self._eval_far_distance - self._eval_near_distance) / spacing)
        self._global_map_dict = {  # run_mapping.py:231-246
            "positions": torch.zeros(32, 3, device=self._device),
            "orientations": torch.zeros(32, 4, device=self._device),
            "training_iterations": torch.zeros(32, device=self._device, dtype=torch.long),
            "num": 0,
        }
        self._learning_rate = config.get("learning_rate", 1e-3)  # run_mapping.py:347-362
        self._adam_eps = config.get("adam_eps", 1e-8)
        self._adam_weight_decay = config.get("adam_weight_decay", 0.0)
        self._optim_state = None
        self.train()

    def eval(self) -> None:  # run_mapping.py:1966-1969
        self._far_distance = self._eval_far_distance
        self._near_distance = self._eval_near_distance
        self._num_samples = self._eval_num_samples

    def train(self) -> None:  # run_mapping.py:1971-1974
        self._far_distance = self._train_far_distance
        self._near_distance = self._train_near_distance
        self._num_samples = self._train_num_samples

    def set_fields(self, all_fields_params: dict, positions: torch.Tensor, orientations: torch.Tensor) -> None:
        """Load a stacked parameter dict (checkpoint layout, run_mapping.py:2152) + field poses."""
        dev = self._device
        self._model.all_fields_params = {k: v.to(dev) for k, v in all_fields_params.items()}
        self._global_map_dict["positions"] = positions.to(dev).float().contiguous()
        self._global_map_dict["orientations"] = orientations.to(dev).float().contiguous()
        self._global_map_dict["num"] = positions.shape[0]
        self._global_map_dict["training_iterations"] = torch.zeros(positions.shape[0], device=dev, dtype=torch.long)

    def save_model(self, path: str) -> None:
        """The checkpoint half of ``NeuralGraphMap.save_model`` (ngm/run_mapping.py:2147-2156): same three keys,
        same tensors (stacked per-field parameters in the reference's state-dict names, the over-allocated map
        tables with their ``num``), so either side loads the other's file.  Verified with a checkpoint written by the
        reference for the in-tree encodings (tests/test_checkpoint.py); for the permutohedral encoding the key names
        and tensor kinds of the third-party module (``_encoding.*``) are assumed, not verified -- its source is not
        in the reference tree (parity unpinned)."""
        torch.save({"map_dict": self._global_map_dict, "all_fields_params": self._model.all_fields_params,
                    "state_dict": self._model.state_dict()}, path)

    def load_model(self, path: str) -> None:
        """``NeuralGraphMap.load_model`` (ngm/run_mapping.py:2166-2173).  The map tables may be over-allocated
        (32 rows doubled on demand, :231-262); ``num`` rows are valid and the renderer reads only those."""
        d = torch.load(path, map_location=self._device)
        self._global_map_dict = d["map_dict"]
        self._model.all_fields_params = d["all_fields_params"]
        self._model.load_state_dict(d["state_dict"])

    def _sample_target_mv(self, current_field_ids: torch.Tensor):
        """Needs the driver's keyframe store on this object: ``_camera``, ``_c_c2w_tensor``, ``_nc_rgbd_tensor``,
        ``_frame_cid_to_ncid``, ``_num_train_fields``, ``_num_rays_per_field`` (run_mapping.py:140-141, 1674-1713)."""
        from .targets import sample_target_mv

        return sample_target_mv(self, current_field_ids)

    def _set_vmap_fields(self, field_ids: torch.Tensor) -> None:
        from . import optim

        optim.set_vmap_fields(self, field_ids)

    def _update_step(self, loss_dict: dict, field_ids: torch.Tensor) -> None:
        from . import optim

        optim.update_step(self, loss_dict, field_ids)

    _render_ijs = render_rays
    _quadrature = quadrature
    render_image = render_image
