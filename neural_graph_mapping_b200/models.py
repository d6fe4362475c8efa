"""``NeuralField`` / ``NeuralFieldSet`` with the reference's names, constructor arguments,
attributes and parameter-dict layout (ngm/models.py:66-411), evaluated by libngm_b200.

Drop-in selection is by YAML type strings, as in the reference (utils.str_to_object):

    model_type: neural_graph_mapping_b200.models.NeuralFieldSet
    model_kwargs.field_type: neural_graph_mapping_b200.models.NeuralField
    field_kwargs.encoding_type: neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Literal, Optional

import torch

from . import _lib
from .utils import str_to_object

_DEFAULT_PRECISION = "fp32"


def set_default_precision(precision: str) -> None:
    """"fp32" (reference arithmetic, FFMA) or "fp16" (fp16 operands, fp32 accumulate, tcgen05)."""
    global _DEFAULT_PRECISION
    if precision not in _lib.PREC:
        raise ValueError(f"precision must be one of {list(_lib.PREC)}")
    _DEFAULT_PRECISION = precision


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


def _no_autograd(*tensors) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "this entry point is forward-only: gradients flow through render_rays(use_vmap=True) "
            "(neural_graph_mapping_b200.autograd); call under torch.no_grad() or detach the parameters"
        )


class NeuralField(torch.nn.Module):
    """Positional encoding + MLP (ngm/models.py:66-182); forward runs ``ngm_field_fwd``."""

    def __init__(
        self,
        encoding_type: str,
        encoding_kwargs: dict,
        num_layers: int,
        dim_out: int,
        dim_mlp_out: Optional[int] = None,
        skip_mode: Literal["no", "add", "concat", "rezero"] = "no",
        initial_geometry_bias: float = 0.0,
        neus_initial_sd: Optional[float] = None,
        precision: Optional[str] = None,
    ) -> None:
        super().__init__()
        self._encoding_type = str_to_object(encoding_type) if isinstance(encoding_type, str) else encoding_type
        if self._encoding_type is None:
            raise ValueError(f"Unknown encoding type {encoding_type}.")
        self._encoding_kwargs = encoding_kwargs
        self._encoding = self._encoding_type(**self._encoding_kwargs)
        self._dim_encoding = self._encoding.get_out_dim()
        self._dim_out = dim_out
        self._dim_mlp_out = dim_mlp_out if dim_mlp_out is not None else self._dim_encoding
        self._skip_mode = skip_mode
        self._initial_geometry_bias = initial_geometry_bias
        self._num_layers = num_layers
        self.precision = precision

        if self._skip_mode in ["no", "add", "rezero"]:
            self._dim_mlp_in = self._dim_mlp_out
        elif self._skip_mode == "concat":
            self._dim_mlp_in = self._dim_mlp_out + self._dim_encoding
        else:
            raise ValueError(f"Skip mode {self._skip_mode} is not available.")  # models.py:110
        if num_layers + 1 > _lib.NGM_MAX_LINEARS:
            raise NotImplementedError(f"num_layers must be <= {_lib.NGM_MAX_LINEARS - 1}")

        if self._skip_mode == "rezero":
            self._rezero = torch.nn.Parameter(torch.zeros(self._num_layers))

        self._dims_in = [self._dim_encoding] + [self._dim_mlp_in for _ in range(self._num_layers)]
        self._dims_out = [self._dim_mlp_out for _ in range(self._num_layers)]
        self._dims_out.append(self._dim_out)

        if neus_initial_sd is not None:
            self._neus_sd = torch.nn.Parameter(torch.tensor(neus_initial_sd))

        self._linears = torch.nn.ModuleList()
        for d_in, d_out in zip(self._dims_in, self._dims_out):
            self._linears.append(torch.nn.Linear(d_in, d_out))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            if self._skip_mode == "rezero":
                self._rezero.zero_()
            self._linears[-1].bias[-1] += self._initial_geometry_bias  # models.py:135

    def numel(self) -> int:
        return sum(p.numel() for p in self.parameters())

    # ---- descriptor ---------------------------------------------------------------------
    def field_desc(self, params: Dict[str, torch.Tensor], stacked: bool, packed: Optional[torch.Tensor] = None):
        """Fill an ``NgmFieldDesc`` from a parameter dict with the reference's state_dict names.
        ``stacked``: tensors carry a leading field dimension (all_fields_params layout).  ``packed``: persistent
        pre-swizzled fp16 images of the same tables (``NeuralFieldSet.packed_images``).
        Returns (desc, keepalive) -- keepalive holds the contiguous tensors the desc points to."""
        d = _lib.NgmFieldDesc()
        keep = []
        if packed is not None:
            keep.append(packed)
            d.packed_weights = packed.data_ptr()
        enc = self._encoding
        d.encoding = _lib.ENC[enc.KIND]
        d.dim_encoding = self._dim_encoding
        d.num_layers = self._num_layers
        d.dim_mlp_out = self._dim_mlp_out
        d.dim_out = self._dim_out
        d.skip_mode = _lib.SKIP[self._skip_mode]

        def table(name):
            t = _lib.dev_f32(params[name], name)
            keep.append(t)
            per_field = t[0].numel() if stacked else 0
            return t.data_ptr(), per_field

        for i in range(self._num_layers + 1):
            d.weights[i], d.weight_stride[i] = table(f"_linears.{i}.weight")
            d.biases[i], d.bias_stride[i] = table(f"_linears.{i}.bias")
            w = params[f"_linears.{i}.weight"]
            if tuple(w.shape[-2:]) != (self._dims_out[i], self._dims_in[i]):
                raise ValueError(f"_linears.{i}.weight has shape {tuple(w.shape)}")
        if self._skip_mode == "rezero":
            d.rezero, d.rezero_stride = table("_rezero")
        if enc.KIND == "nerf":
            d.nerf_num_octaves, d.nerf_start_octave = enc.num_octaves, enc.start_octave
        elif enc.KIND == "fourier":
            d.fourier_num_features = enc._linear.out_features
            d.fourier_raw_coords = int(enc._raw_coords)
            d.enc_param0, d.enc_param0_stride = table("_encoding._linear.weight")
        elif enc.KIND == "triplane":
            d.triplane_resolution, d.triplane_components = enc.resolution, enc.num_components
            d.triplane_mode = _lib.TRIPLANE[enc.mode]
            d.enc_param0, d.enc_param0_stride = table("_encoding.plane_coef")
        elif enc.KIND == "permuto":
            d.permuto_levels, d.permuto_feats = enc.nr_levels, enc.nr_feat_per_level
            d.permuto_log2_capacity = enc.log2_hashmap_size
            d.permuto_concat_points = int(enc.concat_points)
            d.permuto_concat_scaling = float(enc.concat_points_scaling)
            d.enc_param0, d.enc_param0_stride = table("_encoding.lattice_values")
            d.enc_param1, d.enc_param1_stride = table("_encoding.random_shift_per_level")
            sf = _lib.dev_f32(enc.scale_factor.to(keep[0].device), "scale_factor")
            keep.append(sf)
            d.permuto_scale = sf.data_ptr()
        return d, keep

    def _own_params(self) -> Dict[str, torch.Tensor]:
        p = dict(self.named_parameters())
        p.update(dict(self.named_buffers()))
        return p

    def forward(self, query_points: torch.Tensor) -> torch.Tensor:
        """(..., 3) local points -> (..., dim_out)  (ngm/models.py:143-182)."""
        if not query_points.is_cuda:
            raise RuntimeError("query_points is on the CPU: neural_graph_mapping_b200 runs on CUDA "
                               "(sm_100a) only; there is no CPU fallback.")
        params = self._own_params()
        _no_autograd(query_points, *params.values())
        with torch.no_grad():
            return field_forward(self, params, False, query_points.reshape(1, -1, 3), None, None, None,
                                 "no", None, self.precision).reshape(*query_points.shape[:-1], self._dim_out)


def field_forward(proto: NeuralField, params, stacked: bool, points: torch.Tensor, positions, orientations,
                  field_slots, scale_mode: str, field_radius, precision: Optional[str]) -> torch.Tensor:
    """``ngm_field_fwd`` on (F, N, 3) points -> (F, N, dim_out)."""
    pts = _lib.dev_f32(points, "query_points")
    dev = pts.device
    F, N = pts.shape[0], pts.shape[1]
    a = _lib.NgmFieldFwdArgs()
    with torch.cuda.device(dev):
        a.field, keep = proto.field_desc(params, stacked)
        out = torch.empty(F, N, proto._dim_out, device=dev, dtype=torch.float32)
        a.points_per_field = N
        a.num_fields = F
        a.points = pts.data_ptr()
        if positions is not None:
            pos, ori = _lib.dev_f32(positions, "field_positions"), _lib.dev_f32(orientations, "field_orientations")
            keep += [pos, ori]
            a.positions, a.orientations = pos.data_ptr(), ori.data_ptr()
        if field_slots is not None:
            slots = field_slots.to(device=dev, dtype=torch.int64).contiguous()
            keep.append(slots)
            a.field_slots = slots.data_ptr()
        a.out = out.data_ptr()
        a.scale_mode = _lib.SCALE[scale_mode]
        a.field_radius = float(field_radius) if field_radius is not None else 0.0
        a.precision = _lib.PREC[precision or _DEFAULT_PRECISION]
        need = C.c_size_t(0)
        _lib.check(_lib.lib.ngm_field_fwd_workspace_bytes(C.byref(a), C.byref(need)))
        if need.value:
            ws = torch.empty(need.value, device=dev, dtype=torch.uint8)
            keep.append(ws)
            a.workspace, a.workspace_bytes = ws.data_ptr(), need.value
        _lib.check(_lib.lib.ngm_field_fwd(C.byref(a), _lib.stream_ptr(dev)))
    return out


class NeuralFieldSet(torch.nn.Module):
    """Set of posed neural fields (ngm/models.py:185-411)."""

    def __init__(
        self,
        dim_points: int,
        field_type: str,
        field_kwargs: dict,
        num_knn: int,
        distance_factor: float,
        outside_value: float,
        field_radius: Optional[float] = None,
        scale_mode: Literal["no", "unit_ball", "unit_cube"] = "no",
        precision: Optional[str] = None,
    ) -> None:
        super().__init__()
        self._scale_mode = scale_mode
        self._field_radius = field_radius
        if scale_mode != "no" and field_radius is None:
            raise ValueError(f"{scale_mode=} requires field_radius to be specified.")  # models.py:219
        if scale_mode not in _lib.SCALE:
            raise NotImplementedError(f"{scale_mode=} is not available.")  # models.py:285
        if dim_points != 3:
            raise NotImplementedError("Only 3D spaces are supported by the CUDA path.")  # models.py:243
        self._dim_points = dim_points
        self._num_knn = num_knn
        self._distance_factor = distance_factor
        self._outside_value = outside_value
        self.precision = precision
        ft = str_to_object(field_type) if isinstance(field_type, str) else field_type
        if ft is None:
            raise ValueError(f"Unknown field type {field_type}.")
        self._prototype_field = ft(**field_kwargs)
        self.all_fields_params = None
        self.vmap_fields_params = None

    # ---- persistent kernel-friendly weight layout ------------------------------------------------
    def _linear_names(self):
        n = self._prototype_field._num_layers + 1
        return [f"_linears.{i}.{w}" for i in range(n) for w in ("weight", "bias")]

    def _packed_key(self, params):
        return tuple((k, params[k].data_ptr(), params[k]._version, tuple(params[k].shape)) for k in self._linear_names())

    def packed_images(self, params: Dict[str, torch.Tensor]) -> Optional[torch.Tensor]:
        """Pre-swizzled fp16 weight images (one per table row) of ``params`` for the tensor-core kernels, cached
        while the tables are unchanged: the cache key is every linear's storage pointer, torch version counter and
        shape, so in-place updates by torch ops (the reference driver's scatter-back, run_mapping.py:1204) and new
        tables (``add_fields``, ``load_model``) invalidate it; ``ngm_adam_step`` writes through raw pointers, so
        ``repack_rows`` re-packs exactly the rows it touched and keeps the cache valid.  None when the field has no
        tensor-core path or ``params`` is not this set's ``all_fields_params``."""
        if params is not self.all_fields_params or params is None:
            return None
        key = self._packed_key(params)
        cache = getattr(self, "_packed_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        dev = params[self._linear_names()[0]].device
        if dev.type != "cuda":
            return None
        with torch.cuda.device(dev):
            desc, keep = self._prototype_field.field_desc(params, True)
            per = C.c_size_t(0)
            rc = _lib.lib.ngm_packed_weights_bytes(C.byref(desc), C.byref(per))
            if rc == -2:  # field outside the tensor-core path: nothing to cache
                self._packed_cache = (key, None)
                return None
            _lib.check(rc)
            rows = params[self._linear_names()[0]].shape[0]
            buf = torch.empty(max(rows * per.value, 16), dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib.ngm_pack_weights(C.byref(desc), None, rows, buf.data_ptr(), _lib.stream_ptr(dev)))
        self._packed_cache = (key, buf)
        return buf

    def repack_rows(self, field_ids: torch.Tensor) -> None:
        """Re-pack the images of rows ``field_ids`` after an in-place update that torch's version counters cannot
        see (``ngm_adam_step``); no-op without a cache."""
        cache = getattr(self, "_packed_cache", None)
        params = self.all_fields_params
        if cache is None or cache[1] is None or params is None or cache[0] != self._packed_key(params):
            return
        dev = cache[1].device
        ids = field_ids.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            desc, keep = self._prototype_field.field_desc(params, True)
            _lib.check(_lib.lib.ngm_pack_weights(C.byref(desc), ids.data_ptr(), ids.numel(), cache[1].data_ptr(),
                                                 _lib.stream_ptr(dev)))

    def add_fields(self, num_fields: int) -> None:
        """Append ``num_fields`` copies of the prototype's state (ngm/models.py:245-264)."""
        new = {k: v.unsqueeze(0).repeat(num_fields, *([1] * v.dim())).clone()
               for k, v in self._prototype_field.state_dict().items()}
        if self.all_fields_params is None:
            self.all_fields_params = new
        else:
            self.all_fields_params = {k: torch.cat((v, new[k])) for k, v in self.all_fields_params.items()}

    def set_vmap_fields(self, field_ids: Optional[torch.Tensor]) -> None:
        """Active-field subset (ngm/models.py:266-276).  Kept as a gather copy because the driver
        mutates / optimises ``vmap_fields_params`` directly (run_mapping.py:679-707,1204); the
        renderer itself can instead read ``all_fields_params`` through ``field_slots``."""
        if field_ids is None:
            self.vmap_fields_params = self.all_fields_params
        else:
            self.vmap_fields_params = {k: v[field_ids] for k, v in self.all_fields_params.items()}

    def forward(
        self,
        query_points: torch.Tensor,
        field_positions: Optional[torch.Tensor] = None,
        field_orientations: Optional[torch.Tensor] = None,
        field_ids: Optional[torch.Tensor] = None,
        use_vmap: bool = True,
        field_radius: Optional[float] = None,
    ) -> torch.Tensor:
        """Reference semantics of ngm/models.py:287-405."""
        if field_radius is None:
            field_radius = self._field_radius
        if use_vmap:
            params = self.vmap_fields_params
            if params is None:
                raise ValueError("set_vmap_fields() must be called before a vmap forward")
            _no_autograd(query_points, *params.values())
            if (field_positions is None) != (field_orientations is None):
                raise ValueError("field_positions and field_orientations must be given together")
            with torch.no_grad():
                return field_forward(self._prototype_field, params, True, query_points, field_positions,
                                     field_orientations, None, self._scale_mode, self._field_radius,
                                     self.precision)
        from .knn import fieldset_forward_knn  # kNN blend path (models.py:347-405)

        return fieldset_forward_knn(self, query_points, field_positions, field_orientations, field_ids,
                                    field_radius)

    def numel(self) -> int:
        return self._prototype_field.numel() * len(self.all_fields_params)
