"""ctypes binding of ``libngm_b200.so`` (C ABI in ``include/ngm_b200.h``).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module
without the built shared object raises ``ImportError`` (build it with
``python -m neural_graph_mapping_b200._build`` or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# NGM_B200_LIB: load another build of the same library (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("NGM_B200_LIB") or os.path.join(_PKG, "libngm_b200.so")

NGM_ABI_VERSION = 1
NGM_MAX_LINEARS = 9

# enums (ngm_b200.h)
ENC = {"nerf": 0, "fourier": 1, "triplane": 2, "permuto": 3}
SKIP = {"no": 0, "add": 1, "concat": 2, "rezero": 3}
SCALE = {"no": 0, "unit_ball": 1, "unit_cube": 2}
GEOM = {"density": 0, "occupancy": 1, "neus": 2, "nrgbd": 3}
PREC = {"fp32": 0, "fp16": 1}
TRIPLANE = {"sum": 0, "product": 1, "concat": 2}

_fp = C.c_void_p  # device pointers travel as integers


class NgmCamera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx0", C.c_float), ("cy0", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32)]


class NgmFieldDesc(C.Structure):
    _fields_ = [
        ("encoding", C.c_int32), ("dim_encoding", C.c_int32), ("num_layers", C.c_int32),
        ("dim_mlp_out", C.c_int32), ("dim_out", C.c_int32), ("skip_mode", C.c_int32),
        ("nerf_num_octaves", C.c_int32), ("nerf_start_octave", C.c_int32),
        ("fourier_num_features", C.c_int32), ("fourier_raw_coords", C.c_int32),
        ("triplane_resolution", C.c_int32), ("triplane_components", C.c_int32), ("triplane_mode", C.c_int32),
        ("permuto_levels", C.c_int32), ("permuto_feats", C.c_int32), ("permuto_log2_capacity", C.c_int32),
        ("permuto_concat_points", C.c_int32), ("permuto_concat_scaling", C.c_float), ("_pad0", C.c_int32),
        ("weights", _fp * NGM_MAX_LINEARS), ("weight_stride", C.c_int64 * NGM_MAX_LINEARS),
        ("biases", _fp * NGM_MAX_LINEARS), ("bias_stride", C.c_int64 * NGM_MAX_LINEARS),
        ("rezero", _fp), ("rezero_stride", C.c_int64),
        ("enc_param0", _fp), ("enc_param0_stride", C.c_int64),
        ("enc_param1", _fp), ("enc_param1_stride", C.c_int64),
        ("permuto_scale", _fp), ("packed_weights", _fp),
    ]


class NgmSampleArgs(C.Structure):
    _fields_ = [
        ("cam", NgmCamera), ("num_rays", C.c_int64), ("ijs", _fp), ("c2ws", _fp), ("near", _fp), ("far", _fp),
        ("gt", _fp), ("jitter", _fp), ("jitter_guided", _fp), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("near_scalar", C.c_float), ("far_scalar", C.c_float), ("range_guided", C.c_float),
        ("c2w_per_ray", C.c_int32), ("num_samples", C.c_int32), ("num_samples_guided", C.c_int32),
        ("points_cam", _fp), ("points_world", _fp), ("distances", _fp), ("depths", _fp),
    ]


class NgmFieldFwdArgs(C.Structure):
    _fields_ = [
        ("field", NgmFieldDesc), ("points_per_field", C.c_int64), ("points", _fp), ("positions", _fp),
        ("orientations", _fp), ("field_slots", _fp), ("out", _fp), ("workspace", _fp),
        ("workspace_bytes", C.c_size_t), ("field_radius", C.c_float), ("num_fields", C.c_int32),
        ("scale_mode", C.c_int32), ("precision", C.c_int32), ("rows_half", _fp),
    ]


class NgmFieldBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", NgmFieldFwdArgs), ("d_out", _fp), ("d_weights", _fp * NGM_MAX_LINEARS),
        ("d_biases", _fp * NGM_MAX_LINEARS), ("d_encoding", _fp),
    ]


class NgmCompositeArgs(C.Structure):
    _fields_ = [
        ("num_rays", C.c_int64), ("colors", _fp), ("geometries", _fp), ("distances", _fp), ("depths", _fp),
        ("neus_isd", _fp), ("gt", _fp), ("color_stride", C.c_int64), ("geometry_stride", C.c_int64),
        ("rays_per_isd", C.c_int64), ("num_samples", C.c_int32), ("geometry_mode", C.c_int32),
        ("geometry_factor", C.c_float), ("color_factor", C.c_float), ("truncation", C.c_float),
        ("overwrite_behind_camera", C.c_int32), ("overwrite_gate", _fp),
        ("rgbd", _fp), ("color_var", _fp), ("depth_var", _fp), ("term_prob", _fp), ("weights", _fp),
        ("freespace", _fp), ("freespace_mask", _fp), ("tsdf", _fp), ("tsdf_mask", _fp),
        ("mirror_delta", C.c_int64 * 8), ("num_mirrors", C.c_int32), ("_pad_m", C.c_int32),
    ]


class NgmCompositeBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", NgmCompositeArgs), ("g_rgbd", _fp), ("g_color_var", _fp), ("g_depth_var", _fp), ("g_term_prob", _fp),
        ("g_freespace", _fp), ("g_tsdf", _fp), ("d_colors", _fp), ("d_geometries", _fp), ("d_neus_isd", _fp),
        ("workspace", _fp),
    ]


class NgmEncodeArgs(C.Structure):
    _fields_ = [
        ("field", NgmFieldDesc), ("points_per_field", C.c_int64), ("points", _fp), ("field_slots", _fp), ("out", _fp),
        ("d_out", _fp), ("d_param0", _fp), ("num_fields", C.c_int32), ("_pad", C.c_int32),
    ]


class NgmAdamParam(C.Structure):
    _fields_ = [("param_all", _fp), ("exp_avg_all", _fp), ("exp_avg_sq_all", _fp), ("grad", _fp),
                ("param_active", _fp), ("row", C.c_int64)]


class NgmAdamArgs(C.Structure):
    _fields_ = [("params", C.POINTER(NgmAdamParam)), ("field_ids", _fp), ("num_active", C.c_int64),
                ("step", C.c_int64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("weight_decay", C.c_double), ("num_params", C.c_int32), ("_pad", C.c_int32)]


NGM_ADAM_MAX_PARAMS = 24


class NgmTargetVisArgs(C.Structure):
    _fields_ = [("cam", NgmCamera), ("c2ws", _fp), ("rgbds", _fp), ("frame_to_store", _fp), ("positions", _fp),
                ("field_ids", _fp), ("probe_offsets", _fp), ("num_frames", C.c_int64), ("num_fields", C.c_int32),
                ("num_probes", C.c_int32), ("train_radius", C.c_float), ("_pad", C.c_int32),
                ("field_kf_mask", _fp), ("min_xys", _fp), ("max_xys", _fp)]


class NgmObservedArgs(C.Structure):
    _fields_ = [("cam", NgmCamera), ("depth", _fp), ("pixel_ids", _fp), ("c2w", _fp), ("positions", _fp),
                ("pixel_stride", C.c_int64), ("num_points", C.c_int32), ("num_fields", C.c_int32),
                ("field_radius", C.c_float), ("_pad", C.c_int32), ("observed", _fp)]


class NgmTargetRaysArgs(C.Structure):
    _fields_ = [("cam", NgmCamera), ("c2ws", _fp), ("rgbds", _fp), ("frame_to_store", _fp), ("positions", _fp),
                ("field_ids", _fp), ("frame_cids", _fp), ("uv", _fp), ("min_xys", _fp), ("max_xys", _fp),
                ("num_frames", C.c_int64), ("rays_per_field", C.c_int64), ("num_fields", C.c_int32),
                ("train_radius", C.c_float), ("ijs", _fp), ("out_c2ws", _fp), ("near", _fp), ("far", _fp), ("gt", _fp),
                ("out_rgbds", _fp), ("rgb_mask", _fp), ("depth_mask", _fp), ("term_probs", _fp), ("term_mask", _fp)]


class NgmRenderArgs(C.Structure):
    _fields_ = [
        ("field", NgmFieldDesc), ("cam", NgmCamera), ("rays_per_field", C.c_int64), ("ijs", _fp), ("c2ws", _fp),
        ("near", _fp), ("far", _fp), ("gt", _fp), ("jitter", _fp), ("jitter_guided", _fp), ("positions", _fp),
        ("orientations", _fp), ("field_slots", _fp), ("neus_sd", _fp), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("near_scalar", C.c_float), ("far_scalar", C.c_float), ("range_guided", C.c_float),
        ("field_radius", C.c_float), ("geometry_factor", C.c_float), ("color_factor", C.c_float),
        ("truncation", C.c_float), ("c2w_per_ray", C.c_int32), ("num_samples", C.c_int32),
        ("num_samples_guided", C.c_int32), ("num_fields", C.c_int32), ("scale_mode", C.c_int32),
        ("geometry_mode", C.c_int32), ("precision", C.c_int32), ("overwrite_behind_camera", C.c_int32),
        ("_pad1", C.c_int32), ("overwrite_gate", _fp),
        ("rgbd", _fp), ("color_var", _fp), ("depth_var", _fp), ("term_prob", _fp),
        ("freespace", _fp), ("freespace_mask", _fp), ("tsdf", _fp), ("tsdf_mask", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_size_t),
        ("mirror_delta", C.c_int64 * 8), ("num_mirrors", C.c_int32), ("_pad2", C.c_int32),
    ]


class NgmKnnFwdArgs(C.Structure):
    _fields_ = [
        ("field", NgmFieldDesc), ("num_points", C.c_int64), ("points", _fp), ("positions", _fp), ("orientations", _fp),
        ("field_slots", _fp), ("out", _fp), ("workspace", _fp), ("workspace_bytes", C.c_size_t),
        ("field_radius", C.c_float), ("scale_radius", C.c_float), ("distance_factor", C.c_float),
        ("outside_value", C.c_float), ("num_fields", C.c_int32), ("num_knn", C.c_int32), ("scale_mode", C.c_int32),
        ("precision", C.c_int32),
    ]


STRUCTS = [NgmCamera, NgmFieldDesc, NgmSampleArgs, NgmFieldFwdArgs, NgmCompositeArgs, NgmRenderArgs, NgmKnnFwdArgs,
           NgmCompositeBwdArgs, NgmEncodeArgs, NgmAdamParam, NgmAdamArgs,
           NgmTargetVisArgs, NgmTargetRaysArgs, NgmObservedArgs, NgmFieldBwdArgs]
EXPORTS = [
    "ngm_abi_version", "ngm_last_error", "ngm_struct_size", "ngm_launch_count", "ngm_sample_rays", "ngm_field_fwd", "ngm_field_bwd", "ngm_field_bwd_workspace_bytes", "ngm_composite", "ngm_composite_bwd", "ngm_encode_fwd", "ngm_encode_bwd", "ngm_adam_step", "ngm_target_visibility", "ngm_target_rays", "ngm_observed_fields",
    "ngm_render_rays_fwd", "ngm_fieldset_knn_fwd", "ngm_fieldset_knn_workspace_bytes", "ngm_field_fwd_workspace_bytes", "ngm_render_workspace_bytes", "ngm_packed_weights_bytes", "ngm_pack_weights",
]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA library is the product and there is no fallback. "
        "Build it with `python -m neural_graph_mapping_b200._build`."
    )

lib = C.CDLL(LIB_PATH)
lib.ngm_abi_version.restype = C.c_int
lib.ngm_last_error.restype = C.c_char_p
lib.ngm_struct_size.restype = C.c_size_t
lib.ngm_struct_size.argtypes = [C.c_int]
lib.ngm_launch_count.restype = C.c_uint64
for _name, _arg in [("ngm_sample_rays", NgmSampleArgs), ("ngm_field_fwd", NgmFieldFwdArgs),
                    ("ngm_field_bwd", NgmFieldBwdArgs),
                    ("ngm_composite", NgmCompositeArgs), ("ngm_render_rays_fwd", NgmRenderArgs),
                    ("ngm_fieldset_knn_fwd", NgmKnnFwdArgs), ("ngm_composite_bwd", NgmCompositeBwdArgs),
                    ("ngm_encode_fwd", NgmEncodeArgs), ("ngm_encode_bwd", NgmEncodeArgs), ("ngm_adam_step", NgmAdamArgs),
                    ("ngm_target_visibility", NgmTargetVisArgs), ("ngm_target_rays", NgmTargetRaysArgs),
                    ("ngm_observed_fields", NgmObservedArgs)]:
    getattr(lib, _name).restype = C.c_int
    getattr(lib, _name).argtypes = [C.POINTER(_arg), C.c_void_p]
lib.ngm_fieldset_knn_workspace_bytes.restype = C.c_int
lib.ngm_fieldset_knn_workspace_bytes.argtypes = [C.POINTER(NgmKnnFwdArgs), C.POINTER(C.c_size_t)]
lib.ngm_field_bwd_workspace_bytes.restype = C.c_int
lib.ngm_field_bwd_workspace_bytes.argtypes = [C.POINTER(NgmFieldBwdArgs), C.POINTER(C.c_size_t)]
lib.ngm_packed_weights_bytes.restype = C.c_int
lib.ngm_packed_weights_bytes.argtypes = [C.POINTER(NgmFieldDesc), C.POINTER(C.c_size_t)]
lib.ngm_pack_weights.restype = C.c_int
lib.ngm_pack_weights.argtypes = [C.POINTER(NgmFieldDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
lib.ngm_field_fwd_workspace_bytes.restype = C.c_int
lib.ngm_field_fwd_workspace_bytes.argtypes = [C.POINTER(NgmFieldFwdArgs), C.POINTER(C.c_size_t)]
lib.ngm_render_workspace_bytes.restype = C.c_int
lib.ngm_render_workspace_bytes.argtypes = [C.POINTER(NgmRenderArgs), C.POINTER(C.c_size_t)]

if lib.ngm_abi_version() != NGM_ABI_VERSION:
    raise ImportError(f"libngm_b200 ABI {lib.ngm_abi_version()} != binding ABI {NGM_ABI_VERSION}; rebuild")
for _i, _s in enumerate(STRUCTS):
    if lib.ngm_struct_size(_i) != C.sizeof(_s):
        raise ImportError(f"{_s.__name__}: C sizeof {lib.ngm_struct_size(_i)} != ctypes {C.sizeof(_s)}")


def load_debug_lib():
    """The diagnostics build (libngm_b200_debug.so, include/ngm_b200_debug.h): tests and tools only.  The package
    itself never loads it."""
    path = os.path.join(_PKG, "libngm_b200_debug.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m neural_graph_mapping_b200._build --debug`")
    dbg = C.CDLL(path)
    dbg.ngm_last_error.restype = C.c_char_p
    dbg.ngm_debug_tc_trace_peek.restype = C.c_int
    dbg.ngm_debug_tc_trace_peek.argtypes = [C.c_void_p, C.c_int]
    dbg.ngm_debug_tc_trace.restype = C.c_int
    dbg.ngm_debug_tc_trace.argtypes = [C.c_void_p, C.c_int]
    dbg.ngm_debug_bwd_phases.restype = C.c_int
    dbg.ngm_debug_bwd_phases.argtypes = [C.c_void_p]
    dbg.ngm_debug_tc_gemm.restype = C.c_int
    dbg.ngm_debug_tc_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]
    return dbg


class NgmError(RuntimeError):
    pass


def check(rc: int) -> None:
    """Map an NgmStatus to the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib.ngm_last_error().decode(errors="replace")
    if rc == -1:
        raise ValueError(msg)  # reference raises ValueError (run_mapping.py:498, models.py:110,219)
    if rc == -2:
        raise NotImplementedError(msg)  # models.py:243,285
    raise NgmError(f"libngm_b200 status {rc}: {msg}")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """fp32, contiguous, on a CUDA device -- or a loud error (no CPU path)."""
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: neural_graph_mapping_b200 runs on CUDA (sm_100a) only; "
            "there is no CPU fallback."
        )
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
