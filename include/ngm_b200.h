/*
 * ngm_b200.h -- C ABI of libngm_b200.so: the B200 (sm_100a) ray-render hot path of
 * KTH-RPL/neural_graph_mapping.
 *
 * The reference has no FFI: its boundary for this path is Python duck-typing selected by
 * YAML type strings (utils.str_to_object, ngm/utils.py:114-138).  This header is the C
 * boundary that sits directly under that Python surface; every entry point names the
 * reference function it replaces ("ngm/" = /root/reference/src/neural_graph_mapping/).
 * The Python mirror (neural_graph_mapping_b200/{camera,models,renderer}.py) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain C: pointers, sizes, scalars.  No torch / C++ types cross the boundary.
 *   - every pointer is a DEVICE pointer to caller-owned memory (fp32 unless stated,
 *     densely packed row-major); the library never allocates or frees device memory.
 *   - asynchronous on the CUDA stream passed as `void* stream` (a cudaStream_t); no host
 *     synchronisation inside; thread-safe per (device, stream).
 *   - return 0 on success, a negative NgmStatus on failure; ngm_last_error() returns a
 *     thread-local message.  No exceptions cross the ABI.  There is NO CPU fallback.
 */
#ifndef NGM_B200_H_
#define NGM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGM_ABI_VERSION 1
#define NGM_MAX_LINEARS 9 /* num_layers + 1 <= 9 */

typedef enum NgmStatus {
  NGM_OK = 0,
  NGM_ERR_INVALID_ARG = -1,  /* ValueError in the reference (run_mapping.py:498, models.py:110,219,285) */
  NGM_ERR_UNSUPPORTED = -2,  /* configuration outside what this build implements */
  NGM_ERR_CUDA = -3,         /* a CUDA runtime call failed; message has cudaGetErrorString */
  NGM_ERR_WORKSPACE = -4     /* workspace missing or too small (see ngm_*_workspace_bytes) */
} NgmStatus;

/* ngm/positional_encodings.py: PositionalEncodingNeRF :219, PositionalEncodingFourier :164,
 * TriplaneEncoding :69, PermutohedralEncoding :19 (third-party kernel; parity unpinned) */
typedef enum NgmEncoding { NGM_ENC_NERF = 0, NGM_ENC_FOURIER = 1, NGM_ENC_TRIPLANE = 2, NGM_ENC_PERMUTO = 3 } NgmEncoding;
/* ngm/models.py:104-113 skip_mode */
typedef enum NgmSkipMode { NGM_SKIP_NO = 0, NGM_SKIP_ADD = 1, NGM_SKIP_CONCAT = 2, NGM_SKIP_REZERO = 3 } NgmSkipMode;
/* ngm/models.py:278-285 scale_mode */
typedef enum NgmScaleMode { NGM_SCALE_NO = 0, NGM_SCALE_UNIT_BALL = 1, NGM_SCALE_UNIT_CUBE = 2 } NgmScaleMode;
/* ngm/run_mapping.py:746-762 geometry_mode */
typedef enum NgmGeometryMode { NGM_GEOM_DENSITY = 0, NGM_GEOM_OCCUPANCY = 1, NGM_GEOM_NEUS = 2, NGM_GEOM_NRGBD = 3 } NgmGeometryMode;
/* arithmetic of the MLP operands: fp32 FFMA path (reference arithmetic) or fp16 operands with
 * fp32 accumulation on tcgen05 tensor cores.  Geometry, sigmoid, transmittance and every
 * accumulator of the compositor are fp32 in both. */
typedef enum NgmPrecision { NGM_PREC_FP32 = 0, NGM_PREC_FP16 = 1 } NgmPrecision;
typedef enum NgmTriplaneMode { NGM_TRIPLANE_SUM = 0, NGM_TRIPLANE_PRODUCT = 1, NGM_TRIPLANE_CONCAT = 2 } NgmTriplaneMode;

/* Pinhole camera as used by Camera.ijs_to_directions (ngm/camera.py:186-203):
 * cx0/cy0 are the principal point for pixel_center 0, i.e. get_pinhole_camera_parameters(0.0)
 * (ngm/camera.py:98-116) = cx_in - pixel_center. */
typedef struct NgmCamera {
  float fx, fy, cx0, cy0;
  int32_t width, height;
} NgmCamera;

/* One NeuralField architecture (ngm/models.py:66-182) plus the stacked per-field parameter
 * tables in the reference's own layout: all_fields_params[name] = (num_fields, *shape), fp32
 * (ngm/models.py:254-264).  `*_stride` = elements between consecutive fields (normally the
 * per-field numel; 0 broadcasts one field, e.g. a single NeuralField).  Field f of a call reads
 * row field_slots[f] (see the Args structs) -- so neither set_vmap_fields' gather copy
 * (ngm/models.py:274-276) nor a weight repack is needed by the caller. */
typedef struct NgmFieldDesc {
  int32_t encoding;      /* NgmEncoding */
  int32_t dim_encoding;  /* E = encoding.get_out_dim() */
  int32_t num_layers;    /* hidden layers L; linears = L + 1 */
  int32_t dim_mlp_out;   /* W */
  int32_t dim_out;       /* 4 for the renderer (rgb + geometry) */
  int32_t skip_mode;     /* NgmSkipMode */
  /* encoding kwargs */
  int32_t nerf_num_octaves, nerf_start_octave;               /* positional_encodings.py:229 */
  int32_t fourier_num_features, fourier_raw_coords;          /* rows of _encoding._linear.weight; :187-195 */
  int32_t triplane_resolution, triplane_components, triplane_mode; /* :96-121 */
  int32_t permuto_levels, permuto_feats, permuto_log2_capacity, permuto_concat_points; /* :22-62 */
  float permuto_concat_scaling;
  int32_t _pad0;
  /* _linears.{i}.weight (out,in) and .bias (out,), i = 0..num_layers */
  const float* weights[NGM_MAX_LINEARS];
  int64_t weight_stride[NGM_MAX_LINEARS];
  const float* biases[NGM_MAX_LINEARS];
  int64_t bias_stride[NGM_MAX_LINEARS];
  const float* rezero; /* _rezero (num_layers,), skip_mode rezero only */
  int64_t rezero_stride;
  /* encoding parameters: fourier `_encoding._linear.weight` (n,3) | triplane
   * `_encoding.plane_coef` (3,C,res,res) | permuto lattice table (levels,capacity,feats) */
  const float* enc_param0;
  int64_t enc_param0_stride;
  /* permuto: per-level random shift (levels,3) */
  const float* enc_param1;
  int64_t enc_param1_stride;
  /* permuto: per-level, per-axis scale factors (levels,3), shared by all fields */
  const float* permuto_scale;
  /* optional (fp16 path): persistent pre-swizzled fp16 weight images of the table rows, written by
   * ngm_pack_weights and indexed like the tables (image of row r at packed_weights + r * ngm_packed_weights_bytes).
   * When non-NULL the tensor-core kernels read these instead of re-packing the active rows on every call; the
   * caller re-packs the rows it changes (e.g. the rows an Adam step touched). */
  const void* packed_weights;
} NgmFieldDesc;

/* ---- stage: ray sampler --------------------------------------------------------------
 * Replaces Camera.sample_ijs_uniform (ngm/camera.py:215-292, uniform branch), the
 * depth-guided second sample set + sort/gather merge (ngm/run_mapping.py:521-545) and
 * utils.transform_points (ngm/utils.py:276-286 at run_mapping.py:547).
 * Rays are the flattened leading dims of `ijs`.  St = num_samples + num_samples_guided. */
typedef struct NgmSampleArgs {
  NgmCamera cam;
  int64_t num_rays;
  const int64_t* ijs;   /* (num_rays, 2) int64 [row, col] */
  const float* c2ws;    /* (4,4) row-major, or (num_rays,4,4) when c2w_per_ray != 0; OpenGL */
  const float* near;    /* (num_rays) or NULL -> near_scalar */
  const float* far;     /* (num_rays) or NULL -> far_scalar */
  const float* gt;      /* (num_rays) or NULL; 0.0 = unavailable (run_mapping.py:522-526) */
  const float* jitter;        /* (num_rays, num_samples) U[0,1) or NULL -> counter-hash RNG(seed, offset) */
  const float* jitter_guided; /* (num_rays, num_samples_guided) or NULL -> in-kernel RNG */
  uint64_t seed, offset;
  float near_scalar, far_scalar, range_guided;
  int32_t c2w_per_ray;
  int32_t num_samples, num_samples_guided; /* guided set is used only when gt != NULL */
  /* outputs, any may be NULL: */
  float* points_cam;   /* (num_rays, St, 3) */
  float* points_world; /* (num_rays, St, 3) */
  float* distances;    /* (num_rays, St) sorted ascending */
  float* depths;       /* (num_rays, St) = -points_cam.z (run_mapping.py:612) */
} NgmSampleArgs;

/* ---- stage: field evaluation ---------------------------------------------------------
 * Replaces NeuralFieldSet.forward, vmap branch (ngm/models.py:329-345): world->local
 * (quaternion_invert/apply, :331-335), _scale_local_points (:278-285) and NeuralField.forward
 * (:143-182) for `num_fields` fields x `points_per_field` points.  With positions == NULL the
 * points are already local (NeuralField.forward / the :336-337 branch). */
typedef struct NgmFieldFwdArgs {
  NgmFieldDesc field;
  int64_t points_per_field;
  const float* points;        /* (num_fields, points_per_field, 3) */
  const float* positions;     /* (num_slots, 3) or NULL */
  const float* orientations;  /* (num_slots, 4) real-first quaternions, or NULL */
  const int64_t* field_slots; /* (num_fields) row of each field in positions/orientations/param tables; NULL = identity */
  float* out;                 /* (num_fields, points_per_field, dim_out) */
  void* workspace;
  size_t workspace_bytes;
  float field_radius;
  int32_t num_fields;
  int32_t scale_mode; /* NgmScaleMode */
  int32_t precision;  /* NgmPrecision */
  /* optional (fp16 path only): the encoding already evaluated by the caller as fp16 rows
   * (num_fields * points_per_field, EP) with EP = dim_encoding rounded up to 16 and zero padding -- the layer-0
   * operand of the MLP.  `points` / poses are then not read.  NULL: the library evaluates the encoding. */
  const void* rows_half;
} NgmFieldFwdArgs;

/* ---- stage: field evaluation, backward ------------------------------------------------
 * What torch.autograd derives from NeuralField.forward under vmap (ngm/models.py:143-182, :342) when the
 * training step calls loss.backward() (ngm/run_mapping.py:1186): given d_out = dLoss/d out of the forward call
 * `fwd`, the gradients of every `_linears.{i}.weight` / `.bias` of the num_fields evaluated fields, in the layout
 * of the gathered tables (vmap_fields_params: (num_fields, out, in) / (num_fields, out)).  The forward is
 * recomputed inside the kernel; nothing is stored by ngm_field_fwd.  precision FP16: tcgen05 (fp16 operands, the
 * upstream gradient scaled by a per-call power of two, fp32 accumulation); FP32: FFMA kernel, reference arithmetic.
 * Outputs are overwritten (not accumulated).  d_encoding, when given, receives dLoss/d encoding(x)
 * (num_fields, points_per_field, dim_encoding) for the gradient of the encoding's own parameters
 * (ngm_encode_bwd). */
typedef struct NgmFieldBwdArgs {
  NgmFieldFwdArgs fwd;   /* the forward call (its `out` is not read); workspace from ngm_field_bwd_workspace_bytes */
  const float* d_out;    /* (num_fields, points_per_field, dim_out) */
  float* d_weights[NGM_MAX_LINEARS];
  float* d_biases[NGM_MAX_LINEARS];
  float* d_encoding;     /* optional */
} NgmFieldBwdArgs;

/* ---- stage: compositor ---------------------------------------------------------------
 * Replaces the post-MLP split + masks (ngm/run_mapping.py:610-639) and
 * NeuralGraphMap._quadrature (ngm/run_mapping.py:709-799). */
#define NGM_MAX_MIRRORS 8
typedef struct NgmCompositeArgs {
  int64_t num_rays;
  const float* colors;      /* sample colours, element (r,s,c) at colors[(r*S+s)*color_stride + c] */
  const float* geometries;  /* sample geometry, element (r,s) at geometries[(r*S+s)*geometry_stride] */
  const float* distances;   /* (num_rays, S) */
  const float* depths;      /* (num_rays, S) */
  const float* neus_isd;    /* neus: inverse sd per group of rays_per_isd rays (F,), else NULL */
  const float* gt;          /* (num_rays) or NULL */
  int64_t color_stride, geometry_stride; /* (3,1) for separate tensors, (4,4) for packed MLP output */
  int64_t rays_per_isd;
  int32_t num_samples;      /* S */
  int32_t geometry_mode;    /* NgmGeometryMode */
  float geometry_factor, color_factor, truncation;
  int32_t overwrite_behind_camera; /* geometry := fill where depth < 0 (run_mapping.py:614-622) */
  /* device flag gating the overwrite, or NULL = unconditional: the reference switches the overwrite off when
   * every near distance is >= 0 (run_mapping.py:494-495) -- the caller evaluates `(near < 0).any()` on the
   * device and passes its address, so the gate costs no host synchronisation */
  const int32_t* overwrite_gate;
  /* outputs; rgbd is required, the others may be NULL */
  float* rgbd;       /* (num_rays, 4): colour(3), depth */
  float* color_var;  /* (num_rays, 3) */
  float* depth_var;  /* (num_rays) */
  float* term_prob;  /* (num_rays) */
  float* weights;    /* (num_rays, S) sample weights (zero-filled for the dropped last sample) */
  /* dense aux outputs + masks; the host compacts them to the reference's 1-D tensors */
  float* freespace;        /* (num_rays, S): geometry * truncation      (run_mapping.py:625-628) */
  uint8_t* freespace_mask; /* (num_rays, S) */
  float* tsdf;             /* (num_rays, S): geometry*truncation - (gt - dist) (:633-637) */
  uint8_t* tsdf_mask;      /* (num_rays, S) */
  /* multi-GPU tile exchange fused into the compositor (see NgmRenderArgs.mirror_delta): every store of rgbd,
   * color_var, depth_var and term_prob is repeated at `ptr + mirror_delta[i]` bytes, i < num_mirrors, as coalesced
   * 16-byte stores of a warp's 32 rays.  Needs the packed-input compositor (strides (4,4), S % 4 == 0, no weights /
   * aux outputs). */
  int64_t mirror_delta[NGM_MAX_MIRRORS];
  int32_t num_mirrors;
  int32_t _pad_m;
} NgmCompositeArgs;

/* ---- stage: positional encoding on its own ------------------------------------------------
 * ngm/positional_encodings.py (NeRF :245-272, Fourier :197-212, Triplane :132-161, permutohedral
 * wrapper :19-66) as a stand-alone operator, and the permutohedral table gradient -- the two
 * pieces of NeuralField.forward (ngm/models.py:143-145) the training path needs as kernels. */
typedef struct NgmEncodeArgs {
  NgmFieldDesc field;          /* only the encoding members are read */
  int64_t points_per_field;
  const float* points;         /* (num_fields, points_per_field, 3) field-local, scaled coordinates */
  const int64_t* field_slots;  /* row of each field in the encoding tables; NULL = identity */
  float* out;                  /* fwd: (num_fields, points_per_field, dim_encoding) */
  const float* d_out;          /* bwd: upstream gradient, same shape as out */
  float* d_param0;             /* bwd: gradient of field.enc_param0, same layout and per-field stride;
                                  accumulated with atomics -- the caller zeroes it */
  int32_t num_fields;
  int32_t _pad;
} NgmEncodeArgs;

/* ---- stage: compositor backward ---------------------------------------------------------
 * Gradient of the compositor above with respect to the sample colours and geometries (and, in neus
 * mode, the inverse sd): what torch.autograd derives from NeuralGraphMap._quadrature and the
 * post-MLP split (ngm/run_mapping.py:610-639, 709-799) when the training step calls
 * loss.backward() (ngm/run_mapping.py:1186).  Distances / depths carry no gradient (they depend on
 * the camera ray only). */
typedef struct NgmCompositeBwdArgs {
  NgmCompositeArgs fwd;     /* the forward call's inputs (its output pointers are ignored) */
  /* upstream gradients; NULL = zero */
  const float* g_rgbd;      /* (num_rays, 4) */
  const float* g_color_var; /* (num_rays, 3) */
  const float* g_depth_var; /* (num_rays) */
  const float* g_term_prob; /* (num_rays) */
  const float* g_freespace; /* (num_rays, S) dense: zero where the forward's freespace_mask is false */
  const float* g_tsdf;      /* (num_rays, S) dense */
  /* outputs */
  float* d_colors;          /* element (r,s,c) at d_colors[(r*S+s)*color_stride + c] (same layout as fwd.colors) */
  float* d_geometries;      /* element (r,s) at d_geometries[(r*S+s)*geometry_stride] */
  float* d_neus_isd;        /* (num_rays) per-ray gradient of the inverse sd (neus mode; caller sums per field), or NULL */
  float* workspace;         /* >= 2 * num_rays * S floats */
} NgmCompositeBwdArgs;

/* ---- training: Adam on the active fields, in place ---------------------------------------------
 * Replaces the optimizer half of NeuralGraphMap._update_step together with the moment gather of
 * _set_vmap_fields (ngm/run_mapping.py:679-707, 1191-1221): torch.optim.Adam (plain Adam, weight decay
 * added to the gradient; :347-362) on the gathered rows, then the scatter of parameters and moments back
 * into the full per-field tables.  One descriptor per parameter tensor that received a gradient. */
#define NGM_ADAM_MAX_PARAMS 24
typedef struct NgmAdamParam {
  float* param_all;       /* (num_fields_total, row): all_fields_params[name], updated in place at field_ids */
  float* exp_avg_all;     /* (num_fields_total, row): _optim_state[name]["exp_avg"], in place */
  float* exp_avg_sq_all;  /* (num_fields_total, row): _optim_state[name]["exp_avg_sq"], in place */
  const float* grad;      /* (num_active, row): vmap_fields_params[name].grad */
  float* param_active;    /* (num_active, row): vmap_fields_params[name], receives the updated rows; or NULL */
  int64_t row;            /* elements per field */
} NgmAdamParam;
typedef struct NgmAdamArgs {
  const NgmAdamParam* params;  /* HOST array of num_params descriptors (<= NGM_ADAM_MAX_PARAMS) */
  const int64_t* field_ids;    /* device (num_active): row of each active field in the full tables, unique;
                                  NULL = identity */
  int64_t num_active;
  int64_t step;                /* Adam's step count INCLUDING this update (>= 1); shared by all fields of a
                                  parameter tensor, as in the reference (:1213) */
  double lr, beta1, beta2, eps, weight_decay; /* torch.optim.Adam's hyper-parameters (Python floats = doubles) */
  int32_t num_params;
  int32_t _pad;
} NgmAdamArgs;

/* ---- training: multi-view target sampling ------------------------------------------------------
 * The two data-parallel halves of NeuralGraphMap._sample_target_mv (ngm/run_mapping.py:1261-1459), around its one
 * data-dependent step (drop fields no keyframe sees, draw keyframes with torch.multinomial: :1364-1383).
 * Keyframe store as the driver holds it: contiguous poses `_c_c2w_tensor` (num_frames,4,4), the RGB-D buffer
 * `_nc_rgbd_tensor` (num_stored,H,W,4) and the index map `_frame_cid_to_ncid` between them (:1674-1713). */
typedef struct NgmTargetVisArgs {
  NgmCamera cam;
  const float* c2ws;              /* (num_frames, 4, 4) camera-to-world, OpenGL */
  const float* rgbds;             /* (num_stored, H, W, 4): channel 3 = depth */
  const int64_t* frame_to_store;  /* (num_frames) row of each frame in rgbds; NULL = identity */
  const float* positions;         /* field positions table, (>= max field id + 1, 3) */
  const int64_t* field_ids;       /* (num_fields) rows of the candidate fields; NULL = identity */
  const float* probe_offsets;     /* (num_probes, 3) unit vectors (:1322-1323) */
  int64_t num_frames;
  int32_t num_fields;
  int32_t num_probes;
  float train_radius;
  int32_t _pad;
  uint8_t* field_kf_mask;         /* out (num_fields, num_frames): field visible in keyframe (:1356-1362) */
  float* min_xys;                 /* out (num_fields, num_frames, 2): probes' image bounding box, clamped (:1386) */
  float* max_xys;                 /* out (num_fields, num_frames, 2)  (:1387-1389) */
} NgmTargetVisArgs;

typedef struct NgmTargetRaysArgs {
  NgmCamera cam;
  const float* c2ws;              /* as above */
  const float* rgbds;
  const int64_t* frame_to_store;
  const float* positions;
  const int64_t* field_ids;       /* (num_fields) rows of the target fields (those some keyframe sees) */
  const int64_t* frame_cids;      /* (num_fields, rays_per_field) drawn keyframe of every ray (:1381-1383) */
  const float* uv;                /* (num_fields, rays_per_field, 2) uniform draws in [0,1) (:1398-1400) */
  const float* min_xys;           /* (num_fields, num_frames, 2): rows of the target fields */
  const float* max_xys;
  int64_t num_frames;
  int64_t rays_per_field;
  int32_t num_fields;
  float train_radius;
  /* outputs = the members of the reference's Target (:43-58), all (num_fields, rays_per_field, ...) */
  int64_t* ijs;                   /* (.., 2) [row, column] */
  float* out_c2ws;                /* (.., 4, 4) */
  float* near;
  float* far;
  float* gt;                      /* gt_distances */
  float* out_rgbds;               /* (.., 4) */
  uint8_t* rgb_mask;              /* bool */
  uint8_t* depth_mask;            /* bool */
  float* term_probs;
  uint8_t* term_mask;             /* bool */
} NgmTargetRaysArgs;

/* Which fields does the current frame observe: NeuralGraphMap._get_observed_fields (ngm/run_mapping.py:1643-1670)
 * after its torch.nonzero / torch.multinomial choice of valid depth pixels. */
typedef struct NgmObservedArgs {
  NgmCamera cam;
  const float* depth;         /* depth of pixel p (row-major index) at depth[p * pixel_stride]; 4 for channel 3
                                 of an (H, W, 4) RGB-D image */
  const int64_t* pixel_ids;   /* (num_points) row-major indices of the drawn pixels with depth != 0 */
  const float* c2w;           /* (4, 4) pose of the current frame */
  const float* positions;     /* (num_fields, 3): the first num_fields rows of the positions table */
  int64_t pixel_stride;
  int32_t num_points;         /* <= 2048 (the reference draws 500) */
  int32_t num_fields;
  float field_radius;
  int32_t _pad;
  uint8_t* observed;          /* out (num_fields): 1 = observed */
} NgmObservedArgs;

/* ---- fused render --------------------------------------------------------------------
 * Replaces NeuralGraphMap._render_ijs with use_vmap=True (ngm/run_mapping.py:440-666):
 * sampler -> world->local -> encoding -> per-field MLP -> compositor for
 * (num_fields x rays_per_field) rays, ray (f, r) being evaluated by field f only. */
typedef struct NgmRenderArgs {
  NgmFieldDesc field;
  NgmCamera cam;
  int64_t rays_per_field;
  const int64_t* ijs;         /* (num_fields, rays_per_field, 2) */
  const float* c2ws;          /* (4,4) or (num_fields, rays_per_field, 4, 4) */
  const float* near;          /* (num_fields, rays_per_field) or NULL */
  const float* far;
  const float* gt;
  const float* jitter;        /* (num_fields, rays_per_field, num_samples) or NULL -> in-kernel RNG */
  const float* jitter_guided;
  const float* positions;     /* (num_slots, 3)  global field table (_global_map_dict["positions"]) */
  const float* orientations;  /* (num_slots, 4) */
  const int64_t* field_slots; /* (num_fields) = field_ids; NULL = identity */
  const float* neus_sd;       /* (num_slots) `_neus_sd` table, neus mode only (run_mapping.py:641-644) */
  uint64_t seed, offset;
  float near_scalar, far_scalar, range_guided;
  float field_radius;
  float geometry_factor, color_factor, truncation;
  int32_t c2w_per_ray;
  int32_t num_samples, num_samples_guided;
  int32_t num_fields;
  int32_t scale_mode, geometry_mode, precision;
  int32_t overwrite_behind_camera;
  int32_t _pad1;
  const int32_t* overwrite_gate; /* see NgmCompositeArgs.overwrite_gate (run_mapping.py:494-495); NULL = unconditional */
  /* outputs: Prediction (run_mapping.py:59-69); aux ones may be NULL */
  float* rgbd;       /* (num_fields, rays_per_field, 4) */
  float* color_var;  /* (..., 3) */
  float* depth_var;  /* (...) */
  float* term_prob;  /* (...) */
  float* freespace;  uint8_t* freespace_mask; /* (..., St) dense + mask */
  float* tsdf;       uint8_t* tsdf_mask;
  void* workspace;
  size_t workspace_bytes;
  /* multi-GPU tile exchange fused into the render (the reference is single-GPU): every store of the
   * Prediction (rgbd, color_var, depth_var, term_prob) is repeated at `ptr + mirror_delta[i]` BYTES for
   * i < num_mirrors.  The deltas lead from this rank's tile inside a symmetric buffer to the same tile inside the
   * peers' copies of that buffer mapped into this process (NVLink peer memory: num_mirrors = world - 1), or to ONE
   * NVSwitch multicast mapping of it (num_mirrors = 1: the switch replicates the store to every rank).  The caller
   * orders the readers (a barrier over the buffer's signal pads after the kernel).  Performed by the kernel that
   * writes the Prediction: the compositor stage, or the single fused kernel (NGM_RENDER_FUSED=1). */
  int64_t mirror_delta[NGM_MAX_MIRRORS];
  int32_t num_mirrors;
  int32_t _pad2;
} NgmRenderArgs;

/* ---- field set, kNN blend ----------------------------------------------------------------
 * Replaces NeuralFieldSet.forward with use_vmap=False (ngm/models.py:347-405): K nearest field
 * centres per point (pytorch3d knn_points), radius mask on the nearest (:369), per-neighbour
 * local coordinates (:377-381), softmax(-distance_factor*d) blend (:384,399), outside fill (:401).
 * dim_out must be 4 (the reference hard-codes 4 at :388). */
typedef struct NgmKnnFwdArgs {
  NgmFieldDesc field;
  int64_t num_points;
  const float* points;        /* (num_points, 3) world coordinates */
  const float* positions;     /* (num_fields, 3) */
  const float* orientations;  /* (num_fields, 4) */
  const int64_t* field_slots; /* (num_fields) parameter-table row of field i (= field_ids), NULL = identity */
  float* out;                 /* (num_points, 4) */
  void* workspace;
  size_t workspace_bytes;
  float field_radius;   /* radius of the inside test (the `field_radius` argument of forward) */
  float scale_radius;   /* the class radius used by _scale_local_points (models.py:381) */
  float distance_factor, outside_value;
  int32_t num_fields, num_knn, scale_mode, precision;
} NgmKnnFwdArgs;

/* ---- entry points ---------------------------------------------------------------------- */
int ngm_abi_version(void);
const char* ngm_last_error(void);
/* sizeof() of the structs above as compiled, so a binding can verify its mirror:
 * which = 0 NgmCamera, 1 NgmFieldDesc, 2 NgmSampleArgs, 3 NgmFieldFwdArgs, 4 NgmCompositeArgs, 5 NgmRenderArgs,
 * 6 NgmKnnFwdArgs, 7 NgmCompositeBwdArgs, 8 NgmEncodeArgs, 9 NgmAdamParam, 10 NgmAdamArgs,
 * 11 NgmTargetVisArgs, 12 NgmTargetRaysArgs, 13 NgmObservedArgs, 14 NgmFieldBwdArgs */
size_t ngm_struct_size(int which);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t ngm_launch_count(void);

int ngm_sample_rays(const NgmSampleArgs* args, void* stream);   /* camera.py:215-292, run_mapping.py:521-547 */
int ngm_field_fwd(const NgmFieldFwdArgs* args, void* stream);   /* models.py:143-182, 329-345 */
int ngm_field_bwd(const NgmFieldBwdArgs* args, void* stream);   /* autograd of models.py:143-182 under :342 (run_mapping.py:1186) */
int ngm_field_bwd_workspace_bytes(const NgmFieldBwdArgs* args, size_t* out);
int ngm_composite(const NgmCompositeArgs* args, void* stream);  /* run_mapping.py:610-639, 709-799 */
int ngm_composite_bwd(const NgmCompositeBwdArgs* args, void* stream); /* autograd of run_mapping.py:610-639, 709-799 */
int ngm_encode_fwd(const NgmEncodeArgs* args, void* stream);    /* positional_encodings.py forward()s */
int ngm_encode_bwd(const NgmEncodeArgs* args, void* stream);    /* d lattice_values of the permutohedral encoding */
int ngm_adam_step(const NgmAdamArgs* args, void* stream);       /* run_mapping.py:679-707, 1191-1221 */
int ngm_target_visibility(const NgmTargetVisArgs* args, void* stream); /* run_mapping.py:1319-1362, 1386-1389 */
int ngm_target_rays(const NgmTargetRaysArgs* args, void* stream);      /* run_mapping.py:1398-1445 */
int ngm_observed_fields(const NgmObservedArgs* args, void* stream);    /* run_mapping.py:1643-1670 */
int ngm_render_rays_fwd(const NgmRenderArgs* args, void* stream); /* run_mapping.py:440-666 (use_vmap=True) */
int ngm_fieldset_knn_fwd(const NgmKnnFwdArgs* args, void* stream); /* models.py:347-405 (use_vmap=False) */
int ngm_fieldset_knn_workspace_bytes(const NgmKnnFwdArgs* args, size_t* out);

int ngm_field_fwd_workspace_bytes(const NgmFieldFwdArgs* args, size_t* out);
/* persistent kernel-friendly layout of the stacked per-field linears (the tables of
 * models.py:254-264 stay the source of truth): bytes of one row's image, and packing of `num_rows` table rows
 * `rows[i]` (NULL: 0..num_rows-1) into images + rows[i] * bytes. */
int ngm_packed_weights_bytes(const NgmFieldDesc* field, size_t* bytes_per_row);
int ngm_pack_weights(const NgmFieldDesc* field, const int64_t* rows, int32_t num_rows, void* images, void* stream);
int ngm_render_workspace_bytes(const NgmRenderArgs* args, size_t* out);

#ifdef __cplusplus
}
#endif
#endif /* NGM_B200_H_ */
