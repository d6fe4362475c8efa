// extern "C" entry points of libngm_b200.so (see include/ngm_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ngm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// kernels (one translation unit each)
int launch_sample_rays(const NgmSampleArgs& a, cudaStream_t stream);
int launch_composite(const NgmCompositeArgs& a, cudaStream_t stream);
int launch_neus_isd(const float* sd, const int64_t* slots, int num_fields, float* out, cudaStream_t stream);
int launch_field_fwd_simt(const NgmFieldFwdArgs& a, cudaStream_t stream);
size_t field_simt_workspace_bytes(const NgmFieldFwdArgs& a);
size_t field_simt_smem_bytes(const NgmFieldDesc& fd, int* act_stride, int* enc_stride);
// fp16 tcgen05 path (field_tc.cu)
bool field_tc_supported(const NgmFieldDesc& fd, const char** why);
size_t field_tc_workspace_bytes(const NgmFieldDesc& fd, int num_fields);
size_t field_tc_fwd_workspace_bytes(const NgmFieldFwdArgs& a);
bool tc_rows_required(const NgmFieldDesc& fd);
int launch_field_fwd_tc(const NgmFieldFwdArgs& a, cudaStream_t stream);
bool field_bwd_tc_supported(const NgmFieldDesc& fd, const char** why);
size_t field_bwd_tc_workspace_bytes(const NgmFieldBwdArgs& b);
int launch_field_bwd_tc(const NgmFieldBwdArgs& b, cudaStream_t stream);
size_t packed_weights_bytes(const NgmFieldDesc& fd);
int launch_pack_weights(const NgmFieldDesc& fd, const int64_t* rows, int num_rows, void* images, cudaStream_t stream);
bool render_fused_tc_ok(const NgmRenderArgs& a);
int launch_render_fused_tc(const NgmRenderArgs& a, void* tc_ws, float* isd_ws, const void* rows_half, const float* dist,
                           const float* depth, cudaStream_t stream);
bool render_tc_precoded(const NgmRenderArgs& a);
int launch_permuto_rows_half(const PermutoRowsArgs& a, cudaStream_t stream);
#ifdef NGM_DEBUG_EXPORTS
int launch_tc_gemm_debug(const NgmFieldDesc& fd, const void* a_half, long long rows, float* out, void* workspace,
                         cudaStream_t stream);
int tc_trace_read(unsigned long long* out, int max_events);
int tc_trace_peek(unsigned long long* out, int max_events);
int bwd_phases_read(unsigned long long* out16);
#endif

size_t knn_workspace_bytes(const NgmKnnFwdArgs& a);
int launch_fieldset_knn(const NgmKnnFwdArgs& a, cudaStream_t stream);

int launch_composite_bwd(const NgmCompositeBwdArgs& b, cudaStream_t stream);
int launch_encode_fwd(const NgmEncodeArgs& a, cudaStream_t stream);
int launch_encode_bwd(const NgmEncodeArgs& a, cudaStream_t stream);
int launch_adam_step(const NgmAdamArgs& a, cudaStream_t stream);
int launch_target_visibility(const NgmTargetVisArgs& a, cudaStream_t stream);
int launch_target_rays(const NgmTargetRaysArgs& a, cudaStream_t stream);
int launch_observed_fields(const NgmObservedArgs& a, cudaStream_t stream);

static int validate_field(const NgmFieldDesc& fd) {
  NGM_CHECK_ARG(fd.num_layers >= 0 && fd.num_layers + 1 <= NGM_MAX_LINEARS, "num_layers=%d out of range [0,%d]",
                fd.num_layers, NGM_MAX_LINEARS - 1);
  NGM_CHECK_ARG(fd.dim_encoding > 0 && fd.dim_mlp_out > 0 && fd.dim_out > 0, "non-positive layer width");
  NGM_CHECK_ARG(fd.skip_mode >= NGM_SKIP_NO && fd.skip_mode <= NGM_SKIP_REZERO,
                "Skip mode %d is not available.", fd.skip_mode);  // models.py:110
  NGM_CHECK_ARG(fd.encoding >= NGM_ENC_NERF && fd.encoding <= NGM_ENC_PERMUTO, "unknown encoding %d", fd.encoding);
  if (fd.skip_mode == NGM_SKIP_ADD || fd.skip_mode == NGM_SKIP_REZERO)
    NGM_CHECK_ARG(fd.dim_mlp_out >= fd.dim_encoding, "skip_mode add/rezero needs dim_mlp_out >= dim_encoding");
  if (fd.skip_mode == NGM_SKIP_REZERO) NGM_CHECK_ARG(fd.rezero != nullptr, "skip_mode rezero without _rezero");
  for (int i = 0; i <= fd.num_layers; ++i)
    NGM_CHECK_ARG(fd.weights[i] && fd.biases[i], "missing _linears.%d parameters", i);
  switch (fd.encoding) {
    case NGM_ENC_NERF:
      NGM_CHECK_ARG(fd.dim_encoding == 6 * fd.nerf_num_octaves, "nerf: dim_encoding != 2*3*num_octaves");
      break;
    case NGM_ENC_FOURIER:
      NGM_CHECK_ARG(fd.enc_param0 && fd.dim_encoding == fd.fourier_num_features + (fd.fourier_raw_coords ? 3 : 0),
                    "fourier: bad dims or missing weight");
      break;
    case NGM_ENC_TRIPLANE:
      NGM_CHECK_ARG(fd.enc_param0 && fd.triplane_resolution >= 2 &&
                        fd.dim_encoding == fd.triplane_components * (fd.triplane_mode == NGM_TRIPLANE_CONCAT ? 3 : 1),
                    "triplane: bad dims or missing plane_coef");
      break;
    case NGM_ENC_PERMUTO:
      NGM_CHECK_ARG(fd.enc_param0 && fd.enc_param1 && fd.permuto_scale, "permuto: missing table/shift/scale");
      NGM_CHECK_ARG(fd.permuto_feats >= 1 && fd.permuto_feats <= 8 && fd.permuto_log2_capacity >= 1 &&
                        fd.permuto_log2_capacity <= 30 &&
                        fd.dim_encoding == fd.permuto_levels * fd.permuto_feats + (fd.permuto_concat_points ? 3 : 0),
                    "permuto: unsupported dims");
      break;
  }
  return NGM_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct RenderWorkspace {
  size_t points_world, distances, depths, outs, rows_half, rows_f32, rows_f32_bytes, isd, tc, total;
};

static RenderWorkspace render_workspace(const NgmRenderArgs& a) {
  RenderWorkspace w{};
  const int St = a.num_samples + (a.gt ? a.num_samples_guided : 0);
  const size_t n = (size_t)a.num_fields * (size_t)a.rays_per_field * (size_t)St;
  size_t off = 0;
  const bool fused = render_fused_tc_ok(a);
  const bool precoded = fused && render_tc_precoded(a);
  if (!fused || precoded) {  // the fused NeRF kernel keeps every per-sample intermediate on chip
    w.points_world = off; off = align_up(off + n * 3 * sizeof(float), 256);
    w.distances = off;    off = align_up(off + n * sizeof(float), 256);
    w.depths = off;       off = align_up(off + n * sizeof(float), 256);
  }
  if (!fused) { w.outs = off; off = align_up(off + n * 4 * sizeof(float), 256); }
  if (precoded) {  // fp16 A-operand rows of the permutohedral encoding
    const size_t ep = (size_t)(a.field.dim_encoding + 15) / 16 * 16;
    w.rows_half = off; off = align_up(off + n * ep * 2, 256);
  }
  if (!fused && a.precision == NGM_PREC_FP32 && a.field.encoding == NGM_ENC_PERMUTO && a.field.permuto_feats == 2) {
    w.rows_f32 = off;  // fp32 rows for the FFMA field kernel
    w.rows_f32_bytes = n * (size_t)a.field.dim_encoding * sizeof(float);
    off = align_up(off + w.rows_f32_bytes, 256);
  }
  w.isd = off;         off = align_up(off + (size_t)a.num_fields * sizeof(float), 256);
  w.tc = off;
  if (a.precision == NGM_PREC_FP16) {
    off = align_up(off + field_tc_workspace_bytes(a.field, a.num_fields), 256);
    // staged path (> 128 samples per ray): rows of the dense field evaluation, when it encodes them itself
    // (field_tc.cu: fwd_rows_from_encoder -- always for the permutohedral encoding unless NGM_TC_PERMUTO_INKERNEL=1)
    if (!fused && (tc_rows_required(a.field) || a.field.encoding == NGM_ENC_PERMUTO))
      off = align_up(off + n * ((size_t)(a.field.dim_encoding + 15) / 16 * 16) * 2 + 256, 256);
  }
  w.total = off;
  return w;
}

}  // namespace ngm

using namespace ngm;

extern "C" {

int ngm_abi_version(void) { return NGM_ABI_VERSION; }

const char* ngm_last_error(void) { return g_error; }

uint64_t ngm_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

size_t ngm_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(NgmCamera);
    case 1: return sizeof(NgmFieldDesc);
    case 2: return sizeof(NgmSampleArgs);
    case 3: return sizeof(NgmFieldFwdArgs);
    case 4: return sizeof(NgmCompositeArgs);
    case 5: return sizeof(NgmRenderArgs);
    case 6: return sizeof(NgmKnnFwdArgs);
    case 7: return sizeof(NgmCompositeBwdArgs);
    case 8: return sizeof(NgmEncodeArgs);
    case 9: return sizeof(NgmAdamParam);
    case 10: return sizeof(NgmAdamArgs);
    case 11: return sizeof(NgmTargetVisArgs);
    case 12: return sizeof(NgmTargetRaysArgs);
    case 13: return sizeof(NgmObservedArgs);
    case 14: return sizeof(NgmFieldBwdArgs);
    default: return 0;
  }
}

int ngm_sample_rays(const NgmSampleArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_rays >= 0 && a->num_samples > 0 && a->num_samples_guided >= 0, "bad ray/sample counts");
  NGM_CHECK_ARG(a->num_rays == 0 || (a->ijs && a->c2ws), "ijs / c2ws missing");
  NGM_CHECK_ARG(a->cam.fx != 0.f && a->cam.fy != 0.f, "zero focal length");
  return launch_sample_rays(*a, (cudaStream_t)stream);
}

int ngm_field_fwd_workspace_bytes(const NgmFieldFwdArgs* a, size_t* out) {
  NGM_CHECK_ARG(a && out, "null args");
  *out = a->precision == NGM_PREC_FP16 ? field_tc_fwd_workspace_bytes(*a) : field_simt_workspace_bytes(*a);
  return NGM_OK;
}

int ngm_field_fwd(const NgmFieldFwdArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  if (int rc = validate_field(a->field)) return rc;
  NGM_CHECK_ARG(a->num_fields >= 0 && a->points_per_field >= 0, "negative sizes");
  if (a->num_fields == 0 || a->points_per_field == 0) return NGM_OK;
  NGM_CHECK_ARG(a->points && a->out, "points / out missing");
  NGM_CHECK_ARG((a->positions == nullptr) == (a->orientations == nullptr),
                "positions and orientations must be given together");
  NGM_CHECK_ARG(a->scale_mode >= NGM_SCALE_NO && a->scale_mode <= NGM_SCALE_UNIT_CUBE, "scale_mode=%d is not available.",
                a->scale_mode);  // models.py:285
  if (a->scale_mode != NGM_SCALE_NO)
    NGM_CHECK_ARG(a->field_radius > 0.f, "scale_mode requires field_radius to be specified.");  // models.py:219
  if (a->precision == NGM_PREC_FP16) {
    const char* why = nullptr;
    NGM_UNSUPPORTED(!field_tc_supported(a->field, &why), "fp16 tensor-core path unsupported: %s", why);
    size_t need = field_tc_fwd_workspace_bytes(*a);
    if (need > a->workspace_bytes || (need && !a->workspace)) {
      set_error("workspace too small: need %zu B, have %zu B", need, a->workspace_bytes);
      return NGM_ERR_WORKSPACE;
    }
    return launch_field_fwd_tc(*a, (cudaStream_t)stream);
  }
  NGM_CHECK_ARG(a->precision == NGM_PREC_FP32, "unknown precision %d", a->precision);
  return launch_field_fwd_simt(*a, (cudaStream_t)stream);
}

static int validate_field_bwd(const NgmFieldBwdArgs* b) {
  NGM_CHECK_ARG(b != nullptr, "null args");
  const NgmFieldFwdArgs* a = &b->fwd;
  if (int rc = validate_field(a->field)) return rc;
  NGM_CHECK_ARG(a->num_fields >= 0 && a->points_per_field >= 0, "negative sizes");
  NGM_CHECK_ARG((a->positions == nullptr) == (a->orientations == nullptr),
                "positions and orientations must be given together");
  NGM_CHECK_ARG(a->scale_mode >= NGM_SCALE_NO && a->scale_mode <= NGM_SCALE_UNIT_CUBE, "scale_mode=%d is not available.",
                a->scale_mode);
  return NGM_OK;
}

int ngm_field_bwd_workspace_bytes(const NgmFieldBwdArgs* b, size_t* out) {
  NGM_CHECK_ARG(b && out, "null args");
  NGM_UNSUPPORTED(b->fwd.precision != NGM_PREC_FP16, "ngm_field_bwd: only the fp16 tensor-core path is built");
  *out = field_bwd_tc_workspace_bytes(*b);
  return NGM_OK;
}

int ngm_field_bwd(const NgmFieldBwdArgs* b, void* stream) {
  if (int rc = validate_field_bwd(b)) return rc;
  const NgmFieldFwdArgs* a = &b->fwd;
  if (a->num_fields == 0) return NGM_OK;
  NGM_CHECK_ARG((a->points || a->rows_half) && b->d_out, "points / d_out missing");
  NGM_UNSUPPORTED(a->precision != NGM_PREC_FP16, "ngm_field_bwd: only the fp16 tensor-core path is built");
  const char* why = nullptr;
  NGM_UNSUPPORTED(!field_bwd_tc_supported(a->field, &why), "fp16 tensor-core backward unsupported: %s", why);
  const size_t need = field_bwd_tc_workspace_bytes(*b);
  if (need > a->workspace_bytes || !a->workspace) {
    set_error("workspace too small: need %zu B, have %zu B", need, a->workspace_bytes);
    return NGM_ERR_WORKSPACE;
  }
  return launch_field_bwd_tc(*b, (cudaStream_t)stream);
}

int ngm_packed_weights_bytes(const NgmFieldDesc* fd, size_t* out) {
  NGM_CHECK_ARG(fd && out, "null args");
  if (int rc = validate_field(*fd)) return rc;
  const char* why = nullptr;
  NGM_UNSUPPORTED(!field_tc_supported(*fd, &why), "fp16 tensor-core path unsupported: %s", why);
  *out = packed_weights_bytes(*fd);
  return NGM_OK;
}

int ngm_pack_weights(const NgmFieldDesc* fd, const int64_t* rows, int32_t num_rows, void* images, void* stream) {
  NGM_CHECK_ARG(fd && images, "null args");
  NGM_CHECK_ARG(num_rows >= 0, "negative num_rows");
  if (int rc = validate_field(*fd)) return rc;
  const char* why = nullptr;
  NGM_UNSUPPORTED(!field_tc_supported(*fd, &why), "fp16 tensor-core path unsupported: %s", why);
  NGM_CHECK_ARG(((uintptr_t)images & 15) == 0, "images must be 16-byte aligned");
  return launch_pack_weights(*fd, rows, num_rows, images, (cudaStream_t)stream);
}

int ngm_composite(const NgmCompositeArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_rays >= 0 && a->num_samples > 0, "bad ray/sample counts");
  if (a->num_rays == 0) return NGM_OK;
  NGM_CHECK_ARG(a->colors && a->geometries && a->distances && a->depths && a->rgbd, "missing input/output");
  NGM_CHECK_ARG(a->geometry_mode >= NGM_GEOM_DENSITY && a->geometry_mode <= NGM_GEOM_NRGBD, "unknown geometry_mode %d",
                a->geometry_mode);
  if (a->geometry_mode == NGM_GEOM_NEUS)
    NGM_CHECK_ARG(a->neus_isd && a->rays_per_isd > 0, "neus mode needs neus_isd");
  NGM_CHECK_ARG((a->freespace == nullptr) == (a->freespace_mask == nullptr), "freespace and its mask go together");
  NGM_CHECK_ARG((a->tsdf == nullptr) == (a->tsdf_mask == nullptr), "tsdf and its mask go together");
  NGM_CHECK_ARG(((uintptr_t)a->rgbd & 15) == 0, "rgbd must be 16-byte aligned");
  NGM_CHECK_ARG(a->num_mirrors >= 0 && a->num_mirrors <= NGM_MAX_MIRRORS, "num_mirrors=%d outside [0, %d]", a->num_mirrors,
                NGM_MAX_MIRRORS);
  for (int i = 0; i < a->num_mirrors; ++i)
    NGM_CHECK_ARG((a->mirror_delta[i] & 15) == 0, "mirror_delta[%d] must be a multiple of 16 bytes", i);
  return launch_composite(*a, (cudaStream_t)stream);
}

int ngm_composite_bwd(const NgmCompositeBwdArgs* b, void* stream) {
  NGM_CHECK_ARG(b != nullptr, "null args");
  const NgmCompositeArgs* a = &b->fwd;
  NGM_CHECK_ARG(a->num_rays >= 0 && a->num_samples > 0, "bad ray/sample counts");
  if (a->num_rays == 0) return NGM_OK;
  NGM_CHECK_ARG(a->colors && a->geometries && a->distances && a->depths, "missing forward input");
  NGM_CHECK_ARG(b->d_colors && b->d_geometries && b->workspace, "missing output / workspace");
  NGM_CHECK_ARG(a->geometry_mode >= NGM_GEOM_DENSITY && a->geometry_mode <= NGM_GEOM_NRGBD, "unknown geometry_mode %d",
                a->geometry_mode);
  if (a->geometry_mode == NGM_GEOM_NEUS)
    NGM_CHECK_ARG(a->neus_isd && a->rays_per_isd > 0, "neus mode needs neus_isd");
  return launch_composite_bwd(*b, (cudaStream_t)stream);
}

int ngm_adam_step(const NgmAdamArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_params >= 0 && a->num_params <= NGM_ADAM_MAX_PARAMS, "num_params %d outside [0, %d]", a->num_params,
                NGM_ADAM_MAX_PARAMS);
  NGM_CHECK_ARG(a->num_active >= 0, "negative num_active");
  NGM_CHECK_ARG(a->step >= 1, "step must count this update (>= 1)");
  NGM_CHECK_ARG(a->lr >= 0.0 && a->eps >= 0.0 && a->weight_decay >= 0.0, "negative lr / eps / weight_decay");
  NGM_CHECK_ARG(a->beta1 >= 0.0 && a->beta1 < 1.0 && a->beta2 >= 0.0 && a->beta2 < 1.0, "betas outside [0, 1)");
  NGM_CHECK_ARG(a->num_params == 0 || a->params, "params missing");
  for (int i = 0; i < a->num_params; ++i) {
    const NgmAdamParam& d = a->params[i];
    NGM_CHECK_ARG(d.row >= 0, "negative row size");
    NGM_CHECK_ARG(d.row == 0 || a->num_active == 0 || (d.param_all && d.exp_avg_all && d.exp_avg_sq_all && d.grad),
                  "parameter %d: missing table / moment / gradient pointer", i);
  }
  return launch_adam_step(*a, (cudaStream_t)stream);
}

int ngm_target_visibility(const NgmTargetVisArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_fields >= 0 && a->num_frames >= 0 && a->num_probes > 0, "bad field / frame / probe counts");
  NGM_CHECK_ARG(a->cam.fx != 0.f && a->cam.fy != 0.f && a->cam.width > 0 && a->cam.height > 0, "bad camera");
  if (a->num_fields == 0 || a->num_frames == 0) return NGM_OK;
  NGM_CHECK_ARG(a->c2ws && a->rgbds && a->positions && a->probe_offsets, "missing keyframe store / positions / probes");
  NGM_CHECK_ARG(a->field_kf_mask && a->min_xys && a->max_xys, "missing output");
  return launch_target_visibility(*a, (cudaStream_t)stream);
}

int ngm_target_rays(const NgmTargetRaysArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_fields >= 0 && a->rays_per_field >= 0 && a->num_frames > 0, "bad field / ray / frame counts");
  NGM_CHECK_ARG(a->cam.fx != 0.f && a->cam.fy != 0.f && a->cam.width > 0 && a->cam.height > 0, "bad camera");
  if (a->num_fields == 0 || a->rays_per_field == 0) return NGM_OK;
  NGM_CHECK_ARG(a->c2ws && a->rgbds && a->positions && a->frame_cids && a->uv && a->min_xys && a->max_xys,
                "missing input");
  NGM_CHECK_ARG(a->ijs && a->out_c2ws && a->near && a->far && a->gt && a->out_rgbds && a->rgb_mask && a->depth_mask &&
                    a->term_probs && a->term_mask,
                "missing output");
  return launch_target_rays(*a, (cudaStream_t)stream);
}

int ngm_observed_fields(const NgmObservedArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  NGM_CHECK_ARG(a->num_fields >= 0 && a->num_points >= 0 && a->pixel_stride > 0, "bad field / point counts or stride");
  NGM_CHECK_ARG(a->cam.fx != 0.f && a->cam.fy != 0.f && a->cam.width > 0 && a->cam.height > 0, "bad camera");
  if (a->num_fields == 0) return NGM_OK;
  NGM_CHECK_ARG(a->c2w && a->positions && a->observed, "missing pose / positions / output");
  NGM_CHECK_ARG(a->num_points == 0 || (a->depth && a->pixel_ids), "missing depth image / pixel ids");
  return launch_observed_fields(*a, (cudaStream_t)stream);
}

static int validate_encoding(const NgmFieldDesc& fd) {
  NGM_CHECK_ARG(fd.encoding >= NGM_ENC_NERF && fd.encoding <= NGM_ENC_PERMUTO, "unknown encoding %d", fd.encoding);
  NGM_CHECK_ARG(fd.dim_encoding > 0, "non-positive encoding width");
  switch (fd.encoding) {
    case NGM_ENC_NERF:
      NGM_CHECK_ARG(fd.dim_encoding == 6 * fd.nerf_num_octaves, "nerf: dim_encoding != 2*3*num_octaves");
      break;
    case NGM_ENC_FOURIER:
      NGM_CHECK_ARG(fd.enc_param0 && fd.dim_encoding == fd.fourier_num_features + (fd.fourier_raw_coords ? 3 : 0),
                    "fourier: bad dims or missing weight");
      break;
    case NGM_ENC_TRIPLANE:
      NGM_CHECK_ARG(fd.enc_param0 && fd.triplane_resolution >= 2 &&
                        fd.dim_encoding == fd.triplane_components * (fd.triplane_mode == NGM_TRIPLANE_CONCAT ? 3 : 1),
                    "triplane: bad dims or missing plane_coef");
      break;
    default:
      NGM_CHECK_ARG(fd.enc_param0 && fd.enc_param1 && fd.permuto_scale, "permuto: missing table/shift/scale");
      NGM_CHECK_ARG(fd.permuto_feats >= 1 && fd.permuto_feats <= 8 && fd.permuto_log2_capacity >= 1 &&
                        fd.permuto_log2_capacity <= 30 &&
                        fd.dim_encoding == fd.permuto_levels * fd.permuto_feats + (fd.permuto_concat_points ? 3 : 0),
                    "permuto: unsupported dims");
  }
  return NGM_OK;
}

int ngm_encode_fwd(const NgmEncodeArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  if (int rc = validate_encoding(a->field)) return rc;
  NGM_CHECK_ARG(a->num_fields >= 0 && a->points_per_field >= 0, "negative sizes");
  if (a->num_fields == 0 || a->points_per_field == 0) return NGM_OK;
  NGM_CHECK_ARG(a->points && a->out, "points / out missing");
  return launch_encode_fwd(*a, (cudaStream_t)stream);
}

int ngm_encode_bwd(const NgmEncodeArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  if (int rc = validate_encoding(a->field)) return rc;
  NGM_CHECK_ARG(a->num_fields >= 0 && a->points_per_field >= 0, "negative sizes");
  if (a->num_fields == 0 || a->points_per_field == 0) return NGM_OK;
  NGM_CHECK_ARG(a->points && a->d_out && a->d_param0, "points / d_out / d_param0 missing");
  return launch_encode_bwd(*a, (cudaStream_t)stream);
}

int ngm_fieldset_knn_workspace_bytes(const NgmKnnFwdArgs* a, size_t* out) {
  NGM_CHECK_ARG(a && out, "null args");
  NGM_CHECK_ARG(a->num_points >= 0 && a->num_fields >= 1 && a->num_knn >= 1, "bad sizes");
  *out = knn_workspace_bytes(*a);
  return NGM_OK;
}

int ngm_fieldset_knn_fwd(const NgmKnnFwdArgs* a, void* stream) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  if (int rc = validate_field(a->field)) return rc;
  NGM_CHECK_ARG(a->field.dim_out == 4, "the kNN blend path needs dim_out == 4 (models.py:388), got %d", a->field.dim_out);
  NGM_CHECK_ARG(a->num_points >= 0 && a->num_fields >= 1, "bad sizes");
  NGM_CHECK_ARG(a->num_knn >= 1 && a->num_knn <= 8, "num_knn must be in [1, 8], got %d", a->num_knn);
  const int K = a->num_knn < a->num_fields ? a->num_knn : a->num_fields;
  NGM_CHECK_ARG(a->num_points * (long long)K < (1ll << 31), "num_points * K must fit in int32; evaluate in blocks");
  if (a->num_points == 0) return NGM_OK;
  NGM_CHECK_ARG(a->points && a->positions && a->orientations && a->out, "missing input/output");
  NGM_CHECK_ARG(a->scale_mode >= NGM_SCALE_NO && a->scale_mode <= NGM_SCALE_UNIT_CUBE, "scale_mode=%d is not available.",
                a->scale_mode);
  NGM_CHECK_ARG(((uintptr_t)a->out & 15) == 0, "out must be 16-byte aligned");
  if (a->precision == NGM_PREC_FP16) {
    const char* why = nullptr;
    NGM_UNSUPPORTED(!field_tc_supported(a->field, &why), "fp16 tensor-core path unsupported: %s", why);
  } else {
    NGM_CHECK_ARG(a->precision == NGM_PREC_FP32, "unknown precision %d", a->precision);
  }
  const size_t need = knn_workspace_bytes(*a);
  if (need > a->workspace_bytes || !a->workspace) {
    set_error("workspace too small: need %zu B, have %zu B", need, a->workspace_bytes);
    return NGM_ERR_WORKSPACE;
  }
  return launch_fieldset_knn(*a, (cudaStream_t)stream);
}

#ifdef NGM_DEBUG_EXPORTS
/* diagnostics: only in libngm_b200_debug.so (include/ngm_b200_debug.h) */
int ngm_debug_tc_trace_peek(uint64_t* host_out, int max_events) {
  NGM_CHECK_ARG(host_out && max_events > 0, "null args");
  return tc_trace_peek(reinterpret_cast<unsigned long long*>(host_out), max_events);
}

int ngm_debug_bwd_phases(uint64_t* host_out16) {
  NGM_CHECK_ARG(host_out16 != nullptr, "null args");
  return bwd_phases_read(reinterpret_cast<unsigned long long*>(host_out16));
}

int ngm_debug_tc_trace(uint64_t* host_out, int max_events) {
  NGM_CHECK_ARG(host_out && max_events > 0, "null args");
  return tc_trace_read(reinterpret_cast<unsigned long long*>(host_out), max_events);
}

int ngm_debug_tc_gemm(const float* weight, const float* bias, int n, int k, const void* a_half, int64_t rows, float* out,
                      void* workspace, size_t workspace_bytes, void* stream) {
  NGM_CHECK_ARG(weight && bias && a_half && out && workspace, "null pointer");
  NGM_CHECK_ARG(n >= 1 && n <= 128 && k >= 16 && k <= 128 && k % 16 == 0 && rows > 0, "need 1<=n<=128, k%%16==0, 16<=k<=128");
  NgmFieldDesc fd{};
  fd.encoding = NGM_ENC_NERF;
  fd.dim_encoding = k;
  fd.num_layers = 0;
  fd.dim_mlp_out = 16;
  fd.dim_out = n;
  fd.weights[0] = weight;
  fd.biases[0] = bias;
  NGM_CHECK_ARG(workspace_bytes >= (size_t)((k + 63) / 64) * ((n + 15) / 16 * 16) * 128 + 1024, "workspace too small");
  return launch_tc_gemm_debug(fd, a_half, rows, out, workspace, (cudaStream_t)stream);
}
#endif  // NGM_DEBUG_EXPORTS

int ngm_render_workspace_bytes(const NgmRenderArgs* a, size_t* out) {
  NGM_CHECK_ARG(a && out, "null args");
  *out = render_workspace(*a).total;
  return NGM_OK;
}

int ngm_render_rays_fwd(const NgmRenderArgs* a, void* stream_) {
  NGM_CHECK_ARG(a != nullptr, "null args");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = validate_field(a->field)) return rc;
  NGM_CHECK_ARG(a->field.dim_out == 4, "the renderer needs dim_out == 4 (rgb + geometry), got %d", a->field.dim_out);
  NGM_CHECK_ARG(a->num_fields >= 0 && a->rays_per_field >= 0 && a->num_samples > 0 && a->num_samples_guided >= 0,
                "bad sizes");
  const long long num_rays = (long long)a->num_fields * a->rays_per_field;
  if (num_rays == 0) return NGM_OK;
  NGM_CHECK_ARG(a->ijs && a->c2ws && a->positions && a->orientations && a->rgbd, "missing input/output");
  NGM_CHECK_ARG(((uintptr_t)a->rgbd & 15) == 0, "rgbd must be 16-byte aligned");
  NGM_CHECK_ARG(a->geometry_mode >= NGM_GEOM_DENSITY && a->geometry_mode <= NGM_GEOM_NRGBD, "unknown geometry_mode %d",
                a->geometry_mode);
  if (a->geometry_mode == NGM_GEOM_NEUS) NGM_CHECK_ARG(a->neus_sd != nullptr, "neus mode needs the _neus_sd table");
  const RenderWorkspace w = render_workspace(*a);
  if (w.total > a->workspace_bytes || !a->workspace) {
    set_error("workspace too small: need %zu B, have %zu B", w.total, a->workspace_bytes);
    return NGM_ERR_WORKSPACE;
  }
  if (a->precision == NGM_PREC_FP16) {
    const char* why = nullptr;
    NGM_UNSUPPORTED(!field_tc_supported(a->field, &why), "fp16 tensor-core path unsupported: %s", why);
  } else {
    NGM_CHECK_ARG(a->precision == NGM_PREC_FP32, "unknown precision %d", a->precision);
  }
  char* ws = static_cast<char*>(a->workspace);
  const bool fused = render_fused_tc_ok(*a);
  const bool precoded = fused && render_tc_precoded(*a);
  NGM_CHECK_ARG(a->num_mirrors >= 0 && a->num_mirrors <= NGM_MAX_MIRRORS, "num_mirrors=%d outside [0, %d]",
                a->num_mirrors, NGM_MAX_MIRRORS);
  for (int i = 0; i < a->num_mirrors; ++i)
    NGM_CHECK_ARG((a->mirror_delta[i] & 15) == 0, "mirror_delta[%d] must be a multiple of 16 bytes", i);
  if (fused && !precoded)  // one fused tcgen05 kernel per render batch
    return launch_render_fused_tc(*a, ws + w.tc, reinterpret_cast<float*>(ws + w.isd), nullptr, nullptr, nullptr, stream);

  // staged path: the stage kernels over workspace intermediates (fp32 reference arithmetic, fp16 tensor-core
  // field evaluation when a ray has more than 128 samples, or sampler + row encoder ahead of the fused kernel
  // for the permutohedral encoding).
  const int St = a->num_samples + (a->gt ? a->num_samples_guided : 0);
  NgmSampleArgs s{};
  s.cam = a->cam;
  s.num_rays = num_rays;
  s.ijs = a->ijs; s.c2ws = a->c2ws; s.near = a->near; s.far = a->far; s.gt = a->gt;
  s.jitter = a->jitter; s.jitter_guided = a->jitter_guided;
  s.seed = a->seed; s.offset = a->offset;
  s.near_scalar = a->near_scalar; s.far_scalar = a->far_scalar; s.range_guided = a->range_guided;
  s.c2w_per_ray = a->c2w_per_ray;
  s.num_samples = a->num_samples; s.num_samples_guided = a->num_samples_guided;
  s.points_cam = nullptr;
  s.points_world = reinterpret_cast<float*>(ws + w.points_world);
  s.distances = reinterpret_cast<float*>(ws + w.distances);
  s.depths = reinterpret_cast<float*>(ws + w.depths);
  if (int rc = launch_sample_rays(s, stream)) return rc;

  if (precoded) {
    PermutoRowsArgs e{};
    e.field = a->field;
    e.points_world = s.points_world;
    e.positions = a->positions; e.orientations = a->orientations;
    e.field_slots = reinterpret_cast<const long long*>(a->field_slots);
    e.out = reinterpret_cast<uint32_t*>(ws + w.rows_half);
    e.num_points = num_rays * St;
    e.points_per_field = a->rays_per_field * St;
    e.field_radius = a->field_radius;
    e.scale_mode = a->scale_mode;
    e.EP = (a->field.dim_encoding + 15) / 16 * 16;
    if (int rc = launch_permuto_rows_half(e, stream)) return rc;
    return launch_render_fused_tc(*a, ws + w.tc, reinterpret_cast<float*>(ws + w.isd), ws + w.rows_half, s.distances,
                                  s.depths, stream);
  }

  NgmFieldFwdArgs f{};
  f.field = a->field;
  f.points_per_field = a->rays_per_field * St;
  f.points = s.points_world;
  f.positions = a->positions; f.orientations = a->orientations; f.field_slots = a->field_slots;
  f.out = reinterpret_cast<float*>(ws + w.outs);
  f.field_radius = a->field_radius;
  f.num_fields = a->num_fields;
  f.scale_mode = a->scale_mode;
  f.precision = a->precision;
  f.workspace = w.rows_f32_bytes ? ws + w.rows_f32 : ws + w.tc;
  f.workspace_bytes = w.rows_f32_bytes ? w.rows_f32_bytes : w.total - w.tc;
  if (int rc = ngm_field_fwd(&f, stream_)) return rc;

  NgmCompositeArgs c{};
  c.num_rays = num_rays;
  c.colors = f.out; c.geometries = f.out + 3;
  c.color_stride = 4; c.geometry_stride = 4;
  c.distances = s.distances; c.depths = s.depths;
  c.gt = a->gt;
  c.num_samples = St;
  c.geometry_mode = a->geometry_mode;
  c.geometry_factor = a->geometry_factor; c.color_factor = a->color_factor; c.truncation = a->truncation;
  c.overwrite_behind_camera = a->overwrite_behind_camera;
  c.overwrite_gate = a->overwrite_gate;
  c.rgbd = a->rgbd; c.color_var = a->color_var; c.depth_var = a->depth_var; c.term_prob = a->term_prob;
  c.freespace = a->freespace; c.freespace_mask = a->freespace_mask;
  c.tsdf = a->tsdf; c.tsdf_mask = a->tsdf_mask;
  c.num_mirrors = a->num_mirrors;  // the compositor writes the Prediction: it also repeats it for the peers
  for (int i = 0; i < a->num_mirrors; ++i) c.mirror_delta[i] = a->mirror_delta[i];
  if (a->geometry_mode == NGM_GEOM_NEUS) {
    // neus_isds = 1 / |_neus_sd[field_ids]|  (run_mapping.py:641-644)
    float* isd = reinterpret_cast<float*>(ws + w.isd);
    if (int rc = launch_neus_isd(a->neus_sd, a->field_slots, a->num_fields, isd, stream)) return rc;
    c.neus_isd = isd;
    c.rays_per_isd = a->rays_per_field;
  }
  return ngm_composite(&c, stream_);
}

}  // extern "C"
