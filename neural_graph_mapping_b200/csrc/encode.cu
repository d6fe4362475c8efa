// Stage kernels: positional encodings on their own (field-local, scaled points -> features), and the
// permutohedral table gradient.  ngm/positional_encodings.py: NeRF :245-272, Fourier :197-212,
// Triplane :132-161, permutohedral wrapper :19-66.  The training path (autograd.py) evaluates the MLP
// with library GEMMs around these; the render kernels encode in-line with the same device functions.
#include <cuda_fp16.h>

#include "common.cuh"
#include "encodings.cuh"

namespace ngm {

namespace {

__global__ void __launch_bounds__(256) encode_fwd_kernel(NgmEncodeArgs a) {
  const NgmFieldDesc& fd = a.field;
  const int E = fd.dim_encoding;
  const long long n = a.points_per_field;
  if (fd.encoding == NGM_ENC_PERMUTO) {
    const int L = fd.permuto_levels, F = fd.permuto_feats;
    const long long total = (long long)a.num_fields * n * (L + 1);  // L lattice levels + 1 "concat points" item
    const size_t level_elems = ((size_t)1 << fd.permuto_log2_capacity) * F;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int l = (int)(idx % (L + 1));
      const long long pt = idx / (L + 1);
      const long long f = pt / n;
      const long long slot = a.field_slots ? a.field_slots[f] : f;
      const float* src = a.points + pt * 3;
      const float x[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
      float* dst = a.out + pt * E;
      if (l < L) {
        permuto_level<8>(x, fd.enc_param0 + slot * fd.enc_param0_stride + l * level_elems,
                         fd.enc_param1 + slot * fd.enc_param1_stride + l * 3, fd.permuto_scale + l * 3,
                         fd.permuto_log2_capacity, F, dst + l * F);
      } else if (fd.permuto_concat_points) {
        for (int c = 0; c < 3; ++c) dst[L * F + c] = x[c] * fd.permuto_concat_scaling;
      }
    }
    return;
  }
  const long long total = (long long)a.num_fields * n * E;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % E);
    const long long pt = idx / E;
    const long long f = pt / n;
    const long long slot = a.field_slots ? a.field_slots[f] : f;
    const float* src = a.points + pt * 3;
    const float x[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    const float* ep = fd.enc_param0 ? fd.enc_param0 + slot * fd.enc_param0_stride : nullptr;
    float v;
    if (fd.encoding == NGM_ENC_NERF) v = nerf_feature(x, c, fd.nerf_num_octaves, fd.nerf_start_octave);
    else if (fd.encoding == NGM_ENC_FOURIER) v = fourier_feature(x, c, ep, fd.fourier_raw_coords);
    else v = triplane_feature(x, c, ep, fd.triplane_resolution, fd.triplane_components, fd.triplane_mode);
    a.out[idx] = v;
  }
}

// permutohedral: d lattice_values; one thread per (field, point, level)
__global__ void __launch_bounds__(256) permuto_bwd_kernel(NgmEncodeArgs a) {
  const NgmFieldDesc& fd = a.field;
  const int E = fd.dim_encoding, L = fd.permuto_levels, F = fd.permuto_feats;
  const long long n = a.points_per_field;
  const long long total = (long long)a.num_fields * n * L;
  const size_t level_elems = ((size_t)1 << fd.permuto_log2_capacity) * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(idx % L);
    const long long pt = idx / L;
    const long long f = pt / n;
    const long long slot = a.field_slots ? a.field_slots[f] : f;
    const float* src = a.points + pt * 3;
    const float x[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    float g[8];
    bool any = false;
    for (int i = 0; i < F; ++i) {
      g[i] = __ldg(a.d_out + pt * E + l * F + i);
      any |= g[i] != 0.0f;
    }
    if (!any) continue;
    permuto_level_bwd<8>(x, a.d_param0 + slot * fd.enc_param0_stride + l * level_elems,
                         fd.enc_param1 + slot * fd.enc_param1_stride + l * 3, fd.permuto_scale + l * 3,
                         fd.permuto_log2_capacity, F, g);
  }
}

// Permutohedral A-operand rows for the tcgen05 renderer: world point -> field-local -> 2 features per level -> fp16,
// one thread per (point, group of four levels), so that the 4 x 4 table gathers of a thread are independent and the
// whole GPU's warps hide the L2 latency (inside the persistent MMA kernel only 8 warps per SM could).
// Row layout = the kernel's A operand: EP halves per point (levels, optional raw points, zero padding).
// Same rows in fp32 (E floats per point) for the reference-arithmetic field kernel: the FFMA kernel then reads its
// layer-0 input instead of gathering with 256 threads per SM.
__global__ void __launch_bounds__(256) permuto_rows_f32_kernel(PermutoRowsArgs a) {
  const NgmFieldDesc& fd = a.field;
  const int L = fd.permuto_levels, groups = (L + 3) / 4, E = fd.dim_encoding;
  const size_t level_elems = ((size_t)1 << fd.permuto_log2_capacity) * 2;
  const long long total = a.num_points * groups;
  float* out = reinterpret_cast<float*>(a.out);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / a.num_points);  // lanes = consecutive points of one level group (see the fp16 kernel)
    const long long pt = idx - (long long)g * a.num_points;
    long long f = pt / a.points_per_field, src_pt = pt;
    if (a.pair_field) {
      f = __ldg(a.pair_field + pt);
      if (f < 0) continue;
      src_pt = pt / a.knn_k;
    }
    const long long slot = a.field_slots ? a.field_slots[f] : f;
    const long long pose = a.pair_field ? f : slot;
    const float* src = a.points_world + src_pt * 3;
    float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
    if (a.positions) {
      const float* c = a.positions + pose * 3;
      const float* q = a.orientations + pose * 4;
      x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
      x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
    }
    x = scale_local(x, a.scale_mode, a.field_radius);
    const float xs[3] = {x.x, x.y, x.z};
    const float* table = fd.enc_param0 + slot * fd.enc_param0_stride;
    const float* shift = fd.enc_param1 + slot * fd.enc_param1_stride;
    float* row = out + pt * E;
    for (int i = 0; i < 4 && 4 * g + i < L; ++i) {
      const int l = 4 * g + i;
      float f2[2];
      permuto_level<2>(xs, table + (size_t)l * level_elems, shift + l * 3, fd.permuto_scale + l * 3,
                       fd.permuto_log2_capacity, 2, f2);
      row[2 * l] = f2[0];
      row[2 * l + 1] = f2[1];
    }
    if (g == groups - 1 && fd.permuto_concat_points)
      for (int c = 0; c < 3; ++c) row[2 * L + c] = xs[c] * fd.permuto_concat_scaling;
  }
}

__global__ void __launch_bounds__(256) permuto_rows_half_kernel(PermutoRowsArgs a) {
  const NgmFieldDesc& fd = a.field;
  const int L = fd.permuto_levels, words = a.EP / 2;
  const size_t level_elems = ((size_t)1 << fd.permuto_log2_capacity) * 2;
  // One thread per point, all levels in turn (4 per step, one 16-byte store each): the lanes of a warp are
  // neighbouring samples of a ray working on the SAME level, so on the coarse levels they hit the same 32-byte sectors,
  // and the world -> local transform, the index arithmetic and the row address are paid once per point instead of
  // once per (point, 4 levels) -- the encoder is bound by instruction issue (ncu: 82 % issue active).  The four
  // 16-byte stores of a row follow each other closely, so L2 merges them into whole sectors.
  for (long long pt = blockIdx.x * (long long)blockDim.x + threadIdx.x; pt < a.num_points;
       pt += (long long)gridDim.x * blockDim.x) {
    long long f, src_pt = pt;
    if (a.pair_field) {
      f = __ldg(a.pair_field + pt);
      if (f < 0) continue;
      src_pt = pt / a.knn_k;
    } else {
      f = a.num_points < (1ll << 31) ? (long long)((unsigned)pt / (unsigned)a.points_per_field) : pt / a.points_per_field;
    }
    const long long slot = a.field_slots ? a.field_slots[f] : f;
    const long long pose = a.pair_field ? f : slot;
    const float* src = a.points_world + src_pt * 3;
    float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
    const float* c = a.positions + pose * 3;
    const float* q = a.orientations + pose * 4;
    x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
    x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
    x = scale_local(x, a.scale_mode, a.field_radius);
    const float xs[3] = {x.x, x.y, x.z};
    const float* table = fd.enc_param0 + slot * fd.enc_param0_stride;
    const float* shift = fd.enc_param1 + slot * fd.enc_param1_stride;
    uint32_t* row = a.out + pt * words;
    const bool vec_ok = (words & 3) == 0;
#pragma unroll 1
    for (int l0 = 0; l0 < L; l0 += 4) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int l = l0 + i;
        if (l < L) {
          float f2[2];
          permuto_level<2>(xs, table + (size_t)l * level_elems, shift + l * 3, fd.permuto_scale + l * 3,
                           fd.permuto_log2_capacity, 2, f2);
          const __half2 h = __floats2half2_rn(f2[0], f2[1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      if (l0 + 4 <= L && vec_ok) {
        *reinterpret_cast<uint4*>(row + l0) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        for (int i = 0; i < 4 && l0 + i < L; ++i) row[l0 + i] = w[i];
      }
    }
    // raw points (concat_points) and the zero padding up to the K multiple of 16
    const float cs = fd.permuto_concat_scaling;
    for (int k = L; k < words; ++k) {
      __half2 h = __floats2half2_rn(0.f, 0.f);
      if (fd.permuto_concat_points && k == L) h = __floats2half2_rn(x.x * cs, x.y * cs);
      if (fd.permuto_concat_points && k == L + 1) h = __floats2half2_rn(x.z * cs, 0.f);
      row[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
}

unsigned grid_for(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int launch_encode_fwd(const NgmEncodeArgs& a, cudaStream_t stream) {
  const NgmFieldDesc& fd = a.field;
  const long long per_point = fd.encoding == NGM_ENC_PERMUTO ? fd.permuto_levels + 1 : fd.dim_encoding;
  encode_fwd_kernel<<<grid_for((long long)a.num_fields * a.points_per_field * per_point), 256, 0, stream>>>(a);
  return check_launch("encode_fwd_kernel");
}

// fp16 A-operand rows for the encodings whose features are independent functions of the point (NeRF, Fourier,
// Triplane): one thread per (point, 8 features).  This is what puts Fourier and Triplane fields on the tcgen05
// renderer: the fused kernel reads these rows like the permutohedral ones.
__global__ void __launch_bounds__(256) feature_rows_half_kernel(PermutoRowsArgs a) {
  const NgmFieldDesc& fd = a.field;
  const int E = fd.dim_encoding, groups = a.EP / 8;
  const long long total = a.num_points * groups;
  // Lane mapping: consecutive lanes = consecutive points (neighbouring samples of a ray) of ONE group of 4 levels.
  // On the coarse levels neighbouring samples fall into the same or adjacent simplices, so the lanes of a gather
  // instruction hit the same 32-byte sectors and the request count drops; with (point, group) interleaved across
  // lanes every lane of an instruction reads a different level's table (ncu: 2.1 L2 sectors per vertex).
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / a.num_points);
    const long long pt = idx - (long long)g * a.num_points;
    long long f = pt / a.points_per_field, src_pt = pt;
    if (a.pair_field) {
      f = __ldg(a.pair_field + pt);
      if (f < 0) continue;
      src_pt = pt / a.knn_k;
    }
    const long long slot = a.field_slots ? a.field_slots[f] : f;
    const long long pose = a.pair_field ? f : slot;
    const float* src = a.points_world + src_pt * 3;
    float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
    if (a.positions) {
      const float* c = a.positions + pose * 3;
      const float* q = a.orientations + pose * 4;
      x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
      x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
    }
    x = scale_local(x, a.scale_mode, a.field_radius);
    const float xs[3] = {x.x, x.y, x.z};
    const float* ep = fd.enc_param0 ? fd.enc_param0 + slot * fd.enc_param0_stride : nullptr;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = 8 * g + i;
      v[i] = 0.0f;
      if (c < E) {
        if (fd.encoding == NGM_ENC_NERF) v[i] = nerf_feature(xs, c, fd.nerf_num_octaves, fd.nerf_start_octave);
        else if (fd.encoding == NGM_ENC_FOURIER) v[i] = fourier_feature(xs, c, ep, fd.fourier_raw_coords);
        else v[i] = triplane_feature(xs, c, ep, fd.triplane_resolution, fd.triplane_components, fd.triplane_mode);
      }
    }
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(a.out + pt * (a.EP / 2) + 4 * g) = *reinterpret_cast<const uint4*>(h);
  }
}

int launch_permuto_rows_f32(const PermutoRowsArgs& a, cudaStream_t stream) {
  if (a.num_points == 0) return NGM_OK;
  const long long items = a.num_points * ((a.field.permuto_levels + 3) / 4);
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 64;
  permuto_rows_f32_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(a);
  return check_launch("permuto_rows_f32_kernel");
}

int launch_permuto_rows_half(const PermutoRowsArgs& a, cudaStream_t stream) {
  if (a.num_points == 0) return NGM_OK;
  if (a.field.encoding != NGM_ENC_PERMUTO) {
    const long long items = a.num_points * (a.EP / 8);
    long long blocks = (items + 255) / 256;
    const long long cap = (long long)num_sms() * 64;
    feature_rows_half_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(a);
    return check_launch("feature_rows_half_kernel");
  }
  const long long items = a.num_points;
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 64;  // 8 resident CTAs per SM x 8 rounds, then grid-stride
  permuto_rows_half_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(a);
  return check_launch("permuto_rows_half_kernel");
}

int launch_encode_bwd(const NgmEncodeArgs& a, cudaStream_t stream) {
  if (a.field.encoding != NGM_ENC_PERMUTO) {
    set_error("ngm_encode_bwd: only the permutohedral table gradient is a CUDA kernel (nerf has no parameters; "
              "fourier / triplane gradients come from the PyTorch expressions in autograd.py)");
    return NGM_ERR_UNSUPPORTED;
  }
  permuto_bwd_kernel<<<grid_for((long long)a.num_fields * a.points_per_field * a.field.permuto_levels), 256, 0, stream>>>(a);
  return check_launch("permuto_bwd_kernel");
}

}  // namespace ngm
