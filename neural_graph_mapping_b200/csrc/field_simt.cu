// fp32 field evaluation (reference arithmetic): world->local, encoding, MLP.
// Replaces NeuralFieldSet.forward vmap branch (ngm/models.py:329-345) + NeuralField.forward
// (ngm/models.py:143-182) for every encoding and skip mode, in plain fp32 FFMA.  This is the
// exact-parity path (and the only path for shapes the tcgen05 kernel does not cover); the fp16
// tensor-core path lives in field_tc.cu.
//
// One CTA = 128 points of one field, 256 threads: thread t owns point t%128 and the output
// half t/128.  Activations stay in shared memory ([point][channel], row stride chosen so that
// 128-bit row reads are bank-conflict free); the layer's weights stream through shared memory
// in 128(n) x 16(k) chunks; each thread keeps up to 64 accumulators in registers and consumes
// a chunk with 4 activation LDS.128 + 4 broadcast weight LDS.128 per 16 FFMA.
#include "common.cuh"
#include "encodings.cuh"

namespace ngm {

namespace {

constexpr int TP = 128;        // points per tile
constexpr int NTHREADS = 256;
constexpr int KC = 16;         // k-chunk
// accumulators per thread NT (template parameter): 64 -> 128 outputs per pass (wide MLPs, 148 registers, one CTA
// per SM); 16 -> 32 outputs per pass for MLPs of width <= 32 such as the reference's default field
// (neural_graph_map.yaml:15-17): both thread halves busy and three CTAs per SM.  Same summation order either way.
constexpr int NT_WIDE = 64, NT_NARROW = 16;

__host__ __device__ inline int row_stride(int width) {
  int s = (width + KC - 1) / KC * KC;  // readable in whole 16-float chunks
  s += 4;
  if (((s / 4) & 1) == 0) s += 4;      // (stride/4) odd -> conflict-free LDS.128 by rows
  return s;
}

struct SimtParams {
  NgmFieldDesc fd;
  const float* points;
  const float* positions;
  const float* orientations;
  const long long* field_slots;
  float* out;
  long long points_per_field;
  long long tiles_per_field;
  long long total_tiles;
  float field_radius;
  int scale_mode;
  int act_stride, enc_stride;
  // gather mode (kNN path, models.py:386-396): rows are (point, neighbour) entries bucketed by field
  const int* entries;        // [sum counts]  entry = point * K + k
  const int* entry_offsets;  // [F + 1]
  const int* tile_offsets;   // [F + 1]  tiles of 128 entries per field
  int knn_k;
  int num_fields;
  const float* enc_rows;     // precoded layer-0 input (num_fields * points_per_field, E), or nullptr: encode in-kernel
};

__device__ __forceinline__ void encode_tile(const SimtParams& p, long long slot, const float (*xs)[4],
                                            float* dst, int dst_stride) {
  const NgmFieldDesc& fd = p.fd;
  const int E = fd.dim_encoding;
  const int Epad = (E + KC - 1) / KC * KC;
  if (fd.encoding == NGM_ENC_PERMUTO) {
    const int L = fd.permuto_levels, F = fd.permuto_feats;
    const size_t level_elems = ((size_t)1 << fd.permuto_log2_capacity) * F;
    const float* table = fd.enc_param0 + slot * fd.enc_param0_stride;
    const float* shift = fd.enc_param1 + slot * fd.enc_param1_stride;
    for (int idx = threadIdx.x; idx < TP * L; idx += NTHREADS) {
      const int pt = idx / L, l = idx - pt * L;
      permuto_level<8>(xs[pt], table + l * level_elems, shift + l * 3, fd.permuto_scale + l * 3,
                       fd.permuto_log2_capacity, F, dst + pt * dst_stride + l * F);
    }
    const int base = L * F;
    for (int idx = threadIdx.x; idx < TP * (Epad - base); idx += NTHREADS) {
      const int pt = idx / (Epad - base), c = base + idx % (Epad - base);
      dst[pt * dst_stride + c] = (fd.permuto_concat_points && c < base + 3)
                                     ? xs[pt][c - base] * fd.permuto_concat_scaling : 0.0f;
    }
    return;
  }
  const float* ep = fd.enc_param0 ? fd.enc_param0 + slot * fd.enc_param0_stride : nullptr;
  for (int idx = threadIdx.x; idx < TP * Epad; idx += NTHREADS) {
    const int pt = idx / Epad, c = idx - pt * Epad;
    float v = 0.0f;
    if (c < E) {
      if (fd.encoding == NGM_ENC_NERF) v = nerf_feature(xs[pt], c, fd.nerf_num_octaves, fd.nerf_start_octave);
      else if (fd.encoding == NGM_ENC_FOURIER) v = fourier_feature(xs[pt], c, ep, fd.fourier_raw_coords);
      else v = triplane_feature(xs[pt], c, ep, fd.triplane_resolution, fd.triplane_components, fd.triplane_mode);
    }
    dst[pt * dst_stride + c] = v;
  }
}

template <int NT>
__global__ void __launch_bounds__(NTHREADS, NT <= 16 ? 3 : 1) field_fwd_simt_kernel(SimtParams p) {
  constexpr int NPASS = 2 * NT;
  extern __shared__ __align__(16) float smem[];
  const NgmFieldDesc& fd = p.fd;
  const int E = fd.dim_encoding, W = fd.dim_mlp_out, L = fd.num_layers;
  const int AS = p.act_stride, ES = p.enc_stride;
  float(*xs)[4] = reinterpret_cast<float(*)[4]>(smem);
  float* wsm = smem + TP * 4;                 // [NPASS][KC]
  float* act0 = wsm + NPASS * KC;
  float* act1 = act0 + TP * AS;
  float* encs = act1 + TP * AS;               // only when skip_mode != no (ES > 0)
  const bool keep_enc = fd.skip_mode != NGM_SKIP_NO;

  const int tid = threadIdx.x;
  const int pt = tid & (TP - 1), half = tid >> 7;

  // zero both activation buffers once: the 16-float read chunks may extend past a layer's width
  for (int idx = tid; idx < 2 * TP * AS; idx += NTHREADS) act0[idx] = 0.0f;
  __syncthreads();

  const bool gather = p.entries != nullptr;
  const long long total_tiles = gather ? (long long)p.tile_offsets[p.num_fields] : p.total_tiles;
  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    long long f, p0;
    int ent_base = 0, ent_cnt = 0;
    if (gather) {
      int lo = 0, hi = p.num_fields;  // largest f with tile_offsets[f] <= tile
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.tile_offsets[mid] <= tile) lo = mid; else hi = mid;
      }
      f = lo;
      ent_base = p.entry_offsets[f] + (int)(tile - p.tile_offsets[f]) * TP;
      ent_cnt = min(TP, p.entry_offsets[f + 1] - ent_base);
      p0 = 0;
    } else {
      f = tile / p.tiles_per_field;
      p0 = (tile - f * p.tiles_per_field) * TP;
    }
    const long long slot = p.field_slots ? p.field_slots[f] : f;   // row in the parameter tables
    const long long pose_slot = gather ? f : slot;                 // row in positions / orientations

    // ---- local coordinates (models.py:331-339) ----
    if (tid < TP) {
      float3 x = make_float3(0.f, 0.f, 0.f);
      const long long gp = p0 + tid;
      const bool row_ok = gather ? tid < ent_cnt : gp < p.points_per_field;
      if (row_ok) {
        const float* src = gather ? p.points + (long long)(__ldg(p.entries + ent_base + tid) / p.knn_k) * 3
                                  : p.points + (f * p.points_per_field + gp) * 3;
        x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
        if (p.positions) {
          const float* c = p.positions + pose_slot * 3;
          const float* q = p.orientations + pose_slot * 4;
          x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
          x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
        }
        x = scale_local(x, p.scale_mode, p.field_radius);
      }
      xs[tid][0] = x.x; xs[tid][1] = x.y; xs[tid][2] = x.z; xs[tid][3] = 0.f;
    }
    __syncthreads();
    // ---- encoding -> act0 (and a persistent copy for the skip connections) ----
    if (p.enc_rows) {  // rows produced by the whole-GPU row encoder (encode.cu)
      const int Epad = (E + KC - 1) / KC * KC;
      for (int idx = tid; idx < TP * Epad; idx += NTHREADS) {
        const int q = idx / Epad, c = idx - q * Epad;
        const long long gp = p0 + q;
        float v = 0.0f;
        if (c < E) {
          if (gather) {
            if (q < ent_cnt) v = __ldg(p.enc_rows + (long long)__ldg(p.entries + ent_base + q) * E + c);
          } else if (gp < p.points_per_field) {
            v = __ldg(p.enc_rows + (f * p.points_per_field + gp) * E + c);
          }
        }
        act0[q * AS + c] = v;
      }
    } else {
      encode_tile(p, slot, xs, act0, AS);
    }
    __syncthreads();
    if (keep_enc) {
      for (int idx = tid; idx < TP * E; idx += NTHREADS) {
        const int q = idx / E, c = idx - q * E;
        encs[q * ES + c] = act0[q * AS + c];
      }
    }
    float* in = act0;
    float* outb = act1;

    // ---- linears (models.py:148-182) ----
    for (int layer = 0; layer <= L; ++layer) {
      const int K = layer == 0 ? E : (fd.skip_mode == NGM_SKIP_CONCAT ? W + E : W);
      const int N = layer == L ? fd.dim_out : W;
      const float* Wg = fd.weights[layer] + slot * fd.weight_stride[layer];
      const float* Bg = fd.biases[layer] + slot * fd.bias_stride[layer];
      const bool last = layer == L;
      for (int nb = 0; nb < N; nb += NPASS) {
        const int ncount = min(NPASS, N - nb);
        const int jmax = min(NT, ncount - half * NT);  // outputs of this thread in this pass (may be <= 0)
        float acc[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] = 0.0f;
        for (int k0 = 0; k0 < K; k0 += KC) {
          __syncthreads();  // previous chunk consumed / activations of previous layer complete
          {                 // stage W[nb .. nb+128) x [k0 .. k0+16) -> wsm[n][kk], zero padded
            const int n = tid >> 1, kk0 = (tid & 1) * 8;
            if (n < NPASS) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int k = k0 + kk0 + i;
                v[i] = (n < ncount && k < K) ? __ldg(Wg + (size_t)(nb + n) * K + k) : 0.0f;
              }
              float4* d = reinterpret_cast<float4*>(wsm + n * KC + kk0);
              d[0] = make_float4(v[0], v[1], v[2], v[3]);
              d[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
          __syncthreads();
          if (jmax > 0) {
            float a[KC];
            const float4* ar = reinterpret_cast<const float4*>(in + pt * AS + k0);
#pragma unroll
            for (int i = 0; i < KC / 4; ++i) {
              const float4 t = ar[i];
              a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
            }
            const float4* wr = reinterpret_cast<const float4*>(wsm + (half * NT) * KC);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              if (j < jmax) {
                float s = acc[j];
#pragma unroll
                for (int i = 0; i < KC / 4; ++i) {
                  const float4 w = wr[j * (KC / 4) + i];
                  s = fmaf(a[4 * i], w.x, s);
                  s = fmaf(a[4 * i + 1], w.y, s);
                  s = fmaf(a[4 * i + 2], w.z, s);
                  s = fmaf(a[4 * i + 3], w.w, s);
                }
                acc[j] = s;
              }
            }
          }
        }
        // ---- epilogue of this pass: bias, ReLU, skip, store ----
        if (jmax > 0) {
          const float alpha = (fd.skip_mode == NGM_SKIP_REZERO && !last)
                                  ? __ldg(fd.rezero + slot * fd.rezero_stride + layer) : 0.0f;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (j < jmax) {
              const int n = nb + half * NT + j;
              float v = acc[j] + __ldg(Bg + n);
              if (last) {
                if (gather) {
                  if (pt < ent_cnt) p.out[(long long)__ldg(p.entries + ent_base + pt) * fd.dim_out + n] = v;
                } else {
                  const long long gp = p0 + pt;
                  if (gp < p.points_per_field) p.out[(f * p.points_per_field + gp) * fd.dim_out + n] = v;
                }
              } else {
                v = fmaxf(v, 0.0f);
                if (fd.skip_mode == NGM_SKIP_ADD) {
                  if (n < E) v += encs[pt * ES + n];
                } else if (fd.skip_mode == NGM_SKIP_REZERO) {
                  if (layer == 0) v = n < E ? alpha * v + encs[pt * ES + n] : alpha * v;
                  else v = alpha * v + in[pt * AS + n];
                }
                outb[pt * AS + n] = v;
              }
            }
          }
        }
      }
      if (!last) {
        if (fd.skip_mode == NGM_SKIP_CONCAT) {  // outs = cat(outs, enc)  (models.py:160-161)
          for (int idx = tid; idx < TP * E; idx += NTHREADS) {
            const int q = idx / E, c = idx - q * E;
            outb[q * AS + W + c] = encs[q * ES + c];
          }
        }
        float* t = in; in = outb; outb = t;
      }
    }
    __syncthreads();  // before the next tile overwrites xs / act0
  }
}

}  // namespace

size_t field_simt_smem_bytes(const NgmFieldDesc& fd, int* act_stride, int* enc_stride) {
  const int E = fd.dim_encoding, W = fd.dim_mlp_out;
  const int widest = (fd.skip_mode == NGM_SKIP_CONCAT ? W + E : (W > E ? W : E));
  const int AS = row_stride(widest > fd.dim_out ? widest : fd.dim_out);
  const int ES = fd.skip_mode != NGM_SKIP_NO ? row_stride(E) : 0;
  if (act_stride) *act_stride = AS;
  if (enc_stride) *enc_stride = ES;
  return sizeof(float) * ((size_t)TP * 4 + 2 * NT_WIDE * KC + 2 * (size_t)TP * AS + (size_t)TP * ES);
}

int launch_field_fwd_simt_gather(const NgmFieldFwdArgs& a, const int* entries, const int* entry_offsets,
                                 const int* tile_offsets, int knn_k, long long max_tiles, cudaStream_t stream);

int launch_permuto_rows_f32(const PermutoRowsArgs& a, cudaStream_t stream);

// permutohedral rows for the fp32 kernel: worth it when the caller provides the workspace (E floats per point)
size_t field_simt_workspace_bytes(const NgmFieldFwdArgs& a) {
  if (a.field.encoding != NGM_ENC_PERMUTO || a.field.permuto_feats != 2) return 0;
  return (size_t)a.num_fields * (size_t)a.points_per_field * (size_t)a.field.dim_encoding * sizeof(float);
}

int launch_field_fwd_simt(const NgmFieldFwdArgs& a, cudaStream_t stream) {
  const size_t need = field_simt_workspace_bytes(a);
  if (need && a.workspace && a.workspace_bytes >= need) {
    PermutoRowsArgs e{};
    e.field = a.field;
    e.points_world = a.points;
    e.positions = a.positions; e.orientations = a.orientations;
    e.field_slots = reinterpret_cast<const long long*>(a.field_slots);
    e.out = static_cast<uint32_t*>(a.workspace);
    e.num_points = (long long)a.num_fields * a.points_per_field;
    e.points_per_field = a.points_per_field;
    e.field_radius = a.field_radius;
    e.scale_mode = a.scale_mode;
    if (int rc = launch_permuto_rows_f32(e, stream)) return rc;
    return launch_field_fwd_simt_gather(a, nullptr, nullptr, nullptr, -1, 0, stream);  // knn_k = -1: rows precoded
  }
  return launch_field_fwd_simt_gather(a, nullptr, nullptr, nullptr, 1, 0, stream);
}

// `entries` != nullptr: gather mode.  a.points = (num_points,3) world points, a.positions/orientations = pose of
// field f at row f, a.field_slots = parameter row of field f, a.out = (num_points*K, dim_out) per-entry outputs.
int launch_field_fwd_simt_gather(const NgmFieldFwdArgs& a, const int* entries, const int* entry_offsets,
                                 const int* tile_offsets, int knn_k, long long max_tiles, cudaStream_t stream) {
  if (a.num_fields == 0 || (!entries && a.points_per_field == 0)) return NGM_OK;
  SimtParams p;
  p.entries = entries;
  p.entry_offsets = entry_offsets;
  p.tile_offsets = tile_offsets;
  p.knn_k = knn_k;
  // precoded layer-0 rows: knn_k == -1 (dense call) or a.workspace given in gather mode (fp32 call from knn.cu)
  p.enc_rows = ((!entries && knn_k == -1) || (entries && a.precision == NGM_PREC_FP32 && a.workspace))
                   ? static_cast<const float*>(a.workspace) : nullptr;
  p.num_fields = a.num_fields;
  p.fd = a.field;
  p.points = a.points;
  p.positions = a.positions;
  p.orientations = a.orientations;
  p.field_slots = reinterpret_cast<const long long*>(a.field_slots);
  p.out = a.out;
  p.points_per_field = a.points_per_field;
  p.tiles_per_field = (a.points_per_field + TP - 1) / TP;
  p.field_radius = a.field_radius;
  p.scale_mode = a.scale_mode;
  const size_t smem = field_simt_smem_bytes(a.field, &p.act_stride, &p.enc_stride);
  NGM_UNSUPPORTED(smem > 227 * 1024,
                  "fp32 field kernel needs %zu B shared memory (> 227 KB): dim_mlp_out=%d dim_encoding=%d too wide",
                  smem, a.field.dim_mlp_out, a.field.dim_encoding);
  p.total_tiles = entries ? max_tiles : p.tiles_per_field * a.num_fields;
  if (p.total_tiles <= 0) return NGM_OK;
  const long long cap = (long long)num_sms() * 64;
  const unsigned grid = (unsigned)(p.total_tiles < cap ? p.total_tiles : cap);
  const bool narrow = a.field.dim_mlp_out <= 2 * NT_NARROW && a.field.dim_out <= 2 * NT_NARROW;
  if (narrow) {
    NGM_CUDA(cudaFuncSetAttribute(field_fwd_simt_kernel<NT_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    field_fwd_simt_kernel<NT_NARROW><<<grid, NTHREADS, smem, stream>>>(p);
  } else {
    NGM_CUDA(cudaFuncSetAttribute(field_fwd_simt_kernel<NT_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    field_fwd_simt_kernel<NT_WIDE><<<grid, NTHREADS, smem, stream>>>(p);
  }
  return check_launch("field_fwd_simt_kernel");
}

}  // namespace ngm
