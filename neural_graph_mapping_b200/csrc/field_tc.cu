// fp16-operand tensor-core path (tcgen05 / TMEM / TMA-engine bulk copies), sm_100a only.
//
//   * field_fwd (MODE 1): points read from HBM -> world->local -> NeRF encoding -> per-field MLP on
//     tcgen05 -> raw outputs written back (ngm/models.py:329-345).  THE PRODUCT PATH of a render:
//     sample_rays_kernel -> this kernel -> composite_staged_kernel (0.60 of the burst cuBLAS rate on the
//     headline keyframe); also the dense field evaluation and, in gather mode, the kNN path.
//   * render_fused (MODE 0, opt-in: NGM_RENDER_FUSED=1): ONE kernel per render batch = sampler ->
//     world->local -> NeRF encoding -> MLP -> front-to-back alpha composite (ngm/run_mapping.py:440-666,
//     use_vmap=True), ~60 B/ray in + 36 B/ray out.  Measured 13-22 % slower than the three stage kernels
//     (DESIGN.md 5): the in-kernel sampler and compositor lengthen the slot threads' dependency chain.
//
// Design (one persistent CTA per SM, 576 threads):
//   warps 0 / 1     : MMA issuers of tile slot 0 / 1 (one elected lane issues tcgen05.mma after a
//                     blocking mbarrier wait); warp 0 also loads the field's pre-swizzled fp16 weight
//                     image into shared memory with cp.async.bulk (TMA engine)
//   warps 2-9/10-17 : two "tile slots".  A slot's 256 threads own the 128 rows (= sample points)
//                     of one tile, two threads per row (thread <-> TMEM lane, each thread one half
//                     of the columns).  They generate the sample, encode it,
//                     tcgen05.st the fp16 features as the A operand into TMEM, and after every
//                     layer tcgen05.ld the fp32 accumulator, apply bias+ReLU (packed half2),
//                     and tcgen05.st the next A operand.  Activations never touch shared memory
//                     or HBM; shared memory only feeds the B operand (weights), so the MMA reads
//                     64 B/clk of shared memory instead of 128.
//   While slot 0 is in an epilogue the tensor pipe runs slot 1's layer and vice versa.
// TMEM (512 columns): slot s uses columns [256 s, 256 s + 128) for the fp32 accumulator D and
// [256 s + 128, 256 s + 192) for the fp16 A operand (K <= 128).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "encodings.cuh"
#include "tc_ptx.cuh"
#include "field_tc_common.cuh"

namespace ngm {

namespace {

// 2 MMA issuers + 2 x 8 slot warps (+ 1 ray-parameter producer in the fused renderer)
constexpr int threads_of(int mode) { return mode == 0 ? 608 : 576; }
constexpr int kTmemCols = 512;
constexpr int kSlotCols = 256;
constexpr int kACol = 128;
constexpr int kStageCol = 192;  // layer-0 A operand of the slot's NEXT tile (<= 32 columns: K <= 64)

// ---- main kernel ---------------------------------------------------------------------------------
struct TcParams {
  TcImage im;
  const uint8_t* images;
  int images_by_slot;  // persistent images (NgmFieldDesc.packed_weights): indexed by table row, not by field of the call
  int num_fields;
  int E, EP, W, L, dim_out;
  int nerf_start;
  // permutohedral encoding (OCT == 0): per-field table (levels, capacity, 2) fp32, read through L2
  const float* pm_table;
  const float* pm_shift;
  const float* pm_scale;
  long long pm_table_stride, pm_shift_stride;
  int pm_levels, pm_log2cap, pm_concat;
  float pm_concat_scale;
  // field poses
  const float* positions;
  const float* orientations;
  const long long* field_slots;
  int scale_mode;
  float field_radius;
  // tiles
  long long tiles_per_field, total_tiles;
  // MODE 1 (field fwd)
  const float* points;
  long long points_per_field;
  float* out;
  const __half* raw_a;  // layer-0 A operand given directly, (num_fields*points_per_field, EP): debug GEMM, and the
                        // pre-encoded permutohedral rows of the renderer (then raw_dist / raw_depth come with it)
  // MODE 1 gather mode (kNN path, ngm/models.py:386-396): rows are (point, neighbour) entries bucketed by field;
  // the tile count per field is data dependent and stays on the device (knn.cu)
  const int* entries;        // [sum counts]  entry = point * K + k  (also the output row)
  const int* entry_offsets;  // [F + 1]
  const int* tile_offsets;   // [F + 1]  tiles of 128 entries per field
  int knn_k;
  const float* raw_dist;   // MODE 0 with raw_a: sample distances / depths (num_rays, St) of the sampler kernel
  const float* raw_depth;
  // MODE 0 (fused render)
  NgmCamera cam;
  const long long* ijs;
  const float* c2ws;
  const float* near;
  const float* far;
  const float* gt;
  RayJitter jit;
  float near_scalar, far_scalar, range_guided;
  int c2w_per_ray, S, G, St, Sp, sp_shift, rpt;
  float inv_S;  // 1 / S (torch.linspace step)
  long long rays_per_field;
  int geometry_mode, overwrite;
  const int* overwrite_gate;  // device flag of run_mapping.py:494-495 (some near < 0), or nullptr
  float geometry_factor, color_factor, truncation;
  const float* neus_isd;  // (num_fields)
  float* rgbd;
  float* color_var;
  float* depth_var;
  float* term_prob;
  float* freespace;
  uint8_t* freespace_mask;
  float* tsdf;
  uint8_t* tsdf_mask;
  int num_mirrors;  // fused multi-GPU tile exchange (NgmRenderArgs.mirror_delta)
  int mirror_per_ray;
  long long mirror_delta[NGM_MAX_MIRRORS];
  int skip;  // 0 no, 1 "add", 2 "concat" (NgmFieldDesc.skip_mode; rezero stays on the fp32 path)
  int trace;
  int acc16;  // hidden layers accumulate in fp16 inside the tensor core (NGM_TC_ACC16=1): packed accumulator loads
};

// Arrivals on the slot barriers: one per warp (lane 0 after a __syncwarp; each lane has ordered its own tcgen05 / shared
// memory accesses before the warp barrier) instead of one per thread -- 8 arrivals per phase instead of 256.
#ifdef NGM_WARP_ARRIVE
constexpr int kArrivalsPerWarp = 1;
#else
constexpr int kArrivalsPerWarp = 32;
#endif
__device__ __forceinline__ void slot_arrive(uint64_t* bar) {
  if (kArrivalsPerWarp == 1) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(bar);
  } else {
    ptx::mbar_arrive(bar);
  }
}

constexpr int kMaxRaysPerTile = 16;  // Sp >= 8
constexpr int kRayFloats = 12;       // o_local[3], dir_local[3], zscale, near, far, gt, valid, pad

struct Smem {
  uint64_t a0_ready[2];  // layer-0 A operand stored (front-end half) + previous tile's last accumulator drained (compositor half)
  uint64_t a_ready[2];   // hidden-layer A operand stored (all 256 threads of the slot)
  uint64_t d_ready[2];   // accumulator complete (tcgen05.commit)
  uint64_t w_ready;
  uint64_t ray_full[2][4];   // ray parameters of a tile written (producer warp, 32 arrivals)
  uint64_t ray_empty[2][4];  // ... consumed (the 128 compositor-half threads of the slot)
  uint32_t tmem_base;
  uint32_t pad_;
  float ray[2][4][kMaxRaysPerTile][kRayFloats];  // [slot][ring of 4 tiles][ray in tile][param]
  float rowdata[2][3][128][2];  // [slot][ring of 3 tiles][row]{distance, depth}: front end -> compositor
  float sm_x[2][128];      // per slot: front end, depth-guided merge exchange
  float sm_d[2][128];      // per slot: compositor, sample distance (density deltas)
  float sm_g[2][128];      // per slot: geometry after the behind-camera overwrite (neus neighbour)
  float sm_part[2][16][12];  // per slot, per segment (128 / wseg <= 16): 9 partial moments + transmittance leaving the segment
  float comp[2][8][128];     // per slot: deferred compositor inputs of the previous tile {c0,c1,c2,g,d,z,gt,ray_ok} per row
  long long mirror_delta[NGM_MAX_MIRRORS];  // NgmRenderArgs.mirror_delta (fused tile exchange)
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __frcp_rn(1.0f + __expf(-x)); }

// Permutohedral features (2 per level) of lattice levels [l0, l1) of one row -> fp16 -> TMEM A-operand words
// [l0, l1) (one word per level).  The table stays in global memory (L2-resident: 512 KB per field).
__device__ __forceinline__ void encode_permuto_levels(const TcParams& p, uint32_t a_addr, float3 x, long long slot, int l0,
                                                      int l1) {
  const float xs[3] = {x.x, x.y, x.z};
  const size_t level_elems = ((size_t)1 << p.pm_log2cap) * 2;
  const float* table = p.pm_table + slot * p.pm_table_stride;
  const float* shift = p.pm_shift + slot * p.pm_shift_stride;
  for (int l = l0; l < l1; l += 4) {
    if (l + 4 <= l1) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f2[2];
        permuto_level<2>(xs, table + (size_t)(l + i) * level_elems, shift + (l + i) * 3, p.pm_scale + (l + i) * 3,
                         p.pm_log2cap, 2, f2);
        w[i] = ptx::pack_half2(f2[0], f2[1]);
      }
      ptx::tmem_st4(a_addr + l, w);
    } else {
      for (int i = l; i < l1; ++i) {
        float f2[2];
        permuto_level<2>(xs, table + (size_t)i * level_elems, shift + i * 3, p.pm_scale + i * 3, p.pm_log2cap, 2, f2);
        uint32_t w1[1] = {ptx::pack_half2(f2[0], f2[1])};
        ptx::tmem_st1(a_addr + i, w1);
      }
    }
  }
}
// words [levels, EP/2): the optional raw points (concat_points) and the zero padding up to the K multiple of 16
__device__ __forceinline__ void encode_permuto_tail(const TcParams& p, uint32_t a_addr, float3 x) {
  const float cs = p.pm_concat_scale;
  for (int w = p.pm_levels; w < p.EP / 2; ++w) {
    uint32_t v[1] = {0u};
    if (p.pm_concat && w == p.pm_levels) v[0] = ptx::pack_half2(x.x * cs, x.y * cs);
    if (p.pm_concat && w == p.pm_levels + 1) v[0] = ptx::pack_half2(x.z * cs, 0.0f);
    ptx::tmem_st1(a_addr + w, v);
  }
}

// 16 accumulator columns -> relu(x + b) as 8 packed half2 words
__device__ __forceinline__ void cvt16(const uint32_t* v, const uint32_t* bias2, uint32_t* w) {
  const uint4 b0 = *reinterpret_cast<const uint4*>(bias2);
  const uint4 b1 = *reinterpret_cast<const uint4*>(bias2 + 4);
  const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i)
    w[i] = ptx::bias_relu_half2(ptx::pack_half2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), b[i]);
}

// hidden-layer epilogue of one thread: accumulator columns [c0, c0 + n) -> A words [c0/2, (c0+n)/2)
__device__ __forceinline__ void hidden_epilogue(uint32_t d_addr, uint32_t a_addr, const uint32_t* bias2, int c0, int n) {
  int c = c0;
  const int end = c0 + n;
  for (; c + 32 <= end; c += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(d_addr + c, v);
    ptx::tc_wait_ld();
    uint32_t w[16];
    cvt16(v, bias2 + c / 2, w);
    cvt16(v + 16, bias2 + c / 2 + 8, w + 8);
    ptx::tmem_st16(a_addr + c / 2, w);
  }
  if (c < end) {  // 16 columns left
    uint32_t v[16];
    ptx::tmem_ld16(d_addr + c, v);
    ptx::tc_wait_ld();
    uint32_t w[8];
    cvt16(v, bias2 + c / 2, w);
    ptx::tmem_st8(a_addr + c / 2, w);
  }
}

// skip mode "add" (ngm/models.py:162-169): relu(x + b) + encoding on the first E columns.  `enc` = the row's
// pre-encoded fp16 features as half2 words (zero-padded to enc_words), nullptr for a padding row.
__device__ __forceinline__ void add_words(uint32_t* w, const uint32_t* enc, int w0, int n, int enc_words) {
  if (!enc) return;
  for (int i = 0; i < n; i += 4) {
    if (w0 + i >= enc_words) break;  // enc_words is a multiple of 8
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(enc + w0 + i));
    const uint32_t ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&w[i + j]), *reinterpret_cast<const __half2*>(&ev[j]));
      w[i + j] = *reinterpret_cast<const uint32_t*>(&r);
    }
  }
}

__device__ __forceinline__ void hidden_epilogue_add(uint32_t d_addr, uint32_t a_addr, const uint32_t* bias2, int c0, int n,
                                                    const uint32_t* enc, int enc_words) {
  int c = c0;
  const int end = c0 + n;
  for (; c + 16 <= end; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(d_addr + c, v);
    ptx::tc_wait_ld();
    uint32_t w[8];
    cvt16(v, bias2 + c / 2, w);
    add_words(w, enc, c / 2, 8, enc_words);
    ptx::tmem_st8(a_addr + c / 2, w);
  }
}

// the same with 16-bit accumulators: packed loads deliver two columns per register, no fp32 -> fp16 convert
__device__ __forceinline__ void hidden_epilogue_acc16(uint32_t d_addr, uint32_t a_addr, const uint32_t* bias2, int c0, int n) {
  int c = c0;
  const int end = c0 + n;
  // (one packed load of a thread's whole 64-column share -- 32 registers -- was measured slower than two of 32 columns:
  //  2.748 vs 2.713 ms per frame against the fp32-accumulator kernel on the same box, with spills at the 96-register cap)
  for (; c + 32 <= end; c += 32) {  // 32 columns -> 16 packed registers (small register footprint: 96-register cap)
    uint32_t v[16];
    ptx::tmem_ld16_pack16(d_addr + c, v);
    ptx::tc_wait_ld();
    uint32_t w[16];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const uint4 bb = *reinterpret_cast<const uint4*>(bias2 + c / 2 + 4 * q4);
      w[4 * q4] = ptx::bias_relu_half2(v[4 * q4], bb.x);
      w[4 * q4 + 1] = ptx::bias_relu_half2(v[4 * q4 + 1], bb.y);
      w[4 * q4 + 2] = ptx::bias_relu_half2(v[4 * q4 + 2], bb.z);
      w[4 * q4 + 3] = ptx::bias_relu_half2(v[4 * q4 + 3], bb.w);
    }
    ptx::tmem_st16(a_addr + c / 2, w);
  }
  if (c + 16 <= end) {
    uint32_t v[8];
    ptx::tmem_ld8_pack16(d_addr + c, v);
    ptx::tc_wait_ld();
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = ptx::bias_relu_half2(v[i], bias2[c / 2 + i]);
    ptx::tmem_st8(a_addr + c / 2, w);
  }
}

// Per-ray quantities of the fused renderer, computed once per ray (two tiles ahead) instead of per
// sample: the sample point in scaled field-local coordinates is o + d * dir (algebraically the
// reference's  scale(q^-1 (R (dir d) + t - c)), run_mapping.py:547 + models.py:331-339).
__device__ __forceinline__ void compute_ray_params(const TcParams& p, long long f, long long slot, long long r, float* out) {
  const bool ok = r < p.rays_per_field;
  float o[3] = {0.f, 0.f, 0.f}, dl[3] = {0.f, 0.f, 0.f};
  float zs = 0.f, nr = 0.f, fr = 0.f, gt = 0.f;
  if (ok) {
    const long long ray = f * p.rays_per_field + r;
    nr = p.near ? __ldg(p.near + ray) : p.near_scalar;
    fr = p.far ? __ldg(p.far + ray) : p.far_scalar;
    if (p.gt) gt = __ldg(p.gt + ray);
    const longlong2 ij = __ldg(reinterpret_cast<const longlong2*>(p.ijs) + ray);
    const float3 dir = ij_to_direction(ij.x, ij.y, p.cam);
    zs = -dir.z;
    const float* m = p.c2ws + (p.c2w_per_ray ? ray * 16 : 0);
    float mm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) mm[i] = __ldg(m + i);
    float3 dw = make_float3(mm[0] * dir.x + mm[1] * dir.y + mm[2] * dir.z, mm[4] * dir.x + mm[5] * dir.y + mm[6] * dir.z,
                            mm[8] * dir.x + mm[9] * dir.y + mm[10] * dir.z);
    const float* c = p.positions + slot * 3;
    const float* q = p.orientations + slot * 4;
    float3 ow = make_float3(mm[3] - __ldg(c), mm[7] - __ldg(c + 1), mm[11] - __ldg(c + 2));
    const float qw = __ldg(q), qx = __ldg(q + 1), qy = __ldg(q + 2), qz = __ldg(q + 3);
    ow = quat_inv_rotate(qw, qx, qy, qz, ow);
    dw = quat_inv_rotate(qw, qx, qy, qz, dw);
    float sc = 1.0f, sh = 0.0f;
    if (p.scale_mode == NGM_SCALE_UNIT_CUBE) { sc = 1.0f / (2.0f * p.field_radius); sh = 0.5f; }
    else if (p.scale_mode == NGM_SCALE_UNIT_BALL) sc = 1.0f / p.field_radius;
    o[0] = fmaf(ow.x, sc, sh); o[1] = fmaf(ow.y, sc, sh); o[2] = fmaf(ow.z, sc, sh);
    dl[0] = dw.x * sc; dl[1] = dw.y * sc; dl[2] = dw.z * sc;
  }
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
  out[3] = dl[0]; out[4] = dl[1]; out[5] = dl[2];
  out[6] = zs; out[7] = nr; out[8] = fr; out[9] = gt; out[10] = ok ? 1.0f : 0.0f;
  out[11] = __fsub_rn(fr, nr) / (float)p.S;  // stratum width of the coarse set (camera.py:270), once per ray
}

// ---- compositor of one tile slot (the 128 h == 0 threads, one row each, rows of a ray Sp apart) ----
// ngm/run_mapping.py:610-639, 709-799.  The compositor of tile t is DEFERRED: when the last layer of
// tile t completes, the h == 0 threads only pull their row's four outputs out of TMEM into shared
// memory (Smem::comp) and go straight on to the next tile's hidden-layer epilogues; the arithmetic runs
// in two stages slotted into the waits for the next tile's MMAs (stage 1 after the first hidden
// epilogue, stage 2 after the second), exactly like the front end on the other half of the slot.
//   stage 1: occupancy, segmented product scan (transmittance relative to the segment start), the nine
//            weighted moments, a transposed segmented reduction (8 values in 4+2+1(+log) shuffles
//            instead of 8 x log), partial results -> shared memory
//   stage 2: one named barrier, the k == 0 thread of every ray chains its ray's segments
//            (moments are linear in the transmittance entering a segment) and stores the ray.
// wseg = min(Sp, 32) = lanes of one ray inside a warp; rays of 64 / 128 rows span 2 / 4 segments.
// The variances use Sum w (m - x)^2 = Sum w x^2 - m^2 (2 - Sum w): one reduction pass suffices.
__device__ __forceinline__ void comp_stage1(const TcParams& p, Smem& sm, int s, int row, int lane, int barrier_id, int wseg,
                                            long long f, long long tile_in_field) {
  const int St = p.St;
  const int mode = p.geometry_mode;
  const bool drop_last = (mode == NGM_GEOM_DENSITY || mode == NGM_GEOM_NEUS);
  const int Se = drop_last ? St - 1 : St;
  const int k = row & (p.Sp - 1);
  const float c0 = sm.comp[s][0][row], c1 = sm.comp[s][1][row], c2 = sm.comp[s][2][row];
  float g = sm.comp[s][3][row];
  const float d = sm.comp[s][4][row], z = sm.comp[s][5][row], gt = sm.comp[s][6][row];
  const bool ray_ok = sm.comp[s][7][row] != 0.0f;
  const bool valid = ray_ok && k < St;
  const bool has_gt = p.gt != nullptr;
  if (p.overwrite && z < 0.0f && (p.overwrite_gate == nullptr || __ldg(p.overwrite_gate) != 0))
    g = (mode == NGM_GEOM_OCCUPANCY || mode == NGM_GEOM_DENSITY) ? -100.0f : 1.0f;
  if (valid && (p.freespace || p.tsdf)) {
    const long long ray_global = f * p.rays_per_field + tile_in_field * p.rpt + (row >> p.sp_shift);
    const long long idx = ray_global * St + k;
    if (p.freespace) {
      const float thr = has_gt ? (gt - p.truncation) * (gt != 0.0f ? 1.0f : 0.0f) : 0.0f;
      p.freespace[idx] = g * p.truncation;
      p.freespace_mask[idx] = (has_gt && d < thr) ? 1 : 0;
    }
    if (p.tsdf) {
      const float delta = gt - d;
      p.tsdf[idx] = g * p.truncation - delta;
      p.tsdf_mask[idx] = (has_gt && fabsf(delta) < p.truncation && gt != 0.0f) ? 1 : 0;
    }
  }
  float occ = 0.0f;
  if (drop_last) {  // needs the next sample of the same ray
    sm.sm_d[s][row] = d;
    sm.sm_g[s][row] = g;
    ptx::named_bar_sync(barrier_id, 128);
    if (valid && k < Se) {
      if (mode == NGM_GEOM_DENSITY) {
        const float delta = sm.sm_d[s][row + 1] - d;
        occ = 1.0f - __expf(-delta * fmaxf(g, 0.0f));
      } else {
        const float isd_gamma = __ldg(p.neus_isd + f) * p.geometry_factor;
        const float t0 = fast_sigmoid(isd_gamma * g), t1 = fast_sigmoid(isd_gamma * sm.sm_g[s][row + 1]);
        occ = fmaxf(__fdividef(t0 - t1, t0 + 1e-5f), 0.0f);
      }
    }
  } else if (valid) {
    if (mode == NGM_GEOM_NRGBD) {  // 4 s(t) s(-t) = 4u / (1+u)^2, u = exp(-|t|)
      const float u = __expf(-fabsf(p.geometry_factor * g));
      const float r = __frcp_rn(1.0f + u);
      occ = 4.0f * u * r * r;
    } else {
      occ = fast_sigmoid(p.geometry_factor * g);
    }
  }
  // inclusive product scan of (1 - occ) inside the segment
  const int ls = lane & (wseg - 1);
  float incl = 1.0f - occ;
  for (int o = 1; o < wseg; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o, wseg);
    if (ls >= o) incl *= n;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1, wseg);
  if (ls == 0) excl = 1.0f;
  const float wgt = occ * excl;  // weight relative to the transmittance entering this segment
  // transposed reduction of the 8 moments: halve the value set while folding the three top lane bits
  const int h1 = wseg >> 1, h2 = wseg >> 2, h3 = wseg >> 3;
  const bool b1 = (lane & h1) != 0, b2 = (lane & h2) != 0, b3 = (lane & h3) != 0;
  float v[8];
  v[0] = wgt * z; v[1] = wgt * c0; v[2] = wgt * c1; v[3] = wgt * c2;
  v[4] = v[0] * z; v[5] = v[1] * c0; v[6] = v[2] * c1; v[7] = v[3] * c2;
  float a4[4], a2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b1 ? v[i] : v[i + 4], keep = b1 ? v[i + 4] : v[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, h1);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? a4[i] : a4[i + 2], keep = b2 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, h2);
  }
  float r8 = (b3 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b3 ? a2[0] : a2[1], h3);
  float w0 = wgt;
  w0 += __shfl_xor_sync(0xffffffffu, w0, h1);
  w0 += __shfl_xor_sync(0xffffffffu, w0, h2);
  w0 += __shfl_xor_sync(0xffffffffu, w0, h3);
  for (int o = wseg >> 4; o > 0; o >>= 1) {
    r8 += __shfl_xor_sync(0xffffffffu, r8, o);
    w0 += __shfl_xor_sync(0xffffffffu, w0, o);
  }
  // partials of this segment: [0] Sum w, [1] Sum w z, [2..4] Sum w c, [5] Sum w z^2, [6..8] Sum w c^2, [9] transmittance out
  float* part = sm.sm_part[s][row >> (p.sp_shift < 5 ? p.sp_shift : 5)];
  if ((lane & (h3 - 1)) == 0) part[1 + (b1 ? 4 : 0) + (b2 ? 2 : 0) + (b3 ? 1 : 0)] = r8;
  if (ls == 0) part[0] = w0;
  if (ls == wseg - 1) part[9] = incl;
}

// Fused tile exchange (NgmRenderArgs.mirror_delta): the whole CTA repeats rays [r0, r1) of the Prediction arrays --
// the contiguous range its finished field segment has just stored, still in L2 -- at every mirror offset (the
// NVSwitch multicast mapping or the peers' mappings of the symmetric tile buffer) as coalesced 16-byte stores:
// full NVLink packets instead of one 4..16-byte packet per value.  Call after a __syncthreads().
__device__ __forceinline__ void mirror_floats(float* q0, float* q1, const long long* deltas, int nm) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  float* a0 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q0) + 15) & ~uintptr_t(15));
  if (a0 > q1) a0 = q1;
  float* a1 = a0 + ((q1 - a0) & ~3LL);
  for (float* q = q0 + tid; q < a0; q += nthreads) {  // unaligned head
    const float v = __ldcg(q);
    for (int m = 0; m < nm; ++m) *reinterpret_cast<float*>(reinterpret_cast<char*>(q) + deltas[m]) = v;
  }
  for (float4* q = reinterpret_cast<float4*>(a0) + tid; q < reinterpret_cast<float4*>(a1); q += nthreads) {
    const float4 v = __ldcg(q);
    for (int m = 0; m < nm; ++m) *reinterpret_cast<float4*>(reinterpret_cast<char*>(q) + deltas[m]) = v;
  }
  for (float* q = a1 + tid; q < q1; q += nthreads) {  // tail
    const float v = __ldcg(q);
    for (int m = 0; m < nm; ++m) *reinterpret_cast<float*>(reinterpret_cast<char*>(q) + deltas[m]) = v;
  }
}

// (not inlined, arguments by value, the deltas in shared memory: the fused kernel's register allocation and its
// constant-bank parameters stay exactly what they are without the exchange)
__device__ __noinline__ void mirror_range(float* rgbd, float* cvar, float* dvar, float* term, long long r0, long long r1,
                                          const long long* deltas, int nm) {
  mirror_floats(rgbd + r0 * 4, rgbd + r1 * 4, deltas, nm);
  if (cvar) mirror_floats(cvar + r0 * 3, cvar + r1 * 3, deltas, nm);
  if (dvar) mirror_floats(dvar + r0, dvar + r1, deltas, nm);
  if (term) mirror_floats(term + r0, term + r1, deltas, nm);
}

__device__ __forceinline__ void store_prediction(const TcParams& p, long long delta, long long ray, float4 rgbd,
                                                 float cv0, float cv1, float cv2, float dv, float tp) {
#define NGM_AT(q) reinterpret_cast<float*>(reinterpret_cast<char*>(q) + delta)
  reinterpret_cast<float4*>(NGM_AT(p.rgbd))[ray] = rgbd;
  if (p.color_var) {
    float* cv = NGM_AT(p.color_var) + ray * 3;
    cv[0] = cv0; cv[1] = cv1; cv[2] = cv2;
  }
  if (p.depth_var) NGM_AT(p.depth_var)[ray] = dv;
  if (p.term_prob) NGM_AT(p.term_prob)[ray] = tp;
#undef NGM_AT
}

__device__ __forceinline__ void comp_stage2(const TcParams& p, Smem& sm, int s, int row, int barrier_id, long long f,
                                            long long tile_in_field) {
  ptx::named_bar_sync(barrier_id, 128);
  const int k = row & (p.Sp - 1);
  if (k == 0 && sm.comp[s][7][row] != 0.0f) {
    const int nseg = p.Sp > 32 ? p.Sp >> 5 : 1;
    const int sg0 = row >> (p.sp_shift < 5 ? p.sp_shift : 5);
    float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float carry = 1.0f;
    for (int w = 0; w < nseg; ++w) {
      const float* q = sm.sm_part[s][sg0 + w];
#pragma unroll
      for (int i = 0; i < 9; ++i) m[i] = fmaf(carry, q[i], m[i]);
      carry *= q[9];
    }
    const float P = m[0], D = m[1], C0 = m[2], C1 = m[3], C2 = m[4];
    const float t2 = 2.0f - P;
    const long long ray_global = f * p.rays_per_field + tile_in_field * p.rpt + (row >> p.sp_shift);
    const float4 rgbd = make_float4(C0, C1, C2, D);
    const float cv0 = fmaxf(fmaf(-C0 * C0, t2, m[6]), 0.0f);
    const float cv1 = fmaxf(fmaf(-C1 * C1, t2, m[7]), 0.0f);
    const float cv2 = fmaxf(fmaf(-C2 * C2, t2, m[8]), 0.0f);
    const float dv = fmaxf(fmaf(-D * D, t2, m[5]), 0.0f);
    const float tp = 1.0f - (1.0f - P);
    store_prediction(p, 0, ray_global, rgbd, cv0, cv1, cv2, dv, tp);
    // fused tile exchange (NgmRenderArgs.mirror_delta), diagnostic variant NGM_MIRROR_PER_RAY=1: the same small stores
    // straight into the peer / multicast mappings (1.8 M NVLink packets of 4-16 B per keyframe; measured slower than
    // the per-segment bulk repeat below)
    if (p.mirror_per_ray)
      for (int i = 0; i < p.num_mirrors; ++i) store_prediction(p, sm.mirror_delta[i], ray_global, rgbd, cv0, cv1, cv2, dv, tp);
  }
}

// ---- optional phase trace (diagnostics, block 0 only) --------------------------------------------
// NGM_TC_TRACE=1: low-overhead trace -- each role leader appends (event, clock) to its own ring in
//   shared memory (one STS.64, no atomics) and block 0 dumps the rings to g_trace when it finishes.
// NGM_TC_TRACE=2: every event goes straight to global memory with an atomic (perturbs the timing by
//   ~1k cycles per event, but survives a hang: tools/tc_watchdog.py reads it while the kernel runs).
constexpr int kTraceRoles = 6;   // issuer 0/1, slot 0 {compositor, front-end} half, slot 1 {...}
constexpr int kTraceN = 1024;    // events per role (shared-memory ring: 6 x 1024 x 8 B = 48 KB)
__device__ unsigned long long g_trace[16384];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace_ev(int on, unsigned ev) {
  if (on && blockIdx.x == 0) {
    const unsigned i = atomicAdd(&g_trace_n, 1u);
    if (i < 16384u) g_trace[i] = ((unsigned long long)ev << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
  }
}
template <bool ON>
struct Tracer {
  uint2* ring;  // this role's ring in shared memory, or nullptr
  int n, slow;
  __device__ __forceinline__ void operator()(unsigned ev) {
    if constexpr (!ON) return;
    if (ring) {
      if (n < kTraceN) ring[n++] = make_uint2(ev, (unsigned)clock());
    } else if (slow) {
      trace_ev(1, ev);
    }
  }
  __device__ __forceinline__ void dump() {
    if constexpr (!ON) return;
    if (!ring) return;
    const unsigned base = atomicAdd(&g_trace_n, (unsigned)n);
    for (int i = 0; i < n && base + i < 16384u; ++i)
      g_trace[base + i] = ((unsigned long long)ring[i].x << 48) | (unsigned long long)ring[i].y;
  }
};
// event id = role(4b: 0 issuer, 1 front-end thread, 2 compositor thread) | slot(1b) | phase(4b) | layer(4b)
__device__ __forceinline__ unsigned ev_id(int role, int slot, int phase, int layer) {
  return (unsigned)((role << 12) | (slot << 8) | (phase << 4) | layer);
}

// all operands warp-uniform -> the compiler keeps them in uniform registers and issues UTCHMMA directly
// KS = most K-steps of 16 a layer can have: 8 (K <= 128), or 12 with skip mode "concat" (128 activations + 64 encoding
// features).  A template parameter: the four extra predicated-off steps cost the default kernel 3.5 % when the bound
// was simply raised (the issue loop sits on the kernel's critical path).
template <int KS, bool EXACT = false>
__device__ __forceinline__ void issue_layer(uint32_t d_addr, uint32_t a_addr, uint64_t desc0, uint32_t atom_stride16,
                                            uint32_t idesc, int ksteps) {
  if constexpr (!EXACT) {  // the common depths without the per-step predicate
    if (ksteps == 8 && KS >= 8) { issue_layer<8, true>(d_addr, a_addr, desc0, atom_stride16, idesc, 8); return; }
    if (ksteps == 4) { issue_layer<4, true>(d_addr, a_addr, desc0, atom_stride16, idesc, 4); return; }
    if (ksteps == 3) { issue_layer<3, true>(d_addr, a_addr, desc0, atom_stride16, idesc, 3); return; }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    if (EXACT || ks < ksteps) {
      const uint64_t desc = desc0 + (uint64_t)((ks & 3) * 2u + (ks >> 2) * atom_stride16);
      ptx::mma_f16_ts(d_addr, a_addr + ks * 8, desc, idesc, ks > 0 ? 1u : 0u);
    }
  }
}

// Thread layout (576 threads, one CTA per SM):
//   warp 0  : MMA issuer of slot 0 (+ weight-image loader)     warp 1 : MMA issuer of slot 1
//   warps 2-9  : slot 0   (warp w: TMEM quadrant w % 4, half h = ((w-2) / 4) & 1)
//   warps 10-17: slot 1
// Within a slot both halves share the hidden-layer epilogues (h = 0: first column half, h = 1: second);
// h = 1 threads additionally run the FRONT END of the slot's next tile (sample, encode, layer-0 A operand,
// slotted into the waits for the current tile's MMAs), h = 0 threads the COMPOSITOR of the current tile.
// (A lockstep variant in which all 16 warps share every epilogue job of both slots was measured slower --
// 3.85 ms vs 2.83 ms per frame -- because each duty then stalls all 512 threads; profiles/README.md.)
// SKIP: 0 = no skip connection; 1 = "add", 2 = "concat" (ngm/models.py:160-170) -- only with pre-encoded rows
// (p.raw_a): the encoding of a row is read back from them in the hidden epilogues
template <int MODE, int OCT, bool TRACE, bool ACC16, int SKIP = 0>
__global__ void __launch_bounds__(threads_of(MODE), 1) tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // weights image first (1024-B aligned for SWIZZLE_128B), bookkeeping after it; plain pointer
  // arithmetic on smem_raw keeps the shared address space (LDS/STS, not generic LD/ST)
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* wsm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  Smem& sm = *reinterpret_cast<Smem*>(wsm + (p.im.total_bytes + 127) / 128 * 128);
  const uint32_t wsm_addr = ptx::smem_u32(wsm);
  uint2* const trace_rings = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(&sm) + (sizeof(Smem) + 15) / 16 * 16);

  // warp index via broadcast: provably warp-uniform, so role branches and the shuffles inside them are
  // compiled as uniform control flow (no WARPSYNC / ENDCOLLECTIVE wrappers)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 0) ptx::tmem_alloc(&sm.tmem_base, kTmemCols);
  if (MODE == 0 && tid >= 96 && tid < 96 + NGM_MAX_MIRRORS) sm.mirror_delta[tid - 96] = p.mirror_delta[tid - 96];
  if (tid == 64) {
    for (int s = 0; s < 2; ++s) {
      // 128 front-end stores + 128 compositor threads that drained the accumulator.  No parity aliasing on d_ready:
      // the compositor threads arrive only after the slot-wide named barrier that follows the last-layer wait, so
      // the next tile's layer 0 cannot complete before all 256 slot threads have observed the last-layer phase.
      ptx::mbar_init(&sm.a0_ready[s], 8 * kArrivalsPerWarp);
      ptx::mbar_init(&sm.a_ready[s], 8 * kArrivalsPerWarp);
      ptx::mbar_init(&sm.d_ready[s], 1);
    }
    ptx::mbar_init(&sm.w_ready, 1);
    for (int i = 0; i < 8; ++i) {
      ptx::mbar_init(&sm.ray_full[i >> 2][i & 3], 32);
      ptx::mbar_init(&sm.ray_empty[i >> 2][i & 3], 4 * kArrivalsPerWarp);
    }
    ptx::fence_mbar_init();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  // "concat" keeps a tile's encoding right behind its activations (columns kACol + W/2 ..., up to kStageCol + 32) for
  // all of its layers, so the next tile's layer-0 operand is staged in the spare columns behind that
  constexpr uint32_t stage_col = SKIP == 2 ? kStageCol + 32 : kStageCol;
  // contiguous, balanced tile range of this CTA
  const bool gather = MODE == 1 && p.entries != nullptr;
  const long long total_tiles = gather ? (long long)__ldg(p.tile_offsets + p.num_fields) : p.total_tiles;
  const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;

  uint32_t w_phase = 0;
  int ray_n0 = 0, ray_n1 = 0;  // tiles of slot 0 / 1 started by this CTA so far (ray-parameter ring position = n & 3)
  uint32_t pa0 = 0, pa = 0;  // MMA issuer: parities of its slot's a0_ready / a_ready
  uint32_t pd = 0;           // slot thread: parity of its slot's d_ready
  const int L = p.L, W = p.W;

  Tracer<TRACE> tev{nullptr, 0, 0};
  auto tev_role = [&](int role, bool leader) {  // called once per field segment; keeps the ring position
    if constexpr (!TRACE) return;
    if (p.trace == 1 && blockIdx.x == 0 && leader) tev.ring = trace_rings + role * kTraceN;
    tev.slow = p.trace == 2 && blockIdx.x == 0 && leader;
  };

  long long t = t_begin;
  while (t < t_end) {
    long long f, seg_end, seg_begin;
    if (gather) {  // largest f with tile_offsets[f] <= t (skips fields without entries)
      int lo = 0, hi = p.num_fields;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.tile_offsets + mid) <= t) lo = mid; else hi = mid;
      }
      f = lo;
      seg_begin = __ldg(p.tile_offsets + lo);
      seg_end = __ldg(p.tile_offsets + lo + 1);
    } else {
      f = t / p.tiles_per_field;
      seg_begin = f * p.tiles_per_field;
      seg_end = seg_begin + p.tiles_per_field;
    }
    if (seg_end > t_end) seg_end = t_end;
    const int ntiles = (int)(seg_end - t);
    const long long tile0_in_field = t - seg_begin;

    // ---- stage this field's weight image (TMA engine) ----
    if (tid == 0) {
      const long long img = p.images_by_slot ? (p.field_slots ? p.field_slots[f] : f) : f;
      const uint8_t* src = p.images + (size_t)img * p.im.total_bytes;
      ptx::mbar_arrive_expect_tx(&sm.w_ready, p.im.total_bytes);
      for (uint32_t o = 0; o < p.im.total_bytes; o += 32768) {
        const uint32_t n = p.im.total_bytes - o < 32768 ? p.im.total_bytes - o : 32768;
        ptx::bulk_g2s(wsm + o, src + o, n, &sm.w_ready);
      }
    }

    if (MODE == 0 && warp == 18) {
      // ===================== ray-parameter producer (fused render only) =====================
      // Per-ray quantities of every tile (12 floats per ray; dependent global loads, ~1.5k cycles of
      // latency) are produced up to four tiles ahead of each slot by this otherwise idle warp, so no
      // slot warp ever waits on HBM.  Lanes 0-15 serve slot 0's tile, lanes 16-31 slot 1's.
      if (MODE == 0) {
        const long long slot = p.field_slots ? p.field_slots[f] : f;
        const int sl = lane >> 4, r = lane & 15;
        for (int tp = 0; tp < ntiles; tp += 2) {
          const int n0 = ray_n0 + (tp >> 1), n1 = ray_n1 + (tp >> 1);
          const bool have1 = tp + 1 < ntiles;
          if (n0 >= 4) ptx::mbar_wait(&sm.ray_empty[0][n0 & 3], ((n0 >> 2) + 1) & 1);
          if (have1 && n1 >= 4) ptx::mbar_wait(&sm.ray_empty[1][n1 & 3], ((n1 >> 2) + 1) & 1);
          const int n = sl ? n1 : n0;
          if (r < p.rpt && (sl == 0 || have1))
            compute_ray_params(p, f, slot, (tile0_in_field + tp + sl) * p.rpt + r, sm.ray[sl][n & 3][r]);
          ptx::mbar_arrive(&sm.ray_full[0][n0 & 3]);
          if (have1) ptx::mbar_arrive(&sm.ray_full[1][n1 & 3]);
        }
      }
    } else if (warp < 2) {
      // ===================== MMA issuer of slot `warp` =====================
      const int s = warp;
      tev_role(s, lane == 0);
      ptx::mbar_wait(&sm.w_ready, w_phase);
      const int my_tiles = s == 0 ? (ntiles + 1) / 2 : ntiles / 2;
      // broadcast -> provably warp-uniform operands (uniform registers, no per-lane waterfall)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t wsm_u = __shfl_sync(0xffffffffu, wsm_addr, 0);
      const uint32_t d_addr = tmem_u + s * kSlotCols;
      const uint32_t a_addr = d_addr + kACol;
      for (int it = 0; it < my_tiles; ++it) {
        for (int l = 0; l <= L; ++l) {
          if (l == 0) { ptx::mbar_wait_lean(&sm.a0_ready[s], pa0); pa0 ^= 1; }
          else        { ptx::mbar_wait_lean(&sm.a_ready[s], pa);   pa ^= 1; }
          ptx::tc_fence_after();
          tev(ev_id(0, s, 0, l));
          const TcLayer y = p.im.layer[l];
          const uint64_t desc0 = ptx::make_smem_desc_sw128(wsm_u + y.off);
          const uint32_t idesc = (ACC16 && l < L) ? ptx::make_idesc_f16_acc16(y.n_pad) : ptx::make_idesc_f16(y.n_pad);
          if (ptx::elect_one()) {
            issue_layer<SKIP == 2 ? 12 : 8>(d_addr, l == 0 ? d_addr + stage_col : a_addr, desc0, (uint32_t)y.n_pad * 8u, idesc,
                                            y.k_pad / 16);
            ptx::mma_commit(&sm.d_ready[s]);
          }
          __syncwarp();
          tev(ev_id(0, s, 1, l));
        }
      }
    } else {
      // ===================== tile-slot threads =====================
      const int s = (warp - 2) >> 3;
      const int h = ((warp - 2) >> 2) & 1;  // 0: compositor half, 1: front-end half
      const int qwarp = warp & 3;           // TMEM lane quadrant this warp may access
      const int row = qwarp * 32 + lane;
      const uint32_t d_addr = tmem_base + ((uint32_t)(qwarp * 32) << 16) + s * kSlotCols;
      const uint32_t a_addr = d_addr + kACol;
      const uint32_t* bias2 = reinterpret_cast<const uint32_t*>(wsm + p.im.bias_h2_off);
      const float* bias_last = reinterpret_cast<const float*>(wsm + p.im.bias_last_off);
      const long long slot = p.field_slots ? p.field_slots[f] : f;
      const int bar_slot = 1 + s;  // 256 threads of the slot
      const int bar_half = 3 + 2 * s + h;  // the 128 threads of this half (compositor or front end)
      // column split of the hidden epilogues: multiples of 16, h = 0 takes the (larger) first part
      const int w0 = ((W / 16 + 1) / 2) * 16;
      const int my_c0 = h ? w0 : 0, my_n = h ? W - w0 : w0;
      tev_role(2 + 2 * s + h, lane == 0 && qwarp == 2);  // one leader thread per half

      const uint32_t a0_addr = d_addr + stage_col;  // staging columns: layer-0 A operand of the NEXT tile
      // layer-0 A operand read from HBM (pre-encoded rows): compiled out of the fused NeRF renderer
      const bool raw = (MODE == 1 || OCT == 0) && p.raw_a != nullptr;
      float3 fx = make_float3(0.f, 0.f, 0.f);        // sample point of this row for the next tile (fe_a -> fe_b)
      // skip connections: the pre-encoded row that fed this thread's row of tile `ti` (nullptr: padding row)
      auto skip_row = [&](int ti) -> const uint32_t* {
        long long rr;
        bool valid;
        if (MODE == 1) {
          const long long gp = (tile0_in_field + ti) * 128 + row;
          valid = gp < p.points_per_field;
          rr = f * p.points_per_field + gp;
          if (gather) {
            const int e = __ldg(p.entry_offsets + f) + (int)gp;
            valid = e < __ldg(p.entry_offsets + f + 1);
            rr = valid ? __ldg(p.entries + e) : 0;
          }
        } else {
          const long long r = (tile0_in_field + ti) * p.rpt + (row >> p.sp_shift);
          const int k = row & (p.Sp - 1);
          valid = r < p.rays_per_field && k < p.St;
          rr = (f * p.rays_per_field + r) * p.St + k;
        }
        return valid ? reinterpret_cast<const uint32_t*>(p.raw_a + rr * p.EP) : nullptr;
      };

      // Front end of tile `ti` (h == 1 threads), in two parts that are slotted into the waits for
      // the CURRENT tile's MMAs:  fe_a = sample point + row data,  fe_b = encoding -> staging -> arrive.
      auto fe_a = [&](int ti, int rn, int par) {  // rn: slot-tile count of tile ti
        const long long tile_in_field = tile0_in_field + ti;
        tev(ev_id(1, s, 0, 0));
        fx = make_float3(0.f, 0.f, 0.f);
        if (MODE == 1) {
          const long long gp = tile_in_field * 128 + row;
          bool ok = !raw && gp < p.points_per_field;
          const float* src = p.points + (f * p.points_per_field + gp) * 3;
          if (gather) {
            const int e = __ldg(p.entry_offsets + f) + (int)gp;
            ok = e < __ldg(p.entry_offsets + f + 1);
            if (ok) src = p.points + (long long)(__ldg(p.entries + e) / p.knn_k) * 3;
          }
          if (ok) {
            float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
            if (p.positions) {
              const long long pose = gather ? f : slot;  // gather mode: poses are already those of the F fields
              const float* c = p.positions + pose * 3;
              const float* q = p.orientations + pose * 4;
              x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
              x = quat_inv_rotate(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), x);
            }
            fx = scale_local(x, p.scale_mode, p.field_radius);
          }
        } else {
          const int rit = row >> p.sp_shift;
          const int k = row & (p.Sp - 1);
          ptx::mbar_wait(&sm.ray_full[s][rn & 3], (rn >> 2) & 1);  // produced (normally long ago)
          const float* rp = sm.ray[s][rn & 3][rit];
          const bool ray_ok = rp[10] != 0.0f;
          const long long ray = f * p.rays_per_field + tile_in_field * p.rpt + rit;
          const bool valid = ray_ok && k < p.St;
          const float nr = rp[7], fr = rp[8], gt = rp[9];
          const int S = p.S, G = p.G, St = p.St;
          float d = 0.f;
          float z = 0.f;
          if (raw) {  // samples and their encodings were produced by the sampler + encode kernels
            const long long sample = ray * St + k;
            if (valid) { d = __ldg(p.raw_dist + sample); z = __ldg(p.raw_depth + sample); }
            fx.x = __int_as_float(valid ? (int)sample : -1);  // row of the pre-encoded A operand, for fe_b
          } else if (G > 0) {
            // depth-guided merge (run_mapping.py:521-545): own distance + rank, then exchange by rank
            if (valid) {
              float glo, ghi;
              guided_window(nr, fr, gt, p.range_guided, glo, ghi);
              float dk;
              int pos;
              if (k < S) {
                dk = stratified_distance(nr, fr, k, S, p.jit.coarse(ray, k, S, St));
                pos = k + count_before(dk, glo, ghi, G, true, [&](int j) { return p.jit.guided(ray, j, S, G, St); });
              } else {
                const int kg = k - S;
                dk = stratified_distance(glo, ghi, kg, G, p.jit.guided(ray, kg, S, G, St));
                pos = kg + count_before(dk, nr, fr, S, false, [&](int j) { return p.jit.coarse(ray, j, S, St); });
              }
              sm.sm_x[s][rit * p.Sp + pos] = dk;
            }
            ptx::named_bar_sync(bar_half, 128);
            d = sm.sm_x[s][row];
            ptx::named_bar_sync(bar_half, 128);
          } else if (valid) {
            // (delta * u + linspace(0,1,S+1)[k] * (far - near)) + near with the per-ray division hoisted
            const float lin = k < (S + 1) / 2 ? __fmul_rn(p.inv_S, (float)k)
                                              : __fsub_rn(1.0f, __fmul_rn(p.inv_S, (float)(S - k)));
            d = __fadd_rn(__fadd_rn(__fmul_rn(rp[11], p.jit.coarse(ray, k, S, St)), __fmul_rn(lin, __fsub_rn(fr, nr))), nr);
          }
          if (valid && !raw) {
            fx = make_float3(fmaf(d, rp[3], rp[0]), fmaf(d, rp[4], rp[1]), fmaf(d, rp[5], rp[2]));
            z = d * rp[6];
          }
          *reinterpret_cast<float2*>(sm.rowdata[s][par][row]) = make_float2(d, z);
        }
        tev(ev_id(1, s, 1, 0));
      };
      auto fe_b = [&](int ti) {
        if (raw) {
          bool valid;
          long long rr;
          if (MODE == 1) {
            const long long gp = (tile0_in_field + ti) * 128 + row;
            valid = gp < p.points_per_field;
            rr = valid ? f * p.points_per_field + gp : 0;
            if (gather) {  // rows are indexed by the (point, neighbour) entry
              const int e = __ldg(p.entry_offsets + f) + (int)gp;
              valid = e < __ldg(p.entry_offsets + f + 1);
              rr = valid ? __ldg(p.entries + e) : 0;
            }
          } else {
            const int sample = __float_as_int(fx.x);
            valid = sample >= 0;
            rr = valid ? sample : 0;
          }
          const uint32_t* src = reinterpret_cast<const uint32_t*>(p.raw_a + rr * p.EP);
          for (int c = 0; c < p.EP / 2; c += 8) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = valid ? __ldg(src + c + i) : 0u;
            ptx::tmem_st8(a0_addr + c, w);
          }
        } else if constexpr (OCT > 0) {
          encode_nerf_to_tmem<OCT>(a0_addr, fx, p.nerf_start);
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        slot_arrive(&sm.a0_ready[s]);
        tev(ev_id(1, s, 2, 0));
      };
      // permutohedral front end: the 16 x 4 table gathers of a row are spread over the waits of up to four
      // hidden layers (gap g handles a quarter of the levels and stores them straight into the staging columns)
      auto fe_permuto = [&](int l, int ngaps) {
        if (l >= ngaps) return;
        const int Lv = p.pm_levels;
        const int g0 = (Lv + 3) / 4 * l / ngaps * 4, g1 = l + 1 == ngaps ? Lv : (Lv + 3) / 4 * (l + 1) / ngaps * 4;
        encode_permuto_levels(p, a0_addr, fx, slot, g0 < Lv ? g0 : Lv, g1 < Lv ? g1 : Lv);
        if (l + 1 == ngaps) {
          encode_permuto_tail(p, a0_addr, fx);
          ptx::tc_wait_st();
          ptx::tc_fence_before();
          slot_arrive(&sm.a0_ready[s]);
          tev(ev_id(1, s, 2, 0));
        }
      };

      ptx::mbar_wait(&sm.w_ready, w_phase);  // biases live in the image

      // Software-pipelined tile loop.  Iteration `ti` runs the MLP layers of tile ti; the front end of the
      // slot's next tile (ti + 2) is slotted into the waits for tile ti's MMAs (h == 1 threads: fe_a after
      // the first epilogue, fe_b after the second), and the compositor of tile ti (h == 0 threads) runs
      // after the last layer while the next tile's layer 0 is already on the tensor pipe.
      // The first iteration (ti = s - 2) is virtual: it only runs the front end of the first tile.
      int ray_n = (s ? ray_n1 : ray_n0) - 1;  // slot-tile count of tile ti (ray-parameter ring position = ray_n & 3)
      int par = 2;                            // ring slot (of 3) of the row data of tile ti
      int comp_pending = 0;   // h == 0: 1 = inputs of the previous tile parked in sm.comp, 2 = its stage 1 done
      long long comp_tile = 0;
      const int wseg = p.Sp >= 32 ? 32 : p.Sp;
      // one trailing virtual iteration per slot drains the compositor of the slot's last tile
      for (int ti = s - 2; ti < ntiles || comp_pending; ti += 2) {
        const bool real = ti >= 0 && ti < ntiles;
        const bool has_next = ti + 2 < ntiles;
        const int ring = ray_n & 3, npar = par == 2 ? 0 : par + 1;
        const long long tile_in_field = tile0_in_field + ti;
        const int nsteps = real ? L + 1 : 1;
        const int step_b = real ? (L < 1 ? L : 1) : 0;  // fe_a at step 0, fe_b at step min(1, L)
        for (int l = 0; l < nsteps; ++l) {
          if (real) {
            ptx::mbar_wait_lean(&sm.d_ready[s], pd);
            pd ^= 1;
            ptx::tc_fence_after();
            if (l < L) {
              // ---------- hidden layer l (both halves) ----------
              tev(ev_id(1 + (h == 0), s, 3, l));
              if constexpr (SKIP != 0) {
                const uint32_t* enc = skip_row(ti);
                if (SKIP == 1) {
                  if (my_n > 0) hidden_epilogue_add(d_addr, a_addr, bias2 + l * (W / 2), my_c0, my_n, enc, p.EP / 2);
                } else {
                  if (my_n > 0) hidden_epilogue(d_addr, a_addr, bias2 + l * (W / 2), my_c0, my_n);
                  if (l == 0) {  // this tile's encoding behind its activations: words [W/2, W/2 + EP/2), 8 per step
                    const int steps = p.EP / 16, s0 = h ? (steps + 1) / 2 : 0, s1 = h ? steps : (steps + 1) / 2;
                    for (int q = s0; q < s1; ++q) {
                      uint32_t wv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                      if (enc) {
                        const uint4 e0 = __ldg(reinterpret_cast<const uint4*>(enc + q * 8));
                        const uint4 e1 = __ldg(reinterpret_cast<const uint4*>(enc + q * 8 + 4));
                        wv[0] = e0.x; wv[1] = e0.y; wv[2] = e0.z; wv[3] = e0.w;
                        wv[4] = e1.x; wv[5] = e1.y; wv[6] = e1.z; wv[7] = e1.w;
                      }
                      ptx::tmem_st8(a_addr + W / 2 + q * 8, wv);
                    }
                  }
                }
              } else if (my_n > 0) {
                if constexpr (ACC16) hidden_epilogue_acc16(d_addr, a_addr, bias2 + l * (W / 2), my_c0, my_n);
                else                 hidden_epilogue(d_addr, a_addr, bias2 + l * (W / 2), my_c0, my_n);
              }
              ptx::tc_wait_st();
              ptx::tc_fence_before();
              slot_arrive(&sm.a_ready[s]);
              tev(ev_id(1 + (h == 0), s, 4, l));
            } else {
              // ---------- last layer ----------
              tev(ev_id(1 + (h == 0), s, 5, 0));
              // row data / ray parameters written by the front-end half are visible to the compositor half
              ptx::named_bar_sync(bar_slot, 256);
              if (h == 0) {
                if (MODE == 1) {
                  const long long gp = tile_in_field * 128 + row;
                  bool valid = gp < p.points_per_field;
                  float* o = p.out + (f * p.points_per_field + gp) * p.dim_out;
                  if (gather) {
                    const int e = __ldg(p.entry_offsets + f) + (int)gp;
                    valid = e < __ldg(p.entry_offsets + f + 1);
                    if (valid) o = p.out + (long long)__ldg(p.entries + e) * p.dim_out;
                  }
                  for (int c = 0; c < p.im.layer[L].n_pad; c += 16) {
                    uint32_t v[16];
                    ptx::tmem_ld16(d_addr + c, v);
                    ptx::tc_wait_ld();
                    if (valid) {
#pragma unroll
                      for (int i = 0; i < 16; ++i)
                        if (c + i < p.dim_out) o[c + i] = __uint_as_float(v[i]) + bias_last[c + i];
                    }
                  }
                  ptx::tc_fence_before();
                  if (has_next) slot_arrive(&sm.a0_ready[s]);
                } else {
                  uint32_t v[4];
                  ptx::tmem_ld4(d_addr, v);
                  const int rit = row >> p.sp_shift;
                  ptx::mbar_wait(&sm.ray_full[s][ring], (ray_n >> 2) & 1);  // acquire (completed long ago)
                  const float* rp = sm.ray[s][ring][rit];
                  const float2 dz = *reinterpret_cast<const float2*>(sm.rowdata[s][par][row]);
                  sm.comp[s][4][row] = dz.x;
                  sm.comp[s][5][row] = dz.y;
                  sm.comp[s][6][row] = rp[9];
                  sm.comp[s][7][row] = rp[10];
                  slot_arrive(&sm.ray_empty[s][ring]);  // the producer may refill this ring entry
                  ptx::tc_wait_ld();
                  ptx::tc_fence_before();
                  if (has_next) slot_arrive(&sm.a0_ready[s]);  // the next layer-0 MMA may overwrite the accumulator
                  sm.comp[s][0][row] = p.color_factor * (__uint_as_float(v[0]) + bias_last[0]);
                  sm.comp[s][1][row] = p.color_factor * (__uint_as_float(v[1]) + bias_last[1]);
                  sm.comp[s][2][row] = p.color_factor * (__uint_as_float(v[2]) + bias_last[2]);
                  sm.comp[s][3][row] = __uint_as_float(v[3]) + bias_last[3];
                  comp_pending = 1;
                  comp_tile = tile_in_field;
                }
                tev(ev_id(2, s, 7, 0));
              }
            }
          } else if (h == 0) {
            if (has_next) slot_arrive(&sm.a0_ready[s]);  // first tile: no accumulator to drain
          }
          if (h == 1 && has_next) {
            if (l == 0) fe_a(ti + 2, ray_n + 1, npar);
            if (OCT > 0 || raw) {
              if (l == step_b) fe_b(ti + 2);
            } else {
              fe_permuto(l, real ? (L < 4 ? (L < 1 ? 1 : L) : 4) : 1);
            }
          }
          // deferred compositor of the slot's previous tile, in the wait for this tile's next MMA
          if (MODE == 0 && h == 0 && comp_pending && l < 2 && (l < L || !real)) {
            if (l == 0) {
              comp_stage1(p, sm, s, row, lane, bar_half, wseg, f, comp_tile);
              comp_pending = 2;
            }
            if (l == 1 || L == 1 || !real) {
              comp_stage2(p, sm, s, row, bar_half, f, comp_tile);
              comp_pending = 0;
              tev(ev_id(2, s, 6, 0));
            }
          }
        }
        ++ray_n;
        par = npar;
      }
    }
    w_phase ^= 1;
    ray_n0 += (ntiles + 1) / 2;
    ray_n1 += ntiles / 2;
    t = seg_end;
    ptx::fence_proxy_async();  // generic-proxy reads of the image before the next bulk copy overwrites it
    __syncthreads();
    if (MODE == 0 && p.num_mirrors > 0 && !p.mirror_per_ray) {
      const long long r0 = f * p.rays_per_field + tile0_in_field * p.rpt;
      long long r1 = r0 + (long long)ntiles * p.rpt;
      if (r1 > (f + 1) * p.rays_per_field) r1 = (f + 1) * p.rays_per_field;
      mirror_range(p.rgbd, p.color_var, p.depth_var, p.term_prob, r0, r1, sm.mirror_delta, p.num_mirrors);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, kTmemCols);
  tev.dump();
}

// one persistent CTA per SM; NGM_TC_MAX_CTAS (tests) caps the grid so that small problems exercise the
// multi-round / multi-segment paths of a CTA
int tc_grid(long long total_tiles) {
  const char* e = getenv("NGM_TC_MAX_CTAS");  // read per call: tests set it for single cases
  const int cap = e ? atoi(e) : 0;
  long long g = num_sms();
  if (cap > 0 && cap < g) g = cap;
  return (int)(total_tiles < g ? total_tiles : g);
}

// the traced twin of the kernel exists only in the debug build (libngm_b200_debug.so, -DNGM_DEBUG_EXPORTS)
int tc_trace_enabled() {
#ifdef NGM_DEBUG_EXPORTS
  static int v = -1;
  if (v < 0) { const char* e = getenv("NGM_TC_TRACE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
#else
  return 0;
#endif
}

template <int MODE>
int launch_tc(const TcParams& p_in, int octaves, size_t smem, int grid, cudaStream_t stream) {
  TcParams p = p_in;
  p.trace = tc_trace_enabled();
  {
    const char* e = getenv("NGM_TC_ACC16");  // read per call: tests / A-B runs switch it
    p.acc16 = (e && e[0] == '1') ? 1 : 0;
  }
  if (p.trace) {
    unsigned zero = 0;
    cudaMemcpyToSymbolAsync(g_trace_n, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, stream);
  }
  auto go = [&](auto kernel) -> int {
    NGM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, threads_of(MODE), smem, stream>>>(p);
    return check_launch("tc_kernel");
  };
  if (p.skip) {  // skip connections: pre-encoded rows only (field_tc_supported / tc_rows_required)
    if (!p.raw_a) { set_error("tcgen05 path: skip connections need pre-encoded rows"); return NGM_ERR_UNSUPPORTED; }
    return p.skip == 1 ? go(tc_kernel<MODE, 0, false, false, 1>) : go(tc_kernel<MODE, 0, false, false, 2>);
  }
  if (p.acc16 && !p.trace) {  // 16-bit accumulation of the hidden layers (NGM_TC_ACC16=1)
    switch (octaves) {
      case 0: return go(tc_kernel<MODE, 0, false, true>);
      case 4: return go(tc_kernel<MODE, 4, false, true>);
      case 8: return go(tc_kernel<MODE, 8, false, true>);
      default: break;
    }
  }
  switch (octaves * 2 + (p.trace ? 1 : 0)) {  // octaves == 0: permutohedral
    case 0: return go(tc_kernel<MODE, 0, false, false>);
    case 8: return go(tc_kernel<MODE, 4, false, false>);
    case 16: return go(tc_kernel<MODE, 8, false, false>);
#ifdef NGM_DEBUG_EXPORTS
    case 1: return go(tc_kernel<MODE, 0, true, false>);
    case 9: return go(tc_kernel<MODE, 4, true, false>);
    case 17: return go(tc_kernel<MODE, 8, true, false>);
#endif
    default: set_error("tcgen05 path: unsupported num_octaves %d", octaves); return NGM_ERR_UNSUPPORTED;
  }
}

// kernel variant: 4 / 8 = NeRF front end in the kernel; 0 = permutohedral front end or pre-encoded rows
int tc_oct(const NgmFieldDesc& fd) {
  return fd.encoding == NGM_ENC_NERF && nerf_octaves_supported(fd.nerf_num_octaves) ? fd.nerf_num_octaves : 0;
}

int fill_common(TcParams& p, const NgmFieldDesc& fd, int num_fields, const float* positions, const float* orientations,
                const int64_t* slots, int scale_mode, float radius, void* workspace, cudaStream_t stream) {
  p.E = fd.dim_encoding;
  p.EP = ep_of(fd);
  p.W = fd.dim_mlp_out;
  p.L = fd.num_layers;
  p.dim_out = fd.dim_out;
  p.nerf_start = fd.nerf_start_octave;
  p.skip = fd.skip_mode == NGM_SKIP_ADD ? 1 : (fd.skip_mode == NGM_SKIP_CONCAT ? 2 : 0);
  if (fd.encoding == NGM_ENC_PERMUTO) {
    p.pm_table = fd.enc_param0; p.pm_table_stride = fd.enc_param0_stride;
    p.pm_shift = fd.enc_param1; p.pm_shift_stride = fd.enc_param1_stride;
    p.pm_scale = fd.permuto_scale;
    p.pm_levels = fd.permuto_levels; p.pm_log2cap = fd.permuto_log2_capacity;
    p.pm_concat = fd.permuto_concat_points; p.pm_concat_scale = fd.permuto_concat_scaling;
  }
  p.im = make_image(fd, p.EP);
  p.images = static_cast<const uint8_t*>(fd.packed_weights ? fd.packed_weights : workspace);
  p.images_by_slot = fd.packed_weights ? 1 : 0;
  p.num_fields = num_fields;
  p.positions = positions;
  p.orientations = orientations;
  p.field_slots = reinterpret_cast<const long long*>(slots);
  p.scale_mode = scale_mode;
  p.field_radius = radius;
  if (fd.packed_weights) return NGM_OK;  // persistent images: nothing to pack per call
  PackParams pk{fd, p.im, p.field_slots, static_cast<uint8_t*>(workspace), fd.dim_encoding, 0};
  pack_weights_kernel<<<num_fields, 256, 0, stream>>>(pk);
  return check_launch("pack_weights_kernel");
}

// >= 120 KB so that exactly one CTA (one 512-column TMEM allocation) is resident per SM
size_t tc_smem_bytes(const TcImage& im) {
  const size_t need = 1024 + (im.total_bytes + 127) / 128 * 128 + sizeof(Smem) + 128 +
                      (tc_trace_enabled() == 1 ? (size_t)kTraceRoles * kTraceN * sizeof(uint2) : 0);
  return need < 120 * 1024 ? 120 * 1024 : need;
}

}  // namespace

// Encodings without an in-kernel front end (Fourier, Triplane, NeRF with an octave count other than 4 / 8) always
// reach the kernel as pre-encoded fp16 rows.
bool tc_rows_required(const NgmFieldDesc& fd) {
  return fd.encoding == NGM_ENC_FOURIER || fd.encoding == NGM_ENC_TRIPLANE ||
         (fd.encoding == NGM_ENC_NERF && !nerf_octaves_supported(fd.nerf_num_octaves)) ||
         fd.skip_mode == NGM_SKIP_ADD || fd.skip_mode == NGM_SKIP_CONCAT;  // the hidden epilogues re-read the encoding
}

bool field_tc_supported(const NgmFieldDesc& fd, const char** why) {
  const char* w = nullptr;
  if (fd.encoding != NGM_ENC_NERF && fd.encoding != NGM_ENC_PERMUTO && fd.encoding != NGM_ENC_FOURIER &&
      fd.encoding != NGM_ENC_TRIPLANE)
    w = "unknown encoding";
  else if ((fd.encoding == NGM_ENC_FOURIER || fd.encoding == NGM_ENC_TRIPLANE ||
            (fd.encoding == NGM_ENC_NERF && !nerf_octaves_supported(fd.nerf_num_octaves))) && ep_of(fd) > 64)
    w = "pre-encoded rows: encoding wider than 64 features";  // (these encodings reach the kernel as fp16 rows)
  else if (fd.encoding == NGM_ENC_PERMUTO && fd.permuto_feats != 2) w = "permutohedral: nr_feat_per_level must be 2 on the tcgen05 path";
  else if (fd.encoding == NGM_ENC_PERMUTO && ep_of(fd) > 64) w = "permutohedral: encoding wider than 64 features";
  else if (fd.skip_mode == NGM_SKIP_REZERO) w = "skip mode rezero is only on the fp32 path";
  else if (fd.skip_mode != NGM_SKIP_NO && ep_of(fd) > 64) w = "skip connections: encoding wider than 64 features";
  else if (fd.dim_mlp_out % 16 != 0 || fd.dim_mlp_out < 16 || fd.dim_mlp_out > 128) w = "dim_mlp_out must be a multiple of 16 in [16,128]";
  else if (fd.dim_out > 128) w = "dim_out > 128";
  else if (make_image(fd, ep_of(fd)).total_bytes > kMaxImageBytes) w = "weight image exceeds shared memory (too many layers)";
  if (why) *why = w;
  return w == nullptr;
}

size_t packed_weights_bytes(const NgmFieldDesc& fd) {
  const char* why;
  if (!field_tc_supported(fd, &why)) return 0;
  return make_image(fd, ep_of(fd)).total_bytes;
}

int launch_pack_weights(const NgmFieldDesc& fd, const int64_t* rows, int num_rows, void* images, cudaStream_t stream) {
  if (num_rows <= 0) return NGM_OK;
  PackParams pk{fd, make_image(fd, ep_of(fd)), reinterpret_cast<const long long*>(rows), static_cast<uint8_t*>(images),
                fd.dim_encoding, 1};
  pack_weights_kernel<<<num_rows, 256, 0, stream>>>(pk);
  return check_launch("pack_weights_kernel");
}

size_t field_tc_workspace_bytes(const NgmFieldDesc& fd, int num_fields) {
  const char* why;
  if (!field_tc_supported(fd, &why)) return 0;
  if (fd.packed_weights) return 0;
  return (size_t)make_image(fd, ep_of(fd)).total_bytes * (size_t)(num_fields > 0 ? num_fields : 1);
}

// Dense field evaluation reads its layer-0 operand as fp16 rows when the caller supplies them, when the encoding has
// no in-kernel front end, and for the permutohedral encoding (its 64 table gathers per point run ~2x faster from a
// whole-GPU row encoder than from the 8 front-end warps of the persistent MMA kernel; NGM_TC_PERMUTO_INKERNEL=1
// keeps the in-kernel front end).
static bool fwd_rows_from_encoder(const NgmFieldFwdArgs& a) {
  if (a.rows_half) return false;
  if (tc_rows_required(a.field)) return true;
  const char* e = getenv("NGM_TC_PERMUTO_INKERNEL");
  return a.field.encoding == NGM_ENC_PERMUTO && !(e && e[0] == '1') &&
         (long long)a.num_fields * a.points_per_field < (1ll << 31);
}

// workspace: weight images, then (when the library encodes the rows itself) one fp16 row per point
size_t field_tc_fwd_workspace_bytes(const NgmFieldFwdArgs& a) {
  size_t n = (field_tc_workspace_bytes(a.field, a.num_fields) + 255) / 256 * 256;
  if (fwd_rows_from_encoder(a)) n += (size_t)a.num_fields * (size_t)a.points_per_field * (size_t)ep_of(a.field) * 2;
  return n;
}

int launch_permuto_rows_half(const PermutoRowsArgs& a, cudaStream_t stream);

int launch_field_fwd_tc(const NgmFieldFwdArgs& a, cudaStream_t stream) {
  TcParams p{};
  if (int rc = fill_common(p, a.field, a.num_fields, a.positions, a.orientations, a.field_slots, a.scale_mode,
                           a.field_radius, a.workspace, stream))
    return rc;
  p.points = a.points;
  p.points_per_field = a.points_per_field;
  p.out = a.out;
  p.tiles_per_field = (a.points_per_field + 127) / 128;
  p.total_tiles = p.tiles_per_field * a.num_fields;
  if (a.rows_half) {
    p.raw_a = static_cast<const __half*>(a.rows_half);
  } else if (fwd_rows_from_encoder(a)) {
    PermutoRowsArgs e{};
    e.field = a.field;
    e.points_world = a.points;
    e.positions = a.positions; e.orientations = a.orientations;
    e.field_slots = reinterpret_cast<const long long*>(a.field_slots);
    e.out = reinterpret_cast<uint32_t*>(static_cast<char*>(a.workspace) +
                                        (field_tc_workspace_bytes(a.field, a.num_fields) + 255) / 256 * 256);
    e.num_points = (long long)a.num_fields * a.points_per_field;
    e.points_per_field = a.points_per_field;
    e.field_radius = a.field_radius;
    e.scale_mode = a.scale_mode;
    e.EP = p.EP;
    if (int rc = launch_permuto_rows_half(e, stream)) return rc;
    p.raw_a = reinterpret_cast<const __half*>(e.out);
  }
  const int grid = tc_grid(p.total_tiles);
  return launch_tc<1>(p, p.raw_a ? 0 : tc_oct(a.field), tc_smem_bytes(p.im), grid, stream);
}

// kNN path: every tile is 128 (point, neighbour) entries of ONE field; the tile count lives in tile_offsets[F]
int launch_field_fwd_tc_gather(const NgmFieldFwdArgs& a, const int* entries, const int* entry_offsets,
                               const int* tile_offsets, int knn_k, long long max_tiles, const void* rows_half,
                               cudaStream_t stream) {
  TcParams p{};
  p.raw_a = static_cast<const __half*>(rows_half);  // pre-encoded rows (permutohedral), indexed by entry; or nullptr
  if (int rc = fill_common(p, a.field, a.num_fields, a.positions, a.orientations, a.field_slots, a.scale_mode,
                           a.field_radius, a.workspace, stream))
    return rc;
  p.points = a.points;
  p.out = a.out;
  p.entries = entries;
  p.entry_offsets = entry_offsets;
  p.tile_offsets = tile_offsets;
  p.knn_k = knn_k;
  p.points_per_field = 1ll << 40;  // rows are bounded by the entry ranges
  p.tiles_per_field = 1;
  p.total_tiles = max_tiles;
  const int grid = tc_grid(max_tiles);
  return launch_tc<1>(p, tc_oct(a.field), tc_smem_bytes(p.im), grid, stream);
}

#ifdef NGM_DEBUG_EXPORTS
// debug / unit-test entry: D = A (fp16, given) x W^T with the production weight packing, smem
// descriptors, TMEM A operand and epilogue -- isolates the tcgen05 plumbing from the renderer.
int launch_tc_gemm_debug(const NgmFieldDesc& fd, const void* a_half, long long rows, float* out, void* workspace,
                         cudaStream_t stream) {
  TcParams p{};
  NgmFieldDesc f2 = fd;
  if (int rc = fill_common(p, f2, 1, nullptr, nullptr, nullptr, NGM_SCALE_NO, 0.f, workspace, stream)) return rc;
  p.raw_a = static_cast<const __half*>(a_half);
  p.points_per_field = rows;
  p.out = out;
  p.tiles_per_field = (rows + 127) / 128;
  p.total_tiles = p.tiles_per_field;
  const int grid = tc_grid(p.total_tiles);
  return launch_tc<1>(p, 8, tc_smem_bytes(p.im), grid, stream);
}

// diagnostics: copy the phase trace of the last traced launch to the host (synchronises the device)
int tc_trace_read(unsigned long long* out, int max_events) {
  unsigned n = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)) != cudaSuccess) return -1;
  if (n > 16384u) n = 16384u;
  if ((int)n > max_events) n = (unsigned)max_events;
  if (n && cudaMemcpyFromSymbol(out, g_trace, n * sizeof(unsigned long long)) != cudaSuccess) return -1;
  return (int)n;
}

// same, but WITHOUT synchronising the device: reads the trace of a kernel that is still running (or hung)
// through a non-blocking stream -- the watchdog of tools/tc_trace.py uses it to diagnose deadlocks
int tc_trace_peek(unsigned long long* out, int max_events) {
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return -1;
  unsigned n = 0;
  int rc = -1;
  if (cudaMemcpyFromSymbolAsync(&n, g_trace_n, sizeof(n), 0, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
      cudaStreamSynchronize(st) == cudaSuccess) {
    if (n > 16384u) n = 16384u;
    if ((int)n > max_events) n = (unsigned)max_events;
    if (n == 0 || (cudaMemcpyFromSymbolAsync(out, g_trace, n * sizeof(unsigned long long), 0, cudaMemcpyDeviceToHost, st) ==
                       cudaSuccess &&
                   cudaStreamSynchronize(st) == cudaSuccess))
      rc = (int)n;
  }
  cudaStreamDestroy(st);
  return rc;
}

#endif  // NGM_DEBUG_EXPORTS

int launch_neus_isd(const float* sd, const int64_t* slots, int num_fields, float* out, cudaStream_t stream);

bool render_fused_tc_ok(const NgmRenderArgs& a) {
  const int St = a.num_samples + (a.gt ? a.num_samples_guided : 0);
  const char* why;
  {
    // The single fused kernel (sampler + encoding + MLP + compositor in one persistent tcgen05 kernel) is opt-in:
    // measured against the three stage kernels (sample_rays -> tcgen05 field kernel -> compositor) it is 13-22 %
    // SLOWER at every MLP size (4 x 128 NeRF-8 keyframe: 2.71 against 2.32 ms) -- the compositor and the sampling front
    // end lengthen the per-tile dependency chain of the slot threads by more than the ~1 GB of HBM round trips
    // (0.18 ms) costs.  tools/fused_vs_staged.py, profiles/r2_fused_vs_staged.jsonl.  Read per call.
    const char* e = getenv("NGM_RENDER_FUSED");
    if (!(e && e[0] == '1')) return false;
  }
  // num_layers >= 1: the deferred compositor runs in the waits of the hidden layers
  return a.precision == NGM_PREC_FP16 && St <= 128 && a.field.dim_out == 4 && a.field.num_layers >= 1 &&
         field_tc_supported(a.field, &why);
}

// The permutohedral renderer runs as sampler kernel -> row encoder (encode.cu, whole-GPU occupancy for the L2
// table gathers) -> this kernel with the A operand, distances and depths read back (1.3 GB of HBM traffic per
// 640x480x64 frame, 0.4 ms, against 17 ms saved); NGM_TC_PERMUTO_INKERNEL=1 keeps the in-kernel front end.
bool render_tc_precoded(const NgmRenderArgs& a) {
  const char* e = getenv("NGM_TC_PERMUTO_INKERNEL");  // read per call: tests compare both front ends
  const bool inkernel = e && e[0] == '1';
  const long long St = a.num_samples + (a.gt ? a.num_samples_guided : 0);
  const bool fits = (long long)a.num_fields * a.rays_per_field * St < (1ll << 31);
  return tc_rows_required(a.field) || (a.field.encoding == NGM_ENC_PERMUTO && !inkernel && fits);
}

int launch_render_fused_tc(const NgmRenderArgs& a, void* tc_ws, float* isd_ws, const void* rows_half, const float* dist,
                           const float* depth, cudaStream_t stream) {
  TcParams p{};
  p.raw_a = static_cast<const __half*>(rows_half);
  p.raw_dist = dist;
  p.raw_depth = depth;
  if (int rc = fill_common(p, a.field, a.num_fields, a.positions, a.orientations, a.field_slots, a.scale_mode,
                           a.field_radius, tc_ws, stream))
    return rc;
  p.cam = a.cam;
  p.ijs = reinterpret_cast<const long long*>(a.ijs);
  p.c2ws = a.c2ws;
  p.near = a.near; p.far = a.far; p.gt = a.gt;
  p.jit = RayJitter{a.jitter, a.jitter_guided, a.seed, a.offset};
  p.near_scalar = a.near_scalar; p.far_scalar = a.far_scalar; p.range_guided = a.range_guided;
  p.c2w_per_ray = a.c2w_per_ray;
  p.S = a.num_samples;
  p.inv_S = 1.0f / (float)a.num_samples;
  p.G = a.gt ? a.num_samples_guided : 0;
  p.St = p.S + p.G;
  int Sp = 8, shift = 3;  // ray stride in rows: power of two >= max(St, 8)
  while (Sp < p.St) { Sp <<= 1; ++shift; }
  p.Sp = Sp;
  p.sp_shift = shift;
  p.rpt = 128 / Sp;
  p.rays_per_field = a.rays_per_field;
  p.geometry_mode = a.geometry_mode;
  p.overwrite = a.overwrite_behind_camera;
  p.overwrite_gate = a.overwrite_gate;
  p.geometry_factor = a.geometry_factor; p.color_factor = a.color_factor; p.truncation = a.truncation;
  if (a.geometry_mode == NGM_GEOM_NEUS) {
    if (int rc = launch_neus_isd(a.neus_sd, a.field_slots, a.num_fields, isd_ws, stream)) return rc;
    p.neus_isd = isd_ws;
  }
  p.rgbd = a.rgbd; p.color_var = a.color_var; p.depth_var = a.depth_var; p.term_prob = a.term_prob;
  p.num_mirrors = a.num_mirrors;
  { const char* e = getenv("NGM_MIRROR_PER_RAY"); p.mirror_per_ray = (e && e[0] == '1') ? 1 : 0; }  // diagnostics, read per call
  for (int i = 0; i < a.num_mirrors; ++i) p.mirror_delta[i] = a.mirror_delta[i];
  p.freespace = a.freespace; p.freespace_mask = a.freespace_mask; p.tsdf = a.tsdf; p.tsdf_mask = a.tsdf_mask;
  p.tiles_per_field = (a.rays_per_field + p.rpt - 1) / p.rpt;
  p.total_tiles = p.tiles_per_field * a.num_fields;
  const int grid = tc_grid(p.total_tiles);
  return launch_tc<0>(p, tc_oct(a.field), tc_smem_bytes(p.im), grid, stream);
}

}  // namespace ngm
