// Device-side positional encodings (fp32), one feature (or one lattice level) per call.
// Restates ngm/positional_encodings.py: PositionalEncodingNeRF :245-272, PositionalEncodingFourier
// :197-212, TriplaneEncoding :132-161 and the third-party permutohedral encoding wrapped at :19-66
// (published algorithm; parity unpinned -- see oracle/permuto.py).
#pragma once
#include "common.cuh"

namespace ngm {

// feature c of sin/cos(2^o * pi * x): layout [sin: dim-major, octave-minor | cos: same].
__device__ __forceinline__ float nerf_feature(const float* x, int c, int num_octaves, int start_octave) {
  const int half = 3 * num_octaves;
  const bool is_cos = c >= half;
  const int cc = is_cos ? c - half : c;
  const int dim = cc / num_octaves, oct = cc - dim * num_octaves;
  const float mult = ldexpf(3.14159274101257324f, start_octave + oct);  // fl32(pi) * 2^o, exact
  const float arg = __fmul_rn(x[dim], mult);
  return is_cos ? cosf(arg) : sinf(arg);
}

// feature c of [x | sin(W x)] (raw_coords) or sin(W x);  w = (n,3) row-major
__device__ __forceinline__ float fourier_feature(const float* x, int c, const float* __restrict__ w, int raw) {
  if (raw) {
    if (c < 3) return x[c];
    c -= 3;
  }
  const float* r = w + c * 3;
  float s = __ldg(r) * x[0];
  s = fmaf(__ldg(r + 1), x[1], s);
  s = fmaf(__ldg(r + 2), x[2], s);
  return sinf(s);
}

// F.grid_sample(bilinear, align_corners=True, padding_mode="border") of one (res x res) plane
__device__ __forceinline__ float plane_sample(const float* __restrict__ plane, int res, float gx, float gy) {
  const float lim = (float)(res - 1);
  float ix = ((gx + 1.0f) / 2.0f) * lim, iy = ((gy + 1.0f) / 2.0f) * lim;
  ix = fminf(lim, fmaxf(ix, 0.0f));
  iy = fminf(lim, fmaxf(iy, 0.0f));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const float wnw = (fx0 + 1.0f - ix) * (fy0 + 1.0f - iy), wne = (ix - fx0) * (fy0 + 1.0f - iy);
  const float wsw = (fx0 + 1.0f - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
  float out = 0.0f;
  out += __ldg(plane + y0 * res + x0) * wnw;  // (x0,y0) is always in bounds after clipping
  if (x1 < res) out += __ldg(plane + y0 * res + x1) * wne;
  if (y1 < res) out += __ldg(plane + y1 * res + x0) * wsw;
  if (x1 < res && y1 < res) out += __ldg(plane + y1 * res + x1) * wse;
  return out;
}

// feature c of the triplane encoding; coef = (3, C, res, res)
__device__ __forceinline__ float triplane_feature(const float* x, int c, const float* __restrict__ coef,
                                                   int res, int C, int mode) {
  const int plane_elems = res * res;
  // plane 0: (x0,x1)  plane 1: (x0,x2)  plane 2: (x1,x2); grid[...,0] indexes the last (width) dim
  if (mode == NGM_TRIPLANE_CONCAT) {
    const int pl = c / C, comp = c - pl * C;
    const float gx = pl == 2 ? x[1] : x[0], gy = pl == 0 ? x[1] : x[2];
    return plane_sample(coef + ((size_t)pl * C + comp) * plane_elems, res, gx, gy);
  }
  const float a = plane_sample(coef + ((size_t)0 * C + c) * plane_elems, res, x[0], x[1]);
  const float b = plane_sample(coef + ((size_t)1 * C + c) * plane_elems, res, x[0], x[2]);
  const float d = plane_sample(coef + ((size_t)2 * C + c) * plane_elems, res, x[1], x[2]);
  return mode == NGM_TRIPLANE_PRODUCT ? (a * b) * d : (a + b) + d;
}

// Enclosing simplex of a 3-D point on one level of the permutohedral lattice: the hashed table rows of
// its four vertices and their barycentric weights.  shift/scale = 3 floats of this level.
__device__ __forceinline__ void permuto_simplex(const float* x, const float* __restrict__ shift,
                                                const float* __restrict__ scale, int log2_capacity, uint32_t (&row)[4],
                                                float (&weight)[4]) {
  constexpr int D = 3, D1 = 4;
  float cf[D];
#pragma unroll
  for (int i = 0; i < D; ++i) cf[i] = __fmul_rn(__fadd_rn(x[i], __ldg(shift + i)), __ldg(scale + i));
  float E[D1];
  float sm = 0.0f;
#pragma unroll
  for (int i = D; i > 0; --i) {
    E[i] = __fsub_rn(sm, __fmul_rn((float)i, cf[i - 1]));
    sm = __fadd_rn(sm, cf[i - 1]);
  }
  E[0] = sm;
  int rem0[D1], rank[D1] = {0, 0, 0, 0};
  int sum = 0;
#pragma unroll
  for (int i = 0; i < D1; ++i) {
    const float v = __fmul_rn(E[i], 1.0f / D1);
    const float up = ceilf(v) * D1, down = floorf(v) * D1;
    rem0[i] = (__fsub_rn(up, E[i]) < __fsub_rn(E[i], down)) ? (int)up : (int)down;
    sum += rem0[i];
  }
  sum /= D1;  // C truncation
  float resid[D1];
#pragma unroll
  for (int i = 0; i < D1; ++i) resid[i] = __fsub_rn(E[i], (float)rem0[i]);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j < D1; ++j) {
      if (resid[i] < resid[j]) rank[i]++;
      else rank[j]++;
    }
#pragma unroll
  for (int i = 0; i < D1; ++i) {
    rank[i] += sum;
    if (rank[i] < 0) { rank[i] += D1; rem0[i] += D1; }
    else if (rank[i] > D) { rank[i] -= D1; rem0[i] -= D1; }
  }
  float bary[D + 2] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < D1; ++i) {
    const float delta = __fmul_rn(__fsub_rn(E[i], (float)rem0[i]), 1.0f / D1);
    // bary[D - rank] += delta; bary[D + 1 - rank] -= delta  (rank in 0..3; unrolled select keeps registers)
#pragma unroll
    for (int r = 0; r <= D + 1; ++r) {
      if (r == D - rank[i]) bary[r] = __fadd_rn(bary[r], delta);
      if (r == D + 1 - rank[i]) bary[r] = __fsub_rn(bary[r], delta);
    }
  }
  bary[0] = __fadd_rn(bary[0], __fadd_rn(1.0f, bary[D + 1]));
  const uint32_t mask = (1u << log2_capacity) - 1u;
#pragma unroll
  for (int r = 0; r < D1; ++r) {
    uint32_t h = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      int key = rem0[i] + r;
      if (rank[i] > D - r) key -= D1;
      h = (h + (uint32_t)key) * 2531011u;
    }
    row[r] = h & mask;
    weight[r] = bary[r];
  }
}

// One level of the permutohedral-lattice hash encoding for a 3-D point: writes `feats` values.
//   table = (capacity, feats) of this level.
template <int MAXF>
__device__ __forceinline__ void permuto_level(const float* x, const float* __restrict__ table,
                                              const float* __restrict__ shift, const float* __restrict__ scale,
                                              int log2_capacity, int feats, float* out) {
  uint32_t row[4];
  float weight[4];
  permuto_simplex(x, shift, scale, log2_capacity, row, weight);
  float acc[MAXF];
#pragma unroll
  for (int f = 0; f < MAXF; ++f) acc[f] = 0.0f;
  if (feats == 2 && MAXF >= 2 && (reinterpret_cast<uintptr_t>(table) & 7) == 0) {
    // nr_feat_per_level = 2 (the reference's default, neural_graph_map.yaml:11): one 8-byte gather per vertex instead
    // of two 4-byte ones -- the row encoder is bound by L2 sector requests (ncu: 2.1 sectors per vertex before)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(table) + row[r]);
      acc[0] = __fadd_rn(acc[0], __fmul_rn(weight[r], v.x));
      acc[1] = __fadd_rn(acc[1], __fmul_rn(weight[r], v.y));
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* fv = table + (size_t)row[r] * feats;
#pragma unroll
      for (int f = 0; f < MAXF; ++f)
        if (f < feats) acc[f] = __fadd_rn(acc[f], __fmul_rn(weight[r], __ldg(fv + f)));
    }
  }
#pragma unroll
  for (int f = 0; f < MAXF; ++f)
    if (f < feats) out[f] = acc[f];
}

// Gradient of one level with respect to its table: d_table[row_r] += weight_r * d_out  (atomics: many points
// share a vertex).  The point itself carries no gradient (sample positions do not depend on parameters).
template <int MAXF>
__device__ __forceinline__ void permuto_level_bwd(const float* x, float* __restrict__ d_table,
                                                  const float* __restrict__ shift, const float* __restrict__ scale,
                                                  int log2_capacity, int feats, const float* d_out) {
  uint32_t row[4];
  float weight[4];
  permuto_simplex(x, shift, scale, log2_capacity, row, weight);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* dv = d_table + (size_t)row[r] * feats;
#pragma unroll
    for (int f = 0; f < MAXF; ++f)
      if (f < feats) atomicAdd(dv + f, weight[r] * d_out[f]);
  }
}

}  // namespace ngm
