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
//
// Register-only and branch-free.  The textbook form indexes bary[] and the key offsets by the run-time ranks; the
// compiler turned that into local-memory arrays and jump tables (172 LDL + 172 STL + 32 BRX in the row encoder, ~650
// instructions per level).  Here the ranks -- a permutation of 0..3 -- are inverted with selects instead:
// bary[s] = delta[rank == 3-s] - delta[rank == 4-s], which is what the accumulation loop computes (each slot receives
// exactly one +delta and one -delta, and 0 + a - b = 0 - b + a in floating point), so rows and weights are bit-identical.
__device__ __forceinline__ void permuto_simplex(const float* x, const float* __restrict__ shift,
                                                const float* __restrict__ scale, int log2_capacity, uint32_t (&row)[4],
                                                float (&weight)[4]) {
  const float cf0 = __fmul_rn(__fadd_rn(x[0], __ldg(shift + 0)), __ldg(scale + 0));
  const float cf1 = __fmul_rn(__fadd_rn(x[1], __ldg(shift + 1)), __ldg(scale + 1));
  const float cf2 = __fmul_rn(__fadd_rn(x[2], __ldg(shift + 2)), __ldg(scale + 2));
  // elevate: E[i] = sm - i * cf[i-1], sm += cf[i-1], for i = 3, 2, 1; E[0] = sm
  const float E3 = __fsub_rn(0.0f, __fmul_rn(3.0f, cf2));
  const float s2 = __fadd_rn(0.0f, cf2);
  const float E2 = __fsub_rn(s2, __fmul_rn(2.0f, cf1));
  const float s1 = __fadd_rn(s2, cf1);
  const float E1 = __fsub_rn(s1, __fmul_rn(1.0f, cf0));
  const float E0 = __fadd_rn(s1, cf0);
  // nearest lattice point with remainder 0
  auto nearest = [](float e) {
    const float v = __fmul_rn(e, 0.25f);
    const float up = ceilf(v) * 4.0f, down = floorf(v) * 4.0f;
    return (__fsub_rn(up, e) < __fsub_rn(e, down)) ? (int)up : (int)down;
  };
  int m0 = nearest(E0), m1 = nearest(E1), m2 = nearest(E2), m3 = nearest(E3);
  const int sum = (m0 + m1 + m2 + m3) / 4;  // C truncation
  const float r0 = __fsub_rn(E0, (float)m0), r1 = __fsub_rn(E1, (float)m1), r2 = __fsub_rn(E2, (float)m2),
              r3 = __fsub_rn(E3, (float)m3);
  // rank[i] = number of j with resid[i] < resid[j] (j > i) or not resid[j] < resid[i] (j < i)
  const int c01 = r0 < r1, c02 = r0 < r2, c03 = r0 < r3, c12 = r1 < r2, c13 = r1 < r3, c23 = r2 < r3;
  int k0 = c01 + c02 + c03 + sum;
  int k1 = (1 - c01) + c12 + c13 + sum;
  int k2 = (1 - c02) + (1 - c12) + c23 + sum;
  int k3 = (1 - c03) + (1 - c13) + (1 - c23) + sum;
  auto wrap = [](int& k, int& m) {
    const int lo = k < 0, hi = k > 3;
    k += 4 * (lo - hi);
    m += 4 * (lo - hi);
  };
  wrap(k0, m0); wrap(k1, m1); wrap(k2, m2); wrap(k3, m3);
  const float d0 = __fmul_rn(__fsub_rn(E0, (float)m0), 0.25f), d1 = __fmul_rn(__fsub_rn(E1, (float)m1), 0.25f),
              d2 = __fmul_rn(__fsub_rn(E2, (float)m2), 0.25f), d3 = __fmul_rn(__fsub_rn(E3, (float)m3), 0.25f);
  // delta of the coordinate whose rank is r (0 if the ranks are not a permutation, as the indexed form would skip it)
  auto by_rank = [&](int r) {  // a chain of selects (the nested conditional compiled to branches)
    float q = k3 == r ? d3 : 0.0f;
    q = k2 == r ? d2 : q;
    q = k1 == r ? d1 : q;
    q = k0 == r ? d0 : q;
    return q;
  };
  const float q0 = by_rank(0), q1 = by_rank(1), q2 = by_rank(2), q3 = by_rank(3);
  // bary[s] += delta where rank == 3 - s, -= delta where rank == 4 - s;  bary[0] += 1 + bary[4]
  const float b4 = __fsub_rn(0.0f, q0);
  weight[0] = __fadd_rn(q3, __fadd_rn(1.0f, b4));
  weight[1] = __fsub_rn(q2, q3);
  weight[2] = __fsub_rn(q1, q2);
  weight[3] = __fsub_rn(q0, q1);
  const uint32_t mask = (1u << log2_capacity) - 1u;
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // vertex r: key_i = rem0_i + r, minus 4 where rank_i > 3 - r (first three coordinates)
    const int key0 = m0 + r - ((k0 > 3 - r) ? 4 : 0);
    const int key1 = m1 + r - ((k1 > 3 - r) ? 4 : 0);
    const int key2 = m2 + r - ((k2 > 3 - r) ? 4 : 0);
    uint32_t h = (uint32_t)key0 * 2531011u;
    h = (h + (uint32_t)key1) * 2531011u;
    h = (h + (uint32_t)key2) * 2531011u;
    row[r] = h & mask;
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
