// Stage kernel: compositor.
// Replaces the post-MLP split/masks (ngm/run_mapping.py:610-639) and
// NeuralGraphMap._quadrature (ngm/run_mapping.py:709-799).
//
// HBM-bound: algorithmic bytes per ray = S*(4*b + 4 + 4) + 36  [SURVEY 8d].
// One warp per ray; lane l owns samples l, l+32, ... so every load is a contiguous 128 B
// (distances, depths) or 512 B (packed rgb+geometry float4) row segment.  Transmittance is a
// warp-shuffle exclusive product scan with a running carry between 32-sample blocks; the
// expectations and second moments are shuffle reductions (one pass; variances from the moments).
#include "common.cuh"

namespace ngm {

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float fast_sig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct SampleTerms {
  float w, z, c0, c1, c2;
};

struct RayCtx {
  const NgmCompositeArgs& a;
  long long ray;
  int lane, S, Se;
  float isd_gamma;  // neus: isd * geometry_factor
  float gt;
  bool has_gt;

  // geometry of sample k after the behind-camera overwrite (run_mapping.py:614-622)
  __device__ __forceinline__ float geometry(int k, float depth) const {
    float g = __ldg(a.geometries + (ray * S + k) * a.geometry_stride);
    if (a.overwrite_behind_camera && depth < 0.0f)
      g = (a.geometry_mode == NGM_GEOM_OCCUPANCY || a.geometry_mode == NGM_GEOM_DENSITY) ? -100.0f : 1.0f;
    return g;
  }

  // occupancy probability of sample k (run_mapping.py:746-762); k < Se
  __device__ __forceinline__ float occupancy(int k, float g, float dist) const {
    switch (a.geometry_mode) {
      case NGM_GEOM_NRGBD: {
        // 4 s(t) s(-t) = 4u / (1 + u)^2 with u = exp(-|t|): one exponential, no overflow
        // MUFU-based exp / reciprocal: relative error ~2^-21, far inside the fp32 parity tolerance
        const float u = __expf(-fabsf(a.geometry_factor * g));
        const float q = 1.0f + u;
        return __fdividef(4.0f * u, q * q);
      }
      case NGM_GEOM_OCCUPANCY:
        return fast_sig(a.geometry_factor * g);
      case NGM_GEOM_DENSITY: {
        float dn = __ldg(a.distances + ray * S + k + 1);
        float delta = dn - dist;
        return 1.0f - __expf(-delta * fmaxf(g, 0.0f));
      }
      default: {  // NGM_GEOM_NEUS
        float zn = __ldg(a.depths + ray * S + k + 1);
        float gn = geometry(k + 1, zn);
        float t0 = fast_sig(isd_gamma * g), t1 = fast_sig(isd_gamma * gn);
        return fmaxf(__fdividef(t0 - t1, t0 + 1e-5f), 0.0f);
      }
    }
  }
};

struct RawSample {
  float c0, c1, c2, g, dist, z;
};

// global loads of one 32-sample block (issued one block ahead of their use)
template <bool PACKED>
__device__ __forceinline__ RawSample load_block(const RayCtx& c, int blk) {
  const NgmCompositeArgs& a = c.a;
  const int k = blk * 32 + c.lane;
  RawSample r{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (k < c.S) {
    const long long idx = c.ray * c.S + k;
    r.dist = __ldg(a.distances + idx);
    r.z = __ldg(a.depths + idx);
    if (PACKED) {  // rgb + geometry of the MLP output in one 16-byte load
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.colors) + idx);
      r.c0 = v.x; r.c1 = v.y; r.c2 = v.z; r.g = v.w;
    } else {
      const float* col = a.colors + idx * a.color_stride;
      r.c0 = __ldg(col); r.c1 = __ldg(col + 1); r.c2 = __ldg(col + 2);
      r.g = __ldg(a.geometries + idx * a.geometry_stride);
    }
    if (a.overwrite_behind_camera && r.z < 0.0f)
      r.g = (a.geometry_mode == NGM_GEOM_OCCUPANCY || a.geometry_mode == NGM_GEOM_DENSITY) ? -100.0f : 1.0f;
  }
  return r;
}

// One 32-sample block: aux outputs, occupancy, scan.  `carry` = transmittance entering the block,
// updated on exit.
__device__ __forceinline__ SampleTerms eval_block(const RayCtx& c, int blk, const RawSample& r, float& carry) {
  const NgmCompositeArgs& a = c.a;
  const int k = blk * 32 + c.lane;
  SampleTerms t{0.f, 0.f, 0.f, 0.f, 0.f};
  float occ = 0.0f;
  if (k < c.S) {
    const long long idx = c.ray * c.S + k;
    if (a.freespace) {  // run_mapping.py:624-630
      float thr = c.has_gt ? (c.gt - a.truncation) * (c.gt != 0.0f ? 1.0f : 0.0f) : 0.0f;
      a.freespace[idx] = r.g * a.truncation;
      a.freespace_mask[idx] = (c.has_gt && r.dist < thr) ? 1 : 0;
    }
    if (a.tsdf) {  // run_mapping.py:632-639
      float delta = c.gt - r.dist;
      a.tsdf[idx] = r.g * a.truncation - delta;
      a.tsdf_mask[idx] = (c.has_gt && fabsf(delta) < a.truncation && c.gt != 0.0f) ? 1 : 0;
    }
    if (k < c.Se) {
      occ = c.occupancy(k, r.g, r.dist);
      t.z = r.z;
      t.c0 = a.color_factor * r.c0; t.c1 = a.color_factor * r.c1; t.c2 = a.color_factor * r.c2;
    }
  }
  // exclusive product scan of (1 - occ) across the warp, times carry (run_mapping.py:764-771)
  float incl = 1.0f - occ;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (c.lane >= o) incl *= n;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (c.lane == 0) excl = 1.0f;
  t.w = occ * (carry * excl);
  carry *= __shfl_sync(0xffffffffu, incl, 31);
  return t;
}

// One pass per ray: per-lane partial moments over the ray's 32-sample blocks (loads issued one block
// ahead), then nine warp reductions.  The variances use  Sum w (m - x)^2 = Sum w x^2 - m^2 (2 - Sum w)
// (exact algebra; the fp32 cancellation error is ~1e-7 of the colour / depth scale).
template <bool PACKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5) composite_kernel(NgmCompositeArgs a) {
  const int lane = threadIdx.x & 31;
  // broadcast -> provably warp-uniform ray index, so the shuffles below are emitted without
  // per-instruction WARPSYNC / ENDCOLLECTIVE convergence wrappers
  const int warp_in_block = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + warp_in_block;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int S = a.num_samples;
  const bool drop_last = (a.geometry_mode == NGM_GEOM_DENSITY || a.geometry_mode == NGM_GEOM_NEUS);
  const int Se = drop_last ? S - 1 : S;
  const int nblk = (S + 31) / 32;
  for (long long ray = warp0; ray < a.num_rays; ray += nwarps) {
    RayCtx c{a, ray, lane, S, Se, 0.f, 0.f, a.gt != nullptr};
    if (a.geometry_mode == NGM_GEOM_NEUS)
      c.isd_gamma = __ldg(a.neus_isd + ray / a.rays_per_isd) * a.geometry_factor;
    if (c.has_gt) c.gt = __ldg(a.gt + ray);
    float carry = 1.0f;
    float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    RawSample cur = load_block<PACKED>(c, 0);
    for (int b = 0; b < nblk; ++b) {
      RawSample nxt = cur;
      if (b + 1 < nblk) nxt = load_block<PACKED>(c, b + 1);
      const SampleTerms t = eval_block(c, b, cur, carry);
      cur = nxt;
      const float wz = t.w * t.z, w0 = t.w * t.c0, w1 = t.w * t.c1, w2 = t.w * t.c2;
      m[0] += t.w; m[1] += wz; m[2] += w0; m[3] += w1; m[4] += w2;
      m[5] += wz * t.z; m[6] += w0 * t.c0; m[7] += w1 * t.c1; m[8] += w2 * t.c2;
      if (a.weights) {
        const int k = b * 32 + lane;
        if (k < S) a.weights[ray * S + k] = t.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = warp_sum(m[i]);
    if (lane == 0) {
      const float P = m[0], D = m[1], C0 = m[2], C1 = m[3], C2 = m[4];
      const float t2 = 2.0f - P;
      reinterpret_cast<float4*>(a.rgbd)[ray] = make_float4(C0, C1, C2, D);
      if (a.color_var) {
        a.color_var[ray * 3 + 0] = fmaxf(fmaf(-C0 * C0, t2, m[6]), 0.0f);
        a.color_var[ray * 3 + 1] = fmaxf(fmaf(-C1 * C1, t2, m[7]), 0.0f);
        a.color_var[ray * 3 + 2] = fmaxf(fmaf(-C2 * C2, t2, m[8]), 0.0f);
      }
      if (a.depth_var) a.depth_var[ray] = fmaxf(fmaf(-D * D, t2, m[5]), 0.0f);
      // term prob = 1 - bg = 1 - (1 - sum w)  (run_mapping.py:774,796)
      if (a.term_prob) a.term_prob[ray] = 1.0f - (1.0f - P);
    }
  }
}

__global__ void neus_isd_kernel(const float* __restrict__ sd, const long long* __restrict__ slots, int n,
                                float* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n) out[f] = 1.0f / fabsf(sd[slots ? slots[f] : f]);
}

}  // namespace

int launch_neus_isd(const float* sd, const int64_t* slots, int num_fields, float* out, cudaStream_t stream) {
  if (num_fields == 0) return NGM_OK;
  neus_isd_kernel<<<(num_fields + 127) / 128, 128, 0, stream>>>(sd, reinterpret_cast<const long long*>(slots),
                                                                num_fields, out);
  return check_launch("neus_isd_kernel");
}

int launch_composite(const NgmCompositeArgs& a, cudaStream_t stream) {
  if (a.num_rays == 0) return NGM_OK;
  long long blocks = (a.num_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)num_sms() * 64;
  if (blocks > cap) blocks = cap;
  const bool packed = a.color_stride == 4 && a.geometry_stride == 4 && a.geometries == a.colors + 3 &&
                      (reinterpret_cast<uintptr_t>(a.colors) & 15) == 0;
  if (packed) composite_kernel<true><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(a);
  else composite_kernel<false><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(a);
  return check_launch("composite_kernel");
}

}  // namespace ngm
