// Stage kernel: compositor.
// Replaces the post-MLP split/masks (ngm/run_mapping.py:610-639) and
// NeuralGraphMap._quadrature (ngm/run_mapping.py:709-799).
//
// HBM-bound: algorithmic bytes per ray = S*(4*b + 4 + 4) + 36  [SURVEY 8d].
// One warp per ray; lane l owns samples l, l+32, ... so every load is a contiguous 128 B
// (distances, depths) or 512 B (packed rgb+geometry float4) row segment.  Transmittance is a
// warp-shuffle exclusive product scan with a running carry between 32-sample blocks; the
// expectations and second moments are shuffle reductions (one pass; variances from the moments).
#include <stdlib.h>

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
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5) composite_kernel(NgmCompositeArgs a_in) {
  NgmCompositeArgs a = a_in;
  a.overwrite_behind_camera = overwrite_enabled(a_in.overwrite_behind_camera, a_in.overwrite_gate);
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

// ---- staged variant: one THREAD per ray, samples streamed through shared memory --------------------
// The warp-per-ray kernel above is issue-bound (a 32-lane scan and nine 5-step shuffle reductions per
// ray: ~480 warp instructions per ray, ncu: 81% issue-active at 35% of DRAM peak).  Here a warp owns
// 32 consecutive rays; their samples arrive in chunks of 8 through a ring of cp.async (LDGSTS) stages,
// laid out so that both the global side (whole 128-B / 32-B row segments) and the shared side
// (XOR-swizzled 16-B pieces, one row per lane) are conflict-free, and every lane composites its own
// ray sequentially: no shuffles, ~25 instructions per sample, transmittance as a running product.
// Packed (N,S,4) MLP output only, S % 4 == 0, no per-sample outputs (weights / aux): the general
// kernel keeps those.
constexpr int kChunk = 8;                      // samples per stage
constexpr int kStages = 3;                     // ring depth (2 chunks in flight per warp)
constexpr int kStagedWarps = 4;
constexpr int kStageBytes = 32 * kChunk * 16 + 2 * 32 * kChunk * 4;  // colours+geometry, distances, depths = 6 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Moments {
  float T, m[9];
};

// front-to-back accumulation of one sample (run_mapping.py:764-799)
__device__ __forceinline__ void accumulate(Moments& a, float occ, float z, float c0, float c1, float c2) {
  const float w = occ * a.T;
  a.T *= 1.0f - occ;
  const float wz = w * z, w0 = w * c0, w1 = w * c1, w2 = w * c2;
  a.m[0] += w; a.m[1] += wz; a.m[2] += w0; a.m[3] += w1; a.m[4] += w2;
  a.m[5] = fmaf(wz, z, a.m[5]); a.m[6] = fmaf(w0, c0, a.m[6]); a.m[7] = fmaf(w1, c1, a.m[7]); a.m[8] = fmaf(w2, c2, a.m[8]);
}

template <int MODE>
__global__ void __launch_bounds__(kStagedWarps * 32) composite_staged_kernel(NgmCompositeArgs a_in) {
  NgmCompositeArgs a = a_in;
  a.overwrite_behind_camera = overwrite_enabled(a_in.overwrite_behind_camera, a_in.overwrite_gate);
  extern __shared__ __align__(16) uint8_t staged_smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  uint8_t* ring = staged_smem + warp_in_block * (kStages * kStageBytes);
  const int S = a.num_samples;
  const int nchunk = (S + kChunk - 1) / kChunk;
  const long long ntiles = (a.num_rays + 31) / 32;
  const long long warp0 = blockIdx.x * (long long)kStagedWarps + warp_in_block;
  const long long nwarps = (long long)gridDim.x * kStagedWarps;
  if (warp0 >= ntiles) return;
  const long long my_tiles = (ntiles - warp0 + nwarps - 1) / nwarps;
  const long long total_q = my_tiles * nchunk;
  constexpr bool kLag = (MODE == NGM_GEOM_DENSITY || MODE == NGM_GEOM_NEUS);  // need the next sample; last one dropped
  const float4* colors = reinterpret_cast<const float4*>(a.colors);
  const float fill = (MODE == NGM_GEOM_OCCUPANCY || MODE == NGM_GEOM_DENSITY) ? -100.0f : 1.0f;

  // stage q of this warp's (tile, chunk) sequence -> ring slot q % kStages
  auto issue = [&](long long q) {
    if (q < total_q) {
      const long long ray0 = (warp0 + (q / nchunk) * nwarps) * 32;
      const int k0 = (int)(q % nchunk) * kChunk;
      uint8_t* st = ring + (int)(q % kStages) * kStageBytes;
      // colours: 32 rows x 8 pieces of 16 B; 8 lanes cover one row's 128 B, piece p of row r lands at p ^ (r & 7)
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        const int row = i * 4 + (lane >> 3), piece = lane & 7;
        const long long ray = ray0 + row;
        const bool ok = ray < a.num_rays && k0 + piece < S;
        const float4* src = colors + (ok ? ray * S + k0 + piece : 0);
        cp_async16(st + row * 128 + ((piece ^ (row & 7)) << 4), src, ok ? 16 : 0);
      }
      // distances / depths: 32 rows x 2 pieces of 16 B each; piece p of row r lands at p ^ ((r >> 2) & 1)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = i * 16 + (lane >> 1), piece = lane & 1;
        const long long ray = ray0 + row;
        const int k = k0 + piece * 4;
        const int nvalid = ray < a.num_rays ? min(max(S - k, 0), 4) : 0;
        const long long off = nvalid > 0 ? ray * S + k : 0;
        const int dst = row * 32 + ((piece ^ ((row >> 2) & 1)) << 4);
        cp_async16(st + 32 * kChunk * 16 + dst, a.distances + off, nvalid * 4);
        cp_async16(st + 32 * kChunk * 16 + 32 * kChunk * 4 + dst, a.depths + off, nvalid * 4);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int q = 0; q < kStages - 1; ++q) issue(q);

  Moments acc;
  float isd_gamma = 0.0f;
  float pc0 = 0.f, pc1 = 0.f, pc2 = 0.f, pg = 0.f, pd = 0.f, pz = 0.f;  // kLag: the sample waiting for its successor
  for (long long q = 0; q < total_q; ++q) {
    const int chunk = (int)(q % nchunk);
    const long long ray = (warp0 + (q / nchunk) * nwarps) * 32 + lane;
    if (chunk == 0) {
      acc.T = 1.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) acc.m[i] = 0.0f;
      if (MODE == NGM_GEOM_NEUS)
        isd_gamma = ray < a.num_rays ? __ldg(a.neus_isd + ray / a.rays_per_isd) * a.geometry_factor : 0.0f;
    }
    issue(q + kStages - 1);
    cp_async_wait<kStages - 1>();
    __syncwarp();
    const uint8_t* st = ring + (int)(q % kStages) * kStageBytes;
    float dist[kChunk], depth[kChunk];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int o = lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4);
      const float4 d4 = *reinterpret_cast<const float4*>(st + 32 * kChunk * 16 + o);
      const float4 z4 = *reinterpret_cast<const float4*>(st + 32 * kChunk * 16 + 32 * kChunk * 4 + o);
      dist[4 * j] = d4.x; dist[4 * j + 1] = d4.y; dist[4 * j + 2] = d4.z; dist[4 * j + 3] = d4.w;
      depth[4 * j] = z4.x; depth[4 * j + 1] = z4.y; depth[4 * j + 2] = z4.z; depth[4 * j + 3] = z4.w;
    }
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      const int k = chunk * kChunk + j;
      if (k < S) {  // warp-uniform
        const float4 v = *reinterpret_cast<const float4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4));
        const float z = depth[j], d = dist[j];
        float g = v.w;
        if (a.overwrite_behind_camera && z < 0.0f) g = fill;
        const float c0 = a.color_factor * v.x, c1 = a.color_factor * v.y, c2 = a.color_factor * v.z;
        if (!kLag) {
          float occ;
          if (MODE == NGM_GEOM_NRGBD) {  // 4 s(t) s(-t) = 4u / (1 + u)^2, u = exp(-|t|)
            const float u = __expf(-fabsf(a.geometry_factor * g));
            const float qq = 1.0f + u;
            occ = __fdividef(4.0f * u, qq * qq);
          } else {
            occ = fast_sig(a.geometry_factor * g);
          }
          accumulate(acc, occ, z, c0, c1, c2);
        } else {
          if (k > 0) {  // sample k-1, now that its successor is known
            float occ;
            if (MODE == NGM_GEOM_DENSITY) {
              occ = 1.0f - __expf(-(d - pd) * fmaxf(pg, 0.0f));
            } else {
              const float t0 = fast_sig(isd_gamma * pg), t1 = fast_sig(isd_gamma * g);
              occ = fmaxf(__fdividef(t0 - t1, t0 + 1e-5f), 0.0f);
            }
            accumulate(acc, occ, pz, pc0, pc1, pc2);
          }
          pc0 = c0; pc1 = c1; pc2 = c2; pg = g; pd = d; pz = z;
        }
      }
    }
    __syncwarp();  // every lane is done with this ring slot before the next issue() refills it
    if (chunk == nchunk - 1 && ray < a.num_rays) {
      const float P = acc.m[0], D = acc.m[1], C0 = acc.m[2], C1 = acc.m[3], C2 = acc.m[4];
      const float t2 = 2.0f - P;
      const float4 out4 = make_float4(C0, C1, C2, D);
      const float dv = fmaxf(fmaf(-D * D, t2, acc.m[5]), 0.0f), tp = 1.0f - (1.0f - P);
      reinterpret_cast<float4*>(a.rgbd)[ray] = out4;
      if (a.color_var) {
        a.color_var[ray * 3 + 0] = fmaxf(fmaf(-C0 * C0, t2, acc.m[6]), 0.0f);
        a.color_var[ray * 3 + 1] = fmaxf(fmaf(-C1 * C1, t2, acc.m[7]), 0.0f);
        a.color_var[ray * 3 + 2] = fmaxf(fmaf(-C2 * C2, t2, acc.m[8]), 0.0f);
      }
      if (a.depth_var) a.depth_var[ray] = dv;
      if (a.term_prob) a.term_prob[ray] = tp;
      // fused multi-GPU tile exchange (NgmCompositeArgs.mirror_delta): the warp's 32 rays again, into the peer /
      // multicast mappings -- 512 B of rgbd and 128 B each of depth variance and termination probability straight from
      // the registers (coalesced), the 384 B of colour variances re-read below as 16-byte pieces
#pragma unroll
      for (int m = 0; m < NGM_MAX_MIRRORS; ++m) {
        if (m < a.num_mirrors) {
          const long long dl = a.mirror_delta[m];
          reinterpret_cast<float4*>(reinterpret_cast<char*>(a.rgbd) + dl)[ray] = out4;
          if (a.depth_var) reinterpret_cast<float*>(reinterpret_cast<char*>(a.depth_var) + dl)[ray] = dv;
          if (a.term_prob) reinterpret_cast<float*>(reinterpret_cast<char*>(a.term_prob) + dl)[ray] = tp;
        }
      }
    }
    if (chunk == nchunk - 1 && a.num_mirrors > 0 && a.color_var) {  // warp-uniform
      __syncwarp();  // the warp's colour variances above are visible to all its lanes
      const long long ray0 = ray - lane;
      const long long n_valid = a.num_rays - ray0 < 32 ? a.num_rays - ray0 : 32;
      float* cv = a.color_var + ray0 * 3;
      if (n_valid == 32 && (reinterpret_cast<uintptr_t>(cv) & 15) == 0) {
        if (lane < 24) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(cv) + lane);
#pragma unroll
          for (int m = 0; m < NGM_MAX_MIRRORS; ++m)
            if (m < a.num_mirrors)
              reinterpret_cast<float4*>(reinterpret_cast<char*>(cv) + a.mirror_delta[m])[lane] = v;
        }
      } else {
        for (int j = lane; j < n_valid * 3; j += 32) {
          const float v = __ldcg(cv + j);
#pragma unroll
          for (int m = 0; m < NGM_MAX_MIRRORS; ++m)
            if (m < a.num_mirrors) reinterpret_cast<float*>(reinterpret_cast<char*>(cv) + a.mirror_delta[m])[j] = v;
        }
      }
    }
  }
  cp_async_wait<0>();
}

template <int MODE>
int launch_staged(const NgmCompositeArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)kStagedWarps * kStages * kStageBytes;
  NGM_CUDA(cudaFuncSetAttribute(composite_staged_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (a.num_rays + 31) / 32;
  long long blocks = (ntiles + kStagedWarps - 1) / kStagedWarps;
  const long long cap = (long long)num_sms() * 24;  // 3 resident CTAs per SM x 8 rounds; beyond that warps loop over tiles
  if (blocks > cap) blocks = cap;
  composite_staged_kernel<MODE><<<(unsigned)blocks, kStagedWarps * 32, smem, stream>>>(a);
  return check_launch("composite_staged_kernel");
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

// NGM_COMPOSITE_STAGED=0 forces the general warp-per-ray kernel (tests compare the two)
static bool staged_disabled() {
  const char* e = getenv("NGM_COMPOSITE_STAGED");
  return e && e[0] == '0';
}

int launch_composite(const NgmCompositeArgs& a, cudaStream_t stream) {
  if (a.num_rays == 0) return NGM_OK;
  long long blocks = (a.num_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)num_sms() * 64;
  if (blocks > cap) blocks = cap;
  const bool packed = a.color_stride == 4 && a.geometry_stride == 4 && a.geometries == a.colors + 3 &&
                      (reinterpret_cast<uintptr_t>(a.colors) & 15) == 0;
  const bool staged = packed && a.num_samples % 4 == 0 && !a.weights && !a.freespace && !a.tsdf &&
                      ((reinterpret_cast<uintptr_t>(a.distances) | reinterpret_cast<uintptr_t>(a.depths)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.rgbd) & 15) == 0 && !staged_disabled();
  if (staged) {
    switch (a.geometry_mode) {
      case NGM_GEOM_NRGBD: return launch_staged<NGM_GEOM_NRGBD>(a, stream);
      case NGM_GEOM_OCCUPANCY: return launch_staged<NGM_GEOM_OCCUPANCY>(a, stream);
      case NGM_GEOM_DENSITY: return launch_staged<NGM_GEOM_DENSITY>(a, stream);
      case NGM_GEOM_NEUS: return launch_staged<NGM_GEOM_NEUS>(a, stream);
      default: break;
    }
  }
  if (a.num_mirrors > 0) {
    set_error("mirrored Prediction stores need the packed-input compositor (strides (4,4), num_samples %% 4 == 0, 16-byte "
              "aligned inputs, no weights / aux outputs, NGM_COMPOSITE_STAGED not 0)");
    return NGM_ERR_UNSUPPORTED;
  }
  if (packed) composite_kernel<true><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(a);
  else composite_kernel<false><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(a);
  return check_launch("composite_kernel");
}

}  // namespace ngm
