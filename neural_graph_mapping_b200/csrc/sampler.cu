// Stage kernel: ray sampler.
// Replaces Camera.sample_ijs_uniform (ngm/camera.py:215-292), the depth-guided merge
// (ngm/run_mapping.py:521-545) and utils.transform_points (ngm/utils.py:276-286).
//
// HBM-bound: algorithmic bytes per ray = 16 (ijs) + 8 (near, far) + St*(12 + 4 + 4)  [SURVEY 8d].
// A block works on 32 rays at a time.  Phase 1: one thread per ray computes everything that is
// constant along the ray (unit direction, sampling windows, stratum widths) into shared memory --
// the divisions and the square root are paid once per ray, not once per sample.  Phase 2: each
// warp takes 4 of the rays and its lanes stride over the samples, so every output row is written
// with consecutive addresses.  The reference's sort is replaced by a rank computation: both sample
// sets are already sorted (disjoint strata), so the merged position of an element is its own index
// plus the number of elements of the other set below it, found by inspecting <= 3 candidate strata.
#include "common.cuh"

namespace ngm {

namespace {

constexpr int kRaysPerBlock = 32;
constexpr int kThreads = 256;

struct RayConst {
  float dx, dy, dz;     // unit direction (camera frame)
  float lo, span, delta;     // coarse set: near, far - near, (far - near) / S
  float glo, gspan, gdelta;  // guided set
  float hi, ghi;
  float pad;
};

// (delta * u + linspace(0,1,n+1)[k] * span) + lo  -- camera.py:269-276 with the per-ray terms hoisted
__device__ __forceinline__ float strat(float lo, float span, float delta, float step, int k, int n, float u) {
  const int steps = n + 1;
  const float lin = k < steps / 2 ? __fmul_rn(step, (float)k) : __fsub_rn(1.0f, __fmul_rn(step, (float)(steps - k - 1)));
  return __fadd_rn(__fadd_rn(__fmul_rn(delta, u), __fmul_rn(lin, span)), lo);
}

// number of elements of the OTHER stratified set (n_o strata over [lo_o, lo_o + span_o]) before `d`
template <typename JitterFn>
__device__ __forceinline__ int rank_in_other(float d, float lo_o, float span_o, float delta_o, float step_o, int n_o,
                                             bool strict, JitterFn other_jitter) {
  if (!(span_o > 0.0f)) {
    const float v = strat(lo_o, span_o, delta_o, step_o, 0, n_o, 0.0f);
    return (strict ? (v < d) : (v <= d)) ? n_o : 0;
  }
  const float t = (d - lo_o) / span_o * (float)n_o;
  const int j = (int)floorf(fminf(fmaxf(t, -2.0f), (float)n_o + 2.0f));
  int cnt = max(0, min(j - 1, n_o));
#pragma unroll
  for (int c = -1; c <= 1; ++c) {
    const int jj = j + c;
    if (jj >= 0 && jj < n_o) {
      const float v = strat(lo_o, span_o, delta_o, step_o, jj, n_o, other_jitter(jj));
      cnt += (strict ? (v < d) : (v <= d)) ? 1 : 0;
    }
  }
  return cnt;
}

template <bool VEC4>
__global__ void __launch_bounds__(kThreads) sample_rays_kernel(NgmSampleArgs a) {
  __shared__ RayConst rc[kRaysPerBlock];
  __shared__ float c2w[kRaysPerBlock][12];
  const int S = a.num_samples;
  const int G = a.gt ? a.num_samples_guided : 0;
  const int St = S + G;
  const float step_s = 1.0f / (float)S;
  const float step_g = G > 0 ? 1.0f / (float)G : 0.0f;
  const RayJitter jit{a.jitter, a.jitter_guided, a.seed, a.offset};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long num_groups = (a.num_rays + kRaysPerBlock - 1) / kRaysPerBlock;
  if (!a.c2w_per_ray && threadIdx.x < 12) c2w[0][threadIdx.x] = __ldg(a.c2ws + threadIdx.x);

  for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
    const long long ray0 = grp * kRaysPerBlock;
    __syncthreads();  // previous group fully consumed
    if (threadIdx.x < kRaysPerBlock) {
      const long long ray = ray0 + threadIdx.x;
      if (ray < a.num_rays) {
        RayConst r;
        const float nr = a.near ? __ldg(a.near + ray) : a.near_scalar;
        const float fr = a.far ? __ldg(a.far + ray) : a.far_scalar;
        r.lo = nr; r.hi = fr;
        r.span = __fsub_rn(fr, nr);
        r.delta = r.span / (float)S;
        float glo = nr, ghi = fr;
        if (G > 0) guided_window(nr, fr, __ldg(a.gt + ray), a.range_guided, glo, ghi);
        r.glo = glo; r.ghi = ghi;
        r.gspan = __fsub_rn(ghi, glo);
        r.gdelta = G > 0 ? r.gspan / (float)G : 0.0f;
        const longlong2 ij = __ldg(reinterpret_cast<const longlong2*>(a.ijs) + ray);
        const float3 dir = ij_to_direction(ij.x, ij.y, a.cam);
        r.dx = dir.x; r.dy = dir.y; r.dz = dir.z;
        r.pad = 0.f;
        rc[threadIdx.x] = r;
      }
    } else if (a.c2w_per_ray && a.points_world) {
      // the other threads stage the per-ray matrices: 32 rays x 12 floats
      for (int i = threadIdx.x - kRaysPerBlock; i < kRaysPerBlock * 12; i += kThreads - kRaysPerBlock) {
        const int rl = i / 12, e = i - rl * 12;
        if (ray0 + rl < a.num_rays) c2w[rl][e] = __ldg(a.c2ws + (ray0 + rl) * 16 + e);
      }
    }
    __syncthreads();

    if constexpr (VEC4) {
      // No guided set and St % 4 == 0: a lane owns 4 consecutive samples of one ray, so every output leaves as
      // 16-byte vectors (1.25 store instructions per sample instead of 5, each warp store covering whole sectors).
      // A warp owns 4 consecutive rays of the group and its lanes stride over their St / 4 quads jointly.
      const int qpr = St >> 2;  // quads per ray
      for (int idx = lane; idx < 4 * qpr; idx += 32) {
        const int rl = warp * 4 + idx / qpr, k0 = (idx % qpr) << 2;
        const long long ray = ray0 + rl;
        if (ray >= a.num_rays) continue;
        const RayConst r = rc[rl];
        const float* m = c2w[a.c2w_per_ray ? rl : 0];
        float d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = strat(r.lo, r.span, r.delta, step_s, k0 + i, S, jit.coarse(ray, k0 + i, S, St));
        const long long o = ray * St + k0;
        if (a.distances) *reinterpret_cast<float4*>(a.distances + o) = make_float4(d[0], d[1], d[2], d[3]);
        if (a.depths)
          *reinterpret_cast<float4*>(a.depths + o) = make_float4(-(r.dz * d[0]), -(r.dz * d[1]), -(r.dz * d[2]), -(r.dz * d[3]));
        if (a.points_cam) {
          float v[12];
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[3 * i] = r.dx * d[i]; v[3 * i + 1] = r.dy * d[i]; v[3 * i + 2] = r.dz * d[i]; }
          float4* dst = reinterpret_cast<float4*>(a.points_cam + o * 3);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          dst[2] = make_float4(v[8], v[9], v[10], v[11]);
        }
        if (a.points_world) {
          float v[12];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float3 pw = transform_point(m, make_float3(r.dx * d[i], r.dy * d[i], r.dz * d[i]));
            v[3 * i] = pw.x; v[3 * i + 1] = pw.y; v[3 * i + 2] = pw.z;
          }
          float4* dst = reinterpret_cast<float4*>(a.points_world + o * 3);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          dst[2] = make_float4(v[8], v[9], v[10], v[11]);
        }
      }
    } else {
    for (int rl = warp; rl < kRaysPerBlock; rl += kThreads / 32) {
      const long long ray = ray0 + rl;
      if (ray >= a.num_rays) break;
      const RayConst r = rc[rl];
      const float* m = c2w[a.c2w_per_ray ? rl : 0];
      for (int k = lane; k < St; k += 32) {
        float d;
        int pos;
        if (k < S) {
          d = strat(r.lo, r.span, r.delta, step_s, k, S, jit.coarse(ray, k, S, St));
          pos = k;
          if (G > 0)
            pos += rank_in_other(d, r.glo, r.gspan, r.gdelta, step_g, G, /*strict=*/true,
                                 [&](int j) { return jit.guided(ray, j, S, G, St); });
        } else {
          const int kg = k - S;
          d = strat(r.glo, r.gspan, r.gdelta, step_g, kg, G, jit.guided(ray, kg, S, G, St));
          pos = kg + rank_in_other(d, r.lo, r.span, r.delta, step_s, S, /*strict=*/false,
                                   [&](int j) { return jit.coarse(ray, j, S, St); });
        }
        const float3 pc = make_float3(r.dx * d, r.dy * d, r.dz * d);
        const long long o = ray * St + pos;
        if (a.distances) a.distances[o] = d;
        if (a.depths) a.depths[o] = -pc.z;
        if (a.points_cam) {
          a.points_cam[o * 3 + 0] = pc.x;
          a.points_cam[o * 3 + 1] = pc.y;
          a.points_cam[o * 3 + 2] = pc.z;
        }
        if (a.points_world) {
          const float3 pw = transform_point(m, pc);
          a.points_world[o * 3 + 0] = pw.x;
          a.points_world[o * 3 + 1] = pw.y;
          a.points_world[o * 3 + 2] = pw.z;
        }
      }
    }
    }
  }
}

}  // namespace

int launch_sample_rays(const NgmSampleArgs& a, cudaStream_t stream) {
  if (a.num_rays == 0) return NGM_OK;
  long long blocks = (a.num_rays + kRaysPerBlock - 1) / kRaysPerBlock;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const int G = a.gt ? a.num_samples_guided : 0;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec4 = G == 0 && a.num_samples % 4 == 0 && aligned16(a.distances) && aligned16(a.depths) &&
                    aligned16(a.points_cam) && aligned16(a.points_world);
  if (vec4) sample_rays_kernel<true><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
  else      sample_rays_kernel<false><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
  return check_launch("sample_rays_kernel");
}

}  // namespace ngm
