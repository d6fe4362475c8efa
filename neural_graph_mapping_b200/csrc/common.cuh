// Shared helpers for libngm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ngm_b200.h"

namespace ngm {

void set_error(const char* fmt, ...);

#define NGM_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ngm::set_error(__VA_ARGS__);          \
      return NGM_ERR_INVALID_ARG;           \
    }                                       \
  } while (0)

#define NGM_UNSUPPORTED(cond, ...)          \
  do {                                      \
    if (cond) {                             \
      ngm::set_error(__VA_ARGS__);          \
      return NGM_ERR_UNSUPPORTED;           \
    }                                       \
  } while (0)

#define NGM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ngm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),       \
                     __FILE__, __LINE__);                                           \
      return NGM_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

void count_launch();

inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return NGM_ERR_CUDA;
  }
  return NGM_OK;
}

int num_sms();

// behind-camera overwrite (run_mapping.py:614-622) with the reference's gate (:494-495): on only when requested AND
// (no gate given, or the device flag "some near distance is negative" is set)
__device__ __forceinline__ int overwrite_enabled(int requested, const int32_t* gate) {
  return requested && (gate == nullptr || __ldg(gate) != 0);
}

// internal: fp16 A-operand rows of the permutohedral encoding for the tcgen05 renderer (encode.cu)
struct PermutoRowsArgs {
  NgmFieldDesc field;
  const float* points_world;   // (num_points, 3)
  const float* positions;      // field table
  const float* orientations;
  const long long* field_slots;
  uint32_t* out;               // (num_points, EP / 2) half2 words
  long long num_points, points_per_field;
  float field_radius;
  int scale_mode, EP;
  // gather mode (kNN path): row e = (point e / knn_k, neighbour e % knn_k), evaluated by field pair_field[e]
  // (-1: the point is outside every field, no row); positions / orientations are then indexed by that field
  const int* pair_field;
  int knn_k;
};

// ---- Philox4x32-10 counter RNG (in-kernel sampling jitter when no jitter tensor is given) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// U[0,1) for element `index` of stream (seed, offset); 24-bit mantissa, never 1.0.
// Counter-based like Philox but ~6x cheaper (two rounds of the lowbias32 integer hash keyed by the
// 64-bit seed): stratified-sampling jitter needs decorrelation, not cryptographic quality, and the
// fused render kernel draws one value per sample point on its critical path.  The stage sampler
// and the fused kernel share this function, so a seed gives the same samples on every path.
__device__ __forceinline__ uint32_t lowbias32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t offset, uint64_t index) {
  const uint64_t c = offset + index;
  uint32_t h = lowbias32((uint32_t)c ^ (uint32_t)seed);
  h = lowbias32(h + (uint32_t)(c >> 32) * 0x9E3779B1u + (uint32_t)(seed >> 32));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// ---- ray geometry shared by the sampler stage and the fused renderer ------------------------
// Camera.ijs_to_directions, OpenGL convention (ngm/camera.py:186-203).
__device__ __forceinline__ float3 ij_to_direction(long long i, long long j, const NgmCamera& cam) {
  float dx = ((float)j - cam.cx0) / cam.fx;
  float dy = -(((float)i - cam.cy0) / cam.fy);
  float dz = -1.0f;
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1.0f));
  n = fmaxf(n, 1e-12f);  // F.normalize eps
  return make_float3(dx / n, dy / n, dz / n);
}

// torch.linspace(0, 1, n + 1)[idx] (CUDA formula: symmetric about the midpoint).
__device__ __forceinline__ float linspace01(int idx, int n) {
  float step = 1.0f / (float)n;
  int steps = n + 1;
  return idx < steps / 2 ? __fmul_rn(step, (float)idx) : __fsub_rn(1.0f, __fmul_rn(step, (float)(steps - idx - 1)));
}

// Stratified distance k of n in [lo, hi] with jitter u (ngm/camera.py:269-276):
//   (delta * u + linspace(0,1,n+1)[k] * (hi - lo)) + lo,  delta = (hi - lo) / n
__device__ __forceinline__ float stratified_distance(float lo, float hi, int k, int n, float u) {
  float span = __fsub_rn(hi, lo);
  float delta = span / (float)n;
  float b = __fmul_rn(linspace01(k, n), span);
  return __fadd_rn(__fadd_rn(__fmul_rn(delta, u), b), lo);
}

// Depth-guided window (ngm/run_mapping.py:521-530).
__device__ __forceinline__ void guided_window(float nr, float fr, float gt, float range, float& lo, float& hi) {
  bool fallback = (gt == 0.0f) || (nr > gt) || (fr < gt);
  lo = fallback ? nr : __fsub_rn(gt, range);
  hi = fallback ? fr : __fadd_rn(gt, range);
}

struct RayJitter {
  const float* jitter;         // (rays, S) or nullptr
  const float* jitter_guided;  // (rays, G) or nullptr
  uint64_t seed, offset;
  __device__ __forceinline__ float coarse(long long ray, int k, int S, int St) const {
    return jitter ? __ldg(jitter + ray * S + k) : hash_uniform(seed, offset, (uint64_t)ray * St + k);
  }
  __device__ __forceinline__ float guided(long long ray, int k, int S, int G, int St) const {
    return jitter_guided ? __ldg(jitter_guided + ray * G + k)
                         : hash_uniform(seed, offset, (uint64_t)ray * St + S + k);
  }
};

// Position of element `k` of one sorted stratified set inside the merge with the other set
// (the reference concatenates coarse+guided and torch.sort()s, run_mapping.py:540-545).
// Returns the number of elements of the OTHER set that sort before `d`.
//   other set: n_o strata over [lo_o, hi_o]; `strict` = count other < d (else other <= d).
template <typename JitterFn>
__device__ __forceinline__ int count_before(float d, float lo_o, float hi_o, int n_o, bool strict, JitterFn other_jitter) {
  float span = __fsub_rn(hi_o, lo_o);
  if (!(span > 0.0f)) {
    // degenerate other set: all elements equal lo_o (or NaN); compare once
    float v = stratified_distance(lo_o, hi_o, 0, n_o, 0.0f);
    bool before = strict ? (v < d) : (v <= d);
    return before ? n_o : 0;
  }
  float t = (d - lo_o) / span * (float)n_o;
  int j = (int)floorf(fminf(fmaxf(t, -2.0f), (float)n_o + 2.0f));
  int cnt = max(0, min(j - 1, n_o));  // strata < j-1 lie entirely below d
#pragma unroll
  for (int c = -1; c <= 1; ++c) {
    int jj = j + c;
    if (jj >= 0 && jj < n_o) {
      float v = stratified_distance(lo_o, hi_o, jj, n_o, other_jitter(jj));
      cnt += (strict ? (v < d) : (v <= d)) ? 1 : 0;
    }
  }
  return cnt;
}

// world = R p + t  (ngm/utils.py:283-286); m = row-major 4x4
__device__ __forceinline__ float3 transform_point(const float* __restrict__ m, float3 p) {
  return make_float3(__fadd_rn(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[3]),
                     __fadd_rn(m[4] * p.x + m[5] * p.y + m[6] * p.z, m[7]),
                     __fadd_rn(m[8] * p.x + m[9] * p.y + m[10] * p.z, m[11]));
}

// q^-1 (x) v  for a real-first quaternion q (pytorch3d quaternion_invert + quaternion_apply,
// used at ngm/models.py:331-335): vector part of  q* (0,v) q.
__device__ __forceinline__ float3 quat_inv_rotate(float qw, float qx, float qy, float qz, float3 v) {
  // a = q^-1 = (qw,-qx,-qy,-qz);  t = a * (0,v);  out = t * conj(a) = t * q
  float ax = -qx, ay = -qy, az = -qz, aw = qw;
  float tw = -ax * v.x - ay * v.y - az * v.z;
  float tx = aw * v.x + ay * v.z - az * v.y;
  float ty = aw * v.y - ax * v.z + az * v.x;
  float tz = aw * v.z + ax * v.y - ay * v.x;
  // multiply by conj(a) = (aw,-ax,-ay,-az)
  float bx = -ax, by = -ay, bz = -az, bw = aw;
  return make_float3(tw * bx + tx * bw + ty * bz - tz * by,
                     tw * by - tx * bz + ty * bw + tz * bx,
                     tw * bz + tx * by - ty * bx + tz * bw);
}

__device__ __forceinline__ float3 scale_local(float3 x, int scale_mode, float radius) {
  if (scale_mode == NGM_SCALE_UNIT_CUBE) {
    float s = 2.0f * radius;
    return make_float3(x.x / s + 0.5f, x.y / s + 0.5f, x.z / s + 0.5f);
  }
  if (scale_mode == NGM_SCALE_UNIT_BALL) return make_float3(x.x / radius, x.y / radius, x.z / radius);
  return x;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace ngm
