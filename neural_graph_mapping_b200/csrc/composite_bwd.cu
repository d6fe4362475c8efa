// Stage kernel: compositor backward.
// The gradient torch.autograd derives from the post-MLP split (ngm/run_mapping.py:610-639) and
// NeuralGraphMap._quadrature (ngm/run_mapping.py:709-799) when the training step calls
// loss.backward() (ngm/run_mapping.py:1186): d loss / d (sample colours, sample geometries, neus isd)
// from d loss / d (colour, depth, colour variance, depth variance, termination probability, the
// free-space and TSDF sample terms).
//
// With  w_k = o_k T_k,  T_k = prod_{j<k} (1 - o_j),  P = sum w,  C = sum w c,  D = sum w z,
//       V_C = sum w (C - c)^2,  V_D = sum w (D - z)^2  (all sums over the Se composited samples):
//   dL/dc_k = w_k [ gC + gV_C (2 (c_k - C) + 2 C (P - 1)) ]                (dV_C/dC = 2 C (P - 1))
//   dL/dw_k = gP + gC.c_k + gD z_k + gV_C.((C - c_k)^2 + 2 C (P - 1) c_k) + gV_D ((D - z_k)^2 + 2 D (P - 1) z_k)
//   dL/do_k = T_k (dL/dw_k - R_k),   R_k = sum_{j>k} dL/dw_j o_j prod_{k<i<j} (1 - o_i)
//           -> backward recurrence  R_{k-1} = dL/dw_k o_k + (1 - o_k) R_k   (no division by 1 - o_k,
//              so saturated samples, o = 1, are exact)
// and dL/dg from dL/do by the occupancy model of the geometry mode.  One thread per ray: a forward
// sweep stores (o_k, T_k) in the workspace and accumulates P, C, D; a backward sweep applies the
// recurrence.  Training batches are small (tens of thousands of rays x tens of samples), so the
// kernel is latency-, not bandwidth-critical.
#include "common.cuh"

namespace ngm {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct BwdRay {
  const NgmCompositeArgs& a;
  long long ray;
  int S;
  float fill;

  __device__ __forceinline__ float depth(int k) const { return __ldg(a.depths + ray * S + k); }
  __device__ __forceinline__ float dist(int k) const { return __ldg(a.distances + ray * S + k); }
  __device__ __forceinline__ bool overwritten(int k) const { return a.overwrite_behind_camera && depth(k) < 0.0f; }
  // geometry after the behind-camera overwrite (run_mapping.py:614-622)
  __device__ __forceinline__ float geometry(int k) const {
    return overwritten(k) ? fill : __ldg(a.geometries + (ray * S + k) * a.geometry_stride);
  }
  __device__ __forceinline__ float color(int k, int c) const {
    return a.color_factor * __ldg(a.colors + (ray * S + k) * a.color_stride + c);
  }
};

__global__ void __launch_bounds__(128) composite_bwd_kernel(NgmCompositeBwdArgs b_in) {
  NgmCompositeBwdArgs b = b_in;
  b.fwd.overwrite_behind_camera = overwrite_enabled(b_in.fwd.overwrite_behind_camera, b_in.fwd.overwrite_gate);
  const NgmCompositeArgs& a = b.fwd;
  const long long ray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (ray >= a.num_rays) return;
  const int S = a.num_samples;
  const int mode = a.geometry_mode;
  const bool drop_last = (mode == NGM_GEOM_DENSITY || mode == NGM_GEOM_NEUS);
  const int Se = drop_last ? S - 1 : S;
  const float gamma = a.geometry_factor;
  const BwdRay r{a, ray, S, (mode == NGM_GEOM_OCCUPANCY || mode == NGM_GEOM_DENSITY) ? -100.0f : 1.0f};
  float* w_occ = b.workspace + ray * S * 2;
  float* w_T = w_occ + S;
  const float isd = mode == NGM_GEOM_NEUS ? __ldg(a.neus_isd + ray / a.rays_per_isd) : 0.0f;
  const float ag = isd * gamma;  // neus: Phi = sigmoid(ag * g)

  // ---- forward sweep (run_mapping.py:746-779) ----
  float T = 1.0f, P = 0.f, C[3] = {0.f, 0.f, 0.f}, D = 0.f;
  for (int k = 0; k < Se; ++k) {
    const float g = r.geometry(k);
    float occ;
    if (mode == NGM_GEOM_NRGBD) {
      const float s = sigmoidf_(gamma * g);
      occ = 4.0f * s * (1.0f - s);
    } else if (mode == NGM_GEOM_OCCUPANCY) {
      occ = sigmoidf_(gamma * g);
    } else if (mode == NGM_GEOM_DENSITY) {
      occ = 1.0f - expf(-(r.dist(k + 1) - r.dist(k)) * fmaxf(g, 0.0f));
    } else {
      const float t0 = sigmoidf_(ag * g), t1 = sigmoidf_(ag * r.geometry(k + 1));
      occ = fmaxf((t0 - t1) / (t0 + 1e-5f), 0.0f);
    }
    w_occ[k] = occ;
    w_T[k] = T;
    const float w = occ * T;
    P += w;
    D = fmaf(w, r.depth(k), D);
#pragma unroll
    for (int c = 0; c < 3; ++c) C[c] = fmaf(w, r.color(k, c), C[c]);
    T *= 1.0f - occ;
  }

  // ---- upstream gradients ----
  float gC[3] = {0.f, 0.f, 0.f}, gVC[3] = {0.f, 0.f, 0.f}, gD = 0.f, gVD = 0.f, gP = 0.f;
  if (b.g_rgbd) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(b.g_rgbd) + ray);
    gC[0] = v.x; gC[1] = v.y; gC[2] = v.z; gD = v.w;
  }
  if (b.g_color_var) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gVC[c] = __ldg(b.g_color_var + ray * 3 + c);
  }
  if (b.g_depth_var) gVD = __ldg(b.g_depth_var + ray);
  if (b.g_term_prob) gP = __ldg(b.g_term_prob + ray);
  float kC[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) kC[c] = gVC[c] * 2.0f * C[c] * (P - 1.0f);
  const float kD = gVD * 2.0f * D * (P - 1.0f);
  const float tau = a.truncation;
  auto aux_grad = [&](int k) {  // d loss / d geometry through the free-space / TSDF sample terms (:624-639)
    float v = 0.0f;
    if (b.g_freespace) v += __ldg(b.g_freespace + ray * S + k);
    if (b.g_tsdf) v += __ldg(b.g_tsdf + ray * S + k);
    return v * tau;
  };
  auto store_geom = [&](int k, float dg) {
    b.d_geometries[(ray * S + k) * a.geometry_stride] = r.overwritten(k) ? 0.0f : dg + aux_grad(k);
  };
  auto store_color = [&](int k, float d0, float d1, float d2) {
    float* o = b.d_colors + (ray * S + k) * a.color_stride;
    o[0] = a.color_factor * d0; o[1] = a.color_factor * d1; o[2] = a.color_factor * d2;
  };

  // ---- backward sweep ----
  float R = 0.0f;
  float d_isd = 0.0f;
  float phi_next_self = 0.0f;  // neus: d loss / d Phi_{k+1} through occ_{k+1}
  if (drop_last) {             // the dropped last sample only feeds the aux terms (and Phi_{S-1} in neus mode)
    store_color(S - 1, 0.f, 0.f, 0.f);
    if (mode == NGM_GEOM_DENSITY || Se == 0) store_geom(S - 1, 0.0f);
  }
  for (int k = Se - 1; k >= 0; --k) {
    const float occ = w_occ[k], Tk = w_T[k];
    const float w = occ * Tk;
    const float z = r.depth(k);
    const float c0 = r.color(k, 0), c1 = r.color(k, 1), c2 = r.color(k, 2);
    const float e0 = C[0] - c0, e1 = C[1] - c1, e2 = C[2] - c2, ez = D - z;
    const float gw = gP + gC[0] * c0 + gC[1] * c1 + gC[2] * c2 + gD * z
                     + gVC[0] * e0 * e0 + gVC[1] * e1 * e1 + gVC[2] * e2 * e2 + kC[0] * c0 + kC[1] * c1 + kC[2] * c2
                     + gVD * ez * ez + kD * z;
    const float docc = Tk * (gw - R);
    R = fmaf(1.0f - occ, R, gw * occ);
    store_color(k, w * (gC[0] - 2.0f * gVC[0] * e0 + kC[0]), w * (gC[1] - 2.0f * gVC[1] * e1 + kC[1]),
                w * (gC[2] - 2.0f * gVC[2] * e2 + kC[2]));
    const float g = r.geometry(k);
    if (mode == NGM_GEOM_NRGBD) {
      const float s = sigmoidf_(gamma * g);
      store_geom(k, docc * gamma * occ * (1.0f - 2.0f * s));
    } else if (mode == NGM_GEOM_OCCUPANCY) {
      store_geom(k, docc * gamma * occ * (1.0f - occ));
    } else if (mode == NGM_GEOM_DENSITY) {
      const float delta = r.dist(k + 1) - r.dist(k);
      store_geom(k, g > 0.0f ? docc * delta * expf(-delta * g) : 0.0f);
    } else {  // neus: occ_k = max((Phi_k - Phi_{k+1}) / (Phi_k + eps), 0)
      const float gn = r.geometry(k + 1);
      const float t0 = sigmoidf_(ag * g), t1 = sigmoidf_(ag * gn);
      const float den = t0 + 1e-5f;
      const bool active = (t0 - t1) / den >= 0.0f;  // torch.clamp_min passes the gradient where input >= min
      const float d_t1 = phi_next_self + (active ? -docc / den : 0.0f);  // all of d loss / d Phi_{k+1}
      const float dsig1 = t1 * (1.0f - t1);
      store_geom(k + 1, d_t1 * ag * dsig1);
      d_isd = fmaf(d_t1 * dsig1, gamma * gn, d_isd);  // (an overwritten sample's fill value still multiplies isd)
      phi_next_self = active ? docc * (t1 + 1e-5f) / (den * den) : 0.0f;
      if (k == 0) {
        const float dsig0 = t0 * (1.0f - t0);
        store_geom(0, phi_next_self * ag * dsig0);
        d_isd = fmaf(phi_next_self * dsig0, gamma * g, d_isd);
      }
    }
  }
  if (b.d_neus_isd) b.d_neus_isd[ray] = d_isd;
}

}  // namespace

int launch_composite_bwd(const NgmCompositeBwdArgs& b, cudaStream_t stream) {
  const long long n = b.fwd.num_rays;
  if (n == 0) return NGM_OK;
  composite_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(b);
  return check_launch("composite_bwd_kernel");
}

}  // namespace ngm
