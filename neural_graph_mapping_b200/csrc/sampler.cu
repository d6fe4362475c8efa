// Stage kernel: ray sampler.
// Replaces Camera.sample_ijs_uniform (ngm/camera.py:215-292), the depth-guided merge
// (ngm/run_mapping.py:521-545) and utils.transform_points (ngm/utils.py:276-286).
//
// HBM-bound: algorithmic bytes per ray = 16 (ijs) + 8 (near, far) + St*(12 + 4 + 4)  [SURVEY 8d].
// One thread per (ray, sample); a warp covers 32 consecutive samples of (usually) one ray, so
// the per-ray inputs are warp-broadcast loads and every output row is written with consecutive
// addresses.  The sort of the reference is replaced by a rank computation: both sample sets are
// already sorted (disjoint strata), so the merged position of an element is its own index plus
// the number of elements of the other set below it, found by inspecting <= 3 candidate strata.
#include "common.cuh"

namespace ngm {

__global__ void __launch_bounds__(256) sample_rays_kernel(NgmSampleArgs a) {
  const int S = a.num_samples;
  const int G = a.gt ? a.num_samples_guided : 0;
  const int St = S + G;
  const long long total = a.num_rays * (long long)St;
  const RayJitter jit{a.jitter, a.jitter_guided, a.seed, a.offset};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long ray = t / St;
    const int k = (int)(t - ray * St);
    const float nr = a.near ? __ldg(a.near + ray) : a.near_scalar;
    const float fr = a.far ? __ldg(a.far + ray) : a.far_scalar;
    float glo = nr, ghi = fr;
    if (G > 0) guided_window(nr, fr, __ldg(a.gt + ray), a.range_guided, glo, ghi);
    float d;
    int pos;
    if (k < S) {
      d = stratified_distance(nr, fr, k, S, jit.coarse(ray, k, S, St));
      pos = k;
      if (G > 0)
        pos += count_before(d, glo, ghi, G, /*strict=*/true,
                            [&](int j) { return jit.guided(ray, j, S, G, St); });
    } else {
      const int kg = k - S;
      d = stratified_distance(glo, ghi, kg, G, jit.guided(ray, kg, S, G, St));
      pos = kg + count_before(d, nr, fr, S, /*strict=*/false,
                              [&](int j) { return jit.coarse(ray, j, S, St); });
    }
    const longlong2 ij = __ldg(reinterpret_cast<const longlong2*>(a.ijs) + ray);
    const float3 dir = ij_to_direction(ij.x, ij.y, a.cam);
    const float3 pc = make_float3(dir.x * d, dir.y * d, dir.z * d);
    const long long o = ray * St + pos;
    if (a.distances) a.distances[o] = d;
    if (a.depths) a.depths[o] = -pc.z;
    if (a.points_cam) {
      a.points_cam[o * 3 + 0] = pc.x;
      a.points_cam[o * 3 + 1] = pc.y;
      a.points_cam[o * 3 + 2] = pc.z;
    }
    if (a.points_world) {
      const float* m = a.c2ws + (a.c2w_per_ray ? ray * 16 : 0);
      float mm[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) mm[i] = __ldg(m + i);
      const float3 pw = transform_point(mm, pc);
      a.points_world[o * 3 + 0] = pw.x;
      a.points_world[o * 3 + 1] = pw.y;
      a.points_world[o * 3 + 2] = pw.z;
    }
  }
}

int launch_sample_rays(const NgmSampleArgs& a, cudaStream_t stream) {
  const int St = a.num_samples + (a.gt ? a.num_samples_guided : 0);
  const long long total = a.num_rays * (long long)St;
  if (total == 0) return NGM_OK;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 32;  // grid-stride beyond ~4 waves
  if (blocks > cap) blocks = cap;
  sample_rays_kernel<<<(unsigned)blocks, threads, 0, stream>>>(a);
  return check_launch("sample_rays_kernel");
}

}  // namespace ngm
