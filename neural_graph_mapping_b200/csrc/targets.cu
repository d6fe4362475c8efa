// Multi-view target sampling: the step right before the render in every mapping iteration
// (NeuralGraphMap._sample_target_mv, ngm/run_mapping.py:1261-1459).
//
// The reference expresses it as ~60 small torch kernels over (fields x 20 probe points x keyframes)
// intermediates.  Here it is two launches around the one data-dependent step (dropping fields no keyframe
// sees, then torch.multinomial over the visibility mask, :1367-1383):
//
//   target_visibility_kernel  one thread per (field, keyframe): the 20 probe points on the field's training
//     sphere go world -> camera (inverse pose, utils.py:279-282) -> image (Camera.project_points, OpenGL
//     convention, pixel centre 0.5; camera.py:119-154, 173-177); the three `any` masks of :1356-1362
//     (in front of the camera / closer than the keyframe's depth at that pixel / inside the image), and the
//     probes' 2-D bounding box of :1386-1389.
//   target_rays_kernel        one thread per (field, ray): pixel inside the box of the drawn keyframe
//     (:1398-1408), pose, near / far from the field centre's distance along the ray (:1414-1421), the
//     keyframe's RGB-D at the pixel, depth -> distance (camera.py:319-340) and the four masks (:1429-1445).
//
// Both are latency-bound gathers over a few ten thousand threads; what they remove is launch count and
// the (fields x probes x keyframes) intermediates.
#include "common.cuh"

namespace ngm {

namespace {

struct Pose {
  float r[9], t[3];
};

__device__ __forceinline__ Pose load_pose(const float* __restrict__ c2w) {
  Pose p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 row = __ldg(reinterpret_cast<const float4*>(c2w) + i);
    p.r[3 * i] = row.x; p.r[3 * i + 1] = row.y; p.r[3 * i + 2] = row.z; p.t[i] = row.w;
  }
  return p;
}

// utils.transform_points(points, c2w, inv=True): R^T (p - t)   (utils.py:279-282)
__device__ __forceinline__ void world_to_cam(const Pose& p, float x, float y, float z, float& cx, float& cy, float& cz) {
  const float dx = x - p.t[0], dy = y - p.t[1], dz = z - p.t[2];
  cx = p.r[0] * dx + p.r[3] * dy + p.r[6] * dz;
  cy = p.r[1] * dx + p.r[4] * dy + p.r[7] * dz;
  cz = p.r[2] * dx + p.r[5] * dy + p.r[8] * dz;
}

__global__ void __launch_bounds__(128) target_visibility_kernel(NgmTargetVisArgs a) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)a.num_fields * a.num_frames) return;
  const int f = (int)(idx / a.num_frames);
  const long long k = idx % a.num_frames;
  const long long slot = a.field_ids ? a.field_ids[f] : f;
  const float px = __ldg(a.positions + slot * 3), py = __ldg(a.positions + slot * 3 + 1), pz = __ldg(a.positions + slot * 3 + 2);
  const Pose pose = load_pose(a.c2ws + k * 16);
  const long long store = a.frame_to_store ? a.frame_to_store[k] : k;
  const float* __restrict__ image = a.rgbds + store * (long long)a.cam.height * a.cam.width * 4;
  // Camera.get_projection_matrix("opengl", 0.5): [[fx, 0, -cx], [0, -fy, -cy], [0, 0, -1]]  (camera.py:173-177)
  const float cxp = a.cam.cx0 + 0.5f, cyp = a.cam.cy0 + 0.5f;
  bool any_front = false, any_closer = false, any_inside = false;
  float min_x = INFINITY, min_y = INFINITY, max_x = -INFINITY, max_y = -INFINITY;
  for (int s = 0; s < a.num_probes; ++s) {
    const float wx = px + __ldg(a.probe_offsets + s * 3) * a.train_radius;
    const float wy = py + __ldg(a.probe_offsets + s * 3 + 1) * a.train_radius;
    const float wz = pz + __ldg(a.probe_offsets + s * 3 + 2) * a.train_radius;
    float X, Y, Z;
    world_to_cam(pose, wx, wy, wz, X, Y, Z);
    const float depth = -Z;                                  // :1330
    const float hz = -Z;
    const float u = __fdiv_rn(a.cam.fx * X - cxp * Z, hz);   // :148-151
    const float v = __fdiv_rn(-a.cam.fy * Y - cyp * Z, hz);
    const int iu = (int)u, iv = (int)v;                      // .int(): truncation toward zero (:1337)
    const bool inside = iu >= 0 && iu < a.cam.width && iv >= 0 && iv < a.cam.height;  // :1339-1344
    const float kf_depth = inside ? __ldg(image + ((long long)iv * a.cam.width + iu) * 4 + 3) : 0.0f;  // :1345-1354
    any_front |= depth > 0.0f;
    any_closer |= depth < kf_depth;
    any_inside |= inside;
    min_x = fminf(min_x, u); max_x = fmaxf(max_x, u);
    min_y = fminf(min_y, v); max_y = fmaxf(max_y, v);
  }
  a.field_kf_mask[idx] = (any_front && any_closer && any_inside) ? 1 : 0;  // :1356-1362
  reinterpret_cast<float2*>(a.min_xys)[idx] = make_float2(fmaxf(min_x, 0.0f), fmaxf(min_y, 0.0f));  // :1386
  reinterpret_cast<float2*>(a.max_xys)[idx] =
      make_float2(fminf(max_x, (float)a.cam.width), fminf(max_y, (float)a.cam.height));                // :1387-1389
}

__global__ void __launch_bounds__(128) target_rays_kernel(NgmTargetRaysArgs a) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)a.num_fields * a.rays_per_field) return;
  const int f = (int)(idx / a.rays_per_field);
  const long long k = a.frame_cids[idx];
  const long long slot = a.field_ids ? a.field_ids[f] : f;
  const float2 lo = __ldg(reinterpret_cast<const float2*>(a.min_xys) + (long long)f * a.num_frames + k);
  const float2 hi = __ldg(reinterpret_cast<const float2*>(a.max_xys) + (long long)f * a.num_frames + k);
  const float2 uv = __ldg(reinterpret_cast<const float2*>(a.uv) + idx);
  // :1398-1408  (x = column, y = row)
  const float x = __fadd_rn(__fmul_rn(hi.x - lo.x, uv.x), lo.x), y = __fadd_rn(__fmul_rn(hi.y - lo.y, uv.y), lo.y);
  // clamp_max as the reference; the lower clamp never acts for a keyframe the mask allows (its box contains a
  // probe inside the image) and only keeps a caller-supplied frame id outside the mask from indexing out of bounds
  const int j = max(min((int)x, a.cam.width - 1), 0), i = max(min((int)y, a.cam.height - 1), 0);
  reinterpret_cast<longlong2*>(a.ijs)[idx] = make_longlong2(i, j);
  // pose of the drawn keyframe (:1411)
  const float4* src = reinterpret_cast<const float4*>(a.c2ws + k * 16);
  float4* dst = reinterpret_cast<float4*>(a.out_c2ws + idx * 16);
  const float4 r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2), r3 = __ldg(src + 3);
  dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
  // field centre in the camera frame, distance along the ray (:1414-1421)
  const float dx = __ldg(a.positions + slot * 3) - r0.w, dy = __ldg(a.positions + slot * 3 + 1) - r1.w,
              dz = __ldg(a.positions + slot * 3 + 2) - r2.w;
  const float cx = r0.x * dx + r1.x * dy + r2.x * dz;
  const float cy = r0.y * dx + r1.y * dy + r2.y * dz;
  const float cz = r0.z * dx + r1.z * dy + r2.z * dz;
  const float ux = __fdiv_rn((float)j - a.cam.cx0, a.cam.fx), uy = __fdiv_rn((float)i - a.cam.cy0, a.cam.fy);  // camera.py:188-190
  const float norm = __fsqrt_rn(ux * ux + uy * uy + 1.0f);
  const float center = __fdiv_rn(cx * ux - cy * uy - cz, norm);  // OpenGL direction (ux, -uy, -1) / norm
  const float near = fmaxf(center - a.train_radius, 0.0f), far = fmaxf(center + a.train_radius, 0.0f);
  a.near[idx] = near;
  a.far[idx] = far;
  // keyframe RGB-D at the pixel (:1424-1428), depth -> distance along the ray (camera.py:339-340)
  const long long store = a.frame_to_store ? a.frame_to_store[k] : k;
  const float4 rgbd = __ldg(reinterpret_cast<const float4*>(a.rgbds) + (store * a.cam.height + i) * a.cam.width + j);
  reinterpret_cast<float4*>(a.out_rgbds)[idx] = rgbd;
  const float gt = __fdiv_rn(rgbd.w, __fdiv_rn(1.0f, norm));
  a.gt[idx] = gt;
  const bool valid_depth = gt != 0.0f;                          // :1431
  a.depth_mask[idx] = gt > near && gt < far && valid_depth;     // :1432-1436
  a.rgb_mask[idx] = rgbd.x != 0.0f || rgbd.y != 0.0f;           // :1438
  a.term_probs[idx] = gt < far ? 1.0f : 0.0f;                   // :1443
  a.term_mask[idx] = gt > near && valid_depth;                  // :1445
}

// NeuralGraphMap._get_observed_fields (ngm/run_mapping.py:1643-1670): which fields does the current depth image
// observe?  A few hundred valid pixels are back-projected (Camera.depth_to_pointcloud, OpenGL; camera.py:372-385);
// fields whose sphere's AABB misses the points' AABB are dropped (geometry.py:25-42); a remaining field is observed
// if the segment camera -> point passes within field_radius of its centre for any point (geometry.py:67-103).
// Every CTA back-projects the points into shared memory and reduces their AABB; each of its warps then owns one
// field, lanes striding over the points.
constexpr int kObservedWarps = 8;
constexpr int kObservedMaxPoints = 2048;

__global__ void __launch_bounds__(kObservedWarps * 32) observed_fields_kernel(NgmObservedArgs a) {
  __shared__ float pts[kObservedMaxPoints][3];
  __shared__ float red[kObservedWarps][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int p = threadIdx.x; p < a.num_points; p += blockDim.x) {
    const long long pix = a.pixel_ids[p];
    const float i = (float)(pix / a.cam.width), j = (float)(pix % a.cam.width);
    const float d = __ldg(a.depth + pix * a.pixel_stride);
    const float x = __fdiv_rn((j - a.cam.cx0) * d, a.cam.fx);        // camera.py:380
    const float y = __fdiv_rn(-(i - a.cam.cy0) * d, a.cam.fy);       // :381
    const float z = -d;                                              // :382
    pts[p][0] = x; pts[p][1] = y; pts[p][2] = z;
    mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
    mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { red[warp][c] = mn[c]; red[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mn[c] = red[0][c]; mx[c] = red[0][3 + c];
    for (int w = 1; w < kObservedWarps; ++w) { mn[c] = fminf(mn[c], red[w][c]); mx[c] = fmaxf(mx[c], red[w][3 + c]); }
  }
  const int f = blockIdx.x * kObservedWarps + warp;
  if (f >= a.num_fields) return;
  const Pose pose = load_pose(a.c2w);
  float c[3];
  world_to_cam(pose, __ldg(a.positions + f * 3), __ldg(a.positions + f * 3 + 1), __ldg(a.positions + f * 3 + 2), c[0], c[1], c[2]);
  const float r = a.field_radius;
  bool in_box = true;  // sphere AABB [c - r, c + r] against the points' AABB (geometry.py:38-42)
#pragma unroll
  for (int k = 0; k < 3; ++k) in_box = in_box && (c[k] - r <= mx[k]) && (c[k] + r >= mn[k]);
  bool hit = false;
  if (in_box) {
    const float r2 = r * r;
    for (int p = lane; p < a.num_points; p += 32) {
      const float x = pts[p][0], y = pts[p][1], z = pts[p][2];
      float sq = x * x + y * y + z * z;
      if (sq == 0.0f) sq = 1.0f;                                            // geometry.py:100
      const float t = fminf(fmaxf(__fdiv_rn(c[0] * x + c[1] * y + c[2] * z, sq), 0.0f), 1.0f);  // :101-102
      const float dx = c[0] - x * t, dy = c[1] - y * t, dz = c[2] - z * t;
      hit = hit || (dx * dx + dy * dy + dz * dz <= r2);                     // :79-83
    }
  }
  hit = __any_sync(0xffffffffu, hit);
  if (lane == 0) a.observed[f] = hit ? 1 : 0;
}

}  // namespace

int launch_observed_fields(const NgmObservedArgs& a, cudaStream_t stream) {
  if (a.num_fields == 0) return NGM_OK;
  if (a.num_points > kObservedMaxPoints) {
    set_error("num_points %d exceeds %d", a.num_points, kObservedMaxPoints);
    return NGM_ERR_UNSUPPORTED;
  }
  observed_fields_kernel<<<(a.num_fields + kObservedWarps - 1) / kObservedWarps, kObservedWarps * 32, 0, stream>>>(a);
  return check_launch("observed_fields_kernel");
}

int launch_target_visibility(const NgmTargetVisArgs& a, cudaStream_t stream) {
  const long long n = (long long)a.num_fields * a.num_frames;
  if (n == 0) return NGM_OK;
  target_visibility_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(a);
  return check_launch("target_visibility_kernel");
}

int launch_target_rays(const NgmTargetRaysArgs& a, cudaStream_t stream) {
  const long long n = (long long)a.num_fields * a.rays_per_field;
  if (n == 0) return NGM_OK;
  target_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(a);
  return check_launch("target_rays_kernel");
}

}  // namespace ngm
