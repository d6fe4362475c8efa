// kNN branch of the field set: NeuralFieldSet.forward(use_vmap=False), ngm/models.py:347-405.
//
// The reference finds the K nearest field centres per point (pytorch3d knn_points), masks points
// whose nearest centre is outside the radius, evaluates each unique field on its masked subset
// in a Python loop (with host syncs) and blends with softmax(-distance_factor * d).  Here:
//   1. knn_assign   : brute-force K nearest of F centres per point (centres staged in shared
//                     memory), radius mask, softmax weights, per-field entry counts
//   2. knn_scan     : exclusive scans -> entry / tile offsets per field (one block)
//   3. knn_scatter  : bucket the (point, neighbour) entries by field
//   4. field kernel : "gather mode" -- every tile is 128 entries of ONE field (field_simt.cu)
//   5. knn_blend    : out = sum_k w_k f_k(x_k), or outside_value outside the radius
// No host synchronisation anywhere: the data-dependent tile count stays on the device.
#include "common.cuh"

namespace ngm {

int launch_field_fwd_simt_gather(const NgmFieldFwdArgs& a, const int* entries, const int* entry_offsets,
                                 const int* tile_offsets, int knn_k, long long max_tiles, cudaStream_t stream);
int launch_field_fwd_tc_gather(const NgmFieldFwdArgs& a, const int* entries, const int* entry_offsets,
                               const int* tile_offsets, int knn_k, long long max_tiles, const void* rows_half,
                               cudaStream_t stream);
int launch_permuto_rows_f32(const PermutoRowsArgs& a, cudaStream_t stream);
int launch_permuto_rows_half(const PermutoRowsArgs& a, cudaStream_t stream);
size_t field_tc_workspace_bytes(const NgmFieldDesc& fd, int num_fields);
bool tc_rows_required(const NgmFieldDesc& fd);

namespace {

constexpr int kMaxK = 8;
constexpr int kCentreChunk = 1024;

struct KnnWs {
  int* pair_field;  // [N*K]  field index of neighbour k, -1 if the point is outside every radius
  float* pair_w;    // [N*K]  softmax blend weight
  int* counts;      // [F]
  int* entry_offsets;  // [F+1]
  int* tile_offsets;   // [F+1]
  int* cursors;        // [F]
  int* entries;        // [N*K]
  float* pair_out;     // [N*K*4]
  void* tc;            // fp16 path: per-field weight images
  void* rows;          // permutohedral encoding: pre-encoded layer-0 rows per entry (fp32: E floats, fp16: EP halves)
  size_t total;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

KnnWs carve(void* base, long long N, int K, int F, size_t tc_bytes, size_t row_bytes) {
  KnnWs w{};
  char* b = static_cast<char*>(base);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align256(o + bytes); return b ? b + at : nullptr; };
  w.pair_field = reinterpret_cast<int*>(take((size_t)N * K * 4));
  w.pair_w = reinterpret_cast<float*>(take((size_t)N * K * 4));
  w.counts = reinterpret_cast<int*>(take((size_t)(F + 1) * 4));
  w.entry_offsets = reinterpret_cast<int*>(take((size_t)(F + 1) * 4));
  w.tile_offsets = reinterpret_cast<int*>(take((size_t)(F + 1) * 4));
  w.cursors = reinterpret_cast<int*>(take((size_t)(F + 1) * 4));
  w.entries = reinterpret_cast<int*>(take((size_t)N * K * 4));
  w.pair_out = reinterpret_cast<float*>(take((size_t)N * K * 16));
  w.tc = take(tc_bytes);
  w.rows = take((size_t)N * K * row_bytes);
  w.total = o;
  return w;
}

// counters[f] += 1 for every active lane, one atomic per distinct field per warp (neighbouring points share their
// nearest fields, so a warp usually holds 1-3 distinct values: 40 M same-address atomics per frame become ~2 M).
// Returns the lane's slot.  Must be reached by all 32 lanes.
__device__ __forceinline__ int warp_aggregated_inc(int* counters, int f, bool active) {
  const unsigned live = __ballot_sync(0xffffffffu, active);
  int slot = -1;
  if (active) {
    const unsigned peers = __match_any_sync(live, f);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counters + f, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    slot = base + __popc(peers & ((1u << lane) - 1u));
  }
  return slot;
}

// KT = compile-time K (0: runtime K <= kMaxK): the sorted insertion is K predicated compare-swaps per centre; with
// the reference's K = 2 (neural_graph_map.yaml:21) the runtime-K form spent 4x the instructions on dead slots
template <int KT>
__global__ void __launch_bounds__(256) knn_assign_kernel(const float* __restrict__ points, long long N,
                                                         const float* __restrict__ centres, int F, int K_rt, float radius,
                                                         float distance_factor, int* __restrict__ pair_field,
                                                         float* __restrict__ pair_w, int* __restrict__ counts) {
  __shared__ float sc[kCentreChunk * 3];
  constexpr int KM = KT > 0 ? KT : kMaxK;
  const int K = KT > 0 ? KT : K_rt;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool live = i < N;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) { px = __ldg(points + i * 3); py = __ldg(points + i * 3 + 1); pz = __ldg(points + i * 3 + 2); }
  float bd[KM];
  int bi[KM];
#pragma unroll
  for (int j = 0; j < KM; ++j) { bd[j] = INFINITY; bi[j] = -1; }
  for (int c0 = 0; c0 < F; c0 += kCentreChunk) {
    const int cn = min(kCentreChunk, F - c0);
    __syncthreads();
    for (int t = threadIdx.x; t < cn * 3; t += blockDim.x) sc[t] = __ldg(centres + (size_t)c0 * 3 + t);
    __syncthreads();
    if (live) {
      for (int c = 0; c < cn; ++c) {
        const float dx = px - sc[c * 3], dy = py - sc[c * 3 + 1], dz = pz - sc[c * 3 + 2];
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        // sorted insertion into the K best (ascending); strict < keeps the lower index on ties
        float cd = d2;
        int ci = c0 + c;
#pragma unroll
        for (int j = 0; j < KM; ++j) {
          if (j < K && cd < bd[j]) {
            const float td = bd[j]; const int ti = bi[j];
            bd[j] = cd; bi[j] = ci;
            cd = td; ci = ti;
          }
        }
      }
    }
  }
  const float d0 = sqrtf(bd[0]);
  const bool inside = live && d0 < radius;  // models.py:369: only the nearest centre is tested
  // softmax(-distance_factor * d) over the K neighbours (models.py:384)
  float logit[KM], m = -INFINITY;
#pragma unroll
  for (int j = 0; j < KM; ++j)
    if (j < K) { logit[j] = -distance_factor * sqrtf(bd[j]); m = fmaxf(m, logit[j]); }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < KM; ++j)
    if (j < K) { logit[j] = expf(logit[j] - m); sum += logit[j]; }
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    if (j < K) {  // K is warp-uniform: every lane reaches the warp collective
      if (live) {
        pair_field[i * K + j] = inside ? bi[j] : -1;
        pair_w[i * K + j] = logit[j] / sum;
      }
      warp_aggregated_inc(counts, bi[j], inside);
    }
  }
}

__global__ void __launch_bounds__(1024) knn_scan_kernel(const int* __restrict__ counts, int F, int* __restrict__ entry_offsets,
                                                        int* __restrict__ tile_offsets, int* __restrict__ cursors) {
  // single block; F is small (hundreds to a few thousand fields): chunked Hillis-Steele in shared memory
  __shared__ int s_e[1024], s_t[1024];
  __shared__ int carry_e, carry_t;
  if (threadIdx.x == 0) { carry_e = 0; carry_t = 0; }
  __syncthreads();
  for (int base = 0; base < F; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = i < F ? counts[i] : 0;
    const int tl = (c + 127) / 128;
    s_e[threadIdx.x] = c;
    s_t[threadIdx.x] = tl;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int ve = threadIdx.x >= o ? s_e[threadIdx.x - o] : 0;
      const int vt = threadIdx.x >= o ? s_t[threadIdx.x - o] : 0;
      __syncthreads();
      s_e[threadIdx.x] += ve;
      s_t[threadIdx.x] += vt;
      __syncthreads();
    }
    if (i < F) {
      entry_offsets[i] = carry_e + s_e[threadIdx.x] - c;
      tile_offsets[i] = carry_t + s_t[threadIdx.x] - tl;
      cursors[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_e += s_e[1023]; carry_t += s_t[1023]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { entry_offsets[F] = carry_e; tile_offsets[F] = carry_t; }
}

__global__ void __launch_bounds__(256) knn_scatter_kernel(const int* __restrict__ pair_field, long long NK,
                                                          const int* __restrict__ entry_offsets, int* __restrict__ cursors,
                                                          int* __restrict__ entries) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int f = e < NK ? pair_field[e] : -1;
  const int pos = warp_aggregated_inc(cursors, f, f >= 0);
  if (f >= 0) entries[entry_offsets[f] + pos] = (int)e;
}

__global__ void __launch_bounds__(256) knn_blend_kernel(const int* __restrict__ pair_field, const float* __restrict__ pair_w,
                                                        const float* __restrict__ pair_out, long long N, int K,
                                                        float outside_value, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 acc = make_float4(outside_value, outside_value, outside_value, outside_value);  // models.py:401
  if (pair_field[i * K] >= 0) {
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {  // einsum "... k, ... k dc -> ... dc" (models.py:399)
      const float w = pair_w[i * K + k];
      const float4 o = reinterpret_cast<const float4*>(pair_out)[i * K + k];
      acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
    }
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

}  // namespace

static size_t tc_bytes_of(const NgmKnnFwdArgs& a) {
  return a.precision == NGM_PREC_FP16 ? field_tc_workspace_bytes(a.field, a.num_fields) : 0;
}

// bytes of one pre-encoded row (0: the field kernel encodes in-line)
static size_t row_bytes_of(const NgmKnnFwdArgs& a) {
  const size_t half_row = (size_t)((a.field.dim_encoding + 15) / 16 * 16) * 2;
  if (a.precision == NGM_PREC_FP16 && tc_rows_required(a.field)) return half_row;  // Fourier / Triplane: always rows
  if (a.field.encoding != NGM_ENC_PERMUTO || a.field.permuto_feats != 2) return 0;
  if (a.precision == NGM_PREC_FP16) return half_row;
  // fp32: measured no gain on this path (the narrow FFMA kernel already runs three CTAs per SM and hides the
  // gathers: 39.8 ms in-kernel against 42.0 ms with fp32 rows on the 640x480x64 eval frame), so it encodes in-line
  return 0;
}

size_t knn_workspace_bytes(const NgmKnnFwdArgs& a) {
  const int K = a.num_knn < a.num_fields ? a.num_knn : a.num_fields;
  return carve(nullptr, a.num_points, K > 0 ? K : 1, a.num_fields, tc_bytes_of(a), row_bytes_of(a)).total;
}

int launch_fieldset_knn(const NgmKnnFwdArgs& a, cudaStream_t stream) {
  const long long N = a.num_points;
  const int F = a.num_fields;
  const int K = a.num_knn < F ? a.num_knn : F;  // models.py:355-358
  if (N == 0) return NGM_OK;
  const size_t row_bytes = row_bytes_of(a);
  const KnnWs w = carve(a.workspace, N, K, F, tc_bytes_of(a), row_bytes);
  NGM_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)(F + 1) * sizeof(int), stream));
  const unsigned pb = (unsigned)((N + 255) / 256);
  auto assign = K == 1 ? knn_assign_kernel<1> : K == 2 ? knn_assign_kernel<2> : K == 3 ? knn_assign_kernel<3>
                                                                          : K == 4 ? knn_assign_kernel<4> : knn_assign_kernel<0>;
  assign<<<pb, 256, 0, stream>>>(a.points, N, a.positions, F, K, a.field_radius, a.distance_factor,
                                            w.pair_field, w.pair_w, w.counts);
  if (int rc = check_launch("knn_assign_kernel")) return rc;
  knn_scan_kernel<<<1, 1024, 0, stream>>>(w.counts, F, w.entry_offsets, w.tile_offsets, w.cursors);
  if (int rc = check_launch("knn_scan_kernel")) return rc;
  const long long NK = N * K;
  knn_scatter_kernel<<<(unsigned)((NK + 255) / 256), 256, 0, stream>>>(w.pair_field, NK, w.entry_offsets, w.cursors,
                                                                       w.entries);
  if (int rc = check_launch("knn_scatter_kernel")) return rc;

  NgmFieldFwdArgs f{};
  f.field = a.field;
  f.points = a.points;
  f.positions = a.positions;
  f.orientations = a.orientations;
  f.field_slots = a.field_slots;
  f.out = w.pair_out;
  f.field_radius = a.scale_radius;
  f.num_fields = F;
  f.scale_mode = a.scale_mode;
  f.precision = a.precision;
  f.workspace = w.tc;
  if (row_bytes) {  // permutohedral encoding: one whole-GPU pass encodes every (point, neighbour) entry
    PermutoRowsArgs e{};
    e.field = a.field;
    e.points_world = a.points;
    e.positions = a.positions; e.orientations = a.orientations;
    e.field_slots = reinterpret_cast<const long long*>(a.field_slots);
    e.out = static_cast<uint32_t*>(w.rows);
    e.num_points = NK;
    e.points_per_field = 1;
    e.field_radius = a.scale_radius;
    e.scale_mode = a.scale_mode;
    e.EP = (a.field.dim_encoding + 15) / 16 * 16;
    e.pair_field = w.pair_field;
    e.knn_k = K;
    if (int rc = a.precision == NGM_PREC_FP16 ? launch_permuto_rows_half(e, stream) : launch_permuto_rows_f32(e, stream))
      return rc;
  }
  const long long max_tiles = (NK + 127) / 128 + F;  // upper bound of sum_f ceil(count_f / 128)
  if (a.precision == NGM_PREC_FP16) {  // tcgen05 field kernel in gather mode
    if (int rc = launch_field_fwd_tc_gather(f, w.entries, w.entry_offsets, w.tile_offsets, K, max_tiles,
                                            row_bytes ? w.rows : nullptr, stream))
      return rc;
  } else {
    f.workspace = row_bytes ? w.rows : nullptr;  // fp32: the gather-mode kernel reads these rows
    if (int rc = launch_field_fwd_simt_gather(f, w.entries, w.entry_offsets, w.tile_offsets, K, max_tiles, stream)) return rc;
  }

  knn_blend_kernel<<<pb, 256, 0, stream>>>(w.pair_field, w.pair_w, w.pair_out, N, K, a.outside_value, a.out);
  return check_launch("knn_blend_kernel");
}

}  // namespace ngm
