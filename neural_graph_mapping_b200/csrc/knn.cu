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
// Points handled by one block (kAssignIters x 256) / entries handled by one scatter block: the per-field counters are
// bumped in a shared-memory histogram first and reach global memory as ONE atomic per (block, field).  Neighbouring
// rays see the same two or three fields, so at any moment every warp of the GPU targets the same few counters and
// same-address L2 atomics retire about one per clock: 2.5 M warp-aggregated atomics per 640x480x64 frame were
// ~1.1 ms of the assign and of the scatter kernel each; per-block aggregation leaves ~40 k.
constexpr int kAssignIters = 8;
constexpr int kScatterIters = 8;
constexpr int kHistMaxFields = 6144;  // shared-memory histogram (<= 24 KB beside the 21 KB of candidate lists); more fields: global counters directly

// Candidate pruning (exact).  The 32 points of a warp are neighbouring samples of one ray, i.e. a short segment: with
// m their centroid and rho the largest |p - m|, every point's distance to a centre c lies in [|m-c| - rho, |m-c| + rho].
// If U is the K-th smallest upper bound, every point has K centres within U, so a centre whose lower bound exceeds U
// is not among any point's K nearest.  The warp collects the surviving centres (ascending index, so ties resolve as
// in the full scan) in shared memory and only those enter the exact per-point insertion: ~10 of 75 centres on the
// bench scene, and the cost no longer grows with the number of fields of a large map.  Bounds are inflated by more
// than their rounding error; a warp whose list overflows scans all centres instead.
constexpr int kCandCap = 96;
// The same bound one level up: the block's 256 points (a few neighbouring rays) first reduce the F centres to a block
// list, cooperatively (F / 256 centres per thread), and the warps refine from that list instead of each scanning all
// F centres -- the per-warp scans were 3.1 of 5.5 ms per frame on a 2,025-field map.
constexpr int kBlockCandCap = 512;

template <int KT>
__global__ void __launch_bounds__(256) knn_assign_kernel(const float* __restrict__ points, long long N,
                                                         const float* __restrict__ centres, int F, int K_rt, float radius,
                                                         float distance_factor, int* __restrict__ pair_field,
                                                         float* __restrict__ pair_w, int* __restrict__ counts, int use_hist) {
  __shared__ float4 cand[8][kCandCap];
  __shared__ float4 bcand[kBlockCandCap];
  __shared__ float red[8][4];
  __shared__ float kub[8][kMaxK];
  __shared__ int wcnt[8];
  extern __shared__ int hist[];  // [F] when use_hist
  constexpr int KM = KT > 0 ? KT : kMaxK;
  const int K = KT > 0 ? KT : K_rt;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool two_level = F > 4 * kCandCap;  // small maps: the warps scan the centres directly
  if (use_hist) {
    for (int t = threadIdx.x; t < F; t += blockDim.x) hist[t] = 0;
    __syncthreads();
  }
  int* const ctr = use_hist ? hist : counts;
  for (int it = 0; it < kAssignIters; ++it) {
    const long long base = (blockIdx.x * (long long)kAssignIters + it) * blockDim.x;
    if (base >= N) break;  // block-uniform
    const long long i = base + threadIdx.x;
    const bool live = i < N;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = __ldg(points + i * 3); py = __ldg(points + i * 3 + 1); pz = __ldg(points + i * 3 + 2); }
    const unsigned live_mask = __ballot_sync(0xffffffffu, live);
    float bd[KM];
    int bi[KM];
#pragma unroll
    for (int j = 0; j < KM; ++j) { bd[j] = INFINITY; bi[j] = -1; }
    // ---- block level: bounding sphere of the block's points, block candidate list (ascending index) ----
    int nsrc = F;             // centres the warps scan: all of them (global memory) ...
    bool from_list = false;   // ... or the block list (shared memory)
    if (two_level) {          // block-uniform
      float sx = live ? px : 0.f, sy = live ? py : 0.f, sz = live ? pz : 0.f, sn = live ? 1.f : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
      }
      __syncthreads();  // the previous iteration's readers of red / bcand are done
      if (lane == 0) { red[wid][0] = sx; red[wid][1] = sy; red[wid][2] = sz; red[wid][3] = sn; }
      __syncthreads();
      float tx = 0.f, ty = 0.f, tz = 0.f, tn = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { tx += red[w][0]; ty += red[w][1]; tz += red[w][2]; tn += red[w][3]; }
      const float inv = 1.0f / tn;  // base < N: at least one live point
      const float mx = tx * inv, my = ty * inv, mz = tz * inv;
      float rho = live ? sqrtf((px - mx) * (px - mx) + (py - my) * (py - my) + (pz - mz) * (pz - mz)) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rho = fmaxf(rho, __shfl_xor_sync(0xffffffffu, rho, o));
      __syncthreads();
      if (lane == 0) red[wid][0] = rho;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 8; ++w) rho = fmaxf(rho, red[w][0]);
      const float scale = fabsf(mx) + fabsf(my) + fabsf(mz) + rho;
      rho = rho * 1.0001f + 1e-5f * scale + 1e-30f;
      // K smallest |m - c| of this thread's share, merged per warp, then over the 8 warps
      float ub[KM];
#pragma unroll
      for (int j = 0; j < KM; ++j) ub[j] = INFINITY;
      for (int c = threadIdx.x; c < F; c += 256) {
        const float dx = mx - __ldg(centres + c * 3), dy = my - __ldg(centres + c * 3 + 1), dz = mz - __ldg(centres + c * 3 + 2);
        float v = sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
        for (int j = 0; j < KM; ++j)
          if (j < K && v < ub[j]) { const float t = ub[j]; ub[j] = v; v = t; }
      }
      for (int r = 0; r < K; ++r) {  // the warp's K smallest, ascending, into kub[wid][]
        float v = ub[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        const unsigned who = __ballot_sync(0xffffffffu, ub[0] == v);
        if (lane == __ffs(who) - 1) {
          kub[wid][r] = v;
#pragma unroll
          for (int j = 0; j + 1 < KM; ++j) ub[j] = ub[j + 1];
          ub[KM - 1] = INFINITY;
        }
      }
      __syncthreads();
      float U = INFINITY;
      {  // K-th smallest of the 8 sorted lists (every thread, 8 K values)
        int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < K; ++r) {
          float best = INFINITY;
          int bw = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float v = head[w] < K ? kub[w][head[w]] : INFINITY;
            if (v < best) { best = v; bw = w; }
          }
          U = best;
#pragma unroll
          for (int w = 0; w < 8; ++w) head[w] += (w == bw);
        }
      }
      const float limit = (U + rho) * 1.0001f + rho + 1e-5f * scale;
      int total = 0;
      for (int c0 = 0; c0 < F; c0 += 256) {  // ordered compaction, 256 centres per round
        const int c = c0 + threadIdx.x;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        bool keep = false;
        if (c < F) {
          cx = __ldg(centres + c * 3); cy = __ldg(centres + c * 3 + 1); cz = __ldg(centres + c * 3 + 2);
          const float dx = mx - cx, dy = my - cy, dz = mz - cz;
          keep = sqrtf(dx * dx + dy * dy + dz * dz) <= limit;
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) wcnt[wid] = __popc(km);
        __syncthreads();
        int before = 0, round = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { before += w < wid ? wcnt[w] : 0; round += wcnt[w]; }
        const int pos = total + before + __popc(km & ((1u << lane) - 1u));
        if (keep && pos < kBlockCandCap) bcand[pos] = make_float4(cx, cy, cz, __int_as_float(c));
        total += round;
      }
      __syncthreads();
      if (total <= kBlockCandCap) { nsrc = total; from_list = true; }
    }
    if (live_mask != 0u) {  // warp-uniform
      // ---- bounding sphere of the warp's points ----
      float sx = live ? px : 0.f, sy = live ? py : 0.f, sz = live ? pz : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
      }
      const float inv = 1.0f / (float)__popc(live_mask);
      const float mx = sx * inv, my = sy * inv, mz = sz * inv;
      float rho = live ? sqrtf((px - mx) * (px - mx) + (py - my) * (py - my) + (pz - mz) * (pz - mz)) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rho = fmaxf(rho, __shfl_xor_sync(0xffffffffu, rho, o));
      const float scale = fabsf(mx) + fabsf(my) + fabsf(mz) + rho;
      rho = rho * 1.0001f + 1e-5f * scale + 1e-30f;  // far above the rounding error of the bounds below
      // ---- K-th smallest upper bound over all centres ----
      float ub[KM];
#pragma unroll
      for (int j = 0; j < KM; ++j) ub[j] = INFINITY;
      for (int c = lane; c < nsrc; c += 32) {
        float cx, cy, cz;
        if (from_list) { const float4 q = bcand[c]; cx = q.x; cy = q.y; cz = q.z; }
        else { cx = __ldg(centres + c * 3); cy = __ldg(centres + c * 3 + 1); cz = __ldg(centres + c * 3 + 2); }
        const float dx = mx - cx, dy = my - cy, dz = mz - cz;
        float v = sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
        for (int j = 0; j < KM; ++j)
          if (j < K && v < ub[j]) { const float t = ub[j]; ub[j] = v; v = t; }
      }
      float U = INFINITY;
      for (int r = 0; r < K; ++r) {  // pop the warp-wide minimum K times (K <= F)
        float v = ub[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        U = v;
        const unsigned who = __ballot_sync(0xffffffffu, ub[0] == v);
        if (lane == __ffs(who) - 1) {
#pragma unroll
          for (int j = 0; j + 1 < KM; ++j) ub[j] = ub[j + 1];
          ub[KM - 1] = INFINITY;
        }
      }
      const float limit = (U + rho) * 1.0001f + rho + 1e-5f * scale;  // keep c  <=>  |m - c| - rho <= U + rho (inflated)
      // ---- ordered compaction of the surviving centres ----
      int cnt = 0;
      for (int c0 = 0; c0 < nsrc; c0 += 32) {  // warp-uniform trip count
        const int c = c0 + lane;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        int ci = c;
        bool keep = false;
        if (c < nsrc) {
          if (from_list) { const float4 q = bcand[c]; cx = q.x; cy = q.y; cz = q.z; ci = __float_as_int(q.w); }
          else { cx = __ldg(centres + c * 3); cy = __ldg(centres + c * 3 + 1); cz = __ldg(centres + c * 3 + 2); }
          const float dx = mx - cx, dy = my - cy, dz = mz - cz;
          keep = sqrtf(dx * dx + dy * dy + dz * dz) <= limit;
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        const int pos = cnt + __popc(km & ((1u << lane) - 1u));
        if (keep && pos < kCandCap) cand[wid][pos] = make_float4(cx, cy, cz, __int_as_float(ci));
        cnt += __popc(km);
      }
      __syncwarp();
      if (cnt <= kCandCap) {
        if (live) {
          for (int q = 0; q < cnt; ++q) {
            const float4 cc = cand[wid][q];
            const float dx = px - cc.x, dy = py - cc.y, dz = pz - cc.z;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            // sorted insertion into the K best (ascending); strict < keeps the lower index on ties.  After the first
            // few candidates hardly any beats the running K-th best: the insertion (13 of the 21 instructions per
            // candidate) sits behind one compare that is false for the whole warp most of the time
            if (KT > 0 && !(d2 < bd[KM - 1])) continue;
            float cd = d2;
            int ci = __float_as_int(cc.w);
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
      } else if (live) {  // incoherent points (e.g. a random query set): the full scan of the block list / of all centres
        for (int c = 0; c < nsrc; ++c) {
          float cx, cy, cz;
          int ci = c;
          if (from_list) { const float4 q = bcand[c]; cx = q.x; cy = q.y; cz = q.z; ci = __float_as_int(q.w); }
          else { cx = __ldg(centres + c * 3); cy = __ldg(centres + c * 3 + 1); cz = __ldg(centres + c * 3 + 2); }
          const float dx = px - cx, dy = py - cy, dz = pz - cz;
          const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          float cd = d2;
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
      __syncwarp();  // the list is rewritten in the next iteration
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
        warp_aggregated_inc(ctr, bi[j], inside);
      }
    }
  }
  if (use_hist) {
    __syncthreads();
    for (int t = threadIdx.x; t < F; t += blockDim.x) {
      const int h = hist[t];
      if (h) atomicAdd(counts + t, h);
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
                                                          int* __restrict__ entries, int F, int use_hist) {
  extern __shared__ int hist[];  // [F] when use_hist: the block's entries per field, then their base in `entries`
  if (!use_hist) {  // one warp-aggregated global atomic per (warp, field)
    for (int it = 0; it < kScatterIters; ++it) {
      const long long e = (blockIdx.x * (long long)kScatterIters + it) * blockDim.x + threadIdx.x;
      const int f = e < NK ? pair_field[e] : -1;
      const int pos = warp_aggregated_inc(cursors, f, f >= 0);
      if (f >= 0) entries[entry_offsets[f] + pos] = (int)e;
    }
    return;
  }
  for (int t = threadIdx.x; t < F; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  int fs[kScatterIters], rank[kScatterIters];
#pragma unroll
  for (int it = 0; it < kScatterIters; ++it) {  // rank of every entry among the block's entries of its field
    const long long e = (blockIdx.x * (long long)kScatterIters + it) * blockDim.x + threadIdx.x;
    fs[it] = e < NK ? pair_field[e] : -1;
    rank[it] = warp_aggregated_inc(hist, fs[it], fs[it] >= 0);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < F; t += blockDim.x) {  // reserve the block's range of every field it touched
    const int h = hist[t];
    if (h) hist[t] = entry_offsets[t] + atomicAdd(cursors + t, h);
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kScatterIters; ++it) {
    const long long e = (blockIdx.x * (long long)kScatterIters + it) * blockDim.x + threadIdx.x;
    if (fs[it] >= 0) entries[hist[fs[it]] + rank[it]] = (int)e;
  }
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
  const int use_hist = F <= kHistMaxFields ? 1 : 0;
  const size_t hist_bytes = use_hist ? (size_t)F * sizeof(int) : 0;
  auto assign = K == 1 ? knn_assign_kernel<1> : K == 2 ? knn_assign_kernel<2> : K == 3 ? knn_assign_kernel<3>
                                                                          : K == 4 ? knn_assign_kernel<4> : knn_assign_kernel<0>;
  const long long per_assign = 256LL * kAssignIters;
  assign<<<(unsigned)((N + per_assign - 1) / per_assign), 256, hist_bytes, stream>>>(
      a.points, N, a.positions, F, K, a.field_radius, a.distance_factor, w.pair_field, w.pair_w, w.counts, use_hist);
  if (int rc = check_launch("knn_assign_kernel")) return rc;
  knn_scan_kernel<<<1, 1024, 0, stream>>>(w.counts, F, w.entry_offsets, w.tile_offsets, w.cursors);
  if (int rc = check_launch("knn_scan_kernel")) return rc;
  const long long NK = N * K;
  const long long per_scatter = 256LL * kScatterIters;
  knn_scatter_kernel<<<(unsigned)((NK + per_scatter - 1) / per_scatter), 256, hist_bytes, stream>>>(
      w.pair_field, NK, w.entry_offsets, w.cursors, w.entries, F, use_hist);
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
