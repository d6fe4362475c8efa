// tcgen05 backward of the per-field MLP (the training call: ngm/run_mapping.py:1164-1186 -> loss.backward() into
// vmap_fields_params through NeuralField.forward, ngm/models.py:143-182 under torch.vmap, :342), sm_100a only.
//
//   h_0 = enc(x);  z_{l+1} = W_l h_l + b_l,  h_{l+1} = relu(z_{l+1})  (l < L);  out = W_L h_L + b_L
//   given g_out = dLoss/d out (from the compositor backward):
//     dW_L = g_out^T h_L,  db_L = sum_p g_out
//     g_l  = (g_{l+1} W_l) .* [h_l > 0]           (g_{L+1} := g_out)
//     dW_l = g_{l+1}^T h_l,  db_l = sum_p g_{l+1}  (l < L)
//
// One persistent CTA per SM walks a contiguous range of 128-point tiles, ordered by field.  Per tile it RECOMPUTES
// the forward (activations are never stored to HBM by the forward pass), then runs the chain above.  All three GEMM
// families run on tcgen05 with fp16 operands and fp32 accumulation:
//   forward   z = h W^T     : A = h   (TMEM, rows = points),   B = W_l K-major  (the forward kernel's weight image)
//   chain     g' = g W      : A = g   (TMEM, rows = points),   B = W_l MN-major (the SAME image, transposed by the
//                                                                              descriptor -- no second copy)
//   weights   dW += g^T h   : A = g   (smem, MN-major),        B = h (smem, MN-major): the contraction runs over the
//                             tile's 128 points; the accumulator STAYS in TMEM across all tiles of the field that
//                             this CTA owns and is flushed once per field segment (fp32 atomics into HBM).
//   biases    db += g^T 1   : same A, B = a constant ones tile (N = 16).
// Activations / gradients that a weight-gradient GEMM needs live in shared memory as [point][feature] rows in the
// SWIZZLE_128B layout, which one buffer serves both as K-major (K = features) and MN-major (MN = features) operand;
// g_l is written IN PLACE over h_l once the GEMMs that read h_l have completed.
//
// The upstream gradient is scaled by a per-call power of two s (max |g_out| * s in (4, 8]) before the conversion
// to fp16 -- mean-reduced losses give |g_out| ~ 1e-7, far below the fp16 range -- and the accumulators are unscaled
// by 1/s when they are flushed.  The ReLU mask of a layer is re-derived from its stored activation (h > 0).
//
// Measured dead ends (tools/bwd_phases.py, config-4 shard): two accumulator loads in flight (spills at the 168-register
// cap of a 9-warp CTA), spinning row threads (they starve the issuer warp's MMA issue: +4 %), loading the next tile's
// point / upstream gradient one tile ahead (+5 %: the extra live registers cost more than the hidden latency).
//
// TMEM (512 columns): [0,128) chain/forward accumulator, [128,192) fp16 A operand, the rest weight / bias gradient
// accumulators.  Three 128x128 fp32 accumulators do not fit next to those, so a 4-layer x 128 field is processed in
// two launches (linears {4,3,2}, then {1,0}); each launch recomputes the forward and as much of the chain as it needs.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "field_tc_common.cuh"
#include "tc_ptx.cuh"

namespace ngm {

int launch_permuto_rows_half(const PermutoRowsArgs& a, cudaStream_t stream);
bool field_tc_supported(const NgmFieldDesc& fd, const char** why);
bool tc_rows_required(const NgmFieldDesc& fd);

namespace {

constexpr int kBwdThreads = 288;  // warp 0: MMA issuer + weight loader; warps 1-8: two threads per tile row
constexpr int kColWork = 0, kColA = 128, kColAcc = 192;
constexpr uint32_t kSmemLimit = 232448;  // 227 KB of dynamic shared memory per CTA

struct BwdPlan {
  int lo, hi;                         // weight gradients of linears lo..hi are accumulated by this launch
  int col_dw[NGM_MAX_LINEARS];        // TMEM column of dW_l (l in [lo, hi])
  int col_db[NGM_MAX_LINEARS];        // TMEM column of db_l (l in [lo, hi], l < L)
  uint32_t buf_off[NGM_MAX_LINEARS];  // byte offset (from the activation area) of the buffer of h_l / g_l, l in [lo, min(hi+1, L)]
  uint32_t gt_off, ones_off;          // g_out^T tile (16 x 128 K-major) and the ones tile, from the activation area
  uint32_t acts_bytes;                // activation area size (buffers + the two small tiles)
  int want_denc;                      // also emit dLoss/d h_0 (lo == 0 only)
  int emit_g;                         // lo >= 1: also produce g_lo and spill it (fp16 rows) for the next launch
  int g_in;                           // hi < L: the chain starts from the spilled g_{hi+1} (forward only up to h_hi)
};

struct BwdParams {
  TcImage im;
  const uint8_t* images;
  int images_by_slot;  // persistent images (NgmFieldDesc.packed_weights), indexed by table row
  int num_fields;
  int E, EP, W, WP, L, dim_out, nerf_start;
  const float* positions;
  const float* orientations;
  const long long* field_slots;
  int scale_mode;
  float field_radius;
  const float* points;  // (F, N, 3)
  long long points_per_field;
  const __half* raw_a;  // pre-encoded layer-0 rows (F * N, EP) or nullptr (NeRF front end in the kernel)
  const float* d_out;   // (F, N, dim_out)
  const unsigned* absmax_bits;  // max |d_out| as float bits
  float* d_w[NGM_MAX_LINEARS];  // (F, out, in) fp32, zeroed by the host side
  float* d_b[NGM_MAX_LINEARS];  // (F, out)
  float* d_enc;                 // (F, N, E) or nullptr
  uint32_t* g_spill;            // (total_tiles, 128, WP / 2) packed fp16 rows of the spilled gradient
  long long field_base;         // first field of this launch group (row of positions / orientations / field_slots)
  long long tiles_per_field, total_tiles;
  BwdPlan plan;
};

struct BwdSmem {
  uint64_t w_ready, a_ready, d_ready, dw_done;
  uint32_t tmem_base;
  uint32_t pad_;
};

__device__ __forceinline__ int in_dim(const BwdParams& p, int l) { return l == 0 ? p.EP : p.W; }

// Phase accounting (diagnostics build only): CTA 0's issuer lane and one row thread add the cycles they spend in each
// phase to a global table, read back with ngm_debug_bwd_phases.
//   row thread: 0 front end, 1 wait prev dW, 2 wait forward MMA, 3 forward epilogue, 4 wait chain MMA, 5 chain epilogue,
//               6 wait dW of this step, 7 store + arrive, 8 flush, 9 spill epilogue;   issuer: 10 wait for operands, 11 issue
#ifdef NGM_DEBUG_EXPORTS
__device__ unsigned long long g_bwd_phase[16];
#define BWD_PHASE_INIT const bool ph_on = blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32); long long ph_t = clock64();
#define BWD_PHASE(i)                                                                      \
  if (ph_on) {                                                                            \
    const long long now_ = clock64();                                                     \
    atomicAdd(&g_bwd_phase[i], (unsigned long long)(now_ - ph_t));                        \
    ph_t = now_;                                                                          \
  }
#else
#define BWD_PHASE_INIT
#define BWD_PHASE(i)
#endif

// K-major SWIZZLE_128B operand (weights forward, g_out^T, ones): k-step ks of 16 halves
__device__ __forceinline__ uint64_t kmajor_step(uint64_t desc0, int ks, uint32_t atom_stride16) {
  return desc0 + (uint64_t)((ks & 3) * 2u + (ks >> 2) * atom_stride16);
}

// ---- MMA groups (issued by one elected lane; every operand is warp-uniform) ----
// forward: D[work] = A[tmem] x W_l^T
__device__ __forceinline__ void issue_forward(uint32_t d_addr, uint32_t a_addr, uint32_t w_addr, const TcLayer& y) {
  const uint64_t desc0 = ptx::make_smem_desc_sw128(w_addr + y.off);
  const uint32_t idesc = ptx::make_idesc_f16(y.n_pad);
  const int ksteps = y.k_pad / 16;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    if (ks < ksteps) ptx::mma_f16_ts(d_addr, a_addr + ks * 8, kmajor_step(desc0, ks, (uint32_t)y.n_pad * 8u), idesc, ks > 0);
}
// chain: D[work] (points x in) = A[tmem] (points x out) x W_l (out x in), B read MN-major from the forward image
__device__ __forceinline__ void issue_chain(uint32_t d_addr, uint32_t a_addr, uint32_t w_addr, const TcLayer& y, int n) {
  const uint64_t desc0 = ptx::make_smem_desc_sw128_ex(w_addr + y.off, (uint32_t)y.n_pad * 128u, 1024u);
  const uint32_t idesc = ptx::make_idesc_f16_mn(128, n, 0, 1);
  const int ksteps = y.n_pad / 16;  // contraction over the layer's outputs
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    if (ks < ksteps) ptx::mma_f16_ts(d_addr, a_addr + ks * 8, desc0 + (uint64_t)(ks * 128), idesc, ks > 0);
}
// weight gradient: D[acc] (M x N) += A^T B over the tile's 128 points; A, B = [point][feature] buffers (MN-major)
__device__ __forceinline__ void issue_dw(uint32_t d_addr, uint32_t a_buf, uint32_t b_buf, int m, int n, uint32_t first_acc) {
  const uint64_t a0 = ptx::make_smem_desc_sw128_ex(a_buf, 16384u, 1024u);
  const uint64_t b0 = ptx::make_smem_desc_sw128_ex(b_buf, 16384u, 1024u);
  const uint32_t idesc = ptx::make_idesc_f16_mn(m, n, 1, 1);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    ptx::mma_f16_ss(d_addr, a0 + (uint64_t)(ks * 128), b0 + (uint64_t)(ks * 128), idesc, ks > 0 ? 1u : first_acc);
}
// D[acc] (M x 16) += A^T B with B a K-major 16 x 128 tile (g_out^T or ones): dW_L^T and the bias gradients
__device__ __forceinline__ void issue_dw_n16(uint32_t d_addr, uint32_t a_buf, uint32_t b_tile, int m, uint32_t first_acc) {
  const uint64_t a0 = ptx::make_smem_desc_sw128_ex(a_buf, 16384u, 1024u);
  const uint64_t b0 = ptx::make_smem_desc_sw128(b_tile);
  const uint32_t idesc = ptx::make_idesc_f16_mn(m, 16, 1, 0);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    ptx::mma_f16_ss(d_addr, a0 + (uint64_t)(ks * 128), kmajor_step(b0, ks, 128u), idesc, ks > 0 ? 1u : first_acc);
}

// ---- epilogues (thread = one row, one half of the columns) ----
// store 16 packed words (32 features starting at `feat`) of row `row` into a [point][feature] SWIZZLE_128B buffer
__device__ __forceinline__ void store_row_words16(uint8_t* buf, int row, int feat, const uint32_t* w) {
  uint8_t* base = buf + (size_t)(feat >> 6) * 16384 + (size_t)row * 128;
  const int ch0 = (feat & 63) >> 3;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<uint4*>(base + (((ch0 + k) ^ (row & 7)) << 4)) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

// load the 16 packed words (32 features starting at `feat`) of row `row` back from such a buffer
__device__ __forceinline__ void load_row_words16(const uint8_t* buf, int row, int feat, uint32_t* w) {
  const uint8_t* base = buf + (size_t)(feat >> 6) * 16384 + (size_t)row * 128;
  const int ch0 = (feat & 63) >> 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + (((ch0 + k) ^ (row & 7)) << 4));
    w[4 * k] = u.x; w[4 * k + 1] = u.y; w[4 * k + 2] = u.z; w[4 * k + 3] = u.w;
  }
}

// 0xFFFF in every half of x that is > 0 (ReLU derivative of the stored activation)
__device__ __forceinline__ uint32_t positive_mask_half2(uint32_t x) {
  return __hgt2_mask(*reinterpret_cast<const __half2*>(&x), __half2{});
}

// forward epilogue: columns [c0, c0 + CPT) of the accumulator -> h = relu(z + b) as fp16 (one fused fma.relu per pair).
// One 32-column load at a time: two in flight were measured slower (64 more live registers spill at the
// 168-register cap of a 9-warp CTA).  W % 32 == 0, so a 32-column chunk is entirely inside or outside the layer.
template <int CPT>
__device__ __forceinline__ void fwd_epilogue(uint32_t d_addr, uint32_t a_addr, const uint32_t* bias2, int c0, int W, bool to_tmem,
                                             uint8_t* sbuf, int row) {
#pragma unroll
  for (int c = 0; c < CPT / 32; ++c) {
    const int col = c0 + 32 * c;
    uint32_t w[16];
    if (col < W) {
      uint32_t v[32];
      ptx::tmem_ld32(d_addr + col, v);
      ptx::tc_wait_ld();
      const uint4* b4 = reinterpret_cast<const uint4*>(bias2 + col / 2);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint4 bb = b4[q4];
        const uint32_t b[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * q4 + i;
          w[j] = ptx::bias_relu_half2(ptx::pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b[i]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = 0u;
    }
    if (to_tmem) ptx::tmem_st16(a_addr + col / 2, w);
    if (sbuf) store_row_words16(sbuf, row, col, w);
  }
}

// chain epilogue: g = D .* [h > 0] as fp16 words; h is read back from its shared-memory buffer (the same
// locations this thread wrote in the forward epilogue and will overwrite with g)
template <int CPT>
__device__ __forceinline__ void chain_epilogue(uint32_t d_addr, int c0, int W, const uint8_t* hbuf, int row, uint32_t (&w)[CPT / 2]) {
#pragma unroll
  for (int c = 0; c < CPT / 32; ++c) {
    const int col = c0 + 32 * c;
    if (col < W) {
      uint32_t hw[16];
      load_row_words16(hbuf, row, col, hw);
      uint32_t v[32];
      ptx::tmem_ld32(d_addr + col, v);
      ptx::tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        w[16 * c + j] = ptx::pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & positive_mask_half2(hw[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) w[16 * c + j] = 0u;
    }
  }
}

__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }

// spilled gradient rows (written by the previous launch of the group): this thread's CPT / 2 packed words
__device__ __forceinline__ void fetch_spill(const uint32_t* src_words, uint32_t (&gin)[32], int cpt) {
  const uint4* src = reinterpret_cast<const uint4*>(src_words);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < cpt / 8) {
      const uint4 u = __ldcs(src + k);
      gin[4 * k] = u.x; gin[4 * k + 1] = u.y; gin[4 * k + 2] = u.z; gin[4 * k + 3] = u.w;
    }
  }
}
// ... -> A operand of the first chain step (TMEM) and operand of dW_top (shared memory)
__device__ __forceinline__ void place_spill(const uint32_t (&gin)[32], int cpt, bool to_tmem, uint32_t a_words_addr, uint8_t* sbuf,
                                            int row, int c0) {
  if (to_tmem) {
    ptx::tmem_st16(a_words_addr, gin);
    if (cpt == 64) ptx::tmem_st16(a_words_addr + 16, gin + 16);
  }
  store_row_words16(sbuf, row, c0, gin);
  if (cpt == 64) store_row_words16(sbuf, row, c0 + 32, gin + 16);
}

template <int OCT, int L>
__global__ void __launch_bounds__(kBwdThreads, 1) bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* wsm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* acts = wsm + (p.im.total_bytes + 1023) / 1024 * 1024;
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(acts + p.plan.acts_bytes);
  const uint32_t wsm_addr = ptx::smem_u32(wsm);
  const uint32_t acts_addr = ptx::smem_u32(acts);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lo = p.plan.lo, hi = p.plan.hi;
  const int top = p.plan.g_in ? hi : L;          // first backward step
  const int fwd_layers = p.plan.g_in ? hi : L;   // forward layers recomputed per tile
  const int W = p.W, WP = p.WP;
  const int MW = WP == 128 ? 128 : 64;  // M of the weight-gradient GEMMs (rows = features of the A buffer)

  if (warp == 0) ptx::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    ptx::mbar_init(&sm.w_ready, 1);
    ptx::mbar_init(&sm.a_ready, 256);
    ptx::mbar_init(&sm.d_ready, 1);
    ptx::mbar_init(&sm.dw_done, 1);
    ptx::fence_mbar_init();
  }
  // constant tiles: ones (16 x 128, every element 1.0) and g_out^T (rows >= dim_out stay zero)
  for (int i = tid; i < 4096 / 4; i += kBwdThreads) {
    reinterpret_cast<uint32_t*>(acts + p.plan.ones_off)[i] = 0x3C003C00u;
    reinterpret_cast<uint32_t*>(acts + p.plan.gt_off)[i] = 0u;
  }
  ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  // per-call gradient scale: max |g_out| * s in (4, 8]
  float gscale = 1.0f, inv_gscale = 1.0f;
  {
    const float mx = __uint_as_float(__ldg(p.absmax_bits));
    if (mx > 0.0f && mx < 3.0e38f) {
      int e;
      frexpf(mx, &e);  // mx = m * 2^e, m in [0.5, 1)
      e = 3 - e;
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      gscale = ldexpf(1.0f, e);
      inv_gscale = ldexpf(1.0f, -e);
    }
  }

  BWD_PHASE_INIT
  const long long total_tiles = p.total_tiles;
  const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;

  uint32_t w_phase = 0;
  uint32_t pa = 0;            // issuer: parity of a_ready
  uint32_t pd = 0, pdw = 0;   // row threads: parities of d_ready / dw_done

  long long t = t_begin;
  while (t < t_end) {
    const long long f = t / p.tiles_per_field;
    const long long seg_begin = f * p.tiles_per_field;
    long long seg_end = seg_begin + p.tiles_per_field;
    if (seg_end > t_end) seg_end = t_end;
    const int ntiles = (int)(seg_end - t);
    const long long tile0_in_field = t - seg_begin;

    if (tid == 0) {  // stage this field's weight image (TMA engine)
      const long long img = p.images_by_slot ? (p.field_slots ? p.field_slots[p.field_base + f] : p.field_base + f) : f;
      const uint8_t* src = p.images + (size_t)img * p.im.total_bytes;
      ptx::mbar_arrive_expect_tx(&sm.w_ready, p.im.total_bytes);
      for (uint32_t o = 0; o < p.im.total_bytes; o += 32768) {
        const uint32_t n = p.im.total_bytes - o < 32768 ? p.im.total_bytes - o : 32768;
        ptx::bulk_g2s(wsm + o, src + o, n, &sm.w_ready);
      }
    }

    if (warp == 0) {
      // ===================== MMA issuer =====================
      ptx::mbar_wait(&sm.w_ready, w_phase);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t wsm_u = __shfl_sync(0xffffffffu, wsm_addr, 0);
      const uint32_t acts_u = __shfl_sync(0xffffffffu, acts_addr, 0);
      const uint32_t d_work = tmem_u + kColWork, a_tm = tmem_u + kColA;
      for (int ti = 0; ti < ntiles; ++ti) {
        const uint32_t first_acc = ti > 0 ? 1u : 0u;  // the first tile of a field segment overwrites the accumulators
#pragma unroll
        for (int l = 0; l < L; ++l) {  // forward recompute (only up to h_hi when the chain starts from a spilled gradient)
          if (l < fwd_layers) {
            BWD_PHASE(11)
            ptx::mbar_wait_spin(&sm.a_ready, pa);
            BWD_PHASE(10)
            pa ^= 1;
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              issue_forward(d_work, a_tm, wsm_u, p.im.layer[l]);
              ptx::mma_commit(&sm.d_ready);
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int l = L; l >= 0; --l) {  // backward step l: chain to g_l, weight gradient of linear l
          if (l >= lo && l <= top) {
            BWD_PHASE(11)
            ptx::mbar_wait_spin(&sm.a_ready, pa);
            BWD_PHASE(10)
            pa ^= 1;
            ptx::tc_fence_after();
            const bool chain = l > lo || (l == lo && lo >= 1 && p.plan.emit_g) || (l == 0 && p.plan.want_denc);
            const bool dw = l <= hi;
            if (ptx::elect_one()) {
              if (chain) {
                issue_chain(d_work, a_tm, wsm_u, p.im.layer[l], in_dim(p, l));
                ptx::mma_commit(&sm.d_ready);
              }
              if (dw) {
                if (l == L) {  // dW_L^T (in x 16) += h_L^T g_out
                  issue_dw_n16(tmem_u + p.plan.col_dw[l], acts_u + p.plan.buf_off[l], acts_u + p.plan.gt_off, MW, first_acc);
                } else {       // dW_l (out x in) += g_{l+1}^T h_l ; db_l += g_{l+1}^T 1
                  issue_dw(tmem_u + p.plan.col_dw[l], acts_u + p.plan.buf_off[l + 1], acts_u + p.plan.buf_off[l], MW,
                           in_dim(p, l), first_acc);
                  issue_dw_n16(tmem_u + p.plan.col_db[l], acts_u + p.plan.buf_off[l + 1], acts_u + p.plan.ones_off, MW, first_acc);
                }
                ptx::mma_commit(&sm.dw_done);
              }
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ===================== row threads =====================
      const int ew = warp - 1;
      const int q = warp & 3;   // TMEM lane quadrant this warp may access
      const int h = ew >> 2;    // column half
      const int row = q * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      const uint32_t d_addr = lane_base + kColWork, a_addr = lane_base + kColA;
      const uint32_t* bias2 = reinterpret_cast<const uint32_t*>(wsm + p.im.bias_h2_off);
      const long long slot = p.field_slots ? p.field_slots[p.field_base + f] : p.field_base + f;
      const int CPT = WP / 2;   // columns per thread
      const int c0 = h * CPT;
      float dbl[8];             // db_L partial sums (lane 0 of the h == 1 warps)
#pragma unroll
      for (int j = 0; j < 8; ++j) dbl[j] = 0.0f;
      bool dw_pending = false;  // a weight-gradient group whose completion has not been observed yet

      ptx::mbar_wait(&sm.w_ready, w_phase);

      for (int ti = 0; ti < ntiles; ++ti) {
        const long long gp = (tile0_in_field + ti) * 128 + row;
        const bool valid = gp < p.points_per_field;
        const long long prow = f * p.points_per_field + (valid ? gp : 0);
        // ---- front end (h == 0): layer-0 A operand; (h == 1): upstream gradient of this row ----
        uint32_t w0[32];
        float go[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) go[j] = 0.0f;
        if (h == 0) {
          if (OCT == 0) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.raw_a + prow * p.EP);
#pragma unroll
            for (int j = 0; j < 32; ++j) w0[j] = (valid && j < p.EP / 2) ? __ldg(src + j) : 0u;
          } else {
            float3 fx = make_float3(0.f, 0.f, 0.f);
            if (valid) {
              const float* src = p.points + prow * 3;
              float3 x = make_float3(__ldg(src), __ldg(src + 1), __ldg(src + 2));
              if (p.positions) {
                const float* c = p.positions + slot * 3;
                const float* qq = p.orientations + slot * 4;
                x = make_float3(x.x - __ldg(c), x.y - __ldg(c + 1), x.z - __ldg(c + 2));
                x = quat_inv_rotate(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3), x);
              }
              fx = scale_local(x, p.scale_mode, p.field_radius);
            }
            if constexpr (OCT > 0) {
              constexpr int EPW = ((6 * OCT + 15) / 16 * 16) / 2;
              uint32_t we[EPW];
              encode_nerf_words<OCT>(fx, p.nerf_start, we);
#pragma unroll
              for (int j = 0; j < 32; ++j) w0[j] = j < EPW ? we[j < EPW ? j : 0] : 0u;
            }
          }
        } else if (valid && !p.plan.g_in) {
          const float* src = p.d_out + prow * p.dim_out;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < p.dim_out) go[j] = __ldg(src + j);
        }
        // spilled gradient g_{hi+1} of this row (this thread's half of the columns), written by the previous launch
        const size_t spill_row = ((size_t)(t + ti) * 128 + row) * (size_t)(WP / 2) + c0 / 2;
        uint32_t gin[32];
        const uint32_t* g_src = p.g_spill + spill_row;
        const bool chain_top = top > lo || (top == 0 && p.plan.want_denc) || (top == lo && lo >= 1 && p.plan.emit_g);
        uint8_t* g_buf = acts + p.plan.buf_off[top + 1 <= L ? top + 1 : L];
        if (p.plan.g_in && fwd_layers == 0) fetch_spill(g_src, gin, CPT);
        // the previous tile's last weight-gradient group reads the buffers this tile is about to overwrite
        BWD_PHASE(0)
        if (dw_pending) {
          ptx::mbar_wait_lean(&sm.dw_done, pdw);
          pdw ^= 1;
          dw_pending = false;
        }
        BWD_PHASE(1)
        ptx::tc_fence_after();
        if (h == 0) {
          if (fwd_layers > 0) {
            ptx::tmem_st16(a_addr, w0);
            if (p.EP > 32) ptx::tmem_st16(a_addr + 16, w0 + 16);
          }
          if (lo == 0) {
            uint8_t* b0 = acts + p.plan.buf_off[0];
            store_row_words16(b0, row, 0, w0);
            if (p.EP > 32) store_row_words16(b0, row, 32, w0 + 16);
          }
        }
        if (p.plan.g_in && fwd_layers == 0) place_spill(gin, CPT, chain_top, a_addr + c0 / 2, g_buf, row, c0);
        ptx::tc_wait_st();
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&sm.a_ready);
        BWD_PHASE(0)

        // ---- forward recompute ----
#pragma unroll
        for (int l = 0; l < L; ++l) {
          if (l >= fwd_layers) continue;
          if (p.plan.g_in && l + 1 == fwd_layers) fetch_spill(g_src, gin, CPT);
          ptx::mbar_wait_lean(&sm.d_ready, pd);
          BWD_PHASE(2)
          pd ^= 1;
          ptx::tc_fence_after();
          const int j = l + 1;  // this epilogue produces h_j
          const bool to_tmem = j < fwd_layers;
          // h_j goes to shared memory when a weight-gradient GEMM reads it (j in [lo, hi]) or the chain needs its
          // ReLU mask (j <= hi + 1): both are the buffer range [lo, min(hi + 1, L)]
          uint8_t* sbuf = (j >= lo && j <= hi + 1) ? acts + p.plan.buf_off[j] : nullptr;
          if (CPT == 64) fwd_epilogue<64>(d_addr, a_addr, bias2 + l * (W / 2), c0, W, to_tmem, sbuf, row);
          else           fwd_epilogue<32>(d_addr, a_addr, bias2 + l * (W / 2), c0, W, to_tmem, sbuf, row);
          if (j == fwd_layers && p.plan.g_in) place_spill(gin, CPT, chain_top, a_addr + c0 / 2, g_buf, row, c0);
          if (j == L && h == 1 && !p.plan.g_in) {
            // g_out of this row: scaled fp16 into the A operand (K = 16) and, transposed, into the 16 x 128 tile
            uint32_t wg[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) wg[k] = ptx::pack_half2(go[2 * k] * gscale, go[2 * k + 1] * gscale);
#pragma unroll
            for (int k = 4; k < 8; ++k) wg[k] = 0u;
            ptx::tmem_st8(a_addr, wg);
            if (hi == L) {
              uint8_t* gt = acts + p.plan.gt_off + (row >> 6) * 2048;
              const int pc = (row & 63) >> 3, pe = row & 7;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < p.dim_out)
                  *reinterpret_cast<__half*>(gt + k * 128 + ((pc ^ (k & 7)) << 4) + pe * 2) = __float2half_rn(go[k] * gscale);
              // db_L: exact fp32 sums of the unscaled gradient
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float s = go[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                dbl[k] += s;
              }
            }
          }
          ptx::tc_wait_st();
          ptx::fence_proxy_async();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&sm.a_ready);
          BWD_PHASE(3)
        }

        // ---- backward chain ----
#pragma unroll
        for (int l = L; l >= 0; --l) {
          if (l >= lo && l <= top) {
            const bool emit = l == lo && lo >= 1 && p.plan.emit_g;
            const bool chain = l > lo || emit || (l == 0 && p.plan.want_denc);
            const bool dw = l <= hi;
            if (chain) {
              ptx::mbar_wait_lean(&sm.d_ready, pd);
              BWD_PHASE(4)
              pd ^= 1;
              ptx::tc_fence_after();
            }
            if (emit) {
              // g_lo = D .* mask(h_lo) -> HBM (fp16 rows): the next launch continues the chain from it
              uint32_t wg[32];
              const uint8_t* hbuf = acts + p.plan.buf_off[l];
              if (CPT == 64) chain_epilogue<64>(d_addr, c0, W, hbuf, row, reinterpret_cast<uint32_t(&)[32]>(wg));
              else           chain_epilogue<32>(d_addr, c0, W, hbuf, row, reinterpret_cast<uint32_t(&)[16]>(wg));
              uint4* dst = reinterpret_cast<uint4*>(p.g_spill + spill_row);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < CPT / 8) __stcs(dst + k, make_uint4(wg[4 * k], wg[4 * k + 1], wg[4 * k + 2], wg[4 * k + 3]));
              ptx::tc_fence_before();
              if (dw) dw_pending = true;
              BWD_PHASE(9)
            } else if (l >= 1 && chain) {
              // g_l = D .* mask(h_l): A operand of the next chain step, and (in place over h_l) operand of dW_{l-1}
              uint32_t wg[32];
              const uint8_t* hbuf = acts + p.plan.buf_off[l];
              if (CPT == 64) chain_epilogue<64>(d_addr, c0, W, hbuf, row, reinterpret_cast<uint32_t(&)[32]>(wg));
              else           chain_epilogue<32>(d_addr, c0, W, hbuf, row, reinterpret_cast<uint32_t(&)[16]>(wg));
              const bool more = (l - 1 > lo) || (l - 1 == lo && lo >= 1 && p.plan.emit_g) ||
                                (l - 1 == 0 && p.plan.want_denc);  // g_l feeds another chain step
              if (more) {
                ptx::tmem_st16(a_addr + c0 / 2, wg);
                if (CPT == 64) ptx::tmem_st16(a_addr + c0 / 2 + 16, wg + 16);
              }
              BWD_PHASE(5)
              if (dw) {  // dW_l of this step reads h_l: wait before overwriting it
                ptx::mbar_wait_lean(&sm.dw_done, pdw);
                pdw ^= 1;
              }
              BWD_PHASE(6)
              if (l - 1 <= hi) {  // dW_{l-1} is accumulated by this launch: it needs g_l in shared memory
                uint8_t* sbuf = acts + p.plan.buf_off[l];
                store_row_words16(sbuf, row, c0, wg);
                if (CPT == 64) store_row_words16(sbuf, row, c0 + 32, wg + 16);
              }
              ptx::tc_wait_st();
              ptx::fence_proxy_async();
              ptx::tc_fence_before();
              ptx::mbar_arrive(&sm.a_ready);
              BWD_PHASE(7)
            } else {
              // l == lo: nothing feeds a further step
              if (l == 0 && chain) {  // dLoss/d h_0 rows -> HBM (gradient of the encoding's parameters)
                for (int cc = h; cc < p.EP / 16; cc += 2) {
                  uint32_t v[16];
                  ptx::tmem_ld16(d_addr + cc * 16, v);
                  ptx::tc_wait_ld();
                  if (valid) {
                    float* dst = p.d_enc + prow * p.E;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      if (cc * 16 + i < p.E) dst[cc * 16 + i] = __uint_as_float(v[i]) * inv_gscale;
                  }
                }
                ptx::tc_fence_before();
              }
              if (dw) dw_pending = true;
            }
          }
        }
      }

      // ---- flush the accumulators of this field segment ----
      BWD_PHASE(5)
      if (dw_pending) {
        ptx::mbar_wait_lean(&sm.dw_done, pdw);
        pdw ^= 1;
        dw_pending = false;
      }
      ptx::tc_fence_after();
      const int frow = MW == 128 ? row : ((lane < 16) ? (q * 16 + lane) : -1);  // feature row held by this TMEM lane
#pragma unroll
      for (int l = 0; l <= L; ++l) {
        if (l < lo || l > hi) continue;
        if (l == L) {
          // dW_L^T: lane = input feature i, column j = output
          if (h == 0) {
            uint32_t v[16];
            ptx::tmem_ld16(lane_base + p.plan.col_dw[l], v);
            ptx::tc_wait_ld();
            if (frow >= 0 && frow < W) {
              float* dst = p.d_w[l] + (size_t)f * p.dim_out * W + frow;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < p.dim_out) red_add(dst + (size_t)j * W, __uint_as_float(v[j]) * inv_gscale);
            }
          }
          if (h == 1 && lane == 0) {  // db_L: every h == 1 warp adds the fp32 sums of its 32 rows
            float* dst = p.d_b[l] + (size_t)f * p.dim_out;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < p.dim_out) red_add(dst + j, dbl[j]);
          }
        } else {
          const int n_in = l == 0 ? p.E : W;       // columns that exist in the gradient tensor
          const int n_cols = in_dim(p, l);         // accumulator columns (padded)
          for (int cc = h; cc < n_cols / 16; cc += 2) {
            uint32_t v[16];
            ptx::tmem_ld16(lane_base + p.plan.col_dw[l] + cc * 16, v);
            ptx::tc_wait_ld();
            if (frow >= 0 && frow < W) {
              float* dst = p.d_w[l] + ((size_t)f * W + frow) * n_in + cc * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (cc * 16 + i < n_in) red_add(dst + i, __uint_as_float(v[i]) * inv_gscale);
            }
          }
          if (h == 0) {
            uint32_t v[4];
            ptx::tmem_ld4(lane_base + p.plan.col_db[l], v);
            ptx::tc_wait_ld();
            if (frow >= 0 && frow < W) red_add(p.d_b[l] + (size_t)f * W + frow, __uint_as_float(v[0]) * inv_gscale);
          }
        }
      }
      ptx::tc_fence_before();
      BWD_PHASE(8)
    }
    w_phase ^= 1;
    t = seg_end;
    ptx::fence_proxy_async();  // generic-proxy reads of the image before the next bulk copy overwrites it
    __syncthreads();
    ptx::tc_fence_after();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// max |x| over a tensor, as float bits (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
  float m = 0.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

uint32_t buf_bytes(const NgmFieldDesc& fd, int l) { return l == 0 ? 16384u : (fd.dim_mlp_out > 64 ? 32768u : 16384u); }

// Largest set [lo, hi] of linears (from `hi` downwards) whose accumulators fit TMEM and whose operand buffers fit
// shared memory next to the weight image.
bool plan_launch(const NgmFieldDesc& fd, const TcImage& im, int hi, bool want_denc, BwdPlan& plan) {
  const int L = fd.num_layers, W = fd.dim_mlp_out, EP = ep_of(fd);
  plan = BwdPlan{};
  plan.hi = hi;
  int best_lo = -1;
  for (int lo = hi; lo >= 0; --lo) {
    int cols = kColAcc;
    for (int l = lo; l <= hi; ++l) cols += (l == L ? 16 : (l == 0 ? EP : W) + 16);
    uint32_t bytes = 0;
    const int top = hi + 1 <= L ? hi + 1 : L;
    for (int l = lo; l <= top; ++l) bytes += buf_bytes(fd, l);
    const size_t smem = 1024 + (size_t)(im.total_bytes + 1023) / 1024 * 1024 + bytes + 8192 + sizeof(BwdSmem) + 64;
    if (cols > 512 || smem > kSmemLimit) break;
    best_lo = lo;
  }
  if (best_lo < 0) return false;
  plan.lo = best_lo;
  int col = kColAcc;
  for (int l = best_lo; l <= hi; ++l) {
    plan.col_dw[l] = col;
    col += l == L ? 16 : (l == 0 ? EP : W);
    if (l < L) { plan.col_db[l] = col; col += 16; }
  }
  uint32_t off = 0;
  const int top = hi + 1 <= L ? hi + 1 : L;
  for (int l = best_lo; l <= top; ++l) { plan.buf_off[l] = off; off += buf_bytes(fd, l); }
  plan.gt_off = off; off += 4096;
  plan.ones_off = off; off += 4096;
  plan.acts_bytes = off;
  plan.want_denc = (want_denc && best_lo == 0) ? 1 : 0;
  // more than one launch: this one hands g_lo to the next (fp16 rows in HBM) instead of the next one recomputing
  // the whole forward and chain; a launch that starts below the last linear continues from that spill
  plan.emit_g = best_lo >= 1 ? 1 : 0;
  plan.g_in = hi < L ? 1 : 0;
  return true;
}

// >= 120 KB so that exactly one CTA (one 512-column TMEM allocation) is resident per SM
size_t bwd_smem_bytes(const TcImage& im, const BwdPlan& plan) {
  const size_t need = 1024 + (size_t)(im.total_bytes + 1023) / 1024 * 1024 + plan.acts_bytes + sizeof(BwdSmem) + 64;
  return need < 120 * 1024 ? 120 * 1024 : need;
}

int bwd_oct(const NgmFieldDesc& fd) {
  return fd.encoding == NGM_ENC_NERF && nerf_octaves_supported(fd.nerf_num_octaves) ? fd.nerf_num_octaves : 0;
}

template <int OCT>
int launch_bwd_l(const BwdParams& p, size_t smem, int grid, cudaStream_t stream) {
  auto go = [&](auto kernel) -> int {
    NGM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kBwdThreads, smem, stream>>>(p);
    return check_launch("bwd_kernel");
  };
  switch (p.L) {
    case 1: return go(bwd_kernel<OCT, 1>);
    case 2: return go(bwd_kernel<OCT, 2>);
    case 3: return go(bwd_kernel<OCT, 3>);
    case 4: return go(bwd_kernel<OCT, 4>);
    default: set_error("tcgen05 backward: num_layers must be in [1, 4], got %d", p.L); return NGM_ERR_UNSUPPORTED;
  }
}

}  // namespace

#ifdef NGM_DEBUG_EXPORTS
// diagnostics: read (and clear) the phase table of CTA 0 (synchronises the device)
int bwd_phases_read(unsigned long long* out16) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out16, g_bwd_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  unsigned long long zero[16] = {0};
  if (cudaMemcpyToSymbol(g_bwd_phase, zero, sizeof(zero)) != cudaSuccess) return -1;
  return 16;
}
#endif

bool field_bwd_tc_supported(const NgmFieldDesc& fd, const char** why) {
  const char* w = nullptr;
  if (!field_tc_supported(fd, &w)) { if (why) *why = w; return false; }
  if (fd.skip_mode != NGM_SKIP_NO) w = "tcgen05 backward: skip connections are differentiated on the fp32 path";
  else if (fd.num_layers < 1 || fd.num_layers > 4) w = "tcgen05 backward: num_layers must be in [1, 4]";
  else if (fd.dim_out > 8) w = "tcgen05 backward: dim_out must be <= 8";
  else if (fd.dim_mlp_out % 32 != 0) w = "tcgen05 backward: dim_mlp_out must be a multiple of 32";
  else {
    const TcImage im = make_image(fd, ep_of(fd));
    BwdPlan plan;
    for (int hi = fd.num_layers; hi >= 0;) {
      if (!plan_launch(fd, im, hi, false, plan)) { w = "tcgen05 backward: field too large for shared memory / TMEM"; break; }
      hi = plan.lo - 1;
    }
  }
  if (why) *why = w;
  return w == nullptr;
}

// fields per launch group: the gradient spill between two launches of a group stays below ~1 GB
constexpr size_t kSpillBudget = (size_t)1 << 30;
long long group_fields(const NgmFieldFwdArgs& a) {
  const long long tiles_per_field = (a.points_per_field + 127) / 128;
  const size_t per_field = (size_t)tiles_per_field * 128 * (a.field.dim_mlp_out > 64 ? 128 : 64) * 2;
  long long g = (long long)(kSpillBudget / (per_field ? per_field : 1));
  if (g < 1) g = 1;
  return g < a.num_fields ? g : a.num_fields;
}
bool needs_spill(const NgmFieldDesc& fd) {
  BwdPlan plan;
  return plan_launch(fd, make_image(fd, ep_of(fd)), fd.num_layers, false, plan) && plan.lo >= 1;
}
size_t spill_bytes(const NgmFieldFwdArgs& a) {
  if (!needs_spill(a.field)) return 0;
  const long long tiles_per_field = (a.points_per_field + 127) / 128;
  return (size_t)group_fields(a) * (size_t)tiles_per_field * 128 * (a.field.dim_mlp_out > 64 ? 128 : 64) * 2;
}

// workspace: [weight images][fp16 rows of pre-encoded encodings][gradient spill of one launch group][absmax]
size_t field_bwd_tc_workspace_bytes(const NgmFieldBwdArgs& b) {
  const NgmFieldFwdArgs& a = b.fwd;
  size_t n = a.field.packed_weights ? 0 :
      ((size_t)make_image(a.field, ep_of(a.field)).total_bytes * (size_t)(a.num_fields > 0 ? a.num_fields : 1) + 255) / 256 * 256;
  if (bwd_oct(a.field) == 0 && !a.rows_half)
    n += ((size_t)a.num_fields * (size_t)a.points_per_field * (size_t)ep_of(a.field) * 2 + 255) / 256 * 256;
  n += (spill_bytes(a) + 255) / 256 * 256;
  return n + 256;
}

int launch_field_bwd_tc(const NgmFieldBwdArgs& b, cudaStream_t stream) {
  const NgmFieldFwdArgs& a = b.fwd;
  const NgmFieldDesc& fd = a.field;
  BwdParams p{};
  p.E = fd.dim_encoding;
  p.EP = ep_of(fd);
  p.W = fd.dim_mlp_out;
  p.WP = fd.dim_mlp_out > 64 ? 128 : 64;
  p.L = fd.num_layers;
  p.dim_out = fd.dim_out;
  p.nerf_start = fd.nerf_start_octave;
  p.im = make_image(fd, p.EP);
  p.num_fields = a.num_fields;
  p.positions = a.positions;
  p.orientations = a.orientations;
  p.field_slots = reinterpret_cast<const long long*>(a.field_slots);
  p.scale_mode = a.scale_mode;
  p.field_radius = a.field_radius;
  p.points = a.points;
  p.points_per_field = a.points_per_field;
  p.d_out = b.d_out;
  p.d_enc = b.d_encoding;
  p.tiles_per_field = (a.points_per_field + 127) / 128;
  p.total_tiles = p.tiles_per_field * a.num_fields;
  char* ws = static_cast<char*>(a.workspace);
  size_t off = 0;
  if (fd.packed_weights) {
    p.images = static_cast<const uint8_t*>(fd.packed_weights);
    p.images_by_slot = 1;
  } else {
    p.images = reinterpret_cast<const uint8_t*>(ws);
    off = ((size_t)p.im.total_bytes * (size_t)a.num_fields + 255) / 256 * 256;
    PackParams pk{fd, p.im, p.field_slots, reinterpret_cast<uint8_t*>(ws), fd.dim_encoding, 0};
    pack_weights_kernel<<<a.num_fields, 256, 0, stream>>>(pk);
    if (int rc = check_launch("pack_weights_kernel")) return rc;
  }
  const int oct = bwd_oct(fd);
  if (oct == 0) {
    if (a.rows_half) {
      p.raw_a = static_cast<const __half*>(a.rows_half);
    } else {
      PermutoRowsArgs e{};
      e.field = fd;
      e.points_world = a.points;
      e.positions = a.positions; e.orientations = a.orientations;
      e.field_slots = p.field_slots;
      e.out = reinterpret_cast<uint32_t*>(ws + off);
      e.num_points = (long long)a.num_fields * a.points_per_field;
      e.points_per_field = a.points_per_field;
      e.field_radius = a.field_radius;
      e.scale_mode = a.scale_mode;
      e.EP = p.EP;
      if (int rc = launch_permuto_rows_half(e, stream)) return rc;
      p.raw_a = reinterpret_cast<const __half*>(e.out);
      off += ((size_t)e.num_points * (size_t)p.EP * 2 + 255) / 256 * 256;
    }
  }
  p.g_spill = reinterpret_cast<uint32_t*>(ws + off);
  off += (spill_bytes(a) + 255) / 256 * 256;
  unsigned* absmax = reinterpret_cast<unsigned*>(ws + off);
  NGM_CUDA(cudaMemsetAsync(absmax, 0, sizeof(unsigned), stream));
  const long long n_out = (long long)a.num_fields * a.points_per_field * fd.dim_out;
  absmax_kernel<<<num_sms() * 4, 256, 0, stream>>>(b.d_out, n_out, absmax);
  if (int rc = check_launch("absmax_kernel")) return rc;
  p.absmax_bits = absmax;
  const int L = fd.num_layers, W = fd.dim_mlp_out;
  for (int l = 0; l <= L; ++l) {
    const size_t out_l = l == L ? fd.dim_out : W, in_l = l == 0 ? fd.dim_encoding : W;
    NGM_CHECK_ARG(b.d_weights[l] && b.d_biases[l], "d_weights / d_biases of linear %d missing", l);
    NGM_CUDA(cudaMemsetAsync(b.d_weights[l], 0, (size_t)a.num_fields * out_l * in_l * sizeof(float), stream));
    NGM_CUDA(cudaMemsetAsync(b.d_biases[l], 0, (size_t)a.num_fields * out_l * sizeof(float), stream));
    p.d_w[l] = b.d_weights[l];
    p.d_b[l] = b.d_biases[l];
  }
  const char* e = getenv("NGM_TC_MAX_CTAS");
  const int cap = e ? atoi(e) : 0;
  long long sms = num_sms();
  if (cap > 0 && cap < sms) sms = cap;
  // launch groups of fields (bounded gradient spill); inside a group the launches go from the last linear down
  const long long gf = needs_spill(fd) ? group_fields(a) : a.num_fields;
  const BwdParams base = p;
  for (long long f0 = 0; f0 < a.num_fields; f0 += gf) {
    const long long nf = a.num_fields - f0 < gf ? a.num_fields - f0 : gf;
    p = base;
    p.field_base = f0;
    p.num_fields = (int)nf;
    if (!base.images_by_slot) p.images = base.images + (size_t)f0 * p.im.total_bytes;
    const size_t pt0 = (size_t)f0 * (size_t)a.points_per_field;
    if (base.points) p.points = base.points + pt0 * 3;
    if (base.raw_a) p.raw_a = base.raw_a + pt0 * (size_t)p.EP;
    p.d_out = base.d_out + pt0 * (size_t)fd.dim_out;
    if (base.d_enc) p.d_enc = base.d_enc + pt0 * (size_t)fd.dim_encoding;
    for (int l = 0; l <= L; ++l) {
      const size_t out_l = l == L ? fd.dim_out : W, in_l = l == 0 ? fd.dim_encoding : W;
      p.d_w[l] = base.d_w[l] + (size_t)f0 * out_l * in_l;
      p.d_b[l] = base.d_b[l] + (size_t)f0 * out_l;
    }
    p.total_tiles = p.tiles_per_field * nf;
    const int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
    for (int hi = L; hi >= 0;) {
      if (!plan_launch(fd, p.im, hi, b.d_encoding != nullptr, p.plan)) {
        set_error("tcgen05 backward: field too large for shared memory / TMEM");
        return NGM_ERR_UNSUPPORTED;
      }
      const size_t smem = bwd_smem_bytes(p.im, p.plan);
      int rc;
      switch (oct) {
        case 4: rc = launch_bwd_l<4>(p, smem, grid, stream); break;
        case 8: rc = launch_bwd_l<8>(p, smem, grid, stream); break;
        default: rc = launch_bwd_l<0>(p, smem, grid, stream); break;
      }
      if (rc) return rc;
      hi = p.plan.lo - 1;
    }
  }
  return NGM_OK;
}

}  // namespace ngm
