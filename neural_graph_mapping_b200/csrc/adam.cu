// Adam on the ACTIVE rows of the stacked per-field parameter tensors, in place.
//
// The reference trains a subset of fields per iteration: it gathers their rows into
// vmap_fields_params (ngm/models.py:266-276), lets torch.optim.Adam update the gathered copies
// (ngm/run_mapping.py:347-362, 1191-1193; plain Adam, L2 weight decay added to the gradient), and then
// scatters parameters and both moment tensors back into the full tables, after having gathered the
// moments on the way in (ngm/run_mapping.py:679-707, 1195-1221): per parameter tensor one gather and
// one scatter of three tensors around ~8 foreach kernels.  Here ONE launch per training step reads
// (gradient row, parameter row, both moment rows) of every active field straight from the full tables
// and writes them back in place, plus the refreshed active copy the next render reads.
//
// Arithmetic = torch's multi-tensor Adam, operation for operation, in fp32:
//   g += wd * w;  m = m + (g - m) (1 - b1);  v = v b2 + (1 - b2) g g;
//   w -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with the two bias corrections evaluated on the host in double precision (as torch does).
// HBM-bound elementwise work: 16 B read + 16 B written per element (+ 4 B for the active copy).
#include "common.cuh"

namespace ngm {

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamPerThread = 4;
constexpr int kAdamChunk = kAdamThreads * kAdamPerThread;

struct AdamTable {
  NgmAdamParam p[NGM_ADAM_MAX_PARAMS];
  int block_begin[NGM_ADAM_MAX_PARAMS + 1];  // first x-block of each tensor's row
  int n;
};

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const __grid_constant__ AdamTable t,
                                                                 const long long* __restrict__ field_ids,
                                                                 long long num_active, float step_size, float w1,
                                                                 float beta2, float w2, float bc2_sqrt, float eps,
                                                                 float wd) {
  int which = 0;
  while (which + 1 < t.n && (int)blockIdx.x >= t.block_begin[which + 1]) ++which;
  const NgmAdamParam& d = t.p[which];
  const long long e0 = (long long)((int)blockIdx.x - t.block_begin[which]) * kAdamChunk + threadIdx.x;
  for (long long a = blockIdx.y; a < num_active; a += gridDim.y) {
    const long long slot = field_ids ? field_ids[a] : a;
    const float* __restrict__ grad = d.grad + a * d.row;
    float* __restrict__ w_all = d.param_all + slot * d.row;
    float* __restrict__ m_all = d.exp_avg_all + slot * d.row;
    float* __restrict__ v_all = d.exp_avg_sq_all + slot * d.row;
    float* __restrict__ w_act = d.param_active ? d.param_active + a * d.row : nullptr;
#pragma unroll
    for (int i = 0; i < kAdamPerThread; ++i) {
      const long long e = e0 + i * kAdamThreads;
      if (e < d.row) {
        float g = __ldg(grad + e);
        float w = w_all[e], m = m_all[e], v = v_all[e];
        if (wd != 0.0f) g = __fmaf_rn(wd, w, g);
        m = __fmaf_rn(g - m, w1, m);
        v = __fmaf_rn(w2 * g, g, v * beta2);
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
        w = __fmaf_rn(-step_size, __fdiv_rn(m, denom), w);
        w_all[e] = w;
        m_all[e] = m;
        v_all[e] = v;
        if (w_act) w_act[e] = w;
      }
    }
  }
}

}  // namespace

int launch_adam_step(const NgmAdamArgs& a, cudaStream_t stream) {
  AdamTable t;
  t.n = 0;
  int blocks = 0;
  for (int i = 0; i < a.num_params; ++i) {
    const NgmAdamParam& d = a.params[i];
    if (d.row == 0) continue;
    t.p[t.n] = d;
    t.block_begin[t.n] = blocks;
    blocks += (int)((d.row + kAdamChunk - 1) / kAdamChunk);
    ++t.n;
  }
  t.block_begin[t.n] = blocks;
  if (t.n == 0 || a.num_active == 0) return NGM_OK;
  // scalars exactly as torch forms them: Python-float (double) arithmetic, rounded to fp32 at the kernel boundary
  const double bc1 = 1.0 - pow(a.beta1, (double)a.step);
  const double bc2 = 1.0 - pow(a.beta2, (double)a.step);
  const float step_size = (float)(a.lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const dim3 grid((unsigned)blocks, (unsigned)(a.num_active < 65535 ? a.num_active : 65535));
  adam_step_kernel<<<grid, kAdamThreads, 0, stream>>>(t, reinterpret_cast<const long long*>(a.field_ids), a.num_active,
                                                      step_size, (float)(1.0 - a.beta1), (float)a.beta2, (float)(1.0 - a.beta2),
                                                      bc2_sqrt, (float)a.eps, (float)a.weight_decay);
  return check_launch("adam_step_kernel");
}

}  // namespace ngm
