// Diagnostics: TMEM load/store bandwidth micro-benchmark (one CTA per SM).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ngm {
namespace {

__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int iters, int mode, unsigned long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 512;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  ptx::tmem_st16(base, v);
  ptx::tmem_st16(base + 16, v + 16);
  ptx::tmem_st16(base + 32, v);
  ptx::tmem_st16(base + 48, v + 16);
  ptx::tc_wait_st();
  __syncthreads();
  const long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {  // load x32, wait each
      ptx::tmem_ld32(base, v);
      ptx::tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= v[i];
    } else if (mode == 1) {  // two loads in flight
      uint32_t w[32];
      ptx::tmem_ld32(base, v);
      ptx::tmem_ld32(base + 32, w);
      ptx::tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= v[i] ^ w[i];
    } else if (mode == 2) {  // store x16 (half2 activations)
      ptx::tmem_st16(base, v);
      ptx::tc_wait_st();
    } else {  // load x16
      uint32_t w[16];
      ptx::tmem_ld16(base, w);
      ptx::tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; i += 4) acc ^= w[i];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) out[blockIdx.x + gridDim.x] = acc;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

int tmem_bw_bench(int warps, int iters, int mode, unsigned long long* host_cycles) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 2 * 148 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  tmem_bw_kernel<<<num_sms() > 148 ? 148 : num_sms(), warps * 32>>>(iters, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(host_cycles, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { set_error("tmem bench: %s", cudaGetErrorString(e)); return NGM_ERR_CUDA; }
  return NGM_OK;
}

}  // namespace ngm
