// Pieces shared by the tcgen05 forward kernels (field_tc.cu) and the tcgen05 backward kernel (field_tc_bwd.cu):
// the per-field fp16 weight image (UMMA K-major SWIZZLE_128B layout), its packing kernel, and the NeRF front end.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ngm {
namespace {

constexpr uint32_t kMaxImageBytes = 200 * 1024;

struct TcLayer {
  int n_pad, k_pad, atoms;
  uint32_t off;  // byte offset of the layer's B image inside the field image
};

struct TcImage {
  int num_linears;
  TcLayer layer[NGM_MAX_LINEARS];
  uint32_t bias_h2_off;    // uint32 half2 biases of the hidden layers: [L][W/2]
  uint32_t bias_last_off;  // fp32 biases of the last layer: [n_last_pad]
  uint32_t total_bytes;    // multiple of 16
};

inline TcImage make_image(const NgmFieldDesc& fd, int EP) {
  TcImage im{};
  const int L = fd.num_layers, W = fd.dim_mlp_out;
  im.num_linears = L + 1;
  uint32_t off = 0;
  for (int l = 0; l <= L; ++l) {
    TcLayer& y = im.layer[l];
    y.n_pad = l == L ? (fd.dim_out + 15) / 16 * 16 : W;
    // skip mode "concat" (ngm/models.py:160-161): every linear after the first reads [activations | encoding]
    y.k_pad = l == 0 ? EP : (fd.skip_mode == NGM_SKIP_CONCAT ? W + EP : W);
    y.atoms = (y.k_pad + 63) / 64;
    y.off = off;
    off += (uint32_t)y.atoms * y.n_pad * 128;
  }
  im.bias_h2_off = off;
  off += (uint32_t)L * (W / 2) * 4;
  off = (off + 15) / 16 * 16;
  im.bias_last_off = off;
  off += (uint32_t)im.layer[L].n_pad * 4;
  im.total_bytes = (off + 15) / 16 * 16;
  return im;
}

// ---- weight packing: fp32 (out,in) tables -> per-field fp16 image in the UMMA K-major SWIZZLE_128B layout ----
struct PackParams {
  NgmFieldDesc fd;
  TcImage im;
  const long long* field_slots;
  uint8_t* images;
  int E;
  int dst_by_slot;  // persistent images: row r of the tables -> image r (instead of the call's field index)
};

__global__ void __launch_bounds__(256) pack_weights_kernel(PackParams p) {
  const int f = blockIdx.x;
  const long long slot = p.field_slots ? p.field_slots[f] : f;
  uint8_t* img = p.images + (size_t)(p.dst_by_slot ? slot : f) * p.im.total_bytes;
  const int L = p.fd.num_layers, W = p.fd.dim_mlp_out;
  for (int l = 0; l <= L; ++l) {
    const TcLayer y = p.im.layer[l];
    const int N = l == L ? p.fd.dim_out : W;
    const int K = l == 0 ? p.E : (p.fd.skip_mode == NGM_SKIP_CONCAT ? W + p.E : W);
    const float* Wg = p.fd.weights[l] + slot * p.fd.weight_stride[l];
    const int chunks = y.atoms * y.n_pad * 8;  // 16-byte chunks (8 halves)
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
      const int atom = c / (y.n_pad * 8);
      const int n = (c / 8) % y.n_pad;
      const int ck = c % 8;
      __half h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = atom * 64 + ck * 8 + i;
        h[i] = __float2half_rn((n < N && k < K) ? __ldg(Wg + (size_t)n * K + k) : 0.0f);
      }
      uint8_t* dst = img + y.off + (size_t)atom * y.n_pad * 128 + (size_t)n * 128 + ((ck ^ (n & 7)) * 16);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
    }
  }
  uint32_t* bh = reinterpret_cast<uint32_t*>(img + p.im.bias_h2_off);
  for (int l = 0; l < L; ++l) {
    const float* Bg = p.fd.biases[l] + slot * p.fd.bias_stride[l];
    for (int j = threadIdx.x; j < W / 2; j += blockDim.x)
      bh[l * (W / 2) + j] = ptx::pack_half2(__ldg(Bg + 2 * j), __ldg(Bg + 2 * j + 1));
  }
  float* bl = reinterpret_cast<float*>(img + p.im.bias_last_off);
  const float* Bg = p.fd.biases[L] + slot * p.fd.bias_stride[L];
  for (int j = threadIdx.x; j < p.im.layer[L].n_pad; j += blockDim.x) bl[j] = j < p.fd.dim_out ? __ldg(Bg + j) : 0.0f;
}


inline int nerf_octaves_supported(int o) { return o == 4 || o == 8; }
inline int ep_of(const NgmFieldDesc& fd) { return (fd.dim_encoding + 15) / 16 * 16; }

// sin(pi t), cos(pi t) with exact range reduction to [-1, 1], MUFU evaluation
__device__ __forceinline__ void sincospi_fast(float t, float& s, float& c) {
  const float r = fmaf(-2.0f, rintf(0.5f * t), t);
  const float a = 3.14159265358979f * r;
  s = __sinf(a);
  c = __cosf(a);
}

// NeRF features of one row as packed fp16 words (the layer-0 A operand; one thread per row).  Octaves 0 and 4 are
// evaluated directly (exact range reduction + MUFU), the others by the double-angle recurrence: the
// error at most doubles per octave, 3 steps -> < 1e-5, far below fp16 resolution.
template <int OCT>
__device__ __forceinline__ void encode_nerf_words(float3 x, int start_octave, uint32_t (&w)[((6 * OCT + 15) / 16 * 16) / 2]) {
  constexpr int E = 6 * OCT;
  constexpr int EP = (E + 15) / 16 * 16;
  const float base = exp2f((float)start_octave);
  const float xs[3] = {x.x, x.y, x.z};
  float fe[EP];
#pragma unroll
  for (int i = E; i < EP; ++i) fe[i] = 0.0f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float t0 = xs[d] * base;
    float s = 0.f, c = 1.f;
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
      if (o % 4 == 0) {
        sincospi_fast(t0 * (float)(1 << o), s, c);
      } else {
        const float s2 = 2.0f * s * c;
        c = fmaf(-2.0f * s, s, 1.0f);
        s = s2;
      }
      fe[d * OCT + o] = s;
      fe[3 * OCT + d * OCT + o] = c;
    }
  }
#pragma unroll
  for (int j = 0; j < EP / 2; ++j) w[j] = ptx::pack_half2(fe[2 * j], fe[2 * j + 1]);
}

// ... -> TMEM A operand
template <int OCT>
__device__ __forceinline__ void encode_nerf_to_tmem(uint32_t a_addr, float3 x, int start_octave) {
  constexpr int EP = (6 * OCT + 15) / 16 * 16;
  uint32_t w[EP / 2];
  encode_nerf_words<OCT>(x, start_octave, w);
  ptx::tmem_store_n<EP / 2>(a_addr, w);
}

}  // namespace
}  // namespace ngm
