// Inline-PTX wrappers for the sm_100a features the tensor-core path uses:
// mbarrier, cp.async.bulk (TMA engine, 1-D), tcgen05.{alloc,mma,commit,ld,st,fence}.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ngm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the
// hint elapses) instead of spinning and burning issue slots that the working warps need.
__device__ __forceinline__ bool mbar_try_hint(uint64_t* bar, uint32_t parity, uint32_t ticks) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ticks)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (sticky CUDA error, caught by the host) instead of
// hanging the GPU (~4 s).  The clock is only consulted every 256 failed attempts.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_hint(bar, parity, 0x989680u)) return;
  long long t0 = 0;
#pragma unroll 1
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_hint(bar, parity, 0x989680u)) return;
    if ((it & 255u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) {
        printf("ngm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
               smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// Unbounded-in-time but iteration-bounded wait with the smallest possible loop body (used on the
// kernel's critical paths): a waiting warp costs ~4 issue slots per wake-up instead of ~13.
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  uint32_t it = 0;
  long long t0 = 0;
#pragma unroll 1
  while (!mbar_try_hint(bar, parity, 0x989680u)) {
    if ((++it & 63u) == 0u) {  // a failed attempt may sleep for a long time when nothing happens: bound by the clock
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) {
        printf("ngm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
               smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// Pure spin on test_wait (no hardware suspend): lowest wake-up latency, burns issue slots -- only for
// the two MMA-issuer warps, whose wake-up sits on the kernel's critical path.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t it = 0;
  long long t0 = 0;
#pragma unroll 1
  while (!mbar_test(bar, parity)) {
    if ((++it & 0xFFFFu) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) {
        printf("ngm: mbarrier spin wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
               smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// named barrier among a subset of warps
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- async proxy / bulk copy -------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); bytes % 16 == 0, both 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------
// whole warp; writes the TMEM base address to *smem_out
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]   (kind::f16: fp16 operands, fp32 accumulate), one thread issues
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16), one thread issues
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// instruction descriptor, kind::f16: A,B fp16 K-major, D FP16 (16-bit accumulation inside the tensor core), M=128
__host__ __device__ constexpr uint32_t make_idesc_f16_acc16(int n) {
  return ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // c_format = F16 (0)
}
// instruction descriptor, kind::f16: A,B fp16 K-major, D fp32, M=128
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)       // a_format = b_format = F16
         | (0u << 15) | (0u << 16)      // a_major = b_major = K
         | ((uint32_t)(n >> 3) << 17)   // n_dim
         | ((uint32_t)(128 >> 4) << 24);  // m_dim
}

// general instruction descriptor, kind::f16: fp16 operands, fp32 accumulate; a_mn / b_mn = 1 selects an MN-major
// (transposed) shared-memory operand (cute/arch/mma_sm100_desc.hpp InstrDescriptor: a_major bit 15, b_major bit 16)
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// SWIZZLE_128B shared-memory descriptor with explicit leading / stride byte offsets.
//   K-major operand : rows of 128 B (64 halves of K), 8-row groups `sbo` bytes apart (lbo unused)
//   MN-major operand: rows of 128 B hold 64 consecutive MN elements of ONE k index; 8 k rows form a 1024-B group,
//                     groups `sbo` bytes apart, the next 64 MN elements `lbo` bytes apart
//                     (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, cute/atom/mma_traits_sm100.hpp)
// The same [row][128 B] swizzled tile is therefore readable as K-major (K = its 64 columns) and as MN-major
// (MN = its 64 columns, K = its rows).
__device__ __forceinline__ uint64_t make_smem_desc_sw128_ex(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

// TMEM <-> registers, 32 lanes x 32-bit, N consecutive columns; lane of the thread = 32*(warp%4)+laneid
__device__ __forceinline__ void tmem_ld4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}
// 16-bit accumulators: .pack::16b packs the low halves of two adjacent 32-bit columns into one register, so N
// registers cover 2N columns (the accumulator of a kind::f16 MMA with D format F16 keeps one value per column)
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld16_pack16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_st1(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "r"(r[0]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// store N consecutive words (compile-time N) as a sequence of x16/x8/x4/x2/x1 stores
template <int N>
__device__ __forceinline__ void tmem_store_n(uint32_t addr, const uint32_t* r) {
  if constexpr (N >= 16) {
    tmem_st16(addr, r);
    tmem_store_n<N - 16>(addr + 16, r + 16);
  } else if constexpr (N >= 8) {
    tmem_st8(addr, r);
    tmem_store_n<N - 8>(addr + 8, r + 8);
  } else if constexpr (N >= 4) {
    tmem_st4(addr, r);
    tmem_store_n<N - 4>(addr + 4, r + 4);
  } else if constexpr (N >= 2) {
    tmem_st2(addr, r);
    tmem_store_n<N - 2>(addr + 2, r + 2);
  } else if constexpr (N == 1) {
    tmem_st1(addr, r);
  }
}

// two fp32 -> packed half2 (lo = a, hi = b), round-to-nearest
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// x + bias on packed half2 (no activation: the sign bits are the ReLU mask)
__device__ __forceinline__ uint32_t bias_half2(uint32_t x, uint32_t bias) {
  uint32_t d;
  const uint32_t one = 0x3C003C00u;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(bias));
  return d;
}
__device__ __forceinline__ uint32_t relu_half2(uint32_t x) {
  uint32_t d;
  const uint32_t zero = 0u;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(zero));
  return d;
}
// 0xFFFF in every 16-bit half of x whose sign bit is set, 0 elsewhere (byte permute with sign replication)
__device__ __forceinline__ uint32_t sign_mask_half2(uint32_t x) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xBB99u));
  return d;
}
// relu(x + bias) on packed half2
__device__ __forceinline__ uint32_t bias_relu_half2(uint32_t x, uint32_t bias) {
  uint32_t d;
  const uint32_t one = 0x3C003C00u;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(bias));
  return d;
}

}  // namespace ptx
}  // namespace ngm
