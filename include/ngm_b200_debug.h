/*
 * ngm_b200_debug.h -- diagnostic entry points that exist ONLY in libngm_b200_debug.so (the same sources
 * compiled with -DNGM_DEBUG_EXPORTS; `python -m neural_graph_mapping_b200._build --debug`).  The product
 * library libngm_b200.so and its header ngm_b200.h carry only the render path.
 */
#ifndef NGM_B200_DEBUG_H_
#define NGM_B200_DEBUG_H_

#include "ngm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* diagnostics: out(rows,n) = A(rows,k; fp16) x weight(n,k; fp32 -> fp16)^T + bias through the production
 * tcgen05 plumbing (weight packing, SWIZZLE_128B descriptors, A operand in TMEM, TMEM epilogue).
 * k % 16 == 0, 16 <= k <= 128, n <= 128; workspace >= 64 KiB. */
int ngm_debug_tc_gemm(const float* weight, const float* bias, int n, int k, const void* a_half, int64_t rows, float* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* diagnostics: with NGM_TC_TRACE=1 in the environment, CTA 0 of every tcgen05 launch records (event, clock)
 * pairs; this copies them to HOST memory (synchronises the device); returns the event count. */
int ngm_debug_tc_trace(uint64_t* host_out, int max_events);
/* same without synchronising the device (reads the trace of a still-running kernel; deadlock diagnosis) */
int ngm_debug_tc_trace_peek(uint64_t* host_out, int max_events);

/* diagnostics: cycles CTA 0 of the tcgen05 backward launches spent per phase since the last call (16 counters:
 * row thread 0 front end, 1 wait previous dW, 2 wait forward MMA, 3 forward epilogue, 4 wait chain MMA, 5 chain epilogue,
 * 6 wait dW, 7 store + arrive, 8 flush; issuer 10 wait for operands, 11 issue); synchronises the device. */
int ngm_debug_bwd_phases(uint64_t* host_out16);

#ifdef __cplusplus
}
#endif
#endif /* NGM_B200_DEBUG_H_ */
