/* lass_b200 — tcgen05 descriptor probes and issue-rate microbenchmarks (csrc/probe.cu).
 *
 * NOT part of the product library: these entry points live in lass_b200/_lib/liblass_b200_debug.so, which only the GPU
 * tests that pin the hardware rules (tests/test_gpu_umma_probe.py) and the tools/gpu_umma_bench*.py scripts load.
 * Same conventions as include/lass_b200.h (raw device pointers, int return code, lass_last_error of THIS library). */
#ifndef LASS_B200_DEBUG_H_
#define LASS_B200_DEBUG_H_
#include "lass_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------
 * Debug: one tcgen05.mma tile (M = 128) with caller-controlled shared-memory descriptors; used by the GPU
 * tests to pin the descriptor rules the conv kernel relies on.  A (a_rows, kc) and Bm (n, kc) are 16-bit
 * K-major; out (128, n) fp32.  swizzle_mode: 0 none, 2 = 128 B, 4 = 64 B, 6 = 32 B.
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_debug_umma_probe(const void* A, int a_rows, const void* Bm, int n, int kc, int swizzle_mode,
                                   int a_start_bytes, int a_sbo, int a_base_offset, int b_sbo, int fmt_fp16,
                                   float* out, void* stream);

/* Debug: the same for MN-major operands (rows of the tile = contraction index K; the weight-gradient kernel's layouts): A
 * (a_rows, 64 | 32) and Bm (b_rows, 64 | 32) 16-bit (64 elements per row with swizzle 2 = 128 B, 32 with 4 = 64 B); `ksteps`
 * MMAs (M = 128, N = n, K = 16) with start = tile + *_start + ks * *_kstep bytes, leading / stride byte offsets *_lbo /
 * *_sbo; a_fp16 / b_fp16 select the operand formats — they must be EQUAL: a mixed fp16 x bf16
 * instruction descriptor is an illegal instruction on sm_100a (measured).  out (128, n) fp32. */
LASS_API int lass_debug_umma_probe_mn(const void* A, int a_rows, int a_swz, const void* Bm, int b_rows, int b_swz, int n,
                                      int ksteps, int a_start, int a_lbo, int a_sbo, int a_kstep, int b_start, int b_lbo,
                                      int b_sbo, int b_kstep, int a_fp16, int b_fp16, float* out, void* stream);

/* Debug: tcgen05.mma issue/execute throughput for a given operand layout: `iters` back-to-back MMAs (M = 128,
 * N = n, K = 16) over zeroed shared memory, `nacc` accumulators round-robin; cycles_out[grid] receives the SM
 * clock cycles from first issue to completion. */
LASS_API int lass_debug_umma_bench(int n, int kc, int swizzle_mode, int a_start_bytes, int a_sbo, int iters, int nacc,
                                   int grid, long long* cycles_out, void* stream);
/* Same measurement with the issue loop unrolled 8x (about two instructions per MMA from the issuing thread), so
 * that MMAs shorter than the first version's loop overhead are resolved.  nacc in {1, 2}; iters % 8 == 0. */
LASS_API int lass_debug_umma_bench2(int n, int kc, int swizzle_mode, int a_start_bytes, int a_sbo, int iters, int nacc,
                                    int grid, long long* cycles_out, void* stream);
/* Issue-rate benchmark of the conv kernel's own steady-state MMA issue code (one halo chunk = 9 taps x mt m-tiles x
 * ksteps k-steps per item, nothing else running); mode 0 = running descriptors, 1 = per-tap re-derived descriptors. */
LASS_API int lass_debug_umma_bench3(int mt, int bn, int ksteps, int mode, int iters, int grid, long long* cycles_out,
                                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LASS_B200_DEBUG_H_ */
