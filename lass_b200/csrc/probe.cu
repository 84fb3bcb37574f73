// Debug entry point: one tcgen05.mma tile with caller-controlled shared-memory descriptors.
//
// Used by tests/test_umma_probe.py on the B200 to pin down the descriptor rules the implicit-GEMM conv
// kernel relies on (swizzle modes, 8-row-group stride, start addresses that are shifted by whole rows).
// A (rows x kc) and B (n x kc) are 16-bit K-major matrices; A is brought into shared memory by ONE TMA box
// of `a_rows` rows, B by one box of `n` rows; then kc/16 MMAs (M = 128) are issued with
//   desc_a = make_smem_desc(A_smem + a_start_bytes + ks*32, a_sbo, swizzle, a_base_offset)
// and the 128 x n fp32 accumulator is written to `out`.
#include "conv_issue.cuh"
#include "lass_internal.cuh"
#include "../../include/lass_b200_debug.h"
#include "ptx.cuh"

namespace lass {
namespace {

struct ProbeParams {
  CUtensorMap tmA, tmB;
  float* out;
  int a_rows, n, kc, swizzle, a_start_bytes, a_sbo, a_base_offset, b_sbo, fmt;
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int row_bytes = p.kc * 2;
  unsigned char* sA = smem;
  unsigned char* sB = smem + ((p.a_rows * row_bytes + 1023) & ~1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + ((p.n * row_bytes + 1023) & ~1023));
  uint64_t* acc_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)((p.a_rows + p.n) * row_bytes));
    tma_load_2d(sA, &p.tmA, bar, 0, 0);
    tma_load_2d(sB, &p.tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_f16(p.fmt, p.fmt, 128, p.n);
    for (int ks = 0; ks < p.kc / 16; ++ks) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + p.a_start_bytes + ks * 32, p.a_sbo, p.swizzle, p.a_base_offset);
      const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 32, p.b_sbo, p.swizzle, 0);
      umma_f16(tmem_acc, da, db, idesc, ks != 0);
    }
    umma_commit(acc_bar);
  }
  __syncwarp();
  mbar_wait(acc_bar, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < p.n; c += 16) {
    float v[16];
    tmem_ld_x16(taddr + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[(size_t)(warp * 32 + lane) * p.n + c + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, 256);
  }
}

// MN-major probe: A (a_rows x wa) and B (b_rows x wb) 16-bit tiles whose ROWS are the contraction index K (wa / wb = 64
// elements with SWIZZLE_128B, 32 with SWIZZLE_64B), one TMA box each; `ksteps` MMAs (M = 128, K = 16) with
//   desc_a = make_smem_desc_mn(A_smem + a_start + ks*a_kstep, a_lbo, a_sbo, a_swz), desc_b likewise; out (128, n) fp32.
struct ProbeMnParams {
  CUtensorMap tmA, tmB;
  float* out;
  int a_rows, b_rows, wa, wb, n, ksteps;
  int a_swz, a_start, a_lbo, a_sbo, a_kstep, b_swz, b_start, b_lbo, b_sbo, b_kstep, fmt_a, fmt_b;
};

__global__ void __launch_bounds__(128, 1) umma_probe_mn_kernel(const __grid_constant__ ProbeMnParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 1024);
  uint64_t* acc_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;
  if (threadIdx.x == 0) {
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bar, (uint32_t)((p.a_rows * p.wa + p.b_rows * p.wb) * 2));
    tma_load_2d(sA, &p.tmA, bar, 0, 0);
    tma_load_2d(sB, &p.tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_f16_mn(p.fmt_a, p.fmt_b, 128, p.n);
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint64_t da = make_smem_desc_mn(smem_u32(sA) + p.a_start + ks * p.a_kstep, p.a_lbo, p.a_sbo, p.a_swz);
      const uint64_t db = make_smem_desc_mn(smem_u32(sB) + p.b_start + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.b_swz);
      umma_f16(tmem_acc, da, db, idesc, ks != 0);
    }
    umma_commit(acc_bar);
  }
  __syncwarp();
  mbar_wait(acc_bar, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < p.n; c += 16) {
    float v[16];
    tmem_ld_x16(taddr + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[(size_t)(warp * 32 + lane) * p.n + c + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, 256);
  }
}

// Throughput microbenchmark: `iters` back-to-back tcgen05.mma (M = 128) with the given operand layout, alternating
// between `nacc` accumulators and advancing K by 32 B per instruction like a real k-loop (4 steps, then wrap).
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int n, int kc, int swizzle, int a_start_bytes, int a_sbo,
                                                            int iters, int nacc, long long* cycles_out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;               // 64 KiB of (zero) operand data
  unsigned char* sB = smem + 64 * 1024;   // 32 KiB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(kFmtBF16, kFmtBF16, 128, n);
    const int row_bytes = kc * 2;
    const int ksteps = kc / 16;
    const uint64_t da0 = make_smem_desc(smem_u32(sA) + a_start_bytes, a_sbo, swizzle, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(sB), 8 * row_bytes, swizzle, 0);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        const int ks = i % ksteps;
        umma_f16(tmem_acc + (i % nacc) * n, da0 + 2 * ks, db0 + 2 * ks, idesc, 1);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, 512);
  }
}

// Second throughput microbenchmark: the issue loop is unrolled 8x with compile-time k-step / accumulator selection,
// so that the single issuing thread spends ~2 instructions per tcgen05.mma (the first version's runtime modulos cost
// tens of cycles per iteration and hid every MMA shorter than that).
__global__ void __launch_bounds__(128, 1) umma_bench2_kernel(int n, int kc, int swizzle, int a_start_bytes, int a_sbo,
                                                             int iters, int nacc, long long* cycles_out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(kFmtBF16, kFmtBF16, 128, n);
    const int row_bytes = kc * 2;
    const uint32_t kmask = kc / 16 - 1;
    const uint64_t da0 = make_smem_desc(smem_u32(sA) + a_start_bytes, a_sbo, swizzle, 0);
    const uint64_t db0 = make_smem_desc(smem_u32(sB), 8 * row_bytes, swizzle, 0);
    uint64_t da[8], db[8];
    uint32_t acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      da[u] = da0 + 2 * (u & kmask);
      db[u] = db0 + 2 * (u & kmask);
      acc[u] = tmem_acc + ((u >> 2) & (nacc - 1)) * n;
    }
    const long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 1
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) umma_f16(acc[u], da[u], db[u], idesc, 1);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, 512);
  }
}

// Third microbenchmark: the conv kernel's own steady-state issue code (conv_issue.cuh) for one halo chunk per "item"
// (9 taps x MT m-tiles x KSTEPS k-steps), alternating between two A stages and two accumulator stages, one commit per
// item -- the MMA issuer of conv_igemm_kernel with everything else removed.  MODE 0: running descriptors, 1: per-tap
// re-derived descriptors (issue_tap).
template <int MT, int BN, int KSTEPS, int MODE>
__global__ void __launch_bounds__(128, 1) umma_bench3_kernel(int iters, int extras, long long* cycles_out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  constexpr int kc = 16 * KSTEPS;
  constexpr uint32_t a_stage = (((16 * MT + 2) * kHaloPitch * kc * 2) + 1023) & ~1023;
  constexpr uint32_t b_stage = ((BN * kc * 2) + 1023) & ~1023;
  unsigned char* sA = smem;
  constexpr int kBTiles = (9 * b_stage + 2 * a_stage <= 190 * 1024) ? 9 : 1;   // large N: every tap reuses one weight tile
  unsigned char* sB = smem + 2 * a_stage;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBTiles * b_stage);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (int)((2 * a_stage + kBTiles * b_stage) / 4); i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0) {
    constexpr uint32_t kLbo = 1u << 16;
    const uint32_t swz = kc == 64 ? kSwizzle128B : kSwizzle64B;
    SegMma g;
    g.a_hi = static_cast<uint32_t>(make_smem_desc(0, kHaloPitch * kc * 2, swz) >> 32);
    g.b_hi = static_cast<uint32_t>(make_smem_desc(0, 8 * kc * 2, swz) >> 32);
    g.idesc = make_idesc_f16(kFmtBF16, kFmtBF16, 128, BN);
    const uint32_t a_base16 = smem_u32(sA) >> 4, a_stage16 = a_stage >> 4;
    const uint32_t b_lo = (smem_u32(sB) >> 4) | kLbo, b_stage16 = kBTiles == 9 ? (b_stage >> 4) : 0u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_lo = (a_base16 + (i & 1) * a_stage16) | kLbo;
      const uint32_t acc = tmem_acc + (i & 1) * (MT * BN);
      if (extras & 2) tc_fence_after_sync();
      if (extras & 8) mbar_wait(&bar[3], 1);          // a wait that succeeds immediately (fresh barrier, parity 1)
      if (extras & 2) tc_fence_after_sync();
      if (elect_one()) {
        if (MODE == 0) {
          issue_halo_chunk_running<MT, BN, KSTEPS>(acc, a_lo, b_lo, b_stage16, g, 0u);
        } else {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp)
            issue_tap<MT, BN, KSTEPS>(acc, a_lo + ((tp / 3) * kHaloPitch + tp % 3) * (2 * KSTEPS), 16 * kHaloPitch * 2 * KSTEPS,
                                      b_lo + tp * b_stage16, g, (tp == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
      if (elect_one()) umma_commit(&bar[i & 1]);
      __syncwarp();
      if (extras & 4) {
        if (elect_one()) umma_commit(&bar[i & 1]);
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&bar[2]);
    __syncwarp();
    mbar_wait(&bar[2], 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, 512);
  }
}

template <int MT, int BN, int KSTEPS, int MODE>
int launch_bench3(int iters, int extras, int grid, long long* cycles_out, cudaStream_t stream) {
  const size_t smem = 200 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_bench3_kernel<MT, BN, KSTEPS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "umma_bench3 smem attribute");
  umma_bench3_kernel<MT, BN, KSTEPS, MODE><<<grid, 128, smem, stream>>>(iters, extras, cycles_out);
  return set_cuda_error(cudaGetLastError(), "umma_bench3 launch");
}

}  // namespace
}  // namespace lass
using namespace lass;

extern "C" int lass_debug_umma_bench(int n, int kc, int swizzle_mode, int a_start_bytes, int a_sbo, int iters, int nacc,
                                     int grid, long long* cycles_out, void* stream) {
  if (!cycles_out || n % 16 || n < 16 || n > 256 || (kc != 32 && kc != 64) || nacc < 1 || nacc * n > 512 || grid < 1)
    return set_error(LASS_ERR_ARG, "umma_bench: bad arguments");
  const size_t smem = 1024 + 96 * 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "umma_bench smem attribute");
  umma_bench_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(n, kc, swizzle_mode, a_start_bytes, a_sbo, iters, nacc, cycles_out);
  return set_cuda_error(cudaGetLastError(), "umma_bench launch");
}

// Unrolled-issue variant (see umma_bench2_kernel); nacc must be 1 or 2, iters a multiple of 8.
extern "C" int lass_debug_umma_bench2(int n, int kc, int swizzle_mode, int a_start_bytes, int a_sbo, int iters, int nacc,
                                      int grid, long long* cycles_out, void* stream) {
  if (!cycles_out || n % 16 || n < 16 || n > 256 || (kc != 32 && kc != 64) || (nacc != 1 && nacc != 2) || nacc * n > 512 ||
      grid < 1 || iters % 8)
    return set_error(LASS_ERR_ARG, "umma_bench2: bad arguments");
  const size_t smem = 1024 + 96 * 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_bench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "umma_bench2 smem attribute");
  umma_bench2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(n, kc, swizzle_mode, a_start_bytes, a_sbo, iters, nacc, cycles_out);
  return set_cuda_error(cudaGetLastError(), "umma_bench2 launch");
}

extern "C" int lass_debug_umma_probe(const void* A, int a_rows, const void* Bm, int n, int kc, int swizzle_mode,
                                              int a_start_bytes, int a_sbo, int a_base_offset, int b_sbo, int fmt_fp16,
                                              float* out, void* stream) {
  if (!A || !Bm || !out) return set_error(LASS_ERR_ARG, "probe: null pointer");
  if (!(kc == 64 || kc == 32) || n % 16 || n < 16 || n > 256 || a_rows < 8 || a_rows > 256 || a_rows % 8)
    return set_error(LASS_ERR_ARG, "probe: bad shape");
  ProbeParams p;
  const CUtensorMapSwizzle sw = swizzle_mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_mode == 4 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_mode == 6 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                    : CU_TENSOR_MAP_SWIZZLE_NONE;
  {
    uint64_t dims[2] = {(uint64_t)kc, (uint64_t)a_rows};
    uint64_t strides[1] = {(uint64_t)kc * 2};
    uint32_t box[2] = {(uint32_t)kc, (uint32_t)a_rows};
    int e = make_tensor_map(&p.tmA, A, 2, 2, dims, strides, box, sw);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)kc, (uint64_t)n};
    uint64_t strides[1] = {(uint64_t)kc * 2};
    uint32_t box[2] = {(uint32_t)kc, (uint32_t)n};
    int e = make_tensor_map(&p.tmB, Bm, 2, 2, dims, strides, box, sw);
    if (e) return e;
  }
  p.out = out;
  p.a_rows = a_rows;
  p.n = n;
  p.kc = kc;
  p.swizzle = swizzle_mode;
  p.a_start_bytes = a_start_bytes;
  p.a_sbo = a_sbo;
  p.a_base_offset = a_base_offset;
  p.b_sbo = b_sbo;
  p.fmt = fmt_fp16 ? kFmtF16 : kFmtBF16;
  const size_t smem = 1024 + 2 * 256 * 128 + 2048 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "probe smem attribute");
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return set_cuda_error(cudaGetLastError(), "probe launch");
}

// Issue-rate benchmark of the conv kernel's steady-state MMA code; returns total cycles per CTA for `iters` items of
// 9 * mt * ksteps MMAs each (M = 128, N = bn, K = 16).  Supported (mt, bn, ksteps): (2,32,2) (2,64,2) (2,32,4) (1,64,4) (1,256,4).
extern "C" int lass_debug_umma_bench3(int mt, int bn, int ksteps, int mode, int iters, int grid, long long* cycles_out, void* stream) {
  if (!cycles_out || iters < 1 || grid < 1) return set_error(LASS_ERR_ARG, "umma_bench3: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
#define LASS_B3(MT_, BN_, KS_)                                                                       \
  if (mt == MT_ && bn == BN_ && ksteps == KS_)                                                       \
    return (mode & 1) == 0 ? launch_bench3<MT_, BN_, KS_, 0>(iters, mode, grid, cycles_out, st)                  \
                           : launch_bench3<MT_, BN_, KS_, 1>(iters, mode, grid, cycles_out, st);
  LASS_B3(2, 32, 2)
  LASS_B3(2, 64, 2)
  LASS_B3(2, 32, 4)
  LASS_B3(1, 64, 4)
  LASS_B3(1, 256, 4)
#undef LASS_B3
  return set_error(LASS_ERR_ARG, "umma_bench3: unsupported configuration");
}

extern "C" int lass_debug_umma_probe_mn(const void* A, int a_rows, int a_swz, const void* Bm, int b_rows, int b_swz, int n, int ksteps,
                                        int a_start, int a_lbo, int a_sbo, int a_kstep, int b_start, int b_lbo, int b_sbo, int b_kstep,
                                        int a_fp16, int b_fp16, float* out, void* stream) {
  if (!A || !Bm || !out) return set_error(LASS_ERR_ARG, "probe_mn: null pointer");
  if ((a_swz != 2 && a_swz != 4) || (b_swz != 2 && b_swz != 4) || n % 16 || n < 16 || n > 256 || a_rows < 8 || a_rows > 256 || b_rows < 8 ||
      b_rows > 256 || ksteps < 1)
    return set_error(LASS_ERR_ARG, "probe_mn: bad shape");
  ProbeMnParams p;
  p.wa = a_swz == 2 ? 64 : 32;
  p.wb = b_swz == 2 ? 64 : 32;
  {
    uint64_t dims[2] = {(uint64_t)p.wa, (uint64_t)a_rows};
    uint64_t strides[1] = {(uint64_t)p.wa * 2};
    uint32_t box[2] = {(uint32_t)p.wa, (uint32_t)a_rows};
    int e = make_tensor_map(&p.tmA, A, 2, 2, dims, strides, box, a_swz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.wb, (uint64_t)b_rows};
    uint64_t strides[1] = {(uint64_t)p.wb * 2};
    uint32_t box[2] = {(uint32_t)p.wb, (uint32_t)b_rows};
    int e = make_tensor_map(&p.tmB, Bm, 2, 2, dims, strides, box, b_swz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (e) return e;
  }
  p.out = out;
  p.a_rows = a_rows;
  p.b_rows = b_rows;
  p.n = n;
  p.ksteps = ksteps;
  p.a_swz = a_swz;
  p.a_start = a_start;
  p.a_lbo = a_lbo;
  p.a_sbo = a_sbo;
  p.a_kstep = a_kstep;
  p.b_swz = b_swz;
  p.b_start = b_start;
  p.b_lbo = b_lbo;
  p.b_sbo = b_sbo;
  p.b_kstep = b_kstep;
  p.fmt_a = a_fp16 ? kFmtF16 : kFmtBF16;
  p.fmt_b = b_fp16 ? kFmtF16 : kFmtBF16;
  const size_t smem = 1024 + 128 * 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "probe_mn smem attribute");
  umma_probe_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return set_cuda_error(cudaGetLastError(), "probe_mn launch");
}
