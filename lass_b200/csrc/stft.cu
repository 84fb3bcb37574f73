// K1 `stft_fwd`: reflect-pad + framing + Hann window + DFT as ONE tcgen05 GEMM fed by TMA, with the
// magnitude / cos / sin epilogue fused in.
//
// Replaces (reference, per forward): torchlibrosa 0.1.0 `STFT.forward` (F.pad reflect, two
// conv1d(1 -> n_fft/2+1, k = n_fft, stride = hop)) and `Base.spectrogram_phase` models/base.py:83-88
// (clamp(re^2+im^2, 1e-10)**0.5, re/mag, im/mag) — SURVEY.md §8a rows a3, a4.
//
// GEMM view per clip:  C[t, j] = sum_k frames[t, k] * basis[j, k]
//   A = frames (T x n_fft): never materialised.  `stft_prep` writes the reflect-padded waveform once as
//       two bf16 arrays (hi, lo = x - hi); a 3-D TMA tensor map with OVERLAPPING rows
//       (dim0 = k stride 1, dim1 = t stride hop, dim2 = clip) fetches 128-frame x 64-sample K-major tiles.
//   B = windowed DFT basis (the reference's frozen conv_real / conv_imag weights), split hi/lo bf16,
//       row-interleaved per 64-bin tile: rows [0,64) = real basis, rows [64,128) = imag basis of the same
//       bins, so one thread of the epilogue holds re and im of a bin.
//   fp32-parity mode (split = 1): hi*hi + hi*lo + lo*hi, three bf16 MMAs into one fp32 TMEM accumulator
//       (max rel. error ~5e-6); fast mode (split = 0): hi*hi only.
//
// CTA = one 128-frame x 64-bin output tile: warp 0 TMA producer, warp 1 TMEM alloc + MMA issue,
// warps 2..5 epilogue (TMEM -> registers -> smem transpose -> coalesced fp32 stores).
#include "lass_internal.cuh"
#include "ptx.cuh"

namespace lass {

namespace {

constexpr int BM = 128;   // frames per tile
constexpr int BN = 128;   // 64 bins x (re, im)
constexpr int BK = 64;    // samples per pipeline stage (128 B rows, SWIZZLE_128B)
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 2;  // 16 KiB (A and B tiles have the same size)
constexpr int kStageBytes = 4 * kTileBytes;
constexpr int kThreads = 192;
constexpr int kStagePad = 65;  // epilogue transpose row pitch (floats)

struct StftParams {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  float* mag;
  float* cosp;
  float* sinp;
  int T, F, n_fft, split;
  int magphase_mode;  // 0: Base.spectrogram_phase (clamp(re^2+im^2, 1e-10)**0.5, re/mag); 1: torchlibrosa.magphase
};

__global__ void __launch_bounds__(kThreads, 1) stft_gemm_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024 B alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m0 = blockIdx.y * BM;
  const int b = blockIdx.z;
  const int KB = p.n_fft / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA_hi);
    tma_prefetch_desc(&p.tmB_hi);
    if (p.split) {
      tma_prefetch_desc(&p.tmA_lo);
      tma_prefetch_desc(&p.tmB_lo);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t stage_tx = p.split ? 4 * kTileBytes : 2 * kTileBytes;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st = smem + s * kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_tx);
        tma_load_3d(st, &p.tmA_hi, &full_bar[s], kb * BK, m0, b);
        tma_load_2d(st + 2 * kTileBytes, &p.tmB_hi, &full_bar[s], kb * BK, n_tile * BN);
        if (p.split) {
          tma_load_3d(st + kTileBytes, &p.tmA_lo, &full_bar[s], kb * BK, m0, b);
          tma_load_2d(st + 3 * kTileBytes, &p.tmB_lo, &full_bar[s], kb * BK, n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kFmtBF16, kFmtBF16, BM, BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(smem + s * kStageBytes);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t a_hi = make_smem_desc(st + ks * 32, 1024, kSwizzle128B);
          const uint64_t b_hi = make_smem_desc(st + 2 * kTileBytes + ks * 32, 1024, kSwizzle128B);
          umma_f16(tmem_acc, a_hi, b_hi, idesc, (kb | ks) != 0);
          if (p.split) {
            const uint64_t a_lo = make_smem_desc(st + kTileBytes + ks * 32, 1024, kSwizzle128B);
            const uint64_t b_lo = make_smem_desc(st + 3 * kTileBytes + ks * 32, 1024, kSwizzle128B);
            umma_f16(tmem_acc, a_hi, b_lo, idesc, 1);
            umma_f16(tmem_acc, a_lo, b_hi, idesc, 1);
          }
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    // ---- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) ----
    const int q = warp & 3;
    mbar_wait(acc_bar, 0);
    tc_fence_after_sync();
    // all MMAs (and therefore all TMA loads) are done: the pipeline stages are free for the transpose
    float* stage_f = reinterpret_cast<float*>(smem) + (size_t)q * 3 * 32 * kStagePad;
    float* s_mag = stage_f;
    float* s_cos = stage_f + 32 * kStagePad;
    float* s_sin = stage_f + 2 * 32 * kStagePad;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      float re[16], im[16];
      tmem_ld_x16(taddr + c, re);
      tmem_ld_x16(taddr + 64 + c, im);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p2 = re[j] * re[j] + im[j] * im[j];
        float m, d;
        if (p.magphase_mode == 0) {
          // models/base.py:85-87: mag = clamp(re^2 + im^2, 1e-10) ** 0.5 ; cos = re / mag ; sin = im / mag
          m = sqrtf(fmaxf(p2, 1e-10f));
          d = m;
        } else {
          // torchlibrosa.stft.magphase: mag = (re^2 + im^2) ** 0.5 ; cos = re / clamp(mag, 1e-10) ; sin likewise
          m = sqrtf(p2);
          d = fmaxf(m, 1e-10f);
        }
        s_mag[lane * kStagePad + c + j] = m;
        s_cos[lane * kStagePad + c + j] = re[j] / d;
        s_sin[lane * kStagePad + c + j] = im[j] / d;
      }
    }
    __syncwarp();
    const int f0 = n_tile * 64;
    for (int r = 0; r < 32; ++r) {
      const int t = m0 + q * 32 + r;
      if (t >= p.T) break;
      const size_t row = ((size_t)b * p.T + t) * p.F;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int f = f0 + h * 32 + lane;
        if (f < p.F) {
          p.mag[row + f] = s_mag[r * kStagePad + h * 32 + lane];
          p.cosp[row + f] = s_cos[r * kStagePad + h * 32 + lane];
          p.sinp[row + f] = s_sin[r * kStagePad + h * 32 + lane];
        }
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_acc, BN);
  }
}

// reflect-pad (center = True, pad_mode = 'reflect') and split to bf16 hi / lo
__global__ void stft_prep_kernel(const float* __restrict__ wave, __nv_bfloat16* __restrict__ xhi,
                                 __nv_bfloat16* __restrict__ xlo, int L, int Lp, int half) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Lp) return;
  float v = 0.0f;
  int src = i - half;
  if (i < L + 2 * half) {
    if (src < 0) src = -src;
    if (src >= L) src = 2 * (L - 1) - src;
    v = wave[(size_t)b * L + src];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  xhi[(size_t)b * Lp + i] = hi;
  xlo[(size_t)b * Lp + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

}  // namespace

int stft_num_ntiles(int n_fft) { return (n_fft / 2 + 1 + 63) / 64; }

// Row pitch of the padded waveform.  cuTensorMapEncodeTiled wants every stride to be a multiple of 16 B and
// of the preceding stride, so the clip pitch is rounded up to a multiple of hop (hop % 8 == 0).
size_t stft_padded_len(int L, int n_fft, int hop) { return (((size_t)L + n_fft + hop - 1) / hop) * hop; }

size_t stft_workspace_bytes(int B, int L, int n_fft, int hop) {
  return 2 * (size_t)B * stft_padded_len(L, n_fft, hop) * 2 + 256;
}

int launch_stft(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                float* mag, float* cosp, float* sinp, int precision_mode, int magphase_mode, void* workspace,
                cudaStream_t stream) {
  if (n_fft % BK != 0 || hop % 8 != 0 || L <= n_fft / 2) return LASS_ERR_ARG;
  const int T = L / hop + 1;
  const int F = n_fft / 2 + 1;
  const size_t Lp = stft_padded_len(L, n_fft, hop);
  __nv_bfloat16* xhi = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* xlo = xhi + (size_t)B * Lp;
  {
    dim3 grid((unsigned)((Lp + 255) / 256), (unsigned)B);
    stft_prep_kernel<<<grid, 256, 0, stream>>>(wave, xhi, xlo, L, (int)Lp, n_fft / 2);
  }
  StftParams p;
  const int ntn = stft_num_ntiles(n_fft);
  {
    // frames: dim0 = sample within frame, dim1 = frame (stride hop: overlapping rows), dim2 = clip
    uint64_t dims[3] = {(uint64_t)n_fft, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)hop * 2, (uint64_t)Lp * 2};
    uint32_t box[3] = {BK, BM, 1};
    int e = make_tensor_map(&p.tmA_hi, xhi, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    e = make_tensor_map(&p.tmA_lo, xlo, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)n_fft, (uint64_t)ntn * BN};
    uint64_t strides[1] = {(uint64_t)n_fft * 2};
    uint32_t box[2] = {BK, BN};
    int e = make_tensor_map(&p.tmB_hi, basis_hi, 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    e = make_tensor_map(&p.tmB_lo, basis_lo, 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  p.mag = mag;
  p.cosp = cosp;
  p.sinp = sinp;
  p.T = T;
  p.F = F;
  p.n_fft = n_fft;
  p.split = precision_mode == 0 ? 1 : 0;
  p.magphase_mode = magphase_mode ? 1 : 0;
  const size_t smem = (size_t)kStages * kStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stft_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "stft smem attribute");
    configured = true;
  }
  dim3 grid((unsigned)ntn, (unsigned)((T + BM - 1) / BM), (unsigned)B);
  stft_gemm_kernel<<<grid, kThreads, smem, stream>>>(p);
  return set_cuda_error(cudaGetLastError(), "stft launch");
}

}  // namespace lass
