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
//       row-interleaved per 128-bin tile: rows [0,128) = real basis, rows [128,256) = imag basis of the same
//       bins, so one thread of the epilogue holds re and im of a bin.  Im X[0] is identically zero; its row
//       carries the real basis of the Nyquist bin n_fft/2 instead (whose imaginary part is zero as well), so
//       n_fft/2 + 1 bins fit n_fft/256 tiles exactly.
//   fp32-parity mode (split = 1): hi*hi + hi*lo + lo*hi, three bf16 MMAs into one fp32 TMEM accumulator
//       (max rel. error ~5e-6); fast mode (split = 0): hi*hi only.
//
// Persistent CTAs, item = 128-frame x 128-bin output tile (M = 128, N = 256, K = n_fft): warp 0 TMA producer,
// warp 1 TMEM alloc + MMA issue, warps 2..9 epilogue (TMEM -> registers -> smem transpose -> coalesced fp32 stores),
// two TMEM accumulator stages.  Operand traffic per MMA cycle is 1.4x lower than with 128 x 128 tiles (the kernel is
// bound by the L2 -> shared-memory rate of the TMA loads, not by the tensor pipe).
#include "lass_internal.cuh"
#include "ptx.cuh"

namespace lass {

namespace {

constexpr int BM = 128;   // frames per tile
constexpr int BN = 256;   // 128 bins x (re, im)
constexpr int BK = 64;    // samples per k-stage (128 B rows, SWIZZLE_128B)
constexpr int kBinsPerTile = BN / 2;
// Shared-memory rings.  The kernel always runs as CTA pairs (cta_group::2, M = 256 = two neighbouring frame tiles): each CTA
// holds its own 128 frames and HALF of the basis tile (128 of the 256 rows).
//   A ring: "base tiles" of up to kARowsMax frame rows x 64 samples (hi | lo).  Consecutive frames start `hop` samples apart,
//           so k-stage kb + P of frame m holds the SAME samples as k-stage kb of frame m + S whenever S * hop = P * 64 (hop 160:
//           P = 5, S = 2; hop 320: P = 5, S = 1): one base tile of 128 + S (uses - 1) rows serves every k-stage kb = r, r + P,
//           r + 2P, ... through descriptors whose start address is shifted by whole rows (the swizzle is a function of the
//           absolute address bits, as for the conv kernel's taps).  The frames are fetched n_fft / (64 P) times less often.
//   B ring: one k-stage of the basis rows per stage (hi | lo), streamed.
constexpr int kARowsMax = 144;
constexpr int kABase = kARowsMax * BK * 2;   // 18 KiB
constexpr int kAStage = 2 * kABase;
constexpr int kAStages = 2;
constexpr int kBTile = (BN / 2) * BK * 2;    // 16 KiB
constexpr int kBStage = 2 * kBTile;
constexpr int kBStages = 3;
constexpr int kEpiWarps = 8;              // two per TMEM lane quarter, each takes half of the tile's bins
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTrPitch = 33;              // epilogue transpose row pitch (floats)
constexpr int kTrFloats = 32 * kTrPitch;  // per epilogue warp

// One launch serves up to kMaxProblems STFTs of the same batch and hop (the multi-resolution front end of BASELINE config 5:
// n_fft 2048 / 512 / 256 at hop 160): the item list is the concatenation of the problems' tiles, longest K first.
constexpr int kMaxProblems = 3;
struct StftProblem {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  float* mag;
  float* cosp;
  float* sinp;
  int F, n_fft, n_tiles, item_end;   // item_end: one past this problem's last item in the launch's item list
  int period, shift_rows, a_rows;    // A base-tile re-use: k-stages r, r + period, ... share one tile, shifted by shift_rows rows
};
struct StftParams {
  StftProblem pr[kMaxProblems];
  int nprob;
  int T, split;
  int magphase_mode;  // 0: Base.spectrogram_phase (clamp(re^2+im^2, 1e-10)**0.5, re/mag); 1: torchlibrosa.magphase
  int m_tiles, num_items;
};

// item -> (problem, item within the problem)
__device__ __forceinline__ const StftProblem& find_problem(const StftParams& p, int& item) {
  int g = 0;
  while (g + 1 < p.nprob && item >= p.pr[g].item_end) ++g;
  if (g > 0) item -= p.pr[g - 1].item_end;
  return p.pr[g];
}

// magnitude and unit phasor of one bin.  One MUFU.RSQ (2 ulp) instead of sqrt + two divisions: errors of a few 1e-7
// relative, against the 1e-4 parity bar (tests/test_gpu_spectral.py).
__device__ __forceinline__ void magphase(float re, float im, int mode, float& m, float& c, float& sn) {
  if (mode == 2) {      // raw real / imaginary parts (the ISTFT adjoint of the training step, lass_istft_bwd)
    m = re;
    c = 0.0f;
    sn = im;
    return;
  }
  const float p2 = re * re + im * im;
  // mode 0, models/base.py:85-87: mag = clamp(re^2 + im^2, 1e-10) ** 0.5 ; cos = re / mag ; sin = im / mag
  // mode 1, torchlibrosa.stft.magphase: mag = (re^2 + im^2) ** 0.5 ; cos = re / clamp(mag, 1e-10) ; sin likewise
  const float q = fmaxf(p2, mode == 0 ? 1e-10f : 1e-20f);
  const float r = rsqrtf(q);
  m = (mode == 0 ? q : p2) * r;
  c = re * r;
  sn = im * r;
}

// Persistent CTA PAIRS (thread-block cluster of 2) over (clip, pair of 128-frame tiles, 128-bin tile) items, bin tile fastest
// so that the pairs running concurrently share the frames in L2.  Warp 0 TMA producer, warp 1 TMEM alloc + MMA issue (two
// 256-column accumulator stages: the epilogue of item i overlaps the MMAs of item i + 1), warps 2..9 epilogue.
// The leader CTA (rank 0) issues every MMA for both (M = 256: rows 0-127 from its shared memory into its TMEM, rows 128-255
// from / into the peer's); both CTAs load their own frames and their half of the basis rows (transaction bytes counted on the
// leader's barriers), tcgen05.commit multicasts the stage-free / accumulator-full arrivals to both, the peer's epilogue warps
// arrive remotely on the leader's acc_empty.
__global__ void __launch_bounds__(kThreads, 1) stft_gemm_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024 B alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* a_ring = smem;
  unsigned char* b_ring = smem + kAStages * kAStage;
  float* tr_base = reinterpret_cast<float*>(b_ring + kBStages * kBStage);                  // [kEpiWarps][32][33]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tr_base + kEpiWarps * kTrFloats);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* b_full = a_empty + kAStages;
  uint64_t* b_empty = b_full + kBStages;
  uint64_t* acc_full = b_empty + kBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int vbid = (int)(blockIdx.x >> 1), vgrid = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.nprob; ++g) {
      tma_prefetch_desc(&p.pr[g].tmA_hi);
      tma_prefetch_desc(&p.pr[g].tmB_hi);
      if (p.split) {
        tma_prefetch_desc(&p.pr[g].tmA_lo);
        tma_prefetch_desc(&p.pr[g].tmB_lo);
      }
    }
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps * 2);     // the epilogue warps of both CTAs arrive on the leader's
    }
    fence_mbar_init();
  }
  cluster_sync_all();               // both CTAs' barriers exist before any remote arrive / transaction lands on them
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 2 * BN);
    tmem_relinquish_cg2();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t nsplit = p.split ? 2u : 1u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (int gitem = vbid; gitem < p.num_items; gitem += vgrid) {
        int item = gitem;
        const StftProblem& q = find_problem(p, item);
        const int KB = q.n_fft / BK;
        const int n_tile = item % q.n_tiles;
        const int rest = item / q.n_tiles;
        const int m0 = ((rest % p.m_tiles) * 2 + (int)crank) * BM, b = rest / p.m_tiles;
        const int brow = n_tile * BN + (int)crank * (BN / 2);          // this CTA's half of the basis rows
        const uint32_t a_tx = 2u * nsplit * (uint32_t)q.a_rows * (BK * 2);   // both CTAs' boxes land on the leader's barrier
        const int nres = q.period < KB ? q.period : KB;
        for (int r = 0; r < nres; ++r) {
          {
            const uint32_t sa = a_it % kAStages;
            mbar_wait(&a_empty[sa], ((a_it / kAStages) & 1) ^ 1);
            unsigned char* st = a_ring + sa * kAStage;
            if (crank == 0) mbar_arrive_expect_tx(&a_full[sa], a_tx);
            const uint32_t bar = mapa_shared(smem_u32(&a_full[sa]), 0);
            tma_load_3d_cg2(st, &q.tmA_hi, bar, r * BK, m0, b);
            if (p.split) tma_load_3d_cg2(st + kABase, &q.tmA_lo, bar, r * BK, m0, b);
            ++a_it;
          }
          for (int kb = r; kb < KB; kb += q.period, ++b_it) {
            const uint32_t sb = b_it % kBStages;
            mbar_wait(&b_empty[sb], ((b_it / kBStages) & 1) ^ 1);
            unsigned char* st = b_ring + sb * kBStage;
            if (crank == 0) mbar_arrive_expect_tx(&b_full[sb], 2u * nsplit * kBTile);
            const uint32_t bar = mapa_shared(smem_u32(&b_full[sb]), 0);
            tma_load_2d_cg2(st, &q.tmB_hi, bar, kb * BK, brow);
            if (p.split) tma_load_2d_cg2(st + kBTile, &q.tmB_lo, bar, kb * BK, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = make_idesc_f16(kFmtBF16, kFmtBF16, 2 * BM, BN);
      uint32_t a_it = 0, b_it = 0, n = 0;
      for (int gitem = vbid; gitem < p.num_items; gitem += vgrid, ++n) {
        int item = gitem;
        const StftProblem& q = find_problem(p, item);
        const int KB = q.n_fft / BK;
        const int nres = q.period < KB ? q.period : KB;
        const uint32_t as = n & 1u;
        mbar_wait(&acc_empty[as], ((n >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + as * BN;
        uint32_t accumulate = 0;
        for (int r = 0; r < nres; ++r, ++a_it) {
          const uint32_t sa = a_it % kAStages;
          mbar_wait(&a_full[sa], (a_it / kAStages) & 1);
          tc_fence_after_sync();
          uint32_t a_addr = smem_u32(a_ring + sa * kAStage);
          for (int kb = r; kb < KB; kb += q.period, ++b_it, a_addr += (uint32_t)q.shift_rows * (BK * 2)) {
            const uint32_t sb = b_it % kBStages;
            mbar_wait(&b_full[sb], (b_it / kBStages) & 1);
            tc_fence_after_sync();
            const uint32_t bst = smem_u32(b_ring + sb * kBStage);
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              const uint64_t a_hi = make_smem_desc(a_addr + ks * 32, 1024, kSwizzle128B);
              const uint64_t b_hi = make_smem_desc(bst + ks * 32, 1024, kSwizzle128B);
              umma_f16_cg2(tmem_acc, a_hi, b_hi, idesc, accumulate);
              accumulate = 1;
              if (p.split) {
                const uint64_t a_lo = make_smem_desc(a_addr + kABase + ks * 32, 1024, kSwizzle128B);
                const uint64_t b_lo = make_smem_desc(bst + kBTile + ks * 32, 1024, kSwizzle128B);
                umma_f16_cg2(tmem_acc, a_hi, b_lo, idesc, 1);
                umma_f16_cg2(tmem_acc, a_lo, b_hi, idesc, 1);
              }
            }
            umma_commit_cg2(&b_empty[sb]);   // frees the basis stage in both CTAs once these MMAs have read it
          }
          umma_commit_cg2(&a_empty[sa]);     // ... and the frame tile after its last k-stage
        }
        umma_commit_cg2(&acc_full[as]);      // accumulator complete (published in both CTAs)
      }
    }
  } else {
    // ---- epilogue: warps 2..9; TMEM lane quarter = warp % 4, bin half = (warp - 2) / 4 ----
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    float* tr = tr_base + (warp - 2) * kTrFloats;
    uint32_t n = 0;
    for (int gitem = vbid; gitem < p.num_items; gitem += vgrid, ++n) {
      int item = gitem;
      const StftProblem& pq = find_problem(p, item);
      const int half = pq.n_fft / 2;
      const int n_tile = item % pq.n_tiles;
      const int rest = item / pq.n_tiles;
      const int m0 = ((rest % p.m_tiles) * 2 + (int)crank) * BM, b = rest / p.m_tiles;
      float* const g_mag = pq.mag;
      float* const g_cos = pq.cosp;
      float* const g_sin = pq.sinp;
      const int F = pq.F;
      const uint32_t as = n & 1u;
      mbar_wait(&acc_full[as], (n >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
      const int t_row0 = m0 + q * 32;                       // first frame of this warp's 32 rows
      const int t_mine = t_row0 + lane;
      const int rows = min(32, p.T - t_row0);               // valid frames among them (<= 0: none)
      const size_t row0_off = ((size_t)b * p.T + t_row0) * F;
#pragma unroll 1
      for (int cc = 0; cc < kBinsPerTile / 2; cc += 32) {
        const int c = hsel * (kBinsPerTile / 2) + cc;
        float re[32], im[32], cs[32];
        tmem_ld_x32(taddr + c, re);
        tmem_ld_x32(taddr + kBinsPerTile + c, im);
        tmem_ld_wait();
        const int f0 = n_tile * kBinsPerTile + c;
        if (n_tile == 0 && c == 0) {
          // bin 0 has no imaginary part; its slot in the basis carries the (purely real) Nyquist bin n_fft/2
          float m, c1, s1;
          magphase(im[0], 0.0f, p.magphase_mode, m, c1, s1);
          if (t_mine < p.T) {
            const size_t o = row0_off + (size_t)lane * F + half;
            g_mag[o] = m;
            if (g_cos) g_cos[o] = c1;
            g_sin[o] = s1;
          }
          im[0] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) magphase(re[j], im[j], p.magphase_mode, re[j], cs[j], im[j]);   // re <- mag, im <- sin
        // three transposes through the warp's 32 x 33 buffer: lane = frame  ->  lane = bin, 128 B rows to global memory
#pragma unroll
        for (int which = 0; which < 3; ++which) {
          if (which == 1 && !g_cos) continue;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * kTrPitch + j] = which == 0 ? re[j] : (which == 1 ? cs[j] : im[j]);
          __syncwarp();
          float* dst = (which == 0 ? g_mag : (which == 1 ? g_cos : g_sin)) + row0_off + f0 + lane;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            if (r < rows) dst[(size_t)r * F] = tr[r * kTrPitch + lane];
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[as]), 0));    // the leader's barrier
        else mbar_arrive(&acc_empty[as]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // neither CTA leaves while the other may still touch its shared memory / barriers / TMEM
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_cg2(tmem_base, 2 * BN);
  }
}

// reflect-pad (center = True, pad_mode = 'reflect') and split to bf16 hi / lo; grid.z = problem of a multi-resolution launch
struct PrepParams {
  __nv_bfloat16* xhi[kMaxProblems];
  __nv_bfloat16* xlo[kMaxProblems];
  int Lp[kMaxProblems], half[kMaxProblems];
};
// Reflect padding + bf16 hi / lo split of the waveform; a thread owns 8 consecutive padded samples (Lp, n_fft / 2 are multiples
// of 8): two 16-byte loads in the interior (clip length a multiple of 4), one 16-byte store per output array.
__global__ void __launch_bounds__(256) stft_prep_kernel(const float* __restrict__ wave, const PrepParams pp, int L, int vec_ok) {
  const int g = blockIdx.z;
  const int b = blockIdx.y;
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const int Lp = pp.Lp[g], half = pp.half[g];
  if (i0 >= Lp) return;
  const float* wb = wave + (size_t)b * L;
  float v[8];
  const int s0 = i0 - half;
  if (vec_ok && s0 >= 0 && s0 + 8 <= L) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wb + s0)), c = __ldg(reinterpret_cast<const float4*>(wb + s0) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j;
      v[j] = 0.0f;
      if (i < L + 2 * half) {
        int src = i - half;
        if (src < 0) src = -src;
        if (src >= L) src = 2 * (L - 1) - src;
        v[j] = __ldg(wb + src);
      }
    }
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - __bfloat162float(h0), v[2 * j + 1] - __bfloat162float(h1));
    hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(pp.xhi[g] + (size_t)b * Lp + i0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(pp.xlo[g] + (size_t)b * Lp + i0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Adjoint of torchlibrosa ISTFT.forward's overlap-add: the waveform gradient, zero-extended to the frame grid and divided by
// the window-sum (clamped at 1e-11 like the forward), split to bf16 hi / lo.  Sample i of the padded signal is position
// i of y_full (frame t covers [t*hop, t*hop + n_fft)); the forward returned y_full[half : half + L] / wsum.
__global__ void istft_bwd_prep_kernel(const float* __restrict__ dwave, const float* __restrict__ window,
                                      __nv_bfloat16* __restrict__ xhi, __nv_bfloat16* __restrict__ xlo, int L, int Lp, int half,
                                      int n_fft, int hop, int T) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Lp) return;
  float v = 0.0f;
  const int src = i - half;
  if (src >= 0 && src < L) {
    int ta = (i - n_fft) / hop + 1;
    if (i - n_fft < 0) ta = 0;
    int tz = i / hop;
    if (tz > T - 1) tz = T - 1;
    float ws = 0.0f;
    for (int t = ta; t <= tz; ++t) {
      const float w = __ldg(window + (i - t * hop));
      ws = fmaf(w, w, ws);
    }
    v = dwave[(size_t)b * L + src] / fmaxf(ws, 1e-11f);
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  xhi[(size_t)b * Lp + i] = hi;
  xlo[(size_t)b * Lp + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

}  // namespace

// 128 bins per tile; the Nyquist bin travels in the (empty) imaginary slot of bin 0, so n_fft/2 bins need tiles
int stft_num_ntiles(int n_fft) { return (n_fft / 2 + kBinsPerTile - 1) / kBinsPerTile; }

// Row pitch of the padded waveform.  cuTensorMapEncodeTiled wants every stride to be a multiple of 16 B and
// of the preceding stride, so the clip pitch is rounded up to a multiple of hop (hop % 8 == 0).
size_t stft_padded_len(int L, int n_fft, int hop) { return (((size_t)L + n_fft + hop - 1) / hop) * hop; }

// (+4 KiB: the re-used frame tiles are fetched with the largest row count of any k-stage residue, which can run up to
//  (period - 1) * 64 samples past the last clip's padded signal; those rows are never used by an MMA of a valid frame)
size_t stft_workspace_bytes(int B, int L, int n_fft, int hop) {
  return (2 * (size_t)B * stft_padded_len(L, n_fft, hop) * 2 + 4096 + 255) / 256 * 256;
}

// nres STFTs of the same (B, L) waveform batch at the same hop in ONE stft_gemm launch (+ one prep launch): problems are
// queued longest-K first so that the persistent CTAs' static round-robin ends on the short items.  workspace: the problems'
// padded hi / lo arrays back to back (stft_workspace_bytes each, 256 B aligned).
int launch_stft_multi(const float* wave, int B, int L, int hop, int nres, const int* n_ffts, const void* const* basis_hi,
                      const void* const* basis_lo, float* const* mag, float* const* cosp, float* const* sinp, int precision_mode,
                      int magphase_mode, void* workspace, cudaStream_t stream, const float* adjoint_window) {
  if (nres < 1 || nres > kMaxProblems || hop % 8 != 0) return LASS_ERR_ARG;
  int order[kMaxProblems];
  for (int g = 0; g < nres; ++g) order[g] = g;
  for (int a = 0; a < nres; ++a)
    for (int c = a + 1; c < nres; ++c)
      if (n_ffts[order[c]] > n_ffts[order[a]]) {
        const int t = order[a];
        order[a] = order[c];
        order[c] = t;
      }
  const int T = L / hop + 1;
  const int m_tiles = (T + BM - 1) / BM;
  const int m_units = (m_tiles + 1) / 2;      // CTA pairs over two neighbouring frame tiles (an odd last tile pairs with padding)
  int gcd_hk = hop, gb = BK;
  while (gb) {
    const int t = gcd_hk % gb;
    gcd_hk = gb;
    gb = t;
  }
  StftParams p;
  PrepParams pp;
  size_t max_lp = 0;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  int items = 0;
  for (int g = 0; g < nres; ++g) {
    const int r = order[g];
    const int n_fft = n_ffts[r];
    if (n_fft % BK != 0 || (n_fft / 2) % kBinsPerTile != 0 || L <= n_fft / 2) return LASS_ERR_ARG;
    const size_t Lp = stft_padded_len(L, n_fft, hop);
    __nv_bfloat16* xhi = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* xlo = xhi + (size_t)B * Lp;
    ws += stft_workspace_bytes(B, L, n_fft, hop);
    pp.xhi[g] = xhi;
    pp.xlo[g] = xlo;
    pp.Lp[g] = (int)Lp;
    pp.half[g] = n_fft / 2;
    if (Lp > max_lp) max_lp = Lp;
    StftProblem& q = p.pr[g];
    const int ntn = stft_num_ntiles(n_fft);
    {
      // frame-tile re-use across k-stages: S * hop = P * 64 samples
      const int KB = n_fft / BK;
      int period = hop / gcd_hk, shift = BK / gcd_hk;
      int rows = BM + shift * ((KB + period - 1) / period - 1);
      if (rows > kARowsMax || period >= KB) {
        period = KB;
        shift = 0;
        rows = BM;
      }
      q.period = period;
      q.shift_rows = shift;
      q.a_rows = rows;
    }
    {
      // frames: dim0 = sample within frame, dim1 = frame (stride hop: overlapping rows; the shifted re-use reads up to
      // a_rows - BM rows past the last frame, all inside the padded signal), dim2 = clip
      uint64_t dims[3] = {(uint64_t)n_fft, (uint64_t)(T + q.a_rows - BM), (uint64_t)B};
      uint64_t strides[2] = {(uint64_t)hop * 2, (uint64_t)Lp * 2};
      uint32_t box[3] = {BK, (uint32_t)q.a_rows, 1};
      int e = make_tensor_map(&q.tmA_hi, xhi, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (e) return e;
      e = make_tensor_map(&q.tmA_lo, xlo, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (e) return e;
    }
    {
      uint64_t dims[2] = {(uint64_t)n_fft, (uint64_t)ntn * BN};
      uint64_t strides[1] = {(uint64_t)n_fft * 2};
      uint32_t box[2] = {BK, BN / 2};
      int e = make_tensor_map(&q.tmB_hi, basis_hi[r], 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (e) return e;
      e = make_tensor_map(&q.tmB_lo, basis_lo[r] ? basis_lo[r] : basis_hi[r], 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (e) return e;
    }
    q.mag = mag[r];
    q.cosp = cosp ? cosp[r] : nullptr;
    q.sinp = sinp[r];
    q.F = n_fft / 2 + 1;
    q.n_fft = n_fft;
    q.n_tiles = ntn;
    items += ntn * m_units * B;
    q.item_end = items;
  }
  for (int g = nres; g < kMaxProblems; ++g) p.pr[g] = p.pr[0];
  {
    dim3 grid((unsigned)((max_lp + 255) / 256), (unsigned)B, (unsigned)nres);
    const dim3 grid8((unsigned)((max_lp / 8 + 255) / 256), (unsigned)B, (unsigned)nres);       // stft_prep: 8 samples per thread
    if (adjoint_window) {
      if (nres != 1) return LASS_ERR_ARG;
      istft_bwd_prep_kernel<<<dim3(grid.x, grid.y), 256, 0, stream>>>(wave, adjoint_window, pp.xhi[0], pp.xlo[0], L, pp.Lp[0], pp.half[0],
                                                                     n_ffts[0], hop, T);
    } else {
      stft_prep_kernel<<<grid8, 256, 0, stream>>>(wave, pp, L, (L % 4 == 0 && reinterpret_cast<uintptr_t>(wave) % 16 == 0) ? 1 : 0);
    }
  }
  p.nprob = nres;
  p.T = T;
  p.split = precision_mode == 0 ? 1 : 0;
  p.magphase_mode = magphase_mode;
  p.m_tiles = m_units;
  p.num_items = items;
  const int num_sms = device_sm_count();
  const size_t smem = (size_t)kAStages * kAStage + (size_t)kBStages * kBStage + kEpiWarps * kTrFloats * sizeof(float) + 1024 + 256;
  // the opt-in is per device and cheap: set it on every launch for the CURRENT device (no process-wide cache)
  cudaError_t ea = cudaFuncSetAttribute(stft_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ea != cudaSuccess) return set_cuda_error(ea, "stft smem attribute");
  const int npairs = num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)(items < npairs ? items : npairs), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return set_cuda_error(cudaLaunchKernelEx(&cfg, stft_gemm_kernel, p), "stft launch (CTA pairs)");
}

int launch_stft(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                float* mag, float* cosp, float* sinp, int precision_mode, int magphase_mode, void* workspace,
                cudaStream_t stream, const float* adjoint_window) {
  const void* bh[1] = {basis_hi};
  const void* bl[1] = {basis_lo};
  float* m[1] = {mag};
  float* c[1] = {cosp};
  float* sn[1] = {sinp};
  return launch_stft_multi(wave, B, L, hop, 1, &n_fft, bh, bl, m, c, sn, precision_mode, magphase_mode, workspace, stream,
                           adjoint_window);
}

}  // namespace lass
