// Training step: weight gradients on tcgen05 (bf16 operands, fp32 accumulation in TMEM, TMA-fed).
//
//   dW[tap][co][ci] = sum over pixels p of dY[p][co] * X[p + tap][ci]      (zero padding; tap = ky*3 + kx)
//
// GEMM per kernel row ky:  D[(dx, ci), co] = sum_p X[p + (ky-1, dx-1)][ci] * dY[p][co]  with M = (dx, ci), N = co, K = pixels.
// Both operands are NHWC tiles exactly as TMA delivers them ([pixel][channel], hardware swizzle): MN-major operands of
// tcgen05.mma (the contraction index is the row).  The M dimension is filled with HORIZONTAL TAPS of the same channel chunk:
// the X halo tile (18 x 10 pixels of CIW channels) lies in shared memory one pixel per row, so the tile shifted by one pixel is
// the same memory one row further — an MN-major descriptor whose leading byte offset (distance between swizzle atoms along M)
// is ONE ROW reads [dx = 0 | dx = 1 (| dx = 2 | dx = 3)] as one 128-row operand: 2 taps x 64 channels (SWIZZLE_128B) or
// 4 taps x 32 channels (SWIZZLE_64B; the 4th "tap" is discarded).  With 64-channel chunks the atom distance is also used ACROSS
// kernel rows (taps (ky, 2) and (ky + 1, 0) are 8 halo rows apart), so the nine taps need five MMAs per k-step, not six.  The vertical tap is a start-address shift of 10 rows and
// the 8-pixel rows of the 16 x 8 pixel tile are the descriptor's 8-row K groups with a stride of 10 rows — the descriptor
// rules are pinned on hardware by tests/test_gpu_umma_probe.py::test_mn_major_descriptors.  So a 32-channel layer still
// issues full M = 128 MMAs, and no operand is ever transposed or copied.
//
// NB = 128 (co % 128 == 0, 64-channel ci chunks): an MN-major operand is read from shared memory at ~53 B/clk, so the MMA time
// follows the operand BYTES -- per 128 x 64 x 16 MMA 4 KiB of X + 2 KiB of dY (~115 cycles), per 128 x 128 x 16 MMA 4 + 4 KiB
// (~150 cycles) for twice the work.  Five 128-column accumulators do not fit TMEM's 512 columns, so the five MMA groups of the
// 3 x 3 kernel are split over TWO sets of CTAs (groups 0-2 / 3-4, pixel ranges divided 3 : 2 so both finish together); the dY
// tile arrives as two 64-channel boxes one atom (16 KiB) apart, which the N-side descriptor hops with its leading byte offset.
//
// One CTA = one (CIW-channel chunk of ci, NB-channel tile of co) output tile over a contiguous range of pixel tiles
// (split-K; few weights + many pixels -> many ranges): all 3 x 3 taps accumulate in TMEM (5 or 3 accumulators of NB columns),
// then the epilogue adds the valid rows to dW with fp32 red.global.  Warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#include "lass_internal.cuh"
#include "ptx.cuh"

namespace lass {
namespace {

constexpr int kThreads = 192;
constexpr int kTH = 16, kTW = 8, kHaloH = 18, kHaloW = 10;

struct WgradTcParams {
  CUtensorMap tmX, tmY;
  float* dw;
  int co, ci, taps, x_fp16;
  int tiles_h, tiles_w, num_pix_tiles;
  int n_ci_chunks, n_co_tiles, splits;
  int splits0;       // NB = 128, 3 x 3: CTAs [0, splits0) of an output tile run MMA groups 0-2, the rest groups 3-4 (else == splits)
};

template <int CIW, int NB>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  constexpr int NBOX = NB == 128 ? 64 : NB;                      // channels per dY TMA box (one 128-byte swizzle atom at most)
  constexpr int RA = CIW * 2, RB = NBOX * 2;                     // bytes per pixel row of the X / dY tiles
  constexpr uint32_t SWA = CIW == 64 ? kSwizzle128B : kSwizzle64B, SWB = NBOX == 64 ? kSwizzle128B : kSwizzle64B;
  constexpr int YBOX = kTH * kTW * RB;                           // one dY box (the N = 128 tile is two of them, YBOX apart)
  constexpr int XBYTES = kHaloH * kHaloW * RA, YBYTES = (NB / NBOX) * YBOX;
  constexpr int XALLOC = (XBYTES + 1023) & ~1023;
  constexpr int STAGE = XALLOC + YBYTES;
  constexpr int kStages = NB == 128 ? 3 : 4;
  constexpr int TPM = 128 / CIW;                                 // taps per MMA (2 or 4)
  // MMA groups of a 3x3 kernel.  The nine taps sit at halo-row offsets {0,1,2, 10,11,12, 20,21,22}; an MMA's M atoms are LBO
  // bytes apart, so   CIW = 32: one MMA per kernel row = taps (ky,0..2) + one discarded atom (LBO = 1 row);
  //                   CIW = 64: FIVE MMAs (0,1) (2,10) (11,12) (20,21) (22,-): the pair (2,10) spans two kernel rows with
  //                             LBO = 8 rows -- 5 instead of 6 MMAs per k-step.  Accumulator a, atom h holds tap TPM a + h.
  constexpr int NACC = CIW == 64 ? 5 : 3;
  constexpr int NACC_CTA = NB == 128 ? 3 : NACC;                  // accumulators one CTA holds
  constexpr int ACC_COLS = NACC_CTA * NB;
  constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512);
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * STAGE + 1024);   // (+1024: garbage-tap reads run one row past a tile)
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* conv_bar = empty_bar + kStages;      // fp16 X tile converted to bf16 in place (x_fp16 launches only)
  uint64_t* acc_bar = conv_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int item = blockIdx.x;
  int split = item % p.splits;
  item /= p.splits;
  const int co_tile = item % p.n_co_tiles, ci_chunk = item / p.n_co_tiles;
  const bool one_tap = p.taps == 1;
  // MMA groups [acc0, acc0 + nacc) of this CTA and its share of the pixel tiles
  int acc0 = 0, nacc = one_tap ? 1 : NACC, nsplit = p.splits;
  if (p.splits0 < p.splits) {
    if (split < p.splits0) {
      nacc = 3;
      nsplit = p.splits0;
    } else {
      acc0 = 3;
      nacc = NACC - 3;
      split -= p.splits0;
      nsplit = p.splits - p.splits0;
    }
  }
  const int t_begin = (int)((long long)p.num_pix_tiles * split / nsplit);
  const int t_end = (int)((long long)p.num_pix_tiles * (split + 1) / nsplit);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmY);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], 4);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();     // programmatic dependent launch (lass_internal.cuh): the prologue above may overlap the
  griddep_wait();                  // previous launch's tail; nothing below runs before every earlier launch has completed

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int s = it % kStages;
        mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
        const int tw = tile % p.tiles_w;
        const int r = tile / p.tiles_w;
        const int th = r % p.tiles_h, b = r / p.tiles_h;
        unsigned char* st = smem + s * STAGE;
        mbar_arrive_expect_tx(&full_bar[s], XBYTES + YBYTES);
        tma_load_4d(st, &p.tmX, &full_bar[s], ci_chunk * CIW, tw * kTW - 1, th * kTH - 1, b);
        tma_load_4d(st + XALLOC, &p.tmY, &full_bar[s], co_tile * NB, tw * kTW, th * kTH, b);
        if (NB == 128) tma_load_4d(st + XALLOC + YBOX, &p.tmY, &full_bar[s], co_tile * NB + NBOX, tw * kTW, th * kTH, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16_mn(kFmtBF16, kFmtBF16, 128, NB);
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int s = it % kStages;
        mbar_wait(p.x_fp16 ? &conv_bar[s] : &full_bar[s], (it / kStages) & 1);
        tc_fence_after_sync();
        const uint32_t xs = smem_u32(smem + s * STAGE), ys = xs + XALLOC;
        for (int a = acc0; a < acc0 + nacc; ++a) {
          int row0, lbo_rows = 1;                  // halo row of the first atom, atom distance in rows
          if (one_tap) {
            row0 = kHaloW + 1;                     // the centre tap (the second atom is discarded)
          } else if (CIW == 64) {
            row0 = a == 0 ? 0 : a == 1 ? 2 : a == 2 ? kHaloW + 1 : a == 3 ? 2 * kHaloW : 2 * kHaloW + 2;
            if (a == 1) lbo_rows = kHaloW - 2;
          } else {
            row0 = a * kHaloW;
          }
          const uint32_t acc = tmem_base + (uint32_t)((a - acc0) * NB);
#pragma unroll
          for (int ks = 0; ks < kTH / 2; ++ks) {
            const uint64_t da = make_smem_desc_mn(xs + (uint32_t)((2 * ks * kHaloW + row0) * RA), lbo_rows * RA, kHaloW * RA, SWA);
            const uint64_t db = make_smem_desc_mn(ys + (uint32_t)(ks * 2 * kTW * RB), NB == 128 ? YBOX : 0, kTW * RB, SWB);
            umma_f16(acc, da, db, idesc, (it | (uint32_t)ks) != 0u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(acc_bar);
    }
  } else {
    // ---- fp16 activations: tcgen05 kind::f16 takes ONE 16-bit format for both operands (a mixed fp16 x bf16 descriptor is an
    // illegal instruction on sm_100a), and the gradients need bf16's range, so the X tile is converted in place, element by
    // element (the swizzle does not matter), by the warps that otherwise only wait for the epilogue ----
    if (p.x_fp16) {
      const int ct = threadIdx.x - 64;
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int s = it % kStages;
        mbar_wait(&full_bar[s], (it / kStages) & 1);
        unsigned char* st = smem + s * STAGE;
        for (int off = ct * 16; off < XBYTES; off += 128 * 16) {
          uint4 v = *reinterpret_cast<uint4*>(st + off);
          uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            const __nv_bfloat162 b = __float22bfloat162_rn(f);
            w[j] = *reinterpret_cast<const uint32_t*>(&b);
          }
          *reinterpret_cast<uint4*>(st + off) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[s]);
      }
    }
    // ---- epilogue: TMEM lane = (tap within the MMA) * CIW + channel; warp w reads lane quarter w % 4 ----
    mbar_wait(acc_bar, 0);
    tc_fence_after_sync();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int h = row / CIW, cl = row % CIW;       // atom of the MMA's M dimension, channel within the chunk
    const int ci = ci_chunk * CIW + cl;
    if (t_begin < t_end) {
      for (int a = acc0; a < acc0 + nacc; ++a) {
        // tap held by (MMA group a, atom h); CIW = 32: kernel row a, dx = h (h = 3 is the discarded atom)
        const int tap = one_tap ? (h == 0 ? 0 : 9) : (CIW == 64 ? TPM * a + h : (h < 3 ? 3 * a + h : 9));
        const bool valid = tap < p.taps;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (uint32_t)((a - acc0) * NB);
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 32) {
          if (!valid) continue;                    // (warp-uniform: a warp's 32 lanes lie in one atom)
          float v[32];
          tmem_ld_x32(taddr + c0, v);
          tmem_ld_wait();
          float* dst = p.dw + ((size_t)tap * p.co + co_tile * NB + c0) * p.ci + ci;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * p.ci, v[j]);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef void (*WgradTcFn)(const WgradTcParams);

template <int CIW, int NB>
size_t smem_bytes() {
  constexpr int RA = CIW * 2, RB = NB * 2;
  constexpr int XALLOC = (kHaloH * kHaloW * RA + 1023) & ~1023;
  constexpr int kStages = NB == 128 ? 3 : 4;
  return 1024 + (size_t)kStages * (XALLOC + kTH * kTW * RB) + 1024 + 512;
}

}  // namespace
}  // namespace lass

using namespace lass;

static int wgrad_tc_impl(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride, int x_coff, int ci,
                         int B, int H, int W, int taps, float* dw, void* stream_v, bool zero_first) {
  if (!dy || !x || !dw) return set_error(LASS_ERR_ARG, "lass_wgrad_tc: null pointer");
  if (B <= 0 || H <= 0 || W <= 0 || (taps != 9 && taps != 1) || co <= 0 || ci <= 0 || co % 32 || ci % 32 || dy_cstride % 8 || dy_coff % 8 ||
      x_cstride % 8 || x_coff % 8 || dy_coff + co > dy_cstride || x_coff + ci > x_cstride)
    return set_error(LASS_ERR_ARG, "lass_wgrad_tc: bad shape co=%d ci=%d taps=%d B=%d H=%d W=%d", co, ci, taps, B, H, W);
  cudaStream_t s = (cudaStream_t)stream_v;
  const int ciw = ci % 64 == 0 ? 64 : 32;
  const int nb = (co % 128 == 0 && ciw == 64) ? 128 : (co % 64 == 0 ? 64 : 32);
  const int nbox = nb == 128 ? 64 : nb;
  WgradTcParams p;
  int e;
  {
    const char* base = reinterpret_cast<const char*>(x) + (size_t)x_coff * 2;
    uint64_t dims[4] = {(uint64_t)ci, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)x_cstride * 2, (uint64_t)x_cstride * 2 * W, (uint64_t)x_cstride * 2 * W * H};
    uint32_t box[4] = {(uint32_t)ciw, (uint32_t)kHaloW, (uint32_t)kHaloH, 1};
    if ((e = make_tensor_map(&p.tmX, base, 2, 4, dims, strides, box, ciw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))) return e;
  }
  {
    const char* base = reinterpret_cast<const char*>(dy) + (size_t)dy_coff * 2;
    uint64_t dims[4] = {(uint64_t)co, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)dy_cstride * 2, (uint64_t)dy_cstride * 2 * W, (uint64_t)dy_cstride * 2 * W * H};
    uint32_t box[4] = {(uint32_t)nbox, (uint32_t)kTW, (uint32_t)kTH, 1};
    if ((e = make_tensor_map(&p.tmY, base, 2, 4, dims, strides, box, nbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))) return e;
  }
  p.dw = dw;
  p.co = co;
  p.ci = ci;
  p.taps = taps;
  p.x_fp16 = x_fp16 ? 1 : 0;
  p.tiles_h = (H + kTH - 1) / kTH;
  p.tiles_w = (W + kTW - 1) / kTW;
  p.num_pix_tiles = B * p.tiles_h * p.tiles_w;
  p.n_ci_chunks = ci / ciw;
  p.n_co_tiles = co / nb;
  const int out_tiles = p.n_ci_chunks * p.n_co_tiles;
  const int sms = device_sm_count();
  int splits = (2 * sms + out_tiles - 1) / out_tiles;
  if (out_tiles >= sms) splits = 1;
  if (splits > p.num_pix_tiles) splits = p.num_pix_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.splits0 = splits;
  if (nb == 128 && taps == 9) {
    // two CTA sets per output tile (MMA groups 0-2 / 3-4): at least one CTA each, pixel ranges 3 : 2
    if (splits < 2) splits = 2;
    if (splits > 2 * p.num_pix_tiles) splits = 2 * p.num_pix_tiles;
    if (splits < 2) return set_error(LASS_ERR_ARG, "lass_wgrad_tc: empty problem");
    int s0 = (3 * splits + 2) / 5;
    if (s0 < 1) s0 = 1;
    if (s0 > splits - 1) s0 = splits - 1;
    if (s0 > p.num_pix_tiles) s0 = p.num_pix_tiles;
    if (splits - s0 > p.num_pix_tiles) splits = s0 + p.num_pix_tiles;
    p.splits = splits;
    p.splits0 = s0;
  }
  WgradTcFn fn;
  size_t smem;
  if (nb == 128) { fn = wgrad_tc_kernel<64, 128>; smem = smem_bytes<64, 128>(); }
  else if (ciw == 64 && nb == 64) { fn = wgrad_tc_kernel<64, 64>; smem = smem_bytes<64, 64>(); }
  else if (ciw == 64) { fn = wgrad_tc_kernel<64, 32>; smem = smem_bytes<64, 32>(); }
  else if (nb == 64) { fn = wgrad_tc_kernel<32, 64>; smem = smem_bytes<32, 64>(); }
  else { fn = wgrad_tc_kernel<32, 32>; smem = smem_bytes<32, 32>(); }
  if (smem < 120 * 1024) smem = 120 * 1024;       // one CTA per SM: two could not both hold their TMEM accumulators
  cudaError_t ce = zero_first ? cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * co * ci, s) : cudaSuccess;
  if (ce != cudaSuccess) return set_cuda_error(ce, "wgrad_tc memset");
  ce = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (ce != cudaSuccess) return set_cuda_error(ce, "wgrad_tc smem attribute");
  return set_cuda_error(launch_pdl(fn, dim3((unsigned)(out_tiles * splits)), dim3(kThreads), smem, s, p), "wgrad_tc launch");
}

extern "C" int lass_wgrad_tc(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride, int x_coff, int ci,
                             int B, int H, int W, int taps, float* dw, void* stream_v) {
  return wgrad_tc_impl(dy, dy_cstride, dy_coff, co, x, x_fp16, x_cstride, x_coff, ci, B, H, W, taps, dw, stream_v, true);
}

// the same sums ADDED to dw (no memset inside: the training step clears its whole gradient buffer with one memset, and a launch
// that is not preceded by a memset node can overlap its predecessor's tail -- programmatic dependent launch)
extern "C" int lass_wgrad_tc_acc(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride, int x_coff,
                                 int ci, int B, int H, int W, int taps, float* dw, void* stream_v) {
  return wgrad_tc_impl(dy, dy_cstride, dy_coff, co, x, x_fp16, x_cstride, x_coff, ci, B, H, W, taps, dw, stream_v, false);
}
