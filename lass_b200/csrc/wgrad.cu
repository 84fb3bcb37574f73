// Training step: weight gradients of the convolutions (replaces the autograd backward-weight of conv2d 3x3 / 1x1 and
// conv_transpose2d(kernel = stride) of ConvBlockRes / DecoderBlockRes1B, reference models/resunet.py:147-165, 240-264).
//
//   dW[tap][co][ci] = sum over pixels p of dY[p][co] * X[p + tap][ci]        (zero padding; tap = ky*3 + kx, centre = 4)
//
// GEMM view: M = co, N = ci, K = pixels.  Both operands are NHWC, i.e. "MN-major" (the contraction index is the slow one),
// so the tiles go to shared memory as they lie in HBM ([pixel][channel]) and the fragments are read with ldmatrix.trans.
// One block = one output tile (CO_T x CI_T channels x NT taps: a kernel ROW of a 3x3 conv, or the single tap of a 1x1) over a
// contiguous range of 8 x 16 pixel tiles (split-K: a layer with few weights and many pixels is cut into many pixel ranges,
// a layer with many weights into few), accumulated in registers (bf16 mma.sync m16n8k16, fp32 accumulate) and added to
// dW with fp32 atomics once at the end.  fp16 sources (the raw residual stream) are converted to bf16 on the way in.
// A warp owns a 32 co x 32 ci x NT accumulator; warps beyond the tile's (CO_T/32) x (CI_T/32) split the pixel rows of a tile.
#include "lass_internal.cuh"

namespace lass {
namespace {

constexpr int kThreads = 256;
constexpr int kTileH = 8, kTileW = 16;            // pixels per tile: one mma K-step = one tile row of 16 pixels
constexpr int kHaloW = kTileW + 2;

struct WgradParams {
  const uint16_t* dy;
  const uint16_t* x;
  float* dw;
  int dy_cstride, dy_coff, x_cstride, x_coff, x_fp16;
  int co, ci, B, H, W, taps;
  int tiles_h, tiles_w, num_pix_tiles;
  int n_co_tiles, n_ci_tiles, n_rows;             // output tiles: co tiles x ci tiles x kernel rows (3 or 1)
  int splits;                                     // pixel ranges per output tile
};

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 half8_to_bf16(uint4 q) {
  uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    const __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int CO_T, int CI_T, int NT>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const WgradParams p) {
  constexpr int WM = CO_T / 32, WN = CI_T / 32, KS = 8 / (WM * WN);      // warp grid and pixel-row split
  constexpr int YP = CO_T + 8, XP = CI_T + 8;                             // shared-memory row pitches (elements): conflict-free ldmatrix
  constexpr int YV = CO_T / 8, XV = CI_T / 8;                             // 16-byte vectors per pixel
  constexpr int NY = (kTileH * kTileW * YV) / kThreads;                   // dY vectors per thread
  constexpr int NXV = kTileH * kHaloW * XV;
  constexpr int NX = (NXV + kThreads - 1) / kThreads;
  static_assert((kTileH * kTileW * YV) % kThreads == 0, "dY tile must divide over the block");
  extern __shared__ __align__(16) uint16_t smem_w[];
  uint16_t* sY = smem_w;                                  // [kTileH * kTileW][YP]
  uint16_t* sX = smem_w + kTileH * kTileW * YP;           // [kTileH * kHaloW][XP]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % WM, wn = (warp / WM) % WN, kg = warp / (WM * WN);

  // ---- this block's output tile and pixel-tile range ----
  int item = blockIdx.x;
  const int split = item % p.splits;
  item /= p.splits;
  const int krow = item % p.n_rows;
  item /= p.n_rows;
  const int ci_tile = item % p.n_ci_tiles, co_tile = item / p.n_ci_tiles;
  const int co0 = co_tile * CO_T, ci0 = ci_tile * CI_T;
  const int t_begin = (int)((long long)p.num_pix_tiles * split / p.splits);
  const int t_end = (int)((long long)p.num_pix_tiles * (split + 1) / p.splits);
  const int dyrow = (NT == 3) ? krow - 1 : 0;       // input row offset of this kernel row
  const int dx0 = (NT == 3) ? 0 : 1;                // first halo column used (a 1x1 tap reads the centre column)

  float acc[NT][2][4][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][mi][ni][e] = 0.0f;

  uint4 ry[NY], rx[NX];
  auto load_tile = [&](int tile) {
    const int tw = tile % p.tiles_w;
    int r = tile / p.tiles_w;
    const int th = r % p.tiles_h, b = r / p.tiles_h;
    const int h0 = th * kTileH, w0 = tw * kTileW;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const int v = tid + i * kThreads;
      const int cv = v % YV, px = v / YV;
      const int h = h0 + px / kTileW, w = w0 + px % kTileW;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (h < p.H && w < p.W)
        q = __ldg(reinterpret_cast<const uint4*>(p.dy + (((size_t)b * p.H + h) * p.W + w) * p.dy_cstride + p.dy_coff + co0 + cv * 8));
      ry[i] = q;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int v = tid + i * kThreads;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (v < NXV) {
        const int cv = v % XV, px = v / XV;
        const int h = h0 + px / kHaloW + dyrow, w = w0 + px % kHaloW - 1;
        if (h >= 0 && h < p.H && w >= 0 && w < p.W) {
          q = __ldg(reinterpret_cast<const uint4*>(p.x + (((size_t)b * p.H + h) * p.W + w) * p.x_cstride + p.x_coff + ci0 + cv * 8));
          if (p.x_fp16) q = half8_to_bf16(q);
        }
      }
      rx[i] = q;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const int v = tid + i * kThreads;
      const int cv = v % YV, px = v / YV;
      *reinterpret_cast<uint4*>(&sY[px * YP + cv * 8]) = ry[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int v = tid + i * kThreads;
      if (v < NXV) {
        const int cv = v % XV, px = v / XV;
        *reinterpret_cast<uint4*>(&sX[px * XP + cv * 8]) = rx[i];
      }
    }
  };

  // ldmatrix lane addressing (see the file header): matrix id = lane / 8, row within the 8x8 matrix = lane % 8
  const int lid = lane >> 3, lrow = lane & 7;
  // A (dY^T): matrices (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15)
  const int a_k = (lid >> 1) * 8 + lrow, a_m = (lid & 1) * 8;
  // B (X):    matrices (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
  const int b_k = (lid & 1) * 8 + lrow, b_n = (lid >> 1) * 8;
  const uint32_t sY_base = (uint32_t)__cvta_generic_to_shared(sY);
  const uint32_t sX_base = (uint32_t)__cvta_generic_to_shared(sX);

  if (t_begin < t_end) load_tile(t_begin);
  for (int tile = t_begin; tile < t_end; ++tile) {
    __syncthreads();                    // the previous tile's fragments have been read
    store_tile();
    __syncthreads();
    if (tile + 1 < t_end) load_tile(tile + 1);     // global loads of the next tile fly during the MMAs of this one
#pragma unroll 1
    for (int r = kg; r < kTileH; r += KS) {
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const uint32_t addr = sY_base + (uint32_t)(((r * kTileW + a_k) * YP + wm * 32 + mi * 16 + a_m) * 2);
        ldsm_x4_trans(addr, a[mi][0], a[mi][1], a[mi][2], a[mi][3]);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int nj = 0; nj < 2; ++nj) {
          uint32_t b0, b1, b2, b3;
          const uint32_t addr = sX_base + (uint32_t)(((r * kHaloW + dx0 + t + b_k) * XP + wn * 32 + nj * 16 + b_n) * 2);
          ldsm_x4_trans(addr, b0, b1, b2, b3);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            mma_bf16(acc[t][mi][nj * 2], a[mi], b0, b1);
            mma_bf16(acc[t][mi][nj * 2 + 1], a[mi], b2, b3);
          }
        }
      }
    }
  }

  // ---- add this block's partial sums to dW (taps, co, ci) ----
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int tap = (NT == 3) ? krow * 3 + t : 0;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int co = co0 + wm * 32 + mi * 16 + g;
        const int ci = ci0 + wn * 32 + ni * 8 + tq * 2;
        float* d0 = p.dw + ((size_t)tap * p.co + co) * p.ci + ci;
        float* d1 = d0 + (size_t)8 * p.ci;
        atomicAdd(d0, acc[t][mi][ni][0]);
        atomicAdd(d0 + 1, acc[t][mi][ni][1]);
        atomicAdd(d1, acc[t][mi][ni][2]);
        atomicAdd(d1 + 1, acc[t][mi][ni][3]);
      }
  }
}

typedef void (*WgradFn)(const WgradParams);

template <int NT>
WgradFn pick(int co_t, int ci_t) {
  if (co_t == 128 && ci_t == 64) return wgrad_kernel<128, 64, NT>;
  if (co_t == 64 && ci_t == 64) return wgrad_kernel<64, 64, NT>;
  if (co_t == 64 && ci_t == 32) return wgrad_kernel<64, 32, NT>;
  if (co_t == 32 && ci_t == 64) return wgrad_kernel<32, 64, NT>;
  if (co_t == 32 && ci_t == 32) return wgrad_kernel<32, 32, NT>;
  return nullptr;
}

}  // namespace
}  // namespace lass

using namespace lass;

extern "C" int lass_wgrad(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride, int x_coff, int ci,
                          int B, int H, int W, int taps, float* dw, void* stream_v) {
  if (!dy || !x || !dw) return set_error(LASS_ERR_ARG, "lass_wgrad: null pointer");
  if (B <= 0 || H <= 0 || W <= 0 || (taps != 9 && taps != 1) || co <= 0 || ci <= 0 || co % 32 || ci % 32 || dy_cstride % 8 || dy_coff % 8 ||
      x_cstride % 8 || x_coff % 8 || dy_coff + co > dy_cstride || x_coff + ci > x_cstride)
    return set_error(LASS_ERR_ARG, "lass_wgrad: bad shape co=%d ci=%d taps=%d B=%d H=%d W=%d", co, ci, taps, B, H, W);
  cudaStream_t s = (cudaStream_t)stream_v;
  int co_t = (co % 128 == 0) ? 128 : (co % 64 == 0 ? 64 : 32);
  const int ci_t = (ci % 64 == 0) ? 64 : 32;
  if (co_t == 128 && ci_t == 32) co_t = 64;
  WgradFn fn = taps == 9 ? pick<3>(co_t, ci_t) : pick<1>(co_t, ci_t);
  if (!fn) return set_error(LASS_ERR_ARG, "lass_wgrad: no kernel for the %d x %d tile", co_t, ci_t);
  WgradParams p;
  p.dy = reinterpret_cast<const uint16_t*>(dy);
  p.x = reinterpret_cast<const uint16_t*>(x);
  p.dw = dw;
  p.dy_cstride = dy_cstride;
  p.dy_coff = dy_coff;
  p.x_cstride = x_cstride;
  p.x_coff = x_coff;
  p.x_fp16 = x_fp16;
  p.co = co;
  p.ci = ci;
  p.B = B;
  p.H = H;
  p.W = W;
  p.taps = taps;
  p.tiles_h = (H + kTileH - 1) / kTileH;
  p.tiles_w = (W + kTileW - 1) / kTileW;
  p.num_pix_tiles = B * p.tiles_h * p.tiles_w;
  p.n_co_tiles = co / co_t;
  p.n_ci_tiles = ci / ci_t;
  p.n_rows = taps == 9 ? 3 : 1;
  const int out_tiles = p.n_co_tiles * p.n_ci_tiles * p.n_rows;
  const int target = 2 * device_sm_count();
  int splits = (target + out_tiles - 1) / out_tiles;
  if (splits > p.num_pix_tiles) splits = p.num_pix_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * co * ci, s);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad memset");
  const size_t smem = ((size_t)kTileH * kTileW * (co_t + 8) + (size_t)kTileH * kHaloW * (ci_t + 8)) * 2;
  e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad smem attribute");
  fn<<<out_tiles * splits, kThreads, smem, s>>>(p);
  return set_cuda_error(cudaGetLastError(), "wgrad launch");
}
