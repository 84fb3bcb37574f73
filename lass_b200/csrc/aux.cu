// Small memory-bound helpers of the ResUNet30 forward:
//   K2 `film`     all FiLM linears (+ folded BatchNorm shifts) as one skinny fp32 GEMM -> per-(clip, channel)
//                 activation shift table        (reference models/resunet.py:59-81, 38x nn.Linear)
//   `preconv`     bn0 over frequency + zero time padding + Nyquist drop + pre_conv 1x1 (1 -> 32) fused,
//                 writing the raw (fp16) and activated (bf16) NHWC inputs of encoder_block1
//                 (reference models/resunet.py:537-555)
#include "lass_internal.cuh"
#include "ptx.cuh"

namespace lass {

namespace {

// shift[b][j] = bias[j] + sum_k cond[b][k] * W[j][k] as a shared-memory tiled fp32 GEMM: CTA = 64 table rows x 64 clips,
// K in chunks of 64, thread (tx, ty) accumulates the 4 x 4 outputs (j = ty + 16 u, b = tx + 16 v).
constexpr int kFilmTile = 64;
constexpr int kFilmPitch = kFilmTile + 1;
__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ cond, const float* __restrict__ W,
                                                   const float* __restrict__ bias, float* __restrict__ shift,
                                                   int B, int K, int J) {
  __shared__ float Ws[kFilmTile * kFilmPitch];
  __shared__ float Cs[kFilmTile * kFilmPitch];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int j0 = blockIdx.x * kFilmTile, b0 = blockIdx.y * kFilmTile;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += kFilmTile) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + 256 * i, r = e >> 6, c = e & 63;
      const bool kin = k0 + c < K;
      Ws[r * kFilmPitch + c] = (kin && j0 + r < J) ? __ldg(W + (size_t)(j0 + r) * K + k0 + c) : 0.0f;
      Cs[r * kFilmPitch + c] = (kin && b0 + r < B) ? __ldg(cond + (size_t)(b0 + r) * K + k0 + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < kFilmTile; ++c) {
      float wv[4], cv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        wv[u] = Ws[(ty + 16 * u) * kFilmPitch + c];
        cv[u] = Cs[(tx + 16 * u) * kFilmPitch + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(wv[u], cv[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = j0 + ty + 16 * u;
    if (j >= J) continue;
    const float bj = __ldg(bias + j);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int b = b0 + tx + 16 * v;
      if (b < B) shift[(size_t)b * J + j] = acc[u][v] + bj;
    }
  }
}

// 4 threads per pixel, 8 channels each (one 16 B store per tensor per thread).  A CTA works on one clip, so every
// per-channel constant (pre_conv weight / bias, folded BN scale, FiLM shift of that clip) is loaded once per thread and
// the thread then walks over kPixIter pixels.
constexpr int kPixIter = 16;
__global__ void __launch_bounds__(256) preconv_kernel(const float* __restrict__ mag, const float* __restrict__ bn0_scale,
                                                      const float* __restrict__ bn0_shift, const float* __restrict__ pre_w,
                                                      const float* __restrict__ pre_b, const float* __restrict__ act_scale,
                                                      const float* __restrict__ act_shift, int shift_bstride,
                                                      __half* __restrict__ raw, __nv_bfloat16* __restrict__ act, int T,
                                                      int F, int Tp, int Fp, int log2Fp) {
  const int b = blockIdx.y;
  const int cg = (threadIdx.x & 3) * 8;
  // activated output = lrelu(as * (pw * v + pb) + sh) = lrelu(ca * v + cb): one FMA per channel in the pixel loop
  float pw[8], pb[8], ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    pw[j] = __ldg(pre_w + cg + j);
    pb[j] = __ldg(pre_b + cg + j);
    const float as = __ldg(act_scale + cg + j);
    ca[j] = as * pw[j];
    cb[j] = fmaf(as, pb[j], __ldg(act_shift + (size_t)b * shift_bstride + cg + j));
  }
  const int pix_per_clip = Tp * Fp;
  const int pix0 = blockIdx.x * (64 * kPixIter) + (threadIdx.x >> 2);
  const float* magb = mag + (size_t)b * T * F;
  // all of a thread's magnitude loads are issued before any of its stores (memory-level parallelism)
  float vs[kPixIter];
#pragma unroll
  for (int it = 0; it < kPixIter; ++it) {
    const int pix = pix0 + it * 64;
    const int t = pix >> log2Fp, f = pix & (Fp - 1);
    vs[it] = 0.0f;  // time-padding rows are zero AFTER bn0 (models/resunet.py:548)
    if (pix < pix_per_clip && t < T) vs[it] = fmaf(__ldg(bn0_scale + f), __ldg(magb + (size_t)t * F + f), __ldg(bn0_shift + f));
  }
#pragma unroll
  for (int it = 0; it < kPixIter; ++it) {
    const int pix = pix0 + it * 64;
    if (pix >= pix_per_clip) break;
    const float v = vs[it];
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y = fmaf(ca[j], v, cb[j]);
      a[j] = fmaxf(y, 0.01f * y);
    }
    const size_t o = ((size_t)b * pix_per_clip + pix) * 32 + cg;
    uint4 pa;
    pa.x = pack_bf16x2(a[0], a[1]);
    pa.y = pack_bf16x2(a[2], a[3]);
    pa.z = pack_bf16x2(a[4], a[5]);
    pa.w = pack_bf16x2(a[6], a[7]);
    *reinterpret_cast<uint4*>(act + o) = pa;
    if (raw != nullptr) {   // the fused path regenerates the raw tensor where it is needed (conv epilogue) and passes NULL
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(pw[j], v, pb[j]);
      uint4 pr;
      pr.x = pack_f16x2_sat(r[0], r[1]);
      pr.y = pack_f16x2_sat(r[2], r[3]);
      pr.z = pack_f16x2_sat(r[4], r[5]);
      pr.w = pack_f16x2_sat(r[6], r[7]);
      *reinterpret_cast<uint4*>(raw + o) = pr;
    }
  }
}

}  // namespace

cudaError_t launch_film(const float* cond, const float* W, const float* bias, float* shift, int B, int K, int J,
                        cudaStream_t stream) {
  dim3 grid((unsigned)((J + kFilmTile - 1) / kFilmTile), (unsigned)((B + kFilmTile - 1) / kFilmTile));
  film_kernel<<<grid, 256, 0, stream>>>(cond, W, bias, shift, B, K, J);
  return cudaGetLastError();
}

cudaError_t launch_preconv(const float* mag, const float* bn0_scale, const float* bn0_shift, const float* pre_w,
                           const float* pre_b, const float* act_scale, const float* act_shift, int shift_bstride,
                           void* raw, void* act, int B, int T, int F, int Tp, int Fp, cudaStream_t stream) {
  const int pix_per_clip = Tp * Fp;
  int log2Fp = 0;
  while ((1 << log2Fp) < Fp) ++log2Fp;
  if ((1 << log2Fp) != Fp) return cudaErrorInvalidValue;   // Fp = n_fft / 2 is a power of two
  dim3 grid((unsigned)((pix_per_clip + 64 * kPixIter - 1) / (64 * kPixIter)), (unsigned)B);
  preconv_kernel<<<grid, 256, 0, stream>>>(mag, bn0_scale, bn0_shift, pre_w, pre_b, act_scale, act_shift, shift_bstride,
                                           reinterpret_cast<__half*>(raw), reinterpret_cast<__nv_bfloat16*>(act), T, F, Tp,
                                           Fp, log2Fp);
  return cudaGetLastError();
}

}  // namespace lass
