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

// One warp per table row j: shift[b][j] = bias[j] + sum_k cond[b][k] * W[j][k]
__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ cond, const float* __restrict__ W,
                                                   const float* __restrict__ bias, float* __restrict__ shift,
                                                   int B, int K, int J) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= J) return;
  // K <= 1024: up to 32 weights per lane in registers
  float w[32];
  const int per = (K + 31) / 32;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = i * 32 + lane;
    w[i] = (i < per && k < K) ? __ldg(W + (size_t)j * K + k) : 0.0f;
  }
  const float bj = __ldg(bias + j);
  for (int b = 0; b < B; ++b) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = i * 32 + lane;
      if (i < per && k < K) acc = fmaf(w[i], __ldg(cond + (size_t)b * K + k), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) shift[(size_t)b * J + j] = acc + bj;
  }
}

// 4 threads per pixel, 8 channels each (one 16 B store per tensor per thread)
__global__ void __launch_bounds__(256) preconv_kernel(const float* __restrict__ mag, const float* __restrict__ bn0_scale,
                                                      const float* __restrict__ bn0_shift, const float* __restrict__ pre_w,
                                                      const float* __restrict__ pre_b, const float* __restrict__ act_scale,
                                                      const float* __restrict__ act_shift, int shift_bstride,
                                                      __half* __restrict__ raw, __nv_bfloat16* __restrict__ act, int T,
                                                      int F, int Tp, int Fp, long long total_pixels) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gid >> 2;
  if (pix >= total_pixels) return;
  const int cg = (int)(gid & 3) * 8;
  const int f = (int)(pix % Fp);
  const long long bt = pix / Fp;
  const int t = (int)(bt % Tp);
  const int b = (int)(bt / Tp);
  float v = 0.0f;  // time-padding rows are zero AFTER bn0 (models/resunet.py:548)
  if (t < T) v = fmaf(__ldg(bn0_scale + f), __ldg(mag + ((size_t)b * T + t) * F + f), __ldg(bn0_shift + f));
  float r[8], a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg + j;
    r[j] = fmaf(__ldg(pre_w + c), v, __ldg(pre_b + c));
    const float y = fmaf(__ldg(act_scale + c), r[j], __ldg(act_shift + (size_t)b * shift_bstride + c));
    a[j] = y > 0.0f ? y : 0.01f * y;
  }
  uint4 pa;
  pa.x = pack_bf16x2(a[0], a[1]);
  pa.y = pack_bf16x2(a[2], a[3]);
  pa.z = pack_bf16x2(a[4], a[5]);
  pa.w = pack_bf16x2(a[6], a[7]);
  *reinterpret_cast<uint4*>(act + pix * 32 + cg) = pa;
  if (raw != nullptr) {   // the fused path regenerates the raw tensor where it is needed (conv epilogue) and passes NULL
    uint4 pr;
    pr.x = pack_f16x2_sat(r[0], r[1]);
    pr.y = pack_f16x2_sat(r[2], r[3]);
    pr.z = pack_f16x2_sat(r[4], r[5]);
    pr.w = pack_f16x2_sat(r[6], r[7]);
    *reinterpret_cast<uint4*>(raw + pix * 32 + cg) = pr;
  }
}

}  // namespace

cudaError_t launch_film(const float* cond, const float* W, const float* bias, float* shift, int B, int K, int J,
                        cudaStream_t stream) {
  if (K > 1024) return cudaErrorInvalidValue;
  film_kernel<<<(J + 7) / 8, 256, 0, stream>>>(cond, W, bias, shift, B, K, J);
  return cudaGetLastError();
}

cudaError_t launch_preconv(const float* mag, const float* bn0_scale, const float* bn0_shift, const float* pre_w,
                           const float* pre_b, const float* act_scale, const float* act_shift, int shift_bstride,
                           void* raw, void* act, int B, int T, int F, int Tp, int Fp, cudaStream_t stream) {
  const long long pixels = (long long)B * Tp * Fp;
  const long long threads = pixels * 4;
  preconv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
      mag, bn0_scale, bn0_shift, pre_w, pre_b, act_scale, act_shift, shift_bstride, reinterpret_cast<__half*>(raw),
      reinterpret_cast<__nv_bfloat16*>(act), T, F, Tp, Fp, pixels);
  return cudaGetLastError();
}

}  // namespace lass
