// Small memory-bound helpers of the ResUNet30 forward:
//   K2 `film`     all FiLM linears (+ folded BatchNorm shifts) as one skinny fp32 GEMM -> per-(clip, channel)
//                 activation shift table        (reference models/resunet.py:59-81, 38x nn.Linear)
// (bn0 + zero time padding + Nyquist drop + pre_conv, reference models/resunet.py:537-555, no longer have a kernel: both
//  convolutions of encoder_block1 regenerate their operand / residual from the magnitude, see conv.cu "generated A".)
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

}  // namespace

cudaError_t launch_film(const float* cond, const float* W, const float* bias, float* shift, int B, int K, int J,
                        cudaStream_t stream) {
  dim3 grid((unsigned)((J + kFilmTile - 1) / kFilmTile), (unsigned)((B + kFilmTile - 1) / kFilmTile));
  film_kernel<<<grid, 256, 0, stream>>>(cond, W, bias, shift, B, K, J);
  return cudaGetLastError();
}

}  // namespace lass
