// K5 `mask_istft`: complex-mask application + inverse STFT (Hermitian irfft in shared memory) +
// windowed overlap-add + window-sum normalisation + trim, in ONE memory-bound kernel.
//
// Replaces (reference, per forward): `ResUNet30_Base.feature_maps_to_wav` models/resunet.py:436-519
// (sigmoid / tanh / magphase / phase rotation / relu) and torchlibrosa 0.1.0 `ISTFT.forward`
// (Hermitian extension, two dense n_fft x n_fft 1x1 conv1d, F.fold overlap-add, F.fold window sum,
// clamp 1e-11, trim n_fft/2) — SURVEY.md §8a rows a10-a12.
//
// Work decomposition: CTA (chunk, clip) owns the padded-signal samples [chunk*S, (chunk+1)*S) with
// S = FR*hop.  It processes every frame that overlaps its range (FR + ceil(n_fft/hop) - 1 frames, the
// halo frames are recomputed rather than exchanged so the result is deterministic), G frames at a
// time: mask math -> half spectrum in smem -> Hermitian pack -> Stockham inverse FFT of size n_fft/2
// -> each thread accumulates its own output positions (no atomics).
//
// Algorithmic HBM bytes per clip: 6 planes * 4 B * T * F (feat x3, mag, cos, sin) + 4 B * L.
#include "fft.cuh"
#include "lass_internal.cuh"

namespace lass {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

struct MaskIstftArgs {
  const float* feat;  // 3 planes
  long long feat_bstride, feat_cstride;
  int feat_tstride;  // elements between frames in a feat plane
  int feat_F;        // valid bins in feat (bins >= feat_F behave as x = 0)
  const float* mag;
  const float* cosp;
  const float* sinp;  // (B, T, F) contiguous
  const float* window;  // [N] synthesis window (periodic Hann), NOT scaled
  const cpx* tw;        // [N] exp(+2 pi i j / N)
  float* out;           // (B, L)
  int T, F, N, log2M, hop, L, FR;
};

// Fast-math forms (errors ~1e-6 relative, far inside the 1e-4 bar; checked by tests/test_gpu_spectral.py).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// One warp per frame: mask math -> half spectrum -> Hermitian pack -> Stockham inverse FFT, all in the warp's own
// shared-memory buffers with __syncwarp only; block barriers only around the overlap-add of each round of kWarps frames.
__global__ void __launch_bounds__(kThreads) mask_istft_kernel(const MaskIstftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, M = N >> 1, hop = a.hop;
  const int S = a.FR * hop;
  cpx* tw = reinterpret_cast<cpx*>(smem_raw);          // N
  float* win = reinterpret_cast<float*>(tw + N);       // N
  float* ola = win + N;                                // S
  cpx* bufs = reinterpret_cast<cpx*>(ola + S);         // kWarps * (2M + 2)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cpx* bufA = bufs + (size_t)warp * (2 * M + 2);
  cpx* bufB = bufA + M;                                // M + 1 (+1 pad): half spectrum X, then FFT ping-pong

  const int b = blockIdx.y;
  const long long pos0 = (long long)blockIdx.x * S;    // first padded-signal sample owned by this CTA

  for (int i = tid; i < N; i += kThreads) {
    tw[i] = a.tw[i];
    win[i] = a.window[i];
  }
  for (int i = tid; i < S; i += kThreads) ola[i] = 0.0f;

  // frames overlapping [pos0, pos0 + S): t*hop + N > pos0  and  t*hop < pos0 + S
  long long t_lo = (pos0 - N) / hop + 1;
  if (pos0 - N < 0) t_lo = 0;
  long long t_hi = (pos0 + S - 1) / hop;               // inclusive
  if (t_hi > a.T - 1) t_hi = a.T - 1;
  __syncthreads();

  const float* magb = a.mag + (size_t)b * a.T * a.F;
  const float* cosb = a.cosp + (size_t)b * a.T * a.F;
  const float* sinb = a.sinp + (size_t)b * a.T * a.F;
  const float* featb = a.feat + (size_t)b * a.feat_bstride;
  const int passes = fft_num_passes(a.log2M);
  const bool result_in_A = (passes & 1) == 0;

  for (long long tb = t_lo; tb <= t_hi; tb += kWarps) {
    const long long t = tb + warp;
    if (t <= t_hi) {
      // 1. masked half spectrum X[k], k in [0, M]  (models/resunet.py:469-495)
      cpx* X = bufB;
      const float* fp = featb + (size_t)t * a.feat_tstride;
      const size_t o = (size_t)t * a.F;
      // four bins per lane per iteration, all 24 loads issued before any math (memory-level parallelism)
      for (int k0 = lane; k0 <= M; k0 += 128) {
        float x0[4], x1[4], x2[4], sp[4], cs[4], sn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 32 * u;
          const bool inF = k < a.F, inX = k < a.feat_F;
          x0[u] = inX ? __ldg(fp + k) : 0.0f;
          x1[u] = inX ? __ldg(fp + a.feat_cstride + k) : 0.0f;
          x2[u] = inX ? __ldg(fp + 2 * a.feat_cstride + k) : 0.0f;
          sp[u] = inF ? __ldg(magb + o + k) : 0.0f;
          cs[u] = inF ? __ldg(cosb + o + k) : 0.0f;
          sn[u] = inF ? __ldg(sinb + o + k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 32 * u;
          if (k > M) break;
          const float mr = tanh_fast(x1[u]), mi = tanh_fast(x2[u]);
          // torchlibrosa magphase: clamp(|m|, 1e-10)
          const float inv = __fdividef(1.0f, fmaxf(sqrtf(mr * mr + mi * mi), 1e-10f));
          const float mc = mr * inv, ms = mi * inv;
          const float om = fmaxf(sp[u] * sigmoid_fast(x0[u]), 0.0f);   // sp = 0 for bins >= F
          X[k] = cpx{om * (cs[u] * mc - sn[u] * ms), om * (sn[u] * mc + cs[u] * ms)};
        }
      }
      __syncwarp();
      // 2. Hermitian pack -> bufA
      for (int k = lane; k < M; k += 32) bufA[k] = irfft_pack(X, tw, M, k);
      __syncwarp();
      // 3. inverse FFT passes (ping-pong A -> B -> A ...)
      cpx* src = bufA;
      cpx* dst = bufB;
      for (int p = 0; p < passes; ++p) {
        const int nb = fft_pass_butterflies(a.log2M, p);
        for (int j = lane; j < nb; j += 32) ifft_butterfly(src, dst, tw, a.log2M, p, j);
        __syncwarp();
        cpx* tmp = src;
        src = dst;
        dst = tmp;
      }
    }
    __syncthreads();
    // 4. windowed overlap-add: every thread owns output positions tid, tid + kThreads, ... (deterministic, no atomics)
    const int cnt = (int)((t_hi - tb + 1 < kWarps) ? (t_hi - tb + 1) : kWarps);
    // only the positions this round's frames touch: [tb*hop, (tb+cnt-1)*hop + N) clipped to the CTA's range
    long long w0 = tb * hop - pos0;
    if (w0 < 0) w0 = 0;
    long long w1 = (tb + cnt - 1) * hop + N - pos0;
    if (w1 > S) w1 = S;
    const float* zbase = reinterpret_cast<const float*>(bufs + (result_in_A ? 0 : M));
    const int zstride = 2 * (2 * M + 2);                 // floats between consecutive warps' buffers
    for (int pos = (int)w0 + tid; pos < (int)w1; pos += kThreads) {
      float acc = ola[pos];
      const int rel = (int)(pos0 + pos - tb * hop);      // sample offset relative to frame tb's start (>= 0)
      // frames g with 0 <= rel - g*hop < N
      int g_hi = rel / hop;
      if (g_hi > cnt - 1) g_hi = cnt - 1;
      int g_lo = (rel - N) / hop + 1;
      if (rel - N < 0) g_lo = 0;
      for (int g = g_lo; g <= g_hi; ++g) {
        const int n = rel - g * hop;
        acc += zbase[(size_t)g * zstride + n] * win[n];  // interleaved: x[2m] = Re z[m], x[2m+1] = Im z[m]
      }
      ola[pos] = acc;
    }
    __syncthreads();
  }

  // 5. window-sum normalisation (folded hann^2, clamp 1e-11), 1/N of the inverse DFT, trim n_fft/2
  const float invN = 1.0f / (float)N;
  for (int pos = tid; pos < S; pos += kThreads) {
    const long long np = pos0 + pos;          // padded-signal index
    const long long no = np - (N >> 1);       // output index
    if (no < 0 || no >= a.L) continue;
    long long ta = (np - N) / hop + 1;
    if (np - N < 0) ta = 0;
    long long tz = np / hop;
    if (tz > a.T - 1) tz = a.T - 1;
    float ws = 0.0f;
    for (long long t = ta; t <= tz; ++t) {
      const float w = win[np - t * hop];
      ws += w * w;
    }
    a.out[(size_t)b * a.L + no] = (ola[pos] * invN) / fmaxf(ws, 1e-11f);
  }
}

}  // namespace

size_t mask_istft_smem_bytes(int N, int hop, int FR) {
  const int M = N / 2;
  return (size_t)N * sizeof(cpx) + (size_t)N * 4 + (size_t)FR * hop * 4 + (size_t)kWarps * (2 * M + 2) * sizeof(cpx) + 16;
}

cudaError_t launch_mask_istft(const float* feat, long long feat_bstride, long long feat_cstride, int feat_tstride,
                              int feat_F, const float* mag, const float* cosp, const float* sinp,
                              const float* window, const float* tw, float* out, int B, int T, int F, int N,
                              int hop, int L, cudaStream_t stream) {
  MaskIstftArgs a;
  a.feat = feat;
  a.feat_bstride = feat_bstride;
  a.feat_cstride = feat_cstride;
  a.feat_tstride = feat_tstride;
  a.feat_F = feat_F;
  a.mag = mag;
  a.cosp = cosp;
  a.sinp = sinp;
  a.window = window;
  a.tw = reinterpret_cast<const cpx*>(tw);
  a.out = out;
  a.T = T;
  a.F = F;
  a.N = N;
  a.hop = hop;
  a.L = L;
  int log2N = 0;
  while ((1 << log2N) < N) ++log2N;
  a.log2M = log2N - 1;
  // CTA = FR hops of output; as large as fits two CTAs per SM (halo frames are recomputed: ~n_fft/hop per CTA)
  a.FR = 64;
  size_t smem = mask_istft_smem_bytes(N, hop, a.FR);
  while (smem > 110 * 1024 && a.FR > 8) {
    a.FR -= 8;
    smem = mask_istft_smem_bytes(N, hop, a.FR);
  }
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mask_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const long long P = (long long)(T - 1) * hop + N;
  const int S = a.FR * hop;
  dim3 grid((unsigned)((P + S - 1) / S), (unsigned)B);
  mask_istft_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace lass
