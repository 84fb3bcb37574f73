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

constexpr int kThreads = 256;

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
  int T, F, N, log2M, hop, L, FR, G;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kThreads) mask_istft_kernel(const MaskIstftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, M = N >> 1, hop = a.hop;
  const int S = a.FR * hop;
  const int G = a.G;
  // smem carve-up
  cpx* tw = reinterpret_cast<cpx*>(smem_raw);          // N
  float* win = reinterpret_cast<float*>(tw + N);       // N
  float* ola = win + N;                                // S
  cpx* bufA = reinterpret_cast<cpx*>(ola + S);         // G * M
  cpx* bufB = bufA + (size_t)G * M;                    // G * (M + 1)   (also holds the half spectrum X)

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const long long pos0 = (long long)blockIdx.x * S;  // first padded sample owned by this CTA

  for (int i = tid; i < N; i += kThreads) {
    tw[i] = a.tw[i];
    win[i] = a.window[i];
  }
  for (int i = tid; i < S; i += kThreads) ola[i] = 0.0f;

  // frames overlapping [pos0, pos0 + S): t*hop + N > pos0  and  t*hop < pos0 + S
  long long t_lo = (pos0 - N) / hop + 1;
  if (pos0 - N < 0) t_lo = 0;
  long long t_hi = (pos0 + S - 1) / hop;  // inclusive
  if (t_hi > a.T - 1) t_hi = a.T - 1;
  __syncthreads();

  const float* magb = a.mag + (size_t)b * a.T * a.F;
  const float* cosb = a.cosp + (size_t)b * a.T * a.F;
  const float* sinb = a.sinp + (size_t)b * a.T * a.F;
  const float* featb = a.feat + (size_t)b * a.feat_bstride;
  const int passes = fft_num_passes(a.log2M);

  for (long long tb = t_lo; tb <= t_hi; tb += G) {
    const int g_cnt = (int)((t_hi - tb + 1 < G) ? (t_hi - tb + 1) : G);
    // 1. masked half spectrum X[k], k in [0, M]  (models/resunet.py:469-495)
    cpx* X = bufB;
    for (int i = tid; i < g_cnt * (M + 1); i += kThreads) {
      const int g = i / (M + 1), k = i - g * (M + 1);
      const size_t t = (size_t)(tb + g);
      cpx y{0.0f, 0.0f};
      if (k < a.F) {
        float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
        if (k < a.feat_F) {
          const float* fp = featb + t * a.feat_tstride + k;
          x0 = __ldg(fp);
          x1 = __ldg(fp + a.feat_cstride);
          x2 = __ldg(fp + 2 * a.feat_cstride);
        }
        const size_t o = t * a.F + k;
        const float sp = __ldg(magb + o), cs = __ldg(cosb + o), sn = __ldg(sinb + o);
        const float mask_mag = sigmoidf_acc(x0);
        const float mr = tanhf(x1), mi = tanhf(x2);
        // torchlibrosa magphase: clamp(|m|, 1e-10)
        const float mm = fmaxf(sqrtf(mr * mr + mi * mi), 1e-10f);
        const float mc = mr / mm, ms = mi / mm;
        const float oc = cs * mc - sn * ms;
        const float os = sn * mc + cs * ms;
        const float om = fmaxf(sp * mask_mag, 0.0f);
        y.x = om * oc;
        y.y = om * os;
      }
      X[(size_t)g * (M + 1) + k] = y;
    }
    __syncthreads();
    // 2. Hermitian pack -> bufA
    for (int i = tid; i < g_cnt * M; i += kThreads) {
      const int g = i / M, k = i - g * M;
      bufA[(size_t)g * M + k] = irfft_pack(X + (size_t)g * (M + 1), tw, M, k);
    }
    __syncthreads();
    // 3. inverse FFT passes (ping-pong A -> B -> A ...); B rows use stride M here (X no longer needed)
    cpx* src = bufA;
    cpx* dst = bufB;
    for (int p = 0; p < passes; ++p) {
      const int nb = fft_pass_butterflies(a.log2M, p);
      for (int i = tid; i < g_cnt * nb; i += kThreads) {
        const int g = i / nb, j = i - g * nb;
        ifft_butterfly(src + (size_t)g * M, dst + (size_t)g * M, tw, a.log2M, p, j);
      }
      __syncthreads();
      cpx* tmp = src;
      src = dst;
      dst = tmp;
    }
    // 4. windowed overlap-add: every thread owns output positions tid, tid + 256, ...
    const float* z = reinterpret_cast<const float*>(src);  // interleaved: x[2m] = Re z[m], x[2m+1] = Im z[m]
    for (int pos = tid; pos < S; pos += kThreads) {
      float acc = ola[pos];
      for (int g = 0; g < g_cnt; ++g) {
        const long long n = pos0 + pos - (tb + g) * hop;
        if (n >= 0 && n < N) acc += z[(size_t)g * N + n] * win[n];
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

size_t mask_istft_smem_bytes(int N, int hop, int FR, int G) {
  const int M = N / 2;
  return (size_t)N * sizeof(cpx) + (size_t)N * 4 + (size_t)FR * hop * 4 + (size_t)G * M * sizeof(cpx) +
         (size_t)G * (M + 1) * sizeof(cpx) + 16;
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
  // chunk of FR hops per CTA; G frames transformed together
  a.FR = 32;
  a.G = 4;
  size_t smem = mask_istft_smem_bytes(N, hop, a.FR, a.G);
  while (smem > 200 * 1024 && a.FR > 4) {
    a.FR /= 2;
    smem = mask_istft_smem_bytes(N, hop, a.FR, a.G);
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
