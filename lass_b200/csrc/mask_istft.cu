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
#include <stdlib.h>

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
  int T, F, N, log2M, hop, L, FR, B;
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

// =====================================================================================================================
// v2 (n_fft = 1024 / 2048): the warp's frame lives in REGISTERS.  M = n_fft/2 = 32 * E complex points, lane b holds
// Z[32 a + b], a < E (the coalesced load order of the bins).  Cooley-Tukey with k = 32 a + b, n = n1 + E n2:
//     T[b][n1]       = sum_a Z[32 a + b] W_E^(a n1)          E-point inverse DFT inside each lane (radix-4 x radix-4 [x 2])
//     U[b][n1]       = T[b][n1] W_M^(b n1)                    twiddle table tw2[n1][b] in shared memory
//     z[n1 + E n2]   = sum_b U[b][n1] W_32^(b n2)             32-point inverse DFT across lanes: ONE transposition through
//                                                             shared memory (row pitch 33: conflict-free), then again in
//                                                             registers (E = 32: one lane per n1; E = 16: two lanes per n1
//                                                             take the even / odd b and combine with 32 shuffles)
// against v1's five shared-memory Stockham passes.  Measured instruction count per frame: ~1.6 k warp instructions
// (v1: 5.6 k), no shared-memory bank conflicts.  Same CTA decomposition, overlap-add and normalisation as v1.
// =====================================================================================================================
__device__ __forceinline__ void r4_inv(float& ar, float& ai, float& br, float& bi, float& cr, float& ci, float& dr, float& di) {
  // (a, b, c, d) -> (a+b+c+d, a+ib-c-id, a-b+c-d, a-ib-c+id): 4-point inverse DFT, in place
  const float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
  const float s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
  ar = s0r + s2r; ai = s0i + s2i;
  cr = s0r - s2r; ci = s0i - s2i;
  br = s1r - s3i; bi = s1i + s3r;
  dr = s1r + s3i; di = s1i - s3r;
}

// y[n] = sum_{a<16} x[a] exp(+2 pi i a n / 16); x at re[OFF + STR * a]; result y[n] written to (yr[n], yi[n])
template <int OFF, int STR, int NREG>
__device__ __forceinline__ void fft16_inv(float (&re)[NREG], float (&im)[NREG], float (&yr)[16], float (&yi)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, c2 = 0.70710678118654752f;
#pragma unroll
  for (int a0 = 0; a0 < 4; ++a0)
    r4_inv(re[OFF + STR * a0], im[OFF + STR * a0], re[OFF + STR * (4 + a0)], im[OFF + STR * (4 + a0)],
           re[OFF + STR * (8 + a0)], im[OFF + STR * (8 + a0)], re[OFF + STR * (12 + a0)], im[OFF + STR * (12 + a0)]);
  // now element 4 q + a0 holds P[a0][q]; twiddle W16^(a0 q)
  auto rot = [&](int idx, float c, float s) {
    float& xr = re[OFF + STR * idx];
    float& xi = im[OFF + STR * idx];
    const float tr = xr * c - xi * s, ti = xr * s + xi * c;
    xr = tr;
    xi = ti;
  };
  rot(4 * 1 + 1, c1, s1);     // W16^1
  rot(4 * 2 + 1, c2, c2);     // W16^2
  rot(4 * 3 + 1, s1, c1);     // W16^3
  rot(4 * 1 + 2, c2, c2);     // W16^2
  {                           // W16^4 = i
    float& xr = re[OFF + STR * (4 * 2 + 2)];
    float& xi = im[OFF + STR * (4 * 2 + 2)];
    const float t = xr;
    xr = -xi;
    xi = t;
  }
  rot(4 * 3 + 2, -c2, c2);    // W16^6
  rot(4 * 1 + 3, s1, c1);     // W16^3
  rot(4 * 2 + 3, -c2, c2);    // W16^6
  rot(4 * 3 + 3, -c1, -s1);   // W16^9
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    r4_inv(re[OFF + STR * (4 * q + 0)], im[OFF + STR * (4 * q + 0)], re[OFF + STR * (4 * q + 1)], im[OFF + STR * (4 * q + 1)],
           re[OFF + STR * (4 * q + 2)], im[OFF + STR * (4 * q + 2)], re[OFF + STR * (4 * q + 3)], im[OFF + STR * (4 * q + 3)]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      yr[q + 4 * r] = re[OFF + STR * (4 * q + r)];
      yi[q + 4 * r] = im[OFF + STR * (4 * q + r)];
    }
  }
}

// E-point inverse DFT of (re, im)[0..E) in place (natural order in, natural order out)
template <int E>
__device__ __forceinline__ void fftE_inv(float (&re)[E], float (&im)[E]) {
  if constexpr (E == 16) {
    float yr[16], yi[16];
    fft16_inv<0, 1, 16>(re, im, yr, yi);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      re[n] = yr[n];
      im[n] = yi[n];
    }
  } else {
    // radix-2: even / odd inputs -> two 16-point transforms, y[n] = F0[n] + W32^n F1[n], y[n + 16] = F0[n] - W32^n F1[n]
    float f0r[16], f0i[16], f1r[16], f1i[16];
    fft16_inv<0, 2, 32>(re, im, f0r, f0i);
    fft16_inv<1, 2, 32>(re, im, f1r, f1i);
    constexpr float kc[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                              0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
                              -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                              -0.92387953251128674f, -0.98078528040323043f};
    constexpr float ks[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
                              0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f,
                              0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                              0.38268343236508977f, 0.19509032201612825f};
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float tr = f1r[n] * kc[n] - f1i[n] * ks[n], ti = f1r[n] * ks[n] + f1i[n] * kc[n];
      re[n] = f0r[n] + tr;
      im[n] = f0i[n] + ti;
      re[n + 16] = f0r[n] - tr;
      im[n + 16] = f0i[n] - ti;
    }
  }
}

// masked spectrum value of one bin (models/resunet.py:469-495; torchlibrosa magphase clamp 1e-10)
__device__ __forceinline__ cpx mask_bin(float x0, float x1, float x2, float sp, float cs, float sn) {
  // tanh(x) = sgn(x) (1 - t) / (1 + t), t = exp(-2 |x|).  The unit phasor (mc, ms) = (mr, mi) / |m| only needs the RATIO
  // of the two tanh values: (p, q) = (s1 (1 - t1)(1 + t2), s2 (1 - t2)(1 + t1)) is (mr, mi) scaled by (1 + t1)(1 + t2) in [1, 4].
  const float t1 = exp2f(-2.8853900817779268f * fabsf(x1)), t2 = exp2f(-2.8853900817779268f * fabsf(x2));
  const float pm = copysignf((1.0f - t1) * (1.0f + t2), x1), qm = copysignf((1.0f - t2) * (1.0f + t1), x2);
  const float inv = fminf(rsqrtf(pm * pm + qm * qm), 1e10f);      // |m| clamp: exact zeros (padded Nyquist column) give 0
  const float mc = pm * inv, ms = qm * inv;
  const float sig = __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x0));
  const float om = fmaxf(sp * sig, 0.0f);
  return cpx{om * (cs * mc - sn * ms), om * (sn * mc + cs * ms)};
}

// NW warps per CTA: 8 with two CTAs per SM (n_fft 1024), 16 with one (n_fft 2048: the per-warp buffers leave a pair of CTAs only
// 24 hops of output each, i.e. 27 % recomputed halo frames; one CTA of 16 warps holds 64 hops, 10 %)
template <int E, int NW>
__global__ void __launch_bounds__(NW * 32, 16 / NW) mask_istft_v2_kernel(const MaskIstftArgs a) {
  constexpr int kWarps = NW, kThreads = NW * 32;
  constexpr int M = 32 * E, N = 2 * M;
  constexpr int kPitch = 33;                      // transposition rows of 32 lanes, padded
  constexpr int kBuf = kPitch * E;                // complex slots per warp buffer (>= M + 1)
  // Loads are software-pipelined in batches of kLoadBatch bins per lane (6 planes each): batch i + 1 -- for the last batch of
  // a frame: batch 0 of the warp's NEXT frame, which then flies during the whole FFT / overlap-add of this one -- is issued
  // before batch i is consumed (ncu r2c: 29 % of the warp samples sat on the first use of a just-issued load).
  constexpr int kLoadBatch = 4;
  constexpr bool kCrossFrame = (E == 16);         // E = 32: the 24 prefetch registers do not fit next to the 64 of the frame
  constexpr bool kTables = (E == 16);             // Hermitian-pack twiddles + synthesis window in shared memory (E = 32: no room)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = a.hop;
  const int S = a.FR * hop;
  const int SO = S + N - hop;                     // the CTA's samples: its range + the spill of its last frames
  cpx* tw2 = reinterpret_cast<cpx*>(smem_raw);    // [E][32]: exp(+2 pi i b n1 / M)
  cpx* twh = tw2 + E * 32;                        // [M]: exp(+2 pi i k / N) (kTables)
  float2* wins = reinterpret_cast<float2*>(twh + (kTables ? M : 0));   // [M] window pairs (kTables)
  float* ola = reinterpret_cast<float*>(wins + (kTables ? M : 0));     // SO (a multiple of 8)
  float* wsum = ola + SO;                                  // hop: 1 / (N * window-sum) of the interior, by position mod hop
  cpx* bufs = reinterpret_cast<cpx*>(wsum + hop);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cpx* buf = bufs + (size_t)warp * kBuf;
  const float2* win2 = reinterpret_cast<const float2*>(a.window);

  for (int i = tid; i < E * 32; i += kThreads) {
    const float2 w = __ldg(reinterpret_cast<const float2*>(a.tw) + ((2 * (i & 31) * (i >> 5)) & (N - 1)));
    tw2[i] = cpx{w.x, w.y};
  }
  if (kTables)
    for (int i = tid; i < M; i += kThreads) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(a.tw) + i);
      twh[i] = cpx{w.x, w.y};
      wins[i] = __ldg(win2 + i);
    }
  const float invN = 1.0f / (float)N;
  for (int r = tid; r < hop; r += kThreads) {
    float ws = 0.0f;
    for (int n = r; n < N; n += hop) {
      const float w = __ldg(a.window + n);
      ws += w * w;
    }
    wsum[r] = invN / fmaxf(ws, 1e-11f);
  }
  const uint32_t hop_magic = (uint32_t)((0x100000000ull + (uint32_t)hop - 1) / (uint32_t)hop);   // x / hop == umulhi(x, magic) for x * hop < 2^32
  // programmatic dependent launch: the tables above are constants of the model; everything below waits for the previous launches
  griddep_launch_dependents();
  griddep_wait();
  // Persistent CTAs over (clip, chunk of FR frames) items, item i on CTA i mod grid: FR is a constant of the shape (NOT of the
  // batch size: clips come out bit-identical whatever batch they run in), small enough that the static round-robin ends level.
  // An item owns the FR frames that START in [pos0, pos0 + S) -- no halo frames are recomputed.  Their samples reach N - hop past
  // the range: that spill is added to the output with red.global, like the first N - hop samples of the range (which the previous
  // chunk spills into); everything else is stored.  The output is zeroed before the launch, and a sample receives at most two such
  // additions (S >= N - hop), so the result does not depend on their order: 0 + x + y == 0 + y + x in floating point.
  const int chunks = (a.T + a.FR - 1) / a.FR, n_items = chunks * a.B;
  const long long interior_hi = (long long)a.T * hop;
  const bool vec_ok = (a.L & 3) == 0;
  // first frame of this warp in round r of an item (-1: none)
  auto item_frame = [&](int item, int r, int& bb) -> long long {
    if (item >= n_items) return -1;
    bb = item / chunks;
    const long long lo = (long long)(item - bb * chunks) * a.FR;
    long long hi = lo + a.FR - 1;
    if (hi > a.T - 1) hi = a.T - 1;
    const long long t = lo + (long long)r * kWarps + warp;
    return t <= hi ? t : -1;
  };

  struct Batch {
    float x0[kLoadBatch], x1[kLoadBatch], x2[kLoadBatch], sp[kLoadBatch], cs[kLoadBatch], sn[kLoadBatch];
  };
  // bins 32 (a0 + u) + lane of frame t; batch 0 also brings the Nyquist bin k = M (lane 0) in slot `ny`
  float ny[6];
  auto load_batch = [&](Batch& q, int bb, long long t, int a0) {
    const float* fp = a.feat + (size_t)bb * a.feat_bstride + (size_t)t * a.feat_tstride;
    const size_t o = ((size_t)bb * a.T + (size_t)t) * a.F;
    const float *magb = a.mag, *cosb = a.cosp, *sinb = a.sinp;
#pragma unroll
    for (int u = 0; u < kLoadBatch; ++u) {
      const int k = 32 * (a0 + u) + lane;
      const bool inX = k < a.feat_F;
      q.x0[u] = inX ? __ldg(fp + k) : 0.0f;
      q.x1[u] = inX ? __ldg(fp + a.feat_cstride + k) : 0.0f;
      q.x2[u] = inX ? __ldg(fp + 2 * a.feat_cstride + k) : 0.0f;
      q.sp[u] = __ldg(magb + o + k);
      q.cs[u] = __ldg(cosb + o + k);
      q.sn[u] = __ldg(sinb + o + k);
    }
    if (a0 == 0 && lane == 0) {
      const bool inX = M < a.feat_F;
      ny[0] = inX ? __ldg(fp + M) : 0.0f;
      ny[1] = inX ? __ldg(fp + a.feat_cstride + M) : 0.0f;
      ny[2] = inX ? __ldg(fp + 2 * a.feat_cstride + M) : 0.0f;
      ny[3] = __ldg(magb + o + M);
      ny[4] = __ldg(cosb + o + M);
      ny[5] = __ldg(sinb + o + M);
    }
  };
  Batch pre;
  long long pre_tag = -1;          // clip * T + frame whose batch 0 `pre` holds (a warp without a frame in some item breaks the chain)
  if (kCrossFrame) {
    int bb;
    const long long t = item_frame(blockIdx.x, 0, bb);
    if (t >= 0) {
      load_batch(pre, bb, t, 0);
      pre_tag = (long long)bb * a.T + t;
    }
  }

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int b = item / chunks;
  const long long t_lo = (long long)(item - b * chunks) * a.FR;
  long long t_hi = t_lo + a.FR - 1;
  if (t_hi > a.T - 1) t_hi = a.T - 1;
  const long long pos0 = t_lo * hop;
  for (int i = tid; i < SO; i += kThreads) ola[i] = 0.0f;
  __syncthreads();
  for (long long tb = t_lo; tb <= t_hi; tb += kWarps) {
    const long long t = tb + warp;
    if (t <= t_hi) {
      float zr[E], zi[E];
      // ---- 1. masked half spectrum X[32 a + lane]; bins >= feat_F behave as x = 0, bins >= F (never for k < M) as mag = 0 ----
      if (!kCrossFrame || pre_tag != (long long)b * a.T + t) load_batch(pre, b, t, 0);
      if (lane == 0) buf[M] = mask_bin(ny[0], ny[1], ny[2], ny[3], ny[4], ny[5]);     // Nyquist bin k = M
#pragma unroll
      for (int a0 = 0; a0 < E; a0 += kLoadBatch) {
        const Batch cur = pre;
        if (a0 + kLoadBatch < E) {
          load_batch(pre, b, t, a0 + kLoadBatch);
        } else if (kCrossFrame) {
          // this warp's next frame: next round of the item, else round 0 of the CTA's next item
          int bb = b;
          const long long tn = t + kWarps <= t_hi ? t + kWarps : item_frame(item + (int)gridDim.x, 0, bb);
          if (tn >= 0) {
            load_batch(pre, bb, tn, 0);
            pre_tag = (long long)bb * a.T + tn;
          }
        }
#pragma unroll
        for (int u = 0; u < kLoadBatch; ++u) {
          const cpx X = mask_bin(cur.x0[u], cur.x1[u], cur.x2[u], cur.sp[u], cur.cs[u], cur.sn[u]);
          zr[a0 + u] = X.x;
          zi[a0 + u] = X.y;
          buf[32 * (a0 + u) + lane] = X;
        }
      }
      __syncwarp();
      // ---- 2. Hermitian pack Z[k] = (X[k] + conj X[M-k]) + i tw[k] (X[k] - conj X[M-k]) ----
#pragma unroll
      for (int aa = 0; aa < E; ++aa) {
        const int k = 32 * aa + lane;
        cpx xa = cpx{zr[aa], zi[aa]};
        cpx xb = buf[M - k];
        if (k == 0) {        // Im X[0], Im X[M] do not contribute (see irfft_pack)
          xa.y = 0.0f;
          xb.y = 0.0f;
        }
        float2 w;
        if (kTables) {
          const cpx wc = twh[k];
          w = make_float2(wc.x, wc.y);
        } else {
          w = __ldg(reinterpret_cast<const float2*>(a.tw) + k);
        }
        const float er = xa.x + xb.x, ei = xa.y - xb.y;      // X[k] + conj(X[M-k])
        const float dr = xa.x - xb.x, di = xa.y + xb.y;      // X[k] - conj(X[M-k])
        const float orr = w.x * dr - w.y * di, oi = w.x * di + w.y * dr;
        zr[aa] = er - oi;                                     // e + i o
        zi[aa] = ei + orr;
      }
      // ---- 3. E-point inverse DFT inside the lane, twiddle, transposition ----
      fftE_inv<E>(zr, zi);
      __syncwarp();                                           // all reads of X[M-k] done before buf is overwritten
#pragma unroll
      for (int n1 = 0; n1 < E; ++n1) {
        const cpx w = tw2[n1 * 32 + lane];
        buf[n1 * kPitch + lane] = cpx{zr[n1] * w.x - zi[n1] * w.y, zr[n1] * w.y + zi[n1] * w.x};
      }
      __syncwarp();
      // ---- 4. 32-point inverse DFT across the former lanes, window, result to the warp buffer (interleaved samples) ----
      if constexpr (E == 32) {
#pragma unroll
        for (int bb = 0; bb < 32; ++bb) {
          const cpx v = buf[lane * kPitch + bb];
          zr[bb] = v.x;
          zi[bb] = v.y;
        }
        fftE_inv<32>(zr, zi);
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) {
          const int m = lane + 32 * n2;
          const float2 w = kTables ? wins[m] : __ldg(win2 + m);
          buf[m] = cpx{zr[n2] * w.x, zi[n2] * w.y};
        }
      } else {
        const int n1 = lane & 15, h = lane >> 4;
        float fr[16], fi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const cpx v = buf[n1 * kPitch + 2 * j + h];
          zr[j] = v.x;
          zi[j] = v.y;
        }
        fft16_inv<0, 1, 16>(zr, zi, fr, fi);
        __syncwarp();
        constexpr float kc[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                                  0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
                                  -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                                  -0.92387953251128674f, -0.98078528040323043f};
        constexpr float ks[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
                                  0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f,
                                  0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                                  0.38268343236508977f, 0.19509032201612825f};
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
          // lanes with h = 1 hold F1: pre-multiply by W32^n2; then z = F0 + W F1 (h = 0), F0 - W F1 (h = 1)
          const float gr = h ? fr[n2] * kc[n2] - fi[n2] * ks[n2] : fr[n2];
          const float gi = h ? fr[n2] * ks[n2] + fi[n2] * kc[n2] : fi[n2];
          const float pr = __shfl_xor_sync(0xffffffffu, gr, 16), pi = __shfl_xor_sync(0xffffffffu, gi, 16);
          const float vr = h ? pr - gr : gr + pr, vi = h ? pi - gi : gi + pi;
          const int m = n1 + 16 * (n2 + 16 * h);
          const float2 w = kTables ? wins[m] : __ldg(win2 + m);
          buf[m] = cpx{vr * w.x, vi * w.y};
        }
      }
    }
    __syncthreads();
    // ---- 5. overlap-add of this round's (already windowed) frames; each thread owns its positions: deterministic ----
    const int cnt = (int)((t_hi - tb + 1 < kWarps) ? (t_hi - tb + 1) : kWarps);
    const long long w0 = tb * hop - pos0;
    const long long w1 = (tb + cnt - 1) * hop + N - pos0;       // <= SO
    const float* zbase = reinterpret_cast<const float*>(bufs);
    constexpr int zstride = 2 * kBuf;
    // four consecutive samples per thread and step: hop, N and every frame offset are multiples of 4
    for (int pos = (int)w0 + 4 * tid; pos < (int)w1; pos += 4 * kThreads) {
      float4 acc = *reinterpret_cast<const float4*>(ola + pos);
      const int rel = (int)(pos0 + pos - tb * hop);       // offset from frame tb's start, >= 0
      int g_hi = (int)__umulhi((uint32_t)rel, hop_magic);
      if (g_hi > cnt - 1) g_hi = cnt - 1;
      const int g_lo = rel < N ? 0 : (int)__umulhi((uint32_t)(rel - N), hop_magic) + 1;
      for (int g = g_lo; g <= g_hi; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(zbase + (size_t)g * zstride + (rel - g * hop));
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      *reinterpret_cast<float4*>(ola + pos) = acc;
    }
    __syncthreads();
  }

  // ---- 6. window-sum normalisation (folded hann^2, clamp 1e-11), 1/N, trim n_fft/2 ----
  auto edge_scale = [&](long long np) {       // positions some frame that would cover them does not exist for
    long long ta = (np - N) / hop + 1;
    if (np - N < 0) ta = 0;
    long long tz = np / hop;
    if (tz > a.T - 1) tz = a.T - 1;
    float ws = 0.0f;
    for (long long t = ta; t <= tz; ++t) {
      const float w = __ldg(a.window + (np - t * hop));
      ws += w * w;
    }
    return invN / fmaxf(ws, 1e-11f);
  };
  float* const outb = a.out + (size_t)b * a.L;
  // four samples per thread: pos0, hop and N / 2 are multiples of 4, so a quad lies in one hop and (for L % 4 == 0) is one
  // aligned 16-byte store
  for (int pos = 4 * tid; pos < SO; pos += 4 * kThreads) {
    const long long np = pos0 + pos;
    const long long no = np - (N >> 1);
    if (no + 3 < 0 || no >= a.L) continue;
    const bool shared_with_neighbour = pos < N - hop || pos >= S;
    const float4 v = *reinterpret_cast<const float4*>(ola + pos);
    float4 sc;
    if (np >= N - hop && np + 3 < interior_hi) {
      // every frame that can cover these samples exists: the scale depends on the position modulo hop only
      const uint32_t q = __umulhi((uint32_t)pos, hop_magic);
      sc = *reinterpret_cast<const float4*>(wsum + (pos - (int)q * hop));        // pos0 is a multiple of hop
    } else {
      sc = make_float4(edge_scale(np), edge_scale(np + 1), edge_scale(np + 2), edge_scale(np + 3));
    }
    const float4 r = make_float4(v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w);
    const float rr[4] = {r.x, r.y, r.z, r.w};
    if (shared_with_neighbour) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (no + i >= 0 && no + i < a.L) atomicAdd(outb + no + i, rr[i]);
    } else if (vec_ok && no >= 0 && no + 3 < a.L) {
      *reinterpret_cast<float4*>(outb + no) = r;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (no + i >= 0 && no + i < a.L) outb[no + i] = rr[i];
    }
  }
  __syncthreads();      // the accumulator is re-zeroed for the next item
  }
}

template <int E, int NW>
cudaError_t launch_v2(MaskIstftArgs a, int B, int T, cudaStream_t stream) {
  constexpr int N = 64 * E;
  const size_t fixed = (size_t)E * 32 * sizeof(cpx) + (E == 16 ? (size_t)32 * E * (sizeof(cpx) + sizeof(float2)) : 0) +
                       (size_t)NW * 33 * E * sizeof(cpx) + 16;
  // item = 32 frames (4 rounds of 8 warps / 2 of 16), more where the spill rule S >= N - hop wants it
  a.FR = 32;
  while (a.FR * a.hop < N - a.hop) a.FR += NW;
  const size_t smem = fixed + (size_t)(a.FR * a.hop + N) * 4;      // ola: FR hops + the spill (N - hop), + the scale table (hop)
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  {  // per-device attribute, cheap: opted in to the full 227 KiB on every launch for the current device (no process-wide cache)
    cudaError_t e = cudaFuncSetAttribute(mask_istft_v2_kernel<E, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  a.B = B;
  const long long n_items = (long long)((T + a.FR - 1) / a.FR) * B;
  const long long slots = (long long)device_sm_count() * (16 / NW);
  dim3 grid((unsigned)(n_items < slots ? n_items : slots));
  {
    cudaError_t e = cudaMemsetAsync(a.out, 0, (size_t)B * a.L * sizeof(float), stream);    // the CTAs ADD their boundary samples
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(mask_istft_v2_kernel<E, NW>, grid, dim3(NW * 32), smem, stream, a);
}

}  // namespace

// debug switch (lass_debug_set_istft_v1): the shared-memory Stockham kernel for every n_fft
static int g_force_v1 = 0;
void mask_istft_force_v1(int on) { g_force_v1 = on ? 1 : 0; }

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
  a.FR = 0;
  // register-FFT kernel for the two model shapes (hop must keep the window-table reads 8-byte aligned: always true)
  if (!g_force_v1) {
    if (N == 1024) return launch_v2<16, 8>(a, B, T, stream);
    if (N == 2048) return launch_v2<32, 16>(a, B, T, stream);
  }
  // CTA = FR hops of output; as large as fits two CTAs per SM (halo frames are recomputed: ~n_fft/hop per CTA)
  a.FR = 64;
  size_t smem = mask_istft_smem_bytes(N, hop, a.FR);
  while (smem > 110 * 1024 && a.FR > 8) {
    a.FR -= 8;
    smem = mask_istft_smem_bytes(N, hop, a.FR);
  }
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  {
    cudaError_t e = cudaFuncSetAttribute(mask_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  const long long P = (long long)(T - 1) * hop + N;
  const int S = a.FR * hop;
  dim3 grid((unsigned)((P + S - 1) / S), (unsigned)B);
  mask_istft_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace lass
