// Shared-memory inverse real FFT building blocks (Stockham autosort, radix-4 passes + one radix-2
// pass when log2(M) is odd).  Written as __host__ __device__ index math so that the exact same code
// is exercised on the CPU by tests/test_fft_host.py (compiled with g++ through a tiny C shim).
//
// Conventions (N = n_fft, M = N/2, tw[j] = exp(+2*pi*i*j/N) for j in [0, N)):
//   irfft:  x[n] = (1/N) * sum_{k<N} Xfull[k] exp(+2 pi i k n / N),  Xfull Hermitian-extended from X[0..M]
//   pack:   Z[k] = (X[k] + conj(X[M-k])) + i * tw[k] * (X[k] - conj(X[M-k])),  k in [0, M)
//   z = sum_k Z[k] exp(+2 pi i k m / M)  (UNNORMALISED inverse FFT of size M)
//   x[2m] = Re z[m] / N,  x[2m+1] = Im z[m] / N      (the 1/N is folded into the synthesis window)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LASS_HD __host__ __device__ __forceinline__
#else
#define LASS_HD inline
#endif

namespace lass {

struct cpx {
  float x, y;
};

LASS_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
LASS_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
LASS_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
LASS_HD cpx cmul_i(cpx a) { return cpx{-a.y, a.x}; }  // i * a

// Number of passes for a size-M (power of two, M >= 4) inverse FFT.
LASS_HD int fft_num_passes(int log2M) { return (log2M >> 1) + (log2M & 1); }

// One butterfly of pass `pass` (0-based) of the inverse Stockham FFT of size M = 1 << log2M.
// Radix-4 passes come first; if log2M is odd the LAST pass is radix-2.
//   src/dst: ping-pong buffers of M complex values; tw: N = 2M roots exp(+2 pi i j / N)
//   i: butterfly index in [0, M/4) for radix-4 passes, [0, M/2) for the radix-2 pass
// Returns nothing; writes 4 (or 2) outputs.
LASS_HD int fft_pass_is_radix2(int log2M, int pass) { return (log2M & 1) && (pass == (log2M >> 1)); }

LASS_HD int fft_pass_butterflies(int log2M, int pass) {
  return fft_pass_is_radix2(log2M, pass) ? (1 << (log2M - 1)) : (1 << (log2M - 2));
}

LASS_HD void ifft_butterfly(const cpx* src, cpx* dst, const cpx* tw, int log2M, int pass, int i) {
  const int M = 1 << log2M;
  const int t = 2 * pass;   // log2 of the stride s (every earlier pass was radix-4)
  const int s = 1 << t;     // stride
  const int n = M >> t;     // current sub-transform length
  const int q = i & (s - 1);
  const int p = i >> t;
  if (!fft_pass_is_radix2(log2M, pass)) {
    const int n1 = n >> 2;
    // twiddle exp(+2 pi i p / n) = tw[p * (N / n)] = tw[p << (t + 1)]
    const int j1 = p << (t + 1);
    const cpx w1 = tw[j1];
    const cpx w2 = tw[2 * j1];
    const cpx w3 = tw[3 * j1];
    const cpx a = src[q + s * (p)];
    const cpx b = src[q + s * (p + n1)];
    const cpx c = src[q + s * (p + 2 * n1)];
    const cpx d = src[q + s * (p + 3 * n1)];
    const cpx apc = cadd(a, c), amc = csub(a, c);
    const cpx bpd = cadd(b, d), jbmd = cmul_i(csub(b, d));
    dst[q + s * (4 * p + 0)] = cadd(apc, bpd);
    dst[q + s * (4 * p + 1)] = cmul(w1, cadd(amc, jbmd));
    dst[q + s * (4 * p + 2)] = cmul(w2, csub(apc, bpd));
    dst[q + s * (4 * p + 3)] = cmul(w3, csub(amc, jbmd));
  } else {
    // last pass, n == 2: p == 0, twiddle == 1
    const int m = n >> 1;
    const cpx a = src[q + s * (p)];
    const cpx b = src[q + s * (p + m)];
    dst[q + s * (2 * p + 0)] = cadd(a, b);
    dst[q + s * (2 * p + 1)] = cmul(tw[p << (t + 1)], csub(a, b));
  }
}

// Hermitian pack: Z[k] from the half spectrum X[0..M] (k in [0, M)).
LASS_HD cpx irfft_pack(const cpx* X, const cpx* tw, int M, int k) {
  cpx a = X[k];
  cpx b = X[M - k];
  if (k == 0) {
    // Im X[0] and Im X[N/2] do not contribute: the reference's inverse basis has sin(0) = sin(pi n) = 0
    a.y = 0.0f;
    b.y = 0.0f;
  }
  const cpx bc = cpx{b.x, -b.y};
  const cpx e = cadd(a, bc);
  const cpx o = cmul(tw[k], csub(a, bc));
  return cadd(e, cmul_i(o));
}

}  // namespace lass
