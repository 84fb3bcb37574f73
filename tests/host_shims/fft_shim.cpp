// Host build of lass_b200/csrc/fft.cuh: runs the exact index math of the shared-memory inverse real FFT
// used by the mask_istft kernel on the CPU (tests/test_fft_host.py).
#include <vector>
#include "../../lass_b200/csrc/fft.cuh"

extern "C" int lass_host_irfft(const float* X_ri, const float* tw_ri, int N, float* out) {
  using namespace lass;
  const int M = N / 2;
  int log2M = 0;
  while ((1 << log2M) < M) ++log2M;
  const cpx* X = reinterpret_cast<const cpx*>(X_ri);
  const cpx* tw = reinterpret_cast<const cpx*>(tw_ri);
  std::vector<cpx> a(M), b(M);
  for (int k = 0; k < M; ++k) a[k] = irfft_pack(X, tw, M, k);
  cpx* src = a.data();
  cpx* dst = b.data();
  const int passes = fft_num_passes(log2M);
  for (int p = 0; p < passes; ++p) {
    const int nb = fft_pass_butterflies(log2M, p);
    for (int i = 0; i < nb; ++i) ifft_butterfly(src, dst, tw, log2M, p, i);
    cpx* t = src; src = dst; dst = t;
  }
  const float* z = reinterpret_cast<const float*>(src);
  for (int n = 0; n < N; ++n) out[n] = z[n] / (float)N;
  return passes;
}
