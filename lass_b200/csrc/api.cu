// C-ABI layer: argument validation of the entry points (error reporting / tensor maps: common.cu).
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "conv.cuh"
#include "lass_internal.cuh"

using namespace lass;

extern "C" {

int lass_stft_basis_rows(int n_fft) { return stft_num_ntiles(n_fft) * 256; }

size_t lass_stft_workspace_bytes(int B, int L, int n_fft, int hop) {
  if (B <= 0 || L <= 0 || n_fft <= 0 || hop <= 0) return 0;
  return stft_workspace_bytes(B, L, n_fft, hop);
}

int lass_stft_fwd(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                  float* mag, float* cos, float* sin, int precision_mode, int magphase_mode, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (!wave || !basis_hi || !mag || !cos || !sin || !workspace)
    return set_error(LASS_ERR_ARG, "lass_stft_fwd: null pointer");
  if (precision_mode == 0 && !basis_lo) return set_error(LASS_ERR_ARG, "lass_stft_fwd: basis_lo required in mode 0");
  if (B <= 0 || n_fft < 64 || (n_fft & (n_fft - 1)) || hop <= 0 || hop % 8 || L <= n_fft / 2)
    return set_error(LASS_ERR_ARG, "lass_stft_fwd: need power-of-two n_fft >= 64, hop %% 8 == 0, L > n_fft/2 (got B=%d L=%d n_fft=%d hop=%d)",
                     B, L, n_fft, hop);
  if (workspace_bytes < stft_workspace_bytes(B, L, n_fft, hop))
    return set_error(LASS_ERR_WORKSPACE, "lass_stft_fwd: workspace %zu < %zu", workspace_bytes,
                     stft_workspace_bytes(B, L, n_fft, hop));
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return set_error(LASS_ERR_ARG, "lass_stft_fwd: workspace not 256 B aligned");
  return launch_stft(wave, B, L, n_fft, hop, basis_hi, precision_mode == 0 ? basis_lo : basis_hi, mag, cos, sin,
                     precision_mode, magphase_mode, workspace, (cudaStream_t)stream);
}

int lass_stft_multi_fwd(const float* wave, int B, int L, int hop, int nres, const int* n_ffts, const void* const* basis_hi,
                        const void* const* basis_lo, float* const* mag, float* const* cos, float* const* sin, int precision_mode,
                        int magphase_mode, void* workspace, size_t workspace_bytes, void* stream) {
  if (!wave || !n_ffts || !basis_hi || !basis_lo || !mag || !cos || !sin || !workspace)
    return set_error(LASS_ERR_ARG, "lass_stft_multi_fwd: null pointer");
  if (nres < 1 || nres > 3 || B <= 0 || hop <= 0 || hop % 8)
    return set_error(LASS_ERR_ARG, "lass_stft_multi_fwd: need 1..3 resolutions, hop %% 8 == 0 (got nres=%d hop=%d)", nres, hop);
  size_t need = 0;
  for (int r = 0; r < nres; ++r) {
    const int n_fft = n_ffts[r];
    if (n_fft < 64 || (n_fft & (n_fft - 1)) || L <= n_fft / 2 || !basis_hi[r] || !mag[r] || !cos[r] || !sin[r] ||
        (precision_mode == 0 && !basis_lo[r]))
      return set_error(LASS_ERR_ARG, "lass_stft_multi_fwd: bad resolution %d (n_fft=%d, L=%d)", r, n_fft, L);
    need += stft_workspace_bytes(B, L, n_fft, hop);
  }
  if (workspace_bytes < need) return set_error(LASS_ERR_WORKSPACE, "lass_stft_multi_fwd: workspace %zu < %zu", workspace_bytes, need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return set_error(LASS_ERR_ARG, "lass_stft_multi_fwd: workspace not 256 B aligned");
  int e = launch_stft_multi(wave, B, L, hop, nres, n_ffts, basis_hi, precision_mode == 0 ? basis_lo : basis_hi, mag, cos, sin,
                            precision_mode, magphase_mode, workspace, (cudaStream_t)stream, nullptr);
  if (e == LASS_ERR_ARG) return set_error(e, "lass_stft_multi_fwd: unsupported geometry (n_fft %% 256 == 0 required)");
  return e;
}

int lass_mask_istft(const float* feat3, long long feat_bstride, long long feat_cstride, int feat_tstride,
                    int feat_F, const float* mag, const float* cos, const float* sin, const float* window,
                    const float* twiddle, int B, int T, int F, int n_fft, int hop, int L, float* wave_out,
                    void* stream) {
  if (!feat3 || !mag || !cos || !sin || !window || !twiddle || !wave_out)
    return set_error(LASS_ERR_ARG, "lass_mask_istft: null pointer");
  if (B <= 0 || T <= 0 || n_fft < 16 || (n_fft & (n_fft - 1)) || F != n_fft / 2 + 1 || hop <= 0 || L <= 0 ||
      feat_F < 0 || feat_F > F)
    return set_error(LASS_ERR_ARG, "lass_mask_istft: bad shape B=%d T=%d F=%d n_fft=%d hop=%d L=%d feat_F=%d", B, T, F,
                     n_fft, hop, L, feat_F);
  if ((long long)(T - 1) * hop + n_fft < (long long)n_fft / 2 + L)
    return set_error(LASS_ERR_ARG, "lass_mask_istft: %d frames cannot cover %d samples", T, L);
  return set_cuda_error(launch_mask_istft(feat3, feat_bstride, feat_cstride, feat_tstride, feat_F, mag, cos, sin,
                                          window, twiddle, wave_out, B, T, F, n_fft, hop, L, (cudaStream_t)stream),
                        "mask_istft launch");
}

int lass_conv_igemm(const lass_conv_desc* desc_host, void* stream) {
  if (!desc_host) return set_error(LASS_ERR_ARG, "lass_conv_igemm: null descriptor");
  ConvPrepared* cp = nullptr;
  int e = conv_prepare(*desc_host, &cp);
  if (e) return e;
  e = conv_run(cp, (cudaStream_t)stream);
  conv_free(cp);
  return e;
}

int lass_conv_prepare(const lass_conv_desc* desc_host, lass_conv** out) {
  if (!desc_host || !out) return set_error(LASS_ERR_ARG, "lass_conv_prepare: null pointer");
  ConvPrepared* cp = nullptr;
  int e = conv_prepare(*desc_host, &cp);
  *out = reinterpret_cast<lass_conv*>(cp);
  return e;
}

int lass_conv_run(const lass_conv* conv, void* stream) {
  if (!conv) return set_error(LASS_ERR_ARG, "lass_conv_run: null handle");
  return conv_run(reinterpret_cast<const ConvPrepared*>(conv), (cudaStream_t)stream);
}

void lass_conv_destroy(lass_conv* conv) {
  if (conv) conv_free(reinterpret_cast<ConvPrepared*>(conv));
}

int lass_istft_bwd(const float* dwave, int B, int L, int n_fft, int hop, int T, const float* window, const void* basis_hi,
                   const void* basis_lo, float* dre, float* dim, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dwave || !window || !basis_hi || !basis_lo || !dre || !dim || !workspace)
    return set_error(LASS_ERR_ARG, "lass_istft_bwd: null pointer");
  if (B <= 0 || n_fft < 256 || (n_fft & (n_fft - 1)) || hop <= 0 || hop % 8 || L <= n_fft / 2 || T != L / hop + 1)
    return set_error(LASS_ERR_ARG, "lass_istft_bwd: bad shape B=%d L=%d n_fft=%d hop=%d T=%d", B, L, n_fft, hop, T);
  if (workspace_bytes < stft_workspace_bytes(B, L, n_fft, hop)) return set_error(LASS_ERR_WORKSPACE, "lass_istft_bwd: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return set_error(LASS_ERR_ARG, "lass_istft_bwd: workspace not 256 B aligned");
  return launch_stft(dwave, B, L, n_fft, hop, basis_hi, basis_lo, dre, nullptr, dim, 0, 2, workspace, (cudaStream_t)stream, window);
}

int lass_debug_set_conv_profile(long long* device_counters) {
#ifndef LASS_CONV_PROFILE
  if (device_counters)
    return set_error(LASS_ERR_ARG, "conv profile: this build has no role profiler (make prof, LASS_B200_LIB=.../liblass_b200_prof.so)");
#endif
  conv_set_profile_buffer(device_counters);
  return 0;
}

int lass_debug_set_istft_v1(int on) {
  mask_istft_force_v1(on);
  return 0;
}

int lass_debug_set_conv_flags(int flags) {
  conv_set_debug_flags(flags);
  return 0;
}

}  // extern "C"
