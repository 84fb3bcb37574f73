// C-ABI layer: argument validation, error reporting, TMA tensor-map construction.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "conv.cuh"
#include "lass_internal.cuh"

namespace lass {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int set_cuda_error(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int device_sm_count() {
  static std::mutex mu;
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(LASS_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
  if (elem_bytes != 2) return set_error(LASS_ERR_ARG, "tensor maps are built for 16-bit elements only");
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  // bf16 and fp16 tiles move identically; the data type only matters for OOB fill (zeros here)
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(LASS_ERR_DRIVER,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] strides "
                     "[%llu %llu %llu] box [%u %u %u %u] base %p",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                     (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                     (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
  }
  return 0;
}

}  // namespace lass

using namespace lass;

extern "C" {

int lass_version(void) { return LASS_B200_VERSION; }

const char* lass_last_error(void) { return g_err; }

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
