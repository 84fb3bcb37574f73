// Internal (non-ABI) declarations shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lass_b200.h"

namespace lass {

// ---- error reporting (thread-local message behind lass_last_error()) ----
int set_error(int code, const char* fmt, ...);
int set_cuda_error(cudaError_t e, const char* what);

// ---- TMA tensor maps (cuTensorMapEncodeTiled resolved at run time: no link-time libcuda dependency) ----
// dims / box in elements (innermost first); strides in bytes for dims 1..rank-1.
int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// ---- K1 stft ----
int stft_num_ntiles(int n_fft);
size_t stft_padded_len(int L, int n_fft, int hop);
size_t stft_workspace_bytes(int B, int L, int n_fft, int hop);
int launch_stft_multi(const float* wave, int B, int L, int hop, int nres, const int* n_ffts, const void* const* basis_hi,
                      const void* const* basis_lo, float* const* mag, float* const* cosp, float* const* sinp, int precision_mode,
                      int magphase_mode, void* workspace, cudaStream_t stream, const float* adjoint_window);
int launch_stft(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                float* mag, float* cosp, float* sinp, int precision_mode, int magphase_mode, void* workspace,
                cudaStream_t stream, const float* adjoint_window = nullptr);

// SM count of the CURRENT device (cached per device ordinal; thread-safe)
int device_sm_count();

// ---- K5 mask + istft ----
cudaError_t launch_mask_istft(const float* feat, long long feat_bstride, long long feat_cstride, int feat_tstride,
                              int feat_F, const float* mag, const float* cosp, const float* sinp,
                              const float* window, const float* tw, float* out, int B, int T, int F, int N,
                              int hop, int L, cudaStream_t stream);

void mask_istft_force_v1(int on);

// ---- programmatic dependent launch ----
// launch_pdl(kernel, grid, block, smem, stream, args...) launches with cudaLaunchAttributeProgrammaticStreamSerialization: the
// grid may be SCHEDULED while its predecessor in the stream still runs (once all of the predecessor's CTAs have executed
// griddepcontrol.launch_dependents or exited), so launch latency and prologues overlap the predecessor's tail.  Every kernel
// launched this way starts with griddep_launch_dependents(); griddep_wait(); -- nothing it reads or writes is touched before
// every earlier launch of the stream has completed and its memory is visible.
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace lass
