/*
 * lass_b200 — C ABI of the B200-native (sm_100a) LASS/AudioSep separation hot path.
 *
 * The reference (reedrosenbluth/LASS) is pure Python/PyTorch: its "plugin boundary" for this path is the
 * nn.Module API `models.resunet.ResUNet30`.  The Python mirror of that API lives in
 * `lass_b200/models/resunet.py`; everything it executes on the GPU goes through the entry points below,
 * bound with ctypes (`lass_b200/_cabi.py`).  Plain pointers and sizes only — no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream);
 *   - no entry point allocates device memory or synchronises: callers pass workspaces, and every call is
 *     CUDA-graph capturable;
 *   - return value: 0 = success, <0 = LASS_ERR_* argument/state error, >0 = a `cudaError_t`;
 *     `lass_last_error()` returns a thread-local message for the last non-zero return;
 *   - audio layout (B, L) fp32; spectrogram planes (B, T, F) fp32 with F = n_fft/2+1 fastest — byte-identical
 *     to the reference's (B, 1, T, F) tensors.
 */
#ifndef LASS_B200_H_
#define LASS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LASS_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define LASS_API __attribute__((visibility("default")))
#else
#define LASS_API
#endif

#define LASS_OK 0
#define LASS_ERR_ARG (-1)      /* invalid argument (shape, alignment, null pointer) */
#define LASS_ERR_DRIVER (-2)   /* CUDA driver entry point unavailable / tensor-map encode failed */
#define LASS_ERR_STATE (-3)    /* plan used with mismatching shapes */
#define LASS_ERR_WORKSPACE (-4) /* workspace too small */

/* Library version (LASS_B200_VERSION of the build). */
LASS_API int lass_version(void);

/* Thread-local, NUL-terminated description of the last error returned on this thread. */
LASS_API const char* lass_last_error(void);

/* ------------------------------------------------------------------------------------------------------
 * K1  STFT front end  (replaces torchlibrosa 0.1.0 STFT.forward + Base.spectrogram_phase,
 *     reference models/base.py:83-88, constructed at models/resunet.py:284-292)
 *
 *   wave      (B, L) fp32
 *   basis_hi / basis_lo   bf16 (ntiles*128, n_fft): windowed DFT basis split hi/lo, 64-bin tiles with rows
 *             [0,64) = real basis and [64,128) = imaginary basis (see lass_stft_basis_rows; packed from the
 *             reference's frozen `stft.conv_real/conv_imag.weight` by lass_b200.packing.pack_stft_basis)
 *   mag, cos, sin   (B, T, F) fp32 out, T = L/hop + 1, F = n_fft/2 + 1
 *   precision_mode  0 = fp32-parity (3 bf16 MMAs per product, max rel. err ~5e-6), 1 = fast (single bf16 pass)
 *   workspace       >= lass_stft_workspace_bytes(B, L, n_fft, hop) bytes, 256-byte aligned
 * Requirements: n_fft % 64 == 0, hop % 8 == 0, L > n_fft/2 (reflect padding).
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_stft_basis_rows(int n_fft);
LASS_API size_t lass_stft_workspace_bytes(int B, int L, int n_fft, int hop);
LASS_API int lass_stft_fwd(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                  float* mag, float* cos, float* sin, int precision_mode, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K5  complex mask + inverse STFT  (replaces ResUNet30_Base.feature_maps_to_wav, reference
 *     models/resunet.py:436-519, and torchlibrosa 0.1.0 ISTFT.forward incl. window-sum normalisation)
 *
 *   feat3     3 planes of mask features (sigmoid-magnitude, tanh-real, tanh-imag), element (b, k, t, f) at
 *             feat3[b*feat_bstride + k*feat_cstride + t*feat_tstride + f], valid for f < feat_F; bins
 *             f >= feat_F behave as the zero-padded Nyquist column of models/resunet.py:573 (output bin = 0).
 *             The reference layout (B, 3, T, F) is bstride = 3*T*F, cstride = T*F, tstride = F, feat_F = F.
 *   mag, cos, sin   (B, T, F) fp32 mixture spectrogram / phase
 *   window    (n_fft) fp32 synthesis window (periodic Hann, unscaled)
 *   twiddle   (n_fft, 2) fp32: (cos, sin)(2*pi*j/n_fft)
 *   wave_out  (B, L) fp32
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_mask_istft(const float* feat3, long long feat_bstride, long long feat_cstride, int feat_tstride,
                    int feat_F, const float* mag, const float* cos, const float* sin, const float* window,
                    const float* twiddle, int B, int T, int F, int n_fft, int hop, int L, float* wave_out,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K3 / K4  implicit-GEMM convolution on tcgen05  (replaces conv2d 3x3 / 1x1 + conv_transpose2d(kernel = stride)
 *     of ConvBlockRes / EncoderBlockRes1B / DecoderBlockRes1B, reference models/resunet.py:147-165,186-198,
 *     240-264, with batch_norm + FiLM + leaky_relu folded into the producer's epilogue, the residual / shortcut
 *     as a second K-segment, avg_pool2d, torch.cat (channel-slice outputs) and after_conv (:570) fused in)
 *
 * Activations are NHWC 16-bit (bf16, or fp16 for the raw residual stream).  All structs are HOST memory.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct lass_conv_segment {
  const void* src;     /* (B, H, W, src_cstride) 16-bit; channels [src_coff, src_coff + cin) are read            */
  int src_cstride;     /* channels per pixel of the source buffer (multiple of 8)                                 */
  int src_coff;        /* first channel read (multiple of 8)                                                      */
  int cin;             /* channels of this K-segment (multiple of kc)                                             */
  int kc;              /* channels per K-chunk: 32 or 64                                                          */
  int taps;            /* 9 = 3x3 conv with zero padding 1 (tap = ky*3 + kx), 1 = 1x1                             */
  int fp16;            /* 0 = bf16 source and weights, 1 = fp16 source and weights                                */
  const void* weights; /* (taps, ncols, cin) 16-bit                                                               */
} lass_conv_segment;

typedef struct lass_conv_out {
  void* ptr;           /* (B, Ho, Wo, cstride) 16-bit NHWC; NULL = output disabled                                */
  int cstride;         /* channels per pixel of the destination buffer                                            */
  int coff;            /* first channel written                                                                   */
  int fp16;            /* 1 = saturating fp16, 0 = bf16                                                           */
  const float* scale;  /* NULL: store the value; else store leaky_relu(scale[c]*v + shift[b*shift_bstride + c])    */
  const float* shift;
  int shift_bstride;
} lass_conv_out;

typedef struct lass_conv_desc {
  int B, H, W;         /* input grid                                                                              */
  int ncols;           /* GEMM N: Cout, or up_h*up_w*Cout for a transposed conv (column = (dy, dx, c))            */
  int nseg;            /* 1 or 2 K-segments accumulated into the same output                                      */
  lass_conv_segment seg[2];
  const float* bias;   /* (ncols) fp32 or NULL                                                                    */
  int up_h, up_w;      /* transposed conv stride (1 or 2); output pixel (h*up_h + dy, w*up_w + dx)                */
  int group_c;         /* Cout (= ncols / (up_h*up_w))                                                            */
  lass_conv_out full_raw, full_act;  /* outputs on the (H*up_h, W*up_w) grid                                      */
  int pool_h, pool_w;  /* average pooling window (1 or 2) for the pooled outputs                                  */
  lass_conv_out pool_raw, pool_act;  /* outputs on the (H/pool_h, W/pool_w) grid                                  */
  const float* after_w; /* fused after_conv: (3, ncols) fp32 or NULL                                              */
  const float* after_b; /* (3)                                                                                    */
  float* feat;          /* (B, 3, H, W) fp32                                                                      */
} lass_conv_desc;

LASS_API int lass_conv_igemm(const lass_conv_desc* desc_host, void* stream);

/* Debug: shared-memory halo-tile pitch of the conv kernel, 10 (dense, default) or 16 pixels. */
LASS_API int lass_debug_set_halo_pitch(int pitch);

/* ------------------------------------------------------------------------------------------------------
 * Debug: one tcgen05.mma tile (M = 128) with caller-controlled shared-memory descriptors; used by the GPU
 * tests to pin the descriptor rules the conv kernel relies on.  A (a_rows, kc) and Bm (n, kc) are 16-bit
 * K-major; out (128, n) fp32.  swizzle_mode: 0 none, 2 = 128 B, 4 = 64 B, 6 = 32 B.
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_debug_umma_probe(const void* A, int a_rows, const void* Bm, int n, int kc, int swizzle_mode,
                                   int a_start_bytes, int a_sbo, int a_base_offset, int b_sbo, int fmt_fp16,
                                   float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LASS_B200_H_ */
