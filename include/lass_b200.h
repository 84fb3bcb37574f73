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
 *   basis_hi / basis_lo   bf16 (ntiles*256, n_fft): windowed DFT basis split hi/lo, 128-bin tiles with rows
 *             [0,128) = real basis and [128,256) = imaginary basis; row 128 of tile 0 (Im X[0] == 0) carries the real
 *             basis of the Nyquist bin n_fft/2 (see lass_stft_basis_rows; packed from the reference's frozen
 *             `stft.conv_real/conv_imag.weight` by lass_b200.packing.pack_stft_basis)
 *   mag, cos, sin   (B, T, F) fp32 out, T = L/hop + 1, F = n_fft/2 + 1
 *   precision_mode  0 = fp32-parity (3 bf16 MMAs per product, max rel. err ~5e-6), 1 = fast (single bf16 pass)
 *   magphase_mode   0 = Base.spectrogram_phase semantics (mag = clamp(re^2+im^2, 1e-10)**0.5, models/base.py:85-87);
 *                   1 = torchlibrosa.stft.magphase semantics (mag unclamped, cos/sin divided by clamp(mag, 1e-10)) as used by
 *                       the multi-resolution front end (reference scripts/precompute_stfts.py:19-58)
 *   workspace       >= lass_stft_workspace_bytes(B, L, n_fft, hop) bytes, 256-byte aligned
 * Requirements: n_fft % 256 == 0, hop % 8 == 0, L > n_fft/2 (reflect padding).
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_stft_basis_rows(int n_fft);
LASS_API size_t lass_stft_workspace_bytes(int B, int L, int n_fft, int hop);
LASS_API int lass_stft_fwd(const float* wave, int B, int L, int n_fft, int hop, const void* basis_hi, const void* basis_lo,
                  float* mag, float* cos, float* sin, int precision_mode, int magphase_mode, void* workspace,
                  size_t workspace_bytes, void* stream);
/* The multi-resolution front end (reference scripts/precompute_stfts.py:19-58,573-582: win 256 / 512 / 2048 at hop 160 over the
 * same waveforms) as ONE stft_gemm launch (+ one padding launch): nres <= 3 resolutions n_ffts[r] with their own basis pair
 * and (B, T, n_ffts[r]/2 + 1) outputs; the kernel's item list is the concatenation of the resolutions' tiles, longest K first.
 * workspace >= sum of lass_stft_workspace_bytes(B, L, n_ffts[r], hop), 256-byte aligned.  Other arguments as lass_stft_fwd. */
LASS_API int lass_stft_multi_fwd(const float* wave, int B, int L, int hop, int nres, const int* n_ffts,
                                 const void* const* basis_hi, const void* const* basis_lo, float* const* mag, float* const* cos,
                                 float* const* sin, int precision_mode, int magphase_mode, void* workspace,
                                 size_t workspace_bytes, void* stream);

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
  /* Optional rank-1 residual regenerated from a 1-channel fp32 map (the identity residual of encoder_block1, whose
   * input is pre_conv(bn0(mag)), reference models/resunet.py:537-556): for every output pixel (b, h, w), column n
   *   out += resid_w[n] * x + resid_b[n],   x = h < resid_T ? resid_in_scale[w] * resid_src[(b*resid_T + h)*resid_F + w]
   *                                                         + resid_in_shift[w] : 0
   * NULL resid_src disables it. */
  const float* resid_src;
  const float* resid_in_scale;
  const float* resid_in_shift;
  const float* resid_w;
  const float* resid_b;
  int resid_T, resid_F;
  /* Reserved, must be 0 (taps accumulate over K, weights (taps, ncols, cin)).  1 selected a retired "dx-in-N" kernel. */
  int algo;
  /* Optional GENERATED A operand for seg[0] (the first conv of encoder_block1, whose input is the activated pre_conv output
   * of the 1-channel magnitude, reference models/resunet.py:537-556 + ConvBlockRes.bn1): instead of reading seg[0].src the
   * kernel computes, for every input pixel (b, h, w) of the H x W grid and channel c < 32,
   *    x    = h < gen_T ? gen_in_scale[w] * gen_src[(b*gen_T + h)*gen_F + w] + gen_in_shift[w] : 0      (bn0, zero time padding)
   *    a[c] = leaky_relu(gen_scale[c] * (gen_w[c] * x + gen_b[c]) + gen_shift[b*gen_shift_bstride + c])  (pre_conv, BN, FiLM)
   * as bf16 (zero outside the grid).  Needs seg[0] = {cin 32, kc 32, taps 9, bf16}; seg[0].src is ignored.  NULL gen_src
   * disables it. */
  const float* gen_src;
  const float* gen_in_scale;
  const float* gen_in_shift;
  const float* gen_w;
  const float* gen_b;
  const float* gen_scale;
  const float* gen_shift;
  int gen_shift_bstride, gen_T, gen_F;
} lass_conv_desc;

LASS_API int lass_conv_igemm(const lass_conv_desc* desc_host, void* stream);

/* Prepared launches of lass_conv_igemm: tensor maps are encoded and the tile configuration chosen ONCE; lass_conv_run only
 * launches (CUDA-graph capturable).  The device buffers named by the descriptor must stay valid for the handle's life. */
typedef struct lass_conv lass_conv;
LASS_API int lass_conv_prepare(const lass_conv_desc* desc_host, lass_conv** conv_out);
LASS_API int lass_conv_run(const lass_conv* conv, void* stream);
LASS_API void lass_conv_destroy(lass_conv* conv);

/* ------------------------------------------------------------------------------------------------------
 * Training step (BASELINE config 4): what the reference runs per step and rank through autograd --
 * models/audiosep.py:99-111 (ss_model.train(): BatchNorm batch statistics, momentum 0.01; forward; l1_wav, losses.py:4-9;
 * loss.backward()), models/audiosep.py:118-130 (AdamW, amsgrad) -- as explicit kernels.  Convolutions (forward and
 * backward-data) run on lass_conv_*; the entries below are the rest.  Host driver: lass_b200/training.py.
 *
 * Tensors: NHWC 16-bit `x` with `cstride` channels per pixel, channels [coff, coff + C) addressed; `fp16` = 1 for fp16
 * (raw conv outputs and, in training, activations), 0 for bf16 (gradients).  C, cstride, coff multiples of 8.
 * bnp: per-BatchNorm-site block of 6*C floats [scale | shift | mean | rstd | coefA | coefB].
 * ---------------------------------------------------------------------------------------------------- */
/* sums (2, C) float64 = per-channel sum and sum of squares over all pixels (zeroed inside). */
LASS_API int lass_bn_stats(const void* x, int fp16, long long npix, int C, int cstride, int coff, double* sums, void* stream);
/* bn0 runs over the frequency axis (models/resunet.py:537-539): mag (B, T, F) fp32, sums (2, F) over (B, T). */
LASS_API int lass_bn0_stats(const float* mag, int B, int T, int F, double* sums, void* stream);
/* Batch statistics -> bnp[0 .. 4C) = gamma*rstd, beta - mean*gamma*rstd, mean, rstd (biased variance, eps inside the sqrt);
 * running_mean / running_var updated in place with `momentum` and the unbiased variance, like nn.BatchNorm2d in train(). */
LASS_API int lass_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float momentum, float eps, int C, float* bnp, void* stream);
/* lass_bn_stats ADDED into `sums` (no memset inside): the training step zeroes the sums of all its sites with one memset. */
LASS_API int lass_bn_stats_acc(const void* x, int fp16, long long npix, int C, int cstride, int coff, double* sums, void* stream);
/* out = leaky_relu(scale*x + shift + beta[b]) (slope 0.01): BatchNorm + FiLM beta + activation, models/resunet.py:159-160. */
LASS_API int lass_bn_act(const void* x, int x_fp16, int x_cstride, int x_coff, void* out, int out_fp16, int out_cstride,
                         int out_coff, int B, long long pix_per_clip, int C, const float* bnp, const float* beta,
                         int beta_bstride, void* stream);
/* Backward of the same site.  With g' = dact * leaky_relu'(scale*x + shift + beta):
 *   reduce:   sums (B, C, 2) fp32 = per clip [sum g', sum g'*(x - mean)]                       (zeroed inside)
 *   finalize: dgamma = rstd * sum g'(x-mean), dbeta = sum g', dfilm[b] = per-clip sum g' (the FiLM beta gradient, may be NULL),
 *             bnp[4C .. 6C) = coefA, coefB
 *   apply:    dx = scale*g' + coefA*(x - mean) + coefB (+ add)      (native_batch_norm_backward, batch statistics) */
LASS_API int lass_bn_bwd_reduce(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride,
                                int x_coff, int B, long long pix_per_clip, int C, const float* bnp, const float* beta,
                                int beta_bstride, float* sums, void* stream);
LASS_API int lass_bn_bwd_finalize(const float* sums, int B, int C, double count, const float* gamma, float* bnp,
                                  float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride, void* stream);
/* lass_bn_bwd_reduce ADDED into `sums` (no memset inside; the caller zeroes a step's sums at once). */
LASS_API int lass_bn_bwd_reduce_acc(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride,
                                    int x_coff, int B, long long pix_per_clip, int C, const float* bnp, const float* beta,
                                    int beta_bstride, float* sums, void* stream);
/* SyncBatchNorm backward (the reference trains with sync_batchnorm: True, config/audiosep_base.yaml:42, train.py:176,266-283 ->
 * torch.nn.SyncBatchNorm): lass_bn_bwd_totals = this rank's per-channel totals (C, 2) fp64 of the per-clip sums (B, C, 2), which
 * the caller all-reduces over the ranks; lass_bn_bwd_finalize_sync = lass_bn_bwd_finalize with the input-gradient coefficients
 * from those all-reduced totals and the GLOBAL pixel count, while dgamma / dbeta / dfilm stay this rank's sums (DDP averages
 * parameter gradients afterwards).  The forward needs no extra entry: the (2, C) fp64 sums of lass_bn_stats are all-reduced
 * before lass_bn_finalize is called with the global count. */
/* SyncBatchNorm with the statistics exchange fused into the finalize kernels, over NVLink peer memory (no NCCL call, one launch
 * per site and direction).  peer_sums[p] / peer_flags[p] (HOST arrays of `world` device pointers, p = rank): every rank's flat sums
 * buffer and flag table (lass_syncbn_max_peers() x 8-byte epochs per flag_index, zero-initialised once), all mapped on this
 * device (symmetric memory).  The kernel publishes `epoch` (> 0, growing from step to step) to every peer's table, waits for all
 * peers' epochs, then adds every rank's sums at element `sums_offset` in rank order and finalizes like lass_bn_finalize (sums
 * (2, C) fp64) / lass_bn_bwd_finalize_sync.  Backward: the rank first adds ITS per-clip sums (B, C, 2) fp32 at float element
 * `sums_offset` to per-channel totals (C, 2) fp64 which it writes at double element `totals_offset` of its own buffer, publishes,
 * and then reads only the peers' totals (C <= 1024).  The sums must have been written by
 * earlier work of `stream`.  *status (device int, may be NULL) is set to 1 if a peer never arrived (spin limit). */
LASS_API int lass_syncbn_max_peers(void);
LASS_API int lass_bn_finalize_p2p(const void* const* peer_sums, void* const* peer_flags, int world, int rank,
                                  long long sums_offset, int flag_index, unsigned long long epoch, double count_total,
                                  const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  float momentum, float eps, int C, float* bnp, int* status, void* stream);
LASS_API int lass_bn_bwd_finalize_p2p(const void* const* peer_sums, void* const* peer_flags, int world, int rank,
                                      long long sums_offset, long long totals_offset, int flag_index,
                                      unsigned long long epoch, int B, int C, double count_total, const float* gamma,
                                      float* bnp, float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride, int* status,
                                      void* stream);
LASS_API int lass_bn_bwd_totals(const float* sums, int B, int C, double* totals, void* stream);
LASS_API int lass_bn_bwd_finalize_sync(const float* sums, int B, int C, double count_total, const double* totals,
                                       const float* gamma, float* bnp, float* dgamma, float* dbeta, float* dfilm,
                                       int dfilm_bstride, void* stream);
/* A/B variant, NOT used by the training step (measured slower than reduce + the one-block finalize launch, csrc/train.cu):
 * reduce + finalize in ONE launch (count = B * pix_per_clip): the last block to finish (ticket `counter`, one uint32 per site)
 * finalizes.  `sums` and `counter` must be ZERO on entry and are left dirty (the caller clears a step's with one memset each). */
LASS_API int lass_bn_bwd_reduce_finalize(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride,
                                         int x_coff, int B, long long pix_per_clip, int C, float* bnp, const float* beta,
                                         int beta_bstride, float* sums, unsigned int* counter, const float* gamma,
                                         float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride, void* stream);
LASS_API int lass_bn_bwd_apply(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride,
                               int x_coff, const void* add, int add_cstride, int add_coff, void* dx, int dx_cstride,
                               int dx_coff, int B, long long pix_per_clip, int C, const float* bnp, const float* beta,
                               int beta_bstride, void* stream);
/* avg_pool2d backward + skip-connection add (models/resunet.py:196-198): dy (B, H, W, C) = dskip + up(dpool) / (ph*pw). */
LASS_API int lass_pool_bwd(const void* dpool, const void* dskip, int dskip_cstride, int dskip_coff, void* dy, int B, int H,
                           int W, int C, int ph, int pw, void* stream);
/* Gather for the transposed conv's backward (kernel = stride, models/resunet.py:216-224,256): dst (B, H, W, uh*uw*C) with
 * dst[.., (dy*uw+dx)*C + c] = src[b, h*uh+dy, w*uw+dx, coff + c]; dgrad / wgrad of the transposed conv are then 1x1 GEMMs. */
LASS_API int lass_unshuffle(const void* src, int src_cstride, int src_coff, void* dst, int B, int H, int W, int C, int uh,
                            int uw, void* stream);
/* out (C) fp32 = sum over pixels (bias gradient of the 1x1 shortcut convs). */
LASS_API int lass_channel_sum(const void* x, long long npix, int C, int cstride, int coff, float* out, void* stream);
/* Weight gradient dw (taps, co, ci) fp32 (overwritten) = sum_p dy[p, co] * x[p + tap, ci]; taps 9 (3x3, zero padding 1) or 1;
 * dy bf16, x bf16 or fp16 (x_fp16); co, ci multiples of 32.  mma.sync bf16, fp32 accumulate, split over pixel ranges. */
LASS_API int lass_wgrad(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride,
                        int x_coff, int ci, int B, int H, int W, int taps, float* dw, void* stream);
/* The same contract on the 5th-generation tensor cores (tcgen05, accumulators in TMEM, TMA-fed; csrc/wgrad_tc.cu): both
 * operands are the NHWC tiles as they lie in memory (MN-major descriptors, the pixel index is the GEMM K), horizontal taps
 * fill the M = 128 rows, fp16 x tiles are converted to bf16 in shared memory, split over pixel ranges with fp32 red.global. */
LASS_API int lass_wgrad_tc(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride,
                           int x_coff, int ci, int B, int H, int W, int taps, float* dw, void* stream);
/* lass_wgrad_tc ADDED to dw / lass_channel_sum ADDED to out (no memset inside; the training step clears its gradient buffers
 * with one memset per step). */
LASS_API int lass_wgrad_tc_acc(const void* dy, int dy_cstride, int dy_coff, int co, const void* x, int x_fp16, int x_cstride,
                               int x_coff, int ci, int B, int H, int W, int taps, float* dw, void* stream);
LASS_API int lass_channel_sum_acc(const void* x, long long npix, int C, int cstride, int coff, float* out, void* stream);
/* bn0 + zero time padding + Nyquist drop + pre_conv (models/resunet.py:537-555): x0 (B, Tp, Fp, 32) fp16; and its backward
 * from dx0 (bf16): dpre_w, dpre_b (32), dgamma0, dbeta0 (F; the dropped Nyquist bin gets 0).  bnp0 = 6*F block of bn0. */
LASS_API int lass_pre_fwd(const float* mag, int B, int T, int F, int Tp, int Fp, const float* bnp0, const float* pre_w,
                          const float* pre_b, void* x0, void* stream);
LASS_API int lass_pre_bwd(const void* dx0, const float* mag, int B, int T, int F, int Tp, int Fp, const float* bnp0,
                          const float* pre_w, float* dpre_w, float* dpre_b, float* dgamma0, float* dbeta0, void* stream);
/* after_conv (models/resunet.py:570) backward: dfeat (B, 3, npix) fp32, y (B, npix, 32) fp16 -> dy bf16, dw (3, 32), db (3). */
LASS_API int lass_after_bwd(const float* dfeat, const void* y, const float* after_w, void* dy, float* dw, float* db, int B,
                            long long npix, void* stream);
/* Adjoint of torchlibrosa ISTFT.forward (Hermitian extension, IDFT * window, overlap-add, / window-sum, trim) up to the
 * factor c_f / n_fft (applied by lass_mask_bwd): dre, dim (B, T, F) = STFT without padding of the zero-extended
 * dwave / window-sum, on kernel K1 (fp32-parity mode).  Workspace as lass_stft_fwd. */
LASS_API int lass_istft_bwd(const float* dwave, int B, int L, int n_fft, int hop, int T, const float* window,
                            const void* basis_hi, const void* basis_lo, float* dre, float* dim, void* workspace,
                            size_t workspace_bytes, void* stream);
/* Backward of the mask of feature_maps_to_wav (models/resunet.py:457-505): feat / dfeat (B, 3, Tp, Fp) fp32. */
LASS_API int lass_mask_bwd(const float* feat, const float* mag, const float* cos, const float* sin, const float* dre,
                           const float* dim, float* dfeat, int B, int T, int F, int Tp, int Fp, int n_fft, void* stream);
/* l1_wav (losses.py:4-9): *loss_sum += sum |wave - target|, dwave = sign(wave - target) * scale. */
LASS_API int lass_l1_loss(const float* wave, const float* target, long long n, float* loss_sum, float* dwave, float scale,
                          void* stream);
/* FiLM linears backward (models/resunet.py:51-57): dw (J, K) = dbeta^T cond, db (J) = column sums of dbeta (B, J). */
LASS_API int lass_film_bwd(const float* dbeta, const float* cond, float* dw, float* db, int B, int J, int K, void* stream);
/* torch.optim.AdamW(amsgrad=True) single-tensor update over flat fp32 buffers (models/audiosep.py:122-130); step >= 1;
 * the gradient is multiplied by grad_scale first (1 / world size after the NCCL sum). */
LASS_API int lass_adamw_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* fp32 parameter in torch layout -> 16-bit kernel layouts (either may be NULL).  kind 0: conv (co, ci, taps) -> fwd
 * (taps, co, ci) [fp16 if fwd_fp16 else bf16] and dgrad (taps, ci, co) bf16 with flipped taps; kind 1: transposed conv
 * (ci, co, taps) -> fwd (taps*co, ci), dgrad (ci, taps*co).  lass_unpack_grad: packed (taps, co, ci) fp32 -> torch layout. */
LASS_API int lass_pack_weight(const float* w, int kind, int co, int ci, int taps, void* fwd, int fwd_fp16, void* dgrad,
                              void* stream);
LASS_API int lass_unpack_grad(const float* dw, int kind, int co, int ci, int taps, float* grad, void* stream);
/* Multi-tensor forms: ONE launch for every convolution weight of the model (re-pack after the optimizer step) / every weight
 * gradient of an all-reduce bucket.  table_dev: DEVICE array of 8 x int64 per tensor --
 *   pack:   {w, fwd, dgrad, kind, co, ci, taps | fwd_fp16 << 16, first_block}     unpack: {dw, grad, 0, kind, co, ci, taps, first_block}
 * where tensor i owns blocks [first_block_i, first_block_i + n_i) and nblocks is their total; n_i = lass_pack_blocks(kind, co, ci)
 * for pack (32 x 32 tiles of the two leading dimensions) and ceil(co*ci*taps / lass_multi_chunk()) for unpack. */
LASS_API int lass_pack_weights_multi(const long long* table_dev, int nitems, int nblocks, void* stream);
LASS_API int lass_unpack_grads_multi(const long long* table_dev, int nitems, int nblocks, void* stream);
LASS_API int lass_multi_chunk(void);
LASS_API int lass_pack_blocks(int kind, int co, int ci);
/* SegmentMixer.__call__ (data/waveform_mixers.py:19-62, dynamic_loudnorm / get_energy_ratio :65-95), the step in front of the
 * training path (models/audiosep.py:76-78): wave (B, L) fp32 -> mixture, segment (B, L) fp32 (must not alias each other or wave).
 * The reference's random draws are made by the caller in the reference's order and passed as plan (B, max_mix_num + 1) fp32 on the
 * DEVICE: plan[n][0] = mix_num of clip n (2..max_mix_num), plan[n][i] (i = 1..mix_num-1) = gain 10^(dB/20) of the i-th mixed-in
 * clip (n + i) % B, plan[n][max_mix_num] = gain of the summed noise.  scratch: lass_segment_mix_scratch_bytes(B) bytes. */
LASS_API size_t lass_segment_mix_scratch_bytes(int B);
LASS_API int lass_segment_mix(const float* wave, int B, int L, int max_mix_num, const float* plan, float* mixture,
                              float* segment, void* scratch, size_t scratch_bytes, void* stream);
/* Debug: 1 = the shared-memory Stockham iSTFT kernel for every n_fft (default: register-FFT kernel for 1024 / 2048). */
LASS_API int lass_debug_set_istft_v1(int on);

/* Debug: experiments on the conv kernel; read when a launch is PREPARED (lass_conv_igemm, plan creation).  0 = normal operation.
 * Knock-outs, results become wrong: 1 = epilogue skips math and stores, 2 = no tcgen05.mma issued, 4 = no activation (A) TMA
 * loads, 8 = no pooled outputs, 16 = no direct stores, 64 = MMA issuers only.
 * Configuration switches, results stay correct (the parity tests use them to cover both sides of a choice): 32 = direct global
 * stores instead of TMA stores, 128 = single MMA issuer, 256 = generic (unspecialised) epilogue, 512 / 2048 = staged outputs
 * through coalesced / hybrid st.global, 1024 = N 256 x 2 m-tiles, 4096 = no CTA pairs (tcgen05 cta_group::2), 8192 = one tap
 * per streamed weight stage, 16384 = 16-byte instead of 32-byte direct stores, 32768 = CTA pairs for every resident-weight
 * launch, 65536 = the 32-channel transposed conv through TMA stores, 4194304 = conv launches without programmatic dependent launch (default: each conv launch may
 * overlap its prologue with the previous launch's tail and waits with griddepcontrol.wait before it reads anything), 262144 = keep the large tiles for small grids (default: a
 * streamed-weight conv whose items cover less than half of the SMs takes one m-tile and N down to 32). */
LASS_API int lass_debug_set_conv_flags(int flags);
/* Debug: per-CTA role profile of subsequently PREPARED conv launches.  device_counters: >= 16 int64 per CTA
 * (<= 296 CTAs), clock cycles: [0] producer waits for a free A stage, [1] for a free B stage, [2] producer total,
 * [3] MMA issuer waits for a free accumulator, [4] for A data, [5] for B data, [6] MMA issuer total,
 * [7] epilogue waits for an accumulator, [8] epilogue total, [9] items processed.  NULL switches it off. */
LASS_API int lass_debug_set_conv_profile(long long* device_counters);

/* ------------------------------------------------------------------------------------------------------
 * K2  FiLM: all FiLM linears of the model (reference models/resunet.py:59-81) with the eval-mode BatchNorm
 *     shift folded into the bias, as one skinny fp32 GEMM:
 *        shift[b][j] = film_b[j] + sum_k condition[b][k] * film_w[j][k]         (B, J)
 *     The conv epilogues then apply leaky_relu(act_scale[j] * x + shift[b][j]).
 * ---------------------------------------------------------------------------------------------------- */
LASS_API int lass_film(const float* condition, const float* film_w, const float* film_b, int B, int condition_size,
                       int J, float* shift_out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Whole-model entry: ResUNet30.forward for input_channels = output_channels = 1 in eval mode
 * (reference models/resunet.py:522-595,640-653).
 *
 * Row order of the FiLM / activation table (J = 8256 rows for condition -> beta; the six dead
 * `decoder_blockN->beta2` linears are skipped): for encoder block k = 0..6 (encoder_block1..6, conv_block7a):
 * [conv_block1.bn1 (cin_k), conv_block1.bn2 (cout_k)]; then for decoder block j = 0..5:
 * [bn1 (cin_j), conv_block2.bn1 (2*cout_j), conv_block2.bn2 (cout_j)].  lass_resunet30_film_offset() returns the
 * first row of a site: site = 2*k + {0,1} for the encoder, 14 + 3*j + {0,1,2} for the decoder.
 * Channels: encoder cin = {32,32,64,128,256,384,384}, cout = {32,64,128,256,384,384,384};
 *           decoder cin = {384,384,384,256,128,64}, cout = {384,384,256,128,64,32}.
 *
 * Conv weights are 16-bit, packed (taps, ncols, cin): 3x3 convs bf16 with tap = ky*3+kx; transposed convs bf16
 * with column (dy*sw + dx)*cout + co; shortcut (1x1) weights fp16 — an identity matrix where the reference has
 * no shortcut conv (cin == cout).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct lass_resunet30_weights {
  int n_fft, hop, condition_size, film_rows;
  const void* stft_basis_hi;   /* bf16 (lass_stft_basis_rows(n_fft), n_fft) */
  const void* stft_basis_lo;
  const float* istft_window;   /* (n_fft) */
  const float* istft_twiddle;  /* (n_fft, 2) */
  const float* bn0_scale;      /* (n_fft/2 + 1) folded bn0: scale, shift */
  const float* bn0_shift;
  const float* pre_w;          /* (32) pre_conv weight, bias */
  const float* pre_b;
  const float* film_w;         /* (film_rows, condition_size) */
  const float* film_b;         /* (film_rows)  FiLM bias + folded BN shift */
  const float* act_scale;      /* (film_rows)  folded BN scale */
  struct {
    const void* conv1_w;       /* bf16 (9, cout, cin) */
    const void* conv2_w;       /* bf16 (9, cout, cout) */
    const void* sc_w;          /* fp16 (1, cout, cin) shortcut or identity */
    const float* sc_b;         /* (cout) or NULL */
  } enc[7];
  struct {
    const void* up_w;          /* bf16 (1, sh*sw*cout, cin) */
    const void* conv1_w;       /* bf16 (9, cout, 2*cout) */
    const void* conv2_w;       /* bf16 (9, cout, cout) */
    const void* sc_w;          /* fp16 (1, cout, 2*cout) */
    const float* sc_b;         /* (cout) */
  } dec[6];
  const float* after_w;        /* (3, 32) */
  const float* after_b;        /* (3) */
  unsigned int dxn_mask;       /* reserved, must be 0 */
} lass_resunet30_weights;

typedef struct lass_plan lass_plan;

LASS_API int lass_resunet30_film_rows(void);
LASS_API int lass_resunet30_film_offset(int site);
/* Bytes of device workspace a plan for (B clips of L samples) needs. */
LASS_API size_t lass_resunet30_workspace_bytes(int B, int L, int n_fft, int hop);
/* Build a plan (host object: tensor maps, launch list).  `weights_host` is copied; the device buffers it points
 * to and `workspace` must stay valid for the plan's lifetime.  workspace: 1024-byte aligned. */
LASS_API int lass_resunet30_plan_create(const lass_resunet30_weights* weights_host, int B, int L, void* workspace,
                                        size_t workspace_bytes, lass_plan** plan_out);
/* mixture (B, 1, L) fp32, condition (B, condition_size) fp32 -> waveform (B, 1, L) fp32.
 * shift_override: NULL, or a (B, film_rows) fp32 table of precomputed activation shifts (folded BN shift + FiLM
 *   beta) that replaces the FiLM GEMM — the `base(mixtures=, film_dict=)` call of the reference
 *   (models/resunet.py:685-688); `condition` may then be NULL.
 * stft_precision_mode: 0 = fp32-parity STFT, 1 = single-pass bf16 STFT.  Asynchronous on `stream`.
 * lass_resunet30_forward_stages runs a subset (bit mask) of the three stages so a caller can bracket them with
 * its own CUDA events: LASS_STAGE_FRONT (STFT, FiLM), LASS_STAGE_UNET (all convolutions, incl. the fused bn0 + pre_conv),
 * LASS_STAGE_BACK (mask + iSTFT).  lass_resunet30_forward == all stages. */
#define LASS_STAGE_FRONT 1
#define LASS_STAGE_UNET 2
#define LASS_STAGE_BACK 4
#define LASS_STAGE_ALL 7
LASS_API int lass_resunet30_forward(lass_plan* plan, const float* mixture, const float* condition,
                                    const float* shift_override, float* waveform, int stft_precision_mode,
                                    void* stream);
LASS_API int lass_resunet30_forward_stages(lass_plan* plan, int stage_mask, const float* mixture,
                                           const float* condition, const float* shift_override, float* waveform,
                                           int stft_precision_mode, void* stream);
/* Algorithmic FLOPs (2*M*N*K summed over the plan's convolution launches) and launches of the UNET stage. */
LASS_API double lass_resunet30_unet_flops(const lass_plan* plan);
/* Debug: runs the UNET stage once with a CUDA event between launches; ms_out[i] / flops_out[i] (optional) receive the
 * duration and algorithmic FLOPs of launch i (encoder blocks 1..7: conv1, conv2; decoder blocks 1..6: transposed conv,
 * conv1, conv2).  Returns the number of launches (<= capacity) or a negative error.  Synchronises the stream. */
LASS_API int lass_debug_time_unet_launches(lass_plan* plan, float* ms_out, double* flops_out, int capacity, void* stream);
/* Number of kernel launches one lass_resunet30_forward issues. */
LASS_API int lass_resunet30_num_launches(const lass_plan* plan);
/* Debug / test access to intermediates in the workspace: name in {"mag","cos","sin","shift","feat",
 * "x_raw0".."x_raw6","x_act0".."x_act6","a2_0".."a2_6","cat_raw0".."cat_raw5","cat_act0".."cat_act5",
 * "d_act1".."d_act6"}; returns a device pointer or NULL; dims = {d0,d1,d2,d3} elements, *elem_bytes 2 or 4. */
LASS_API void* lass_resunet30_buffer(const lass_plan* plan, const char* name, int dims[4], int* elem_bytes);
LASS_API void lass_resunet30_plan_destroy(lass_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* LASS_B200_H_ */
