// Training step: the memory-bound kernels around the convolutions (BASELINE config 4).
//
// Replaces, per training step of the reference (models/audiosep.py:99-111 with ss_model.train(), autograd, AdamW):
//   nn.BatchNorm2d in batch-statistics mode (33 sites) + FiLM add + leaky_relu  -> bn_stats / bn_finalize / bn_act
//   their autograd backward (native_batch_norm_backward, leaky_relu_backward)     -> bn_bwd_reduce / finalize / apply
//   avg_pool2d backward + the skip-connection gradient add                         -> pool_bwd
//   conv_transpose2d(kernel = stride) backward's gather                            -> unshuffle (the GEMMs run in conv.cu / wgrad.cu)
//   bn0 + pad + pre_conv forward / backward (models/resunet.py:537-555)            -> bn0_stats / pre_fwd / pre_bwd
//   after_conv backward (:570), mask backward (:457-505), ISTFT adjoint            -> after_bwd / mask_bwd / istft_bwd (stft.cu)
//   l1_wav (losses.py:4-9), FiLM linears backward, AdamW(amsgrad) (models/audiosep.py:122-130)
//
// Tensors: NHWC 16-bit with a channel stride / offset (slices of the concat buffers), fp16 or bf16 per flag; all math fp32.
// Every kernel maps one thread to a 16-byte vector of 8 channels, consecutive threads to consecutive vectors.
#include <string.h>

#include "lass_internal.cuh"

namespace lass {
namespace {

constexpr float kSlope = 0.01f;
constexpr int kRedThreads = 384;   // divisible by C/8 for every channel count of the model (4 .. 96 vectors)

struct V8 {
  float v[8];
};

__device__ __forceinline__ V8 load8(const void* base, size_t elem_off, int fp16) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + elem_off));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  V8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    } else {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  return r;
}

__device__ __forceinline__ uint4 loadq(const void* base, size_t elem_off) {
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + elem_off));
}
// 8 packed 16-bit values -> fp32 (kept packed in 4 registers until they are used: the elementwise kernels hold several
// loads in flight per thread and must stay under 85 registers for two 384-thread CTAs per SM)
__device__ __forceinline__ V8 unpack8(const uint4& q, int fp16) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  V8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    } else {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  return r;
}

__device__ __forceinline__ void store8(void* base, size_t elem_off, int fp16, const V8& r) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) {
      const float a = fminf(fmaxf(r.v[2 * i], -65504.0f), 65504.0f), b = fminf(fmaxf(r.v[2 * i + 1], -65504.0f), 65504.0f);
      const __half2 h = __floats2half2_rn(a, b);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + elem_off) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ V8 ldf8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  V8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// per-channel reductions over pixels: block = 384 threads = R pixel rows x CV channel vectors
// ---------------------------------------------------------------------------------------------------------------
// NACC accumulators per channel.  After the grid-stride loop the R partial rows of a block are summed through shared memory.
template <int NACC>
__device__ __forceinline__ void block_reduce_rows(float (&acc)[NACC][8], int CV, int R, int cv, int prow, float* red /* [R][CV*8*NACC] */) {
  const int width = CV * 8 * NACC;
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[prow * width + (a * CV + cv) * 8 + i] = acc[a][i];
  __syncthreads();
}

// Optional tail of bn_bwd_reduce (kept for A/B, lass_bn_bwd_reduce_finalize): the LAST block to finish (ticket counter,
// __threadfence) finalizes.  Measured on B200 it is SLOWER than the separate one-block finalize launch it replaces (+8.5 us per
// site against ~5 us for kernel + launch gap), and so is deriving the forward tables in bn_act's prologue (+12 us: the fp64
// divisions / square roots repeated by every block) -- the training step therefore keeps the separate finalize launches.
__device__ __forceinline__ bool last_block_done(unsigned int* counter, unsigned int nblocks) {
  __shared__ unsigned int ticket;
  __threadfence();                                  // this block's atomics are visible before its ticket
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(counter, 1u);
  __syncthreads();
  const bool last = ticket == nblocks - 1;
  if (last) __threadfence();
  return last;
}

__device__ __forceinline__ void bn_finalize_channel(double s, double q, int c, int C, double count, const float* gamma, const float* beta,
                                                    float* running_mean, float* running_var, float momentum, float eps, float* bnp) {
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double scale = (double)gamma[c] * rstd;
  bnp[c] = (float)scale;
  bnp[C + c] = (float)((double)beta[c] - mean * scale);
  bnp[2 * C + c] = (float)mean;
  bnp[3 * C + c] = (float)rstd;
  const double unbiased = var * (count / fmax(count - 1.0, 1.0));
  running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
  running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
}

__global__ void __launch_bounds__(kRedThreads, 2) bn_stats_kernel(const void* __restrict__ x, int fp16, long long npix, int C, int cstride,
                                                               int coff, double* __restrict__ sums) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  extern __shared__ float red[];
  const int CV = C / 8, R = kRedThreads / CV;
  const int cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.0f;
  const long long stride = (long long)gridDim.x * R;
  for (long long p = (long long)blockIdx.x * R + prow; p < npix; p += 4 * stride) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)                 // four independent 16-byte loads in flight per thread
      if (p + u * stride < npix) raw[u] = loadq(x, (size_t)(p + u * stride) * cstride + coff + cv * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * stride < npix) {
        const V8 v = unpack8(raw[u], fp16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += v.v[i];
          acc[1][i] = fmaf(v.v[i], v.v[i], acc[1][i]);
        }
      }
  }
  block_reduce_rows<2>(acc, CV, R, cv, prow, red);
  const int width = C * 2;
  for (int e = threadIdx.x; e < width; e += kRedThreads) {
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += (double)red[r * width + e];
    atomicAdd(&sums[e], s);          // e < C: sum, e >= C: sum of squares  (layout (2, C))
  }
}

__global__ void bn0_stats_kernel(const float* __restrict__ mag, int rows, int F, int rows_per_block, double* __restrict__ sums) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s = 0.0f, q = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const float v = __ldg(mag + (size_t)r * F + f);
    s += v;
    q = fmaf(v, v, q);
  }
  atomicAdd(&sums[f], (double)s);
  atomicAdd(&sums[F + f], (double)q);
}

// bnp = [scale | shift | mean | rstd | coefA | coefB] x C
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, int C, float* __restrict__ bnp) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  bn_finalize_channel(sums[c], sums[C + c], c, C, count, gamma, beta, running_mean, running_var, momentum, eps, bnp);
}

// Elementwise kernels: block = 384 threads = R pixels x CV channel vectors, grid.y = clip.  A thread keeps ITS channel vector for
// the whole launch, so the per-channel tables live in registers; it walks pixels with a stride of gridDim.x * R and has kU
// independent 16-byte loads per operand in flight (no 64-bit index divisions, no table reloads).
__global__ void __launch_bounds__(kRedThreads, 2) bn_act_kernel(const void* __restrict__ x, int x_fp16, int x_cstride, int x_coff,
                                                             void* __restrict__ out, int out_fp16, int out_cstride, int out_coff,
                                                             long long pix_per_clip, int C, const float* __restrict__ bnp,
                                                             const float* __restrict__ beta, int beta_bstride) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  constexpr int kU = 4;
  const int CV = C / 8, R = kRedThreads / CV;
  const int cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  if (prow >= R) return;
  const int b = blockIdx.y, c = cv * 8;
  const V8 sc = ldf8(bnp + c);
  V8 sh = ldf8(bnp + C + c);
  {
    const V8 be = ldf8(beta + (size_t)b * beta_bstride + c);
#pragma unroll
    for (int k = 0; k < 8; ++k) sh.v[k] += be.v[k];
  }
  const size_t p0 = (size_t)b * pix_per_clip;
  const long long stride = (long long)gridDim.x * R;
  for (long long q = (long long)blockIdx.x * R + prow; q < pix_per_clip; q += kU * stride) {
    uint4 raw[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (q + u * stride < pix_per_clip) raw[u] = loadq(x, (p0 + q + u * stride) * x_cstride + x_coff + c);
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (q + u * stride < pix_per_clip) {
        const V8 v = unpack8(raw[u], x_fp16);
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float pre = fmaf(sc.v[k], v.v[k], sh.v[k]);
          r.v[k] = pre > 0.0f ? pre : kSlope * pre;
        }
        store8(out, (p0 + q + u * stride) * out_cstride + out_coff + c, out_fp16, r);
      }
  }
}

struct BnBwdFinalize {
  unsigned int* counter;
  double count;
  const float* gamma;
  float* bnp;           // reads scale / rstd, writes coefA / coefB
  float* dgamma;
  float* dbeta;
  float* dfilm;         // (B, dfilm_bstride) or nullptr
  int dfilm_bstride;
};

__device__ __forceinline__ void bn_bwd_finalize_totals(double t1, double t2, int c, int C, double count, float* bnp, float* dgamma,
                                                       float* dbeta) {
  const double scale = (double)bnp[c], rstd = (double)bnp[3 * C + c];
  dbeta[c] = (float)t1;
  dgamma[c] = (float)(rstd * t2);
  bnp[4 * C + c] = (float)(-scale * rstd * rstd * t2 / count);
  bnp[5 * C + c] = (float)(-scale * t1 / count);
}

__device__ __forceinline__ void bn_bwd_finalize_channel(const float* sums, int c, int B, int C, double count, float* bnp, float* dgamma,
                                                        float* dbeta, float* dfilm, int dfilm_bstride) {
  double t1 = 0.0, t2 = 0.0;
  for (int b = 0; b < B; ++b) {
    const float s1 = sums[((size_t)b * C + c) * 2], s2 = sums[((size_t)b * C + c) * 2 + 1];
    t1 += (double)s1;
    t2 += (double)s2;
    if (dfilm) dfilm[(size_t)b * dfilm_bstride + c] = s1;
  }
  bn_bwd_finalize_totals(t1, t2, c, C, count, bnp, dgamma, dbeta);
}

// sums (B, C, 2): [sum g', sum g' (x - mean)] per clip; grid.y = clip
__global__ void __launch_bounds__(kRedThreads, 2) bn_bwd_reduce_kernel(const void* __restrict__ dact, int d_cstride, int d_coff,
                                                                    const void* __restrict__ x, int x_fp16, int x_cstride, int x_coff,
                                                                    long long pix_per_clip, int C, const float* __restrict__ bnp,
                                                                    const float* __restrict__ beta, int beta_bstride,
                                                                    float* __restrict__ sums, const BnBwdFinalize fin) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  extern __shared__ float red[];
  const int CV = C / 8, R = kRedThreads / CV;
  const int cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  const int b = blockIdx.y, c = cv * 8;
  const V8 sc = ldf8(bnp + c), mean = ldf8(bnp + 2 * C + c);
  V8 sh = ldf8(bnp + C + c);
  {
    const V8 be = ldf8(beta + (size_t)b * beta_bstride + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) sh.v[i] += be.v[i];
  }
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.0f;
  const long long stride = (long long)gridDim.x * R;
  for (long long q = (long long)blockIdx.x * R + prow; q < pix_per_clip; q += 2 * stride) {
    uint4 xq[2], dq[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (q + u * stride < pix_per_clip) {
        const size_t p = (size_t)b * pix_per_clip + q + u * stride;
        xq[u] = loadq(x, p * x_cstride + x_coff + c);
        dq[u] = loadq(dact, p * d_cstride + d_coff + c);
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (q + u * stride < pix_per_clip) {
        const V8 xv = unpack8(xq[u], x_fp16), dv = unpack8(dq[u], 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pre = fmaf(sc.v[i], xv.v[i], sh.v[i]);
          const float g = pre > 0.0f ? dv.v[i] : kSlope * dv.v[i];
          acc[0][i] += g;
          acc[1][i] = fmaf(g, xv.v[i] - mean.v[i], acc[1][i]);
        }
      }
  }
  block_reduce_rows<2>(acc, CV, R, cv, prow, red);
  const int width = C * 2;
  for (int e = threadIdx.x; e < width; e += kRedThreads) {
    float s = 0.0f;
    for (int r = 0; r < R; ++r) s += red[r * width + e];
    const int a = e / C, ch = e - a * C;
    atomicAdd(&sums[((size_t)b * C + ch) * 2 + a], s);
  }
  if (fin.counter == nullptr) return;
  if (!last_block_done(fin.counter, gridDim.x * gridDim.y)) return;
  // phase 1: every (clip, channel) pair is an independent 8-byte load (no chain of L2 round trips); per-channel totals in shared memory
  double* tot = reinterpret_cast<double*>(red);           // [C][2]; red holds R * C * 2 floats with R >= 4
  for (int e = threadIdx.x; e < 2 * C; e += kRedThreads) tot[e] = 0.0;
  __syncthreads();
  const int nb = gridDim.y;
  for (int i = threadIdx.x; i < nb * C; i += kRedThreads) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(sums) + i);
    const int bb = i / C, ch = i - bb * C;
    if (fin.dfilm) fin.dfilm[(size_t)bb * fin.dfilm_bstride + ch] = v.x;
    atomicAdd(&tot[2 * ch], (double)v.x);
    atomicAdd(&tot[2 * ch + 1], (double)v.y);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += kRedThreads)
    bn_bwd_finalize_totals(tot[2 * ch], tot[2 * ch + 1], ch, C, fin.count, fin.bnp, fin.dgamma, fin.dbeta);
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sums, int B, int C, double count, const float* __restrict__ gamma,
                                       float* __restrict__ bnp, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dfilm, int dfilm_bstride) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  bn_bwd_finalize_channel(sums, c, B, C, count, bnp, dgamma, dbeta, dfilm, dfilm_bstride);
}

// SyncBatchNorm backward (the reference trains with sync_batchnorm: True, config/audiosep_base.yaml:42 -> torch.nn.SyncBatchNorm):
// the input gradient needs the two sums over ALL ranks' pixels, the parameter gradients (dgamma, dbeta, dFiLM) stay this rank's
// (DDP averages them afterwards).  bn_bwd_totals = this rank's per-channel totals (C, 2) fp64 for the all-reduce;
// bn_bwd_finalize_sync = the finalize with coefficients from the all-reduced totals and the global pixel count.
__global__ void bn_bwd_totals_kernel(const float* __restrict__ sums, int B, int C, double* __restrict__ totals) {
  griddep_launch_dependents();
  griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int b = 0; b < B; ++b) {
    t1 += (double)sums[((size_t)b * C + c) * 2];
    t2 += (double)sums[((size_t)b * C + c) * 2 + 1];
  }
  totals[2 * c] = t1;
  totals[2 * c + 1] = t2;
}

__global__ void bn_bwd_finalize_sync_kernel(const float* __restrict__ sums, int B, int C, double count_total,
                                            const double* __restrict__ totals, float* __restrict__ bnp, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, float* __restrict__ dfilm, int dfilm_bstride) {
  griddep_launch_dependents();
  griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int b = 0; b < B; ++b) {
    const float s1 = sums[((size_t)b * C + c) * 2], s2 = sums[((size_t)b * C + c) * 2 + 1];
    t1 += (double)s1;
    t2 += (double)s2;
    if (dfilm) dfilm[(size_t)b * dfilm_bstride + c] = s1;
  }
  const double scale = (double)bnp[c], rstd = (double)bnp[3 * C + c];
  dbeta[c] = (float)t1;                                  // this rank's share (DDP averages the parameter gradients)
  dgamma[c] = (float)(rstd * t2);
  bnp[4 * C + c] = (float)(-scale * rstd * rstd * totals[2 * c + 1] / count_total);
  bnp[5 * C + c] = (float)(-scale * totals[2 * c] / count_total);
}

// ---- SyncBatchNorm over NVLink peer memory: the statistics exchange fused INTO the finalize kernels ----
// Every rank keeps its per-site sums in a symmetric-memory buffer that all ranks of the node have mapped (peer pointers over
// NVLink / NVSwitch).  Instead of [all-reduce launch -> finalize launch] the finalize kernel itself (i) tells every peer "my sums
// of this site are complete" by writing the step's epoch number into its slot of the peer's flag table (st.release.sys after a
// system fence; the sums were written by the previous kernel of this stream), (ii) waits until all peers' epochs have arrived
// in its own table (ld.acquire.sys), (iii) loads every rank's sums through the peer pointers and adds them IN RANK ORDER in fp64
// -- the same numbers in the same order on every rank, so all ranks derive bit-identical tables -- and (iv) finalizes.
// One launch per site and direction, no NCCL call, ~2 NVLink round trips of latency.  Flags only ever grow (epoch = step
// counter), so nothing is reset; the sums are double-buffered by epoch parity, so a fast rank's memset for step t + 1 cannot
// touch what a slow rank still reads for step t.  A spin that exceeds kSpinLimit polls sets *status and gives up (wrong numbers,
// loudly reported by the host) rather than hanging the GPU.
constexpr int kMaxPeers = 16;
constexpr long long kSpinLimit = 1ll << 26;
struct PeerTable {
  const void* sums[kMaxPeers];            // base of every rank's flat sums buffer
  unsigned long long* flags[kMaxPeers];   // base of every rank's flag table: [flag_index][kMaxPeers] epochs
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// publish this rank's epoch to all peers (block 0 only) and wait for all peers' epochs (every block); false on timeout
__device__ __forceinline__ bool peer_exchange(const PeerTable& pt, int world, int rank, int flag_index, unsigned long long epoch, int* status) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const int p = threadIdx.x;
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(pt.flags[p] + (size_t)flag_index * kMaxPeers + rank, epoch);
    }
    const unsigned long long* mine = pt.flags[rank] + (size_t)flag_index * kMaxPeers + p;
    long long spins = 0;
    // once an exchange has timed out (*status set) no later one waits again: the step is lost, but it ends
    const bool dead = status != nullptr && *reinterpret_cast<volatile int*>(status) != 0;
    while (!dead && ld_acquire_sys(mine) < epoch) {
      if (++spins > kSpinLimit) {
        ok = 0;
        if (status) atomicExch(status, 1);
        break;
      }
    }
  }
  __syncthreads();
  return ok != 0;
}

// one 8-byte load from every peer, all in flight before the first use (a peer load is an NVLink round trip)
template <typename T>
__device__ __forceinline__ void load_peers(const T* const (&ptr)[kMaxPeers], int world, T (&v)[kMaxPeers]) {
#pragma unroll
  for (int p = 0; p < kMaxPeers; ++p)
    if (p < world) v[p] = __ldcv(ptr[p]);
}

__global__ void bn_finalize_p2p_kernel(PeerTable pt, int world, int rank, long long sums_offset, int flag_index, unsigned long long epoch,
                                       double count_total, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps, int C,
                                       float* __restrict__ bnp, int* __restrict__ status) {
  griddep_launch_dependents();
  griddep_wait();
  peer_exchange(pt, world, rank, flag_index, epoch, status);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* ps[kMaxPeers];
  const double* pq[kMaxPeers];
#pragma unroll
  for (int p = 0; p < kMaxPeers; ++p) {
    const double* base = reinterpret_cast<const double*>(pt.sums[p < world ? p : 0]) + sums_offset;     // (2, C) of this site
    ps[p] = base + c;
    pq[p] = base + C + c;
  }
  double vs[kMaxPeers], vq[kMaxPeers];
  load_peers(ps, world, vs);
  load_peers(pq, world, vq);
  double s = 0.0, q = 0.0;
#pragma unroll
  for (int p = 0; p < kMaxPeers; ++p)
    if (p < world) s += vs[p], q += vq[p];               // rank order: the same sum on every rank
  bn_finalize_channel(s, q, c, C, count_total, gamma, beta, running_mean, running_var, momentum, eps, bnp);
}

// ONE block: (i) this rank's per-channel totals of its per-clip sums (local loads) -> its totals area, (ii) publish / wait,
// (iii) every rank's totals through the peer pointers (world loads per channel instead of world x B), finalize.
constexpr int kP2PBwdThreads = 256;
constexpr int kP2PBwdMaxPerThread = 4;        // C <= 1024
__global__ void __launch_bounds__(kP2PBwdThreads) bn_bwd_finalize_p2p_kernel(
    PeerTable pt, int world, int rank, long long sums_offset, long long totals_offset, int flag_index, unsigned long long epoch, int B, int C,
    double count_total, float* __restrict__ bnp, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dfilm,
    int dfilm_bstride, int* __restrict__ status) {
  griddep_launch_dependents();
  griddep_wait();
  const float2* mine = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(pt.sums[rank]) + sums_offset);   // (B, C) x [g, g (x - mean)]
  double* mytot = reinterpret_cast<double*>(const_cast<void*>(pt.sums[rank])) + totals_offset;                        // (C, 2)
  double l1[kP2PBwdMaxPerThread], l2[kP2PBwdMaxPerThread];
#pragma unroll
  for (int i = 0; i < kP2PBwdMaxPerThread; ++i) {
    const int c = threadIdx.x + i * kP2PBwdThreads;
    l1[i] = l2[i] = 0.0;
    if (c < C) {
      for (int b = 0; b < B; ++b) {
        const float2 v = __ldcg(mine + (size_t)b * C + c);
        l1[i] += (double)v.x;
        l2[i] += (double)v.y;
        if (dfilm) dfilm[(size_t)b * dfilm_bstride + c] = v.x;
      }
      mytot[2 * c] = l1[i];
      mytot[2 * c + 1] = l2[i];
    }
  }
  __threadfence_system();
  __syncthreads();                    // all totals of this rank are written (and fenced) before its epoch is published
  peer_exchange(pt, world, rank, flag_index, epoch, status);
#pragma unroll
  for (int i = 0; i < kP2PBwdMaxPerThread; ++i) {
    const int c = threadIdx.x + i * kP2PBwdThreads;
    if (c >= C) continue;
    const double* p1[kMaxPeers];
    const double* p2[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      const double* base = reinterpret_cast<const double*>(pt.sums[p < world ? p : 0]) + totals_offset;
      p1[p] = base + 2 * c;
      p2[p] = base + 2 * c + 1;
    }
    double v1[kMaxPeers], v2[kMaxPeers];
    load_peers(p1, world, v1);
    load_peers(p2, world, v2);
    double g1 = 0.0, g2 = 0.0;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < world) g1 += v1[p], g2 += v2[p];
    const double scale = (double)bnp[c], rstd = (double)bnp[3 * C + c];
    dbeta[c] = (float)l1[i];                              // this rank's share (the gradient all-reduce averages parameter gradients)
    dgamma[c] = (float)(rstd * l2[i]);
    bnp[4 * C + c] = (float)(-scale * rstd * rstd * g2 / count_total);
    bnp[5 * C + c] = (float)(-scale * g1 / count_total);
  }
}

__global__ void __launch_bounds__(kRedThreads, 2) bn_bwd_apply_kernel(const void* __restrict__ dact, int d_cstride, int d_coff,
                                                                   const void* __restrict__ x, int x_fp16, int x_cstride, int x_coff,
                                                                   const void* __restrict__ add, int add_cstride, int add_coff,
                                                                   void* __restrict__ dx, int dx_cstride, int dx_coff, long long pix_per_clip,
                                                                   int C, const float* __restrict__ bnp, const float* __restrict__ beta,
                                                                   int beta_bstride) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  constexpr int kU = 2;
  const int CV = C / 8, R = kRedThreads / CV;
  const int cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  if (prow >= R) return;
  const int b = blockIdx.y, c = cv * 8;
  const V8 sc = ldf8(bnp + c), ca = ldf8(bnp + 4 * C + c);
  V8 sh = ldf8(bnp + C + c), cb = ldf8(bnp + 5 * C + c);
  {
    const V8 mean = ldf8(bnp + 2 * C + c), be = ldf8(beta + (size_t)b * beta_bstride + c);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cb.v[k] = fmaf(-ca.v[k], mean.v[k], cb.v[k]);      // coefA (x - mean) + coefB = coefA x + (coefB - coefA mean)
      sh.v[k] += be.v[k];
    }
  }
  const size_t p0 = (size_t)b * pix_per_clip;
  const long long stride = (long long)gridDim.x * R;
  for (long long q = (long long)blockIdx.x * R + prow; q < pix_per_clip; q += kU * stride) {
    uint4 xq[kU], dq[kU], aq[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (q + u * stride < pix_per_clip) {
        const size_t p = p0 + q + u * stride;
        xq[u] = loadq(x, p * x_cstride + x_coff + c);
        dq[u] = loadq(dact, p * d_cstride + d_coff + c);
        if (add) aq[u] = loadq(add, p * add_cstride + add_coff + c);
      }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (q + u * stride < pix_per_clip) {
        const V8 xv = unpack8(xq[u], x_fp16), dv = unpack8(dq[u], 0);
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float pre = fmaf(sc.v[k], xv.v[k], sh.v[k]);
          const float g = pre > 0.0f ? dv.v[k] : kSlope * dv.v[k];
          r.v[k] = fmaf(sc.v[k], g, fmaf(ca.v[k], xv.v[k], cb.v[k]));
        }
        if (add) {
          const V8 av = unpack8(aq[u], 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) r.v[k] += av.v[k];
        }
        store8(dx, (p0 + q + u * stride) * dx_cstride + dx_coff + c, 0, r);
      }
  }
}

// The three copy-like kernels below: grid.y = output image row (b, h), 32-bit index arithmetic inside the row, four independent
// 16-byte loads per operand in flight per thread (the first versions ran a 64-bit div / mod chain per vector with one load in flight:
// 3.7-4.0 TB/s).
__global__ void __launch_bounds__(256) pool_bwd_kernel(const void* __restrict__ dpool, const void* __restrict__ dskip, int s_cstride,
                                                       int s_coff, void* __restrict__ dy, int B, int H, int W, int C, int ph, int pw) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  constexpr int kU = 4;
  const int CV = C / 8;
  const float inv = 1.0f / (float)(ph * pw);
  const int Hp = H / ph, Wp = W / pw;
  const int nv = W * CV, stride = gridDim.x * 256;
  for (int row = blockIdx.y; row < B * H; row += gridDim.y) {
  const int b = row / H, h = row - b * H;
  const size_t prow = ((size_t)b * Hp + h / ph) * Wp;      // first pooled pixel of the source row
  const size_t orow = (size_t)row * W;                      // first output pixel of this row
  for (int i0 = blockIdx.x * 256 + threadIdx.x; i0 < nv; i0 += kU * stride) {
    uint4 dp[kU], ds[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      if (i < nv) {
        const int w = i / CV, c = (i - w * CV) * 8;
        dp[u] = loadq(dpool, (prow + w / pw) * C + c);
        if (dskip) ds[u] = loadq(dskip, (orow + w) * s_cstride + s_coff + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      if (i < nv) {
        const V8 a = unpack8(dp[u], 0);
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] = a.v[k] * inv;
        if (dskip) {
          const V8 d = unpack8(ds[u], 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) r.v[k] += d.v[k];
        }
        store8(dy, orow * C + (size_t)i * 8, 0, r);
      }
    }
  }
  }
}

// dst (B, H, W, uh*uw*C)[(dy*uw + dx)*C + c] = src (B, H*uh, W*uw, cstride)[h*uh + dy, w*uw + dx, coff + c]   (16-byte copies)
__global__ void __launch_bounds__(256) unshuffle_kernel(const uint16_t* __restrict__ src, int cstride, int coff, uint16_t* __restrict__ dst,
                                                        int B, int H, int W, int C, int uh, int uw) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  constexpr int kU = 4;
  const int CV = C / 8, G = uh * uw;
  const int nv = W * G * CV, stride = gridDim.x * 256;
  const size_t Ws = (size_t)W * uw;
  for (int row = blockIdx.y; row < B * H; row += gridDim.y) {      // destination rows
  const int b = row / H, h = row - b * H;
  const size_t srow = ((size_t)b * H + h) * uh * Ws;              // first source pixel of source row h * uh
  uint4* drow = reinterpret_cast<uint4*>(dst) + (size_t)row * nv;
  for (int i0 = blockIdx.x * 256 + threadIdx.x; i0 < nv; i0 += kU * stride) {
    uint4 q[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      if (i < nv) {
        const int t = i / CV, cv = i - t * CV;
        const int w = t / G, g = t - w * G;
        const int dy = g / uw, dx = g - dy * uw;
        q[u] = __ldg(reinterpret_cast<const uint4*>(src + (srow + (size_t)dy * Ws + (size_t)w * uw + dx) * cstride + coff + cv * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      if (i < nv) drow[i] = q[u];
    }
  }
  }
}

__global__ void __launch_bounds__(kRedThreads) channel_sum_kernel(const void* __restrict__ x, long long npix, int C, int cstride, int coff,
                                                                  float* __restrict__ out) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  extern __shared__ float red[];
  const int CV = C / 8, R = kRedThreads / CV;
  const int cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.0f;
  const long long stride = (long long)gridDim.x * R;
  for (long long p = (long long)blockIdx.x * R + prow; p < npix; p += 4 * stride) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * stride < npix) raw[u] = loadq(x, (size_t)(p + u * stride) * cstride + coff + cv * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * stride < npix) {
        const V8 v = unpack8(raw[u], 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[0][i] += v.v[i];
      }
  }
  block_reduce_rows<1>(acc, CV, R, cv, prow, red);
  for (int e = threadIdx.x; e < C; e += kRedThreads) {
    float s = 0.0f;
    for (int r = 0; r < R; ++r) s += red[r * C + e];
    atomicAdd(&out[e], s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bn0 + zero time padding + Nyquist drop + pre_conv (1 -> 32), reference models/resunet.py:537-555
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pre_fwd_kernel(const float* __restrict__ mag, int B, int T, int F, int Tp, int Fp,
                                                      const float* __restrict__ bnp0, const float* __restrict__ pre_w,
                                                      const float* __restrict__ pre_b, void* __restrict__ x0) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  // grid.y = output row (b, h); a thread keeps its 8-channel vector of the pre_conv weights, walks the row's pixels
  const int cv = threadIdx.x & 3;
  const V8 wv = ldf8(pre_w + cv * 8), bv = ldf8(pre_b + cv * 8);
  const int stride = gridDim.x * 64;
  for (int row = blockIdx.y; row < B * Tp; row += gridDim.y) {
  const int b = row / Tp, h = row - b * Tp;
  const float* mrow = mag + ((size_t)b * T + h) * F;
  const bool live = h < T;                                   // zero time padding AFTER bn0 (models/resunet.py:548)
  for (int w0 = blockIdx.x * 64 + (threadIdx.x >> 2); w0 < Fp; w0 += 4 * stride) {
    float xb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int w = w0 + u * stride;
      xb[u] = (live && w < Fp) ? fmaf(__ldg(bnp0 + w), __ldg(mrow + w), __ldg(bnp0 + F + w)) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int w = w0 + u * stride;
      if (w < Fp) {
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] = fmaf(wv.v[k], xb[u], bv.v[k]);
        store8(x0, ((size_t)row * Fp + w) * 32 + cv * 8, 1, r);
      }
    }
  }
  }
}

// One thread per (pixel, 8-channel vector); a block covers kPreRows image rows x 64 columns (256 threads = 64 px x 4 vectors).
// Outputs (all accumulated with atomics into zeroed buffers): dpre_w[32], dpre_b[32], dgamma0[F], dbeta0[F].
constexpr int kPreRows = 16;
__global__ void __launch_bounds__(256) pre_bwd_kernel(const void* __restrict__ dx0, const float* __restrict__ mag, int B, int T, int F, int Tp,
                                                      int Fp, const float* __restrict__ bnp0, const float* __restrict__ pre_w,
                                                      float* __restrict__ dpre_w, float* __restrict__ dpre_b,
                                                      float* __restrict__ dgamma0, float* __restrict__ dbeta0) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  __shared__ float red[2][8][32];     // [w|b][warp][channel]
  const int cv = threadIdx.x & 3, col = threadIdx.x >> 2;
  const int w = blockIdx.x * 64 + col;
  const int b = blockIdx.z;
  const int h0 = blockIdx.y * kPreRows;
  const V8 pw = ldf8(pre_w + cv * 8);
  float aw[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) aw[k] = ab[k] = 0.0f;
  float dg = 0.0f, db = 0.0f;
  const bool wok = w < Fp;
  const float sc0 = wok ? __ldg(bnp0 + w) : 0.0f, sh0 = wok ? __ldg(bnp0 + F + w) : 0.0f;
  const float mean0 = wok ? __ldg(bnp0 + 2 * F + w) : 0.0f, rstd0 = wok ? __ldg(bnp0 + 3 * F + w) : 0.0f;
  for (int h = h0; h < min(Tp, h0 + kPreRows); ++h) {
    float part = 0.0f, xbn = 0.0f, xhat = 0.0f;
    if (wok) {
      const V8 d = load8(dx0, (((size_t)b * Tp + h) * Fp + w) * 32 + cv * 8, 0);
      if (h < T) {
        const float m = __ldg(mag + ((size_t)b * T + h) * F + w);
        xbn = fmaf(sc0, m, sh0);
        xhat = (m - mean0) * rstd0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ab[k] += d.v[k];
        aw[k] = fmaf(d.v[k], xbn, aw[k]);
        part = fmaf(d.v[k], pw.v[k], part);
      }
    }
    // dxbn of the pixel = sum over its four channel vectors (lanes cv = 0..3 are adjacent)
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (h < T) {
      dg = fmaf(part, xhat, dg);
      db += part;
    }
  }
  if (wok && cv == 0) {
    atomicAdd(&dgamma0[w], dg);
    atomicAdd(&dbeta0[w], db);
  }
  // channel sums: reduce over the 8 pixels of a warp (lanes with equal cv), then over the 8 warps through shared memory
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      aw[k] += __shfl_xor_sync(0xffffffffu, aw[k], o);
      ab[k] += __shfl_xor_sync(0xffffffffu, ab[k], o);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[0][warp][lane * 8 + k] = aw[k];
      red[1][warp][lane * 8 + k] = ab[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, c = threadIdx.x & 31;
    float s = 0.0f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) s += red[which][wp][c];
    atomicAdd(which == 0 ? &dpre_w[c] : &dpre_b[c], s);
  }
}

// after_conv (32 -> 3, 1x1) backward: thread = (pixel, 8-channel vector), 64 pixels per 256-thread block iteration
__global__ void __launch_bounds__(256) after_bwd_kernel(const float* __restrict__ dfeat, const void* __restrict__ y,
                                                        const float* __restrict__ after_w, void* __restrict__ dy,
                                                        float* __restrict__ dw, float* __restrict__ db, int B, long long npix) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  __shared__ float red[8][4][27];
  const int cv = threadIdx.x & 3;
  V8 w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = ldf8(after_w + k * 32 + cv * 8);
  float aw[3][8], ab[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[k][i] = 0.0f;
  // grid.y = clip; two pixels per thread and step so that eight loads are in flight
  const int b = blockIdx.y;
  const float* dfb = dfeat + (size_t)b * 3 * npix;
  const size_t pix0 = (size_t)b * npix;
  const long long stride = (long long)gridDim.x * 64;
  for (long long q0 = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); q0 < npix; q0 += 2 * stride) {
    float df[2][3];
    uint4 yq[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long q = q0 + u * stride;
      if (q < npix) {
#pragma unroll
        for (int k = 0; k < 3; ++k) df[u][k] = __ldg(dfb + (size_t)k * npix + q);
        yq[u] = loadq(y, (pix0 + q) * 32 + cv * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long q = q0 + u * stride;
      if (q < npix) {
        const V8 yv = unpack8(yq[u], 1);
        V8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          r.v[i] = df[u][0] * w[0].v[i] + df[u][1] * w[1].v[i] + df[u][2] * w[2].v[i];
#pragma unroll
          for (int k = 0; k < 3; ++k) aw[k][i] = fmaf(df[u][k], yv.v[i], aw[k][i]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) ab[k] += df[u][k];
        store8(dy, (pix0 + q) * 32 + cv * 8, 0, r);
      }
    }
  }
  // reduce over the 8 pixels of a warp (lanes with equal cv), then over warps
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ab[k] += __shfl_xor_sync(0xffffffffu, ab[k], o);
#pragma unroll
      for (int i = 0; i < 8; ++i) aw[k][i] += __shfl_xor_sync(0xffffffffu, aw[k][i], o);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane][k * 8 + i] = aw[k][i];
      red[warp][lane][24 + k] = ab[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 96) {
    const int k = threadIdx.x / 32, c = threadIdx.x % 32;
    float s = 0.0f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) s += red[wp][c >> 3][k * 8 + (c & 7)];
    atomicAdd(&dw[k * 32 + c], s);
  } else if (threadIdx.x < 99) {
    const int k = threadIdx.x - 96;
    float s = 0.0f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) s += red[wp][0][24 + k];
    atomicAdd(&db[k], s);
  }
}

// backward of feature_maps_to_wav's mask (reference models/resunet.py:457-505) for one (b, t, f < Fp)
__global__ void __launch_bounds__(256) mask_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ mag,
                                                       const float* __restrict__ cosp, const float* __restrict__ sinp,
                                                       const float* __restrict__ dre, const float* __restrict__ dim_,
                                                       float* __restrict__ dfeat, int B, int T, int F, int Tp, int Fp, float inv_n) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const long long total = (long long)B * Tp * Fp;
  const size_t plane = (size_t)Tp * Fp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % Fp);
    const long long tt = i / Fp;
    const int t = (int)(tt % Tp), b = (int)(tt / Tp);
    const size_t fo = (size_t)b * 3 * plane + (size_t)t * Fp + f;
    float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
    if (t < T) {
      const size_t so = ((size_t)b * T + t) * F + f;
      const float cf = (f == 0 ? 1.0f : 2.0f) * inv_n;      // bins 1 .. n/2-1 appear twice in the Hermitian extension (f < Fp = n/2)
      const float gre = __ldg(dre + so) * cf, gim = __ldg(dim_ + so) * cf;
      const float x0 = __ldg(feat + fo), x1 = __ldg(feat + fo + plane), x2 = __ldg(feat + fo + 2 * plane);
      const float mg = __ldg(mag + so), cs = __ldg(cosp + so), sn = __ldg(sinp + so);
      const float m = 1.0f / (1.0f + __expf(-x0));
      const float a = tanhf(x1), bb = tanhf(x2);
      const float r = sqrtf(a * a + bb * bb);
      const float rc = fmaxf(r, 1e-10f);
      const float mc = a / rc, ms = bb / rc;
      const float cy = cs * mc - sn * ms, sy = sn * mc + cs * ms;
      const float absy = mg * m;
      const float dabs = gre * cy + gim * sy;
      const float dcy = gre * absy, dsy = gim * absy;
      const float dmc = dcy * cs + dsy * sn, dms = -dcy * sn + dsy * cs;
      float da, dbb;
      if (r > 1e-10f) {
        const float inv3 = 1.0f / (rc * rc * rc);
        da = bb * inv3 * (dmc * bb - dms * a);
        dbb = a * inv3 * (dms * a - dmc * bb);
      } else {
        da = dmc * 1e10f;
        dbb = dms * 1e10f;
      }
      d0 = dabs * mg * m * (1.0f - m);
      d1 = da * (1.0f - a * a);
      d2 = dbb * (1.0f - bb * bb);
    }
    dfeat[fo] = d0;
    dfeat[fo + plane] = d1;
    dfeat[fo + 2 * plane] = d2;
  }
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ wave, const float* __restrict__ target, long long n,
                                                      float* __restrict__ loss_sum, float* __restrict__ dwave, float scale) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  __shared__ float red[8];
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = wave[i] - target[i];
    s += fabsf(d);
    dwave[i] = d > 0.0f ? scale : (d < 0.0f ? -scale : 0.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss_sum, t);
  }
}

// dw (J, K) = dbeta^T cond ; db (J) = column sums of dbeta.  Thread = (j, 4 consecutive k).
__global__ void __launch_bounds__(128) film_bwd_kernel(const float* __restrict__ dbeta, const float* __restrict__ cond,
                                                       float* __restrict__ dw, float* __restrict__ db, int B, int J, int K) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const int j = blockIdx.x;
  float s = 0.0f;
  for (int k0 = threadIdx.x * 4; k0 < K; k0 += 128 * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float d = __ldg(dbeta + (size_t)b * J + j);
      const float4 c = __ldg(reinterpret_cast<const float4*>(cond + (size_t)b * K + k0));
      acc.x = fmaf(d, c.x, acc.x);
      acc.y = fmaf(d, c.y, acc.y);
      acc.z = fmaf(d, c.z, acc.z);
      acc.w = fmaf(d, c.w, acc.w);
    }
    *reinterpret_cast<float4*>(dw + (size_t)j * K + k0) = acc;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < B; ++b) s += __ldg(dbeta + (size_t)b * J + j);
    db[j] = s;
  }
}

// Same operation order as torch.optim.adamw's single-tensor path (see oracle/train_oracle.py::adamw_amsgrad_step).
__device__ __forceinline__ void adamw_one(float& pi, float gi, float& mi, float& vi, float& vm, float decay, float one_minus_b1, float b2,
                                          float one_minus_b2, float bc2_sqrt, float eps, float neg_step_size, float grad_scale) {
  gi = __fmul_rn(gi, grad_scale);
  pi = __fmul_rn(pi, decay);
  mi = __fadd_rn(mi, __fmul_rn(one_minus_b1, __fsub_rn(gi, mi)));                         // lerp_
  vi = __fmul_rn(vi, b2);
  vi = __fadd_rn(vi, __fmul_rn(__fmul_rn(one_minus_b2, gi), gi));                         // addcmul_
  vm = fmaxf(vm, vi);
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vm), bc2_sqrt), eps);
  pi = __fadd_rn(pi, __fmul_rn(neg_step_size, __fdiv_rn(mi, denom)));                     // addcdiv_
}

// n4 float4 groups (16-byte loads / stores), then the scalar tail
__global__ void __launch_bounds__(256) adamw_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                            float* __restrict__ v, float* __restrict__ vmax, long long n, long long n4,
                                                            float decay, float one_minus_b1, float b2, float one_minus_b2, float bc2_sqrt,
                                                            float eps, float neg_step_size, float grad_scale) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i],
           xv = reinterpret_cast<float4*>(vmax)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    adamw_one(pv.x, gv.x, mv.x, vv.x, xv.x, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, grad_scale);
    adamw_one(pv.y, gv.y, mv.y, vv.y, xv.y, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, grad_scale);
    adamw_one(pv.z, gv.z, mv.z, vv.z, xv.z, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, grad_scale);
    adamw_one(pv.w, gv.w, mv.w, vv.w, xv.w, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, grad_scale);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(vmax)[i] = xv;
  }
  for (long long i = 4 * n4 + t0; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i], vm = vmax[i];
    adamw_one(pi, g[i], mi, vi, vm, decay, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, grad_scale);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    vmax[i] = vm;
  }
}

// fp32 parameter (torch layout) -> 16-bit kernel layouts; thread per source element
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int kind, int co, int ci, int taps,
                                                          void* __restrict__ fwd, int fwd_fp16, void* __restrict__ dgrad) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const long long n = (long long)co * ci * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    int o, c;
    size_t fi, di;
    if (kind == 0) {            // conv (co, ci, taps): fwd (taps, co, ci); dgrad (taps, ci, co) with flipped taps
      c = (int)(r % ci);
      o = (int)(r / ci);
      fi = ((size_t)t * co + o) * ci + c;
      di = ((size_t)(taps - 1 - t) * ci + c) * co + o;
    } else {                    // transposed conv (ci, co, taps): fwd (taps*co, ci); dgrad (ci, taps*co)
      o = (int)(r % co);
      c = (int)(r / co);
      fi = ((size_t)t * co + o) * ci + c;
      di = (size_t)c * taps * co + (size_t)t * co + o;
    }
    const float val = w[i];
    if (fwd) {
      if (fwd_fp16) reinterpret_cast<__half*>(fwd)[fi] = __float2half_rn(fminf(fmaxf(val, -65504.0f), 65504.0f));
      else reinterpret_cast<__nv_bfloat16*>(fwd)[fi] = __float2bfloat16_rn(val);
    }
    if (dgrad) reinterpret_cast<__nv_bfloat16*>(dgrad)[di] = __float2bfloat16_rn(val);
  }
}

// packed fp32 gradient (taps, co, ci) -> torch layout; thread per destination element
__global__ void __launch_bounds__(256) unpack_grad_kernel(const float* __restrict__ dw, int kind, int co, int ci, int taps,
                                                          float* __restrict__ grad) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const long long n = (long long)co * ci * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    int o, c;
    if (kind == 0) {
      c = (int)(r % ci);
      o = (int)(r / ci);
    } else {
      o = (int)(r % co);
      c = (int)(r / co);
    }
    grad[i] = dw[((size_t)t * co + o) * ci + c];
  }
}

// Multi-tensor forms of the two kernels above: ONE launch re-packs every convolution weight after the optimizer step (42 tensors)
// / un-packs every weight gradient of a bucket.  table: 8 x int64 per tensor
//   pack:   [w, fwd, dgrad, kind, co, ci, taps | fwd_fp16 << 16, first block]      unpack: [dw, grad, 0, kind, co, ci, taps, first block]
// unpack: a block handles kMultiChunk consecutive destination elements of its tensor (the source is the contiguous-in-ci packed
// layout, read through L2); pack: see below.
constexpr int kMultiChunk = 2048;
__device__ __forceinline__ int multi_find(const long long* __restrict__ table, int nitems, int block) {
  int lo = 0, hi = nitems - 1;
  while (lo < hi) {                       // last tensor whose first block <= block
    const int mid = (lo + hi + 1) >> 1;
    if ((int)table[mid * 8 + 7] <= block) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// pack: a block owns a 32 (outer) x 32 (inner) tile of the parameter's two leading dimensions with all its taps, staged through
// shared memory so that the read (taps innermost in the source) and both writes (ci innermost in fwd, co innermost in dgrad) move
// whole 64-byte runs: the thread-per-element form wrote 2-byte values at strides of co / ci elements (245 us for the model's 25.6 M
// weights; this form ~3x faster).  first block / block count per tensor: ceil(d0 / 32) * ceil(d1 / 32) (lass_pack_blocks).
constexpr int kPackTile = 32;
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const long long* __restrict__ table, int nitems) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  constexpr int kRow = kPackTile * 9 + 1;                // odd row pitch: lanes along a tile ROW hit 32 different banks
  __shared__ float tile[kPackTile * kRow];
  const int it = multi_find(table, nitems, blockIdx.x);
  const long long* e = table + it * 8;
  const float* w = reinterpret_cast<const float*>(e[0]);
  void* fwd = reinterpret_cast<void*>(e[1]);
  void* dgrad = reinterpret_cast<void*>(e[2]);
  const int kind = (int)e[3], co = (int)e[4], ci = (int)e[5], taps = (int)(e[6] & 0xffff), fwd_fp16 = (int)(e[6] >> 16);
  // source (d0, d1, taps): conv d0 = co, d1 = ci; transposed conv d0 = ci, d1 = co
  const int d0 = kind == 0 ? co : ci, d1 = kind == 0 ? ci : co;
  const int tiles1 = (d1 + kPackTile - 1) / kPackTile;
  const int tb = blockIdx.x - (int)e[7];
  const int r0 = (tb / tiles1) * kPackTile, c0 = (tb % tiles1) * kPackTile;
  const int nr = min(kPackTile, d0 - r0), nc = min(kPackTile, d1 - c0);
  const int row_len = nc * taps;                         // contiguous floats of one source row inside the tile
  if (taps > 9) {                                        // not a shape of this model: plain per-element path
    for (int i = threadIdx.x; i < nr * row_len; i += 256) {
      const int r = i / row_len, q = i - r * row_len, c = q / taps, t = q - c * taps;
      const float val = w[((size_t)(r0 + r) * d1 + c0) * taps + q];
      const int o = kind == 0 ? r0 + r : c0 + c, cc = kind == 0 ? c0 + c : r0 + r;
      const size_t fi = ((size_t)t * co + o) * ci + cc;
      const size_t di = kind == 0 ? ((size_t)(taps - 1 - t) * ci + cc) * co + o : (size_t)cc * taps * co + (size_t)t * co + o;
      if (fwd) {
        if (fwd_fp16) reinterpret_cast<__half*>(fwd)[fi] = __float2half_rn(fminf(fmaxf(val, -65504.0f), 65504.0f));
        else reinterpret_cast<__nv_bfloat16*>(fwd)[fi] = __float2bfloat16_rn(val);
      }
      if (dgrad) reinterpret_cast<__nv_bfloat16*>(dgrad)[di] = __float2bfloat16_rn(val);
    }
    return;
  }
  for (int i = threadIdx.x; i < nr * row_len; i += 256) {
    const int r = i / row_len, q = i - r * row_len;
    tile[r * kRow + q] = w[((size_t)(r0 + r) * d1 + c0) * taps + q];
  }
  __syncthreads();
  auto at = [&](int r, int c, int t) { return tile[r * kRow + c * taps + t]; };
  auto put_fwd = [&](size_t idx, float val) {
    if (fwd_fp16) reinterpret_cast<__half*>(fwd)[idx] = __float2half_rn(fminf(fmaxf(val, -65504.0f), 65504.0f));
    else reinterpret_cast<__nv_bfloat16*>(fwd)[idx] = __float2bfloat16_rn(val);
  };
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (kind == 0) {
    // fwd (taps, co, ci): lanes along ci (= tile column);  dgrad (taps, ci, co) with flipped taps: lanes along co (= tile row)
    if (fwd)
      for (int j = wrp; j < taps * nr; j += 8) {
        const int t = j / nr, r = j - t * nr;
        if (lane < nc) put_fwd(((size_t)t * co + r0 + r) * ci + c0 + lane, at(r, lane, t));
      }
    if (dgrad)
      for (int j = wrp; j < taps * nc; j += 8) {
        const int t = j / nc, c = j - t * nc;
        if (lane < nr)
          reinterpret_cast<__nv_bfloat16*>(dgrad)[((size_t)(taps - 1 - t) * ci + c0 + c) * co + r0 + lane] = __float2bfloat16_rn(at(lane, c, t));
      }
  } else {
    // source (ci, co, taps).  fwd (taps * co, ci): lanes along ci (= tile row);  dgrad (ci, taps * co): lanes along co (= tile column)
    if (fwd)
      for (int j = wrp; j < taps * nc; j += 8) {
        const int t = j / nc, c = j - t * nc;
        if (lane < nr) put_fwd(((size_t)t * co + c0 + c) * ci + r0 + lane, at(lane, c, t));
      }
    if (dgrad)
      for (int j = wrp; j < taps * nr; j += 8) {
        const int t = j / nr, r = j - t * nr;
        if (lane < nc)
          reinterpret_cast<__nv_bfloat16*>(dgrad)[(size_t)(r0 + r) * taps * co + (size_t)t * co + c0 + lane] = __float2bfloat16_rn(at(r, lane, t));
      }
  }
}

__global__ void __launch_bounds__(256) unpack_grads_multi_kernel(const long long* __restrict__ table, int nitems) {
  griddep_launch_dependents();   // programmatic dependent launch: see launch_pdl (lass_internal.cuh)
  griddep_wait();
  const int it = multi_find(table, nitems, blockIdx.x);
  const long long* e = table + it * 8;
  const float* dw = reinterpret_cast<const float*>(e[0]);
  float* grad = reinterpret_cast<float*>(e[1]);
  const int kind = (int)e[3], co = (int)e[4], ci = (int)e[5], taps = (int)e[6];
  const long long n = (long long)co * ci * taps;
  const long long i0 = (long long)(blockIdx.x - (int)e[7]) * kMultiChunk;
  for (long long i = i0 + threadIdx.x; i < i0 + kMultiChunk && i < n; i += 256) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    int o, c;
    if (kind == 0) {
      c = (int)(r % ci);
      o = (int)(r / ci);
    } else {
      o = (int)(r % co);
      c = (int)(r / co);
    }
    grad[i] = dw[((size_t)t * co + o) * ci + c];
  }
}

int grid_for(long long work_items, int threads, int max_blocks = 148 * 8) {
  long long b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

bool chan_ok(int C, int cstride, int coff) { return C > 0 && C % 8 == 0 && cstride % 8 == 0 && coff % 8 == 0 && coff + C <= cstride && kRedThreads % (C / 8) == 0; }

}  // namespace
}  // namespace lass

using namespace lass;

#define LASS_LAUNCH_CHECK(what) return set_cuda_error(cudaGetLastError(), what)

extern "C" {

int lass_bn_stats(const void* x, int fp16, long long npix, int C, int cstride, int coff, double* sums, void* stream_v) {
  if (!x || !sums || npix <= 0 || !chan_ok(C, cstride, coff)) return set_error(LASS_ERR_ARG, "lass_bn_stats: bad argument (C=%d cstride=%d coff=%d)", C, cstride, coff);
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
  const int R = kRedThreads / (C / 8);
  const int grid = grid_for(npix, R * 8, 148 * 4);
  launch_pdl(bn_stats_kernel, grid, kRedThreads, (size_t)R * C * 2 * sizeof(float), s, x, fp16, npix, C, cstride, coff, sums);
  LASS_LAUNCH_CHECK("bn_stats launch");
}

// the same reduction into sums the CALLER zeroed (one memset for all of a step's sites)
int lass_bn_stats_acc(const void* x, int fp16, long long npix, int C, int cstride, int coff, double* sums, void* stream_v) {
  if (!x || !sums || npix <= 0 || !chan_ok(C, cstride, coff)) return set_error(LASS_ERR_ARG, "lass_bn_stats_acc: bad argument (C=%d cstride=%d coff=%d)", C, cstride, coff);
  const int R = kRedThreads / (C / 8);
  const int grid = grid_for(npix, R * 8, 148 * 4);
  launch_pdl(bn_stats_kernel, grid, kRedThreads, (size_t)R * C * 2 * sizeof(float), (cudaStream_t)stream_v, x, fp16, npix, C, cstride, coff, sums);
  LASS_LAUNCH_CHECK("bn_stats launch");
}

int lass_bn0_stats(const float* mag, int B, int T, int F, double* sums, void* stream_v) {
  if (!mag || !sums || B <= 0 || T <= 0 || F <= 0) return set_error(LASS_ERR_ARG, "lass_bn0_stats: bad argument");
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * F, s);
  const int rows = B * T, rpb = 64;
  dim3 grid((unsigned)((F + 127) / 128), (unsigned)((rows + rpb - 1) / rpb));
  launch_pdl(bn0_stats_kernel, grid, 128, 0, s, mag, rows, F, rpb, sums);
  LASS_LAUNCH_CHECK("bn0_stats launch");
}

int lass_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int C, float* bnp, void* stream_v) {
  if (!sums || !gamma || !beta || !running_mean || !running_var || !bnp || C <= 0 || count <= 0) return set_error(LASS_ERR_ARG, "lass_bn_finalize: bad argument");
  launch_pdl(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_v, sums, count, gamma, beta, running_mean, running_var, momentum, eps, C, bnp);
  LASS_LAUNCH_CHECK("bn_finalize launch");
}

int lass_bn_act(const void* x, int x_fp16, int x_cstride, int x_coff, void* out, int out_fp16, int out_cstride, int out_coff, int B,
                long long pix_per_clip, int C, const float* bnp, const float* beta, int beta_bstride, void* stream_v) {
  if (!x || !out || !bnp || !beta || B <= 0 || pix_per_clip <= 0 || !chan_ok(C, x_cstride, x_coff) || !chan_ok(C, out_cstride, out_coff) || beta_bstride % 4)
    return set_error(LASS_ERR_ARG, "lass_bn_act: bad argument");
  const int R = kRedThreads / (C / 8);
  dim3 grid((unsigned)grid_for(pix_per_clip, R * 4, (148 * 8 + B - 1) / B), (unsigned)B);
  launch_pdl(bn_act_kernel, grid, kRedThreads, 0, (cudaStream_t)stream_v, x, x_fp16, x_cstride, x_coff, out, out_fp16, out_cstride, out_coff,
                                                                  pix_per_clip, C, bnp, beta, beta_bstride);
  LASS_LAUNCH_CHECK("bn_act launch");
}

static int bn_bwd_reduce_launch(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride, int x_coff, int B,
                                long long pix_per_clip, int C, const float* bnp, const float* beta, int beta_bstride, float* sums,
                                const BnBwdFinalize& fin, cudaStream_t s) {
  const int R = kRedThreads / (C / 8);
  int gx = grid_for(pix_per_clip, R * 8, (148 * 4 + B - 1) / B);
  dim3 grid((unsigned)gx, (unsigned)B);
  launch_pdl(bn_bwd_reduce_kernel, grid, kRedThreads, (size_t)R * C * 2 * sizeof(float), s, dact, d_cstride, d_coff, x, x_fp16, x_cstride, x_coff,
                                                                                    pix_per_clip, C, bnp, beta, beta_bstride, sums, fin);
  LASS_LAUNCH_CHECK("bn_bwd_reduce launch");
}

int lass_bn_bwd_reduce(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride, int x_coff, int B,
                       long long pix_per_clip, int C, const float* bnp, const float* beta, int beta_bstride, float* sums, void* stream_v) {
  if (!dact || !x || !bnp || !beta || !sums || B <= 0 || pix_per_clip <= 0 || !chan_ok(C, x_cstride, x_coff) || !chan_ok(C, d_cstride, d_coff) || beta_bstride % 4)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_reduce: bad argument");
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)B * C, s);
  BnBwdFinalize fin;
  memset(&fin, 0, sizeof(fin));
  return bn_bwd_reduce_launch(dact, d_cstride, d_coff, x, x_fp16, x_cstride, x_coff, B, pix_per_clip, C, bnp, beta, beta_bstride, sums, fin, s);
}

// the same reduction into sums the CALLER zeroed (one memset for all of a step's sites)
int lass_bn_bwd_reduce_acc(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride, int x_coff, int B,
                           long long pix_per_clip, int C, const float* bnp, const float* beta, int beta_bstride, float* sums, void* stream_v) {
  if (!dact || !x || !bnp || !beta || !sums || B <= 0 || pix_per_clip <= 0 || !chan_ok(C, x_cstride, x_coff) || !chan_ok(C, d_cstride, d_coff) || beta_bstride % 4)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_reduce_acc: bad argument");
  BnBwdFinalize fin;
  memset(&fin, 0, sizeof(fin));
  return bn_bwd_reduce_launch(dact, d_cstride, d_coff, x, x_fp16, x_cstride, x_coff, B, pix_per_clip, C, bnp, beta, beta_bstride, sums, fin,
                              (cudaStream_t)stream_v);
}

int lass_bn_bwd_reduce_finalize(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride, int x_coff, int B,
                                long long pix_per_clip, int C, float* bnp, const float* beta, int beta_bstride, float* sums,
                                unsigned int* counter, const float* gamma, float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride,
                                void* stream_v) {
  if (!dact || !x || !bnp || !beta || !sums || !counter || !gamma || !dgamma || !dbeta || B <= 0 || pix_per_clip <= 0 ||
      !chan_ok(C, x_cstride, x_coff) || !chan_ok(C, d_cstride, d_coff) || beta_bstride % 4)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_reduce_finalize: bad argument");
  BnBwdFinalize fin;
  fin.counter = counter;
  fin.count = (double)B * (double)pix_per_clip;
  fin.gamma = gamma;
  fin.bnp = bnp;
  fin.dgamma = dgamma;
  fin.dbeta = dbeta;
  fin.dfilm = dfilm;
  fin.dfilm_bstride = dfilm_bstride;
  return bn_bwd_reduce_launch(dact, d_cstride, d_coff, x, x_fp16, x_cstride, x_coff, B, pix_per_clip, C, bnp, beta, beta_bstride, sums, fin,
                              (cudaStream_t)stream_v);
}

int lass_bn_bwd_finalize(const float* sums, int B, int C, double count, const float* gamma, float* bnp, float* dgamma, float* dbeta,
                         float* dfilm, int dfilm_bstride, void* stream_v) {
  if (!sums || !gamma || !bnp || !dgamma || !dbeta || B <= 0 || C <= 0 || count <= 0) return set_error(LASS_ERR_ARG, "lass_bn_bwd_finalize: bad argument");
  launch_pdl(bn_bwd_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_v, sums, B, C, count, gamma, bnp, dgamma, dbeta, dfilm, dfilm_bstride);
  LASS_LAUNCH_CHECK("bn_bwd_finalize launch");
}

int lass_bn_bwd_totals(const float* sums, int B, int C, double* totals, void* stream_v) {
  if (!sums || !totals || B <= 0 || C <= 0) return set_error(LASS_ERR_ARG, "lass_bn_bwd_totals: bad argument");
  launch_pdl(bn_bwd_totals_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_v, sums, B, C, totals);
  LASS_LAUNCH_CHECK("bn_bwd_totals launch");
}

int lass_bn_bwd_finalize_sync(const float* sums, int B, int C, double count_total, const double* totals, const float* gamma, float* bnp,
                              float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride, void* stream_v) {
  if (!sums || !totals || !gamma || !bnp || !dgamma || !dbeta || B <= 0 || C <= 0 || count_total <= 0)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_finalize_sync: bad argument");
  launch_pdl(bn_bwd_finalize_sync_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_v, sums, B, C, count_total, totals, bnp, dgamma, dbeta,
             dfilm, dfilm_bstride);
  LASS_LAUNCH_CHECK("bn_bwd_finalize_sync launch");
}

static int fill_peer_table(PeerTable* pt, const void* const* peer_sums, void* const* peer_flags, int world, int rank, const char* who) {
  if (!peer_sums || !peer_flags || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return set_error(LASS_ERR_ARG, "%s: bad peer table (world %d, rank %d, at most %d peers)", who, world, rank, kMaxPeers);
  memset(pt, 0, sizeof(*pt));
  for (int p = 0; p < world; ++p) {
    if (!peer_sums[p] || !peer_flags[p]) return set_error(LASS_ERR_ARG, "%s: null peer pointer %d", who, p);
    pt->sums[p] = peer_sums[p];
    pt->flags[p] = reinterpret_cast<unsigned long long*>(peer_flags[p]);
  }
  return LASS_OK;
}

int lass_syncbn_max_peers(void) { return kMaxPeers; }

int lass_bn_finalize_p2p(const void* const* peer_sums, void* const* peer_flags, int world, int rank, long long sums_offset, int flag_index,
                         unsigned long long epoch, double count_total, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, int C, float* bnp, int* status, void* stream_v) {
  PeerTable pt;
  if (int e = fill_peer_table(&pt, peer_sums, peer_flags, world, rank, "lass_bn_finalize_p2p")) return e;
  if (!gamma || !beta || !running_mean || !running_var || !bnp || C <= 0 || count_total <= 0 || sums_offset < 0 || flag_index < 0 || epoch == 0)
    return set_error(LASS_ERR_ARG, "lass_bn_finalize_p2p: bad argument");
  launch_pdl(bn_finalize_p2p_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_v, pt, world, rank, sums_offset, flag_index, epoch, count_total,
             gamma, beta, running_mean, running_var, momentum, eps, C, bnp, status);
  LASS_LAUNCH_CHECK("bn_finalize_p2p launch");
}

int lass_bn_bwd_finalize_p2p(const void* const* peer_sums, void* const* peer_flags, int world, int rank, long long sums_offset,
                             long long totals_offset, int flag_index, unsigned long long epoch, int B, int C, double count_total,
                             const float* gamma, float* bnp, float* dgamma, float* dbeta, float* dfilm, int dfilm_bstride, int* status,
                             void* stream_v) {
  PeerTable pt;
  if (int e = fill_peer_table(&pt, peer_sums, peer_flags, world, rank, "lass_bn_bwd_finalize_p2p")) return e;
  if (!gamma || !bnp || !dgamma || !dbeta || B <= 0 || C <= 0 || C > kP2PBwdThreads * kP2PBwdMaxPerThread || count_total <= 0 || sums_offset < 0 ||
      (sums_offset & 1) || totals_offset < 0 || flag_index < 0 || epoch == 0)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_finalize_p2p: bad argument");
  launch_pdl(bn_bwd_finalize_p2p_kernel, 1, kP2PBwdThreads, 0, (cudaStream_t)stream_v, pt, world, rank, sums_offset, totals_offset, flag_index, epoch,
             B, C, count_total, bnp, dgamma, dbeta, dfilm, dfilm_bstride, status);
  LASS_LAUNCH_CHECK("bn_bwd_finalize_p2p launch");
}

int lass_bn_bwd_apply(const void* dact, int d_cstride, int d_coff, const void* x, int x_fp16, int x_cstride, int x_coff, const void* add,
                      int add_cstride, int add_coff, void* dx, int dx_cstride, int dx_coff, int B, long long pix_per_clip, int C,
                      const float* bnp, const float* beta, int beta_bstride, void* stream_v) {
  if (!dact || !x || !dx || !bnp || !beta || B <= 0 || pix_per_clip <= 0 || !chan_ok(C, x_cstride, x_coff) || !chan_ok(C, d_cstride, d_coff) ||
      !chan_ok(C, dx_cstride, dx_coff) || (add && !chan_ok(C, add_cstride, add_coff)) || beta_bstride % 4)
    return set_error(LASS_ERR_ARG, "lass_bn_bwd_apply: bad argument");
  const int R = kRedThreads / (C / 8);
  dim3 grid((unsigned)grid_for(pix_per_clip, R * 2, (148 * 8 + B - 1) / B), (unsigned)B);
  launch_pdl(bn_bwd_apply_kernel, grid, kRedThreads, 0, (cudaStream_t)stream_v, dact, d_cstride, d_coff, x, x_fp16, x_cstride, x_coff, add, add_cstride,
                                                                        add_coff, dx, dx_cstride, dx_coff, pix_per_clip, C, bnp, beta,
                                                                        beta_bstride);
  LASS_LAUNCH_CHECK("bn_bwd_apply launch");
}

int lass_pool_bwd(const void* dpool, const void* dskip, int dskip_cstride, int dskip_coff, void* dy, int B, int H, int W, int C, int ph,
                  int pw, void* stream_v) {
  if (!dpool || !dy || B <= 0 || H <= 0 || W <= 0 || C % 8 || ph < 1 || pw < 1 || H % ph || W % pw || (dskip && !chan_ok(C, dskip_cstride, dskip_coff)))
    return set_error(LASS_ERR_ARG, "lass_pool_bwd: bad argument");
  const int nv = W * (C / 8);
  launch_pdl(pool_bwd_kernel, dim3((unsigned)grid_for(nv, 256 * 4, 64), (unsigned)(B * H < 65535 ? B * H : 65535)), 256, 0, (cudaStream_t)stream_v, dpool, dskip, dskip_cstride, dskip_coff, dy, B, H, W, C, ph, pw);
  LASS_LAUNCH_CHECK("pool_bwd launch");
}

int lass_unshuffle(const void* src, int src_cstride, int src_coff, void* dst, int B, int H, int W, int C, int uh, int uw, void* stream_v) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || uh < 1 || uw < 1 || C % 8 || src_cstride % 8 || src_coff % 8 || src_coff + C > src_cstride)
    return set_error(LASS_ERR_ARG, "lass_unshuffle: bad argument");
  const int nv = W * uh * uw * (C / 8);
  launch_pdl(unshuffle_kernel, dim3((unsigned)grid_for(nv, 256 * 4, 64), (unsigned)(B * H < 65535 ? B * H : 65535)), 256, 0, (cudaStream_t)stream_v, 
      reinterpret_cast<const uint16_t*>(src), src_cstride, src_coff, reinterpret_cast<uint16_t*>(dst), B, H, W, C, uh, uw);
  LASS_LAUNCH_CHECK("unshuffle launch");
}

int lass_channel_sum_acc(const void* x, long long npix, int C, int cstride, int coff, float* out, void* stream_v) {
  if (!x || !out || npix <= 0 || !chan_ok(C, cstride, coff)) return set_error(LASS_ERR_ARG, "lass_channel_sum_acc: bad argument");
  const int R = kRedThreads / (C / 8);
  launch_pdl(channel_sum_kernel, grid_for(npix, R * 8, 148 * 4), kRedThreads, (size_t)R * C * sizeof(float), (cudaStream_t)stream_v, x, npix, C, cstride,
             coff, out);
  LASS_LAUNCH_CHECK("channel_sum launch");
}

int lass_channel_sum(const void* x, long long npix, int C, int cstride, int coff, float* out, void* stream_v) {
  if (!x || !out || npix <= 0 || !chan_ok(C, cstride, coff)) return set_error(LASS_ERR_ARG, "lass_channel_sum: bad argument");
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(out, 0, sizeof(float) * C, s);
  const int R = kRedThreads / (C / 8);
  launch_pdl(channel_sum_kernel, grid_for(npix, R * 8, 148 * 4), kRedThreads, (size_t)R * C * sizeof(float), s, x, npix, C, cstride, coff, out);
  LASS_LAUNCH_CHECK("channel_sum launch");
}

int lass_pre_fwd(const float* mag, int B, int T, int F, int Tp, int Fp, const float* bnp0, const float* pre_w, const float* pre_b, void* x0,
                 void* stream_v) {
  if (!mag || !bnp0 || !pre_w || !pre_b || !x0 || B <= 0 || T <= 0 || Tp < T || Fp <= 0 || Fp > F) return set_error(LASS_ERR_ARG, "lass_pre_fwd: bad argument");
  launch_pdl(pre_fwd_kernel, dim3((unsigned)grid_for(Fp, 64 * 4, 16), (unsigned)(B * Tp < 65535 ? B * Tp : 65535)), 256, 0, (cudaStream_t)stream_v, mag, B, T, F, Tp, Fp, bnp0, pre_w,
                                                                                                         pre_b, x0);
  LASS_LAUNCH_CHECK("pre_fwd launch");
}

int lass_pre_bwd(const void* dx0, const float* mag, int B, int T, int F, int Tp, int Fp, const float* bnp0, const float* pre_w, float* dpre_w,
                 float* dpre_b, float* dgamma0, float* dbeta0, void* stream_v) {
  if (!dx0 || !mag || !bnp0 || !pre_w || !dpre_w || !dpre_b || !dgamma0 || !dbeta0 || B <= 0 || T <= 0 || Tp < T || Fp <= 0 || Fp > F)
    return set_error(LASS_ERR_ARG, "lass_pre_bwd: bad argument");
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(dpre_w, 0, 32 * sizeof(float), s);
  cudaMemsetAsync(dpre_b, 0, 32 * sizeof(float), s);
  cudaMemsetAsync(dgamma0, 0, F * sizeof(float), s);
  cudaMemsetAsync(dbeta0, 0, F * sizeof(float), s);
  dim3 grid((unsigned)((Fp + 63) / 64), (unsigned)((Tp + kPreRows - 1) / kPreRows), (unsigned)B);
  launch_pdl(pre_bwd_kernel, grid, 256, 0, s, dx0, mag, B, T, F, Tp, Fp, bnp0, pre_w, dpre_w, dpre_b, dgamma0, dbeta0);
  LASS_LAUNCH_CHECK("pre_bwd launch");
}

int lass_after_bwd(const float* dfeat, const void* y, const float* after_w, void* dy, float* dw, float* db, int B, long long npix, void* stream_v) {
  if (!dfeat || !y || !after_w || !dy || !dw || !db || B <= 0 || npix <= 0) return set_error(LASS_ERR_ARG, "lass_after_bwd: bad argument");
  cudaStream_t s = (cudaStream_t)stream_v;
  cudaMemsetAsync(dw, 0, 96 * sizeof(float), s);
  cudaMemsetAsync(db, 0, 3 * sizeof(float), s);
  launch_pdl(after_bwd_kernel, dim3((unsigned)grid_for(npix, 64 * 2, (148 * 4 + B - 1) / B), (unsigned)B), 256, 0, s, dfeat, y, after_w, dy, dw, db, B, npix);
  LASS_LAUNCH_CHECK("after_bwd launch");
}

int lass_mask_bwd(const float* feat, const float* mag, const float* cos, const float* sin, const float* dre, const float* dim, float* dfeat,
                  int B, int T, int F, int Tp, int Fp, int n_fft, void* stream_v) {
  if (!feat || !mag || !cos || !sin || !dre || !dim || !dfeat || B <= 0 || T <= 0 || Tp < T || F != n_fft / 2 + 1 || Fp != n_fft / 2)
    return set_error(LASS_ERR_ARG, "lass_mask_bwd: bad argument");
  const long long total = (long long)B * Tp * Fp;
  launch_pdl(mask_bwd_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream_v, feat, mag, cos, sin, dre, dim, dfeat, B, T, F, Tp, Fp, 1.0f / (float)n_fft);
  LASS_LAUNCH_CHECK("mask_bwd launch");
}

int lass_l1_loss(const float* wave, const float* target, long long n, float* loss_sum, float* dwave, float scale, void* stream_v) {
  if (!wave || !target || !loss_sum || !dwave || n <= 0) return set_error(LASS_ERR_ARG, "lass_l1_loss: bad argument");
  launch_pdl(l1_loss_kernel, grid_for(n, 256, 148 * 4), 256, 0, (cudaStream_t)stream_v, wave, target, n, loss_sum, dwave, scale);
  LASS_LAUNCH_CHECK("l1_loss launch");
}

int lass_film_bwd(const float* dbeta, const float* cond, float* dw, float* db, int B, int J, int K, void* stream_v) {
  if (!dbeta || !cond || !dw || !db || B <= 0 || J <= 0 || K <= 0 || K % 4) return set_error(LASS_ERR_ARG, "lass_film_bwd: bad argument");
  launch_pdl(film_bwd_kernel, J, 128, 0, (cudaStream_t)stream_v, dbeta, cond, dw, db, B, J, K);
  LASS_LAUNCH_CHECK("film_bwd launch");
}

int lass_adamw_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float grad_scale, void* stream_v) {
  if (!p || !g || !m || !v || !vmax || n <= 0 || step < 1) return set_error(LASS_ERR_ARG, "lass_adamw_amsgrad: bad argument");
  // host scalars exactly as torch.optim.adamw computes them (python floats = doubles, then cast)
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float decay = (float)(1.0 - (double)lr * (double)weight_decay);
  const float neg_step = (float)(-((double)lr / bc1));
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vmax)) & 15) == 0;
  const long long n4 = aligned ? n / 4 : 0;
  launch_pdl(adamw_amsgrad_kernel, grid_for(n4 > 0 ? n4 : n, 256), 256, 0, (cudaStream_t)stream_v, p, g, m, v, vmax, n, n4, decay, (float)(1.0 - (double)beta1), beta2,
                                                                             (float)(1.0 - (double)beta2), (float)sqrt(bc2), eps, neg_step,
                                                                             grad_scale);
  LASS_LAUNCH_CHECK("adamw launch");
}

int lass_pack_weight(const float* w, int kind, int co, int ci, int taps, void* fwd, int fwd_fp16, void* dgrad, void* stream_v) {
  if (!w || (!fwd && !dgrad) || co <= 0 || ci <= 0 || taps <= 0 || (kind != 0 && kind != 1)) return set_error(LASS_ERR_ARG, "lass_pack_weight: bad argument");
  launch_pdl(pack_weight_kernel, grid_for((long long)co * ci * taps, 256), 256, 0, (cudaStream_t)stream_v, w, kind, co, ci, taps, fwd, fwd_fp16, dgrad);
  LASS_LAUNCH_CHECK("pack_weight launch");
}

int lass_unpack_grad(const float* dw, int kind, int co, int ci, int taps, float* grad, void* stream_v) {
  if (!dw || !grad || co <= 0 || ci <= 0 || taps <= 0 || (kind != 0 && kind != 1)) return set_error(LASS_ERR_ARG, "lass_unpack_grad: bad argument");
  launch_pdl(unpack_grad_kernel, grid_for((long long)co * ci * taps, 256), 256, 0, (cudaStream_t)stream_v, dw, kind, co, ci, taps, grad);
  LASS_LAUNCH_CHECK("unpack_grad launch");
}

int lass_pack_weights_multi(const long long* table_dev, int nitems, int nblocks, void* stream_v) {
  if (!table_dev || nitems <= 0 || nblocks <= 0) return set_error(LASS_ERR_ARG, "lass_pack_weights_multi: bad argument");
  launch_pdl(pack_weights_multi_kernel, nblocks, 256, 0, (cudaStream_t)stream_v, table_dev, nitems);
  LASS_LAUNCH_CHECK("pack_weights_multi launch");
}

int lass_unpack_grads_multi(const long long* table_dev, int nitems, int nblocks, void* stream_v) {
  if (!table_dev || nitems <= 0 || nblocks <= 0) return set_error(LASS_ERR_ARG, "lass_unpack_grads_multi: bad argument");
  launch_pdl(unpack_grads_multi_kernel, nblocks, 256, 0, (cudaStream_t)stream_v, table_dev, nitems);
  LASS_LAUNCH_CHECK("unpack_grads_multi launch");
}

int lass_multi_chunk(void) { return kMultiChunk; }
int lass_pack_blocks(int kind, int co, int ci) {
  const int d0 = kind == 0 ? co : ci, d1 = kind == 0 ? ci : co;
  return ((d0 + kPackTile - 1) / kPackTile) * ((d1 + kPackTile - 1) / kPackTile);
}

}  // extern "C"
