// SegmentMixer (reference data/waveform_mixers.py:9-62 + dynamic_loudnorm / get_energy_ratio :65-95): the step in front of the
// training path (models/audiosep.py:76-78).  Per clip n: noise = sum_i gain_i * w[(n+i) % B] / ratio(w[(n+i) % B], w[n]),
// noise = gain * noise / ratio(noise, w[n]), mixture = w[n] + noise, both scaled by 0.9 / max|mixture| when that exceeds 1.
//
// The reference runs a Python loop over the batch with ~6 elementwise torch ops and 2 reductions per mixed-in clip.  Here it is
// two launches over a batch that lives in L2 (16 x 80 000 fp32 = 5 MB):
//   mixer_energy_kernel : sum x^2 of every clip as kSlices partials (summed in a FIXED order by the consumer: deterministic)
//   mixer_mix_kernel    : one thread-block CLUSTER of kCluster CTAs per clip.  The two clip-wide reductions the algorithm needs
//                         (energy of the summed noise, max |mixture|) are exchanged through distributed shared memory: every CTA
//                         writes its partial into all CTAs' tables, one cluster barrier, every CTA sums the same values in the
//                         same order.  A CTA keeps its slice of the summed noise in shared memory between the three passes
//                         (clips longer than 8 x 51 200 samples recompute it from the L2-resident sources instead): nothing
//                         but the two outputs is ever written to memory.
// The random draws (mix_num, the loudness offsets) stay on the host in the reference's order (lass_b200/waveform_mixers.py);
// they arrive as one small device table.
#include <cooperative_groups.h>

#include "lass_internal.cuh"

namespace cg = cooperative_groups;

namespace lass {

namespace {

constexpr int kSlices = 16;     // energy partials per clip
constexpr int kCluster = 8;     // CTAs per clip in mixer_mix_kernel
constexpr int kThreads = 256;    // mixer_energy_kernel
constexpr int kMixThreads = 1024; // mixer_mix_kernel: 128 CTAs for a 16-clip batch, so latency is hidden by warps per SM
constexpr int kMaxTerms = 16;   // max_mix_num - 1 mixed-in clips at most

// block-wide sum / max in a fixed order (warp shuffles, then the 8 warp results serially): deterministic
template <bool kMax, int kT = kThreads>
__device__ __forceinline__ float block_reduce(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = kMax ? fmaxf(v, u) : v + u;
  }
  __syncthreads();                       // sh may still be read by a previous reduction
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
#pragma unroll
  for (int w = 1; w < kT / 32; ++w) r = kMax ? fmaxf(r, sh[w]) : r + sh[w];
  return r;
}

__global__ void __launch_bounds__(kThreads) mixer_energy_kernel(const float* __restrict__ wave, int L, float* __restrict__ partial) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float sh[kThreads / 32];
  const int n = blockIdx.y, s = blockIdx.x;
  const float* x = wave + (size_t)n * L;
  const int per = (((L + kSlices - 1) / kSlices) + 3) & ~3;
  const int lo = min(L, s * per), hi = min(L, lo + per);
  float acc = 0.0f;
  if (((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
    const int hi4 = lo + ((hi - lo) & ~3);
    for (int i = lo + 4 * threadIdx.x; i < hi4; i += 4 * kThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
      acc = fmaf(v.x, v.x, acc);
      acc = fmaf(v.y, v.y, acc);
      acc = fmaf(v.z, v.z, acc);
      acc = fmaf(v.w, v.w, acc);
    }
    for (int i = hi4 + threadIdx.x; i < hi; i += kThreads) acc = fmaf(x[i], x[i], acc);
  } else {
    for (int i = lo + threadIdx.x; i < hi; i += kThreads) acc = fmaf(x[i], x[i], acc);
  }
  const float r = block_reduce<false>(acc, sh);
  if (threadIdx.x == 0) partial[n * kSlices + s] = r;
}

__device__ __forceinline__ float clip_energy(const float* partial, int k, float inv_len) {
  float e = 0.0f;
#pragma unroll
  for (int s = 0; s < kSlices; ++s) e += partial[k * kSlices + s];
  return e * inv_len;                     // torch.mean(x ** 2) (get_energy, data/waveform_mixers.py:72-73)
}
// get_energy_ratio (data/waveform_mixers.py:76-82): ((e1 / max(e2, 1e-10)) ** 0.5).clamp(0.02, 50)
__device__ __forceinline__ float energy_ratio(float e1, float e2) {
  return fminf(fmaxf(sqrtf(e1 / fmaxf(e2, 1e-10f)), 0.02f), 50.0f);
}

struct Terms {
  const float* src[kMaxTerms];
  float ratio[kMaxTerms], gain[kMaxTerms];
  int n;
};

// V samples per access: 4 (16-byte loads / stores; every row 16-byte aligned and slices whole float4s) or 1
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&o)[V]) {
  if constexpr (V == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  } else {
    o[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void ldsv(const float* p, float (&o)[V]) {          // shared memory
  if constexpr (V == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  } else {
    o[0] = p[0];
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&o)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    p[0] = o[0];
  }
}

// The mixed-in clips of one segment with their ratios / gains in REGISTERS (NT = 1, 2, 4 slots, unrolled; NT = 0: any number,
// read from the shared-memory table).  noise sample exactly in the reference's operation order:
// noise = 0; noise += gain * (next / ratio)   (data/waveform_mixers.py:38-41, :65-69, :93)
template <int NT>
struct Mixed {
  const float* src[NT ? NT : 1];
  float ratio[NT ? NT : 1], gain[NT ? NT : 1];
  const Terms* t;
  int n;
  __device__ __forceinline__ explicit Mixed(const Terms& tt) : t(&tt), n(tt.n) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const bool on = k < n;
      src[k] = tt.src[on ? k : 0];
      ratio[k] = on ? tt.ratio[k] : 1.0f;
      gain[k] = on ? tt.gain[k] : 0.0f;
    }
  }
  template <int V>
  __device__ __forceinline__ void noise(int i, float (&a)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = 0.0f;
    if constexpr (NT > 0) {
      float x[NT][V];
#pragma unroll
      for (int k = 0; k < NT; ++k)
        if (k < n) ldv<V>(src[k] + i, x[k]);
#pragma unroll
      for (int k = 0; k < NT; ++k)
        if (k < n) {
#pragma unroll
          for (int j = 0; j < V; ++j) a[j] += gain[k] * (x[k][j] / ratio[k]);
        }
    } else {
      for (int k = 0; k < n; ++k) {
        float x[V];
        ldv<V>(t->src[k] + i, x);
        const float g = t->gain[k], r = t->ratio[k];
#pragma unroll
        for (int j = 0; j < V; ++j) a[j] += g * (x[j] / r);
      }
    }
  }
};

// every CTA of the cluster contributes `mine`; returns the clip-wide sum / max, identical in all of them
template <bool kMax>
__device__ __forceinline__ float cluster_reduce(cg::cluster_group& cluster, float mine, float* table) {
  if (threadIdx.x < kCluster) {
    float* remote = cluster.map_shared_rank(table, threadIdx.x);
    remote[cluster.block_rank()] = mine;
  }
  cluster.sync();
  float r = table[0];
#pragma unroll
  for (int c = 1; c < kCluster; ++c) r = kMax ? fmaxf(r, table[c]) : r + table[c];
  return r;
}

struct MixShared {
  float sh[kMixThreads / 32];
  float tab_energy[kCluster], tab_max[kCluster];
  Terms t;
};

// The three passes of one CTA over its slice [lo, hi) of clip n.
template <int V, bool kCache, int NT>
__device__ __forceinline__ void mix_passes(cg::cluster_group& cluster, MixShared& S, float* cache, const float* __restrict__ seg,
                                           float* __restrict__ out_m, float* __restrict__ out_s, int lo, int hi, float e_seg,
                                           float gain_noise, float inv_len) {
  const Mixed<NT> mixed(S.t);
  constexpr int kStep = V * kMixThreads;
  const int first = lo + V * (int)threadIdx.x;

  // pass 1: energy of the summed noise over the whole clip
  float acc = 0.0f;
#pragma unroll 2
  for (int i = first; i < hi; i += kStep) {
    float v[V];
    mixed.template noise<V>(i, v);
    if (kCache) stv<V>(cache + (i - lo), v);
#pragma unroll
    for (int j = 0; j < V; ++j) acc = fmaf(v[j], v[j], acc);
  }
  const float e_noise = cluster_reduce<false>(cluster, block_reduce<false, kMixThreads>(acc, S.sh), S.tab_energy) * inv_len;
  const float ratio_noise = energy_ratio(e_noise, e_seg);

  // pass 2: max |segment + noise| over the whole clip (a thread reads back only the cache entries it wrote itself)
  float mx = 0.0f;
#pragma unroll 2
  for (int i = first; i < hi; i += kStep) {
    float v[V], s[V];
    ldv<V>(seg + i, s);
    if (kCache) {
      ldsv<V>(cache + (i - lo), v);
    } else {
      mixed.template noise<V>(i, v);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) mx = fmaxf(mx, fabsf(s[j] + gain_noise * (v[j] / ratio_noise)));
  }
  const float max_value = cluster_reduce<true>(cluster, block_reduce<true, kMixThreads>(mx, S.sh), S.tab_max);
  const bool declip = max_value > 1.0f;                  // data/waveform_mixers.py:50-53
  const float scale = declip ? 0.9f / max_value : 1.0f;

  // pass 3: the two outputs
#pragma unroll 2
  for (int i = first; i < hi; i += kStep) {
    float v[V], s[V], m[V];
    ldv<V>(seg + i, s);
    if (kCache) {
      ldsv<V>(cache + (i - lo), v);
    } else {
      mixed.template noise<V>(i, v);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      m[j] = s[j] + gain_noise * (v[j] / ratio_noise);
      if (declip) m[j] *= scale, s[j] *= scale;
    }
    stv<V>(out_m + i, m);
    stv<V>(out_s + i, s);
  }
}

// kCache: this CTA's slice of the summed noise is kept in (dynamic) shared memory after pass 1; otherwise (clips too long for
// that) passes 2 and 3 recompute it from the sources.
template <bool kCache>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kMixThreads)
    mixer_mix_kernel(const float* __restrict__ wave, int B, int L, int max_mix_num, const float* __restrict__ plan,
                     const float* __restrict__ partial, float* __restrict__ mixture, float* __restrict__ segment) {
  griddep_launch_dependents();
  griddep_wait();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float cache[];
  __shared__ MixShared S;
  const int n = blockIdx.y;
  const int rank = (int)cluster.block_rank();
  const float inv_len = 1.0f / (float)L;
  const float* seg = wave + (size_t)n * L;
  const float* row = plan + (size_t)n * (max_mix_num + 1);
  const float e_seg = clip_energy(partial, n, inv_len);
  const int terms = min(max((int)row[0] - 1, 0), kMaxTerms);
  if (threadIdx.x < terms) {
    const int k = threadIdx.x, src = (n + k + 1) % B;
    S.t.src[k] = wave + (size_t)src * L;
    S.t.ratio[k] = energy_ratio(clip_energy(partial, src, inv_len), e_seg);
    S.t.gain[k] = row[1 + k];
  }
  if (threadIdx.x == 0) S.t.n = terms;
  __syncthreads();
  const float gain_noise = row[max_mix_num];
  float* out_m = mixture + (size_t)n * L;
  float* out_s = segment + (size_t)n * L;
  const bool vec = (L % 4) == 0 && ((reinterpret_cast<uintptr_t>(wave) | reinterpret_cast<uintptr_t>(mixture) |
                                     reinterpret_cast<uintptr_t>(segment)) & 15) == 0;
  const int per = vec ? ((((L + kCluster - 1) / kCluster) + 3) & ~3) : (L + kCluster - 1) / kCluster;
  const int lo = min(L, rank * per), hi = min(L, lo + per);
#define LASS_MIX_PASSES(V, NT) mix_passes<V, kCache, NT>(cluster, S, cache, seg, out_m, out_s, lo, hi, e_seg, gain_noise, inv_len)
  if (vec) {
    if (terms <= 1) LASS_MIX_PASSES(4, 1);
    else if (terms <= 2) LASS_MIX_PASSES(4, 2);
    else if (terms <= 4) LASS_MIX_PASSES(4, 4);
    else LASS_MIX_PASSES(4, 0);
  } else {
    if (terms <= 2) LASS_MIX_PASSES(1, 2);
    else LASS_MIX_PASSES(1, 0);
  }
#undef LASS_MIX_PASSES
  cluster.sync();     // no CTA exits while a neighbour may still write into its tables
}

constexpr size_t kMaxCacheBytes = 200 * 1024;

}  // namespace

}  // namespace lass

using namespace lass;

extern "C" {

size_t lass_segment_mix_scratch_bytes(int B) { return B > 0 ? (size_t)B * kSlices * sizeof(float) : 0; }

int lass_segment_mix(const float* wave, int B, int L, int max_mix_num, const float* plan, float* mixture, float* segment,
                     void* scratch, size_t scratch_bytes, void* stream_v) {
  if (!wave || !plan || !mixture || !segment || !scratch || B <= 0 || L <= 0 || max_mix_num < 2 || max_mix_num - 1 > kMaxTerms)
    return set_error(LASS_ERR_ARG, "lass_segment_mix: bad argument (B %d, L %d, max_mix_num %d: 2..%d)", B, L, max_mix_num, kMaxTerms + 1);
  if (scratch_bytes < lass_segment_mix_scratch_bytes(B)) return set_error(LASS_ERR_ARG, "lass_segment_mix: scratch too small");
  if (B > 65535) return set_error(LASS_ERR_ARG, "lass_segment_mix: at most 65535 clips per call");
  if (mixture == wave || segment == wave || mixture == segment) return set_error(LASS_ERR_ARG, "lass_segment_mix: outputs must not alias");
  cudaStream_t s = (cudaStream_t)stream_v;
  float* partial = (float*)scratch;
  cudaError_t e = launch_pdl(mixer_energy_kernel, dim3(kSlices, (unsigned)B), kThreads, 0, s, wave, L, partial);
  if (e != cudaSuccess) return set_cuda_error(e, "mixer_energy launch");
  const size_t cache_bytes = (size_t)((((L + kCluster - 1) / kCluster) + 3) & ~3) * sizeof(float);
  if (cache_bytes <= kMaxCacheBytes) {
    // cudaFuncSetAttribute is per device and cheap: set on every call rather than cached in process-wide state
    if (cache_bytes > 48 * 1024) {
      e = cudaFuncSetAttribute(mixer_mix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxCacheBytes);
      if (e != cudaSuccess) return set_cuda_error(e, "mixer_mix shared-memory opt-in");
    }
    e = launch_pdl(mixer_mix_kernel<true>, dim3(kCluster, (unsigned)B), kMixThreads, cache_bytes, s, wave, B, L, max_mix_num, plan,
                   (const float*)partial, mixture, segment);
  } else {
    e = launch_pdl(mixer_mix_kernel<false>, dim3(kCluster, (unsigned)B), kMixThreads, 0, s, wave, B, L, max_mix_num, plan,
                   (const float*)partial, mixture, segment);
  }
  if (e != cudaSuccess) return set_cuda_error(e, "mixer_mix launch");
  return LASS_OK;
}

}  // extern "C"
