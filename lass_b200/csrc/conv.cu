// K3 / K4: implicit-GEMM convolution on tcgen05 (sm_100a) for the FiLM-conditioned ResUNet30.
//
// Replaces (reference, per forward): every conv2d 3x3 / 1x1 and conv_transpose2d (kernel = stride) of
// `ConvBlockRes` / `EncoderBlockRes1B` / `DecoderBlockRes1B` (models/resunet.py:147-165,186-198,240-264)
// together with the ops around them: batch_norm + FiLM add + leaky_relu (as the PRODUCER's epilogue),
// the residual / 1x1 shortcut add (as one more K-segment of the same accumulator), avg_pool2d, torch.cat
// (outputs are written straight into the concat buffer's channel slice) and after_conv (:570).
//
// GEMM view:  D[pixel, cout] = sum over segments, taps (dy,dx), channels  A[pixel + (dy,dx), c] * W[tap][cout][c]
//   * activations are NHWC 16-bit, already activated by their producer, so the conv's zero padding is the
//     TMA out-of-bounds zero fill;
//   * A: ONE 4-D TMA box per K-chunk brings the (16*MT + 2) x 10 pixel halo tile of kc channels into shared
//     memory (rows = pixels, kc*2 bytes each, hardware swizzle).  The nine taps are nine shared-memory matrix
//     descriptors into that same tile: start address shifted by (dy*10 + dx) rows, 8-row groups (8 pixels of
//     one image row) strided by the halo pitch.  tcgen05 applies the swizzle on absolute smem address bits,
//     so row-shifted starts are legal (verified on B200 by tests/test_gpu_umma_probe.py);
//   * B: weight tiles (BN couts x kc channels) per (tap, chunk), streamed through a ring — or, when all tiles of
//     a work item fit in the ring, loaded once and kept resident for the CTA's lifetime;
//   * D: fp32 accumulators in TMEM, a ring of 4 (or 2) stages so the epilogue of item i overlaps the MMAs of items i+1, i+2.
// Persistent CTAs (grid = SM count x CTAs/SM), static round-robin over work items
// (item = pixel tile (16*MT x 8) x N tile).  Warp roles (384 threads): 0-1 = TMA producers, 2-3 = MMA issuers (+ TMEM
// alloc), 4-7 / 8-11 = two epilogue groups (TMEM -> registers -> fp32 math -> 16-bit NHWC stores / TMA stores).
// CG2 instantiations run as CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256): see the comment at `cg2` in the kernel.
#include <string.h>

#include <new>

#include "conv.cuh"
#include "conv_issue.cuh"
#include "ptx.cuh"

namespace lass {

namespace {

template <uint32_t V>
struct UTag {
  static constexpr uint32_t value = V;
};
// epilogue features of conv_igemm_kernel (compile-time sets for the launch shapes of the ResUNet30 plan)
enum : uint32_t {
  kFRaw = 1u, kFAct = 2u, kFTma = 4u, kFTmaPool = 8u, kFPool = 16u, kFPoolH2 = 32u, kFPoolRaw = 64u, kFPoolAct = 128u,
  kFAfter = 256u, kFResid = 512u, kFBias = 1024u, kFUp = 2048u, kFCoal = 4096u, kFGeneric = 0x80000000u,
  // encoder conv2: raw + activated skip into the concat buffers (TMA stores) + pooled raw / activated block output
  kFEnc2Common = kFRaw | kFAct | kFTma | kFPool | kFPoolRaw | kFPoolAct,
  kFEnc2Resid = kFEnc2Common | kFPoolH2 | kFResid,       // encoder_block1 (rank-1 identity residual)
  kFEnc2Bias = kFEnc2Common | kFPoolH2 | kFBias,         // encoder_block2..5 (1x1 shortcut segment, bias)
  kFEnc2BiasP12 = kFEnc2Common | kFBias,                 // encoder_block6 (pooling (1, 2))
  kFUpconv = kFRaw | kFAct | kFTma | kFUp,               // transposed convs into the concat buffers
  kFUpDirect = kFRaw | kFAct | kFUp,                     // the last one (32 channels): 32-byte direct stores
  kFAfterBias = kFAfter | kFBias,                        // decoder_block6 conv2 + after_conv
};
constexpr int kThreadsK = 384;  // conv_igemm_kernel: warps 0-1 TMA producers, 2-3 MMA issuers, 4-7 / 8-11 epilogue groups
// accumulator stages / TMEM columns of conv_igemm_kernel<BN, MT>
constexpr int acc_stages(int BN, int MT) { return (4 * MT * BN <= 256) ? 4 : (2 * MT * BN <= 512) ? 2 : 1; }
constexpr int tmem_cols(int BN, int MT) {
  const int c = acc_stages(BN, MT) * MT * BN;
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}
constexpr int kMaxA = 8;
constexpr int kMaxB = 40;
constexpr float kSlope = 0.01f;
constexpr int kStageWarpBytes = 6 * 1024;   // per epilogue warp: full_raw 2K | full_act 2K | pool_raw 1K | pool_act 1K
constexpr int kEpiWarps = 8;

struct SegDev {
  CUtensorMap tmA;  // (C, W, H, B) view of the source channels
  CUtensorMap tmB;  // (cin, ncols, taps)
  int nchunks, taps, kc, fmt;
};

struct OutDev {
  void* ptr;
  const float* scale;
  const float* shift;
  int cstride, coff, fp16, shift_bstride;
  int st256;   // every 32-channel piece of this output is 32 B aligned: direct stores use 32-byte st.global
};

struct ConvParams {
  SegDev seg[2];
  // TMA-store tensor maps: [0,1] full_raw (dy = 0,1), [2,3] full_act (dy = 0,1), [4] pool_raw, [5] pool_act
  CUtensorMap tm_out[6];
  OutDev full_raw, full_act, pool_raw, pool_act;
  const float* bias;
  const float* after_w;
  const float* after_b;
  float* feat;
  const float* resid_src;       // rank-1 residual from a 1-channel fp32 map (see lass_conv_desc)
  const float* resid_in_scale;
  const float* resid_in_shift;
  const float* resid_w;
  const float* resid_b;
  int resid_T, resid_F;
  // generated A operand of segment 0 (lass_conv_desc.gen_src)
  const float* gen_src;
  const float* gen_in_scale;
  const float* gen_in_shift;
  const float* gen_w;
  const float* gen_b;
  const float* gen_scale;
  const float* gen_shift;
  int gen_shift_bstride, gen_T, gen_F;
  long long* prof;  // optional per-CTA cycle counters (kProfSlots each), nullptr = off
  int nseg;
  int B, H, W, ncols;
  int tiles_h, tiles_w, pix_tiles, n_tiles, num_items;
  int a_stages, b_stages, b_resident;
  int b_tps;       // streamed weights: taps per weight stage of a 3x3 segment (1, or 3 = the three taps of a kernel row in one
                   // TMA box and one full / empty handshake)
  int epi_mode;    // conv_igemm_kernel: 0 = generic epilogue (run-time feature tests), 1 = lean path (one activated bf16 output,
                   // direct stores), 2..6 = generic code specialised at compile time for a feature set (see kFEnc2Resid ...)
  int cg2;         // conv_igemm_kernel: CTA pairs (cluster of 2, tcgen05 cta_group::2) -- streamed weights, N >= 128
  int dual_issue;  // conv_igemm_kernel: two MMA-issuing warps, each with half of the A ring (resident weights, >= 4 A stages)
  int tma_store;   // 1: full-resolution 16-bit outputs leave through per-warp shared-memory staging + TMA stores
  int tma_pool;    // 1: pooled outputs too (only when they are channel slices; whole-pixel pooled outputs store directly)
  uint32_t a_stage_bytes, b_stage_bytes;
  int up_h, up_w, group_c;
  int pool_h, pool_w;
  int debug_flags;  // timing experiments only: 1 epilogue idle, 2 no MMA, 4 no A loads, 8 no pooled outputs, 16 no stores,
                    // 32 (host side) direct global stores instead of TMA stores
};

struct Item {
  int b, h0, w0, n0;
};

template <int MT>
__device__ __forceinline__ Item decode_item(const ConvParams& p, int item, int BN) {
  Item it;
  const int nt = item / p.pix_tiles;
  int pix = item - nt * p.pix_tiles;
  const int per_img = p.tiles_h * p.tiles_w;
  it.b = pix / per_img;
  pix -= it.b * per_img;
  const int th = pix / p.tiles_w;
  it.h0 = th * (16 * MT);
  it.w0 = (pix - th * p.tiles_w) * TW;
  it.n0 = nt * BN;
  return it;
}

// Epilogue parameters of the current (clip, N tile), staged in shared memory once per change so that the
// per-column-chunk code reads them with broadcast LDS instead of dependent global loads.
template <int BN>
struct EpiTables {
  float bias[BN];
  float sc_full[BN], sh_full[BN];
  float sc_pool[BN], sh_pool[BN];
  float resid_w[BN];
  float after_w[3 * 32];
  float after_b[4];
};

__device__ __forceinline__ void store32(const OutDev& o, int b, int ho, int wo, int Ho, int Wo, int c, const uint32_t* w,
                                        bool valid) {
  if (valid) {
    uint16_t* base = reinterpret_cast<uint16_t*>(o.ptr) + (((size_t)b * Ho + ho) * Wo + wo) * o.cstride + o.coff + c;
    if (o.st256) {
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(base), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                   "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(base + 16), "r"(w[8]), "r"(w[9]), "r"(w[10]),
                   "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
                   : "memory");
      return;
    }
    uint4* dst = reinterpret_cast<uint4*>(base);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
    dst[3] = make_uint4(w[12], w[13], w[14], w[15]);
  }
}

// raw output: the value itself as saturating fp16 (residual / skip stream), or as bf16 (backward-data launches of the
// training step: gradients span too many decades for fp16)
__device__ __forceinline__ void pack_raw32(const float* v, uint32_t* w, bool f16 = true) {
  if (f16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_f16x2_sat(v[2 * j], v[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  }
}

// activated output: lrelu(sc * v + sh) as bf16 (operand of the next convolution)
__device__ __forceinline__ void pack_act32(const float* sc, const float* sh, const float* v, uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 a = *reinterpret_cast<const float4*>(sc + j);
    const float4 s = *reinterpret_cast<const float4*>(sh + j);
    const float t0 = fmaf(a.x, v[j + 0], s.x), t1 = fmaf(a.y, v[j + 1], s.y);
    const float t2 = fmaf(a.z, v[j + 2], s.z), t3 = fmaf(a.w, v[j + 3], s.w);
    w[j / 2] = lrelu_bf16x2(pack_bf16x2(t0, t1));          // fp32 affine, LeakyReLU on the packed bf16 pair
    w[j / 2 + 1] = lrelu_bf16x2(pack_bf16x2(t2, t3));
  }
}

// Staging tile of one warp: rows of 64 B (32 channels), CU_TENSOR_MAP_SWIZZLE_64B pattern (16 B chunk index XOR
// address bits [7,8]) so that 32 lanes writing one row each are bank-conflict free.
__device__ __forceinline__ uint4* stage_slot(unsigned char* tile, int row, int piece) {
  return reinterpret_cast<uint4*>(tile + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
}
// explicit st.shared (a store through a generic pointer into the shared window is several times slower)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t stage_slot_addr(uint32_t tile_addr, int row, int piece) {
  return tile_addr + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void stage_row32(unsigned char* tile, int row, const uint32_t* w) {
  const uint32_t t = smem_u32(tile);
#pragma unroll
  for (int pc = 0; pc < 4; ++pc) sts128(stage_slot_addr(t, row, pc), w[4 * pc], w[4 * pc + 1], w[4 * pc + 2], w[4 * pc + 3]);
}

// profiling slots (clock cycles, per CTA); documented at lass_debug_set_conv_profile in include/lass_b200.h
enum { kProfProdAEmpty = 0, kProfProdBEmpty, kProfProdTotal, kProfMmaAccEmpty, kProfMmaAFull, kProfMmaBFull, kProfMmaTotal,
       kProfEpiAccFull, kProfEpiTotal, kProfItems, kProfSlots = 16 };

// The role profiler is compiled in only with -DLASS_CONV_PROFILE (make prof -> liblass_b200_prof.so): even switched off
// at run time its counters cost registers and a few per cent in the epilogue-bound layers.
#ifdef LASS_CONV_PROFILE
#define LASS_PROF_ON(p) ((p).prof != nullptr)
#else
#define LASS_PROF_ON(p) false
#endif

#define LASS_TIMED_WAIT(bar, parity, slot)              \
  do {                                                  \
    const long long _t = prof ? clock64() : 0;          \
    mbar_wait(bar, parity);                             \
    if (prof) pc[slot] += clock64() - _t;               \
  } while (0)
#define LASS_TIMED_WAIT_RELAXED(bar, parity, slot)      \
  do {                                                  \
    const long long _t = prof ? clock64() : 0;          \
    mbar_wait_relaxed(bar, parity);                     \
    if (prof) pc[slot] += clock64() - _t;               \
  } while (0)

// 16-byte shared-memory load by 32-bit shared address.  Volatile: stays where it is written relative to the named
// barriers that publish the epilogue tables and to the TMEM loads it is meant to overlap with.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

// 32-byte global store (sm_100 STG.256): a lane's piece covers whole 32 B sectors.  The per-lane pieces of these epilogues lie
// one pixel (>= 64 B) apart, so a 16 B store writes HALF a sector per lane and costs twice the LSU transactions.
__device__ __forceinline__ void stg256(void* dst, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// y = lrelu(sc * v + sh) for 32 channels of one pixel, packed to bf16 and stored as 2 x 32 B (st256) or 4 x 16 B.
__device__ __forceinline__ void act_store32(const float* v, const float4* sc, const float4* sh, uint16_t* dst, bool valid,
                                            bool st256) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t0 = fmaf(sc[j].x, v[4 * j + 0], sh[j].x), t1 = fmaf(sc[j].y, v[4 * j + 1], sh[j].y);
    const float t2 = fmaf(sc[j].z, v[4 * j + 2], sh[j].z), t3 = fmaf(sc[j].w, v[4 * j + 3], sh[j].w);
    w[2 * j] = lrelu_bf16x2(pack_bf16x2(t0, t1));          // fp32 affine, LeakyReLU on the packed bf16 pair
    w[2 * j + 1] = lrelu_bf16x2(pack_bf16x2(t2, t3));
  }
  if (valid) {
    if (st256) {
      stg256(dst, w);
      stg256(dst + 16, w + 8);
    } else {
      uint4* d = reinterpret_cast<uint4*>(dst);
      d[0] = make_uint4(w[0], w[1], w[2], w[3]);
      d[1] = make_uint4(w[4], w[5], w[6], w[7]);
      d[2] = make_uint4(w[8], w[9], w[10], w[11]);
      d[3] = make_uint4(w[12], w[13], w[14], w[15]);
    }
  }
}

template <int BN, int MT, bool CG2 = false>
__global__ void __launch_bounds__(kThreadsK, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  // accumulator ring: item n of this CTA uses stage n % NS; MMA warp (n & 1) issues it, epilogue group (n & 1) drains it.
  // Four stages where they fit in 256 columns, so that the MMAs of item n + 2 do not have to wait for the epilogue of item n.
  constexpr int NS = acc_stages(BN, MT);
  constexpr int kTmemCols = tmem_cols(BN, MT);
  // With a single accumulator stage (N = 256, two m-tiles) both epilogue groups work on EVERY item, one m-tile each:
  // groups that alternate items would have to poll a phase of the one acc_full barrier whose predecessor is still open.
  constexpr bool kSplitMT = (NS == 1);
  static_assert(!kSplitMT || MT == 2, "single-stage accumulators are split between the two epilogue groups by m-tile");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* a_buf = smem;
  unsigned char* b_buf = smem + (size_t)p.a_stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.b_stages * p.b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxA;
  uint64_t* b_full = a_empty + kMaxA;
  uint64_t* b_empty = b_full + kMaxB;
  uint64_t* acc_full = b_empty + kMaxB;
  uint64_t* acc_empty = acc_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  EpiTables<BN>* tabs = reinterpret_cast<EpiTables<BN>*>(reinterpret_cast<unsigned char*>(tmem_slot) + 16);  // [group][2]
  unsigned char* stage_base = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(tabs + 4) + 1023) & ~uintptr_t(1023));                                     // [8 warps][6 KiB]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA-pair mode (p.cg2, thread-block cluster of 2, tcgen05 cta_group::2): the two CTAs take two pixel tiles of the SAME
  // N tile; every MMA covers both (M = 256) and each CTA holds only half of the weight rows, so the operand reads per CTA
  // drop from 4 KiB + 32 N to 4 KiB + 16 N bytes per MMA (shared-memory bandwidth is what bounds the N >= 128 layers).  The
  // leader CTA (rank 0) issues the MMAs and owns the full / acc_empty barriers; both CTAs load, and drain their own TMEM.
  constexpr bool cg2 = CG2;        // compile time: the single-CTA instantiations carry none of the pair code
  const uint32_t crank = cg2 ? cluster_ctarank() : 0u;
  // the item loops below run over VIRTUAL indices: a pair shares one index and decodes it to two neighbouring pixel tiles
  const int vbid = cg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int vgrid = cg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int vitems = cg2 ? p.num_items >> 1 : p.num_items;
  auto dec = [&](int v) {
    int item = v;
    if (cg2) {
      const int half = p.pix_tiles >> 1;
      const int nt = v / half;
      item = nt * p.pix_tiles + 2 * (v - nt * half) + (int)crank;
    }
    return decode_item<MT>(p, item, BN);
  };
  const bool prof = LASS_PROF_ON(p);
  long long pc[kProfSlots];
#pragma unroll
  for (int i = 0; i < kProfSlots; ++i) pc[i] = 0;
  const long long t_start = prof ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      if (!(s == 0 && p.gen_src != nullptr)) tma_prefetch_desc(&p.seg[s].tmA);
      tma_prefetch_desc(&p.seg[s].tmB);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], (kSplitMT ? 8 : 4) * (cg2 ? 2 : 1));     // pair: the epilogue warps of both CTAs arrive
    }
    fence_mbar_init();
  }
  if (cg2) cluster_sync_all();      // both CTAs' barriers exist before any remote arrive / transaction lands on them
  if (warp == 2) {
    if (cg2) {
      tmem_alloc_cg2(tmem_slot, kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (conv_run): everything above -- barriers, TMEM, tensor-map prefetch -- touches nothing a
  // previous launch writes and may overlap its tail; the NEXT conv launch may now be scheduled onto free SMs and do the same.
  // Nothing below runs before every earlier launch of the stream has completed.
  griddep_launch_dependents();
  griddep_wait();

  const bool mma_only = (p.debug_flags & 64) != 0;   // timing experiment: the MMA issuer runs free, nothing else runs
  if (warp < 2) {
    // =========================== TMA producers ===========================
    // Two producer warps, one per MMA issuer: producer w loads this CTA's items w, w + 2, ... into ITS half of the A ring
    // (and of the weight ring when the weights are streamed), so the two item streams advance independently.  Without
    // p.dual_issue producer 0 owns the whole rings and producer 1 idles.  Warp-uniform loops; the elected lane issues.
    const bool dual = p.dual_issue != 0;
    const uint32_t pw = (uint32_t)warp;
    if (!mma_only && (dual || pw == 0)) {
      const uint32_t ring_a = dual ? (uint32_t)p.a_stages >> 1 : (uint32_t)p.a_stages;
      const uint32_t ring_b = dual ? (uint32_t)p.b_stages >> 1 : (uint32_t)p.b_stages;
      const uint32_t a0 = dual ? pw * ring_a : 0u, b0 = dual ? pw * ring_b : 0u;
      const int step = dual ? 2 : 1;
      uint32_t a_it = 0, b_it = 0;
      bool first_item = (pw == 0);     // resident weights are loaded once, by producer 0
      uint32_t gen_ca[16], gen_cb[16]; // generated-A mode: per-clip affine of the 32 channels, packed bf16 pairs
      int gen_b = -1;
      constexpr int kGenRows = (16 * MT + 2) * kHaloPitch;          // 64 B rows (32 channels) of the halo tile
      constexpr int kGenPer = (kGenRows + 31) / 32;
      // raw loads of this lane's pixels of the item about to be generated; they are only CONSUMED (bn0 affine) one item
      // later, so the in-order warp never waits for them
      float gen_m[kGenPer], gen_s[kGenPer], gen_t[kGenPer];
      bool gen_have = false;
      // per-lane constants of the kGenPer pixels this lane generates: (row, column) inside the halo tile and the element
      // offset of the magnitude relative to the tile's first pixel -- the per-item address work is then one add per load
      int gen_hh[kGenPer], gen_ww[kGenPer], gen_off[kGenPer];
#pragma unroll
      for (int i = 0; i < kGenPer; ++i) {
        const int r = lane + 32 * i;
        gen_hh[i] = (r * 205) >> 11;                           // r / 10 for r < 1029
        gen_ww[i] = r - gen_hh[i] * kHaloPitch;
        gen_off[i] = gen_hh[i] * p.gen_F + gen_ww[i];
      }
      auto gen_load = [&](const Item& gi) {
        const int h1 = gi.h0 - 1, w1 = gi.w0 - 1;              // first pixel of the halo tile
        const float* base = p.gen_src + ((size_t)gi.b * p.gen_T + h1) * p.gen_F + w1;
        const float* sc = p.gen_in_scale + w1;
        const float* sh = p.gen_in_shift + w1;
        // interior tile (every halo pixel inside the grid and above the time padding): no per-pixel tests
        const bool interior = h1 >= 0 && w1 >= 0 && w1 + kHaloPitch <= p.W && h1 + 16 * MT + 2 <= min(p.H, p.gen_T);
        if (interior) {
#pragma unroll
          for (int i = 0; i < kGenPer; ++i) {
            const bool on = lane + 32 * i < kGenRows;
            gen_m[i] = on ? __ldg(base + gen_off[i]) : 0.0f;
            gen_s[i] = on ? __ldg(sc + gen_ww[i]) : 0.0f;
            gen_t[i] = on ? __ldg(sh + gen_ww[i]) : 0.0f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < kGenPer; ++i) {
            const int h = h1 + gen_hh[i], w = w1 + gen_ww[i];
            const bool inside = lane + 32 * i < kGenRows && h >= 0 && h < p.H && w >= 0 && w < p.W;
            const bool live = inside && h < p.gen_T;                    // zero time padding AFTER bn0 (models/resunet.py:548)
            gen_m[i] = live ? __ldg(base + gen_off[i]) : 0.0f;
            gen_s[i] = live ? __ldg(sc + gen_ww[i]) : 0.0f;
            gen_t[i] = inside ? (live ? __ldg(sh + gen_ww[i]) : 0.0f) : __int_as_float(0x7fc00000);   // NaN = outside the grid
          }
        }
      };
      for (int item = vbid + (dual ? (int)pw : 0) * vgrid; item < vitems; item += step * vgrid) {
        const Item it = dec(item);
        uint32_t b_slot_res = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const SegDev& sg = p.seg[s];
          const uint32_t row_bytes = sg.kc * 2;
          const bool halo = sg.taps == 9;
          const uint32_t a_bytes = halo ? (uint32_t)(16 * MT + 2) * kHaloPitch * row_bytes : (uint32_t)(16 * MT) * TW * row_bytes;
          const uint32_t b_bytes = BN * row_bytes;
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            const uint32_t sa = a0 + a_it % ring_a;
            LASS_TIMED_WAIT_RELAXED(&a_empty[sa], ((a_it / ring_a) & 1) ^ 1, kProfProdAEmpty);
            if (s == 0 && p.gen_src != nullptr) {
              // ---- generated A tile: activated pre_conv output of the 1-channel magnitude, written by this warp ----
              if (it.b != gen_b) {
                gen_b = it.b;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                  float ca[2], cb[2];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const float as = __ldg(p.gen_scale + c + u);
                    ca[u] = as * __ldg(p.gen_w + c + u);
                    cb[u] = fmaf(as, __ldg(p.gen_b + c + u), __ldg(p.gen_shift + (size_t)it.b * p.gen_shift_bstride + c + u));
                  }
                  gen_ca[c / 2] = pack_bf16x2(ca[0], ca[1]);
                  gen_cb[c / 2] = pack_bf16x2(cb[0], cb[1]);
                }
              }
              const uint32_t tile = smem_u32(a_buf + (size_t)sa * p.a_stage_bytes);
              if (!gen_have) gen_load(it);                                 // first item: nothing was prefetched
              float xs[kGenPer];
#pragma unroll
              for (int i = 0; i < kGenPer; ++i) xs[i] = fmaf(gen_s[i], gen_m[i], gen_t[i]);
              {
                // prefetch the magnitudes of this producer's NEXT item: their latency hides behind the tile written below
                const int nxt = item + step * vgrid;
                gen_have = nxt < vitems;
                if (gen_have) gen_load(dec(nxt));
              }
#pragma unroll
              for (int i = 0; i < kGenPer; ++i) {
                const int r = lane + 32 * i;
                if (r < kGenRows) {
                  const float x = xs[i];
                  if (x == x) {
                    // Packed bf16 arithmetic (3 instructions per channel pair; the fp32 form needs 5 and the two producer
                    // warps cannot keep up with the tensor pipe): x, scale and shift are rounded to bf16 before the fused
                    // multiply-add, i.e. this operand carries about three bf16 roundings instead of one.
                    const uint32_t xx = pack_bf16x2(x, x);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                      uint32_t wv[4];
#pragma unroll
                      for (int j = 0; j < 4; ++j) wv[j] = affine_lrelu_bf16x2(gen_ca[4 * pc + j], xx, gen_cb[4 * pc + j]);
                      sts128(stage_slot_addr(tile, r, pc), wv[0], wv[1], wv[2], wv[3]);
                    }
                  } else {                       // outside the grid: the convolution's zero padding
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) sts128(stage_slot_addr(tile, r, pc), 0u, 0u, 0u, 0u);
                  }
                }
              }
              fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core's operand reads
              __syncwarp();
              if (elect_one()) mbar_arrive(&a_full[sa]);
              __syncwarp();
            } else if (elect_one()) {
              if (p.debug_flags & 4) {
                mbar_arrive(&a_full[sa]);
              } else if (cg2) {
                // both CTAs load their own tile; the bytes of both are counted on the LEADER's barrier
                if (crank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * a_bytes);
                tma_load_4d_cg2(a_buf + (size_t)sa * p.a_stage_bytes, &sg.tmA, mapa_shared(smem_u32(&a_full[sa]), 0), ch * sg.kc,
                                halo ? it.w0 - 1 : it.w0, halo ? it.h0 - 1 : it.h0, it.b);
              } else {
                mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                tma_load_4d(a_buf + (size_t)sa * p.a_stage_bytes, &sg.tmA, &a_full[sa], ch * sg.kc,
                            halo ? it.w0 - 1 : it.w0, halo ? it.h0 - 1 : it.h0, it.b);
              }
            }
            __syncwarp();
            ++a_it;
            if (p.b_resident) {
              if (first_item) {
                for (int tp = 0; tp < sg.taps; ++tp, ++b_slot_res) {
                  if (elect_one()) {
                    if (cg2) {
                      if (crank == 0) mbar_arrive_expect_tx(&b_full[b_slot_res], b_bytes);
                      tma_load_3d_cg2(b_buf + (size_t)b_slot_res * p.b_stage_bytes, &sg.tmB,
                                      mapa_shared(smem_u32(&b_full[b_slot_res]), 0), ch * sg.kc, it.n0 + (int)crank * (BN / 2), tp);
                    } else {
                      mbar_arrive_expect_tx(&b_full[b_slot_res], b_bytes);
                      tma_load_3d(b_buf + (size_t)b_slot_res * p.b_stage_bytes, &sg.tmB, &b_full[b_slot_res], ch * sg.kc,
                                  it.n0, tp);
                    }
                  }
                  __syncwarp();
                }
              }
            } else {
              const int tps = halo ? p.b_tps : 1;         // taps per stage: the box of tmB covers tps taps
              for (int tp = 0; tp < sg.taps; tp += tps) {
                const uint32_t sb = b0 + b_it % ring_b;
                LASS_TIMED_WAIT_RELAXED(&b_empty[sb], ((b_it / ring_b) & 1) ^ 1, kProfProdBEmpty);
                if (elect_one()) {
                  if (cg2) {
                    // each CTA holds half of the tile's weight rows: rows [n0 + rank * BN/2, + BN/2)
                    if (crank == 0) mbar_arrive_expect_tx(&b_full[sb], (uint32_t)tps * b_bytes);
                    tma_load_3d_cg2(b_buf + (size_t)sb * p.b_stage_bytes, &sg.tmB, mapa_shared(smem_u32(&b_full[sb]), 0),
                                    ch * sg.kc, it.n0 + (int)crank * (BN / 2), tp);
                  } else {
                    mbar_arrive_expect_tx(&b_full[sb], (uint32_t)tps * b_bytes);
                    tma_load_3d(b_buf + (size_t)sb * p.b_stage_bytes, &sg.tmB, &b_full[sb], ch * sg.kc, it.n0, tp);
                  }
                }
                __syncwarp();
                ++b_it;
              }
            }
          }
        }
        first_item = false;
      }
      if (prof && lane == 0 && pw == 0) {
        long long* dst = p.prof + (size_t)blockIdx.x * kProfSlots;
        dst[kProfProdAEmpty] = pc[kProfProdAEmpty];
        dst[kProfProdBEmpty] = pc[kProfProdBEmpty];
        dst[kProfProdTotal] = clock64() - t_start;
      }
    }
  } else if (warp < 4) {
    // =========================== MMA issuers ===========================
    // With p.dual_issue there are TWO issuing warps: warp 2 takes this CTA's items 0, 2, 4, ..., warp 3 the items
    // 1, 3, 5, ...  Measured on B200 (tools/gpu_umma_bench3.py, lass_debug_set_conv_flags 64): the tensor pipe does not
    // run ahead of the issuing thread, so everything a single issuer does between two items (commits, barrier polls, loop
    // bookkeeping, descriptor set-up: 500-700 cycles) shows up as idle tensor pipe -- a third of the time when an item is
    // 36 MMAs of 40 cycles.  With two issuers that work hides behind the other warp's MMAs.  The items of the two warps
    // accumulate into different TMEM stages, so their relative order in the pipe does not matter.  Each issuer owns HALF of
    // the A ring and of the streamed-weight ring, filled by its own producer warp: a shared ring would make a warp poll an
    // mbarrier phase whose predecessor has not completed yet, which a parity wait cannot express.
    // The whole warp runs the (warp-uniform) loop so that descriptors live in uniform registers; only the elected
    // lane issues tcgen05.mma / tcgen05.commit.  Steady state with resident weights: one wait + one elected region
    // per K-chunk that issues all taps back to back (compile-time tap offsets); streaming weights: per-tap
    // full/empty handshake on the B ring.
    {
      constexpr uint32_t kLbo = 1u << 16;
      const uint32_t a_base16 = smem_u32(a_buf) >> 4, a_stage16 = p.a_stage_bytes >> 4;
      const uint32_t b_base16 = smem_u32(b_buf) >> 4, b_stage16 = p.b_stage_bytes >> 4;
      const bool dual = p.dual_issue != 0;
      const uint32_t mw = (uint32_t)warp - 2u;     // which of the two issuers
      const uint32_t n_a = dual ? (uint32_t)p.a_stages >> 1 : (uint32_t)p.a_stages;
      const uint32_t n_b = (dual && p.b_resident == 0) ? (uint32_t)p.b_stages >> 1 : (uint32_t)p.b_stages;
      const uint32_t ring0 = dual ? mw * n_a : 0u;   // first stage of this issuer's part of the A ring
      const uint32_t bring0 = (dual && p.b_resident == 0) ? mw * n_b : 0u;   // ... and of the streamed-weight ring
      const bool resident = p.b_resident != 0;
      const bool no_mma = (p.debug_flags & 2) != 0;
      // per-segment constants, hoisted out of the item loop
      SegMma g[2];
      uint32_t seg_kc[2], seg_chunks[2], seg_halo[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const bool on = s < p.nseg;
        const uint32_t kc = on ? p.seg[s].kc : 64;
        const uint32_t halo = (on && p.seg[s].taps == 9) ? 1u : 0u;
        const uint32_t swz = kc == 64 ? kSwizzle128B : kSwizzle64B;
        g[s].a_hi = static_cast<uint32_t>(make_smem_desc(0, (halo ? kHaloPitch : TW) * kc * 2, swz) >> 32);
        g[s].b_hi = static_cast<uint32_t>(make_smem_desc(0, 8 * kc * 2, swz) >> 32);
        g[s].idesc = make_idesc_f16(on ? p.seg[s].fmt : 0, on ? p.seg[s].fmt : 0, cg2 ? 256 : 128, BN);
        seg_kc[s] = kc;
        seg_chunks[s] = on ? p.seg[s].nchunks : 0;
        seg_halo[s] = halo;
      }
      uint32_t sa = 0, pa = 0;          // A ring (this issuer's part): stage, phase
      uint32_t sb = 0, pb = 0;          // B ring (streaming mode) / running slot (resident mode)
      uint32_t n_items = 0;
      bool first_item = true;
      const int step = dual ? 2 : 1;
      uint32_t n = dual ? mw : 0u;      // index of the item within this CTA's sequence
      for (int item = ((dual || mw == 0) && crank == 0) ? vbid + (int)n * vgrid : vitems; item < vitems;
           item += step * vgrid, n += step) {
        const uint32_t as = n % NS, pacc = (n / NS) & 1u;
        if (!mma_only) LASS_TIMED_WAIT(&acc_empty[as], pacc ^ 1, kProfMmaAccEmpty);
        tc_fence_after_sync();
        const uint32_t acc_addr = tmem_base + as * (MT * BN);
        uint32_t accumulate = 0;
        if (resident) sb = 0;
        const bool need_wait = !resident || (first_item && !mma_only);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const SegMma gs = g[s];
          const uint32_t kc = seg_kc[s];
          const bool halo = seg_halo[s] != 0;
          const uint32_t row16 = kc >> 3;                       // row bytes / 16
          const uint32_t pitch = halo ? kHaloPitch : TW;
          const uint32_t mt_step16 = 16 * pitch * row16;
          const bool row_stage = !resident && halo && p.b_tps == 3;
          const uint32_t b_tile16 = (cg2 ? BN / 2 : BN) * row16;       // one tap's weight tile inside a row stage
#pragma unroll 1
          for (uint32_t ch = 0; ch < seg_chunks[s]; ++ch) {
            if (!mma_only) LASS_TIMED_WAIT(&a_full[ring0 + sa], pa, kProfMmaAFull);
            tc_fence_after_sync();
            uint32_t a_lo = (a_base16 + (ring0 + sa) * a_stage16) | kLbo;
            if (!need_wait) {
              // ---- resident weights, steady state: all taps of the chunk in one elected region ----
              const uint32_t b_lo = (b_base16 + sb * b_stage16) | kLbo;
              if (!no_mma && elect_one()) {
                if (halo) {
                  if (kc == 64) issue_halo_chunk_running<MT, BN, 4, cg2>(acc_addr, a_lo, b_lo, b_stage16, gs, accumulate);
                  else issue_halo_chunk_running<MT, BN, 2, cg2>(acc_addr, a_lo, b_lo, b_stage16, gs, accumulate);
                } else {
                  if (kc == 64) issue_tap<MT, BN, 4, cg2>(acc_addr, a_lo, 16 * TW * 8, b_lo, gs, accumulate);
                  else issue_tap<MT, BN, 2, cg2>(acc_addr, a_lo, 16 * TW * 4, b_lo, gs, accumulate);
                }
              }
              __syncwarp();
              accumulate = 1;
              sb += halo ? 9u : 1u;
            } else {
              const uint32_t tap_rows = halo ? 3u : 1u;
#pragma unroll 1
              for (uint32_t dy = 0; dy < tap_rows; ++dy) {
#pragma unroll
                for (uint32_t dx = 0; dx < 3; ++dx) {
                  if (dx > 0 && !halo) break;
                  // row_stage: the three taps of this kernel row share one weight stage (one wait, one commit)
                  if (!mma_only && (!row_stage || dx == 0)) LASS_TIMED_WAIT(&b_full[bring0 + sb], resident ? 0u : pb, kProfMmaBFull);
                  tc_fence_after_sync();
                  const uint32_t b_lo = (b_base16 + (bring0 + sb) * b_stage16 + (row_stage ? dx * b_tile16 : 0u)) | kLbo;
                  if (!no_mma && elect_one()) {
                    if (kc == 64) issue_tap<MT, BN, 4, cg2>(acc_addr, a_lo + dx * 8, mt_step16, b_lo, gs, accumulate);
                    else issue_tap<MT, BN, 2, cg2>(acc_addr, a_lo + dx * 4, mt_step16, b_lo, gs, accumulate);
                  }
                  __syncwarp();
                  accumulate = 1;
                  if (resident) {
                    ++sb;
                  } else if (!row_stage || dx == 2) {
                    if (elect_one()) {
                      if (cg2) umma_commit_cg2(&b_empty[bring0 + sb]);
                      else umma_commit(&b_empty[bring0 + sb]);
                    }
                    __syncwarp();
                    if (++sb == n_b) {
                      sb = 0;
                      pb ^= 1;
                    }
                  }
                }
                a_lo += pitch * row16;
              }
            }
            if (elect_one()) {
              if (cg2) umma_commit_cg2(&a_empty[ring0 + sa]);
              else umma_commit(&a_empty[ring0 + sa]);
            }
            __syncwarp();
            if (++sa == n_a) {
              sa = 0;
              pa ^= 1;
            }
          }
        }
        if (elect_one()) {
          if (cg2) umma_commit_cg2(&acc_full[as]);
          else umma_commit(&acc_full[as]);
        }
        __syncwarp();
        ++n_items;
        first_item = false;
      }
      if (mma_only) {
        // all MMAs done before the TMEM is released: one more commit on a barrier nobody else uses in this mode
        if (elect_one()) umma_commit(&b_empty[mw]);
        __syncwarp();
        mbar_wait(&b_empty[mw], 0);
      }
      if (prof && lane == 0 && mw == 0) {
        long long* dst = p.prof + (size_t)blockIdx.x * kProfSlots;
        dst[kProfMmaAccEmpty] = pc[kProfMmaAccEmpty];
        dst[kProfMmaAFull] = pc[kProfMmaAFull];
        dst[kProfMmaBFull] = pc[kProfMmaBFull];
        dst[kProfMmaTotal] = clock64() - t_start;
        dst[kProfItems] = n_items;
      }
    }
  } else if (!mma_only) {
    // =========================== epilogue ===========================
    // Two groups of four warps; group g drains accumulator stage g (items g, g + 2, ... of this CTA), so two
    // items are in the epilogue at once and every SM sub-partition has two epilogue warps to overlap latencies.
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int et = threadIdx.x & 127;                // thread index within the group
    const int hl = q * 4 + (lane >> 3);
    const int wl = lane & 7;
    const int Ho = p.H * p.up_h, Wo = p.W * p.up_w;
    const int Hp = p.H / p.pool_h, Wp = p.W / p.pool_w;
    const bool pooling = ((p.pool_raw.ptr != nullptr) || (p.pool_act.ptr != nullptr)) && !(p.debug_flags & 8);
    const bool no_store = (p.debug_flags & 16) != 0;
    const float pool_scale = 1.0f / (float)(p.pool_h * p.pool_w);
    EpiTables<BN>* gtabs = tabs + 2 * grp;
    const bool tma_store = p.tma_store != 0;
    const bool tma_pool = p.tma_pool != 0;
    unsigned char* stg = stage_base + (size_t)(warp - 4) * kStageWarpBytes;
    uint32_t n = (uint32_t)grp;                      // index of the item within this CTA's sequence
    int tab_b = -1, tab_n0 = -1;
    uint32_t tab_sel = 0;
    if (p.epi_mode == 1) {
      // ---- lean path: ONE activated bf16 output written with direct stores (first conv of every block, decoder conv2) ----
      // The generic loop below spends ~450 instructions per 32 pixels x 32 channels on run-time feature tests; this one
      // ~150, and the table reads (ld.shared, issued before tcgen05.wait::ld) overlap the TMEM load.
      uint16_t* const out = reinterpret_cast<uint16_t*>(p.full_act.ptr) + p.full_act.coff;
      const int cstride = p.full_act.cstride;
      const bool idle = (p.debug_flags & 1) != 0;
      const bool st256 = p.full_act.st256 != 0;
      if (kSplitMT) n = 0;
      for (int item = vbid + (kSplitMT ? 0 : grp) * vgrid; item < vitems;
           item += (kSplitMT ? 1 : 2) * vgrid, n += (kSplitMT ? 1 : 2)) {
        const uint32_t as = n % NS, acc_parity = (n / NS) & 1u;
        const Item it = dec(item);
        if (it.b != tab_b || it.n0 != tab_n0) {
          tab_sel ^= 1u;
          EpiTables<BN>& t = gtabs[tab_sel];
          for (int c = et; c < BN; c += 128) {
            const int nn = it.n0 + c;
            const bool in = nn < p.ncols;
            const float scv = in ? __ldg(p.full_act.scale + nn) : 0.0f;
            const float shv = in ? __ldg(p.full_act.shift + (size_t)it.b * p.full_act.shift_bstride + nn) : 0.0f;
            t.sc_full[c] = scv;
            t.sh_full[c] = fmaf(scv, (in && p.bias) ? __ldg(p.bias + nn) : 0.0f, shv);   // conv bias folded into the shift
          }
          tab_b = it.b;
          tab_n0 = it.n0;
          if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        const uint32_t sc_addr = smem_u32(gtabs[tab_sel].sc_full), sh_addr = smem_u32(gtabs[tab_sel].sh_full);
        const int w = it.w0 + wl;
        uint16_t* dst[MT];
        bool valid[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int h = it.h0 + mt * 16 + hl;
          valid[mt] = (h < p.H) && (w < p.W) && !no_store;
          dst[mt] = out + (((size_t)it.b * p.H + h) * p.W + w) * cstride + it.n0;
        }
        const uint32_t taddr = tmem_base + as * (MT * BN) + (static_cast<uint32_t>(q * 32) << 16);
        const int nchunk = min(BN, p.ncols - it.n0);
        mbar_wait_relaxed(&acc_full[as], acc_parity);
        tc_fence_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < nchunk; c0 += 32) {
          float v[MT][32];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (!kSplitMT || mt == grp) tmem_ld_x32(taddr + mt * BN + c0, v[mt]);
          float4 sc[8], sh[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sc[j] = lds128(sc_addr + (c0 + 4 * j) * 4);
            sh[j] = lds128(sh_addr + (c0 + 4 * j) * 4);
          }
          tmem_ld_wait();
          if (idle) continue;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (!kSplitMT || mt == grp) act_store32(v[mt], sc, sh, dst[mt] + c0, valid[mt], st256);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (cg2 && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[as]), 0));   // the leader's barrier
          else mbar_arrive(&acc_empty[as]);
        }
      }
    } else {
    auto generic_items = [&](auto ftag) {
    // F == kFGeneric: every feature is tested at run time; otherwise F is the exact feature set of the launch and the
    // tests fold away (the specialised copies run ~2x fewer instructions per output tile)
    constexpr uint32_t F = decltype(ftag)::value;
#define FEAT(bit, cond) ((F == kFGeneric) ? (cond) : ((F & (bit)) != 0u))
    if (kSplitMT) n = 0;
    constexpr bool every = kSplitMT;
    // prefetched operands of the rank-1 residual (see below)
    float pf_s[MT], pf_m[MT], pf_t[MT];
    bool pf_have = false;
    auto resid_load = [&](const Item& gi) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int hh = gi.h0 + mt * 16 + hl, ww = gi.w0 + wl;
        const bool on = hh < p.resid_T && ww < p.W;
        pf_s[mt] = on ? __ldg(p.resid_in_scale + ww) : 0.0f;
        pf_m[mt] = on ? __ldg(p.resid_src + ((size_t)gi.b * p.resid_T + hh) * p.resid_F + ww) : 0.0f;
        pf_t[mt] = on ? __ldg(p.resid_in_shift + ww) : 0.0f;
      }
    };
    for (int item = vbid + (kSplitMT ? 0 : grp) * vgrid; item < vitems;
         item += (kSplitMT ? 1 : 2) * vgrid, n += (kSplitMT ? 1 : 2)) {
      const uint32_t as = n % NS, acc_parity = (n / NS) & 1u;
      const Item it = dec(item);
      // ---- (re)stage the per-(clip, N tile) tables; double-buffered so one named barrier per change suffices ----
      if (it.b != tab_b || it.n0 != tab_n0) {
        tab_sel ^= 1u;
        EpiTables<BN>& t = gtabs[tab_sel];
        for (int c = et; c < BN; c += 128) {
          const int n = it.n0 + c;
          const bool in = n < p.ncols;
          const int cc = (FEAT(kFUp, p.up_h * p.up_w > 1)) ? n % p.group_c : n;
          t.bias[c] = ((in && FEAT(kFBias, p.bias != nullptr)) ? __ldg(p.bias + n) : 0.0f) + ((in && FEAT(kFResid, p.resid_src != nullptr)) ? __ldg(p.resid_b + n) : 0.0f);
          t.resid_w[c] = (in && FEAT(kFResid, p.resid_src != nullptr)) ? __ldg(p.resid_w + n) : 0.0f;
          t.sc_full[c] = (in && FEAT(kFAct, p.full_act.scale != nullptr)) ? __ldg(p.full_act.scale + cc) : 0.0f;
          t.sh_full[c] = (in && FEAT(kFAct, p.full_act.scale != nullptr)) ? __ldg(p.full_act.shift + (size_t)it.b * p.full_act.shift_bstride + cc) : 0.0f;
          t.sc_pool[c] = (in && FEAT(kFPoolAct, p.pool_act.scale != nullptr)) ? __ldg(p.pool_act.scale + cc) : 0.0f;
          t.sh_pool[c] = (in && FEAT(kFPoolAct, p.pool_act.scale != nullptr)) ? __ldg(p.pool_act.shift + (size_t)it.b * p.pool_act.shift_bstride + cc) : 0.0f;
        }
        if (FEAT(kFAfter, p.after_w != nullptr) && et < 3 * 32)
          t.after_w[et] = (et % 32 < p.ncols) ? __ldg(p.after_w + (et / 32) * p.ncols + et % 32) : 0.0f;
        if (FEAT(kFAfter, p.after_w != nullptr) && et < 3) t.after_b[et] = __ldg(p.after_b + et);
        tab_b = it.b;
        tab_n0 = it.n0;
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      const EpiTables<BN>& tb = gtabs[tab_sel];
      __builtin_assume(__isShared(&tb));   // table reads become ld.shared instead of generic loads
      // rank-1 residual operand of this thread's pixels.  The magnitudes come from DRAM (they were last touched two launches
      // ago): they are loaded one ITEM ahead and only consumed here, because an in-order warp that multiplies them right
      // after the load sits out the full memory latency -- 36 % of this epilogue's time in the ncu source view
      float resid_xs[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) resid_xs[mt] = 0.0f;
      if (FEAT(kFResid, p.resid_src != nullptr)) {
        if (!pf_have) resid_load(it);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) resid_xs[mt] = fmaf(pf_s[mt], pf_m[mt], pf_t[mt]);
        const int nxt = item + (every ? 1 : 2) * vgrid;
        pf_have = nxt < vitems;
        if (pf_have) resid_load(dec(nxt));
      }
      if (q == 0 && lane == 0 && grp == 0) {
        LASS_TIMED_WAIT_RELAXED(&acc_full[as], acc_parity, kProfEpiAccFull);
      }
      __syncwarp();
      mbar_wait_relaxed(&acc_full[as], acc_parity);
      tc_fence_after_sync();
#pragma unroll 1
      for (int mt = kSplitMT ? grp : 0; mt < (kSplitMT ? grp + 1 : MT); ++mt) {
        const int h = it.h0 + mt * 16 + hl;
        const int w = it.w0 + wl;
        const bool valid = (h < p.H) && (w < p.W) && !(no_store && h >= 0);
        const uint32_t taddr = tmem_base + as * (MT * BN) + mt * BN + (static_cast<uint32_t>(q * 32) << 16);
        float fa0 = 0.0f, fa1 = 0.0f, fa2 = 0.0f;
        const float resid_x = (MT == 2 && mt == 1) ? resid_xs[MT - 1] : resid_xs[0];
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int n = it.n0 + c0;
          if (n >= p.ncols) break;
          float v[32];
          tmem_ld_x32(taddr + c0, v);
          tmem_ld_wait();
          if (p.debug_flags & 1) continue;
          if (FEAT(kFResid, p.resid_src != nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 rw = *reinterpret_cast<const float4*>(tb.resid_w + c0 + j);
              v[j + 0] = fmaf(rw.x, resid_x, v[j + 0]);
              v[j + 1] = fmaf(rw.y, resid_x, v[j + 1]);
              v[j + 2] = fmaf(rw.z, resid_x, v[j + 2]);
              v[j + 3] = fmaf(rw.w, resid_x, v[j + 3]);
            }
          }
          if (FEAT(kFBias, p.bias != nullptr) || FEAT(kFResid, p.resid_src != nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(tb.bias + c0 + j);
              v[j + 0] += bb.x;
              v[j + 1] += bb.y;
              v[j + 2] += bb.z;
              v[j + 3] += bb.w;
            }
          }
          int c = n, ho = h, wo = w;
          if (FEAT(kFUp, p.up_h * p.up_w > 1)) {
            const int g = n / p.group_c;
            c = n - g * p.group_c;
            const int dy = g / p.up_w;
            ho = h * p.up_h + dy;
            wo = w * p.up_w + (g - dy * p.up_w);
          }
          const int hw0 = it.h0 + mt * 16 + q * 4;            // first image row of this warp's 4 x 8 pixel patch
          int grp_dy = 0, grp_dx = 0;
          if (FEAT(kFUp, p.up_h * p.up_w > 1)) {
            const int g = n / p.group_c;
            grp_dy = g / p.up_w;
            grp_dx = g - grp_dy * p.up_w;
          }
          if (FEAT(kFTma, tma_store) && FEAT(kFTmaPool, tma_pool)) {
            // pooled outputs are staged too: the staging buffers must be free before the pooling block below
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
          if (FEAT(kFAfter, p.after_w != nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              fa0 = fmaf(tb.after_w[j], v[j], fa0);
              fa1 = fmaf(tb.after_w[32 + j], v[j], fa1);
              fa2 = fmaf(tb.after_w[64 + j], v[j], fa2);
            }
          }
          if (FEAT(kFPool, pooling)) {
            // Butterfly transpose-reduce: after exchanging with the horizontal neighbour (lane ^ 1) each lane owns the
            // pair sums of 16 of the 32 channels; after the vertical exchange (lane ^ 8) the 2x2 sums of 8 channels.
            // Every lane then finishes and stores its own 8 (or 16) channels of the pooled pixel.
            const bool odd_w = (lane & 1) != 0;
            float s1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float send = odd_w ? v[j] : v[16 + j];
              const float mine = odd_w ? v[16 + j] : v[j];
              s1[j] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            const int hp = h / p.pool_h, wp = w >> 1;
            if (FEAT(kFPoolH2, p.pool_h == 2)) {
              const bool odd_h = (lane & 8) != 0;
              float s2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float send = odd_h ? s1[j] : s1[8 + j];
                const float mine = odd_h ? s1[8 + j] : s1[j];
                s2[j] = (mine + __shfl_xor_sync(0xffffffffu, send, 8)) * pool_scale;
              }
              const int cb = (odd_w ? 16 : 0) + (odd_h ? 8 : 0);     // first of this lane's 8 channels within the chunk
              const int pp = ((lane >> 4) << 2) + (wl >> 1);         // pooled pixel within the warp's 2 x 4 pooled patch
              uint4 praw, pact;
              praw = make_uint4(pack_f16x2_sat(s2[0], s2[1]), pack_f16x2_sat(s2[2], s2[3]), pack_f16x2_sat(s2[4], s2[5]),
                                pack_f16x2_sat(s2[6], s2[7]));
              {
                uint32_t wv[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  const float t0 = fmaf(tb.sc_pool[c0 + cb + j], s2[j], tb.sh_pool[c0 + cb + j]);
                  const float t1 = fmaf(tb.sc_pool[c0 + cb + j + 1], s2[j + 1], tb.sh_pool[c0 + cb + j + 1]);
                  wv[j / 2] = lrelu_bf16x2(pack_bf16x2(t0, t1));
                }
                pact = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              }
              if (FEAT(kFTmaPool, tma_pool)) {
                if (FEAT(kFPoolRaw, p.pool_raw.ptr != nullptr)) *stage_slot(stg + 4096, pp, cb >> 3) = praw;
                if (FEAT(kFPoolAct, p.pool_act.ptr != nullptr)) *stage_slot(stg + 5120, pp, cb >> 3) = pact;
              } else if (valid) {
                if (FEAT(kFPoolRaw, p.pool_raw.ptr != nullptr))
                  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.pool_raw.ptr) +
                                            (((size_t)it.b * Hp + hp) * Wp + wp) * p.pool_raw.cstride + p.pool_raw.coff + c + cb) = praw;
                if (FEAT(kFPoolAct, p.pool_act.ptr != nullptr))
                  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.pool_act.ptr) +
                                            (((size_t)it.b * Hp + hp) * Wp + wp) * p.pool_act.cstride + p.pool_act.coff + c + cb) = pact;
              }
            } else {
              const int cb = odd_w ? 16 : 0;                        // this lane's 16 channels within the chunk
              const int pp = ((lane >> 3) << 2) + (wl >> 1);         // pooled pixel within the warp's 4 x 4 pooled patch
              uint32_t wr[8], wa[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float u0 = s1[j] * pool_scale, u1 = s1[j + 1] * pool_scale;
                wr[j / 2] = pack_f16x2_sat(u0, u1);
                const float t0 = fmaf(tb.sc_pool[c0 + cb + j], u0, tb.sh_pool[c0 + cb + j]);
                const float t1 = fmaf(tb.sc_pool[c0 + cb + j + 1], u1, tb.sh_pool[c0 + cb + j + 1]);
                wa[j / 2] = lrelu_bf16x2(pack_bf16x2(t0, t1));
              }
              if (FEAT(kFTmaPool, tma_pool)) {
                if (FEAT(kFPoolRaw, p.pool_raw.ptr != nullptr)) {
                  *stage_slot(stg + 4096, pp, cb >> 3) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
                  *stage_slot(stg + 4096, pp, (cb >> 3) + 1) = make_uint4(wr[4], wr[5], wr[6], wr[7]);
                }
                if (FEAT(kFPoolAct, p.pool_act.ptr != nullptr)) {
                  *stage_slot(stg + 5120, pp, cb >> 3) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
                  *stage_slot(stg + 5120, pp, (cb >> 3) + 1) = make_uint4(wa[4], wa[5], wa[6], wa[7]);
                }
              } else if (valid) {
                if (FEAT(kFPoolRaw, p.pool_raw.ptr != nullptr)) {
                  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.pool_raw.ptr) +
                                                        (((size_t)it.b * Hp + hp) * Wp + wp) * p.pool_raw.cstride + p.pool_raw.coff + c + cb);
                  dst[0] = make_uint4(wr[0], wr[1], wr[2], wr[3]);
                  dst[1] = make_uint4(wr[4], wr[5], wr[6], wr[7]);
                }
                if (FEAT(kFPoolAct, p.pool_act.ptr != nullptr)) {
                  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.pool_act.ptr) +
                                                        (((size_t)it.b * Hp + hp) * Wp + wp) * p.pool_act.cstride + p.pool_act.coff + c + cb);
                  dst[0] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
                  dst[1] = make_uint4(wa[4], wa[5], wa[6], wa[7]);
                }
              }
            }
            if (FEAT(kFTmaPool, tma_pool)) {
              fence_proxy_async_smem();
              __syncwarp();
              if (elect_one()) {
                if (FEAT(kFPoolRaw, p.pool_raw.ptr != nullptr)) tma_store_4d(&p.tm_out[4], stg + 4096, c, it.w0 >> 1, hw0 / p.pool_h, it.b);
                if (FEAT(kFPoolAct, p.pool_act.ptr != nullptr)) tma_store_4d(&p.tm_out[5], stg + 5120, c, it.w0 >> 1, hw0 / p.pool_h, it.b);
                tma_store_commit();
              }
              __syncwarp();
            }
          }
          // Full-resolution outputs LAST: their staging buffers are reused chunk after chunk, and the pooling work above
          // gives the previous chunk's TMA stores time to read them before this wait.
          const bool coal = FEAT(kFCoal, p.tma_store == 2);   // staged, but written with coalesced st.global instead of TMA
          const bool act_direct = p.tma_store == 3;           // hybrid: raw tensor through TMA, activated tensor st.global
          if (FEAT(kFTma, tma_store) && !FEAT(kFTmaPool, tma_pool)) {
            if (!coal && lane == 0) tma_store_wait_read();
            __syncwarp();                                     // (coal: the previous chunk's staging reads are done)
          }
          if (FEAT(kFRaw, p.full_raw.ptr != nullptr)) {
            uint32_t wv[16];
            pack_raw32(v, wv, p.full_raw.fp16 != 0);
            if (FEAT(kFTma, tma_store)) stage_row32(stg, lane, wv);
            else store32(p.full_raw, it.b, ho, wo, Ho, Wo, c, wv, valid);
          }
          if (FEAT(kFAct, p.full_act.ptr != nullptr)) {
            uint32_t wv[16];
            pack_act32(tb.sc_full + c0, tb.sh_full + c0, v, wv);
            if (FEAT(kFTma, tma_store) && !act_direct) stage_row32(stg + 2048, lane, wv);
            else store32(p.full_act, it.b, ho, wo, Ho, Wo, c, wv, valid);
          }
          if (FEAT(kFTma, tma_store) && coal) {
            // The warp's 4 x 8 pixel patch sits in shared memory, one 64 B row per pixel.  Four lanes now write the four
            // 16 B pieces of ONE pixel, so every st.global covers whole 32 B sectors (8 pixels x 64 B per instruction) --
            // the SM's TMA engine moves only about one 64 B box row per 4-5 cycles, which bounded these launches.
            __syncwarp();
            const int pc = lane & 3, col = lane >> 2;
            const int wpx = it.w0 + col;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int hpx = hw0 + i;
              if (hpx < p.H && wpx < p.W && !no_store) {
                const size_t pix = ((size_t)it.b * Ho + (size_t)(hpx * p.up_h + grp_dy)) * Wo + (size_t)(wpx * p.up_w + grp_dx);
                if (FEAT(kFRaw, p.full_raw.ptr != nullptr))
                  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.full_raw.ptr) + pix * p.full_raw.cstride + p.full_raw.coff + c + pc * 8) =
                      *stage_slot(stg, i * 8 + col, pc);
                if (FEAT(kFAct, p.full_act.ptr != nullptr))
                  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.full_act.ptr) + pix * p.full_act.cstride + p.full_act.coff + c + pc * 8) =
                      *stage_slot(stg + 2048, i * 8 + col, pc);
              }
            }
          } else if (FEAT(kFTma, tma_store) && (FEAT(kFRaw, p.full_raw.ptr != nullptr) || FEAT(kFAct, p.full_act.ptr != nullptr))) {
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              if (FEAT(kFRaw, p.full_raw.ptr != nullptr)) tma_store_5d(&p.tm_out[grp_dy], stg, c, grp_dx, it.w0, hw0, it.b);
              if (FEAT(kFAct, p.full_act.ptr != nullptr) && !act_direct)
                tma_store_5d(&p.tm_out[2 + grp_dy], stg + 2048, c, grp_dx, it.w0, hw0, it.b);
              tma_store_commit();
            }
            __syncwarp();
          }
        }
        if (FEAT(kFAfter, p.after_w != nullptr) && valid && !(p.debug_flags & 1)) {
          const size_t plane = (size_t)p.H * p.W;
          float* fp = p.feat + (size_t)it.b * 3 * plane + (size_t)h * p.W + w;
          fp[0] = fa0 + tb.after_b[0];
          fp[plane] = fa1 + tb.after_b[1];
          fp[2 * plane] = fa2 + tb.after_b[2];
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
          if (cg2 && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[as]), 0));   // the leader's barrier
          else mbar_arrive(&acc_empty[as]);
        }
    }
#undef FEAT
    };
    switch (p.epi_mode) {
      case 2: generic_items(UTag<kFEnc2Resid>{}); break;
      case 3: generic_items(UTag<kFEnc2Bias>{}); break;
      case 4: generic_items(UTag<kFEnc2BiasP12>{}); break;
      case 5: generic_items(UTag<kFUpconv>{}); break;
      case 6: generic_items(UTag<kFAfterBias>{}); break;
      case 7: generic_items(UTag<kFUpDirect>{}); break;
      default: generic_items(UTag<kFGeneric>{}); break;
    }
    }
    if (tma_store && lane == 0) tma_store_wait_all();
    if (prof && q == 0 && lane == 0 && grp == 0) {
      long long* dst = p.prof + (size_t)blockIdx.x * kProfSlots;
      dst[kProfEpiAccFull] = pc[kProfEpiAccFull];
      dst[kProfEpiTotal] = clock64() - t_start;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (cg2) cluster_sync_all();      // the peer's MMAs / remote arrivals are done before shared memory and TMEM go away
  if (warp == 2) {
    __syncwarp();
    tc_fence_after_sync();
    if (cg2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef void (*ConvKernelFn)(const ConvParams);

struct KernelChoice {
  ConvKernelFn fn;
  int BN, MT;
};

template <int BN, int MT, bool CG2 = false>
KernelChoice make_choice() {
  return KernelChoice{conv_igemm_kernel<BN, MT, CG2>, BN, MT};
}

int g_debug_flags = 0;
long long* g_prof_buffer = nullptr;

}  // namespace

struct ConvPrepared {
  ConvParams params;
  ConvKernelFn fn;
  int grid;
  int threads;
  int cluster;   // 2: launched as thread-block clusters of two CTAs (params.cg2)
  int pdl;       // launched with programmatic stream serialization (debug flag 4194304: off)
  size_t smem;
};

void conv_set_debug_flags(int flags) { g_debug_flags = flags; }
void conv_set_profile_buffer(long long* buf) { g_prof_buffer = buf; }

double conv_flops(const ConvLaunch& l) {
  double k = 0;
  for (int s = 0; s < l.nseg; ++s) k += (double)l.seg[s].cin * l.seg[s].taps;
  return 2.0 * l.B * l.H * l.W * (double)l.ncols * k;
}

static void fill_out(OutDev& d, const ConvOut& o) {
  d.ptr = o.ptr;
  d.scale = o.scale;
  d.shift = o.shift;
  d.cstride = o.cstride;
  d.coff = o.coff;
  d.fp16 = o.fp16;
  d.shift_bstride = o.shift_bstride;
  // (transposed convs write group_c-channel pieces at channel offsets that are multiples of 32 within cstride, so the same
  // test covers them; debug flag 16384 keeps 16-byte stores)
  d.st256 = (o.ptr && o.cstride % 16 == 0 && o.coff % 16 == 0 && reinterpret_cast<uintptr_t>(o.ptr) % 32 == 0 &&
             !(g_debug_flags & 16384)) ? 1 : 0;
}

static int check_out(const ConvOut& o, const char* name, int ncols_eff) {
  if (!o.ptr) return 0;
  if (o.cstride % 8 || o.coff % 8 || o.coff + ncols_eff > o.cstride)
    return set_error(LASS_ERR_ARG, "conv: output %s has bad channel slice (cstride %d coff %d cols %d)", name, o.cstride,
                     o.coff, ncols_eff);
  if (reinterpret_cast<uintptr_t>(o.ptr) % 16) return set_error(LASS_ERR_ARG, "conv: output %s not 16 B aligned", name);
  if ((o.scale == nullptr) != (o.shift == nullptr)) return set_error(LASS_ERR_ARG, "conv: output %s needs scale AND shift", name);
  const bool is_raw = name[5] == 'r';   // "full_raw" / "pool_raw"
  // (bf16 raw stores exist for the full-resolution output only: the backward-data launches of the training step)
  if (is_raw && ((!o.fp16 && name[0] != 'f') || o.scale)) return set_error(LASS_ERR_ARG, "conv: %s must be fp16 without activation", name);
  if (!is_raw && (o.fp16 || !o.scale)) return set_error(LASS_ERR_ARG, "conv: %s must be bf16 with an activation table", name);
  return 0;
}

int conv_prepare(const ConvLaunch& l, ConvPrepared** out) {
  *out = nullptr;
  if (l.B <= 0 || l.H <= 0 || l.W <= 0 || l.ncols <= 0 || l.ncols % 16 || l.nseg < 1 || l.nseg > 2)
    return set_error(LASS_ERR_ARG, "conv: bad shape B=%d H=%d W=%d ncols=%d nseg=%d", l.B, l.H, l.W, l.ncols, l.nseg);
  const int up = l.up_h * l.up_w;
  if (l.up_h < 1 || l.up_w < 1 || l.up_h > 2 || l.up_w > 2 || l.group_c <= 0 || l.group_c % 16 || l.group_c * up != l.ncols)
    return set_error(LASS_ERR_ARG, "conv: bad upsample spec up=(%d,%d) group_c=%d ncols=%d", l.up_h, l.up_w, l.group_c, l.ncols);
  if ((l.pool_h != 1 && l.pool_h != 2) || (l.pool_w != 1 && l.pool_w != 2) || l.H % l.pool_h || l.W % l.pool_w)
    return set_error(LASS_ERR_ARG, "conv: bad pooling (%d,%d) for %dx%d", l.pool_h, l.pool_w, l.H, l.W);
  if ((l.pool_raw.ptr || l.pool_act.ptr) && up > 1) return set_error(LASS_ERR_ARG, "conv: pooling with upsampling");
  if ((l.pool_raw.ptr || l.pool_act.ptr) && l.pool_w != 2) return set_error(LASS_ERR_ARG, "conv: pooled outputs need pool_w == 2");
  if (l.after_w && (!l.after_b || !l.feat || l.ncols > 256 || up > 1))
    return set_error(LASS_ERR_ARG, "conv: fused after_conv needs after_b, feat and a single N tile");
  if (l.resid_src && (!l.resid_in_scale || !l.resid_in_shift || !l.resid_w || !l.resid_b || up > 1 || l.resid_T <= 0 ||
                      l.resid_T > l.H || l.resid_F < l.W))
    return set_error(LASS_ERR_ARG, "conv: bad rank-1 residual spec");
  if (l.algo != 0) return set_error(LASS_ERR_ARG, "conv: lass_conv_desc.algo is reserved and must be 0");
  int e;
  if ((e = check_out(l.full_raw, "full_raw", l.group_c))) return e;
  if ((e = check_out(l.full_act, "full_act", l.group_c))) return e;
  if ((e = check_out(l.pool_raw, "pool_raw", l.group_c))) return e;
  if ((e = check_out(l.pool_act, "pool_act", l.group_c))) return e;

  // ---- tile configuration ----
  int BN;
  if (l.ncols <= 32) BN = 32;
  else if (l.ncols <= 64) BN = 64;
  else if (l.ncols % 256 == 0) BN = 256;
  else BN = 128;
  // N = 256 tiles take one m-tile.  Two m-tiles (one 512-column accumulator, epilogue not overlapped; debug flag 1024)
  // halve the weight traffic from L2 but are not faster (measured: 512 -> 256 conv 0.228 vs 0.222 ms): at N = 256 the MMA
  // operand reads alone take 96 of the 128 B/clk of shared-memory bandwidth, and the TMA fill needs most of the rest --
  // the fix for these layers is cta_group::2 (each CTA holds half of the weight tile), not a larger tile.
  int big_tiles = 0;
  for (int s = 0; s < l.nseg; ++s)
    if (l.seg[s].kc > 0) big_tiles += (l.seg[s].cin / l.seg[s].kc) * l.seg[s].taps;
  int MT = (BN == 256) ? ((big_tiles >= 36 && (g_debug_flags & 1024)) ? 2 : 1) : 2;
  if (l.H < 32 || l.H % 32) MT = 1;
  if (MT == 2 && l.ncols <= BN) {
    // Resident weights (no per-tap ring handshake in the MMA issuer) beat the larger tile: if the weights of an item
    // fit next to two A stages only with one m-tile per item, take MT = 1 (measured on the 128 -> 64 decoder conv).
    size_t b_stage_max = 0, a2 = 0, a1 = 0;
    int tiles = 0;
    for (int s = 0; s < l.nseg; ++s) {
      const size_t row = (size_t)l.seg[s].kc * 2;
      const bool halo = l.seg[s].taps == 9;
      const size_t t2 = ((halo ? (size_t)34 * kHaloPitch : (size_t)32 * TW) * row + 1023) & ~size_t(1023);
      const size_t t1 = ((halo ? (size_t)18 * kHaloPitch : (size_t)16 * TW) * row + 1023) & ~size_t(1023);
      if (t2 > a2) a2 = t2;
      if (t1 > a1) a1 = t1;
      const size_t bt = ((size_t)BN * row + 1023) & ~size_t(1023);
      if (bt > b_stage_max) b_stage_max = bt;
      if (l.seg[s].kc > 0) tiles += (l.seg[s].cin / l.seg[s].kc) * l.seg[s].taps;
    }
    const size_t avail = 220 * 1024 - (12 * 1024 + 4 * ((size_t)6 * BN + 100) * sizeof(float));
    const size_t bw = (size_t)tiles * b_stage_max;
    if (tiles <= kMaxB && 2 * a2 + bw > avail && 2 * a1 + bw <= avail) MT = 1;
  }
  // Small grids (small batches, the deep levels): a launch whose items cover less than half of the SMs streams its weights
  // through a handful of CTAs, each at the ~50 B/clk one SM gets from L2 -- 30-47 us for the 384-channel levels at batch 1 against
  // 1-5 us of tensor work.  Narrower tiles (one m-tile, N down to 32) spread the same weight bytes over up to 8x as many SMs;
  // the accumulation order of an output element does not depend on the tile, so results stay bit-identical across batch sizes.
  // Streamed-weight convolutions only (ncols >= 128; resident weights want their single N tile); debug flag 262144 = off.
  // (not the decoder conv2 launches: their 1x1 shortcut segment over the 2C-channel concat makes an item a chain of
  //  activation-tile loads that every additional N tile repeats -- measured 33 -> 41 us at batch 1)
  const bool a_bound = l.nseg == 2 && l.seg[1].cin > l.seg[0].cin;
  if (!(g_debug_flags & 262144) && up == 1 && !l.after_w && !l.gen_src && l.ncols >= 128 && !a_bound) {
    const long long sms = device_sm_count();
    auto n_items = [&](int bn, int mt) {
      return (long long)l.B * ((l.H + 16 * mt - 1) / (16 * mt)) * ((l.W + TW - 1) / TW) * ((l.ncols + bn - 1) / bn);
    };
    while (n_items(BN, MT) * 2 <= sms) {
      if (MT == 2) MT = 1;
      else if (BN > 32 && l.ncols % (BN / 2) == 0) BN /= 2;
      else break;
    }
  }
  if (l.after_w && BN < l.ncols) return set_error(LASS_ERR_ARG, "conv: fused after_conv needs ncols <= BN");
  KernelChoice kc;
  if (BN == 32) kc = MT == 2 ? make_choice<32, 2>() : make_choice<32, 1>();
  else if (BN == 64) kc = MT == 2 ? make_choice<64, 2>() : make_choice<64, 1>();
  else if (BN == 128) kc = MT == 2 ? make_choice<128, 2>() : make_choice<128, 1>();
  else kc = MT == 2 ? make_choice<256, 2>() : make_choice<256, 1>();
  // CTA-pair variants (cta_group::2) of the N >= 128 tiles, taken below when the weights are streamed
  KernelChoice kc_pair = kc;
  if (BN == 128) kc_pair = MT == 2 ? make_choice<128, 2, true>() : make_choice<128, 1, true>();
  else if (BN == 256 && MT == 1) kc_pair = make_choice<256, 1, true>();
  else if (BN == 64) kc_pair = MT == 2 ? make_choice<64, 2, true>() : make_choice<64, 1, true>();
  else if (BN == 32 && MT == 2) kc_pair = make_choice<32, 2, true>();

  ConvPrepared* cp = new (std::nothrow) ConvPrepared();
  if (!cp) return set_error(LASS_ERR_ARG, "conv: out of host memory");
  ConvParams& p = cp->params;
  memset(&p, 0, sizeof(p));
  p.nseg = l.nseg;
  p.B = l.B;
  p.H = l.H;
  p.W = l.W;
  p.ncols = l.ncols;
  p.debug_flags = g_debug_flags;
  p.prof = g_prof_buffer;
  p.tiles_h = (l.H + 16 * MT - 1) / (16 * MT);
  p.tiles_w = (l.W + TW - 1) / TW;
  p.pix_tiles = l.B * p.tiles_h * p.tiles_w;
  p.n_tiles = (l.ncols + BN - 1) / BN;
  p.num_items = p.pix_tiles * p.n_tiles;
  p.bias = l.bias;
  p.up_h = l.up_h;
  p.up_w = l.up_w;
  p.group_c = l.group_c;
  p.pool_h = l.pool_h;
  p.pool_w = l.pool_w;
  p.after_w = l.after_w;
  p.after_b = l.after_b;
  p.feat = l.feat;
  p.resid_src = l.resid_src;
  p.resid_in_scale = l.resid_in_scale;
  p.resid_in_shift = l.resid_in_shift;
  p.resid_w = l.resid_w;
  p.resid_b = l.resid_b;
  p.resid_T = l.resid_T;
  p.resid_F = l.resid_F;
  if (l.gen_src) {
    if (!l.gen_in_scale || !l.gen_in_shift || !l.gen_w || !l.gen_b || !l.gen_scale || !l.gen_shift || l.gen_T <= 0 ||
        l.gen_T > l.H || l.gen_F < l.W || l.seg[0].cin != 32 || l.seg[0].kc != 32 || l.seg[0].taps != 9 || l.seg[0].fp16) {
      delete cp;
      return set_error(LASS_ERR_ARG, "conv: bad generated-A spec (needs a 32-channel bf16 3x3 segment 0)");
    }
    p.gen_src = l.gen_src;
    p.gen_in_scale = l.gen_in_scale;
    p.gen_in_shift = l.gen_in_shift;
    p.gen_w = l.gen_w;
    p.gen_b = l.gen_b;
    p.gen_scale = l.gen_scale;
    p.gen_shift = l.gen_shift;
    p.gen_shift_bstride = l.gen_shift_bstride;
    p.gen_T = l.gen_T;
    p.gen_F = l.gen_F;
  }
  fill_out(p.full_raw, l.full_raw);
  fill_out(p.full_act, l.full_act);
  fill_out(p.pool_raw, l.pool_raw);
  fill_out(p.pool_act, l.pool_act);

  uint32_t a_stage = 0, b_stage = 0;
  int b_tiles_per_item = 0;
  for (int s = 0; s < l.nseg; ++s) {
    const ConvSegment& sg = l.seg[s];
    const bool generated = (s == 0 && l.gen_src != nullptr);
    if ((!sg.src && !generated) || !sg.weights || (sg.kc != 32 && sg.kc != 64) || sg.cin <= 0 || sg.cin % sg.kc ||
        (sg.taps != 9 && sg.taps != 1) ||
        (!generated && (sg.src_cstride % 8 || sg.src_coff % 8 || sg.src_coff + sg.cin > sg.src_cstride))) {
      delete cp;
      return set_error(LASS_ERR_ARG, "conv: bad segment %d (cin %d kc %d taps %d cstride %d coff %d)", s, sg.cin, sg.kc,
                       sg.taps, sg.src_cstride, sg.src_coff);
    }
    SegDev& d = p.seg[s];
    d.nchunks = sg.cin / sg.kc;
    d.taps = sg.taps;
    d.kc = sg.kc;
    d.fmt = sg.fp16 ? kFmtF16 : kFmtBF16;
    const CUtensorMapSwizzle swz = sg.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const bool halo = sg.taps == 9;
    if (!generated) {
      const char* base = reinterpret_cast<const char*>(sg.src) + (size_t)sg.src_coff * 2;
      uint64_t dims[4] = {(uint64_t)sg.cin, (uint64_t)l.W, (uint64_t)l.H, (uint64_t)l.B};
      uint64_t strides[3] = {(uint64_t)sg.src_cstride * 2, (uint64_t)sg.src_cstride * 2 * l.W,
                             (uint64_t)sg.src_cstride * 2 * l.W * l.H};
      uint32_t box[4] = {(uint32_t)sg.kc, (uint32_t)(halo ? kHaloPitch : TW), (uint32_t)(halo ? 16 * MT + 2 : 16 * MT), 1};
      if ((e = make_tensor_map(&d.tmA, base, 2, 4, dims, strides, box, swz))) {
        delete cp;
        return e;
      }
    }
    {
      uint64_t dims[3] = {(uint64_t)sg.cin, (uint64_t)l.ncols, (uint64_t)sg.taps};
      uint64_t strides[2] = {(uint64_t)sg.cin * 2, (uint64_t)sg.cin * 2 * l.ncols};
      uint32_t box[3] = {(uint32_t)sg.kc, (uint32_t)BN, 1};
      if ((e = make_tensor_map(&d.tmB, sg.weights, 2, 3, dims, strides, box, swz))) {
        delete cp;
        return e;
      }
    }
    const uint32_t row_bytes = sg.kc * 2;
    const uint32_t a_bytes = halo ? (16 * MT + 2) * kHaloPitch * row_bytes : 16 * MT * TW * row_bytes;
    if (a_bytes > a_stage) a_stage = a_bytes;
    if (BN * row_bytes > b_stage) b_stage = BN * row_bytes;
    b_tiles_per_item += d.nchunks * d.taps;
  }
  p.a_stage_bytes = (a_stage + 1023u) & ~1023u;
  p.b_stage_bytes = (b_stage + 1023u) & ~1023u;

  // ---- shared-memory budget: weights resident if every tile of an item fits, else a streaming ring; the per-warp
  //      TMA-store staging (48 KiB) is taken when it still leaves a healthy pipeline ----
  const size_t kBudget = 220 * 1024;
  const size_t fixed_base = 1024 /*alignment slack*/ + (2 * kMaxA + 2 * kMaxB + 8) * 8 + 64 +
                            4 * ((size_t)6 * BN + 3 * 32 + 4) * sizeof(float) + 64;
  const size_t stage_bytes = 1024 + (size_t)kEpiWarps * kStageWarpBytes;
  const bool has_16bit_out = l.full_raw.ptr || l.full_act.ptr || l.pool_raw.ptr || l.pool_act.ptr;
  size_t fixed = fixed_base;
  // Measured (tools/gpu_conv_timing.py, whole-model launch lists): the per-warp TMA-store path wins where a warp's
  // destination is scattered — transposed convs (64 B pieces at a 2-pixel stride, -30 %) and channel slices of the
  // concat buffers (-15 %) — and loses a few per cent against direct 16 B stores for whole-pixel outputs, where loads
  // and stores then compete for the SM's TMA request rate (~1 box row / 4 clk).
  const bool sliced = (l.full_raw.ptr && l.full_raw.cstride != l.group_c) || (l.full_act.ptr && l.full_act.cstride != l.group_c);
  // ... except the 32-channel transposed conv: its 64 B pieces are bound by the TMA row rate, and two 32-byte st.global per
  // piece (whole sectors) beat it (-7 %); with 16-byte stores they did not.
  const bool up32_direct = up > 1 && l.group_c == 32 && (!l.full_raw.ptr || p.full_raw.st256) && (!l.full_act.ptr || p.full_act.st256) &&
                           !(g_debug_flags & 65536);
  const bool want_tma = has_16bit_out && (up > 1 || sliced) && !up32_direct && !(g_debug_flags & 32);
  // preference order: resident weights + TMA stores, resident weights, streaming + TMA stores, streaming
  for (int attempt = 0; attempt < 4; ++attempt) {
    const bool try_resident = attempt < 2;
    p.tma_store = ((attempt & 1) == 0 && want_tma) ? 1 : 0;
    if ((attempt & 1) == 0 && !want_tma) continue;
    fixed = fixed_base + (p.tma_store ? stage_bytes : 0);
    const size_t min_a = 2 * (size_t)p.a_stage_bytes;
    if (try_resident) {
      if (!(p.n_tiles == 1 && b_tiles_per_item <= kMaxB &&
            fixed + min_a + (size_t)b_tiles_per_item * p.b_stage_bytes <= kBudget))
        continue;
      p.b_resident = 1;
      p.b_stages = b_tiles_per_item;
      size_t rest = kBudget - fixed - (size_t)p.b_stages * p.b_stage_bytes;
      p.a_stages = (int)(rest / p.a_stage_bytes);
      if (p.a_stages > kMaxA) p.a_stages = kMaxA;      // deep A rings where the weights leave room (enc0.c2 -5 % vs 4 stages)
      break;
    }
    p.b_resident = 0;
    p.a_stages = 2;
    size_t rest = kBudget > fixed + min_a ? kBudget - fixed - min_a : 0;
    p.b_stages = (int)(rest / p.b_stage_bytes);
    if (p.b_stages > 12) p.b_stages = 12;
    if (p.b_stages >= 5 || !p.tma_store) break;      // staging would starve the weight ring: fall back to direct stores
  }
  // CTA pairs for streamed weights with N >= 128 (debug flag 4096 switches them off): two neighbouring pixel tiles of one
  // N tile per pair, every MMA M = 256 over both CTAs, each CTA streams and holds only HALF of every weight tile -- the
  // operand reads per CTA and MMA drop from 4 KiB + 32 N to 4 KiB + 16 N bytes and the weight fill per CTA halves.
  p.cg2 = 0;
  p.b_tps = 1;
  // weight tensor maps for `rows` couts and `tps` taps per box; returns the largest stage (bytes) over the segments
  auto weight_maps = [&](int rows, int tps, uint32_t* stage_out) -> int {
    uint32_t stage = 0;
    for (int s = 0; s < l.nseg; ++s) {
      const ConvSegment& sg = l.seg[s];
      const CUtensorMapSwizzle swz = sg.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      const int t = sg.taps == 9 ? tps : 1;
      uint64_t dims[3] = {(uint64_t)sg.cin, (uint64_t)l.ncols, (uint64_t)sg.taps};
      uint64_t strides[2] = {(uint64_t)sg.cin * 2, (uint64_t)sg.cin * 2 * l.ncols};
      uint32_t box[3] = {(uint32_t)sg.kc, (uint32_t)rows, (uint32_t)t};
      int err = make_tensor_map(&p.seg[s].tmB, sg.weights, 2, 3, dims, strides, box, swz);
      if (err) return err;
      const uint32_t bytes = (uint32_t)t * rows * sg.kc * 2;
      if (bytes > stage) stage = bytes;
    }
    *stage_out = (stage + 1023u) & ~1023u;
    return 0;
  };
  // Not for the transposed convs: they are bound by their stores, and the coupled pair loses (measured +12 %).
  const bool pair_ok = kc_pair.fn != kc.fn && (p.pix_tiles % 2) == 0 && !l.gen_src && !(g_debug_flags & 4096) &&
                       l.ncols % BN == 0 && up == 1;
  // Resident weights (half of every tile per CTA) pair up only where the MMA issue rate is the bound -- activated-output-only
  // launches with N <= 64 and >= 36 MMAs per m-tile: the decoder's 128 -> 64 conv1 (-15 %), its conv2 (-17 %) and the 64 -> 32
  // conv1 (-5 %).  Launches bound by their epilogue or stores lose 10-60 % in a coupled pair (measured; flag 32768 forces it).
  int mmas_per_mtile = 0;
  for (int s = 0; s < l.nseg; ++s) mmas_per_mtile += (l.seg[s].cin / 16) * l.seg[s].taps;
  const bool act_only = l.full_act.ptr && !l.full_raw.ptr && !l.pool_raw.ptr && !l.pool_act.ptr && !l.after_w && !l.resid_src;
  if (p.b_resident && pair_ok && ((act_only && BN <= 64 && mmas_per_mtile >= 36) || (g_debug_flags & 32768))) {
    p.cg2 = 1;
    if ((e = weight_maps(BN / 2, 1, &p.b_stage_bytes))) {
      delete cp;
      return e;
    }
    const size_t rest = kBudget - fixed - (size_t)p.b_stages * p.b_stage_bytes;
    p.a_stages = (int)(rest / p.a_stage_bytes);
    if (p.a_stages > 4) p.a_stages = 4;
    kc = kc_pair;
  }
  if (!p.b_resident && BN >= 128 && pair_ok) {
    p.cg2 = 1;
    if ((e = weight_maps(BN / 2, 1, &p.b_stage_bytes))) {
      delete cp;
      return e;
    }
    // a decoder conv2's shortcut segment (1x1 over the 2C-channel concat) moves twice the activation bytes of the 3x3 segment
    // for an eighth of its MMAs: a third A stage keeps its loads ahead (measured -5..-9 %; +2 % on the other launches)
    if (l.nseg == 2 && l.seg[1].taps == 1 && l.seg[1].cin > l.seg[0].cin) p.a_stages = 3;
    const size_t min_a = (size_t)p.a_stages * p.a_stage_bytes;
    const size_t rest = kBudget > fixed + min_a ? kBudget - fixed - min_a : 0;
    // what the halved weight stages free goes to a deeper weight ring (a third A stage instead: no gain, measured)
    p.b_stages = (int)(rest / p.b_stage_bytes);
    if (p.b_stages > 16) p.b_stages = 16;
    kc = kc_pair;
  }
  // Row stages for streamed 3x3 weights (debug flag 8192 switches them off): the three taps of a kernel row arrive in one
  // TMA box and cost the MMA issuer one full / empty handshake instead of three -- what it does between two MMAs is exposed
  // tensor-pipe time.  Taken when at least three such stages fit.
  if (!p.b_resident && !(g_debug_flags & 8192)) {
    const int rows = p.cg2 ? BN / 2 : BN;
    bool any9 = false;
    uint32_t stage3 = 0;
    for (int s = 0; s < l.nseg; ++s) {
      any9 = any9 || l.seg[s].taps == 9;
      const uint32_t bytes = (uint32_t)(l.seg[s].taps == 9 ? 3 : 1) * rows * l.seg[s].kc * 2;
      if (bytes > stage3) stage3 = bytes;
    }
    stage3 = (stage3 + 1023u) & ~1023u;
    const size_t used = fixed + (size_t)p.a_stages * p.a_stage_bytes;
    const int n3 = kBudget > used ? (int)((kBudget - used) / stage3) : 0;
    if (any9 && n3 >= 3) {
      p.b_tps = 3;
      if ((e = weight_maps(rows, 3, &p.b_stage_bytes))) {
        delete cp;
        return e;
      }
      p.b_stages = n3 > 8 ? 8 : n3;
    }
  }
  // Two MMA issuers + two producers, each pair with half of the rings.  Measured per layer (tools/gpu_conv_timing.py):
  // a win when every issuer keeps two A stages (resident weights, >= 4 stages), and for streamed weights with N <= 128 and
  // long items (>= 36 weight tiles: the second issuer hides the per-tap ring handshakes); a loss with one A stage per
  // issuer and short items, and for N = 256 tiles, where a weight ring of two 32 KiB stages per issuer is too shallow.
  p.dual_issue = 0;
  if (!(g_debug_flags & 128)) {
    // ... or, with one A stage per issuer, for the longer resident items without pooled outputs (measured: the 128 -> 64
    // decoder conv -10 %, the 64 -> 64 + shortcut one -3 %; the encoder conv2 launches and short items lose)
    if (p.b_resident)
      p.dual_issue = ((p.a_stages >= 4 || (p.a_stages >= 2 && b_tiles_per_item >= 11 && !l.pool_raw.ptr && !l.pool_act.ptr)) &&
                      !(p.cg2 && l.nseg == 2)) ? 1 : 0;      // (a pair with a shortcut segment: one issuer, measured)
    // (CTA pairs: one issuer with the whole ring beats two with half each by 15-30 %, measured on every eligible layer)
    else p.dual_issue = (!p.cg2 && BN <= 128 && p.b_stages >= 4 && b_tiles_per_item >= 36) ? 1 : 0;
  }
  if (p.dual_issue) {
    p.a_stages &= ~1;
    if (!p.b_resident) p.b_stages &= ~1;
  }
  if (!p.b_resident && p.b_stages < 2) {
    delete cp;
    return set_error(LASS_ERR_ARG, "conv: tile does not fit in shared memory");
  }
  p.tma_pool = (p.tma_store && ((l.pool_raw.ptr && l.pool_raw.cstride != l.ncols) || (l.pool_act.ptr && l.pool_act.cstride != l.ncols))) ? 1 : 0;
  // debug flag 512: staged outputs leave through coalesced st.global (mode 2) instead of TMA stores.  Measured slower
  // than the TMA stores on every launch of the plan (enc conv2 +20 %, transposed convs +5..20 %), kept for experiments.
  if (p.tma_store && !p.tma_pool && (g_debug_flags & 512)) p.tma_store = 2;
  // debug flag 2048: hybrid -- the raw tensor through TMA stores, the activated one with per-lane st.global
  if (p.tma_store == 1 && !p.tma_pool && (g_debug_flags & 2048) && l.full_raw.ptr && l.full_act.ptr) p.tma_store = 3;
  {
    // epilogue specialisation: the exact feature set of this launch, matched against the compile-time sets of the kernel
    const bool pooling = (l.pool_raw.ptr || l.pool_act.ptr) && !(g_debug_flags & 8);
    uint32_t f = 0;
    if (l.full_raw.ptr) f |= kFRaw;
    if (l.full_act.ptr) f |= kFAct;
    if (p.tma_store) f |= kFTma;
    if (p.tma_store == 2) f |= kFCoal;
    if (p.tma_pool) f |= kFTmaPool;
    if (pooling) f |= kFPool;
    if (pooling && l.pool_h == 2) f |= kFPoolH2;
    if (l.pool_raw.ptr) f |= kFPoolRaw;
    if (l.pool_act.ptr) f |= kFPoolAct;
    if (l.after_w) f |= kFAfter;
    if (l.resid_src) f |= kFResid;
    if (l.bias) f |= kFBias;
    if (up > 1) f |= kFUp;
    if (g_debug_flags & 256) p.epi_mode = 0;
    else if ((f & ~(uint32_t)kFBias) == kFAct) p.epi_mode = 1;          // lean path: one activated output, direct stores
    else if (f == kFEnc2Resid) p.epi_mode = 2;
    else if (f == kFEnc2Bias) p.epi_mode = 3;
    else if (f == kFEnc2BiasP12) p.epi_mode = 4;
    else if (f == kFUpconv) p.epi_mode = 5;
    else if (f == kFAfterBias) p.epi_mode = 6;
    else if (f == kFUpDirect) p.epi_mode = 7;
    else p.epi_mode = 0;
  }
  if (p.tma_store == 1 || p.tma_store == 3) {
    const int Ho = l.H * l.up_h, Wo = l.W * l.up_w;
    auto full_map = [&](CUtensorMap* tm, const ConvOut& o, int dy) -> int {
      const char* base = reinterpret_cast<const char*>(o.ptr) + (size_t)o.coff * 2 + (size_t)dy * Wo * o.cstride * 2;
      uint64_t dims[5] = {(uint64_t)l.group_c, (uint64_t)l.up_w, (uint64_t)l.W, (uint64_t)l.H, (uint64_t)l.B};
      uint64_t strides[4] = {(uint64_t)o.cstride * 2, (uint64_t)l.up_w * o.cstride * 2,
                             (uint64_t)l.up_h * Wo * o.cstride * 2, (uint64_t)Ho * Wo * o.cstride * 2};
      uint32_t box[5] = {32, 1, TW, 4, 1};
      return make_tensor_map(tm, base, 2, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    auto pool_map = [&](CUtensorMap* tm, const ConvOut& o) -> int {
      const int Hp = l.H / l.pool_h, Wp = l.W / l.pool_w;
      const char* base = reinterpret_cast<const char*>(o.ptr) + (size_t)o.coff * 2;
      uint64_t dims[4] = {(uint64_t)l.ncols, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)l.B};
      uint64_t strides[3] = {(uint64_t)o.cstride * 2, (uint64_t)o.cstride * 2 * Wp, (uint64_t)o.cstride * 2 * Wp * Hp};
      uint32_t box[4] = {32, TW / 2, (uint32_t)(l.pool_h == 2 ? 2 : 4), 1};
      return make_tensor_map(tm, base, 2, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    for (int dy = 0; dy < l.up_h; ++dy) {
      if (l.full_raw.ptr && (e = full_map(&p.tm_out[dy], l.full_raw, dy))) { delete cp; return e; }
      if (l.full_act.ptr && (e = full_map(&p.tm_out[2 + dy], l.full_act, dy))) { delete cp; return e; }
    }
    if (p.tma_pool && l.pool_raw.ptr && (e = pool_map(&p.tm_out[4], l.pool_raw))) { delete cp; return e; }
    if (p.tma_pool && l.pool_act.ptr && (e = pool_map(&p.tm_out[5], l.pool_act))) { delete cp; return e; }
  }
  cp->smem = fixed + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes;
  cp->fn = kc.fn;
  const int g_num_sms = device_sm_count();   // of the current device (per-device cache in api.cu)
  // opt every instantiation in to the full 227 KiB (a later, smaller request must not lower the limit that an
  // already prepared launch of the same kernel relies on)
  cudaError_t ce = cudaFuncSetAttribute(reinterpret_cast<const void*>(kc.fn), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (ce != cudaSuccess) {
    delete cp;
    return set_cuda_error(ce, "conv smem attribute");
  }
  // CTAs per SM: limited by shared memory / registers (occupancy query) and by TMEM (AS * MT * BN columns each)
  const int tmem_need = tmem_cols(BN, MT);
  int per_sm = 1;
  cp->threads = kThreadsK;
  cp->pdl = (g_debug_flags & 4194304) ? 0 : 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(kc.fn), kThreadsK, cp->smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  if (per_sm > 512 / tmem_need) per_sm = 512 / tmem_need;
  if (per_sm > 2) per_sm = 2;
  cp->grid = p.num_items < g_num_sms * per_sm ? p.num_items : g_num_sms * per_sm;
  cp->cluster = 1;
  if (p.cg2) {
    // one pair per TPC: as many clusters as the device keeps resident at once (persistent CTAs, static round robin)
    cp->cluster = 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (unsigned)(g_num_sms / 2), 1, 1);
    cfg.blockDim = dim3(kThreadsK, 1, 1);
    cfg.dynamicSmemBytes = cp->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int pairs = 0;
    if (cudaOccupancyMaxActiveClusters(&pairs, reinterpret_cast<const void*>(kc.fn), &cfg) != cudaSuccess || pairs < 1) {
      cudaGetLastError();
      pairs = g_num_sms / 2;
    }
    if (pairs > g_num_sms / 2) pairs = g_num_sms / 2;
    const int vitems = p.num_items / 2;
    cp->grid = 2 * (vitems < pairs ? vitems : pairs);
  }
  *out = cp;
  return 0;
}

int conv_run(const ConvPrepared* cp, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)cp->grid, 1, 1);
  cfg.blockDim = dim3((unsigned)cp->threads, 1, 1);
  cfg.dynamicSmemBytes = cp->smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cp->cluster == 2) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (cp->pdl) {
    // the kernel's prologue may overlap the tail of the previous launch in the stream (it calls griddepcontrol.wait before
    // it touches anything that launch wrote); a predecessor without griddepcontrol.launch_dependents releases it on completion
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)na;
  return set_cuda_error(cudaLaunchKernelEx(&cfg, cp->fn, cp->params), cp->cluster == 2 ? "conv launch (CTA pairs)" : "conv launch");
}

void conv_free(ConvPrepared* p) { delete p; }

}  // namespace lass
