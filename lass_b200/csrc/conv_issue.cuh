// tcgen05.mma issue helpers of the implicit-GEMM conv kernels (shared with the issue-rate microbenchmark in probe.cu).
#pragma once
#include "ptx.cuh"

namespace lass {

constexpr int TW = 8;          // pixels per tile row == rows of one 8-row descriptor group
constexpr int kHaloPitch = TW + 2;  // pixels per image row of the halo tile in shared memory

// One tcgen05.mma with descriptors given as (low word = start address >> 4 | LBO, high word = SBO | version | swizzle).
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The same for a CTA pair (cta_group::2, M = 256: 128 rows from each CTA's shared memory at these offsets, each CTA holds
// half of the N weight rows); issued by the leader CTA only.
__device__ __forceinline__ void umma_lohi_cg2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool CG2>
__device__ __forceinline__ void umma_sel(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  if (CG2) umma_lohi_cg2(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else umma_lohi(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
}

// Per-segment constants of the MMA issuer (uniform registers).
struct SegMma {
  uint32_t a_hi, b_hi;      // descriptor high words
  uint32_t idesc;
};

// All MMAs of one (chunk, tap) for MT m-tiles: KSTEPS k-steps of 16 channels each.  a_lo / b_lo already contain the
// LBO field; start addresses advance by 2 (x16 B) per k-step and by mt_step16 per m-tile.
template <int MT, int BN, int KSTEPS, bool CG2 = false>
__device__ __forceinline__ void issue_tap(uint32_t acc, uint32_t a_lo, uint32_t mt_step16, uint32_t b_lo, const SegMma& g,
                                          uint32_t accumulate) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      if (CG2)
        umma_lohi_cg2(acc + mt * BN, a_lo + mt * mt_step16 + 2 * ks, g.a_hi, b_lo + 2 * ks, g.b_hi, g.idesc,
                      ks == 0 ? accumulate : 1u);
      else
        umma_lohi(acc + mt * BN, a_lo + mt * mt_step16 + 2 * ks, g.a_hi, b_lo + 2 * ks, g.b_hi, g.idesc,
                  ks == 0 ? accumulate : 1u);
    }
  }
}

// Same MMAs for all nine taps of a halo chunk, but with RUNNING descriptor low words that are advanced in place by
// compile-time deltas (opaque to the optimiser), so that the issuing thread needs one uniform add per operand and MMA
// instead of re-deriving every descriptor from the chunk base (which costs uniform-register moves and spills).
template <int D>
__device__ __forceinline__ void desc_add(uint32_t& lo) {
  asm volatile("add.s32 %0, %0, %1;" : "+r"(lo) : "n"(D));
}
__device__ __forceinline__ void desc_add_r(uint32_t& lo, uint32_t d) {
  asm volatile("add.s32 %0, %0, %1;" : "+r"(lo) : "r"(d));
}
template <int MT, int BN, int KSTEPS, int TP, bool CG2>
__device__ __forceinline__ void issue_halo_tap_running(uint32_t acc, uint32_t& a, uint32_t& b, uint32_t b_next_tap,
                                                       const SegMma& g, uint32_t accumulate) {
  constexpr int ROW16 = 2 * KSTEPS;                       // bytes of one pixel row / 16
  constexpr int MT_STEP = 16 * kHaloPitch * ROW16;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    umma_sel<CG2>(acc, a, g.a_hi, b, g.b_hi, g.idesc, (TP == 0 && ks == 0) ? accumulate : 1u);
    if (MT == 2) {
      desc_add<MT_STEP>(a);
      umma_sel<CG2>(acc + BN, a, g.a_hi, b, g.b_hi, g.idesc, (TP == 0 && ks == 0) ? accumulate : 1u);
    }
    if (ks + 1 < KSTEPS) {
      desc_add<2 - (MT - 1) * MT_STEP>(a);
      desc_add<2>(b);
    }
  }
  if (TP + 1 < 9) {
    constexpr int cur = ((TP / 3) * kHaloPitch + TP % 3) * ROW16;
    constexpr int nxt = (((TP + 1) / 3) * kHaloPitch + (TP + 1) % 3) * ROW16;
    desc_add<nxt - cur - 2 * (KSTEPS - 1) - (MT - 1) * MT_STEP>(a);
    desc_add_r(b, b_next_tap);
  }
}
template <int MT, int BN, int KSTEPS, bool CG2 = false>
__device__ __forceinline__ void issue_halo_chunk_running(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t b_stage16,
                                                         const SegMma& g, uint32_t accumulate) {
  uint32_t a = a_lo, b = b_lo;
  const uint32_t b_next_tap = b_stage16 - 2 * (KSTEPS - 1);
  issue_halo_tap_running<MT, BN, KSTEPS, 0, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 1, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 2, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 3, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 4, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 5, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 6, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 7, CG2>(acc, a, b, b_next_tap, g, accumulate);
  issue_halo_tap_running<MT, BN, KSTEPS, 8, CG2>(acc, a, b, b_next_tap, g, accumulate);
}

}  // namespace lass
