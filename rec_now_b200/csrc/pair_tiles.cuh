// Pair-scoring tiles of the pair kernel (device code shared by pairwise.cu and the dev micro-benchmarks).
//
// pairwise_loss_from_batch.py:117-126 (bpr_loss_func) + TF autodiff, evaluated on 32-wide J blocks that rotate
// through the lanes of a warp with shfl.bfly.
#pragma once
#include "common.cuh"

namespace rn {

enum { M_HASW = 1, M_DIFF = 2, M_RWN = 4, M_WRONG = 8 };

// General (masked) 32x32 tile for one positive row per lane, rotation steps [t0, t1) (multiples of 4; the whole tile is
// [0, 32)).  Lane l meets negative l ^ t at step t, so disjoint step ranges score disjoint pair sets: a tile can be
// split between warps at a granularity of 4 steps.
template <int MODE, bool FULL>
__device__ __forceinline__ void tile_general(const float si, const float yi, const float wpi, const u32 lo, const u32 hi,
                                             const u32 pjm, const float sjm, const float yjm, const float wnjm,
                                             const float c, const int t0, const int t1, float& li, float& gi, u32& cnt,
                                             float& accj) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, WRONG = MODE & M_WRONG;
  float gi_t = 0.f, li_t = 0.f;
#pragma unroll 2
  for (int tb = t0; tb < t1; tb += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = tb + k;
      const float sj = __shfl_xor_sync(0xFFFFFFFFu, sjm, t);
      const float x = si - sj;                           // PW:117 (float32 subtract, as the reference)
      const float xs = x * c;                            // (x * factor) in log2 units
      const float e = mufu_ex2(-fabsf(xs));              // exp(-|x|)
      const float t1p = 1.0f + e;
      const float L = mufu_lg2(t1p);                     // log1p(exp(-|x|)) / ln2
      const float r = mufu_rcp(t1p);
      const float lo2 = fmaxf(-xs, 0.f) + L;             // softplus(-x) / ln2        (PW:120-121, TF stable form)
      float d = (xs >= 0.f ? e : 1.0f) * r;              // sigma(-x)
      bool valid = true;
      if (!FULL) { const u32 pj = pjm ^ (u32)t; valid = (pj >= lo) && (pj < hi); }
      if (WRONG) valid = valid && (x < 0.f);             // PW:200-202  s_i < s_j
      float wv = 1.f;
      if (HASW) {
        wv = wpi;
        if (DIFF) { const float yj = __shfl_xor_sync(0xFFFFFFFFu, yjm, t); wv = (yi - yj) * wpi; }
        if (RWN) { const float wn = __shfl_xor_sync(0xFFFFFFFFu, wnjm, t); wv = wv * wn; valid = valid && (wv > 0.f); }
        d *= wv;
      }
      if (!FULL || WRONG || RWN) d = valid ? d : 0.f;
      if (WRONG || RWN) cnt += valid ? 1u : 0u;
      if (HASW) { if (valid) li_t = fmaf(wv, lo2, li_t); } else { if (valid) li_t += lo2; }
      gi_t += d;
      accj += __shfl_xor_sync(0xFFFFFFFFu, d, t);
    }
  }
  li += li_t; gi += gi_t;
}

// Fast 64x32 tile: both rows of every lane pair with ALL 32 negatives and their pair weight (wv0 / wv1) is constant
// over the tile.  Per pair: ex2 + rcp (SFU), ~11 FP32 ops; per row and tile one lg2 of the product of the (1+e).
// PART: only the rotation steps [ts, te) (multiples of 4) -- see tile_general.
template <bool HASW, bool PART = false, int KB = 4>
__device__ __forceinline__ void tile_fast(const float si0, const float si1, const float wv0, const float wv1,
                                          const float sjm, const float c, float& li0, float& li1, float& gi0,
                                          float& gi1, float& accj, const int ts = 0, const int te = 32) {
  // Four rotation steps (8 independent pair chains) are issued in lock-step so that the SHFL / MUFU latencies of
  // one chain are covered by the other seven: a warp issues in order, without this batching every step exposes
  // its whole dependency chain.
  float p0 = 1.f, p1 = 1.f, m0 = 0.f, m1 = 0.f, g0 = 0.f, g1 = 0.f;
#pragma unroll (PART ? 1 : 32 / KB)
  for (int tb = (PART ? ts : 0); tb < (PART ? te : 32); tb += KB) {
    float sj[KB], x0[KB], x1[KB], e0[KB], e1[KB], t0[KB], t1[KB], d0[KB], d1[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) sj[k] = __shfl_xor_sync(0xFFFFFFFFu, sjm, tb + k);
#pragma unroll
    for (int k = 0; k < KB; ++k) { x0[k] = (si0 - sj[k]) * c; x1[k] = (si1 - sj[k]) * c; }   // PW:117-119, log2 units
#pragma unroll
    for (int k = 0; k < KB; ++k) { e0[k] = mufu_ex2(-fabsf(x0[k])); e1[k] = mufu_ex2(-fabsf(x1[k])); }
#pragma unroll
    for (int k = 0; k < KB; ++k) { t0[k] = 1.0f + e0[k]; t1[k] = 1.0f + e1[k]; }
#pragma unroll
    for (int k = 0; k < KB; ++k) { d0[k] = mufu_rcp(t0[k]); d1[k] = mufu_rcp(t1[k]); }        // sigma(-x) for x < 0
    {                                                                                       // prod (1+e) <= 2^32
      float q0 = (t0[0] * t0[1]) * (t0[2] * t0[3]), q1 = (t1[0] * t1[1]) * (t1[2] * t1[3]);
      if (KB == 8) { q0 *= (t0[4 % KB] * t0[5 % KB]) * (t0[6 % KB] * t0[7 % KB]); q1 *= (t1[4 % KB] * t1[5 % KB]) * (t1[6 % KB] * t1[7 % KB]); }
      p0 *= q0; p1 *= q1;
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      if (x0[k] >= 0.f) d0[k] *= e0[k]; else m0 -= x0[k];          // sigma(-x) = e/(1+e) for x >= 0; max(-x,0)
      if (x1[k] >= 0.f) d1[k] *= e1[k]; else m1 -= x1[k];
      if (HASW) { d0[k] *= wv0; d1[k] *= wv1; }
    }
    g0 += (d0[0] + d0[1]) + (d0[2] + d0[3]);
    g1 += (d1[0] + d1[1]) + (d1[2] + d1[3]);
    if (KB == 8) { g0 += (d0[4 % KB] + d0[5 % KB]) + (d0[6 % KB] + d0[7 % KB]); g1 += (d1[4 % KB] + d1[5 % KB]) + (d1[6 % KB] + d1[7 % KB]); }
    float back[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) back[k] = __shfl_xor_sync(0xFFFFFFFFu, d0[k] + d1[k], tb + k);
    accj += (back[0] + back[1]) + (back[2] + back[3]);
    if (KB == 8) accj += (back[4 % KB] + back[5 % KB]) + (back[6 % KB] + back[7 % KB]);
  }
  const float L0 = m0 + mufu_lg2(p0), L1 = m1 + mufu_lg2(p1);       // sum softplus(-x) / ln2   (PW:120-121)
  li0 += HASW ? wv0 * L0 : L0; li1 += HASW ? wv1 * L1 : L1;
  gi0 += g0; gi1 += g1;
}

}  // namespace rn
