// Pair-scoring tiles of the pair kernel (device code shared by pairwise.cu and the dev micro-benchmarks).
//
// pairwise_loss_from_batch.py:117-126 (bpr_loss_func) + TF autodiff, evaluated on 32-wide J blocks that rotate
// through the lanes of a warp with shfl.bfly.
#pragma once
#include "common.cuh"

namespace rn {

enum { M_HASW = 1, M_DIFF = 2, M_RWN = 4, M_WRONG = 8, M_LUT = 64, M_LAMBDA = 128 };      // (16, 32: dispatch-only bits of pairwise.cu)
// M_LAMBDA (RN_LABEL_LAMBDA, with M_DIFF on the 2^y label column): the pair weight also carries |D_i - D_j|, the difference
// of the rows' rank discounts (the negatives' discounts travel in the negative-side weight column, the row's own in `di`);
// every tile is a general tile then -- the weight changes from pair to pair.

// Label part of a pair weight under M_DIFF: the difference of the (transformed) labels, or -- M_LUT, RN_LABEL_LUT -- the
// entry of the 8 x 8 level table (shared memory; the sorted label column then holds the label LEVEL 0 .. 7 as a float).
template <bool LUT>
__device__ __forceinline__ float label_weight(const float* lut, const float yi, const float yj) {
  if (LUT) return lut[(int)yi * 8 + (int)yj];
  return yi - yj;
}

// General (masked) 32x32 tile for one positive row per lane, rotation steps [t0, t1) (multiples of 4; the whole tile is
// [0, 32)).  Lane l meets negative l ^ t at step t, so disjoint step ranges score disjoint pair sets: a tile can be
// split between warps at a granularity of 4 steps.
// HINGE: the pair loss is max(0, margin - x) with x = (s_i - s_j) * factor (c = factor, plain units) instead of the
// logistic softplus(-x): its "sigma" is the step [margin - x > 0], no SFU operation at all.
template <int MODE, bool FULL, bool HINGE = false>
__device__ __forceinline__ void tile_general(const float si, const float yi, const float wpi, const u32 lo, const u32 hi,
                                             const u32 pjm, const float sjm, const float yjm, const float wnjm,
                                             const float c, const int t0, const int t1, float& li, float& gi, u32& cnt,
                                             float& accj, const float margin = 0.f, const float* lut = nullptr,
                                             const float di = 0.f) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, WRONG = MODE & M_WRONG, LUT = MODE & M_LUT;
  constexpr bool LAMBDA = MODE & M_LAMBDA;
  float gi_t = 0.f, li_t = 0.f;
#pragma unroll 2
  for (int tb = t0; tb < t1; tb += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = tb + k;
      const float sj = __shfl_xor_sync(0xFFFFFFFFu, sjm, t);
      const float x = si - sj;                           // PW:117 (float32 subtract, as the reference)
      const float xs = x * c;                            // (x * factor) in log2 units
      float lo2, d;
      if (HINGE) {
        const float v = fmaf(-x, c, margin);             // hinge: max(0, margin - x), d/dx = -[margin - x > 0] (as tile_hinge)
        lo2 = fmaxf(v, 0.f); d = v > 0.f ? 1.0f : 0.f;
      } else {
        const float e = mufu_ex2(-fabsf(xs));            // exp(-|x|)
        const float t1p = 1.0f + e;
        const float L = mufu_lg2(t1p);                   // log1p(exp(-|x|)) / ln2
        const float r = mufu_rcp(t1p);
        lo2 = fmaxf(-xs, 0.f) + L;                       // softplus(-x) / ln2        (PW:120-121, TF stable form)
        d = (xs >= 0.f ? e : 1.0f) * r;                  // sigma(-x)
      }
      bool valid = true;
      if (!FULL) { const u32 pj = pjm ^ (u32)t; valid = (pj >= lo) && (pj < hi); }
      if (WRONG) valid = valid && (x < 0.f);             // PW:200-202  s_i < s_j
      float wv = 1.f;
      if (HASW) {
        wv = wpi;
        if (DIFF) { const float yj = __shfl_xor_sync(0xFFFFFFFFu, yjm, t); wv = label_weight<LUT>(lut, yi, yj) * wpi; }
        if (LAMBDA) { const float dj = __shfl_xor_sync(0xFFFFFFFFu, wnjm, t); wv *= fabsf(di - dj); }
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

// Fast 64x32 tile of the hinge loss max(0, margin - x): same contract as tile_fast (both rows of a lane pair with all
// 32 negatives, pair weight constant over the tile, out-of-range negatives carry the sentinel score -3e38 so that x = +huge
// and the pair is inactive).  No SFU operation: ~7 FP32 / compare operations per pair.
template <bool PART = false>
__device__ __forceinline__ void tile_hinge(const float si0, const float si1, const float wv0, const float wv1,
                                           const float sjm, const float c, const float margin, float& li0, float& li1,
                                           float& gi0, float& gi1, float& accj, const int ts = 0, const int te = 32) {
  float l0 = 0.f, l1 = 0.f, g0 = 0.f, g1 = 0.f;
#pragma unroll (PART ? 1 : 8)
  for (int tb = (PART ? ts : 0); tb < (PART ? te : 32); tb += 4) {
    float sj[4], back[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sj[k] = __shfl_xor_sync(0xFFFFFFFFu, sjm, tb + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // margin - (s_i - s_j) * c: the difference first, as the reference (PW:117) -- tied scores give exactly `margin`
      const float v0 = fmaf(sj[k] - si0, c, margin), v1 = fmaf(sj[k] - si1, c, margin);
      const bool a0 = v0 > 0.f, a1 = v1 > 0.f;
      l0 += a0 ? v0 : 0.f; l1 += a1 ? v1 : 0.f;
      g0 += a0 ? 1.0f : 0.f; g1 += a1 ? 1.0f : 0.f;
      back[k] = (a0 ? wv0 : 0.f) + (a1 ? wv1 : 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) back[k] = __shfl_xor_sync(0xFFFFFFFFu, back[k], tb + k);
    accj += (back[0] + back[1]) + (back[2] + back[3]);
  }
  li0 = fmaf(wv0, l0, li0); li1 = fmaf(wv1, l1, li1);
  gi0 = fmaf(wv0, g0, gi0); gi1 = fmaf(wv1, g1, gi1);
}

// ---- product-form fast tile ------------------------------------------------------------------------------------
// Packed FP32x2 arithmetic (sm_100a FMUL2 / FADD2 / FFMA2: two FP32 operations per issue slot).
__device__ __forceinline__ u64 f2_pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// |c * (s - m)| bound under which a tile may use the product form: u <= 2^24, four factors (1 + u) <= 2^97
constexpr float kProdRange = 12.0f;

// The same 64x32 tile as tile_fast with ONE SFU operation per pair.  With a reference score m of the tile,
//   2^(-c (s_i - s_j)) = E_i * F_j,   E_i = 2^(-c (s_i - m))  (once per row),  F_j = 2^(c (s_j - m))  (once per negative),
// so exp(-x) of a pair is one multiplication; sigma(-x) = u / (1 + u) needs the reciprocal (the one SFU operation) and
// softplus(-x) = log(1 + u) is taken once per row and tile on the running product of the (1 + u), whose exponent is
// split off after every four factors (integer operations) so that it cannot overflow.  Preconditions (checked by the
// caller, else tile_fast): |c (s - m)| <= kProdRange for every row and negative of the tile; F = 0 for the negatives
// outside the overlap.  The arithmetic runs on pairs of rotation steps in packed FP32x2 instructions.
template <bool PART = false>
__device__ __forceinline__ void tile_prod(const float E0, const float E1, const float wv0, const float wv1, const float Fm,
                                          float& li0, float& li1, float& gi0, float& gi1, float& accj,
                                          const int ts = 0, const int te = 32) {
  const u64 EE0 = f2_pack(E0, E0), EE1 = f2_pack(E1, E1), one2 = f2_pack(1.f, 1.f);
  const u64 W0 = f2_pack(wv0, wv0), W1 = f2_pack(wv1, wv1);
  float p0 = 1.f, p1 = 1.f;                 // mantissas of the running products, in [1, 2)
  u32 x0 = 0, x1 = 0;                       // their exponents (biased, summed)
  u64 g0 = 0, g1 = 0, acc = 0;              // packed partial sums (two rotation steps side by side)
#pragma unroll (PART ? 1 : 8)
  for (int tb = (PART ? ts : 0); tb < (PART ? te : 32); tb += 4) {
    float f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = __shfl_xor_sync(0xFFFFFFFFu, Fm, tb + k);
    const u64 fa = f2_pack(f[0], f[1]), fb = f2_pack(f[2], f[3]);
    const u64 u0a = f2_mul(EE0, fa), u0b = f2_mul(EE0, fb), u1a = f2_mul(EE1, fa), u1b = f2_mul(EE1, fb);   // exp(-x)
    const u64 t0a = f2_add(u0a, one2), t0b = f2_add(u0b, one2), t1a = f2_add(u1a, one2), t1b = f2_add(u1b, one2);
    float t0[4], t1[4], r0[4], r1[4];
    f2_unpack(t0a, t0[0], t0[1]); f2_unpack(t0b, t0[2], t0[3]); f2_unpack(t1a, t1[0], t1[1]); f2_unpack(t1b, t1[2], t1[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) { r0[k] = mufu_rcp(t0[k]); r1[k] = mufu_rcp(t1[k]); }
    {
      float ql, qh;
      f2_unpack(f2_mul(t0a, t0b), ql, qh); p0 *= ql * qh;
      f2_unpack(f2_mul(t1a, t1b), ql, qh); p1 *= ql * qh;
      const u32 b0 = __float_as_uint(p0), b1 = __float_as_uint(p1);
      x0 += b0 >> 23; x1 += b1 >> 23;
      p0 = __uint_as_float((b0 & 0x007FFFFFu) | 0x3F800000u); p1 = __uint_as_float((b1 & 0x007FFFFFu) | 0x3F800000u);
    }
    const u64 d0a = f2_mul(u0a, f2_pack(r0[0], r0[1])), d0b = f2_mul(u0b, f2_pack(r0[2], r0[3]));            // sigma(-x)
    const u64 d1a = f2_mul(u1a, f2_pack(r1[0], r1[1])), d1b = f2_mul(u1b, f2_pack(r1[2], r1[3]));
    g0 = f2_add(g0, f2_add(d0a, d0b)); g1 = f2_add(g1, f2_add(d1a, d1b));
    float b[4];
    f2_unpack(f2_fma(d1a, W1, f2_mul(d0a, W0)), b[0], b[1]); f2_unpack(f2_fma(d1b, W1, f2_mul(d0b, W0)), b[2], b[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = __shfl_xor_sync(0xFFFFFFFFu, b[k], tb + k);
    acc = f2_add(acc, f2_add(f2_pack(b[0], b[1]), f2_pack(b[2], b[3])));
  }
  const int nb = ((PART ? te : 32) - (PART ? ts : 0)) >> 2;
  const float L0 = (float)((int)x0 - 127 * nb) + mufu_lg2(p0), L1 = (float)((int)x1 - 127 * nb) + mufu_lg2(p1);
  float lo, hi;
  li0 = fmaf(wv0, L0, li0); li1 = fmaf(wv1, L1, li1);
  f2_unpack(g0, lo, hi); gi0 = fmaf(wv0, lo + hi, gi0);
  f2_unpack(g1, lo, hi); gi1 = fmaf(wv1, lo + hi, gi1);
  f2_unpack(acc, lo, hi); accj += lo + hi;
}

// General (masked) tile in product form: exp(-x) = E_i * F_j as in tile_prod (same range precondition), so a pair costs
// TWO SFU operations (reciprocal, log) instead of three, and no max(-x, 0) branch: softplus(-x) = log(1 + u) directly
// (u <= 2^24).  Not for the wrong-order filter: s_i < s_j must be decided on the scores themselves.
template <int MODE, bool FULL>
__device__ __forceinline__ void tile_general_prod(const float Ei, const float yi, const float wpi, const u32 lo, const u32 hi,
                                                  const u32 pjm, const float Fm, const float yjm, const float wnjm,
                                                  const int t0, const int t1, float& li, float& gi, u32& cnt, float& accj,
                                                  const float* lut = nullptr, const float di = 0.f) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, LUT = MODE & M_LUT;
  constexpr bool LAMBDA = MODE & M_LAMBDA;
  float gi_t = 0.f, li_t = 0.f;
#pragma unroll 2
  for (int tb = t0; tb < t1; tb += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = tb + k;
      const float u = Ei * __shfl_xor_sync(0xFFFFFFFFu, Fm, t);        // exp(-x)
      const float t1p = 1.0f + u;
      const float L = mufu_lg2(t1p);                                   // softplus(-x) / ln2
      float d = u * mufu_rcp(t1p);                                     // sigma(-x)
      bool valid = true;
      if (!FULL) { const u32 pj = pjm ^ (u32)t; valid = (pj >= lo) && (pj < hi); }
      float wv = 1.f;
      if (HASW) {
        wv = wpi;
        if (DIFF) { const float yj = __shfl_xor_sync(0xFFFFFFFFu, yjm, t); wv = label_weight<LUT>(lut, yi, yj) * wpi; }
        if (LAMBDA) { const float dj = __shfl_xor_sync(0xFFFFFFFFu, wnjm, t); wv *= fabsf(di - dj); }
        if (RWN) { const float wn = __shfl_xor_sync(0xFFFFFFFFu, wnjm, t); wv = wv * wn; valid = valid && (wv > 0.f); }
        d *= wv;
      }
      if (!FULL || RWN) d = valid ? d : 0.f;
      if (RWN) cnt += valid ? 1u : 0u;
      if (HASW) { if (valid) li_t = fmaf(wv, L, li_t); } else { if (valid) li_t += L; }
      gi_t += d;
      accj += __shfl_xor_sync(0xFFFFFFFFu, d, t);
    }
  }
  li += li_t; gi += gi_t;
}

// One row of focal_crossentropy_loss (rec_block/focal_loss.py:12-66 of the reference): value and d / d logit.
//   ce = sigmoid_cross_entropy_with_logits(y, z) = max(z, 0) - z y + log1p(exp(-|z|))                  (focal_loss.py:48)
//   alpha factor y alpha + (1 - y)(1 - alpha)                                                          (:50-53)
//   modulating factor (1 - (y p + (1 - y)(1 - p)))^gamma, p = sigmoid(z); optionally without gradient  (:55-62)
// (not inlined: it runs once per row outside the pair loop, and the pair kernel has no registers to spare)
static __device__ __noinline__ float2 focal_row(float z, float y, float alpha, float gamma, int stop) {
  const float e = expf(-fabsf(z));
  const float ce = fmaxf(z, 0.f) - z * y + log1pf(e);
  const float p = z >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
  const float af = alpha != 0.f ? y * alpha + (1.0f - y) * (1.0f - alpha) : 1.0f;
  float mod = 1.0f, dmod = 0.f;
  if (gamma != 0.f) {
    const float om = 1.0f - (y * p + (1.0f - y) * (1.0f - p));
    mod = powf(om, gamma);
    if (!stop) dmod = -gamma * powf(om, gamma - 1.0f) * (2.0f * y - 1.0f) * p * (1.0f - p);
  }
  return make_float2(af * mod * ce, af * (mod * (p - y) + ce * dmod));
}

}  // namespace rn
