// K2/K3 — fused in-batch pairwise logistic (BPR) loss, forward + backward, on the segmented batch.
//
// Replaces pairwise_loss_from_batch.py:254-274 + TF autodiff.  After K1 the rows are
// sorted by (group, label, row), so for a row at sorted position p with group start a(p) and label-level
// start l(p) the negatives of p are exactly the positions [a(p), l(p)): pair enumeration is position
// arithmetic, pair COUNTS are exact integers (l - a), and the kept pairs of a group form a staircase of
// dense rectangles.  The pair kernel walks that staircase in 32x32 micro-tiles, one warp per tile:
// lane = one positive row i (kept in registers), the 32 negatives of the tile rotate through the lanes with
// shfl.bfly, so each lane meets each negative exactly once; the dL/ds_j contribution travels back with a
// second shfl.bfly and is summed in the lane that owns j.  No shared memory, no block barriers, no dense
// contraction: per pair 3 MUFU (ex2, lg2, rcp) + ~14 FP32/INT ops, i.e. SFU-bound.
#include "common.cuh"

namespace rn {

constexpr u32 kTargetUnits = 12288;     // work-list granularity target (units of <= C micro-tiles)
constexpr u32 kMaxUnitC = 64;

struct PairParams {
  int64_t B; int K; int gbits;
  const float* logits; const float* labels; const float* rw_pos; const float* rw_neg;
  float c_log2;          // factor * log2(e)
  float factor, power; int reduce_mean; int dyn_count;
  int part_rank, part_count;
  float* loss; float* n_pair_f32; int64_t* n_pair; float* dlogits; int64_t* row_pairs;
};

// ---- heads: group / label-level starts, gather into sorted order, per-I-block J ranges, work list ------
__device__ __forceinline__ u32 lower_bound_gid(const u64* __restrict__ k, u32 n, u32 gid) {
  u32 lo = 0, hi = n;            // first p with (k[p] >> 32) >= gid
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if ((u32)(k[mid] >> 32) < gid) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ u32 lower_bound_key(const u64* __restrict__ k, u32 n, u64 key) {
  u32 lo = 0, hi = n;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (k[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

// inclusive max-scan over the 1024 threads of the block of two values at once
__device__ __forceinline__ void block_maxscan2(u32& x, u32& y, u32 (*sm)[2]) {
  const u32 ln = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 tx = __shfl_up_sync(0xFFFFFFFFu, x, o), ty = __shfl_up_sync(0xFFFFFFFFu, y, o);
    if (ln >= (u32)o) { x = max(x, tx); y = max(y, ty); }
  }
  if (ln == 31) { sm[w][0] = x; sm[w][1] = y; }
  __syncthreads();
  u32 cx = 0, cy = 0;
  for (u32 k = 0; k < w; ++k) { cx = max(cx, sm[k][0]); cy = max(cy, sm[k][1]); }
  x = max(x, cx); y = max(y, cy);
}

__global__ void __launch_bounds__(1024) k_heads(PairParams P, const u64* keyA, const u64* keyB, const u32* valA,
                                                const u32* valB, uint2* __restrict__ aj, float* __restrict__ ss,
                                                float* __restrict__ sy, float* __restrict__ swp,
                                                float* __restrict__ swn, uint2* __restrict__ blk,
                                                u32* __restrict__ ustart, u32 nblk, Ctl* ctl, int use_label) {
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, P.gbits, use_label != 0);
  const u64* __restrict__ key = (pl.npass & 1) ? keyB : keyA;
  const u32* __restrict__ val = (pl.npass & 1) ? valB : valA;
  __shared__ u32 sm[32][2];
  __shared__ u32 carry[2];
  __shared__ u32 s_last;
  const u32 B = (u32)P.B;
  const u32 t0 = blockIdx.x * 1024u;
  const u32 p = t0 + threadIdx.x;
  const bool in = p < B;
  u64 k = in ? key[p] : ~0ull;
  u64 kp = (in && p > 0) ? key[p - 1] : ~k;
  bool head = in && (p == 0 || (u32)(k >> 32) != (u32)(kp >> 32));
  bool lvl = in && (head || k != kp);
  if (threadIdx.x == 0) carry[0] = lower_bound_gid(key, t0 < B ? t0 + 1 : B, (u32)(k >> 32));
  if (threadIdx.x == 32) { u64 k0 = key[t0 < B ? t0 : B - 1]; carry[1] = lower_bound_key(key, t0 < B ? t0 + 1 : B, k0); }
  u32 xa = head ? p + 1 : 0, xl = lvl ? p + 1 : 0;
  block_maxscan2(xa, xl, sm);
  __syncthreads();
  u32 a = xa ? xa - 1 : carry[0];
  u32 l = xl ? xl - 1 : carry[1];
  u32 n = l - a;
  u32 row = in ? val[p] : 0;
  float wp = 1.f, wn = 1.f;
  if (in) {
    if (P.rw_pos) { wp = P.rw_pos[row]; if (!(wp > 0.f)) n = 0; }
    if (P.rw_neg) wn = P.rw_neg[row];
    aj[p] = make_uint2(a, n);
    ss[p] = P.logits[row];
    sy[p] = P.labels[row];
    if (P.rw_pos) swp[p] = wp;
    if (P.rw_neg) swn[p] = wn;
  } else {
    n = 0;
  }
  // J range needed by this I-block (warp)
  u32 jlo = warp_min(n ? a : 0xFFFFFFFFu), jhi = warp_max(n ? a + n : 0u);
  if (lane_id() == 0 && (p >> 5) < nblk) blk[p >> 5] = make_uint2(jlo, jhi);
  // ---- last block builds the work list ---------------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->heads_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  volatile uint2* vb = blk;
  auto ntile = [&](u32 b) -> u32 {
    u32 lo = vb[b].x, hi = vb[b].y;
    return hi > lo ? ((hi + 31) >> 5) - (lo >> 5) : 0u;
  };
  __shared__ u64 s_red[32];
  __shared__ u32 s_scan[32];
  __shared__ u32 s_carry;
  u64 m = 0;
  for (u32 b = threadIdx.x; b < nblk; b += 1024) m += ntile(b);
  m = warp_sum(m);
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) { u64 v = s_red[threadIdx.x]; v = warp_sum(v); if (threadIdx.x == 0) s_red[0] = v; }
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const u64 M = s_red[0];
  u32 C = (u32)((M + kTargetUnits - 1) / kTargetUnits);
  C = C < 1 ? 1 : (C > kMaxUnitC ? kMaxUnitC : C);
  for (u32 b0 = 0; b0 < nblk; b0 += 1024) {
    u32 b = b0 + threadIdx.x;
    u32 v = b < nblk ? (ntile(b) + C - 1) / C : 0u;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane_id() >= (u32)o) inc += t; }
    if (lane_id() == 31) s_scan[threadIdx.x >> 5] = inc;
    __syncthreads();
    u32 off = s_carry;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) off += s_scan[w];
    if (b < nblk) ustart[b] = off + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) { ustart[nblk] = s_carry; ctl->n_units = s_carry; ctl->unit_c = C; ctl->n_tiles = M; }
}

// ---- the pair kernel ------------------------------------------------------------------------------
enum { M_HASW = 1, M_DIFF = 2, M_RWN = 4, M_WRONG = 8 };

template <int MODE, bool FULL>
__device__ __forceinline__ void tile32(const float si, const float yi, const float wpi, const u32 lo, const u32 hi,
                                       const u32 pjm, const float sjm, const float yjm, const float wnjm,
                                       const float c, float& li, float& gi, u32& cnt, float& accj) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, WRONG = MODE & M_WRONG;
  float gi_t = 0.f, li_t = 0.f;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const float sj = __shfl_xor_sync(0xFFFFFFFFu, sjm, t);
    const float x = si - sj;                           // PW:117 (float32 subtract, as the reference)
    const float xs = x * c;                            // (x * factor) in log2 units
    const float e = mufu_ex2(-fabsf(xs));              // exp(-|x|)
    const float t1 = 1.0f + e;
    const float L = mufu_lg2(t1);                      // log1p(exp(-|x|)) / ln2
    const float r = mufu_rcp(t1);
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
  li += li_t; gi += gi_t;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_pair(PairParams P, const uint2* __restrict__ aj,
                                              const float* __restrict__ ss, const float* __restrict__ sy,
                                              const float* __restrict__ swp, const float* __restrict__ swn,
                                              const uint2* __restrict__ blk, const u32* __restrict__ ustart,
                                              u32 nblk, float* gacc, float* lossrow, u32* cntrow, Ctl* ctl) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, WRONG = MODE & M_WRONG;
  const u32 ln = lane_id();
  const u32 U = ctl->n_units, C = ctl->unit_c;
  const u32 B = (u32)P.B;
  const u32 u_begin = (u32)(((u64)U * (u32)P.part_rank) / (u32)P.part_count);
  const u32 u_end = (u32)(((u64)U * ((u32)P.part_rank + 1)) / (u32)P.part_count);
  const float c = P.c_log2;
  for (;;) {
    u32 u = 0;
    if (ln == 0) u = atomicAdd(&ctl->k2_ticket, 1u) + u_begin;
    u = __shfl_sync(0xFFFFFFFFu, u, 0);
    if (u >= u_end) break;
    // unit -> (I-block b, chunk): largest b with ustart[b] <= u
    u32 lo_b = 0, hi_b = nblk;
    while (hi_b - lo_b > 1) { u32 mid = (lo_b + hi_b) >> 1; if (ustart[mid] <= u) lo_b = mid; else hi_b = mid; }
    const u32 b = lo_b;
    const uint2 bj = blk[b];
    const u32 chunk = u - ustart[b];
    const u32 jb0 = (bj.x >> 5) + chunk * C;
    u32 jb1 = (bj.y + 31) >> 5; if (jb1 > jb0 + C) jb1 = jb0 + C;
    // positive side: one row per lane
    const u32 pi = b * 32 + ln;
    uint2 an = make_uint2(0, 0); float si = 0.f, yi = 0.f, wpi = 1.f;
    if (pi < B) { an = aj[pi]; si = ss[pi]; if (DIFF) yi = sy[pi]; if (HASW && swp) wpi = swp[pi]; }
    const u32 lo = an.x, hi = an.x + an.y;
    float li = 0.f, gi = 0.f; u32 cnt = 0;
    u32 pjm = jb0 * 32 + ln;
    float sjm = pjm < B ? ss[pjm] : 0.f, yjm = 0.f, wnjm = 1.f;
    if (DIFF) yjm = pjm < B ? sy[pjm] : 0.f;
    if (RWN) wnjm = pjm < B ? swn[pjm] : 0.f;
    for (u32 jb = jb0; jb < jb1; ++jb) {
      // prefetch the next J-block while this one is being scored
      const u32 pjn = pjm + 32;
      const bool more = (jb + 1 < jb1) && pjn < B;
      float sjn = more ? ss[pjn] : 0.f, yjn = 0.f, wnjn = 1.f;
      if (DIFF) yjn = more ? sy[pjn] : 0.f;
      if (RWN) wnjn = more ? swn[pjn] : 0.f;
      float accj = 0.f;
      const u32 j0 = jb * 32;
      const bool full = __all_sync(0xFFFFFFFFu, lo <= j0 && j0 + 32 <= hi);
      if (full) tile32<MODE, true>(si, yi, wpi, lo, hi, pjm, sjm, yjm, wnjm, c, li, gi, cnt, accj);
      else      tile32<MODE, false>(si, yi, wpi, lo, hi, pjm, sjm, yjm, wnjm, c, li, gi, cnt, accj);
      if (accj != 0.f) atomicAdd(gacc + pjm, accj);
      pjm = pjn; sjm = sjn; yjm = yjn; wnjm = wnjn;
    }
    if (pi < B && an.y) {
      if (gi != 0.f) atomicAdd(gacc + pi, -gi);
      if (li != 0.f) atomicAdd(lossrow + pi, li);
      if ((WRONG || RWN) && cnt) atomicAdd(cntrow + pi, cnt);
    }
  }
}

// ---- finalisation ---------------------------------------------------------------------------------
// F_a: exact counts.  c_row = pairs with the row on the positive side; c_h accumulates per PRIMARY key
// (pairwise_loss_from_batch.py:286-289); n = sum.
__global__ void __launch_bounds__(256) k_fin_counts(PairParams P, const u32* valA, const u32* valB,
                                                    const uint2* __restrict__ aj, const u32* __restrict__ cntrow,
                                                    const u32* __restrict__ slotp, u64* cprim, Ctl* ctl,
                                                    int use_label) {
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, P.gbits, use_label != 0);
  const u32* __restrict__ val = (pl.npass & 1) ? valB : valA;
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  const bool in = p < (u32)P.B;
  u32 c = 0, ps = kEmpty, row = 0;
  if (in) {
    row = val[p];
    c = P.dyn_count ? cntrow[p] : aj[p].y;
    if (c) ps = slotp[row];
    if (P.row_pairs) P.row_pairs[row] = (int64_t)c;
  }
  const u32 m = __match_any_sync(0xFFFFFFFFu, ps);
  const u32 tot = __reduce_add_sync(m, c);
  if (ps != kEmpty && lane_id() == (u32)(__ffs(m) - 1)) atomicAdd(cprim + ps, (u64)tot);
  __shared__ u64 red[8];
  u64 s = warp_sum((u64)c);
  if (lane_id() == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 t = 0;
    for (int k = 0; k < 8; ++k) t += red[k];
    if (t) atomicAdd(&ctl->n_pair, t);
  }
}

// F_b: scale, apply the occurrence weight, un-permute the gradient, reduce the loss.
__global__ void __launch_bounds__(256) k_fin_scale(PairParams P, const u32* valA, const u32* valB,
                                                   const float* __restrict__ gacc, const float* __restrict__ lossrow,
                                                   const u32* __restrict__ slotp, const u64* __restrict__ cprim,
                                                   Ctl* ctl, int use_label) {
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, P.gbits, use_label != 0);
  const u32* __restrict__ val = (pl.npass & 1) ? valB : valA;
  const u64 n = ctl->n_pair;
  const float denom = P.reduce_mean ? ((float)n + 1.0e-10f) : 1.0f;       // PW:125-126, PW:13
  const float gscale = P.factor / denom;
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  double lp = 0.0;
  if (p < (u32)P.B) {
    const u32 row = val[p];
    const float g = gacc[p], l = lossrow[p];
    float wocc = 1.f;
    if (P.power != 0.f && (g != 0.f || l != 0.f)) {
      const u32 ps = slotp[row];
      const u64 ch = ps != kEmpty ? cprim[ps] : 0ull;
      wocc = ch ? ((P.power == 1.0f) ? (float)ch : powf((float)ch, P.power)) : 0.f;   // PW:147-149
    }
    P.dlogits[row] = g * wocc * gscale;
    lp = (double)l * (double)wocc;
  }
  __shared__ double red[8];
  __shared__ u32 s_last;
  lp = warp_sum(lp);
  if (lane_id() == 0) red[threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < 8; ++k) t += red[k];
    if (t != 0.0) atomicAdd(&ctl->loss_sum, t);
    __threadfence();
    s_last = (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (s_last) {
      __threadfence();
      const double tot = *((volatile double*)&ctl->loss_sum) * 0.6931471805599453;
      *P.loss = (float)(tot / (double)denom);
      *P.n_pair_f32 = (float)n;                  // PW:276
      *P.n_pair = (int64_t)n;
    }
  }
}

template <int MODE>
static cudaError_t launch_pair(const PairParams& P, const Layout& L, char* base, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pair<MODE>, 256, 0);
    if (e != cudaSuccess) return e;
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  k_pair<MODE><<<sms * blocks_per_sm, 256, 0, st>>>(P, at<uint2>(base, L.aj), at<float>(base, L.ss),
                                                    at<float>(base, L.sy), P.rw_pos ? at<float>(base, L.swp) : nullptr,
                                                    at<float>(base, L.swn), at<uint2>(base, L.blk),
                                                    at<u32>(base, L.ustart), L.nblk, at<float>(base, L.gacc),
                                                    at<float>(base, L.lossrow), at<u32>(base, L.cnt),
                                                    at<Ctl>(base, L.ctl));
  return cudaGetLastError();
}

static cudaError_t dispatch_pair(int mode, const PairParams& P, const Layout& L, char* base, cudaStream_t st) {
  switch (mode) {
#define RN_CASE(m) case m: return launch_pair<m>(P, L, base, st);
    RN_CASE(0) RN_CASE(M_WRONG)
    RN_CASE(M_HASW) RN_CASE(M_HASW | M_WRONG)
    RN_CASE(M_HASW | M_DIFF) RN_CASE(M_HASW | M_DIFF | M_WRONG)
    RN_CASE(M_HASW | M_RWN) RN_CASE(M_HASW | M_RWN | M_WRONG)
    RN_CASE(M_HASW | M_DIFF | M_RWN) RN_CASE(M_HASW | M_DIFF | M_RWN | M_WRONG)
#undef RN_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace rn

using namespace rn;

// ---- pair-kernel timing (measurement aid, see recnow_b200.h) ------------------------------------------
static struct { bool on = false; int cap = 0; int n = 0; cudaEvent_t* ev = nullptr; } g_prof;

extern "C" int rn_profile_disable(void) {
  if (g_prof.ev) { for (int i = 0; i < 2 * g_prof.cap; ++i) cudaEventDestroy(g_prof.ev[i]); delete[] g_prof.ev; }
  g_prof.ev = nullptr; g_prof.on = false; g_prof.cap = g_prof.n = 0;
  return RN_OK;
}
extern "C" int rn_profile_enable(int32_t max_calls) {
  rn_profile_disable();
  if (max_calls <= 0 || max_calls > (1 << 20)) return RN_ERR_ARG;
  g_prof.ev = new cudaEvent_t[2 * max_calls];
  for (int i = 0; i < 2 * max_calls; ++i)
    if (cudaEventCreate(&g_prof.ev[i]) != cudaSuccess) return RN_ERR_LAUNCH;
  g_prof.cap = max_calls; g_prof.n = 0; g_prof.on = true;
  return RN_OK;
}
extern "C" int rn_profile_collect(float* ms_out_host, int32_t capacity, int32_t* n_out_host) {
  if (!ms_out_host || !n_out_host) return RN_ERR_ARG;
  int n = g_prof.n < capacity ? g_prof.n : capacity;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_prof.ev[2 * i + 1]) != cudaSuccess) return RN_ERR_LAUNCH;
    if (cudaEventElapsedTime(&ms_out_host[i], g_prof.ev[2 * i], g_prof.ev[2 * i + 1]) != cudaSuccess) return RN_ERR_LAUNCH;
  }
  *n_out_host = n; g_prof.n = 0;
  return RN_OK;
}

extern "C" size_t rn_pairwise_scratch_bytes(int64_t B, int32_t K) {
  if (B <= 0 || K <= 0) return 0;
  return make_layout(B, K).total;
}

extern "C" int rn_pairwise_launch_count(int64_t B, int32_t K) {
  if (B <= 0 || K <= 0) return 0;
  return seg_launch_count(make_layout(B, K)) + 4;
}

static int validate_pairwise(const rn_pairwise_args* a) {
  if (!a || a->B <= 0 || a->B > (1ll << 28) || a->K <= 0 || a->K > 8) return RN_ERR_ARG;
  if (!a->keys || !a->logits || !a->labels || !a->loss || !a->n_pair_f32 || !a->n_pair || !a->dlogits) return RN_ERR_ARG;
  if (a->label_func != RN_LABEL_STEP && a->label_func != RN_LABEL_DIFF) return RN_ERR_UNSUPPORTED;
  if (a->part_count < 1 || a->part_rank < 0 || a->part_rank >= a->part_count) return RN_ERR_ARG;
  const void* ptrs[] = {a->keys, a->logits, a->labels, a->row_ok, a->rw_pos, a->rw_neg, a->dlogits, a->row_pairs};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  return RN_OK;
}

extern "C" int rn_pairwise_fwd_bwd(const rn_pairwise_args* a, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = validate_pairwise(a);
  if (rc) return rc;
  const bool dyn = a->only_wrong || a->rw_neg;
  // partial (multi-GPU) evaluation needs globally consistent counts: position arithmetic gives them, the
  // score- / weight-dependent filters do not (they would need an all-reduce between counting and weighting)
  if (a->part_count > 1 && dyn) return RN_ERR_UNSUPPORTED;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  const Layout L = make_layout(a->B, a->K);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  SegInputs in{a->B, a->K, a->keys, a->labels, a->row_ok, true, true};
  if (seg_run(L, scratch, in, st) != cudaSuccess) return RN_ERR_LAUNCH;
  PairParams P{};
  P.B = a->B; P.K = a->K; P.gbits = L.gbits;
  P.logits = a->logits; P.labels = a->labels; P.rw_pos = a->rw_pos; P.rw_neg = a->rw_neg;
  P.factor = a->factor; P.power = a->power; P.reduce_mean = a->reduce_mean; P.dyn_count = dyn ? 1 : 0;
  P.c_log2 = a->factor * 1.4426950408889634f;
  P.part_rank = a->part_rank; P.part_count = a->part_count;
  P.loss = a->loss; P.n_pair_f32 = a->n_pair_f32; P.n_pair = a->n_pair; P.dlogits = a->dlogits; P.row_pairs = a->row_pairs;
  Ctl* ctl = at<Ctl>(base, L.ctl);
  const u32 nblk1024 = (u32)((a->B + 1023) / 1024);
  k_heads<<<nblk1024, 1024, 0, st>>>(P, at<u64>(base, L.keyA), at<u64>(base, L.keyB), at<u32>(base, L.valA),
                                     at<u32>(base, L.valB), at<uint2>(base, L.aj), at<float>(base, L.ss),
                                     at<float>(base, L.sy), at<float>(base, L.swp), at<float>(base, L.swn),
                                     at<uint2>(base, L.blk), at<u32>(base, L.ustart), L.nblk, ctl, 1);
  int mode = 0;
  if (a->label_func == RN_LABEL_DIFF || a->rw_pos || a->rw_neg) mode |= M_HASW;
  if (a->label_func == RN_LABEL_DIFF) mode |= M_DIFF;
  if (a->rw_neg) mode |= M_RWN;
  if (a->only_wrong) mode |= M_WRONG;
  const bool prof = g_prof.on && g_prof.n < g_prof.cap;
  if (prof) cudaEventRecord(g_prof.ev[2 * g_prof.n], st);
  if (dispatch_pair(mode, P, L, base, st) != cudaSuccess) return RN_ERR_LAUNCH;
  if (prof) { cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], st); ++g_prof.n; }
  const u32* slotp = (a->K > 1) ? at<u32>(base, L.slot1) : at<u32>(base, L.slot);
  const u32 g256 = (u32)((a->B + 255) / 256);
  k_fin_counts<<<g256, 256, 0, st>>>(P, at<u32>(base, L.valA), at<u32>(base, L.valB), at<uint2>(base, L.aj),
                                     at<u32>(base, L.cnt), slotp, at<u64>(base, L.cprim), ctl, 1);
  k_fin_scale<<<g256, 256, 0, st>>>(P, at<u32>(base, L.valA), at<u32>(base, L.valB), at<float>(base, L.gacc),
                                    at<float>(base, L.lossrow), slotp, at<u64>(base, L.cprim), ctl, 1);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
