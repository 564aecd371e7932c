// Small batches (B <= 1024): the whole pairwise call in ONE ordinary launch of ONE CTA, everything in shared memory.
//
// Replaces the same reference code as the general path (pairwise_loss_from_batch.py:228-279 with bpr_loss_func :96-127,
// the masks :16-74, 154-203, the occurrence weights :282-291) for the batch sizes the reference's own tests and its
// CPU-runnable configuration use (BASELINE.json configs[0]: B = 1024, 64 groups).  At that size the general path is pure
// latency -- two cooperative kernels, two grid barriers in each, ~30 us for 3 000 pairs; here there is no grid barrier,
// no arena traffic, no second launch, no floating-point atomic (the result is bit-identical from run to run):
//   load      one row per thread (key, score, label, weights, mask)
//   group     tile-local hashing exactly as the count phase of the counting path: the first thread to claim a key's cell
//             represents the group
//   place     counting sort in shared memory on (representative, label level) -- integer labels -1 .. 6, the menu of the
//             counting path: one atomicAdd per row hands out its rank, a 1024-wide scan of the group totals the bases.
//             Labels outside the menu: bitonic sort of (representative, order-preserving label bits, row), heads by search
//   pairs     rows of a group are contiguous, ascending label levels: the negatives of position p are [a, l) (group start,
//             level start), its positives [le, ge) (level end, group end).  The positions that have negatives are
//             compacted and walked by L lanes each (loss, own dL/ds, exact count; shuffle reduction), then the positions
//             that have positives (the dL/ds received as a negative): each pair is evaluated twice instead of once plus
//             an atomic, in a fixed order
//   finish    per-group pair totals -> occurrence weights, 1/n, gradient by original row, loss in float64
// One SM issues ~4 warp instructions per clock, so the budget of a 5 us kernel is ~1 200 instructions per row: the code
// below is written for instruction count (no unrolled sort network, no per-row searches on the usual path).
#include "pair_tiles.cuh"
#include "group_count.cuh"

namespace rn {

constexpr int kSmallRows = 1024;
constexpr int kSmallThreads = 1024;

struct SmallArgs {
  u32 B; int mode;                 // M_* bits of pair_tiles.cuh
  const int64_t* keys; const float *logits, *labels, *rw_pos, *rw_neg; const uint8_t* row_ok;
  float c, factor, power, margin; double loss_unit; int reduce_mean, hinge, gain2;
  float focal_w, focal_alpha, focal_gamma; int focal_stop;
  float* loss; float* n_pair_f32; int64_t* n_pair; float* dlogits; int64_t* row_pairs;
  Ctl* ctl; int persistent;
  const float* lut;                // gain2 == 2 (RN_LABEL_LUT): the 8 x 8 level weight table, W = lut[l_i][l_j] for l_i > l_j
  u64* dbg;                        // RN_SMALL_DEBUG: per-warp phase stamps [phase][32] (developer aid), else nullptr
};

struct SmallSmem {
  union {
    u32 cnt[kSmallRows][kLevels];                                  // counting sort: rows per (representative, level) -> level starts
    struct { u64 xch[2][kSmallRows]; u64 skey[kSmallRows]; } srt;   // fallback: exchange buffers and the sorted keys
  } u;
  u64 key[kSmallRows];             // canonical keys by row
  u32 tab[2 * kSmallRows];         // grouping hash table (representative row of the key hashed here)
  u32 gend[kSmallRows];            // end of the group of representative r (counting sort)
  float ss[kSmallRows], sy[kSmallRows], swp[kSmallRows], swn[kSmallRows];   // by sorted position
  u32 srow[kSmallRows];            // original row at the position
  uint4 sext[kSmallRows];          // (a, l, le, ge) of the position
  u32 sgrp[kSmallRows];            // representative of the position's group (kSmallRows: cannot pair)
  u32 gc[kSmallRows + 1];          // kept pairs per group (by representative)
  u32 lpos[kSmallRows], lneg[kSmallRows];                  // positions that have negatives / positives to walk
  float r_li[kSmallRows], r_gi[kSmallRows], r_gn[kSmallRows]; u32 r_cnt[kSmallRows];   // per position: results of the walks
  double red_d[2][32]; u32 red_u[32]; u32 scan[34];
  double tot_d[2]; u32 tot_u; u32 bad; u32 trash;
  float lut[64]; u32 lutbad;       // RN_LABEL_LUT: the level weight table (entries with l_i <= l_j zeroed); an entry was not usable
};

__device__ __forceinline__ u32 small_lower_bound(const u64* a, u32 n, u64 key) {   // first idx with a[idx] >= key
  u32 lo = 0, hi = n;
  while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

// sigma-like factor of one pair (d loss / d x = -d) and, if WITH_LOSS, its loss in the kernel's unit
template <bool WITH_LOSS>
__device__ __forceinline__ float small_pair(const float x, const float c, const int hinge, const float margin, float& lo2) {
  if (hinge) {
    const float hv = fmaf(-x, c, margin);                                   // max(0, margin - x), as tile_hinge
    if (WITH_LOSS) lo2 = fmaxf(hv, 0.f);
    return hv > 0.f ? 1.0f : 0.f;
  }
  const float xs = x * c;                                                   // (x * factor) in log2 units
  const float e = mufu_ex2(-fabsf(xs));
  const float t1p = 1.0f + e;
  if (WITH_LOSS) lo2 = fmaxf(-xs, 0.f) + mufu_lg2(t1p);                     // softplus(-x) / ln2   (PW:120-121)
  return (xs >= 0.f ? e : 1.0f) * mufu_rcp(t1p);                            // sigma(-x)
}

// SIMPLE: the default call (no weights of any kind, no wrong-order filter, logistic loss) -- a second instantiation whose
// pair walks carry none of the option tests (the kernel is bound by the issue slots of one SM).
template <bool SIMPLE>
__global__ void __launch_bounds__(kSmallThreads, 1) k_small(SmallArgs A) {
  extern __shared__ __align__(16) unsigned char small_raw[];
  SmallSmem& S = *reinterpret_cast<SmallSmem*>(small_raw);
  const u32 tid = threadIdx.x, ln = tid & 31u, w = tid >> 5, B = A.B;
  const bool HASW = !SIMPLE && (A.mode & M_HASW), DIFF = !SIMPLE && (A.mode & M_DIFF), RWN = !SIMPLE && (A.mode & M_RWN),
             WRONG = !SIMPLE && (A.mode & M_WRONG);
  const int hinge = SIMPLE ? 0 : A.hinge;
  const bool in = tid < B;
  auto stamp_s = [&](int i) {                                                     // (phase stamps, as the general path's)
    if (tid == 0) A.ctl->ts[i] = globaltimer();
    if (A.dbg && ln == 0) A.dbg[i * 32 + w] = globaltimer();
  };
  stamp_s(0);
  // ---- load ----------------------------------------------------------------------------------------------------
  const u64 key = in ? (u64)A.keys[tid] : 0ull;
  const float s_row = in ? A.logits[tid] : 0.f, y_row = in ? A.labels[tid] : 0.f;
  const float wp_row = (in && A.rw_pos) ? A.rw_pos[tid] : 1.f, wn_row = (in && A.rw_neg) ? A.rw_neg[tid] : 1.f;
  const bool ok = in && (A.row_ok ? A.row_ok[tid] != 0 : true) && !(y_row != y_row);       // (a NaN label pairs with nothing)
  S.key[tid] = key;
  S.tab[tid] = kEmpty; S.tab[tid + kSmallRows] = kEmpty;
  *reinterpret_cast<uint4*>(&S.u.cnt[tid][0]) = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<uint4*>(&S.u.cnt[tid][4]) = make_uint4(0, 0, 0, 0);
  S.gc[tid] = 0;
  if (tid == 0) { S.gc[kSmallRows] = 0; S.bad = 0; S.trash = 0; S.lutbad = 0; }
  // RN_LABEL_LUT: the table, checked as in k_pair (entries with l_i > l_j finite and > 0, the others zeroed)
  const bool lutmode = !SIMPLE && A.gain2 == 2;
  bool lut_entry_bad = false;
  if (lutmode && tid < 64) {
    float v = A.lut[tid];
    if ((tid >> 3) > (tid & 7u)) { if (!(v > 0.f) || v > 3.0e38f) { v = 0.f; lut_entry_bad = true; } }
    else v = 0.f;
    S.lut[tid] = v;
  }
  double fsum = 0.0;
  if (A.focal_w != 0.f && in) fsum = (double)focal_row(s_row, y_row, A.focal_alpha, A.focal_gamma, A.focal_stop).x;
  float yv_row = (y_row != y_row) ? 0.f : (A.gain2 == 1 ? exp2f(y_row) : y_row);          // the label the weights see
  __syncthreads();
  if (lut_entry_bad) S.lutbad = 1;
  // ---- group: the first thread to claim a key's cell represents it --------------------------------------------------
  u32 rep = kSmallRows;                            // rows that cannot pair: one group behind all others
  int lev = 0;
  if (ok) {
    const u64 h = mix64(0x9E3779B97F4A7C15ull ^ key);
    u32 ls = (u32)(h >> 40) & (2 * kSmallRows - 1);
    for (;;) {
      u32 cur = S.tab[ls];
      if (cur == kEmpty) {
        const u32 prev = atomicCAS(&S.tab[ls], kEmpty, tid);
        if (prev == kEmpty) { rep = tid; break; }
        cur = prev;
      }
      if (S.key[cur] == key) { rep = cur; break; }
      ls = (ls + 1) & (2 * kSmallRows - 1);
    }
    if (!label_level(y_row, lev)) { S.bad = 1; lev = 0; }
  }
  if (lutmode) yv_row = (float)lev;                // (the weights see the label LEVEL; a label off the menu fails the call below)
  // label part of a pair weight under M_DIFF (as label_weight of pair_tiles.cuh)
  auto lw = [&](const float yi, const float yj) -> float { return lutmode ? S.lut[(int)yi * 8 + (int)yj] : yi - yj; };
  // rank inside (group, level): the counting sort's only atomic
  u32 rank = 0;
  if (ok) rank = atomicAdd(&S.u.cnt[rep][lev], 1u);
  else if (in) rank = atomicAdd(&S.trash, 1u);
  __syncthreads();
  stamp_s(1);
  u32 pos = tid;                                   // sorted position of this thread's row
  if (!S.bad) {
    // ---- place: level starts of every group from a scan over the representatives' totals ------------------------------
    uint4 c0 = *reinterpret_cast<uint4*>(&S.u.cnt[tid][0]), c1 = *reinterpret_cast<uint4*>(&S.u.cnt[tid][4]);
    const u32 tot = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
    u32 inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
    if (ln == 31) S.scan[w] = inc;
    __syncthreads();
    if (w == 0) {
      const u32 x = S.scan[ln];
      u32 xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y2 = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (ln >= (u32)o) xi += y2; }
      S.scan[ln] = xi - x;
      if (ln == 31) S.scan[32] = xi;               // rows that can pair: the others sit behind them
    }
    __syncthreads();
    u32 run = S.scan[w] + inc - tot;               // base of the group this thread represents (if any)
    uint4 p0, p1;
    p0.x = run; run += c0.x; p0.y = run; run += c0.y; p0.z = run; run += c0.z; p0.w = run; run += c0.w;
    p1.x = run; run += c1.x; p1.y = run; run += c1.y; p1.z = run; run += c1.z; p1.w = run; run += c1.w;
    *reinterpret_cast<uint4*>(&S.u.cnt[tid][0]) = p0; *reinterpret_cast<uint4*>(&S.u.cnt[tid][4]) = p1;
    S.gend[tid] = run;
    __syncthreads();
    u32 a = 0, l = 0, le = 0, ge = 0;
    if (ok) {
      a = S.u.cnt[rep][0]; l = S.u.cnt[rep][lev]; ge = S.gend[rep];
      le = lev + 1 < kLevels ? S.u.cnt[rep][lev + 1] : ge;
      pos = l + rank;
    } else if (in) pos = S.scan[32] + rank;
    if (in) S.sext[pos] = make_uint4(a, l, le, ge);
  } else {
    // ---- labels outside the level menu: bitonic sort of (representative, label bits, row), heads by binary search ---------
    __syncthreads();                               // (everybody has read S.bad and is done with the counters)
    u64 v = ~0ull;                                 // (threads beyond B: behind everything)
    if (in) v = ((u64)rep << 43) | ((u64)(ok ? enc_label(y_row) : 0u) << 11) | (u64)tid;
    int buf = 0;
#pragma unroll 1
    for (u32 k = 2; k <= (u32)kSmallThreads; k <<= 1) {
#pragma unroll 1
      for (u32 j = k >> 1; j > 0; j >>= 1) {
        u64 o;
        if (j >= 32) {
          S.u.srt.xch[buf][tid] = v;
          __syncthreads();
          o = S.u.srt.xch[buf][tid ^ j];
          buf ^= 1;
        } else {
          o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
        }
        const bool keep_min = ((tid & j) == 0) == ((tid & k) == 0);
        v = keep_min ? (v < o ? v : o) : (v < o ? o : v);
      }
    }
    S.u.srt.skey[tid] = v;                         // position tid holds row (v & 0x7FF): tell that row where it went
    __syncthreads();
    if (tid < B) {
      const u32 grp = (u32)(v >> 43);
      u32 a = 0, l = 0, le = 0, ge = 0;
      if (grp != (u32)kSmallRows) {
        a = small_lower_bound(S.u.srt.skey, B, (u64)grp << 43);
        ge = small_lower_bound(S.u.srt.skey, B, (u64)(grp + 1u) << 43);
        l = small_lower_bound(S.u.srt.skey, B, v & ~0x7FFull);
        le = small_lower_bound(S.u.srt.skey, B, (v | 0x7FFull) + 1ull);
      }
      S.sext[tid] = make_uint4(a, l, le, ge);
      S.tab[(u32)v & 0x7FFu] = tid;                // (the hash table is free now: row -> position)
    }
    __syncthreads();
    if (in) pos = S.tab[tid];
  }
  // scatter the row into its position
  if (in) {
    S.ss[pos] = s_row; S.sy[pos] = yv_row; S.swp[pos] = wp_row; S.swn[pos] = wn_row; S.srow[pos] = tid; S.sgrp[pos] = rep;
  }
  __syncthreads();
  stamp_s(2);
  // ---- sorted position p = tid -----------------------------------------------------------------------------------------
  const bool in_p = tid < B;
  float si = 0.f, wp = 1.f;
  u32 row = 0, grp = kSmallRows;
  uint4 ext = make_uint4(0, 0, 0, 0);
  if (in_p) { si = S.ss[tid]; wp = S.swp[tid]; row = S.srow[tid]; grp = S.sgrp[tid]; ext = S.sext[tid]; }
  const bool wp_test = !SIMPLE && A.rw_pos != nullptr && !RWN;                       // PW:193 C = W > 0 with a positive-side factor only
  const float c = A.c;
  // Work per position: negatives [a, l) as the positive side, positives [le, ge) as the negative side.  Most positions
  // have none (binary labels: the 25 % positives hold all the negative ranges), so the positions that have work are
  // COMPACTED first (warps stay full, idle positions cost no issue slots), and every listed position gets L lanes
  // (a power of two, as many as the block can spare) that stride over its range and reduce with shuffles.
  const u32 wpos = (in_p && !(wp_test && !(wp > 0.f))) ? ext.y - ext.x : 0u, wneg = in_p ? ext.w - ext.z : 0u;
  {
    const u32 bp = __ballot_sync(0xFFFFFFFFu, wpos != 0u), bn = __ballot_sync(0xFFFFFFFFu, wneg != 0u);
    if (ln == 0) S.scan[w] = (u32)__popc(bp) | ((u32)__popc(bn) << 16);       // (both counts <= 1024: no carry between the halves)
    S.r_li[tid] = 0.f; S.r_gi[tid] = 0.f; S.r_gn[tid] = 0.f; S.r_cnt[tid] = 0u;
    __syncthreads();
    if (w == 0) {
      const u32 x = S.scan[ln];
      u32 xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y2 = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (ln >= (u32)o) xi += y2; }
      S.scan[ln] = xi - x;
      if (ln == 31) S.scan[32] = xi;
    }
    __syncthreads();
    const u32 base = S.scan[w];
    if (wpos) S.lpos[(base & 0xFFFFu) + (u32)__popc(bp & lanemask_lt())] = tid;
    if (wneg) S.lneg[(base >> 16) + (u32)__popc(bn & lanemask_lt())] = tid;
    __syncthreads();
  }
  stamp_s(5);
  const u32 npos = S.scan[32] & 0xFFFFu, nneg = S.scan[32] >> 16;
  auto lanes_for = [](u32 n) -> u32 { u32 L = 32; while (L > 1u && n * L > (u32)kSmallThreads) L >>= 1; return L; };
  {
    // listed position as the POSITIVE side
    const u32 L = lanes_for(npos), r = tid / L, sub = tid & (L - 1u);
    float li = 0.f, gi = 0.f; u32 cnt = 0;
    u32 p = 0;
    if (r < npos) {
      p = S.lpos[r];
      const float s_p = S.ss[p], y_p = S.sy[p], w_p = S.swp[p];
      const uint4 e = S.sext[p];
      // (four negatives per round, their chains -- shared-memory load, ex2, lg2 / rcp -- in flight together)
      for (u32 q0 = e.x + sub; q0 < e.y; q0 += 4u * L) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const u32 q = q0 + (u32)k * L;
          bool valid = q < e.y;
          const u32 qq = valid ? q : e.x;
          const float x = s_p - S.ss[qq];                                 // PW:117 (float32 subtract)
          float wv = 1.f;
          if (HASW) {
            wv = w_p;
            if (DIFF) wv = lw(y_p, S.sy[qq]) * w_p;
            if (RWN) { wv *= S.swn[qq]; valid = valid && wv > 0.f; }      // PW:193
          }
          if (WRONG) valid = valid && (x < 0.f);                          // PW:200-202
          float lo2 = 0.f;
          const float d = small_pair<true>(x, c, hinge, A.margin, lo2) * wv;
          const float l = wv * lo2;
          li += valid ? l : 0.f; gi += valid ? d : 0.f; cnt += valid ? 1u : 0u;
        }
      }
    }
    for (u32 o = L >> 1; o; o >>= 1) {
      li += __shfl_xor_sync(0xFFFFFFFFu, li, o); gi += __shfl_xor_sync(0xFFFFFFFFu, gi, o); cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    if (r < npos && sub == 0) { S.r_li[p] = li; S.r_gi[p] = gi; S.r_cnt[p] = cnt; }
  }
  stamp_s(6);
  {
    // listed position as the NEGATIVE side: the same pairs, seen from the other end (dL/ds it receives)
    const u32 L = lanes_for(nneg), r = tid / L, sub = tid & (L - 1u);
    float gn = 0.f;
    u32 p = 0;
    if (r < nneg) {
      p = S.lneg[r];
      const float s_p = S.ss[p], y_p = S.sy[p], wn_p = S.swn[p];
      const uint4 e = S.sext[p];
      for (u32 q0 = e.z + sub; q0 < e.w; q0 += 4u * L) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const u32 q = q0 + (u32)k * L;
          bool valid = q < e.w;
          const u32 qq = valid ? q : e.z;
          const float wq = S.swp[qq];
          if (wp_test) valid = valid && (wq > 0.f);
          const float x = S.ss[qq] - s_p;
          float wv = 1.f;
          if (HASW) {
            wv = wq;
            if (DIFF) wv = lw(S.sy[qq], y_p) * wq;
            if (RWN) { wv *= wn_p; valid = valid && wv > 0.f; }
          }
          if (WRONG) valid = valid && (x < 0.f);
          float unused;
          const float d = small_pair<false>(x, c, hinge, A.margin, unused) * wv;
          gn += valid ? d : 0.f;
        }
      }
    }
    for (u32 o = L >> 1; o; o >>= 1) gn += __shfl_xor_sync(0xFFFFFFFFu, gn, o);
    if (r < nneg && sub == 0) S.r_gn[p] = gn;
  }
  stamp_s(7);
  __syncthreads();
  const float li = S.r_li[tid], gi = S.r_gi[tid], gn = S.r_gn[tid];
  const u32 cnt = S.r_cnt[tid];
  stamp_s(4);
  // ---- counts -> occurrence weights, totals ------------------------------------------------------------------------------
  if (A.power != 0.f && cnt) atomicAdd(&S.gc[grp], cnt);                  // (cnt is exact in every mode: one per kept pair)
  {
    const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (ln == 0) S.red_u[w] = wsum;
  }
  __syncthreads();
  float wocc = 1.f;
  if (A.power != 0.f) {
    const u32 ch = S.gc[grp];                                             // PW:286-289: pairs of the row's (primary) group
    wocc = ch ? ((A.power == 1.0f) ? (float)ch : powf((float)ch, A.power)) : 0.f;      // PW:147-149
  }
  {
    const double lp = warp_sum((double)(li * wocc)), fp = warp_sum(fsum);
    if (ln == 0) { S.red_d[0][w] = lp; S.red_d[1][w] = fp; }
  }
  __syncthreads();
  if (w == 0) {
    const u32 nt = __reduce_add_sync(0xFFFFFFFFu, S.red_u[ln]);
    const double lt = warp_sum(S.red_d[0][ln]), ft = warp_sum(S.red_d[1][ln]);
    if (ln == 0) { S.tot_u = nt; S.tot_d[0] = lt; S.tot_d[1] = ft; }
  }
  __syncthreads();
  const u32 n = S.tot_u;
  const float denom = A.reduce_mean ? ((float)n + 1.0e-10f) : 1.0f;       // PW:125-126, PW:13
  const float gscale = A.factor / denom;
  if (in_p) {
    float g = (gn - gi) * wocc * gscale;
    if (A.focal_w != 0.f) g += A.focal_w / (float)B * focal_row(si, A.labels[row], A.focal_alpha, A.focal_gamma, A.focal_stop).y;
    A.dlogits[row] = g;
    if (A.row_pairs) A.row_pairs[row] = (int64_t)cnt;
  }
  if (tid == 0) {
    float lossv = (float)(S.tot_d[0] * A.loss_unit / (double)denom);
    if (A.focal_w != 0.f) lossv += A.focal_w * (float)(S.tot_d[1] / (double)B);
    // RN_LABEL_LUT: a label without a level or an unusable table entry fails the call (as in k_pair: never a plausible number)
    const u32 lut_err = (lutmode && (S.bad || S.lutbad)) ? 8u : 0u;
    if (lut_err) lossv = __int_as_float(0x7FC00000);
    *A.loss = lossv;
    *A.n_pair_f32 = (float)n;                    // PW:276
    *A.n_pair = (int64_t)n;
    // report for rn_last_device_error / the path query (3 = this kernel); the working fields stay as they are (clean)
    A.ctl->rep_err = lut_err; A.ctl->rep_path = 3; A.ctl->ts[23] = globaltimer();
    if (!A.persistent) { A.ctl->err = 0; A.ctl->path = 0; }
  }
}

// Host side: true if the call was taken (rc then holds its status).
bool small_pairwise(const rn_pairwise_args* a, void* scratch, cudaStream_t st, int mode, bool hinge, int* rc) {
  const char* env = getenv("RN_SMALL");             // (read per call: the tests switch between the two paths)
  const bool on = !(env && *env) || atoi(env) != 0;
  if (!on || a->B > kSmallRows || a->K != 1 || a->block_rows || a->part_count != 1 || a->deterministic || a->gather_dst)
    return false;
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *rc = RN_ERR_LAUNCH; return true; }
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute((const void*)k_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)) != cudaSuccess ||
        cudaFuncSetAttribute((const void*)k_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)) != cudaSuccess) {
      cudaGetLastError(); *rc = RN_ERR_LAUNCH; return true;
    }
    attr_set[dev] = true;
  }
  SmallArgs A{};
  A.B = (u32)a->B; A.mode = mode;
  A.keys = a->keys; A.logits = a->logits; A.labels = a->labels; A.rw_pos = a->rw_pos; A.rw_neg = a->rw_neg; A.row_ok = a->row_ok;
  A.c = hinge ? a->factor : a->factor * 1.4426950408889634f;
  A.factor = a->factor; A.power = a->power; A.margin = a->margin; A.loss_unit = hinge ? 1.0 : 0.6931471805599453;
  A.reduce_mean = a->reduce_mean; A.hinge = hinge ? 1 : 0;
  A.gain2 = a->label_func == RN_LABEL_GAIN2 ? 1 : (a->label_func == RN_LABEL_LUT ? 2 : 0); A.lut = a->weight_lut;
  A.focal_w = a->focal_weight; A.focal_alpha = a->focal_alpha; A.focal_gamma = a->focal_gamma; A.focal_stop = a->focal_stop_weight_gradient;
  A.loss = a->loss; A.n_pair_f32 = a->n_pair_f32; A.n_pair = a->n_pair; A.dlogits = a->dlogits; A.row_pairs = a->row_pairs;
  A.ctl = static_cast<Ctl*>(scratch); A.persistent = a->scratch_persistent;
  {
    const char* dv = getenv("RN_SMALL_DEBUG");
    if (dv && *dv && atoi(dv)) A.dbg = at<u64>(scratch, make_layout(a->scratch_rows ? a->scratch_rows : a->B, 1, a->B).gstat);
  }
  if (mode == 0 && !hinge && !a->rw_pos) k_small<true><<<1, kSmallThreads, sizeof(SmallSmem), st>>>(A);
  else k_small<false><<<1, kSmallThreads, sizeof(SmallSmem), st>>>(A);
  *rc = cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
  return true;
}

}  // namespace rn
