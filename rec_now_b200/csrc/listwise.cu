// K5/K6 — in-batch listwise loss: segmented log-sum-exp softmax cross-entropy, forward + backward.
//
// Replaces to_listwise_sample (listwise_loss_from_batch.py:89-148: unique_with_counts + four (G,B)
// densifications + row filters) and listwise_loss_via_softmax_cross_entropy_with_logits (LW:151-173).
// After K1 (sort by (first-occurrence id, row)) a list is a contiguous run of sorted positions and the runs
// appear in first-occurrence order, so valid-list ranks (the row order of the reference's dense outputs) are
// an exclusive scan over run heads.  HBM-bound: 20 algorithmic bytes per sample.
#include <stdlib.h>
#include "segment.cuh"

namespace rn {

struct LwParams {
  u32 B; int gbits;
  const float* list_w; float th; int do_reduce;
  float inv_t;                  // 1 / temperature: the softmax runs on logits * inv_t, the gradient is scaled back
  float* loss; float* list_loss; int32_t* n_valid; int32_t* n_group; float* dlogits;
};

// per-list record, SoA over head positions (gstat + k*B)
enum { R_MAX = 0, R_LSE = 1, R_SY = 2, R_LOSS = 3, R_VALID = 4, R_RANK = 5 };

__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// Listwise tail of k_seg: bounds -> per-list statistics -> ranks of the valid lists + weighted losses -> gradient,
// separated by grid barriers (one launch instead of five).
struct ListwiseTail {
  static constexpr bool kFast = false;
  BoundsTail bounds;            // astart / gend / perm + gathers of logits (ss) and labels (sy)
  LwParams P;
  const float* ss; const float* sy; float* rec; u32* chunkcnt;

  // One warp per 32 sorted positions; the warp processes every list whose head falls in its block.
  __device__ __forceinline__ u32 group_block(u32 blk, Ctl* ctl) const {
    const u32 ln = lane_id();
    const u32 p = blk * 32 + ln;
    const u32* astart = bounds.astart; const u32* gend = bounds.gend;
    const bool head = p < P.B && astart[p] == p;
    u32 heads = __ballot_sync(0xFFFFFFFFu, head);
    u32 ng = __popc(heads), nv = 0;
    while (heads) {
      const u32 h = __ffs(heads) - 1; heads &= heads - 1;
      const u32 a = blk * 32 + h, e = gend[a];
      float m = -INFINITY;
      for (u32 q = a + ln; q < e; q += 32) m = fmaxf(m, ss[q] * P.inv_t);
      m = warp_maxf(m);
      float z = 0.f, sumy = 0.f; int hp = 0, hn = 0;
      for (u32 q = a + ln; q < e; q += 32) {
        const float s = ss[q] * P.inv_t, y = sy[q];
        z += expf(s - m); sumy += y;
        hp |= (y > P.th); hn |= ((y - P.th) < 0.f);                // LW:135-136
      }
      z = warp_sum(z); sumy = warp_sum(sumy);
      const bool valid = __any_sync(0xFFFFFFFFu, hp) && __any_sync(0xFFFFFFFFu, hn);   // LW:137
      const float lse = logf(z);
      float dot = 0.f;
      if (valid)
        for (u32 q = a + ln; q < e; q += 32) dot += (sy[q] / sumy) * (lse - (ss[q] * P.inv_t - m));   // LW:144, LW:167
      dot = warp_sum(dot);
      if (ln == 0) {
        rec[(size_t)R_MAX * P.B + a] = m; rec[(size_t)R_LSE * P.B + a] = lse; rec[(size_t)R_SY * P.B + a] = sumy;
        rec[(size_t)R_LOSS * P.B + a] = dot; rec[(size_t)R_VALID * P.B + a] = valid ? 1.f : 0.f;
      }
      nv += valid ? 1u : 0u;
    }
    if (ln == 0 && ng) atomicAdd(&ctl->n_groups, ng);
    return nv;
  }

  __device__ __forceinline__ void run(const SegParams& S, const Plan& pl, const u64* key, const u32* val, u32* smem, u32& epoch) const {
    Ctl* ctl = S.ctl;
    const u32 B = P.B, ln = lane_id(), w = threadIdx.x >> 5;
    const u32 nchunks = (B + kSegThreads - 1) / kSegThreads;
    const u32* astart = bounds.astart;
    bounds.run(S, pl, key, val, smem, epoch);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    // ---- per-list statistics; valid lists per 512-position chunk -------------------------------------------
    u32* sm_cnt = smem;                    // [kSegWarps]
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const u32 nv = group_block(c * kSegWarps + w, ctl);
      if (ln == 0) sm_cnt[w] = nv;
      __syncthreads();
      if (threadIdx.x == 0) {
        u32 t = 0;
        for (int q = 0; q < kSegWarps; ++q) t += sm_cnt[q];
        chunkcnt[c] = t;
        if (t) atomicAdd(&ctl->n_valid, t);
      }
      __syncthreads();
    }
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    // ---- rank of every valid list (exclusive scan of the valid flags over head positions = first-occurrence
    //      order, LW:109), per-list weighted losses, their sum ------------------------------------------------
    u32* sm_scan = smem;                   // [kSegWarps]
    u32* sm_off = smem + kSegWarps;        // [1]
    double* sm_d = reinterpret_cast<double*>(smem + 32);   // [kSegWarps]
    double acc = 0.0;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
      u32 part = 0;
      for (u32 q = threadIdx.x; q < c; q += kSegThreads) part += chunkcnt[q];
      part = warp_sum(part);
      if (ln == 0) sm_scan[w] = part;
      __syncthreads();
      if (threadIdx.x == 0) { u32 t = 0; for (int q = 0; q < kSegWarps; ++q) t += sm_scan[q]; *sm_off = t; }
      __syncthreads();
      const u32 base = *sm_off;
      __syncthreads();
      const u32 p = c * kSegThreads + threadIdx.x;
      const bool v = p < B && astart[p] == p && rec[(size_t)R_VALID * B + p] != 0.f;
      u32 inc = v ? 1u : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
      if (ln == 31) sm_scan[w] = inc;
      __syncthreads();
      u32 off = base;
      for (u32 q = 0; q < w; ++q) off += sm_scan[q];
      if (v) {
        const u32 r = off + inc - 1;
        float l = rec[(size_t)R_LOSS * B + p];
        if (P.list_w) l *= P.list_w[r];                              // LW:168-169
        rec[(size_t)R_RANK * B + p] = __uint_as_float(r);
        if (P.list_loss) P.list_loss[r] = l;
        acc += (double)l;
      }
      __syncthreads();
    }
    acc = warp_sum(acc);
    if (ln == 0) sm_d[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int q = 0; q < kSegWarps; ++q) t += sm_d[q];
      if (t != 0.0) atomicAdd(&ctl->loss_sum, t);
    }
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    // ---- gradient + scalars -------------------------------------------------------------------------------
    const u32 V = ld_relaxed(&ctl->n_valid);
    const u32 gtid = blockIdx.x * kSegThreads + threadIdx.x, gthreads = gridDim.x * kSegThreads;
    // LW:172 nan_to_zero on the reduced loss: tf.cond takes the constant branch, so no gradient flows either
    const double tsum = *reinterpret_cast<volatile double*>(&ctl->loss_sum);
    const bool nanloss = P.do_reduce && V && (tsum != tsum);
    for (u32 p = gtid; p < B; p += gthreads) {
      const u32 a = astart[p];
      float g = 0.f;
      if (!nanloss && rec[(size_t)R_VALID * B + a] != 0.f) {
        const float m = rec[(size_t)R_MAX * B + a], lse = rec[(size_t)R_LSE * B + a], sumy = rec[(size_t)R_SY * B + a];
        float wr = 1.f;
        if (P.list_w) wr = P.list_w[__float_as_uint(rec[(size_t)R_RANK * B + a])];
        if (P.do_reduce) wr /= (float)V;
        g = P.inv_t * wr * (expf(ss[p] * P.inv_t - m - lse) - sy[p] / sumy);               // xent backprop: softmax - labels
      }
      P.dlogits[bounds.perm[p]] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const double t = *reinterpret_cast<volatile double*>(&ctl->loss_sum);
      if (P.do_reduce) *P.loss = (V && !nanloss) ? (float)(t / (double)V) : 0.f;   // LW:171-172 (mean; NaN -> 0)
      *P.n_valid = (int32_t)V;
      *P.n_group = (int32_t)ld_relaxed(&ctl->n_groups);
    }
  }
};


// ---- K5, counting form: the whole listwise loss in ONE cooperative kernel without a sort or a scatter ---------------
// (same reference lines: LW:107-148, 166-172).  The loss and its gradient only need PER-LIST statistics -- max logit,
// sum exp, label sum, sum y (s - max), has-positive / has-negative -- and every row's own (s, y): rows never have to
// move.  The lists live in the record table of group_count.cuh (one 128-bit compare-and-swap per distinct id and
// 512-row tile); the statistics are accumulated with atomics on the records:
//   phase 1  find / create the list's record; max logit (atomicMax on the order-preserving encoding), label sum,
//            validity flags -- aggregated per tile in shared memory first
//   phase 2  sum exp(s - max), sum y (s - max) per list; the records' creators count the valid lists (LW:135-137)
//   phase 3  gradient (softmax - p) / V written in ROW order (coalesced), per-list losses summed; the last CTA out
//            writes the scalars (mean, NaN -> 0 as LW:172) and leaves the arena clean.
// Two grid barriers.  Taken when the valid lists' RANKS are not needed (do_reduce, no per-list weights, no per-list
// output) and the arena is persistent; everything else runs the sorted path above.
struct __align__(16) LRec { u64 key; u32 rep1; u32 flags; u32 maxenc; float sumy; float Z; float dotm; u32 pad[8]; };
static_assert(sizeof(LRec) == sizeof(GRec), "the listwise records share the group record table");

__device__ __forceinline__ float dec_label(u32 e) {     // inverse of enc_label
  const u32 u = (e & 0x80000000u) ? (e ^ 0x80000000u) : ~e;
  return __uint_as_float(u);
}

struct LwCountArgs {
  u32 B; u32 capmask;
  const int64_t* keys; const uint8_t* row_ok; const float* labels; const float* logits;
  float th; float inv_t;
  float* loss; int32_t* n_valid; int32_t* n_group; float* dlogits;
  GRec* rec; u32 *glist, *gcount, *rslot; Ctl* ctl;
};

__global__ void __launch_bounds__(kSegThreads, 1) k_lw_count(LwCountArgs A) {
  constexpr u32 kLoc = 2 * kGTile;
  __shared__ __align__(16) u32 smem[kLoc + 2 * kGTile + 5 * kGTile + 64];
  u32* sm_tab = smem;                                           // [kLoc]
  u64* sm_key = reinterpret_cast<u64*>(smem + kLoc);            // [kGTile]
  u32* sm_max = smem + kLoc + 2 * kGTile;                       // [kGTile] per representative
  float* sm_sumy = reinterpret_cast<float*>(sm_max + kGTile);   // [kGTile]
  u32* sm_flags = sm_max + 2 * kGTile;                          // [kGTile]
  u32* sm_gslot = sm_max + 3 * kGTile;                          // [kGTile]
  u32* sm_misc = sm_max + 4 * kGTile;                           // [0] created, [1] singleton rows, [2] valid lists, [3] lists
  double* sm_d = reinterpret_cast<double*>(sm_max + 4 * kGTile + 16);   // [kSegWarps]
  Ctl* ctl = A.ctl;
  LRec* rec = reinterpret_cast<LRec*>(A.rec);
  const u32 B = A.B, tid = threadIdx.x, ln = lane_id(), w = tid >> 5;
  const u32 ntile = (B + kGTile - 1) / kGTile;
  const bool single = ntile <= gridDim.x;
  u32 epoch = 0;
  grid_dep_wait();
  stamp(ctl, 0);
  if (tid < 4) sm_misc[tid] = 0;              // (CTAs without rows still take part in the reductions below)
  __syncthreads();
  u32 k_slot = kEmpty; float k_s = 0.f, k_y = 0.f;
  // ---- phase 1 ----------------------------------------------------------------------------------------------------
  for (u32 t = blockIdx.x; t < ntile; t += gridDim.x) {
    const u32 i = t * kGTile + tid;
    const bool in = i < B;
    const u64 key = in ? (u64)A.keys[i] : 0ull;
    const float y = in ? A.labels[i] : 0.f, s = in ? A.logits[i] * A.inv_t : 0.f;
    const bool ok = in && (A.row_ok ? A.row_ok[i] != 0 : true);       // (0: the id was NaN / inf -- a singleton list, never valid)
    sm_tab[tid] = kEmpty; sm_tab[tid + kGTile] = kEmpty;
    sm_max[tid] = 0u; sm_sumy[tid] = 0.f; sm_flags[tid] = 0u;
    if (tid < 2) sm_misc[tid] = 0;
    sm_key[tid] = key;
    __syncthreads();
    const u64 h = mix64(0x9E3779B97F4A7C15ull ^ key);
    u32 rep = tid;
    if (ok) {
      u32 ls = (u32)(h >> 40) & (kLoc - 1);
      for (;;) {
        u32 cur = sm_tab[ls];
        if (cur == kEmpty) {
          const u32 prev = atomicCAS(&sm_tab[ls], kEmpty, tid);
          if (prev == kEmpty) { rep = tid; break; }
          cur = prev;
        }
        if (sm_key[cur] == key) { rep = cur; break; }
        ls = (ls + 1) & (kLoc - 1);
      }
      atomicMax(&sm_max[rep], enc_label(s));
      atomicAdd(&sm_sumy[rep], y);
      const u32 f = ((y > A.th) ? 1u : 0u) | (((y - A.th) < 0.f) ? 2u : 0u);          // LW:135-136
      if (f) atomicOr(&sm_flags[rep], f);
    } else if (in) {
      atomicAdd(&sm_misc[1], 1u);
    }
    bool created = false; u32 slot = 0;
    const bool isrep = ok && rep == tid;
    if (isrep) { slot = grec_insert(A.rec, A.capmask, h, key, i, created, &ctl->err); sm_gslot[tid] = slot; }
    if (created) A.glist[(size_t)t * kGTile + atomicAdd(&sm_misc[0], 1u)] = slot;
    __syncthreads();
    if (isrep) {
      atomicMax(&rec[slot].maxenc, sm_max[tid]);
      atomicAdd(&rec[slot].sumy, sm_sumy[tid]);
      if (sm_flags[tid]) atomicOr(&rec[slot].flags, sm_flags[tid]);
    }
    if (tid == 0) { A.gcount[t] = sm_misc[0]; if (sm_misc[1]) atomicAdd(&ctl->n_groups, sm_misc[1]); }
    if (in) {
      const u32 rslot = ok ? sm_gslot[rep] : kEmpty;
      if (single) { k_slot = rslot; k_s = s; k_y = y; } else A.rslot[i] = rslot;
    }
    __syncthreads();
  }
  stamp(ctl, 1);
  grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
  stamp(ctl, 2);
  // ---- phase 2: sum exp and label-weighted logit sum per list; the creators count the valid lists ----------------------
  float k_e = 0.f;
  auto accumulate = [&](u32 slot, float s, float y) -> float {
    if (slot == kEmpty) return 0.f;
    const float m = dec_label(rec[slot].maxenc);
    const float e = expf(s - m);
    atomicAdd(&rec[slot].Z, e);
    if (y != 0.f) atomicAdd(&rec[slot].dotm, y * (s - m));
    return e;
  };
  u32 nval = 0, nlist = 0;
  for (u32 t = blockIdx.x; t < ntile; t += gridDim.x) {
    const u32 i = t * kGTile + tid;
    if (single) { if (i < B) k_e = accumulate(k_slot, k_s, k_y); }
    else if (i < B) accumulate(A.rslot[i], A.logits[i] * A.inv_t, A.labels[i]);
    const u32 ncr = A.gcount[t];
    for (u32 k = tid; k < ncr; k += kSegThreads) {
      ++nlist;
      if (rec[A.glist[(size_t)t * kGTile + k]].flags == 3u) ++nval;                   // LW:137
    }
  }
  nval = __reduce_add_sync(0xFFFFFFFFu, nval); nlist = __reduce_add_sync(0xFFFFFFFFu, nlist);
  if (ln == 0) { if (nval) atomicAdd(&sm_misc[2], nval); if (nlist) atomicAdd(&sm_misc[3], nlist); }
  __syncthreads();
  if (tid == 0) { if (sm_misc[2]) atomicAdd(&ctl->n_valid, sm_misc[2]); if (sm_misc[3]) atomicAdd(&ctl->n_groups, sm_misc[3]); }
  stamp(ctl, 3);
  grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
  stamp(ctl, 4);
  // ---- phase 3: gradient in row order, per-list losses ------------------------------------------------------------------
  const u32 V = ld_relaxed(&ctl->n_valid);
  const float invV = V ? 1.0f / (float)V : 0.f;
  auto grad = [&](u32 slot, float s, float y, float e, bool have_e) -> float {
    if (slot == kEmpty) return 0.f;
    const LRec r = rec[slot];
    if (r.flags != 3u) return 0.f;
    if (!have_e) e = expf(s - dec_label(r.maxenc));
    return invV * (e / r.Z - y / r.sumy);                       // xent backprop: softmax - labels (LW:144, LW:167); mean over V
  };
  double lsum = 0.0;
  for (u32 t = blockIdx.x; t < ntile; t += gridDim.x) {
    const u32 i = t * kGTile + tid;
    if (i < B) A.dlogits[i] = A.inv_t * (single ? grad(k_slot, k_s, k_y, k_e, true) : grad(A.rslot[i], A.logits[i] * A.inv_t, A.labels[i], 0.f, false));
    const u32 ncr = A.gcount[t];
    for (u32 k = tid; k < ncr; k += kSegThreads) {
      const LRec r = rec[A.glist[(size_t)t * kGTile + k]];
      // lse - sum p s, max-subtracted.  A valid list whose labels sum to zero holds labels of both signs: p = y / 0 is
      // +inf and -inf (and 0 / 0), the reference's sum p (lse - z) is inf - inf = NaN (-> 0 by LW:172)
      if (r.flags == 3u) lsum += (r.sumy == 0.f) ? (double)NAN : (double)(logf(r.Z) - r.dotm / r.sumy);
    }
  }
  lsum = warp_sum(lsum);
  if (ln == 0) sm_d[w] = lsum;
  __syncthreads();
  if (tid == 0) {
    double tsum = 0;
    for (int q = 0; q < kSegWarps; ++q) tsum += sm_d[q];
    if (tsum != 0.0) atomicAdd(&ctl->loss_sum, tsum);           // (a NaN sum compares unequal to 0: it is added, and stays NaN)
  }
  stamp(ctl, 5);
  grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
  stamp(ctl, 6);
  // ---- behind the last barrier: scalars, NaN -> 0 (LW:172, value and gradient), clean arena (all CTAs) -----------------
  const double tot = *reinterpret_cast<volatile double*>(&ctl->loss_sum);
  const float lossv = V ? (float)(tot / (double)V) : 0.f;
  const bool isnan_ = lossv != lossv;
  if (blockIdx.x == 0 && tid == 0) {
    *A.loss = isnan_ ? 0.f : lossv;
    *A.n_valid = (int32_t)V;
    *A.n_group = (int32_t)ld_relaxed(&ctl->n_groups);
  }
  if (isnan_)                                                   // tf.cond takes the constant branch: no gradient
    for (u32 i = blockIdx.x * kSegThreads + tid; i < B; i += gridDim.x * kSegThreads) A.dlogits[i] = 0.f;
  clean_records_grid(A.rec, 0, A.glist, A.gcount, ntile);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) { __threadfence(); ctl->ts[23] = globaltimer(); ctl->path = 1; ctl_finish(ctl); }
  }
}

__global__ void __launch_bounds__(256) k_lw_dense_fill(size_t n, uint8_t* __restrict__ dm, float* __restrict__ dl,
                                                       float* __restrict__ dz, float pad) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (dm) dm[i] = 0; if (dl) dl[i] = 0.f; if (dz) dz[i] = pad;
  }
}

__global__ void __launch_bounds__(256) k_lw_dense_scatter(u32 B, int64_t V, const u32* __restrict__ astart,
                                                          const u32* __restrict__ perm, const float* __restrict__ ss,
                                                          const float* __restrict__ sy, const float* __restrict__ rec,
                                                          uint8_t* __restrict__ dm, float* __restrict__ dl,
                                                          float* __restrict__ dz) {
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= B) return;
  const u32 a = astart[p];
  if (rec[(size_t)R_VALID * B + a] == 0.f) return;
  const u32 r = __float_as_uint(rec[(size_t)R_RANK * B + a]);
  if ((int64_t)r >= V) return;
  const size_t o = (size_t)r * B + perm[p];
  if (dm) dm[o] = 1;
  if (dl) dl[o] = sy[p] / rec[(size_t)R_SY * B + a];              // LW:144
  if (dz) dz[o] = ss[p];                                          // LW:139-140: member logits unchanged
}

}  // namespace rn

using namespace rn;

extern "C" size_t rn_listwise_scratch_bytes(int64_t B) { return B > 0 ? make_layout(B, 1).total : 0; }
extern "C" int rn_listwise_launch_count(int64_t B) {
  if (B <= 0) return 0;
  const Layout L = make_layout(B, 1);
  (void)L;
  return 2;        // k_init, k_seg<ListwiseTail>
}

static int validate_listwise(const rn_listwise_args* a) {
  if (!a || a->B <= 0 || a->B > (1ll << 28)) return RN_ERR_ARG;
  if (!a->keys || !a->labels || !a->logits || !a->n_valid || !a->n_group || !a->dlogits) return RN_ERR_ARG;
  if (a->do_reduce && !a->loss) return RN_ERR_ARG;
  if (!(a->pos_neg_th >= 0.f)) return RN_ERR_UNSUPPORTED;   // th < 0 also fires on the dense zero padding
  if (a->inv_temperature < 0.f || a->inv_temperature != a->inv_temperature) return RN_ERR_ARG;
  const void* ptrs[] = {a->keys, a->labels, a->logits, a->row_ok, a->list_w, a->list_loss, a->dlogits};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  return RN_OK;
}

extern "C" int rn_listwise_fwd_bwd(const rn_listwise_args* a, void* scratch, size_t scratch_bytes, void* stream) {
  RN_NVTX_RANGE("rn_listwise_fwd_bwd");
  int rc = validate_listwise(a);
  if (rc) return rc;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  const Layout L = make_layout(a->B, 1);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  const float inv_t = a->inv_temperature == 0.f ? 1.0f : a->inv_temperature;       // (0 = a zero-initialised struct: no temperature)
  static const char* lwc = getenv("RN_LW_COUNT");
  if (a->scratch_persistent && a->do_reduce && !a->list_w && !a->list_loss && !(lwc && *lwc == '0')) {
    // counting form: one cooperative kernel, no sort, no scatter (the ranks of the valid lists are not needed)
    LwCountArgs A{(u32)a->B, L.cap - 1, a->keys, a->row_ok, a->labels, a->logits, a->pos_neg_th, inv_t, a->loss, a->n_valid,
                  a->n_group, a->dlogits, at<GRec>(base, L.rec), at<u32>(base, L.glist), at<u32>(base, L.gcount),
                  at<u32>(base, L.slot), at<Ctl>(base, L.ctl)};
    void* args[] = {&A};
    if (launch_coop((const void*)k_lw_count, device_sm_count(), kSegThreads, args, st) != cudaSuccess) return RN_ERR_LAUNCH;
    return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
  }
  SegInputs in{a->B, 1, a->keys, nullptr, a->row_ok, false, false};
  u32* astart = at<u32>(base, L.aj);
  u32* gend = at<u32>(base, L.cnt);
  u32* perm = at<u32>(base, L.misc);
  float* ss = at<float>(base, L.ss);
  float* sy = at<float>(base, L.sy);
  float* rec = at<float>(base, L.gstat);
  GatherCols gc{{a->logits, a->labels, nullptr, nullptr}, {ss, sy, nullptr, nullptr}};
  LwParams P{(u32)a->B, L.gbits, a->list_w, a->pos_neg_th, a->do_reduce, inv_t, a->loss, a->list_loss, a->n_valid,
             a->n_group, a->dlogits};
  ListwiseTail T{BoundsTail{astart, gend, perm, gc}, P, ss, sy, rec, at<u32>(base, L.blk)};
  if (seg_run(L, scratch, in, T, st) != cudaSuccess) return RN_ERR_LAUNCH;
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_listwise_dense(const rn_listwise_args* a, void* scratch, size_t scratch_bytes, int64_t V,
                                 uint8_t* dense_mask, float* dense_labels, float* dense_logits,
                                 int32_t do_mask_logits, float value_of_masked_logit, void* stream) {
  RN_NVTX_RANGE("rn_listwise_dense");
  if (!a || a->B <= 0 || V < 0 || !scratch) return RN_ERR_ARG;
  const Layout L = make_layout(a->B, 1);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  if (V == 0) return RN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  const size_t n = (size_t)V * (size_t)a->B;
  // LW:139-140: non-members get 0 + 1*value_of_masked_logit when masking is on, stay 0 otherwise
  k_lw_dense_fill<<<148 * 8, 256, 0, st>>>(n, dense_mask, dense_labels, dense_logits,
                                          do_mask_logits ? value_of_masked_logit : 0.f);
  k_lw_dense_scatter<<<(u32)((a->B + 255) / 256), 256, 0, st>>>((u32)a->B, V, at<u32>(base, L.aj), at<u32>(base, L.misc),
                                                                at<float>(base, L.ss), at<float>(base, L.sy),
                                                                at<float>(base, L.gstat), dense_mask, dense_labels,
                                                                dense_logits);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
