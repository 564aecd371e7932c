// K5/K6 — in-batch listwise loss: segmented log-sum-exp softmax cross-entropy, forward + backward.
//
// Replaces to_listwise_sample (listwise_loss_from_batch.py:89-148: unique_with_counts + four (G,B)
// densifications + row filters) and listwise_loss_via_softmax_cross_entropy_with_logits (LW:151-173).
// After K1 (sort by (first-occurrence id, row)) a list is a contiguous run of sorted positions and the runs
// appear in first-occurrence order, so valid-list ranks (the row order of the reference's dense outputs) are
// an exclusive scan over run heads.  HBM-bound: 20 algorithmic bytes per sample.
#include "segment.cuh"

namespace rn {

struct LwParams {
  u32 B; int gbits;
  const float* list_w; float th; int do_reduce;
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
      for (u32 q = a + ln; q < e; q += 32) m = fmaxf(m, ss[q]);
      m = warp_maxf(m);
      float z = 0.f, sumy = 0.f; int hp = 0, hn = 0;
      for (u32 q = a + ln; q < e; q += 32) {
        const float s = ss[q], y = sy[q];
        z += expf(s - m); sumy += y;
        hp |= (y > P.th); hn |= ((y - P.th) < 0.f);                // LW:135-136
      }
      z = warp_sum(z); sumy = warp_sum(sumy);
      const bool valid = __any_sync(0xFFFFFFFFu, hp) && __any_sync(0xFFFFFFFFu, hn);   // LW:137
      const float lse = logf(z);
      float dot = 0.f;
      if (valid)
        for (u32 q = a + ln; q < e; q += 32) dot += (sy[q] / sumy) * (lse - (ss[q] - m));   // LW:144, LW:167
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

  __device__ __forceinline__ void run(const SegParams& S, const Plan& pl, const u64* key, const u32* val, u32* smem) const {
    Ctl* ctl = S.ctl;
    const u32 B = P.B, ln = lane_id(), w = threadIdx.x >> 5;
    const u32 nchunks = (B + kSegThreads - 1) / kSegThreads;
    const u32* astart = bounds.astart;
    bounds.run(S, pl, key, val, smem);
    grid_sync(&ctl->bar_cnt, &ctl->bar_gen, &ctl->err);
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
    grid_sync(&ctl->bar_cnt, &ctl->bar_gen, &ctl->err);
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
    grid_sync(&ctl->bar_cnt, &ctl->bar_gen, &ctl->err);
    // ---- gradient + scalars -------------------------------------------------------------------------------
    const u32 V = ld_relaxed(&ctl->n_valid);
    const u32 gtid = blockIdx.x * kSegThreads + threadIdx.x, gthreads = gridDim.x * kSegThreads;
    for (u32 p = gtid; p < B; p += gthreads) {
      const u32 a = astart[p];
      float g = 0.f;
      if (rec[(size_t)R_VALID * B + a] != 0.f) {
        const float m = rec[(size_t)R_MAX * B + a], lse = rec[(size_t)R_LSE * B + a], sumy = rec[(size_t)R_SY * B + a];
        float wr = 1.f;
        if (P.list_w) wr = P.list_w[__float_as_uint(rec[(size_t)R_RANK * B + a])];
        if (P.do_reduce) wr /= (float)V;
        g = wr * (expf(ss[p] - m - lse) - sy[p] / sumy);               // xent backprop: softmax - labels
      }
      P.dlogits[bounds.perm[p]] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const double t = *reinterpret_cast<volatile double*>(&ctl->loss_sum);
      if (P.do_reduce) *P.loss = V ? (float)(t / (double)V) : 0.f;   // LW:171-172 (mean; NaN -> 0)
      *P.n_valid = (int32_t)V;
      *P.n_group = (int32_t)ld_relaxed(&ctl->n_groups);
    }
  }
};

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
  const void* ptrs[] = {a->keys, a->labels, a->logits, a->row_ok, a->list_w, a->list_loss, a->dlogits};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  return RN_OK;
}

extern "C" int rn_listwise_fwd_bwd(const rn_listwise_args* a, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = validate_listwise(a);
  if (rc) return rc;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  const Layout L = make_layout(a->B, 1);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  SegInputs in{a->B, 1, a->keys, nullptr, a->row_ok, false, false};
  u32* astart = at<u32>(base, L.aj);
  u32* gend = at<u32>(base, L.cnt);
  u32* perm = at<u32>(base, L.misc);
  float* ss = at<float>(base, L.ss);
  float* sy = at<float>(base, L.sy);
  float* rec = at<float>(base, L.gstat);
  GatherCols gc{{a->logits, a->labels, nullptr, nullptr}, {ss, sy, nullptr, nullptr}};
  LwParams P{(u32)a->B, L.gbits, a->list_w, a->pos_neg_th, a->do_reduce, a->loss, a->list_loss, a->n_valid,
             a->n_group, a->dlogits};
  ListwiseTail T{BoundsTail{astart, gend, perm, gc}, P, ss, sy, rec, at<u32>(base, L.blk)};
  if (seg_run(L, scratch, in, T, st) != cudaSuccess) return RN_ERR_LAUNCH;
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_listwise_dense(const rn_listwise_args* a, void* scratch, size_t scratch_bytes, int64_t V,
                                 uint8_t* dense_mask, float* dense_labels, float* dense_logits,
                                 int32_t do_mask_logits, float value_of_masked_logit, void* stream) {
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
