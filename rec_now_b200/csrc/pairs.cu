// K4 — pair materialisation in the reference's row-major order, and occurance_power_weight.
//
// Replaces tf.boolean_mask(reshape(mat,[-1]), pair_mask) (pairwise_loss_from_batch.py:206-217), i.e. the
// tf.where scan over B^2 booleans + gathers, for callers that hand pairwise_loss an arbitrary pairloss_func
// or label_pair_to_weight_func (as the reference's own test does, tests/rec_block/
// test_pairwise_loss_from_batch.py:38-45).  The batch is sorted by (group, row), so a row's candidates are
// one contiguous run whose rows are already ascending: count per row, exclusive scan in ORIGINAL row order,
// fill -> pairs come out by i ascending then j ascending, exactly PW:217's order.
#include "segment.cuh"

namespace rn {

struct PiParams {
  u32 B; int label_cond; int only_wrong; int label_func; int has_rwp, has_rwn;
};

// predicate + weight of candidate (p -> q), both sorted positions of the same group
__device__ __forceinline__ bool pair_pred(const PiParams& P, u32 p, u32 q, float si, float yi, float wpi, float sj,
                                          float yj, float wnj, float& w) {
  w = 1.f;
  if (q == p) return false;                                        // PW:36 identity removed
  if (!P.label_cond) return true;
  bool c = yi > yj;                                                // PW:189
  if (P.label_func == RN_LABEL_DIFF || P.has_rwp || P.has_rwn) {   // PW:192-193: C = W > 0
    w = (P.label_func == RN_LABEL_DIFF) ? (yi - yj) * (c ? 1.f : 0.f) : (c ? 1.f : 0.f);
    if (P.has_rwp) w *= wpi;
    if (P.has_rwn) w *= wnj;
    c = w > 0.f;
  }
  if (P.only_wrong) c = c && (si < sj);                            // PW:200-202
  return c;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_pi(PiParams P, const u32* __restrict__ astart, const u32* __restrict__ gend,
                                            const u32* __restrict__ perm, const float* __restrict__ ss,
                                            const float* __restrict__ sy, const float* __restrict__ swp,
                                            const float* __restrict__ swn, u32* __restrict__ rowcnt,
                                            const u64* __restrict__ offs, int32_t* __restrict__ pos_idx,
                                            int32_t* __restrict__ neg_idx, float* __restrict__ wout, u64 capacity,
                                            Ctl* ctl) {
  const u32 ln = lane_id();
  const u32 p = (blockIdx.x * 256u + threadIdx.x) >> 5;           // one warp per sorted row
  if (p >= P.B) return;
  const u32 a = astart[p], e = gend[a], row = perm[p];
  const float si = ss[p], yi = sy[p], wpi = P.has_rwp ? swp[p] : 1.f;
  u64 base = FILL ? offs[row] : 0;
  u32 cnt = 0;
  for (u32 q0 = a; q0 < e; q0 += 32) {
    const u32 q = q0 + ln;
    bool c = false; float w = 1.f;
    if (q < e) c = pair_pred(P, p, q, si, yi, wpi, ss[q], sy[q], P.has_rwn ? swn[q] : 1.f, w);
    const u32 bal = __ballot_sync(0xFFFFFFFFu, c);
    if (FILL) {
      if (c) {
        const u64 o = base + __popc(bal & lanemask_lt());
        if (o < capacity) {
          pos_idx[o] = (int32_t)row; neg_idx[o] = (int32_t)perm[q];
          if (wout) wout[o] = w;
        } else {
          atomicOr(&ctl->err, 2u);
        }
      }
      base += __popc(bal);
    } else {
      cnt += __popc(bal);
    }
  }
  if (!FILL && ln == 0) rowcnt[row] = cnt;
}

// single-block exclusive scan u32 -> u64 over rows in original order; total to offs[B] and ctl->n_pair
__global__ void __launch_bounds__(1024) k_scan_rows(u32 B, const u32* __restrict__ rowcnt, u64* __restrict__ offs,
                                                    Ctl* ctl) {
  __shared__ u64 wsum[32];
  __shared__ u64 s_carry;
  const u32 ln = lane_id(), w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u32 i0 = 0; i0 < B; i0 += 1024) {
    const u32 i = i0 + threadIdx.x;
    const u64 v = i < B ? rowcnt[i] : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += t; }
    if (ln == 31) wsum[w] = inc;
    __syncthreads();
    u64 off = s_carry;
    for (u32 k = 0; k < w; ++k) off += wsum[k];
    if (i < B) offs[i] = off + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) { offs[B] = s_carry; ctl->n_pair = s_carry; }
}

// ---- occurance_power_weight -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_occ_insert(const int64_t* __restrict__ ids, u32 N, u32* table, u32* count,
                                                    u32* __restrict__ slot, u32 capmask) {
  const u32 i = blockIdx.x * 256u + threadIdx.x;
  if (i >= N) return;
  const int64_t key = ids[i];
  u32 s = (u32)mix64(0x9E3779B97F4A7C15ull ^ (u64)key) & capmask;
  for (;;) {
    u32 cur = ld_relaxed(table + s);
    if (cur == kEmpty) { u32 prev = atomicCAS(table + s, kEmpty, i); cur = (prev == kEmpty) ? i : prev; }
    if (cur == i || ids[cur] == key) break;
    s = (s + 1) & capmask;
  }
  slot[i] = s;
  atomicAdd(count + s, 1u);
}
__global__ void __launch_bounds__(256) k_occ_out(u32 N, const u32* __restrict__ count, const u32* __restrict__ slot,
                                                 float power, float* __restrict__ out) {
  const u32 i = blockIdx.x * 256u + threadIdx.x;
  if (i >= N) return;
  const float c = (float)count[slot[i]];                           // PW:147
  out[i] = (power == 1.0f) ? c : powf(c, power);                   // PW:148-149
}
__global__ void __launch_bounds__(256) k_fill_u32(u32* p, size_t n, u32 v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace rn

using namespace rn;

extern "C" size_t rn_pair_indices_scratch_bytes(int64_t B, int32_t K) { return (B > 0 && K > 0) ? make_layout(B, K).total : 0; }

static int pi_common(const rn_pairwise_args* a, void* scratch, size_t scratch_bytes, Layout& L) {
  if (a && a->block_rows) return RN_ERR_UNSUPPORTED;      // pair materialisation reads contiguous columns only
  if (!a || a->B <= 0 || a->B > (1ll << 28) || a->K <= 0 || a->K > 8) return RN_ERR_ARG;
  if (!a->keys || !a->logits || !a->labels) return RN_ERR_ARG;
  if (a->label_func != RN_LABEL_STEP && a->label_func != RN_LABEL_DIFF) return RN_ERR_UNSUPPORTED;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  L = make_layout(a->B, a->K);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  return RN_OK;
}

static PiParams pi_params(const rn_pairwise_args* a, int label_cond) {
  return PiParams{(u32)a->B, label_cond, a->only_wrong, a->label_func, a->rw_pos ? 1 : 0, a->rw_neg ? 1 : 0};
}

extern "C" int rn_pair_indices_count(const rn_pairwise_args* a, int32_t label_cond, void* scratch, size_t scratch_bytes,
                                     int64_t* n_pairs_host, void* stream) {
  RN_NVTX_RANGE("rn_pair_indices_count");
  Layout L;
  int rc = pi_common(a, scratch, scratch_bytes, L);
  if (rc) return rc;
  if (!n_pairs_host) return RN_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  // NaN labels: with the label condition they pair with nothing (trash); as raw candidates they stay
  SegInputs in{a->B, a->K, a->keys, a->labels, a->row_ok, false, label_cond != 0};
  GatherCols gc{{a->logits, a->labels, a->rw_pos, a->rw_neg},
                {at<float>(base, L.ss), at<float>(base, L.sy), at<float>(base, L.swp), at<float>(base, L.swn)}};
  u32* perm = at<u32>(base, L.slot1);       // slot1 is free here (K > 1 primary slots are not needed)
  BoundsTail T{at<u32>(base, L.aj), at<u32>(base, L.cnt), perm, gc};
  if (seg_run(L, scratch, in, T, st) != cudaSuccess) return RN_ERR_LAUNCH;
  const PiParams P = pi_params(a, label_cond);
  const u32 gw = (u32)((a->B * 32 + 255) / 256);
  k_pi<false><<<gw, 256, 0, st>>>(P, at<u32>(base, L.aj), at<u32>(base, L.cnt), perm, at<float>(base, L.ss),
                                  at<float>(base, L.sy), at<float>(base, L.swp), at<float>(base, L.swn),
                                  at<u32>(base, L.slot), nullptr, nullptr, nullptr, nullptr, 0, at<Ctl>(base, L.ctl));
  k_scan_rows<<<1, 1024, 0, st>>>((u32)a->B, at<u32>(base, L.slot), at<u64>(base, L.misc), at<Ctl>(base, L.ctl));
  if (cudaGetLastError() != cudaSuccess) return RN_ERR_LAUNCH;
  u64 total = 0;
  if (cudaMemcpyAsync(&total, at<u64>(base, L.misc) + a->B, sizeof(u64), cudaMemcpyDeviceToHost, st) != cudaSuccess) return RN_ERR_LAUNCH;
  if (cudaStreamSynchronize(st) != cudaSuccess) return RN_ERR_LAUNCH;
  *n_pairs_host = (int64_t)total;
  return RN_OK;
}

extern "C" int rn_pair_indices_fill(const rn_pairwise_args* a, int32_t label_cond, void* scratch, size_t scratch_bytes,
                                    int32_t* pos_idx, int32_t* neg_idx, float* w, int64_t capacity, void* stream) {
  RN_NVTX_RANGE("rn_pair_indices_fill");
  Layout L;
  int rc = pi_common(a, scratch, scratch_bytes, L);
  if (rc) return rc;
  if (capacity < 0 || (capacity > 0 && (!pos_idx || !neg_idx))) return RN_ERR_ARG;
  if (capacity == 0) return RN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  const PiParams P = pi_params(a, label_cond);
  const u32 gw = (u32)((a->B * 32 + 255) / 256);
  k_pi<true><<<gw, 256, 0, st>>>(P, at<u32>(base, L.aj), at<u32>(base, L.cnt), at<u32>(base, L.slot1),
                                 at<float>(base, L.ss), at<float>(base, L.sy), at<float>(base, L.swp),
                                 at<float>(base, L.swn), nullptr, at<u64>(base, L.misc), pos_idx, neg_idx, w,
                                 (u64)capacity, at<Ctl>(base, L.ctl));
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

// occurrence arena: table[cap] | count[cap] | slot[N]
static void occ_layout(int64_t N, u32& cap, size_t& o_table, size_t& o_count, size_t& o_slot, size_t& total) {
  cap = 1024; while ((int64_t)cap < 2 * N) cap <<= 1;
  o_table = 0; o_count = align_up(sizeof(u32) * cap); o_slot = o_count + align_up(sizeof(u32) * cap);
  total = o_slot + align_up(sizeof(u32) * N);
}

extern "C" size_t rn_occurrence_scratch_bytes(int64_t N) {
  if (N <= 0) return 0;
  u32 cap; size_t a, b, c, t; occ_layout(N, cap, a, b, c, t); return t;
}

extern "C" int rn_occurrence_power_weight(const int64_t* ids, int64_t N, float power, float* out, void* scratch,
                                          size_t scratch_bytes, void* stream) {
  RN_NVTX_RANGE("rn_occurrence_power_weight");
  if (!ids || !out || N <= 0 || N > (1ll << 28) || !scratch) return RN_ERR_ARG;
  if (check_align(ids) || check_align(out) || check_align(scratch)) return RN_ERR_ALIGN;
  u32 cap; size_t o_table, o_count, o_slot, total;
  occ_layout(N, cap, o_table, o_count, o_slot, total);
  if (scratch_bytes < total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  k_fill_u32<<<148, 256, 0, st>>>(at<u32>(base, o_table), cap, kEmpty);
  k_fill_u32<<<148, 256, 0, st>>>(at<u32>(base, o_count), cap, 0u);
  const u32 g = (u32)((N + 255) / 256);
  k_occ_insert<<<g, 256, 0, st>>>(ids, (u32)N, at<u32>(base, o_table), at<u32>(base, o_count), at<u32>(base, o_slot), cap - 1);
  k_occ_out<<<g, 256, 0, st>>>((u32)N, at<u32>(base, o_count), at<u32>(base, o_slot), power, out);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
