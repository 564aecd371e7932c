// Segment pooling of slot embeddings: the step that feeds the model whose logits the ranking losses score.
//
// Replaces rec_block/embedding_util.py:239-324 of the reference (embedding_using_sparse_batch_segment_ids and its helper
// sparse_batch_segment_ids_of_targets :127-198): a StaticHashTable lookup of every slot, boolean_mask, unique, gather,
// weight multiply and unsorted_segment_sum / _mean over (row, target slot) segments -- seven stock TF ops and three
// materialised [kept, D] intermediates -- as ONE gather-accumulate kernel: output segment (b, t) walks the C columns of row
// b and adds weight * E[id] for the columns whose slot is target_slots[t], in column order (the order of TF's CPU
// kernel, so sums are bit-identical to it).  HBM bound: D * 4 bytes per kept id in, B * T * D * 4 bytes out, nothing
// materialised in between.  Backward: d E[id] += w * d out[b, t] (vector atomics: ids repeat), d w = <E[id], d out[b, t]>.
#include "common.cuh"

namespace rn {

struct PoolParams {
  u32 B, C, T, D; int mean;
  const int32_t* slots; const int64_t* ids; const float* weights; const int32_t* tslots;
  const float* table; int64_t V;
};

// One thread per VEC consecutive floats of an output segment; the D / VEC threads of a segment are neighbours, so the
// slot / id / weight loads are warp broadcasts and the table rows are read as contiguous vectors.
template <int VEC>
__global__ void __launch_bounds__(256) k_pool_fwd(PoolParams P, float* __restrict__ out, u32* __restrict__ err) {
  const u32 dv = P.D / VEC;
  const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 seg = gid / dv;
  if (seg >= (u64)P.B * P.T) return;
  const u32 v = (u32)(gid - seg * dv);
  const u32 b = (u32)(seg / P.T), t = (u32)(seg - (u64)b * P.T);
  const int32_t target = P.tslots[t];
  const int32_t* srow = P.slots + (size_t)b * P.C;
  const int64_t* irow = P.ids + (size_t)b * P.C;
  const float* wrow = P.weights ? P.weights + (size_t)b * P.C : nullptr;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  u32 cnt = 0;
  for (u32 c = 0; c < P.C; ++c) {
    if (srow[c] != target) continue;
    const int64_t id = irow[c];
    if (id < 0 || id >= P.V) { if (err) atomicOr(err, 8u); continue; }   // (tf.gather would raise: flagged, skipped)
    const float w = wrow ? wrow[c] : 1.0f;
    const float* e = P.table + (size_t)id * P.D + (size_t)v * VEC;
    float ev[VEC];
    if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(e); ev[0] = q.x; ev[1] = q.y; ev[2 % VEC] = q.z; ev[3 % VEC] = q.w; }
    else { for (int k = 0; k < VEC; ++k) ev[k] = e[k]; }
    // embeddings * weights, then the segment sum: two roundings, as the reference's two ops (:310-317)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], wrow ? __fmul_rn(ev[k], w) : ev[k]);
    ++cnt;
  }
  if (P.mean && cnt > 1) {                                                // unsorted_segment_mean: sum / max(count, 1)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = __fdiv_rn(acc[k], (float)cnt);
  }
  float* o = out + (size_t)seg * P.D + (size_t)v * VEC;
  if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2 % VEC], acc[3 % VEC]);
  else { for (int k = 0; k < VEC; ++k) o[k] = acc[k]; }
}

// Warp-per-row forward kernel (the bandwidth path): a warp owns row b.  Its lanes load the row's slots / ids / weights
// 32 columns at a time (coalesced), every target slot's columns are found with ONE ballot, and the warp then streams the
// kept table rows -- all 32 lanes on one row (D floats contiguous: full sectors), four rows in flight per warp before the
// first add, added in column order (same sums as k_pool_fwd).  Lane l holds elements l * VEC + k of every 32 * VEC chunk.
template <int VEC, int CH>
__global__ void __launch_bounds__(256) k_pool_fwd_rows(PoolParams P, float* __restrict__ out, u32* __restrict__ err) {
  const u32 ln = threadIdx.x & 31u;
  const u64 b = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  const int32_t* srow = P.slots + (size_t)b * P.C;
  const int64_t* irow = P.ids + (size_t)b * P.C;
  const float* wrow = P.weights ? P.weights + (size_t)b * P.C : nullptr;
  for (u32 t = 0; t < P.T; ++t) {
    const int32_t target = P.tslots[t];
    float acc[CH][VEC];
#pragma unroll
    for (int h = 0; h < CH; ++h)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[h][k] = 0.f;
    u32 cnt = 0;
    for (u32 c0 = 0; c0 < P.C; c0 += 32) {
      const u32 c = c0 + ln;
      const bool in = c < P.C;
      const int32_t sl = in ? srow[c] : 0;
      int64_t id = 0; float w = 1.f;
      bool hit = in && sl == target;
      if (hit) { id = irow[c]; if (wrow) w = wrow[c]; }
      if (hit && (id < 0 || id >= P.V)) { if (err) atomicOr(err, 8u); hit = false; }     // (tf.gather would raise: flagged, skipped)
      u32 m = __ballot_sync(0xFFFFFFFFu, hit);
      cnt += (u32)__popc(m);
      while (m) {
        // up to four kept columns per round: their row loads are issued before the first add
        int src[4]; float wv[4]; float ev[4][CH][VEC];
#pragma unroll
        for (int r = 0; r < 4; ++r) { src[r] = m ? __ffs(m) - 1 : -1; if (m) m &= m - 1; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int sl_ = src[r] < 0 ? 0 : src[r];
          const int64_t rid = __shfl_sync(0xFFFFFFFFu, id, sl_);
          wv[r] = __shfl_sync(0xFFFFFFFFu, w, sl_);
          if (src[r] >= 0) {
            const float* e = P.table + (size_t)rid * P.D + (size_t)ln * VEC;
#pragma unroll
            for (int h = 0; h < CH; ++h) {
              if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(e + h * 32 * VEC); ev[r][h][0] = q.x; ev[r][h][1 % VEC] = q.y; ev[r][h][2 % VEC] = q.z; ev[r][h][3 % VEC] = q.w; }
              else if (VEC == 2) { const float2 q = *reinterpret_cast<const float2*>(e + h * 32 * VEC); ev[r][h][0] = q.x; ev[r][h][1 % VEC] = q.y; }
              else ev[r][h][0] = e[h * 32 * VEC];
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (src[r] >= 0) {
#pragma unroll
            for (int h = 0; h < CH; ++h)
#pragma unroll
              for (int k = 0; k < VEC; ++k) acc[h][k] = __fadd_rn(acc[h][k], wrow ? __fmul_rn(ev[r][h][k], wv[r]) : ev[r][h][k]);
          }
        }
      }
    }
    if (P.mean && cnt > 1) {
#pragma unroll
      for (int h = 0; h < CH; ++h)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[h][k] = __fdiv_rn(acc[h][k], (float)cnt);
    }
    float* o = out + ((size_t)b * P.T + t) * P.D + (size_t)ln * VEC;
#pragma unroll
    for (int h = 0; h < CH; ++h) {
      if (VEC == 4) *reinterpret_cast<float4*>(o + h * 32 * VEC) = make_float4(acc[h][0], acc[h][1 % VEC], acc[h][2 % VEC], acc[h][3 % VEC]);
      else if (VEC == 2) *reinterpret_cast<float2*>(o + h * 32 * VEC) = make_float2(acc[h][0], acc[h][1 % VEC]);
      else o[h * 32 * VEC] = acc[h][0];
    }
  }
}

// The same warp-per-row walk with ALL target slots in one pass: the accumulators of the row's T output segments live in
// shared memory (T * D floats per warp; a lane only ever touches its own columns, so there is nothing to synchronise),
// every lane classifies its column against the target list once, and the kept columns of the whole row are streamed RIN
// table rows at a time -- one dependent phase per 32 columns instead of one per target slot.  Column order per segment is
// unchanged (same sums).  T <= 32 (lane t keeps the segment's column count for the mean).
template <int VEC, int CH, int RIN>
__global__ void __launch_bounds__(256) k_pool_fwd_rows_all(PoolParams P, float* __restrict__ out, u32* __restrict__ err) {
  extern __shared__ __align__(16) float pool_acc[];
  const u32 ln = threadIdx.x & 31u, wq = threadIdx.x >> 5;
  const u64 b = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  float* acc = pool_acc + (size_t)wq * P.T * P.D;
  const int32_t* srow = P.slots + (size_t)b * P.C;
  const int64_t* irow = P.ids + (size_t)b * P.C;
  const float* wrow = P.weights ? P.weights + (size_t)b * P.C : nullptr;
  for (u32 t = 0; t < P.T; ++t)
#pragma unroll
    for (int h = 0; h < CH; ++h)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[(size_t)t * P.D + h * 32 * VEC + ln * VEC + k] = 0.f;
  const int32_t my_target = ln < P.T ? P.tslots[ln] : 0;
  u32 cnt = 0;                                  // lane t: kept columns of target t
  for (u32 c0 = 0; c0 < P.C; c0 += 32) {
    const u32 c = c0 + ln;
    const bool in = c < P.C;
    const int32_t sl = in ? srow[c] : 0;
    // which target (if any) is this column's slot?  every lane's slot is compared with lane t's target
    int tix = -1;
    for (u32 t = 0; t < P.T; ++t) {
      const int32_t tg = __shfl_sync(0xFFFFFFFFu, my_target, (int)t);
      const u32 bt = __ballot_sync(0xFFFFFFFFu, in && sl == tg);
      if (in && sl == tg) tix = (int)t;
      if (ln == t) cnt += (u32)__popc(bt);
    }
    int64_t id = 0; float w = 1.f;
    bool hit = tix >= 0;
    if (hit) { id = irow[c]; if (wrow) w = wrow[c]; }
    if (hit && (id < 0 || id >= P.V)) { if (err) atomicOr(err, 8u); hit = false; }     // (tf.gather would raise: flagged, skipped)
    u32 m = __ballot_sync(0xFFFFFFFFu, hit);
    while (m) {
      int src[RIN]; float wv[RIN]; int tt[RIN]; float ev[RIN][CH][VEC];
#pragma unroll
      for (int r = 0; r < RIN; ++r) { src[r] = m ? __ffs(m) - 1 : -1; if (m) m &= m - 1; }
#pragma unroll
      for (int r = 0; r < RIN; ++r) {
        const int sl_ = src[r] < 0 ? 0 : src[r];
        const int64_t rid = __shfl_sync(0xFFFFFFFFu, id, sl_);
        wv[r] = __shfl_sync(0xFFFFFFFFu, w, sl_);
        tt[r] = __shfl_sync(0xFFFFFFFFu, tix, sl_);
        if (src[r] >= 0) {
          const float* e = P.table + (size_t)rid * P.D + (size_t)ln * VEC;
#pragma unroll
          for (int h = 0; h < CH; ++h) {
            if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(e + h * 32 * VEC); ev[r][h][0] = q.x; ev[r][h][1 % VEC] = q.y; ev[r][h][2 % VEC] = q.z; ev[r][h][3 % VEC] = q.w; }
            else if (VEC == 2) { const float2 q = *reinterpret_cast<const float2*>(e + h * 32 * VEC); ev[r][h][0] = q.x; ev[r][h][1 % VEC] = q.y; }
            else ev[r][h][0] = e[h * 32 * VEC];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RIN; ++r) {
        if (src[r] >= 0) {
          float* a = acc + (size_t)tt[r] * P.D + ln * VEC;
#pragma unroll
          for (int h = 0; h < CH; ++h)
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              a[h * 32 * VEC + k] = __fadd_rn(a[h * 32 * VEC + k], wrow ? __fmul_rn(ev[r][h][k], wv[r]) : ev[r][h][k]);
        }
      }
    }
  }
  for (u32 t = 0; t < P.T; ++t) {
    const u32 ct = __shfl_sync(0xFFFFFFFFu, cnt, (int)t);
    float* o = out + ((size_t)b * P.T + t) * P.D + (size_t)ln * VEC;
    const float* a = acc + (size_t)t * P.D + ln * VEC;
#pragma unroll
    for (int h = 0; h < CH; ++h) {
      float v[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) { v[k] = a[h * 32 * VEC + k]; if (P.mean && ct > 1) v[k] = __fdiv_rn(v[k], (float)ct); }
      if (VEC == 4) *reinterpret_cast<float4*>(o + h * 32 * VEC) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
      else if (VEC == 2) *reinterpret_cast<float2*>(o + h * 32 * VEC) = make_float2(v[0], v[1 % VEC]);
      else o[h * 32 * VEC] = v[0];
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) k_pool_bwd(PoolParams P, const float* __restrict__ d_out, float* __restrict__ d_table,
                                                  float* __restrict__ d_weights) {
  const u32 dv = P.D / VEC;
  const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 seg = gid / dv;
  if (seg >= (u64)P.B * P.T) return;
  const u32 v = (u32)(gid - seg * dv);
  const u32 b = (u32)(seg / P.T), t = (u32)(seg - (u64)b * P.T);
  const int32_t target = P.tslots[t];
  const int32_t* srow = P.slots + (size_t)b * P.C;
  const int64_t* irow = P.ids + (size_t)b * P.C;
  const float* wrow = P.weights ? P.weights + (size_t)b * P.C : nullptr;
  float g[VEC];
  {
    const float* go = d_out + (size_t)seg * P.D + (size_t)v * VEC;
    if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(go); g[0] = q.x; g[1] = q.y; g[2 % VEC] = q.z; g[3 % VEC] = q.w; }
    else { for (int k = 0; k < VEC; ++k) g[k] = go[k]; }
  }
  if (P.mean) {
    u32 cnt = 0;
    for (u32 c = 0; c < P.C; ++c) cnt += srow[c] == target ? 1u : 0u;
    if (cnt > 1) { const float r = 1.0f / (float)cnt; for (int k = 0; k < VEC; ++k) g[k] *= r; }
  }
  for (u32 c = 0; c < P.C; ++c) {
    if (srow[c] != target) continue;
    const int64_t id = irow[c];
    if (id < 0 || id >= P.V) continue;
    const float w = wrow ? wrow[c] : 1.0f;
    if (d_table) {
      float* dt = d_table + (size_t)id * P.D + (size_t)v * VEC;
      if (VEC == 4) atomicAdd(reinterpret_cast<float4*>(dt), make_float4(g[0] * w, g[1] * w, g[2 % VEC] * w, g[3 % VEC] * w));
      else { for (int k = 0; k < VEC; ++k) atomicAdd(dt + k, g[k] * w); }
    }
    if (d_weights) {
      const float* e = P.table + (size_t)id * P.D + (size_t)v * VEC;
      float dot = 0.f;
      for (int k = 0; k < VEC; ++k) dot = fmaf(e[k], g[k], dot);
      atomicAdd(d_weights + (size_t)b * P.C + c, dot);
    }
  }
}

// Vector reductions (sm_90+: REDG.E.ADD.F32x2 / x4): one instruction adds a lane's 8 / 16 contiguous bytes, so a warp
// covers whole sectors of a table row with one atomic operation each.
__device__ __forceinline__ void red_add_v2(float* a, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* a, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// Warp-per-row backward (same walk as k_pool_fwd_rows_all): for every kept column the warp reads the segment's slice of
// d out (one contiguous row), scales it by the column's weight (and 1 / count for the mean) and adds it to the id's row
// of d table with vector reductions -- eight full sectors per row instead of one scalar atomic per element; d weights is
// a warp-reduced dot product stored by lane 0 (one warp owns a (row, column): no atomic).  T <= 32.
template <int VEC, int CH>
__global__ void __launch_bounds__(256) k_pool_bwd_rows(PoolParams P, const float* __restrict__ d_out, float* __restrict__ d_table,
                                                       float* __restrict__ d_weights) {
  const u32 ln = threadIdx.x & 31u;
  const u64 b = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  const int32_t* srow = P.slots + (size_t)b * P.C;
  const int64_t* irow = P.ids + (size_t)b * P.C;
  const float* wrow = P.weights ? P.weights + (size_t)b * P.C : nullptr;
  const int32_t my_target = ln < P.T ? P.tslots[ln] : 0;
  u32 cnt = 0;                                  // lane t: kept columns of target t (mean only)
  if (P.mean) {
    for (u32 c0 = 0; c0 < P.C; c0 += 32) {
      const u32 c = c0 + ln;
      const int32_t sl = c < P.C ? srow[c] : 0;
      for (u32 t = 0; t < P.T; ++t) {
        const int32_t tg = __shfl_sync(0xFFFFFFFFu, my_target, (int)t);
        const u32 bt = __ballot_sync(0xFFFFFFFFu, c < P.C && sl == tg);
        if (ln == t) cnt += (u32)__popc(bt);
      }
    }
  }
  for (u32 c0 = 0; c0 < P.C; c0 += 32) {
    const u32 c = c0 + ln;
    const bool in = c < P.C;
    const int32_t sl = in ? srow[c] : 0;
    int tix = -1;
    for (u32 t = 0; t < P.T; ++t) {
      const int32_t tg = __shfl_sync(0xFFFFFFFFu, my_target, (int)t);
      if (in && sl == tg) tix = (int)t;
    }
    int64_t id = 0; float w = 1.f;
    bool hit = tix >= 0;
    if (hit) { id = irow[c]; if (wrow) w = wrow[c]; }
    if (hit && (id < 0 || id >= P.V)) hit = false;
    u32 m = __ballot_sync(0xFFFFFFFFu, hit);
    float dw_mine = 0.f;                        // d weights of this lane's column (filled when its turn comes)
    while (m) {
      const int src = __ffs(m) - 1; m &= m - 1;
      const int64_t rid = __shfl_sync(0xFFFFFFFFu, id, src);
      const float wv = __shfl_sync(0xFFFFFFFFu, w, src);
      const int tt = __shfl_sync(0xFFFFFFFFu, tix, src);
      float scale = wv;
      if (P.mean) { const u32 ct = __shfl_sync(0xFFFFFFFFu, cnt, tt); if (ct > 1) scale = wv / (float)ct; }
      const float* go = d_out + ((size_t)b * P.T + (u32)tt) * P.D + (size_t)ln * VEC;
      float* dt = d_table ? d_table + (size_t)rid * P.D + (size_t)ln * VEC : nullptr;
      const float* e = P.table + (size_t)rid * P.D + (size_t)ln * VEC;
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < CH; ++h) {
        float g[VEC];
        if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(go + h * 32 * VEC); g[0] = q.x; g[1 % VEC] = q.y; g[2 % VEC] = q.z; g[3 % VEC] = q.w; }
        else if (VEC == 2) { const float2 q = *reinterpret_cast<const float2*>(go + h * 32 * VEC); g[0] = q.x; g[1 % VEC] = q.y; }
        else g[0] = go[h * 32 * VEC];
        if (dt) {
          if (VEC == 4) red_add_v4(dt + h * 32 * VEC, g[0] * scale, g[1 % VEC] * scale, g[2 % VEC] * scale, g[3 % VEC] * scale);
          else if (VEC == 2) red_add_v2(dt + h * 32 * VEC, g[0] * scale, g[1 % VEC] * scale);
          else atomicAdd(dt + h * 32 * VEC, g[0] * scale);
        }
        if (d_weights) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) dot = fmaf(e[h * 32 * VEC + k], g[k], dot);
        }
      }
      if (d_weights) {
        dot = warp_sum(dot);
        if (P.mean) { const u32 ct = __shfl_sync(0xFFFFFFFFu, cnt, tt); if (ct > 1) dot /= (float)ct; }
        if ((int)ln == src) dw_mine = dot;
      }
    }
    if (d_weights && hit) d_weights[(size_t)b * P.C + c] += dw_mine;       // (this warp is the only writer of the row's columns)
  }
}

static int pool_params(const rn_pool_args* a, PoolParams& P) {
  if (!a || a->B <= 0 || a->C <= 0 || a->T <= 0 || a->D <= 0 || a->V <= 0) return RN_ERR_ARG;
  if (a->B > 0x7FFFFFFFll || a->C > 0x7FFFFFFFll) return RN_ERR_ARG;
  if (!a->slots || !a->ids || !a->target_slots || !a->table) return RN_ERR_ARG;
  P.B = (u32)a->B; P.C = (u32)a->C; P.T = (u32)a->T; P.D = (u32)a->D; P.mean = a->mean;
  P.slots = a->slots; P.ids = a->ids; P.weights = a->weights; P.tslots = a->target_slots; P.table = a->table; P.V = a->V;
  return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" int rn_segment_pool_fwd(const rn_pool_args* a, float* out, uint32_t* err_flag, void* stream) {
  RN_NVTX_RANGE("rn_segment_pool_fwd");
  PoolParams P;
  int rc = pool_params(a, P);
  if (rc) return rc;
  if (!out) return RN_ERR_ARG;
  const bool vec = (P.D % 4) == 0 && check_align(a->table) == RN_OK && check_align(out) == RN_OK;
  const u64 threads = (u64)P.B * P.T * (vec ? P.D / 4 : P.D);
  const u64 grid = (threads + 255) / 256;
  if (grid > 0x7FFFFFFFull) return RN_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // widths that are whole warps of float / float2 / float4: the warp-per-row kernel (one table row per load instruction)
  if (check_align(a->table) == RN_OK && check_align(out) == RN_OK) {
    const u64 rgrid = ((u64)P.B * 32 + 255) / 256;
    if (rgrid <= 0x7FFFFFFFull) {
      // all target slots in one pass when their accumulators fit the warps' shared memory
      const size_t acc_bytes = (size_t)8 * P.T * P.D * sizeof(float);
      if (P.T <= 32 && acc_bytes <= 48 * 1024) {
        bool done = true;
        switch (P.D) {
          case 32:  k_pool_fwd_rows_all<1, 1, 8><<<(unsigned)rgrid, 256, acc_bytes, st>>>(P, out, err_flag); break;
          case 64:  k_pool_fwd_rows_all<2, 1, 8><<<(unsigned)rgrid, 256, acc_bytes, st>>>(P, out, err_flag); break;
          case 128: k_pool_fwd_rows_all<4, 1, 4><<<(unsigned)rgrid, 256, acc_bytes, st>>>(P, out, err_flag); break;
          case 256: k_pool_fwd_rows_all<4, 2, 4><<<(unsigned)rgrid, 256, acc_bytes, st>>>(P, out, err_flag); break;
          default: done = false;
        }
        if (done) return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
      }
      bool done = true;
      switch (P.D) {
        case 32:  k_pool_fwd_rows<1, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, out, err_flag); break;
        case 64:  k_pool_fwd_rows<2, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, out, err_flag); break;
        case 128: k_pool_fwd_rows<4, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, out, err_flag); break;
        case 256: k_pool_fwd_rows<4, 2><<<(unsigned)rgrid, 256, 0, st>>>(P, out, err_flag); break;
        default: done = false;
      }
      if (done) return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
    }
  }
  if (vec) k_pool_fwd<4><<<(unsigned)grid, 256, 0, st>>>(P, out, err_flag);
  else k_pool_fwd<1><<<(unsigned)grid, 256, 0, st>>>(P, out, err_flag);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_segment_pool_bwd(const rn_pool_args* a, const float* d_out, float* d_table, float* d_weights, void* stream) {
  RN_NVTX_RANGE("rn_segment_pool_bwd");
  PoolParams P;
  int rc = pool_params(a, P);
  if (rc) return rc;
  if (!d_out || (!d_table && !d_weights)) return RN_ERR_ARG;
  if (d_weights && !a->weights) return RN_ERR_ARG;
  const bool vec = (P.D % 4) == 0 && check_align(a->table) == RN_OK && check_align(d_out) == RN_OK &&
                   (!d_table || check_align(d_table) == RN_OK);
  const u64 threads = (u64)P.B * P.T * (vec ? P.D / 4 : P.D);
  const u64 grid = (threads + 255) / 256;
  if (grid > 0x7FFFFFFFull) return RN_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec && P.T <= 32) {
    const u64 rgrid = ((u64)P.B * 32 + 255) / 256;
    if (rgrid <= 0x7FFFFFFFull) {
      bool done = true;
      switch (P.D) {
        case 32:  k_pool_bwd_rows<1, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, d_out, d_table, d_weights); break;
        case 64:  k_pool_bwd_rows<2, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, d_out, d_table, d_weights); break;
        case 128: k_pool_bwd_rows<4, 1><<<(unsigned)rgrid, 256, 0, st>>>(P, d_out, d_table, d_weights); break;
        case 256: k_pool_bwd_rows<4, 2><<<(unsigned)rgrid, 256, 0, st>>>(P, d_out, d_table, d_weights); break;
        default: done = false;
      }
      if (done) return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
    }
  }
  if (vec) k_pool_bwd<4><<<(unsigned)grid, 256, 0, st>>>(P, d_out, d_table, d_weights);
  else k_pool_bwd<1><<<(unsigned)grid, 256, 0, st>>>(P, d_out, d_table, d_weights);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
