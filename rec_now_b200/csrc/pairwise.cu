// K2/K3 -- fused in-batch pairwise logistic (BPR) loss, forward + backward, on the segmented batch.
//
// Replaces pairwise_loss_from_batch.py:254-274 + TF autodiff.  After K1 the rows are
// sorted by (group, label, row), so for a row at sorted position p with group start a(p) and label-level
// start l(p) the negatives of p are exactly the positions [a(p), l(p)): pair enumeration is position
// arithmetic, pair COUNTS are exact integers (l - a), and the kept pairs of a group form a staircase of
// dense rectangles.
//
// Two launches per call after k_init:
//   k_seg<HeadsTail>  segmentation (segment.cuh) + heads: group / level starts, gathers into sorted order,
//                     exact counts, per-I-block J ranges and the work list.
//   k_pair            persistent cooperative kernel.  A warp owns an I-block of 64 positive rows (two per
//                     lane, in registers) and walks its J range in blocks of 32 negatives that rotate through
//                     the lanes with shfl.bfly: every lane meets every negative exactly once, the two rows share
//                     the shuffled score and send ONE summed dL/ds_j back.  Fast path (both rows cover the whole
//                     J block, weights constant over it): per pair ex2 + rcp on the SFU, the log is taken once
//                     per 32 pairs of the running product of (1+e); ~12 FP32 ops per pair.  Edge tiles take the
//                     masked general path.  After a grid barrier the same kernel finalises: exact counts (when
//                     they depend on scores), occurrence weights, 1/n, un-permutation of the gradient, loss.
// No shared memory in the pair loop, no dense contraction, no tensor cores: the roofline is the SFU pipe.
#include <stdlib.h>
#include "segment.cuh"
#include "pair_tiles.cuh"

namespace rn {

// developer tuning knobs (read once from the environment; defaults are the shipped configuration)
static int tune_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

constexpr u32 kMaxUnitC = 64;
constexpr u32 kMaxNibS = 8192;           // I-blocks whose cost prefixes fit k_pair's shared memory (B <= 524288)
constexpr int kPairThreads = 1024;      // threads per CTA of the pair kernel (one CTA per SM, <= 64 registers)
constexpr int kPairWarps = kPairThreads / 32;

struct PairParams {
  u32 B; int K; int gbits;
  const float* logits; const float* labels; const float* rw_pos; const float* rw_neg;
  float c_log2;          // factor * log2(e)
  float factor, power; int reduce_mean; int dyn_count; int debug;
  int det;               // deterministic mode: fixed-point gradient accumulators, ordered loss partials
  float focal_w, focal_alpha, focal_gamma; int focal_stop;      // fused focal term (rn_pairwise_args.focal_*); focal_w = 0: off
  float margin;          // hinge pair loss (rn_pairwise_args.pair_loss = RN_LOSS_HINGE): max(0, margin - x); c_log2 is then the plain factor
  double loss_unit;      // unit of the accumulated pair losses: ln 2 (logistic, log2 units) or 1 (hinge)
  int lambda;            // RN_LABEL_LAMBDA: the scatter zeroes the pre-pass's per-group accumulators (cnt, g64)
  int gain2;             // what the sorted label column holds: 0 the label; 1 (RN_LABEL_GAIN2) 2^y (W = 2^y_i - 2^y_j);
                         // 2 (RN_LABEL_LUT) the label LEVEL 0 .. 7 (W = weight_lut[l_i][l_j])
  int part_rank, part_count; int ascending;
  float* loss; float* n_pair_f32; int64_t* n_pair; float* dlogits; int64_t* row_pairs;
  RowMap rm; u32 out_chunk;       // blocked input rows; floats per output chunk (0 = dlogits[B]), see rn_pairwise_args
};

// Cost of one J-block (32 negatives) of a virtual block; a fast tile costs 8 * kCostUnit.  Range R1 of an I-block (the
// rows of the group that began before the block) is fast tiles plus the blocks that hold a label-level boundary; range
// R2 (groups that begin inside the block) is general tiles (a warp that finishes early costs little -- its SMSP
// neighbours speed up -- while a late one runs alone and cannot fill the SFU pipe, so the estimates lean to the
// pessimistic side).
constexpr u32 kCostUnit = 4;                 // cost units per eighth of a fast tile (sub-eighth resolution for the averages below)
constexpr u32 kCostFast = 8 * kCostUnit;
// A range R1 of nt J-blocks crosses up to `lv` label-level boundaries of its group; a block that holds one is scored in
// 2-3 passes (runs_tile) or as a general tile.  Short ranges (medium-sized groups) consist mostly of such blocks, long
// ones hardly notice: cost per block = 8 + cstr * min(nt, lv) / nt (rounded up).  lv = 0: label levels play no role.
__device__ __forceinline__ u32 vcost(u32 v, u32 nt, u32 cgen, u32 cstr, u32 lv) {
  if (v & 1u) return cgen;
  return kCostFast + (cstr * min(nt, lv) + nt - 1u) / max(nt, 1u);
}

// Position `pos` of the cost line -> (virtual block, J-block, eighth).  s_pi: cost prefix over the virtual blocks,
// s_jn: first J-block | J-block count << 16 of every virtual block.
__device__ __forceinline__ void resolve_pos(u32 pos, u32 tot, u32 nvb, const u32* s_pi, const u32* s_jn, u32 cgen, u32 csw,
                                            u32 cstr, u32 lv, u32& v, u32& jb, u32& e) {
  if (pos >= tot) { v = nvb; jb = 0; e = 0; return; }
  v = last_le(s_pi, 0, nvb, pos);
  u32 o = pos - s_pi[v];
  o = o > csw ? o - csw : 0u;                      // (the first csw units of a block stand for its row loads / flush)
  const u32 c = vcost(v, s_jn[v] >> 16, cgen, cstr, lv), q = o / c;
  jb = (s_jn[v] & 0xFFFFu) + q;
  e = ((o - q * c) * 8u) / c;
}

// Partition of the pair kernel's cost line, computed ahead of it (see HeadsTail::partition): boundary r of the
// wpr = (pair-kernel warps of the grid) equal pieces as (virtual block, J-block, eighth), and per virtual block its first
// J-block | J-block count << 16.
struct PrePart {
  uint4* bnd; u32* jn;        // [wpr + 1] (v, jb, e, -), [nvb]
  u32 wpr, pair_grid;         // pieces = warps of the pair kernel's grid; its CTAs
  u32 cost_gen, cost_switch, cost_straddle, lv_cost;
  int on;                     // the arrays exist and the batch's cost prefix fits k_seg's shared memory
};

// ---- heads tail of k_seg ---------------------------------------------------------------------------------
// Work units: an I-block (64 sorted rows) x up to C consecutive J-blocks (32 sorted rows each) of its J range.
// Record = (I-block, first J-block | J-block count << 24).
struct HeadsTail {
  static constexpr bool kFast = true;
  PairParams P;
  uint2* aj; float *ss, *sy, *swp, *swn, *gacc, *lossrow; u32 *cnt, *perm, *sgrp;
  uint2* blk; uint2* units; u32 nib; u64* cprim; const u32* pgid;
  u32 target_units;          // work-list granularity target (units of <= C J-blocks)
  PrePart pp;
  unsigned long long* g64;   // deterministic mode: fixed-point gradient accumulators (zeroed here)

  // ---- counting path (group_count.cuh): count -> offsets -> scatter; false = outside its menu (radix path instead) ----
  // One row's scatter: its sorted position from the group record, its sorted columns, its negative range and the
  // J ranges of its I-block.
  __device__ __forceinline__ void scatter_row(const SegParams& S, u32 i, u32 slot, u32 meta, float s, float y, float wp,
                                              float wn) const {
    const GRec* r = S.rec + slot;
    const u32 li = meta >> 29, rank = meta & 0x1FFFFFFFu;
    const u32 a = r->base, pre = r->cnt[li], own = r->own;
    const float wocc = r->wocc;
    // global mode: inside every label level the rank's OWN rows come first, then the other ranks' rows.  An own row pairs
    // with all rows below its level, wherever they live; a remote row is only ever a negative here (every pair is scored
    // by the rank that owns its positive row, so no two ranks need to agree on the order of the rows)
    // Small groups are not split: ONE rank (a hash of the key picks it) takes all their rows as positives, the others
    // none -- otherwise every rank would walk every I-block of every small group for a few own rows each.
    const bool remote = S.rm.Bl && (i / S.rm.Bl != (u32)P.part_rank) && slot <= S.capmask;
    const u32 pos = a + (remote ? S.rec2[slot].cr[li] : pre) + rank;
    const u32 n = (own == 2u || (remote && own != 1u)) ? 0u : pre;      // rows of the group below this row's level
    aj[pos] = make_uint2(a, n);
    // (a NaN label pairs with nothing, but its row shares I-blocks with rows that do: under label-gain weights the tile
    // multiplies a zero row weight by (y_i - y_ref), and 0 * NaN would poison the block's sums)
    ss[pos] = s; sy[pos] = (y != y) ? 0.f : (P.gain2 == 2 ? (float)li : (P.gain2 ? exp2f(y) : y));      // (counting path: every label is on the level menu)
    if (P.rw_pos) swp[pos] = wp;
    if (P.rw_neg) swn[pos] = wn;
    gacc[pos] = 0.f; perm[pos] = i; sgrp[pos] = slot;
    lossrow[pos] = P.dyn_count ? 0.f : wocc;                 // (non-dynamic: the row's occurrence weight, read by k_pair)
    if (P.dyn_count) cnt[pos] = 0;
    else if (P.row_pairs) P.row_pairs[i] = (int64_t)n;
    if (P.lambda) { cnt[pos] = 0; g64[pos] = 0ull; }         // (the LambdaRank pre-pass's per-group accumulators)
  }

  // J ranges of the I-blocks a group touches, from its geometry alone (offsets phase: base, level starts `pre`, level
  // counts `c`, rows `tot`).  Range = (group start, level start of the group's LAST row inside the block): the levels
  // ascend along the positions, so that row has the longest negative range.  The block in which the group begins gets
  // it as R2 (hull with the other groups that begin there: atomicMax on (~lo, hi), zero = empty), every later block
  // as R1 (exactly one group can reach into a block from the left: plain store).
  // Geometry of one group as the offsets phase knows it.  The group's rows are laid out level by level; inside a level the
  // rank's OWN rows come first, then (global mode) the other ranks' rows.  ls[q]: start of level q relative to the group's
  // base, co[q]: own rows of the level (they sit at [ls[q], ls[q] + co[q])).  The negative range of an own row of level q
  // is [base, base + ls[q]); remote rows have none.
  struct Geo { u32 base, tot; u32 ls[kLevels], co[kLevels]; };
  __device__ __forceinline__ void block_range(u32 b, u32 b0, const Geo& g) const {
    // rows of the group inside I-block b: relative positions [x0, x1]; the longest range belongs to the highest level
    // that has an own row in there
    const u32 x0 = b == b0 ? 0u : b * kIB - g.base, x1 = min(g.tot, (b + 1u) * kIB - g.base) - 1u;
    u32 hi = 0;
#pragma unroll
    for (int q = 0; q < kLevels; ++q) if (g.co[q] && g.ls[q] <= x1 && g.ls[q] + g.co[q] > x0) hi = g.ls[q];
    if (!hi) return;
    if (b == b0) { atomicMax(&blk[2 * b + 1].x, ~g.base); atomicMax(&blk[2 * b + 1].y, g.base + hi); }
    else blk[2 * b] = make_uint2(~g.base, g.base + hi);
  }

  // Cost prefix of the pair kernel's work line and the boundaries of its equal pieces (what k_pair's prologue
  // otherwise computes in every CTA), by helper h of H: every helper scans the costs, resolves a slice of the boundaries.
  __device__ __forceinline__ void partition(const SegParams& S, u32* smem, u32 h, u32 H) const {
    const u32 nvb = 2 * nib;
    u32* s_pi = smem;                          // [nvb + 1] cost prefix
    u32* s_jn = smem + nvb + 1;                // [nvb] first J-block | J-block count << 16
    u32* s_sc = s_jn + nvb;                    // scan partials
    u32 msum = 0;
    // (four virtual blocks per round: their range loads are all in flight together)
    for (u32 v0 = 0; v0 < nvb; v0 += 4 * kSegThreads) {
      uint2 r[4], rl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 v = v0 + k * kSegThreads + threadIdx.x;
        r[k] = make_uint2(0, 0); rl[k] = make_uint2(0, 0);
        if (v < nvb) { r[k] = blk[v]; if (v & 1u) rl[k] = blk[v - 1]; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 v = v0 + k * kSegThreads + threadIdx.x;
        if (v < nvb) {
          u32 lo = 0, hi = 0;
          if (r[k].y > ~r[k].x) { lo = ~r[k].x >> 5; hi = (r[k].y + 31) >> 5; }
          if ((v & 1u) && rl[k].y > ~rl[k].x) lo = max(lo, (rl[k].y + 31) >> 5);      // (as vblock_tiles)
          const u32 nt = hi > lo ? hi - lo : 0u;
          const u32 jn = nt ? (lo | (nt << 16)) : 0u;
          s_jn[v] = jn;
          if (h == 0) pp.jn[v] = jn;
          s_pi[v] = nt ? nt * vcost(v, nt, pp.cost_gen, pp.cost_straddle, pp.lv_cost) + pp.cost_switch : 0u;
          msum += nt;
        }
      }
    }
    __syncthreads();
    if (S.dbgts && threadIdx.x == 0) S.dbgts[(size_t)5 * gridDim.x + blockIdx.x] = globaltimer();
    block_excl_scan(s_pi, nvb, s_sc);
    if (S.dbgts && threadIdx.x == 0) S.dbgts[(size_t)6 * gridDim.x + blockIdx.x] = globaltimer();
    const u32 tot = s_pi[nvb];
    const u32 q = tot / pp.wpr, rem = tot - q * pp.wpr;
    for (u32 r = h * kSegThreads + threadIdx.x; r <= pp.wpr; r += H * kSegThreads) {
      const u32 pos = q * r + (u32)(((u64)rem * r) / pp.wpr);
      u32 v, jb, e;
      resolve_pos(pos, tot, nvb, s_pi, s_jn, pp.cost_gen, pp.cost_switch, pp.cost_straddle, pp.lv_cost, v, jb, e);
      pp.bnd[r] = make_uint4(v, jb, e, 0u);
    }
    if (h == 0) {
      msum = warp_sum(msum);
      if (lane_id() == 0 && msum) atomicAdd((u64*)&S.ctl->n_tiles, (u64)msum);
      if (threadIdx.x == 0) { S.ctl->n_units = tot; S.ctl->unit_c = 0; }
    }
  }

  __device__ __forceinline__ bool count_run(const SegParams& S, u32* smem, u32& epoch) const {
    constexpr u32 kLoc = 2 * kGTile;
    u32* sm_tab = smem;                                           // [kLoc] representative thread of the key hashed here
    u64* sm_key = reinterpret_cast<u64*>(smem + kLoc);            // [kGTile] keys of the tile
    u32* sm_gslot = smem + kLoc + 2 * kGTile;                     // [kGTile] record of the representative's key
    u32* sm_cnt = sm_gslot + kGTile;                              // [kGTile][kLevels] rows per (representative, level) -> rank base
    u32* sm_misc = sm_cnt + kGTile * kLevels;                     // [0] created records, [1] rows that cannot pair, [2] their rank base
    u64* sm_red = reinterpret_cast<u64*>(sm_misc + 4);            // [kSegWarps]
    Ctl* ctl = S.ctl;
    GRec* rec = S.rec;
    const u32 B = S.B, tid = threadIdx.x, ln = lane_id(), w = tid >> 5;
    const u32 ntile = (B + kGTile - 1) / kGTile;
    const bool single = ntile <= gridDim.x;                       // one tile per CTA: the rows stay in registers
    const u32 trash = S.capmask + 1u;
    const bool dyn = P.dyn_count != 0;
    // global mode (blocked rows): tiles never straddle two ranks' blocks (block_rows is a multiple of the tile); a tile
    // of another rank's rows counts into the REMOTE half of the groups' second records
    const bool glob = S.rm.Bl != 0;
    GRec2* rec2 = S.rec2;
    bool bad;
    u32 k_slot = 0, k_meta = 0; float k_s = 0.f, k_y = 0.f, k_wp = 1.f, k_wn = 1.f;
    // (one tile per CTA: the thread that created a record keeps it for the offsets phase -- no list to read back)
    u32 k_cslot = 0; bool k_created = false, k_trash = false;
    bad = false;
    // ---- count --------------------------------------------------------------------------------------------------
    for (u32 t = blockIdx.x; t < ntile; t += gridDim.x) {
      const u32 i = t * kGTile + tid;
      const bool in = i < B;
      const bool tile_remote = glob && (t * kGTile / S.rm.Bl != (u32)P.part_rank);
      const size_t i8 = S.rm.i8(i), i4 = S.rm.i4(i), i1 = S.rm.i1(i);
      // all loads of the row first (independent, one round trip)
      const u64 key = in ? (u64)S.keys[i8] : 0ull;
      const float y = in ? S.labels[i4] : 0.f;
      const bool okb = in && (S.row_ok ? S.row_ok[i1] != 0 : true);
      if (single) {
        k_s = in ? P.logits[i4] : 0.f;
        if (P.rw_pos) k_wp = in ? P.rw_pos[i4] : 1.f;
        if (P.rw_neg) k_wn = in ? P.rw_neg[i4] : 1.f;
        k_y = y;
      }
      float wp = 1.f;
      if (!single && P.rw_pos) wp = in ? P.rw_pos[i4] : 1.f; else wp = k_wp;
      sm_tab[tid] = kEmpty; sm_tab[tid + kGTile] = kEmpty;
#pragma unroll
      for (int q = 0; q < kLevels; q += 4) *reinterpret_cast<uint4*>(sm_cnt + tid * kLevels + q) = make_uint4(0, 0, 0, 0);
      if (tid < 4) sm_misc[tid] = 0;
      sm_key[tid] = key;
      const bool ok = okb && !(y != y);
      int li = 0;
      if (ok) {
        if (!label_level(y, li)) { bad = true; li = 0; }
        if (P.rw_pos && !(wp > 0.f)) bad = true;              // PW:193 C = W > 0 removes the row's pairs: counts are no longer position arithmetic
      }
      __syncthreads();
      // tile-local grouping: the first thread to claim a key's cell represents it
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
      }
      // rank inside (representative, level) of the tile
      u32 lr = 0;
      if (ok) lr = atomicAdd(&sm_cnt[rep * kLevels + li], 1u);
      else if (in) lr = atomicAdd(&sm_misc[1], 1u);
      // representatives: find / create the group's record
      bool created = false; u32 slot = 0;
      const bool isrep = ok && rep == tid;
      if (isrep) { slot = grec_insert(rec, S.capmask, h, key, i, created, &ctl->err); sm_gslot[tid] = slot; }
      if (created) S.glist[(size_t)t * kGTile + atomicAdd(&sm_misc[0], 1u)] = slot;
      if (single) { k_created = created; k_cslot = slot; }
      __syncthreads();
      // reserve the tile's ranks inside every (group, level): one atomicAdd each, all in flight together
      if (isrep) {
        u32 c[kLevels];
#pragma unroll
        for (int q = 0; q < kLevels; ++q) c[q] = sm_cnt[tid * kLevels + q];
        u32* tgt = glob ? (tile_remote ? rec2[slot].cr : rec2[slot].cl) : rec[slot].cnt;
#pragma unroll
        for (int q = 0; q < kLevels; ++q) if (c[q]) c[q] = atomicAdd(tgt + q, c[q]);
#pragma unroll
        for (int q = 0; q < kLevels; ++q) sm_cnt[tid * kLevels + q] = c[q];
      }
      if (tid == 0 && sm_misc[1]) {
        // rows that cannot pair (row_ok = 0, NaN label): one group of one level behind the table
        if (atomicCAS(&rec[trash].rep1, 0u, 1u) == 0u) { S.glist[(size_t)t * kGTile + atomicAdd(&sm_misc[0], 1u)] = trash; k_trash = true; }
        sm_misc[2] = atomicAdd(&rec[trash].cnt[0], sm_misc[1]);
      }
      __syncthreads();
      if (tid == 0) S.gcount[t] = sm_misc[0];
      if (in) {
        const u32 rslot = ok ? sm_gslot[rep] : trash;
        const u32 rank = (ok ? sm_cnt[rep * kLevels + li] : sm_misc[2]) + lr;
        const u32 meta = ((u32)li << 29) | rank;
        if (single) { k_slot = rslot; k_meta = meta; }
        else { S.rslot[i] = rslot; S.rmeta[i] = meta; }
      }
      __syncthreads();
    }
    if (__syncthreads_or(bad) && tid == 0) st_relaxed(&ctl->fallback, 1u);
    auto dbg = [&](u32 ph) { if (S.dbgts && tid == 0) S.dbgts[(size_t)ph * gridDim.x + blockIdx.x] = globaltimer(); };
    stamp(ctl, 1); dbg(0);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    stamp(ctl, 2); dbg(1);
    if (ld_relaxed(&ctl->fallback)) return false;
    // ---- offsets: one thread per record created by this CTA's tiles (all tiles' lists walked as one) ---------------------
    u64 npsum = 0;
    u32* sm_toff = smem;                     // [tiles of this CTA + 1] prefix of their list lengths (<= 64 tiles; else tile by tile)
    const u32 my_tiles = blockIdx.x < ntile ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const bool flat = my_tiles <= 64u;
    u32 total = 0;
    const bool trash_round = single && __syncthreads_or(k_trash);       // (this CTA created the record of the rows that cannot pair)
    if (flat && !single) {
      if (tid == 0) {
        u32 acc = 0;
        for (u32 q = 0; q < my_tiles; ++q) { sm_toff[q] = acc; acc += S.gcount[blockIdx.x + q * gridDim.x]; }
        sm_toff[my_tiles] = acc;
      }
      __syncthreads();
      total = sm_toff[my_tiles];
    }
    constexpr u32 kWideBlocks = 128, kWideCap = 48, kWideWords = 4 + 2 * kLevels;
    u32* sm_qn = smem + 100;                 // queue of very long groups (behind sm_toff): count, entries
    u32* sm_q = smem + 128;
    if (tid == 0) *sm_qn = 0;
    __syncthreads();
    const u32 nrounds = single ? (trash_round ? 2u : 1u) : flat ? (total + kSegThreads - 1) / kSegThreads : my_tiles;
    for (u32 rd = 0; rd < nrounds; ++rd) {
      // (single: round 0 = every thread's own created record, round 1 = the record of the unpairable rows;
      //  flat: round rd handles items [rd * 512, +512) of the concatenated lists; else round rd = tile rd, lists <= 512 long)
      u32 t = 0, k = 0; bool act = false;
      if (single) {
        act = rd == 0 ? k_created : (tid == 0 && k_trash);
      } else if (flat) {
        const u32 item = rd * kSegThreads + tid;
        if (item < total) {
          u32 q = 0;
          while (sm_toff[q + 1] <= item) ++q;
          t = blockIdx.x + q * gridDim.x; k = item - sm_toff[q]; act = true;
        }
      } else {
        t = blockIdx.x + rd * gridDim.x; k = tid; act = k < S.gcount[t];
      }
      {
        u32 slot = 0; u64 pairs = 0; u32 own = 0;
        Geo g{};                    // (inactive lanes: an empty group)
        u32 cl[kLevels];            // own rows per level (the split of every level: own rows first, then the others')
#pragma unroll
        for (int q = 0; q < kLevels; ++q) cl[q] = 0;
        if (act) {
          slot = single ? (rd == 0 ? k_cslot : trash) : S.glist[(size_t)t * kGTile + k];
          u32 cr[kLevels];
          if (glob && slot <= S.capmask) {
            const uint4 a0 = *reinterpret_cast<const uint4*>(rec2[slot].cl), a1 = *reinterpret_cast<const uint4*>(rec2[slot].cl + 4);
            const uint4 b0 = *reinterpret_cast<const uint4*>(rec2[slot].cr), b1 = *reinterpret_cast<const uint4*>(rec2[slot].cr + 4);
            cl[0] = a0.x; cl[1] = a0.y; cl[2] = a0.z; cl[3] = a0.w; cl[4] = a1.x; cl[5] = a1.y; cl[6] = a1.z; cl[7] = a1.w;
            cr[0] = b0.x; cr[1] = b0.y; cr[2] = b0.z; cr[3] = b0.w; cr[4] = b1.x; cr[5] = b1.y; cr[6] = b1.z; cr[7] = b1.w;
          } else {
            const uint4 a0 = *reinterpret_cast<const uint4*>(rec[slot].cnt), a1 = *reinterpret_cast<const uint4*>(rec[slot].cnt + 4);
            cl[0] = a0.x; cl[1] = a0.y; cl[2] = a0.z; cl[3] = a0.w; cl[4] = a1.x; cl[5] = a1.y; cl[6] = a1.z; cl[7] = a1.w;
#pragma unroll
            for (int q = 0; q < kLevels; ++q) cr[q] = 0;
          }
#pragma unroll
          for (int q = 0; q < kLevels; ++q) {
            g.ls[q] = g.tot; g.co[q] = cl[q];
            pairs += (u64)(cl[q] + cr[q]) * g.tot;             // pairs of a row = ALL rows of the group below its level
            g.tot += cl[q] + cr[q];
          }
          if (glob && slot <= S.capmask && g.tot < kOwnRows) {
            // a small group is scored whole by one rank: every rank picks the same one from the key
            own = ((u32)(mix64(0xD6E8FEB86659FD93ull ^ rec[slot].key) >> 32) % (u32)P.part_count == (u32)P.part_rank) ? 1u : 2u;
#pragma unroll
            for (int q = 0; q < kLevels; ++q) g.co[q] = own == 1u ? cl[q] + cr[q] : 0u;
          }
        }
        u32 inc = g.tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
        const u32 wtot = __shfl_sync(0xFFFFFFFFu, inc, 31);
        u32 wbase = 0;
        if (ln == 0 && wtot) wbase = atomicAdd(&ctl->cursor, wtot);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        g.base = wbase + inc - g.tot;
        // J ranges of the I-blocks the group touches: up to three blocks by its own thread, longer groups by the warp
        u32 b0 = 0, nblk = 0;
        if (act && g.tot) {
          b0 = g.base / kIB; nblk = (g.base + g.tot - 1u) / kIB - b0 + 1u;
          for (u32 b = b0; b < b0 + min(nblk, 3u); ++b) block_range(b, b0, g);
        }
        // Very long groups (the head of a Zipf batch; in the global mode tens of thousands of rows) go to a queue the
        // whole CTA works off after the round: their creators' first rows all sit in the batch's first tile, so one
        // warp of one CTA used to write the ranges of all of them, one after the other, with the grid waiting
        bool queued = false;
        if (act && nblk > kWideBlocks) {
          const u32 qi = atomicAdd(sm_qn, 1u);
          if (qi < kWideCap) {
            u32* e = sm_q + qi * kWideWords;
            e[0] = g.base; e[1] = g.tot; e[2] = b0; e[3] = nblk;
#pragma unroll
            for (int q = 0; q < kLevels; ++q) { e[4 + q] = g.ls[q]; e[4 + kLevels + q] = g.co[q]; }
            queued = true;
          }
        }
        u32 big = __ballot_sync(0xFFFFFFFFu, nblk > 3u && !queued);
        while (big) {
          const int src = __ffs(big) - 1; big &= big - 1;
          Geo h;
          h.base = __shfl_sync(0xFFFFFFFFu, g.base, src); h.tot = __shfl_sync(0xFFFFFFFFu, g.tot, src);
#pragma unroll
          for (int q = 0; q < kLevels; ++q) { h.ls[q] = __shfl_sync(0xFFFFFFFFu, g.ls[q], src); h.co[q] = __shfl_sync(0xFFFFFFFFu, g.co[q], src); }
          const u32 gb0 = __shfl_sync(0xFFFFFFFFu, b0, src), gnb = __shfl_sync(0xFFFFFFFFu, nblk, src);
          for (u32 b = gb0 + 3u + ln; b < gb0 + gnb; b += 32) block_range(b, gb0, h);
        }
        if (act) {
          GRec* r = rec + slot;
          r->base = g.base;
          *reinterpret_cast<uint4*>(r->cnt) = make_uint4(g.ls[0], g.ls[1], g.ls[2], g.ls[3]);
          *reinterpret_cast<uint4*>(r->cnt + 4) = make_uint4(g.ls[4], g.ls[5], g.ls[6], g.ls[7]);
          if (glob && slot <= S.capmask) {       // (start of the remote rows of every level; the split of the group)
            *reinterpret_cast<uint4*>(rec2[slot].cr) = make_uint4(g.ls[0] + cl[0], g.ls[1] + cl[1], g.ls[2] + cl[2], g.ls[3] + cl[3]);
            *reinterpret_cast<uint4*>(rec2[slot].cr + 4) = make_uint4(g.ls[4] + cl[4], g.ls[5] + cl[5], g.ls[6] + cl[6], g.ls[7] + cl[7]);
            r->own = own;
          }
          if (!dyn) {
            r->npair = pairs;
            r->wocc = (P.power != 0.f && pairs) ? ((P.power == 1.0f) ? (float)pairs : powf((float)pairs, P.power)) : 0.f;   // PW:147-149
            npsum += pairs;
          }
        }
      }
      // the queued long groups: every thread of the CTA takes I-blocks of each
      __syncthreads();
      {
        const u32 nq = min(*sm_qn, kWideCap);
        for (u32 qi = 0; qi < nq; ++qi) {
          const u32* e = sm_q + qi * kWideWords;
          Geo h;
          h.base = e[0]; h.tot = e[1];
          const u32 gb0 = e[2], gnb = e[3];
#pragma unroll
          for (int q = 0; q < kLevels; ++q) { h.ls[q] = e[4 + q]; h.co[q] = e[4 + kLevels + q]; }
          for (u32 b = gb0 + 3u + tid; b < gb0 + gnb; b += kSegThreads) block_range(b, gb0, h);
        }
      }
      __syncthreads();
      if (tid == 0) *sm_qn = 0;
      __syncthreads();
    }
    if (!dyn) {
      npsum = warp_sum(npsum);
      if (ln == 0) sm_red[w] = npsum;
      __syncthreads();
      if (tid == 0) {
        u64 tsum = 0;
        for (int q = 0; q < kSegWarps; ++q) tsum += sm_red[q];
        if (tsum) atomicAdd(&ctl->n_pair, tsum);
      }
    }
    if (blockIdx.x == 0 && tid == 0) ctl->path = 1;
    stamp(ctl, 3); dbg(2);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    stamp(ctl, 4); dbg(3);
    // ---- scatter (+ the pair kernel's partition, beside it) ----------------------------------------------------------
    // The J ranges are complete: CTAs without rows (the grid is always full) work out the partition of the pair kernel's
    // cost line while the others scatter; without enough spare CTAs everybody takes a slice after its rows.
    const u32 spare = gridDim.x > ntile ? gridDim.x - ntile : 0u;
    if (pp.on && spare >= 4u && blockIdx.x >= ntile) { partition(S, smem, blockIdx.x - ntile, spare); dbg(4); return true; }
    if (single) {
      const u32 i = blockIdx.x * kGTile + tid;
      if (i < B) scatter_row(S, i, k_slot, k_meta, k_s, k_y, k_wp, k_wn);
    } else {
      // (four tiles per round: the loads of their rows and of the rows' records are all in flight together)
      for (u32 t0 = blockIdx.x; t0 < ntile; t0 += 4 * gridDim.x) {
        u32 ri[4], rs[4], rmt[4]; float fs[4], fy[4], fwp[4], fwn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const u32 t = t0 + q * gridDim.x, i = t * kGTile + tid;
          ri[q] = (t < ntile && i < B) ? i : kEmpty;
          if (ri[q] != kEmpty) {
            const size_t i4 = S.rm.i4(i);
            rs[q] = S.rslot[i]; rmt[q] = S.rmeta[i]; fs[q] = P.logits[i4]; fy[q] = S.labels[i4];
            fwp[q] = P.rw_pos ? P.rw_pos[i4] : 1.f; fwn[q] = P.rw_neg ? P.rw_neg[i4] : 1.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (ri[q] != kEmpty) scatter_row(S, ri[q], rs[q], rmt[q], fs[q], fy[q], fwp[q], fwn[q]);
      }
    }
    if (pp.on && spare < 4u) { __syncthreads(); partition(S, smem, blockIdx.x, gridDim.x); }
    stamp(ctl, 17); dbg(4);
    return true;
  }

  __device__ __forceinline__ void run(const SegParams& S, const Plan& pl, const u64* key, const u32* val, u32* smem, u32& epoch) const {
    u32* sm_scan = smem;                       // [kSegWarps][2]
    u32* sm_carry = smem + 2 * kSegWarps;      // [2]
    u32* sm_wj = smem + 36;                    // [kSegWarps][4]
    u64* sm_red = reinterpret_cast<u64*>(smem + 104);  // [kSegWarps]
    Ctl* ctl = S.ctl;
    const u32 B = S.B, ln = lane_id(), w = threadIdx.x >> 5;
    const bool count_now = !P.dyn_count;
    const u32 nchunks = (B + kSegThreads - 1) / kSegThreads;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const u32 t0 = c * kSegThreads, p = t0 + threadIdx.x;
      const bool in = p < B;
      const u64 k = in ? key[p] : ~0ull;
      const u64 kp = (in && p > 0) ? key[p - 1] : ~k;
      const u32 gid = (u32)(k >> pl.labbits);
      const bool head = in && (p == 0 || gid != (u32)(kp >> pl.labbits));
      const bool lvl = in && (head || k != kp);
      if (w == 0) { const u32 lb = coop_lower_bound(key, t0, pl.labbits); if (ln == 0) sm_carry[0] = lb; }
      if (w == 1) { const u32 lb = coop_lower_bound(key, t0, 0); if (ln == 0) sm_carry[1] = lb; }
      u32 xa = head ? p + 1 : 0, xl = lvl ? p + 1 : 0;
      block_maxscan2(xa, xl, sm_scan);
      const u32 a = xa ? xa - 1 : sm_carry[0];
      const u32 l = xl ? xl - 1 : sm_carry[1];
      u32 n = in ? l - a : 0;
      const u32 row = in ? val[p] : 0;
      if (in) {
        float wp = 1.f, wn = 1.f;
        const size_t ro = P.rm.i4(row);
        // PW:193  C = W > 0: a non-positive positive-side weight removes the row's pairs -- unless a negative-side
        // weight may flip the sign back (then the tile decides pair by pair)
        if (P.rw_pos) { wp = P.rw_pos[ro]; if (!P.rw_neg && !(wp > 0.f)) n = 0; }
        if (P.rw_neg) wn = P.rw_neg[ro];
        aj[p] = make_uint2(a, n);
        ss[p] = P.logits[ro];
        {
          const float yv = P.labels[ro];
          float yc = (yv != yv) ? 0.f : (P.gain2 == 1 ? exp2f(yv) : yv);      // (NaN labels pair with nothing; see scatter_row)
          if (P.gain2 == 2 && yv == yv) {                                      // RN_LABEL_LUT: the level; off the menu = a failed call
            int lv;
            if (!label_level(yv, lv)) { lv = 0; if (!S.row_ok || S.row_ok[row]) atomicOr(&ctl->err, 8u); }      // (not with blocked rows)
            yc = (float)lv;
          }
          sy[p] = yc;
        }
        if (P.rw_pos) swp[p] = wp;
        if (P.rw_neg) swn[p] = wn;
        gacc[p] = 0.f; perm[p] = row;
        lossrow[p] = 0.f;
        if (P.dyn_count) cnt[p] = 0;
        if (P.lambda) { cnt[p] = 0; g64[p] = 0ull; }           // (see scatter_row)
        if (P.det) {
          // bound of one pair weight, for the scale of the fixed-point accumulators: max |w_i| and max |y| over the rows
          g64[p] = 0ull;
          const u32 wb = __float_as_uint(fabsf(wp)), yb = __float_as_uint(fabsf(sy[p]));
          const u32 wm = __reduce_max_sync(__activemask(), wb), ym = __reduce_max_sync(__activemask(), yb);
          if (ln == (u32)(__ffs(__activemask()) - 1)) { atomicMax(&ctl->det_w, wm); atomicMax(&ctl->det_y, ym); }
        }
      }
      // J ranges needed by this warp's 32 rows (two warps make one I-block): R1 = the range of the rows whose group
      // began before the I-block (one group: it can reach far to the left), R2 = the hull of the ranges of the rows
      // whose group begins inside the I-block (those lie inside the I-block's own 64 positions).  Keeping them apart
      // keeps the gap between them -- the top label level of a big group -- out of the work list.
      {
        const u32 p0 = p & ~(u32)(kIB - 1);
        const bool r1 = n && a < p0, r2 = n && a >= p0;
        const u32 r1lo = warp_min(r1 ? a : 0xFFFFFFFFu), r1hi = warp_max(r1 ? a + n : 0u);
        const u32 r2lo = warp_min(r2 ? a : 0xFFFFFFFFu), r2hi = warp_max(r2 ? a + n : 0u);
        if (ln == 0) { sm_wj[4 * w] = r1lo; sm_wj[4 * w + 1] = r1hi; sm_wj[4 * w + 2] = r2lo; sm_wj[4 * w + 3] = r2hi; }
      }
      // exact counts (position arithmetic): per row, per PRIMARY group (PW:286-289), total
      u32 cn = 0;
      // (merged first phase: ids are table slots; the id `cap` collects the rows that cannot pair)
      const u32 pgi = (in && !(S.merged && gid > S.capmask)) ? (S.K > 1 ? pgid[row] : gid) : kEmpty;
      if (in) sgrp[p] = pgi;                    // index of the row's occurrence count: k_pair weights the row by c_h^power
      if (count_now) {
        cn = n;
        if (in && P.row_pairs) P.row_pairs[row] = (int64_t)cn;
        const u32 pg = cn ? pgi : kEmpty;
        const u32 m = __match_any_sync(0xFFFFFFFFu, pg);
        const u32 tot = __reduce_add_sync(m, cn);
        if (pg != kEmpty && ln == (u32)(__ffs(m) - 1)) atomicAdd(cprim + pg, (u64)tot);
      }
      const u64 s = warp_sum((u64)cn);
      if (ln == 0) sm_red[w] = s;
      __syncthreads();
      if (!(w & 1) && ln == 0) {
        const u32 ib = (t0 + w * 32) / kIB;
        if (ib < nib) {
          blk[2 * ib] = make_uint2(~min(sm_wj[4 * w], sm_wj[4 * w + 4]), max(sm_wj[4 * w + 1], sm_wj[4 * w + 5]));
          blk[2 * ib + 1] = make_uint2(~min(sm_wj[4 * w + 2], sm_wj[4 * w + 6]), max(sm_wj[4 * w + 3], sm_wj[4 * w + 7]));
        }
      }
      if (threadIdx.x == 0 && count_now) {
        u64 t = 0;
        for (int q = 0; q < kSegWarps; ++q) t += sm_red[q];
        if (t) atomicAdd(&ctl->n_pair, t);
      }
      __syncthreads();
    }
    stamp(ctl, 17);
    // Batches up to 524288 rows: k_pair partitions the work itself (cost prefixes in shared memory), the kernel ends here
    // without another grid barrier.  Larger batches: explicit unit records.
    if (nib <= kMaxNibS) return;
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    stamp(ctl, 18);
    // ---- work list: every CTA scans the per-I-block tile counts (redundantly, it is ~nib/512 block scans) and emits
    //      the unit records of its own I-blocks ------------------------------------------------------------------
    const u32 nvb = 2 * nib;
    auto ntile = [&](u32 v) -> u32 { u32 jf; return vblock_tiles(blk, v, jf); };
    u64 m = 0;
    for (u32 b = threadIdx.x; b < nvb; b += kSegThreads) m += ntile(b);
    m = warp_sum(m);
    __syncthreads();
    if (ln == 0) sm_red[w] = m;
    __syncthreads();
    u64 M = 0;
    for (int q = 0; q < kSegWarps; ++q) M += sm_red[q];
    u32 C = (u32)((M + target_units - 1) / target_units);
    C = C < 1 ? 1 : (C > kMaxUnitC ? kMaxUnitC : C);
    u32* s_scan = smem;            // [kSegWarps]
    u32* s_carry = smem + kSegWarps;
    __syncthreads();
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (u32 b0 = 0; b0 < nvb; b0 += kSegThreads) {
      const u32 b = b0 + threadIdx.x;
      u32 jfirst = 0;
      const u32 nt = b < nvb ? vblock_tiles(blk, b, jfirst) : 0u;
      const u32 v = (nt + C - 1) / C;
      u32 inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
      if (ln == 31) s_scan[w] = inc;
      __syncthreads();
      u32 off = *s_carry;
      for (u32 q = 0; q < w; ++q) off += s_scan[q];
      if (v && (b % gridDim.x) == blockIdx.x) {
        uint2* dst = units + (off + inc - v);
        for (u32 q = 0; q < v; ++q) dst[q] = make_uint2(b >> 1, (jfirst + q * C) | (min(C, nt - q * C) << 24));
      }
      __syncthreads();
      if (threadIdx.x == kSegThreads - 1) *s_carry = off + inc;
      __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->n_units = *s_carry; ctl->unit_c = C; ctl->n_tiles = M; }
  }
};

// ---- the pair kernel ------------------------------------------------------------------------------
struct KpArgs {
  const uint2* aj; const float *ss, *sy, *swp, *swn; const uint2* units; const uint2* blk; u32 nib; u32 target_units;
  u32 cost_gen;              // cost of a J-block of a range R2 in eighths of a fast tile (see vcost)
  u32 cost_switch;           // fixed cost of a virtual block (row loads, flush of the row accumulators), same unit
  u32 cost_straddle, cost_levels;   // extra cost of a block that holds a label-level boundary; boundaries per range (see vcost)
  float *gacc, *lossrow; u32* cnt; const u32* perm; Ctl* ctl;
  const float* lut;          // RN_LABEL_LUT: the caller's 8 x 8 level weight table (copied to shared memory by every CTA)
  const u32* sgrp;           // per sorted row: index of its (primary) group's pair total -- cprim[] or, counting path, rec[]
  u64* cprim;
  // counting path (group_count.cuh): the call ran without k_init; the records its count phase created are listed per
  // 512-row tile and zeroed again by the last phase of this kernel
  int fast; GRec* rec; u32 rec2_off; const u32* glist; const u32* gcount; u32 ngt;     // (rec2_off: global mode only, else 0)
  uint2* blk_w;              // (the J ranges are zeroed again as well)
  const uint4* bnd; const u32* jn; int pre_on;     // partition computed by k_seg's helper CTAs (PrePart)
  // global mode with a score- / weight-dependent pair set: the per-row pair counts are partial (this rank's negatives).
  // k_pair publishes them by ORIGINAL row (xcnt_out) and stops; after a cross-rank barrier k_fin_dyn sums the ranks'
  // arrays (xcnt_peer, peer-mapped) and finishes: counts -> occurrence weights -> gradient (PW:197-203, 282-291)
  u32* xcnt_out; const u32* xcnt_peer[8]; int xworld; u32* xsum;      // xsum: local [B] scratch for the summed counts
  // deterministic mode: per-row gradient sums as 64-bit fixed point (integer adds commute: any arrival order gives the
  // same bits), one loss partial per CTA summed in CTA order
  unsigned long long* g64; double* lpart;
  u64* dbgbuf;      // RN_PAIR_DEBUG: per warp {first segment start, loop exit, busy cycles, segments | general tiles << 32}
};

// Occurrence weight c_h^power of a row's primary group (PW:147-149, 285-290).  Both rows of a pair share the group, so
// the weight is applied per ROW, not per pair: to the row's loss sum when its segment ends and to its gradient in the
// final pass.  The counts are final before k_pair starts.
__device__ __forceinline__ float occ_pow(u64 ch, float power) {
  return ch ? ((power == 1.0f) ? (float)ch : powf((float)ch, power)) : 0.f;
}

// ---- end of a call: leave the arena clean for the next one (persistent arenas run without k_init) --------------------
// Zero the group records created by the count phase of tiles t0, t0 + tstep, ... (four threads per 64-byte record).
__device__ __forceinline__ void clean_records(const KpArgs& A, u32 t0, u32 tstep) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (u32 t = t0; t < A.ngt; t += tstep) {
    const u32 n = A.gcount[t];
    for (u32 k = threadIdx.x; k < 4 * n; k += blockDim.x) {
      const u32 slot = A.glist[(size_t)t * kGTile + (k >> 2)];
      reinterpret_cast<uint4*>(A.rec + slot)[k & 3u] = z;
      // (the second records sit rec2_off records behind the first ones: one base pointer for both -- the pair kernel
      // runs at its register limit, and a second pointer kept for this tail costs it hundreds of bytes of spills)
      if (A.rec2_off) reinterpret_cast<uint4*>(A.rec + A.rec2_off + slot)[k & 3u] = z;
    }
  }
}

// Positive-side rows of an I-block (two per lane) and the first J block of a segment: loaded one segment ahead.
struct UnitRows {
  uint2 an0, an1; float si0, si1, yi0, yi1, wp0, wp1; float sjm, yjm, wnjm;
  float di0, di1;          // M_LAMBDA: the rows' rank discounts
};

// One piece of work of a warp: I-block b x J-blocks [jb0, jb1); of the first J-block only the rotation-step eighths
// from ea on, of the last one only those below eb (a whole tile is [0, 8)).
struct Seg { u32 b, jb0, jb1, ea, eb; };

template <int MODE>
__device__ __forceinline__ void load_unit_rows(const KpArgs& A, u32 B, u32 b, u32 jb0, u32 ln, UnitRows& r) {
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, LAMBDA = MODE & M_LAMBDA;
  const u32 pi0 = b * kIB + ln, pi1 = pi0 + 32;
  r.an0 = make_uint2(0, 0); r.an1 = make_uint2(0, 0);
  r.si0 = r.si1 = r.yi0 = r.yi1 = 0.f; r.wp0 = r.wp1 = 1.f; r.di0 = r.di1 = 0.f;
  if (LAMBDA) { if (pi0 < B) r.di0 = A.swn[pi0]; if (pi1 < B) r.di1 = A.swn[pi1]; }
  if (pi0 < B) { r.an0 = A.aj[pi0]; r.si0 = A.ss[pi0]; if (DIFF) r.yi0 = A.sy[pi0]; if (HASW && A.swp) r.wp0 = A.swp[pi0]; }
  if (pi1 < B) { r.an1 = A.aj[pi1]; r.si1 = A.ss[pi1]; if (DIFF) r.yi1 = A.sy[pi1]; if (HASW && A.swp) r.wp1 = A.swp[pi1]; }
  const u32 pjm = jb0 * 32 + ln;
  r.sjm = pjm < B ? A.ss[pjm] : 0.f; r.yjm = 0.f; r.wnjm = 1.f;
  if (DIFF) r.yjm = pjm < B ? A.sy[pjm] : 0.f;
  if (RWN || LAMBDA) r.wnjm = pjm < B ? A.swn[pjm] : 0.f;
}

// A block whose negatives span several label levels of ONE group (label-gain weights): every row covers a level of the
// block entirely or not at all, so the block is the sum of one product-form pass per level, each with the other levels'
// negatives switched off (F = 0).  Up to three levels (a general tile costs about four fast ones); false = not taken.
template <bool LUT>
__device__ __forceinline__ bool runs_tile(const float* lut, const bool in0, const bool in1, const u32 lo0, const u32 lo1, const u32 hi0,
                                          const u32 hi1, const float si0, const float si1, const float yi0, const float yi1,
                                          const float wp0, const float wp1, const u32 pjm, const float sjm, const float yjm,
                                          const bool jin, const u32 smin, const u32 j0, const float c, const int ts,
                                          const int te, float& li0, float& li1, float& gi0, float& gi1, float& accj) {
  const u32 ln = lane_id();
  const u32 gl = __reduce_min_sync(0xFFFFFFFFu, min(in0 ? lo0 : 0xFFFFFFFFu, in1 ? lo1 : 0xFFFFFFFFu));
  const u32 gh = __reduce_max_sync(0xFFFFFFFFu, max(in0 ? lo0 : 0u, in1 ? lo1 : 0u));
  if (gl != gh) return false;                                        // rows of several groups touch the block
  const float yprev = __shfl_up_sync(0xFFFFFFFFu, yjm, 1);
  const u32 heads = __ballot_sync(0xFFFFFFFFu, jin && (pjm == smin || yjm != yprev));
  if (__popc(heads) > 3) return false;
  const float mref = __shfl_sync(0xFFFFFFFFu, sjm, smin - j0);
  const float aj = (sjm - mref) * c, a0 = (si0 - mref) * c, a1 = (si1 - mref) * c;
  if (!__all_sync(0xFFFFFFFFu, (!jin || fabsf(aj) <= kProdRange) && (!in0 || fabsf(a0) <= kProdRange) &&
                               (!in1 || fabsf(a1) <= kProdRange))) return false;
  const float Fm = jin ? mufu_ex2(aj) : 0.f;
  const float E0 = in0 ? mufu_ex2(-a0) : 0.f, E1 = in1 ? mufu_ex2(-a1) : 0.f;
  u32 rem = __ballot_sync(0xFFFFFFFFu, jin);
  while (rem) {
    const int lead = __ffs(rem) - 1;
    const float yv = __shfl_sync(0xFFFFFFFFu, yjm, lead);
    const u32 runm = __ballot_sync(0xFFFFFFFFu, jin && yjm == yv) | (1u << lead);
    const u32 rs = j0 + (u32)(__ffs(runm) - 1), re = j0 + 32u - (u32)__clz(runm);
    const float wv0 = (in0 && lo0 <= rs && re <= hi0) ? wp0 * label_weight<LUT>(lut, yi0, yv) : 0.f;
    const float wv1 = (in1 && lo1 <= rs && re <= hi1) ? wp1 * label_weight<LUT>(lut, yi1, yv) : 0.f;
    tile_prod<true>(E0, E1, wv0, wv1, ((runm >> ln) & 1u) ? Fm : 0.f, li0, li1, gi0, gi1, accj, ts, te);
    rem &= ~runm;
  }
  return true;
}

__device__ __forceinline__ void dyn_finalize(const PairParams& P, const KpArgs& A, const u64* cprim, u32 cstride, u32 nvb,
                                             u32& epoch2, u64* red_u, double* red_d);

// ---- LambdaRank weights (RN_LABEL_LAMBDA) -----------------------------------------------------------------
// W_ij = |delta NDCG_ij| = (2^y_i - 2^y_j) * |D(r_i) - D(r_j)| / IDCG_g for y_i > y_j, with D(r) = 1 / log2(1 + r), r_i the
// 1-based rank of row i among the rows of its group by score (descending; ties by original row) and IDCG_g the DCG of
// the group's labels in descending order (gain 2^y - 1).  The weights are constants of the step (PW:270 stops the gradient
// through every weight).  This pre-pass of the pair kernel works the per-row parts out on the sorted columns:
//   1  rows per group (one atomic per group and warp; the per-group accumulators, indexed by the group's first sorted
//      position, were zeroed by the segmentation kernel's scatter)
//   2  one warp per row: its rank by a scan of the group's scores (coalesced; rows are dealt to the CTAs round-robin, so
//      the rows of a big group are spread evenly over all SMs), its discount D(r) -> the negative-side weight column;
//      then one thread per row: its term of the ideal DCG -- the labels ascend along the sorted positions, so the row's
//      ideal rank is its distance from the group's end -- summed in double precision
//   3  row weight = rw_pos (or 1) / IDCG -> the positive-side weight column
// Discounts are rounded to float32 from double precision (|D_i - D_j| of neighbouring ranks of a long group cancels
// six digits: the oracle rounds the same way).  Groups of one label level (and the rows that cannot pair) are skipped.
constexpr u32 kLambdaBarriers = 3;
static __device__ __noinline__ void lambda_prepare(const PairParams& P, const KpArgs& A) {
  Ctl* ctl = A.ctl;
  const u32 B = P.B, ln = lane_id();
  const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  u32* gsz = A.cnt;                                     // (the dynamic modes' pair counters: free here)
  double* idcg = reinterpret_cast<double*>(A.g64);      // (the deterministic mode's accumulators: free here)
  float* sd = const_cast<float*>(A.swn);
  float* swp = const_cast<float*>(A.swp);
  u32 epoch = 0;
  for (u32 p0 = 0; p0 < B; p0 += gthreads) {
    const u32 p = p0 + gtid;
    const u32 a = p < B ? A.aj[p].x : kEmpty;
    const u32 m = __match_any_sync(0xFFFFFFFFu, a);
    if (a != kEmpty && ln == (u32)(__ffs(m) - 1)) atomicAdd(gsz + a, (u32)__popc(m));
  }
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
  stamp(ctl, 13);
  u32* sdu = reinterpret_cast<u32*>(sd);                // (the rank count first; the discount is worked out row-parallel in step 3)
  // (row p -> CTA p % grid: every CTA takes the same share of every big group's rows -- whole 32-row windows per CTA left
  // some SMs with two windows of the biggest group and others with one; a CTA's warps still scan the same few groups)
  for (u32 p = blockIdx.x + gridDim.x * (threadIdx.x >> 5); p < B; p += gridDim.x * (blockDim.x >> 5)) {
    const u32 a = A.aj[p].x, sz = gsz[a];
    if (A.aj[a + sz - 1u].y == 0u) { if (ln == 0) sdu[p] = 0xFFFFFFFFu; continue; }      // one label level: the group has no pairs
    const float sp = A.ss[p];
    // Rows ranked above this one = scores greater + tied scores of earlier rows.  The scan is bound by its instruction
    // count (L1 serves 94 % of its loads): whole rounds of 256 positions run without bounds tests, eight loads in flight,
    // two compares and two predicated adds per load; the tie-break by row index is a second scan taken only when a tie
    // exists (the first version -- bounds test, compare and tie branch per element -- took 17 instructions per load).
    u32 c = 0, e = 0;
    const u32 end = a + sz;
    u32 base = a;
    for (; base + 256u <= end; base += 256u) {
      const float* ps = A.ss + base + ln;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ps[32 * k];
#pragma unroll
      for (int k = 0; k < 8; ++k) { c += v[k] > sp ? 1u : 0u; e += v[k] == sp ? 1u : 0u; }
    }
    for (u32 q = base + ln; q < end; q += 32u) { const float v = A.ss[q]; c += v > sp ? 1u : 0u; e += v == sp ? 1u : 0u; }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    e = __reduce_add_sync(0xFFFFFFFFu, e);
    if (e > 1u) {                                       // (the row itself is one of the ties)
      const u32 rp = A.perm[p];
      u32 t = 0;
      for (u32 q = a + ln; q < end; q += 32u) t += (A.ss[q] == sp && A.perm[q] < rp) ? 1u : 0u;
      c += __reduce_add_sync(0xFFFFFFFFu, t);
    }
    if (ln == 0) sdu[p] = c;                            // r = c + 1
  }
  stamp(ctl, 14);
  // (the ideal DCG: one thread per row; consecutive lanes hold consecutive rows of a group -- a segmented warp sum, then
  // one atomic per group and warp instead of 7 000 on the accumulator of a big group)
  for (u32 p0 = 0; p0 < B; p0 += gthreads) {
    const u32 p = p0 + gtid;
    u32 a = kEmpty; double t = 0.0;
    if (p < B) {
      a = A.aj[p].x;
      const u32 sz = gsz[a];
      if (A.aj[a + sz - 1u].y != 0u) t = ((double)A.sy[p] - 1.0) / log2(1.0 + (double)(a + sz - p));      // (the label column holds 2^y)
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tv = __shfl_down_sync(0xFFFFFFFFu, t, o);
      const u32 av = __shfl_down_sync(0xFFFFFFFFu, a, o);
      if (ln + (u32)o < 32u && av == a) t += tv;
    }
    const u32 ap = __shfl_up_sync(0xFFFFFFFFu, a, 1);
    if (a != kEmpty && (ln == 0 || ap != a) && t != 0.0) atomicAdd(idcg + a, t);
  }
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
  stamp(ctl, 15);
  for (u32 p = gtid; p < B; p += gthreads) {
    const double t = idcg[A.aj[p].x];
    swp[p] = t > 0.0 ? (P.rw_pos ? swp[p] : 1.0f) * (float)(1.0 / t) : 0.f;
    const u32 c = sdu[p];
    sd[p] = c == 0xFFFFFFFFu ? 0.f : (float)(1.0 / log2(2.0 + (double)c));      // D(r) = 1 / log2(1 + r), r = c + 1
  }
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
}

// DET: deterministic mode (fixed-point gradient accumulators, ordered loss partials) -- a separate instantiation, so that
// the default kernel carries none of it through the pair loop.
// HINGE: the pair loss is the hinge max(0, margin - x) (tile_hinge / tile_general<.., HINGE>) instead of the logistic loss.
template <int MODE, bool DET = false, bool HINGE = false>
__global__ void __launch_bounds__(kPairThreads) k_pair(PairParams P, KpArgs A) {
  // Work = the (I-block x J-block) tiles of the staircase, laid out on a COST LINE: virtual block after virtual block
  // (the two J ranges of every I-block), J-block after J-block, each with the cost vcost() estimates for it (a fast tile
  // = 8 eighths; more for general tiles and for blocks that hold a label-level boundary).  For B <= 524288 the kernel partitions the line STATICALLY: the costs are prefix-summed
  // over the virtual blocks (shared memory, every CTA redundantly) and every warp of the grid takes ONE contiguous
  // piece of exactly the same length, cut at a granularity of one eighth of a tile (4 of the 32 rotation steps).
  // Pieces are dealt to the warps interleaved over the SMs (piece r -> CTA r % grid), so every SM gets the same mix.
  // No tickets, no atomics for scheduling, a few row loads per warp (the rows of the next I-block of a piece are
  // fetched while the current one is scored).  Larger batches use explicit unit records (built by k_seg) that the
  // warps of a CTA take dynamically.
  extern __shared__ __align__(16) u32 dsm[];
  __shared__ u32 s_cnt;
  __shared__ u64 red_u[kPairWarps];
  __shared__ double red_d[kPairWarps];
  __shared__ u32 s_scan[kPairWarps + 2];
  __shared__ u32 s_bnd[6 * kPairWarps];
  constexpr bool HASW = MODE & M_HASW, DIFF = MODE & M_DIFF, RWN = MODE & M_RWN, WRONG = MODE & M_WRONG, LUT = MODE & M_LUT;
  constexpr bool LAMBDA = MODE & M_LAMBDA;
  constexpr bool DYN = RWN || WRONG;
  Ctl* ctl = A.ctl;
  const u32 ln = lane_id();
  const u32 B = P.B;
  // RN_LABEL_LUT: the level weight table in shared memory.  Only entries with l_i > l_j are ever a kept pair's weight;
  // they must be finite and > 0 (else the pair set would not be [y_i > y_j]: the call fails, loss = NaN), the others are
  // zeroed (a row without overlap multiplies its zero weight with whatever entry its level addresses).
  __shared__ float s_lut[LUT ? 64 : 1];
  bool lut_bad = false;
  if (LUT) {
    if (threadIdx.x < 64) {
      float v = A.lut[threadIdx.x];
      if ((threadIdx.x >> 3) > (threadIdx.x & 7u)) {
        if (!(v > 0.f) || v > 3.0e38f) { v = 0.f; lut_bad = true; }
      } else v = 0.f;
      s_lut[threadIdx.x] = v;
    }
    __syncthreads();
  }
  grid_dep_wait();
  if (LUT && lut_bad) atomicOr(&ctl->err, 8u);            // (the control block is the previous kernel's until here)
  if (LAMBDA) lambda_prepare(P, A);                       // rank discounts and 1 / IDCG of every row (kLambdaBarriers grid barriers)
  stamp(ctl, 20);
  if (threadIdx.x == 0) s_cnt = 0;
  const bool own_list = A.nib <= kMaxNibS;
  const float fold_power = DYN ? 0.f : P.power;
  const u32 nvb = 2 * A.nib;                 // virtual blocks: the two J ranges of every I-block
  u32* s_pi = dsm;                           // [nvb + 1]
  u32* s_jn = s_pi + nvb + 1;                // [nvb]
  // static partition: this warp's piece [cur, end) of the cost line; explicit list: U units
  u32 cb = 0, cj = 0, ce = 0, zb = 0, zj = 0, ze = 0;
  bool st_done = true;
  u32 U = 0;
  const u32 lv_cost = DIFF ? A.cost_levels : 0u;     // (label levels only split blocks under label-gain weights)
  // Occurrence weight of every row for the final pass: the chain cnt -> cprim -> pow is two dependent round trips, on
  // the critical path if the last warp to leave the pair loop has to walk it in front of the grid barrier.  It is
  // walked here instead, its loads spread over the prologue, and the weight parked in the (otherwise unused) lossrow
  // column: each thread reads back what it wrote itself.
  // (Counting path: the offsets phase of k_seg knows every group's total and the scatter phase has already parked the
  // weights; the pair totals then live in the group records.)
  const u32 gtid0 = blockIdx.x * blockDim.x + threadIdx.x;
  // (only `counted` stays live through the pair loop -- the kernel runs at its register limit; the radix path's totals
  // are A.cprim with stride 1, the counting path needs no totals inside the loop)
  const bool counted = A.fast && ld_relaxed(&ctl->fallback) == 0u;
  const bool park = !DYN && fold_power != 0.f && !counted;
  u32 park_pg = kEmpty; u64 park_ch = 0;
  if (park && gtid0 < B) park_pg = A.sgrp[gtid0];
  // Deterministic mode: the scale of the fixed-point gradient accumulators.  A row's sum is bounded by (pairs it takes part
  // in <= 2 B) x (largest pair weight W); the scale leaves that bound just inside 62 bits, i.e. a resolution of
  // W x 2^-(61 - log2 2B) per contribution (W x 2^-44 at B = 65536).
  // (kept in shared memory: the pair loop has no registers to spare)
  __shared__ double s_dscale[2];
  if (DET && threadIdx.x == 0) {
    float W = fmaxf(__uint_as_float(ld_relaxed(&ctl->det_w)), 1.0e-30f);
    if (DIFF) W *= 2.0f * fmaxf(__uint_as_float(ld_relaxed(&ctl->det_y)), 1.0e-30f);
    int ew; frexpf(W, &ew);                                  // W < 2^ew
    const int e = 61 - (33 - __clz(B)) - ew;                 // (2 B < 2^(33 - clz B))
    s_dscale[0] = ldexp(1.0, e); s_dscale[1] = ldexp(1.0, -e);
  }
  auto gadd = [&](u32 p, float v) {
    if (DET) atomicAdd(A.g64 + p, (unsigned long long)__double2ll_rn((double)v * s_dscale[0]));
    else atomicAdd(A.gacc + p, v);
  };
  const bool pre = counted && A.pre_on && own_list;     // the partition was worked out by k_seg's helper CTAs
  const u32* jn_ = pre ? A.jn : s_jn;
  if (pre) {
    const u32 wq = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    const u32 r = wq * gridDim.x + blockIdx.x;
    const uint4 ba = A.bnd[r], bz = A.bnd[r + 1];
    cb = ba.x; cj = ba.y; ce = ba.z; zb = bz.x; zj = bz.y; ze = bz.z;
    st_done = !(cb < zb || (cb == zb && (cj < zj || (cj == zj && ce < ze))));
  } else if (own_list) {
    u32 msum = 0;
    // (one 16-byte load per I-block = both of its J ranges; four I-blocks per round, their loads in flight together)
    for (u32 b0 = 0; b0 < A.nib; b0 += 4 * kPairThreads) {
      uint4 rr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 b = b0 + k * kPairThreads + threadIdx.x;
        rr[k] = b < A.nib ? reinterpret_cast<const uint4*>(A.blk)[b] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 b = b0 + k * kPairThreads + threadIdx.x;
        if (b < A.nib) {
          // (as vblock_tiles: R1 = (x, y), R2 = (z, w) trimmed so that no tile is listed twice)
          u32 lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
          if (rr[k].y > ~rr[k].x) { lo1 = ~rr[k].x >> 5; hi1 = (rr[k].y + 31) >> 5; }
          if (rr[k].w > ~rr[k].z) { lo2 = ~rr[k].z >> 5; hi2 = (rr[k].w + 31) >> 5; }
          if (hi1 > lo1) lo2 = max(lo2, hi1);
          const u32 nt1 = hi1 > lo1 ? hi1 - lo1 : 0u, nt2 = hi2 > lo2 ? hi2 - lo2 : 0u;
          s_jn[2 * b] = nt1 ? (lo1 | (nt1 << 16)) : 0u;
          s_jn[2 * b + 1] = nt2 ? (lo2 | (nt2 << 16)) : 0u;
          s_pi[2 * b] = nt1 ? nt1 * vcost(2 * b, nt1, A.cost_gen, A.cost_straddle, lv_cost) + A.cost_switch : 0u;
          s_pi[2 * b + 1] = nt2 ? nt2 * vcost(2 * b + 1, nt2, A.cost_gen, A.cost_straddle, lv_cost) + A.cost_switch : 0u;
          msum += nt1 + nt2;
        }
      }
    }
    __syncthreads();
    stamp(ctl, 11);
    if (park_pg != kEmpty) park_ch = A.cprim[park_pg];
    block_excl_scan(s_pi, nvb, s_scan);
    stamp(ctl, 12);
    const u32 tot = s_pi[nvb];
    // this rank's share [r0, r1) of the cost line, one equal piece per warp; the 2 x 32 piece boundaries of this CTA's
    // warps are resolved by 64 threads at once (s_bnd: [64][v, jb, e])
    u32 r0 = 0, r1 = tot;
    if (P.part_count > 1 && !counted) {      // (counting path: the ranks split the pairs by the owner of the negative row instead)
      r0 = (u32)(((u64)tot * (u32)P.part_rank) / (u32)P.part_count);
      r1 = (u32)(((u64)tot * ((u32)P.part_rank + 1)) / (u32)P.part_count);
    }
    if (threadIdx.x < 2 * kPairWarps) {
      const u32 wpr = gridDim.x * kPairWarps;                       // pieces per rank (< 65536)
      const u32 r = (threadIdx.x >> 1) * gridDim.x + blockIdx.x + (threadIdx.x & 1u);
      const u32 len = r1 - r0, q = len / wpr, rem = len - q * wpr;  // len * r / wpr without 64-bit division
      const u32 pos = r0 + q * r + (rem * r) / wpr;
      u32 v, jb, e;
      resolve_pos(pos, tot, nvb, s_pi, s_jn, A.cost_gen, A.cost_switch, A.cost_straddle, lv_cost, v, jb, e);
      s_bnd[3 * threadIdx.x] = v; s_bnd[3 * threadIdx.x + 1] = jb; s_bnd[3 * threadIdx.x + 2] = e;
    }
    __syncthreads();
    {
      // (warp-uniform values are passed through shfl so that the compiler keeps them in uniform registers)
      const u32 wq = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
      const u32* bd = s_bnd + 6 * wq;
      cb = bd[0]; cj = bd[1]; ce = bd[2]; zb = bd[3]; zj = bd[4]; ze = bd[5];
      st_done = !(cb < zb || (cb == zb && (cj < zj || (cj == zj && ce < ze))));
    }
    if (blockIdx.x == 0) {
      msum = warp_sum(msum);
      if (ln == 0 && msum) atomicAdd((u64*)&ctl->n_tiles, (u64)msum);
      if (threadIdx.x == 0) { ctl->n_units = tot; ctl->unit_c = 0; }
    }
  } else {
    U = ld_relaxed(&ctl->n_units);
    if (park_pg != kEmpty) park_ch = A.cprim[park_pg];
    __syncthreads();
  }
  if (park) {
    if (gtid0 < B) A.lossrow[gtid0] = occ_pow(park_ch, fold_power);
    for (u32 p = gtid0 + gridDim.x * blockDim.x; p < B; p += gridDim.x * blockDim.x) {
      const u32 pg = A.sgrp[p];
      A.lossrow[p] = pg != kEmpty ? occ_pow(A.cprim[pg], fold_power) : 0.f;
    }
  }
  if (!DYN && P.focal_w != 0.f) {
    // fused focal term, value: sum over ALL rows (original order; the mean is over B), one atomic per CTA
    double fs = 0.0;
    for (u32 i = gtid0; i < B; i += gridDim.x * blockDim.x) fs += (double)focal_row(P.logits[i], P.labels[i], P.focal_alpha, P.focal_gamma, P.focal_stop).x;
    fs = warp_sum(fs);
    if (ln == 0) red_d[threadIdx.x >> 5] = fs;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int q = 0; q < kPairWarps; ++q) t += red_d[q];
      atomicAdd(&ctl->acc_d[3], t);
    }
    __syncthreads();
  }
  stamp(ctl, 16);
  double lsum = 0.0;
  {
    const u32 u_begin = (u32)(((u64)U * (u32)P.part_rank) / (u32)P.part_count);
    const u32 u_end = (u32)(((u64)U * ((u32)P.part_rank + 1)) / (u32)P.part_count);
    const u32 n_mine = (u_end - u_begin > blockIdx.x) ? (u_end - u_begin - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const float c = P.c_log2;
    // debug tallies (per warp, flushed once): busy cycles, segments, fast / general tiles, eighths
    u64 d_busy = 0; u32 d_units = 0, d_fast = 0, d_gen = 0, d_eighths = 0; u64 d_gencyc = 0, d_fastcyc = 0;
    const bool tally = (P.debug & 3) != 0;          // (bits 1, 2: per-warp records / grid tallies; the higher bits only switch code paths)
    const u64 d_start = tally ? globaltimer() : 0;
#ifdef RN_TRACE
    // dev build: where a warp's time goes (clock cycles): taking segments, per-J-block preamble, tiles, RED + loop end, flush
    u64 tr_take = 0, tr_pre = 0, tr_tile = 0, tr_post = 0, tr_flush = 0; long long tr_t = 0;
#define TR_BEGIN() (tr_t = clock64())
#define TR_ADD(acc) do { const long long now_ = clock64(); acc += (u64)(now_ - tr_t); tr_t = now_; } while (0)
#else
#define TR_BEGIN() do {} while (0)
#define TR_ADD(acc) do {} while (0)
#endif
    auto take = [&](Seg& sg) -> bool {
      if (own_list) {
        // next I-block segment of this warp's piece
        for (;;) {
          if (st_done || cb >= nvb) return false;
          if (cb > zb || (cb == zb && (cj > zj || (cj == zj && ce >= ze)))) { st_done = true; return false; }
          const u32 jn = jn_[cb], jend = (jn & 0xFFFFu) + (jn >> 16);
          if (cj >= jend) {                       // (also skips virtual blocks without tiles)
            // the next virtual block that has tiles, 32 candidates per look: a piece can hold a long run of empty blocks
            // (rows of pairless small groups cost nothing on the cost line), and walking them one dependent load at a
            // time made single warps the last to reach the barrier
            u32 nb = cb + 1u;
            while (nb < nvb && nb <= zb) {
              const u32 idx = nb + ln;
              const u32 m = __ballot_sync(0xFFFFFFFFu, idx >= nvb || (jn_[idx] >> 16) != 0u);
              if (m) { nb += (u32)__ffs(m) - 1u; break; }
              nb += 32u;
            }
            cb = nb; cj = cb < nvb ? (jn_[cb] & 0xFFFFu) : 0u; ce = 0;
            continue;
          }
          sg.b = cb >> 1; sg.jb0 = cj; sg.ea = ce;
          if (cb == zb) { sg.jb1 = ze ? zj + 1 : zj; sg.eb = ze ? ze : 8u; st_done = true; }
          else { sg.jb1 = jend; sg.eb = 8u; ++cb; cj = cb < nvb ? (jn_[cb] & 0xFFFFu) : 0u; ce = 0; }
          sg.b = __shfl_sync(0xFFFFFFFFu, sg.b, 0); sg.jb0 = __shfl_sync(0xFFFFFFFFu, sg.jb0, 0);
          sg.jb1 = __shfl_sync(0xFFFFFFFFu, sg.jb1, 0);
          const u32 ee = __shfl_sync(0xFFFFFFFFu, sg.ea | (sg.eb << 4), 0);
          sg.ea = ee & 15u; sg.eb = ee >> 4;
          return true;
        }
      }
      u32 q = 0;
      if (ln == 0) q = atomicAdd(&s_cnt, 1u);
      q = __shfl_sync(0xFFFFFFFFu, q, 0);
      if (q >= n_mine) return false;
      const u32 idx = q * gridDim.x + blockIdx.x;
      const uint2 rec = A.units[P.ascending ? u_begin + idx : u_end - 1u - idx];
      sg.b = rec.x; sg.jb0 = rec.y & 0xFFFFFFu; sg.jb1 = sg.jb0 + (rec.y >> 24); sg.ea = 0; sg.eb = 8u;
      return true;
    };
    Seg sg, sg_n; UnitRows R, Rn;
    bool have = take(sg);
    const bool fold = fold_power != 0.f;
    const bool use_prod = !(P.debug & 16);        // (debug bit 16: fast tiles without the product form)
    u32 occ_a = 0xFFFFFFFFu; float occ_w = 0.f;   // group start the cached occurrence weight belongs to
    if (have) load_unit_rows<MODE>(A, B, sg.b, sg.jb0, ln, R);
    while (have) {
      const long long d_t0 = tally ? clock64() : 0;
      TR_BEGIN();
      const bool have_n = take(sg_n);
      if (have_n) load_unit_rows<MODE>(A, B, sg_n.b, sg_n.jb0, ln, Rn);      // in flight while this segment is scored
      TR_ADD(tr_take);
      const u32 b = sg.b, jb0 = sg.jb0, jb1 = sg.jb1;
      const u32 pi0 = b * kIB + ln, pi1 = pi0 + 32;
      const uint2 an0 = R.an0, an1 = R.an1;
      const float si0 = R.si0, si1 = R.si1, yi0 = R.yi0, yi1 = R.yi1, wp0 = R.wp0, wp1 = R.wp1;
      const float di0 = R.di0, di1 = R.di1;
      const u32 lo0 = an0.x, hi0 = an0.x + an0.y, lo1 = an1.x, hi1 = an1.x + an1.y;
      float li0 = 0.f, li1 = 0.f, gi0 = 0.f, gi1 = 0.f; u32 cnt0 = 0, cnt1 = 0;
      u32 pjm = jb0 * 32 + ln;
      float sjm = R.sjm, yjm = R.yjm, wnjm = R.wnjm;
      // J blocks inside [seg_lomax, seg_himin) are covered entirely by every row of the I-block that has pairs
      const bool act0 = an0.y != 0, act1 = an1.y != 0;
      const u32 seg_lomax = __reduce_max_sync(0xFFFFFFFFu, max(act0 ? lo0 : 0u, act1 ? lo1 : 0u));
      const u32 seg_himin = __reduce_min_sync(0xFFFFFFFFu, min(act0 ? hi0 : 0xFFFFFFFFu, act1 ? hi1 : 0xFFFFFFFFu));
      for (u32 jb = jb0; jb < jb1; ++jb) {
        // prefetch the next J-block while this one is being scored
        const u32 pjn = pjm + 32;
        const bool more = (jb + 1 < jb1) && pjn < B;
        float sjn = more ? A.ss[pjn] : 0.f, yjn = 0.f, wnjn = 1.f;
        if (DIFF) yjn = more ? A.sy[pjn] : 0.f;
        if (RWN || LAMBDA) wnjn = more ? A.swn[pjn] : 0.f;
        float accj = 0.f;
        const u32 j0 = jb * 32;
        const int ts = (jb == jb0) ? 4 * (int)sg.ea : 0, te = (jb + 1 == jb1) ? 4 * (int)sg.eb : 32;   // rotation steps
        const bool part = (te - ts) != 32 || (P.debug & 8);      // (debug bit 8: every fast tile through the looped variant)
        // Overlap of every row's negative range with this J block.  Row ranges start at a group start and end at
        // a level start of the same group, so a row covers a (group, level) run of the block entirely or not at
        // all: if all rows that touch the block share ONE overlap [s, e) and (label weights) its labels are one
        // value, the block is scored by the fast tile with out-of-range negatives replaced by a sentinel score
        // (x = +inf: e = 0, sigma = 0, factor 1 in the product) and rows that do not touch it weighted 0.
        bool in0, in1; u32 smin, smax, emin, emax;
        if (j0 >= seg_lomax && j0 + 32 <= seg_himin && seg_himin != 0xFFFFFFFFu && !(P.debug & 64)) {   // (debug bit 64: always the general overlap test)
          // the usual block of a big group: every row of the I-block that has pairs covers all 32 negatives
          in0 = act0; in1 = act1; smin = smax = j0; emin = emax = j0 + 32;
        } else {
          const u32 s0 = max(lo0, j0), e0 = min(hi0, j0 + 32), s1 = max(lo1, j0), e1 = min(hi1, j0 + 32);
          in0 = s0 < e0; in1 = s1 < e1;
          smin = __reduce_min_sync(0xFFFFFFFFu, min(in0 ? s0 : 0xFFFFFFFFu, in1 ? s1 : 0xFFFFFFFFu));
          smax = __reduce_max_sync(0xFFFFFFFFu, max(in0 ? s0 : 0u, in1 ? s1 : 0u));
          emin = __reduce_min_sync(0xFFFFFFFFu, min(in0 ? e0 : 0xFFFFFFFFu, in1 ? e1 : 0xFFFFFFFFu));
          emax = __reduce_max_sync(0xFFFFFFFFu, max(in0 ? e0 : 0u, in1 ? e1 : 0u));
        }
        const bool any_in = smin != 0xFFFFFFFFu;
        const bool jin = pjm >= smin && pjm < emax;                  // this lane's negative is inside the overlap
        bool fast = any_in && smin == smax && emin == emax && !RWN && !WRONG && !LAMBDA;
        float yref = 0.f;
        if (DIFF && fast) {
          yref = __shfl_sync(0xFFFFFFFFu, yjm, smin - j0);
          fast = __all_sync(0xFFFFFFFFu, !jin || yjm == yref);
        }
        if (tally) { if (fast) ++d_fast; else if (any_in) ++d_gen; d_eighths += (u32)(te - ts) >> 2; }
        const long long d_g0 = tally ? clock64() : 0;
        TR_ADD(tr_pre);
        if (fast) {
          float wv0 = in0 ? wp0 : 0.f, wv1 = in1 ? wp1 : 0.f;
          if (DIFF) { wv0 *= label_weight<LUT>(s_lut, yi0, yref); wv1 *= label_weight<LUT>(s_lut, yi1, yref); }
          const float sje = jin ? sjm : -3.0e38f;
          if (HINGE) {
            if (part) tile_hinge<true>(si0, si1, wv0, wv1, sje, c, P.margin, li0, li1, gi0, gi1, accj, ts, te);
            else      tile_hinge<false>(si0, si1, wv0, wv1, sje, c, P.margin, li0, li1, gi0, gi1, accj);
          } else {
          // product form (one SFU operation per pair) when the scores of the tile lie within 2^+-kProdRange of a
          // reference score -- here the first negative of the overlap; otherwise exp(-|x|) per pair
          const float mref = __shfl_sync(0xFFFFFFFFu, sjm, smin - j0);
          const float aj = (sjm - mref) * c, a0 = (si0 - mref) * c, a1 = (si1 - mref) * c;
          const bool prod = use_prod && __all_sync(0xFFFFFFFFu, (!jin || fabsf(aj) <= kProdRange) &&
                                                   (!in0 || fabsf(a0) <= kProdRange) && (!in1 || fabsf(a1) <= kProdRange));
          if (prod) {
            const float Fm = jin ? mufu_ex2(aj) : 0.f;
            const float E0 = in0 ? mufu_ex2(-a0) : 0.f, E1 = in1 ? mufu_ex2(-a1) : 0.f;
            if (part) tile_prod<true>(E0, E1, wv0, wv1, Fm, li0, li1, gi0, gi1, accj, ts, te);
            else      tile_prod<false>(E0, E1, wv0, wv1, Fm, li0, li1, gi0, gi1, accj);
          } else {
            if (part) tile_fast<true, true>(si0, si1, wv0, wv1, sje, c, li0, li1, gi0, gi1, accj, ts, te);
            else      tile_fast<true>(si0, si1, wv0, wv1, sje, c, li0, li1, gi0, gi1, accj);
          }
          }
        } else if (!HINGE && any_in && DIFF && !RWN && !WRONG && !LAMBDA && use_prod && !(P.debug & 32) &&      // (debug bit 32: without the level passes)
                   runs_tile<LUT>(s_lut, in0, in1, lo0, lo1, hi0, hi1, si0, si1, yi0, yi1, wp0, wp1, pjm, sjm, yjm, jin, smin, j0, c, ts, te,
                             li0, li1, gi0, gi1, accj)) {
          // (scored as up to three product-form passes, one per label level of the negatives)
        } else if (any_in) {
          const bool full0 = __all_sync(0xFFFFFFFFu, lo0 <= j0 && j0 + 32 <= hi0);
          const bool full1 = __all_sync(0xFFFFFFFFu, lo1 <= j0 && j0 + 32 <= hi1);
          const bool any0 = __any_sync(0xFFFFFFFFu, in0);
          const bool any1 = __any_sync(0xFFFFFFFFu, in1);
          // product form of the general tile (two SFU operations per pair) under the same range test as tile_prod
          bool gprod = false; float Fg = 0.f, Eg0 = 0.f, Eg1 = 0.f;
          if (!HINGE && !WRONG && use_prod && !(P.debug & 128)) {                  // (debug bit 128: general tiles with exp per pair)
            const float mref = __shfl_sync(0xFFFFFFFFu, sjm, smin - j0);
            const float aj = (sjm - mref) * c, a0 = (si0 - mref) * c, a1 = (si1 - mref) * c;
            gprod = __all_sync(0xFFFFFFFFu, (!jin || fabsf(aj) <= kProdRange) && (!in0 || fabsf(a0) <= kProdRange) &&
                                            (!in1 || fabsf(a1) <= kProdRange));
            if (gprod) { Fg = jin ? mufu_ex2(aj) : 0.f; Eg0 = in0 ? mufu_ex2(-a0) : 0.f; Eg1 = in1 ? mufu_ex2(-a1) : 0.f; }
          }
          if (gprod) {
            if (any0) {
              if (full0) tile_general_prod<MODE, true>(Eg0, yi0, wp0, lo0, hi0, pjm, Fg, yjm, wnjm, ts, te, li0, gi0, cnt0, accj, s_lut, di0);
              else       tile_general_prod<MODE, false>(Eg0, yi0, wp0, lo0, hi0, pjm, Fg, yjm, wnjm, ts, te, li0, gi0, cnt0, accj, s_lut, di0);
            }
            if (any1) {
              if (full1) tile_general_prod<MODE, true>(Eg1, yi1, wp1, lo1, hi1, pjm, Fg, yjm, wnjm, ts, te, li1, gi1, cnt1, accj, s_lut, di1);
              else       tile_general_prod<MODE, false>(Eg1, yi1, wp1, lo1, hi1, pjm, Fg, yjm, wnjm, ts, te, li1, gi1, cnt1, accj, s_lut, di1);
            }
          } else {
          if (any0) {
            if (full0) tile_general<MODE, true, HINGE>(si0, yi0, wp0, lo0, hi0, pjm, sjm, yjm, wnjm, c, ts, te, li0, gi0, cnt0, accj, P.margin, s_lut, di0);
            else       tile_general<MODE, false, HINGE>(si0, yi0, wp0, lo0, hi0, pjm, sjm, yjm, wnjm, c, ts, te, li0, gi0, cnt0, accj, P.margin, s_lut, di0);
          }
          if (any1) {
            if (full1) tile_general<MODE, true, HINGE>(si1, yi1, wp1, lo1, hi1, pjm, sjm, yjm, wnjm, c, ts, te, li1, gi1, cnt1, accj, P.margin, s_lut, di1);
            else       tile_general<MODE, false, HINGE>(si1, yi1, wp1, lo1, hi1, pjm, sjm, yjm, wnjm, c, ts, te, li1, gi1, cnt1, accj, P.margin, s_lut, di1);
          }
          }
        }
        if (tally) { if (fast) d_fastcyc += (u64)(clock64() - d_g0); else d_gencyc += (u64)(clock64() - d_g0); }
        TR_ADD(tr_tile);
        if (accj != 0.f && !(P.debug & 4)) gadd(pjm, accj);      // (debug bit 4: timing experiment without the RED)
        pjm = pjn; sjm = sjn; yjm = yjn; wnjm = wnjn;
        TR_ADD(tr_post);
      }
      if (pi0 < B && an0.y) {
        if (gi0 != 0.f) gadd(pi0, -gi0);
        if (DYN && li0 != 0.f) atomicAdd(A.lossrow + pi0, li0);
        if (DYN && cnt0) atomicAdd(A.cnt + pi0, cnt0);
      }
      if (pi1 < B && an1.y) {
        if (gi1 != 0.f) gadd(pi1, -gi1);
        if (DYN && li1 != 0.f) atomicAdd(A.lossrow + pi1, li1);
        if (DYN && cnt1) atomicAdd(A.cnt + pi1, cnt1);
      }
      if (!DYN) {
        if (fold) {
          // Occurrence weight of the rows' primary group.  Usually all rows of the I-block that have pairs belong to ONE
          // group (equal group start): its weight is computed once per warp and kept while the group stays the same.
          const u32 qmin = __reduce_min_sync(0xFFFFFFFFu, min(act0 ? lo0 : 0xFFFFFFFFu, act1 ? lo1 : 0xFFFFFFFFu));
          const u32 qmax = __reduce_max_sync(0xFFFFFFFFu, max(act0 ? lo0 : 0u, act1 ? lo1 : 0u));
          if (qmin == qmax) {
            if (qmin != occ_a) {
              if (counted) occ_w = A.lossrow[qmin];                 // (parked by k_seg's scatter phase)
              else {
                const u32 pg = A.sgrp[qmin];
                occ_w = pg != kEmpty ? occ_pow(A.cprim[pg], fold_power) : 0.f;
              }
              occ_a = qmin;
            }
            li0 *= occ_w; li1 *= occ_w;
          } else if (qmin != 0xFFFFFFFFu) {
            if (counted) {
              li0 *= act0 ? A.lossrow[pi0] : 0.f; li1 *= act1 ? A.lossrow[pi1] : 0.f;
            } else {
              const u32 pg0 = act0 ? A.sgrp[pi0] : kEmpty, pg1 = act1 ? A.sgrp[pi1] : kEmpty;
              const u64 ch0 = pg0 != kEmpty ? A.cprim[pg0] : 0ull, ch1 = pg1 != kEmpty ? A.cprim[pg1] : 0ull;
              li0 *= occ_pow(ch0, fold_power); li1 *= occ_pow(ch1, fold_power);
            }
          }
        }
        lsum += (double)li0 + (double)li1;
      }
      TR_ADD(tr_flush);
      if (tally) { d_busy += (u64)(clock64() - d_t0); ++d_units; }
      have = have_n; sg = sg_n; R = Rn;
    }
    if ((P.debug & 1) && ln == 0) {
      const u64 now = globaltimer();
      u64* rec = A.dbgbuf + 8 * (size_t)(blockIdx.x * kPairWarps + (threadIdx.x >> 5));
      rec[0] = d_start; rec[1] = now; rec[2] = d_busy; rec[3] = (u64)d_units | ((u64)d_gen << 32);
      rec[4] = d_fast; rec[5] = d_fastcyc; rec[6] = d_gencyc; rec[7] = d_eighths;
#ifdef RN_TRACE
      rec[2] = d_busy | ((u64)d_units << 40) | ((u64)(d_fast + d_gen) << 48);
      rec[3] = tr_take; rec[4] = tr_pre; rec[5] = tr_tile; rec[6] = tr_post; rec[7] = tr_flush;
#endif
      if (P.debug & 2) {      // grid-wide tallies: ~33 000 same-address atomics when the warps finish -- they perturb the timeline
        atomicAdd(&ctl->dbg[1], d_busy); atomicMax(&ctl->dbg[2], now);
        atomicAdd(&ctl->dbg[4], (u64)d_units); atomicAdd(&ctl->dbg[5], (u64)d_fast);
        atomicAdd(&ctl->dbg[6], (u64)d_gen); atomicAdd(&ctl->dbg[7], d_gencyc); atomicAdd(&ctl->dbg[3], d_fastcyc);
      }
    }
  }
  const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  u32 epoch2 = LAMBDA ? kLambdaBarriers * gridDim.x : 0u;        // (lambda_prepare's barriers came first)
  if (!DYN) {
    // ---- this CTA's share of sum w * loss, then (after the grid barrier) un-permute and scale the gradient ----
    lsum = warp_sum(lsum);
    if (ln == 0) red_d[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int q = 0; q < kPairWarps; ++q) t += red_d[q];
      if (DET) A.lpart[blockIdx.x] = t;                       // (summed in CTA order behind the barrier)
      else if (t != 0.0) atomicAdd(&ctl->loss_sum, t);
    }
    // nothing else in front of the barrier: the last warp to leave the pair loop sets the pace
    stamp(ctl, 21);
    grid_sync(&ctl->bar2_cnt, epoch2, &ctl->err);
    stamp(ctl, 22);
    // one round trip of independent loads (pair count, permutation, parked occurrence weight, gradient sum), then the store
    const u64 n = *reinterpret_cast<volatile u64*>(&ctl->n_pair);
    const float denom = P.reduce_mean ? ((float)n + 1.0e-10f) : 1.0f;       // PW:125-126, PW:13
    const float gscale = P.factor / denom;
    auto row_scale = [&](u32 p) -> float { return fold_power == 0.f ? gscale : A.lossrow[p] * gscale; };
    // output index: plain, or chunked for a following reduce-scatter (row i -> (i / Bl) * out_chunk + i % Bl)
    auto out_at = [&](u32 row) -> size_t { return P.out_chunk ? (size_t)(row / P.rm.Bl) * P.out_chunk + (row % P.rm.Bl) : row; };
    if (P.focal_w != 0.f) {
      // ... plus the focal term's gradient, focal_w / B * d focal_i / d logit_i (the mean is over all rows)
      const float fsc = P.focal_w / (float)B;
      for (u32 p = gtid; p < B; p += gthreads) {
        const u32 row = A.perm[p];
        const float gp = DET ? (float)((double)(long long)A.g64[p] * s_dscale[1]) : A.gacc[p];
        P.dlogits[row] = gp * row_scale(p) + fsc * focal_row(P.logits[row], P.labels[row], P.focal_alpha, P.focal_gamma, P.focal_stop).y;
      }
    } else if (DET) for (u32 p = gtid; p < B; p += gthreads) P.dlogits[out_at(A.perm[p])] = (float)((double)(long long)A.g64[p] * s_dscale[1]) * row_scale(p);
    else for (u32 p = gtid; p < B; p += gthreads) P.dlogits[out_at(A.perm[p])] = A.gacc[p] * row_scale(p);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      double lsum_all = *reinterpret_cast<volatile double*>(&ctl->loss_sum);
      if (DET) { lsum_all = 0.0; for (u32 c = 0; c < gridDim.x; ++c) lsum_all += A.lpart[c]; }
      const double tot = lsum_all * P.loss_unit;
      float lossv = (float)(tot / (double)denom);
      if (P.focal_w != 0.f) lossv += P.focal_w * (float)(*reinterpret_cast<volatile double*>(&ctl->acc_d[3]) / (double)B);
      // a device-side failure (grid barrier timed out, hash table full: the arena was not clean) must not look like a result
      if (*reinterpret_cast<volatile u32*>(&ctl->err)) lossv = __int_as_float(0x7FC00000);
      *P.loss = lossv;
      *P.n_pair_f32 = (float)n;                  // PW:276
      *P.n_pair = (int64_t)n;
      if (P.out_chunk)                           // the (partial) loss rides in every chunk of the reduce-scatter
        for (u32 r = 0; r * P.rm.Bl < B; ++r) {
          float* tail = P.dlogits + (size_t)r * P.out_chunk + P.rm.Bl;
          tail[0] = lossv;
          for (u32 q = 1; P.rm.Bl + q < P.out_chunk; ++q) tail[q] = 0.f;
        }
      ctl->ts[23] = globaltimer();
    }
    // leave the arena clean: J ranges, the group records of the counting path, the control block (last CTA out)
    {
      const uint2 z2 = make_uint2(0, 0);
      for (u32 v = gtid; v < nvb; v += gthreads) A.blk_w[v] = z2;
      if (A.fast) clean_records(A, blockIdx.x, gridDim.x);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) { __threadfence(); ctl_finish(ctl); }
      }
    }
    return;
  }
  // ---- finalisation when the pair set depends on scores / negative-side weights (all CTAs, after a grid barrier) ----
  stamp(ctl, 21);
  grid_sync(&ctl->bar2_cnt, epoch2, &ctl->err);
  stamp(ctl, 22);
  if (A.xcnt_out) {
    // global mode: publish this rank's partial per-row counts by original row and stop; k_fin_dyn finishes the call
    for (u32 p = gtid; p < B; p += gthreads) A.xcnt_out[A.perm[p]] = A.cnt[p];
    const uint2 z2 = make_uint2(0, 0);
    for (u32 v = gtid; v < nvb; v += gthreads) A.blk_w[v] = z2;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) { ctl->bar2_cnt = 0; ctl->fin_done = 0; }     // (everybody is past the barrier)
    }
    return;
  }
  dyn_finalize(P, A, counted ? &A.rec->npair : A.cprim, counted ? (u32)(sizeof(GRec) / sizeof(u64)) : 1u, nvb, epoch2, red_u, red_d);
}

// Counts -> occurrence weights -> gradient and loss for a score- / weight-dependent pair set (the tail of k_pair, or
// k_fin_dyn in the global mode).  All CTAs of a cooperative grid; one grid barrier inside.
__device__ __forceinline__ void dyn_finalize(const PairParams& P, const KpArgs& A, const u64* cprim, u32 cstride, u32 nvb,
                                             u32& epoch2, u64* red_u, double* red_d) {
  Ctl* ctl = A.ctl;
  const u32 B = P.B, ln = lane_id();
  const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  {
    // F_a: exact counts from the kernel's per-row tallies: per row, per PRIMARY group (PW:286-289), total
    u64 tot_c = 0;
    for (u32 p0 = 0; p0 < B; p0 += gthreads) {             // uniform trip count: the body uses warp collectives
      const u32 p = p0 + gtid;
      u32 cn = 0, pg = kEmpty;
      if (p < B) {
        if (A.xworld) {                                    // global mode: the ranks' partial counts of this row
          const u32 row = A.perm[p];
          for (int r = 0; r < A.xworld; ++r) cn += A.xcnt_peer[r][row];
        } else {
          cn = A.cnt[p];
        }
        if (cn) pg = A.sgrp[p];
        if (P.row_pairs) P.row_pairs[A.perm[p]] = (int64_t)cn;
      }
      const u32 m = __match_any_sync(0xFFFFFFFFu, pg);
      const u32 tot = __reduce_add_sync(m, cn);
      if (pg != kEmpty && ln == (u32)(__ffs(m) - 1)) atomicAdd(const_cast<u64*>(cprim) + (size_t)pg * cstride, (u64)tot);
      tot_c += cn;
    }
    tot_c = warp_sum(tot_c);
    if (ln == 0) red_u[threadIdx.x >> 5] = tot_c;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t = 0;
      for (int q = 0; q < kPairWarps; ++q) t += red_u[q];
      if (t) atomicAdd(&ctl->n_pair, t);
    }
    const uint2 z2 = make_uint2(0, 0);
    for (u32 v = gtid; v < nvb; v += gthreads) A.blk_w[v] = z2;        // (the J ranges were last read in the prologue)
    grid_sync(&ctl->bar2_cnt, epoch2, &ctl->err);
  }
  // F_b: scale, apply the occurrence weight, un-permute the gradient, reduce the loss
  const u64 n = *reinterpret_cast<volatile u64*>(&ctl->n_pair);
  const float denom = P.reduce_mean ? ((float)n + 1.0e-10f) : 1.0f;       // PW:125-126, PW:13
  const float gscale = P.factor / denom;
  auto out_at = [&](u32 row) -> size_t { return P.out_chunk ? (size_t)(row / P.rm.Bl) * P.out_chunk + (row % P.rm.Bl) : row; };
  double lp = 0.0;
  for (u32 p = gtid; p < B; p += gthreads) {
    const float g = A.gacc[p], l = A.lossrow[p];
    float wocc = 1.f;
    if (P.power != 0.f && (g != 0.f || l != 0.f)) {
      const u64 ch = cprim[(size_t)A.sgrp[p] * cstride];
      wocc = ch ? ((P.power == 1.0f) ? (float)ch : powf((float)ch, P.power)) : 0.f;   // PW:147-149
    }
    P.dlogits[out_at(A.perm[p])] = g * wocc * gscale;
    lp += (double)l * (double)wocc;
  }
  lp = warp_sum(lp);
  if (ln == 0) red_d[threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int q = 0; q < kPairWarps; ++q) t += red_d[q];
    if (t != 0.0) atomicAdd(&ctl->loss_sum, t);
  }
  grid_sync(&ctl->bar2_cnt, epoch2, &ctl->err);
  // behind the last barrier: the scalars (CTA 0) and, now that nobody reads the group records any more, a clean arena
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->ts[23] = globaltimer();
    const double tot = *reinterpret_cast<volatile double*>(&ctl->loss_sum) * P.loss_unit;
    float lossv = (float)(tot / (double)denom);
    if (*reinterpret_cast<volatile u32*>(&ctl->err)) lossv = __int_as_float(0x7FC00000);      // (see k_pair: a failed call returns NaN)
    *P.loss = lossv;
    *P.n_pair_f32 = (float)n;                  // PW:276
    *P.n_pair = (int64_t)n;
    if (P.out_chunk)                           // the (partial) loss rides in every chunk of the reduce-scatter
      for (u32 r = 0; r * P.rm.Bl < B; ++r) {
        float* tl = P.dlogits + (size_t)r * P.out_chunk + P.rm.Bl;
        tl[0] = lossv;
        for (u32 q = 1; P.rm.Bl + q < P.out_chunk; ++q) tl[q] = 0.f;
      }
  }
  if (A.fast) clean_records_grid(A.rec, A.rec2_off, A.glist, A.gcount, A.ngt);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) { __threadfence(); ctl_finish(ctl); }
  }
}

// Second stage of a global-mode call whose pair set depends on scores / weights (see KpArgs::xcnt_out).
__global__ void __launch_bounds__(kPairThreads) k_fin_dyn(PairParams P, KpArgs A) {
  __shared__ u64 red_u[kPairWarps];
  __shared__ double red_d[kPairWarps];
  grid_dep_wait();
  const bool counted = A.fast && ld_relaxed(&A.ctl->fallback) == 0u;
  const u64* cprim = counted ? &A.rec->npair : A.cprim;
  const u32 cstride = counted ? (u32)(sizeof(GRec) / sizeof(u64)) : 1u;
  u32 epoch2 = 0;
  // the ranks' partial counts summed in ROW order first (coalesced 16-byte peer loads), then gathered locally
  const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const u32 n4 = P.B / 4;                                    // (the rows per rank are a multiple of 16)
  for (u32 i = gtid; i < n4; i += gthreads) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < A.xworld; ++r) {
      const uint4 v = reinterpret_cast<const uint4*>(A.xcnt_peer[r])[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<uint4*>(A.xsum)[i] = acc;
  }
  grid_sync(&A.ctl->bar2_cnt, epoch2, &A.ctl->err);
  KpArgs A1 = A;
  A1.xworld = 1; A1.xcnt_peer[0] = A.xsum;
  dyn_finalize(P, A1, cprim, cstride, 2 * A.nib, epoch2, red_u, red_d);
}

// CTAs per SM of the pair kernel's cooperative grid on the current device (all co-resident), cached per thread.
static int pair_blocks_per_sm(const void* func, int dev) {
  struct E { const void* f; int dev; int bps; };
  static thread_local E cache[32];
  static thread_local int n = 0;
  for (int i = 0; i < n; ++i) if (cache[i].f == func && cache[i].dev == dev) return cache[i].bps;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, func, kPairThreads, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (nb < 1) nb = 1;
  static const int want = tune_int("RN_PAIR_BPS", 1);
  const int bps = want < nb ? (want < 1 ? 1 : want) : nb;
  if (n < 32) cache[n++] = E{func, dev, bps};
  return bps;
}

template <int MODE, bool DET = false, bool HINGE = false>
static cudaError_t launch_pair(const PairParams& P, const KpArgs& A, cudaStream_t st) {
  // (the opt-in to large dynamic shared memory is a per-device function attribute)
  static size_t smem_set[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const void* fn = (const void*)k_pair<MODE, DET, HINGE>;
  const int bps = pair_blocks_per_sm(fn, dev);
  if (bps < 1) return cudaErrorUnknown;
  PairParams p = P; KpArgs a = A;
  void* args[] = {&p, &a};
  // cost prefix of the static partition: s_pi[2 nib + 1], s_jn[2 nib]
  const size_t smem = A.nib <= kMaxNibS ? sizeof(u32) * (4 * (size_t)A.nib + 2) : 0;
  if (smem > smem_set[dev]) {
    const size_t want = sizeof(u32) * (4 * (size_t)kMaxNibS + 2);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) return e;
    smem_set[dev] = want;
  }
  return launch_coop(fn, device_sm_count() * bps, kPairThreads, args, st, smem);
}

constexpr int M_DET = 16;          // (dispatch only: deterministic instantiation of a non-dynamic mode)
constexpr int M_HINGE = 32;        // (dispatch only: hinge pair loss)
static const void* pair_func(int mode) {
  switch (mode) {
#define RN_CASE(m) case M_HINGE | (m): return (const void*)k_pair<m, false, true>;
    RN_CASE(0) RN_CASE(M_WRONG)
    RN_CASE(M_HASW) RN_CASE(M_HASW | M_WRONG)
    RN_CASE(M_HASW | M_DIFF) RN_CASE(M_HASW | M_DIFF | M_WRONG)
    RN_CASE(M_HASW | M_RWN) RN_CASE(M_HASW | M_RWN | M_WRONG)
    RN_CASE(M_HASW | M_DIFF | M_RWN) RN_CASE(M_HASW | M_DIFF | M_RWN | M_WRONG)
#undef RN_CASE
    case M_HINGE | M_HASW | M_DIFF | M_LUT: return (const void*)k_pair<M_HASW | M_DIFF | M_LUT, false, true>;
    case M_HASW | M_DIFF | M_LUT: return (const void*)k_pair<M_HASW | M_DIFF | M_LUT>;
    case M_HASW | M_DIFF | M_LAMBDA: return (const void*)k_pair<M_HASW | M_DIFF | M_LAMBDA>;
    case M_DET: return (const void*)k_pair<0, true>;
    case M_DET | M_HASW: return (const void*)k_pair<M_HASW, true>;
    case M_DET | M_HASW | M_DIFF: return (const void*)k_pair<M_HASW | M_DIFF, true>;
#define RN_CASE(m) case m: return (const void*)k_pair<m>;
    RN_CASE(0) RN_CASE(M_WRONG)
    RN_CASE(M_HASW) RN_CASE(M_HASW | M_WRONG)
    RN_CASE(M_HASW | M_DIFF) RN_CASE(M_HASW | M_DIFF | M_WRONG)
    RN_CASE(M_HASW | M_RWN) RN_CASE(M_HASW | M_RWN | M_WRONG)
    RN_CASE(M_HASW | M_DIFF | M_RWN) RN_CASE(M_HASW | M_DIFF | M_RWN | M_WRONG)
#undef RN_CASE
  }
  return nullptr;
}

static cudaError_t dispatch_pair(int mode, const PairParams& P, const KpArgs& A, cudaStream_t st) {
  switch (mode) {
#define RN_CASE(m) case M_HINGE | (m): return launch_pair<m, false, true>(P, A, st);
    RN_CASE(0) RN_CASE(M_WRONG)
    RN_CASE(M_HASW) RN_CASE(M_HASW | M_WRONG)
    RN_CASE(M_HASW | M_DIFF) RN_CASE(M_HASW | M_DIFF | M_WRONG)
    RN_CASE(M_HASW | M_RWN) RN_CASE(M_HASW | M_RWN | M_WRONG)
    RN_CASE(M_HASW | M_DIFF | M_RWN) RN_CASE(M_HASW | M_DIFF | M_RWN | M_WRONG)
#undef RN_CASE
    case M_HINGE | M_HASW | M_DIFF | M_LUT: return launch_pair<M_HASW | M_DIFF | M_LUT, false, true>(P, A, st);
    case M_HASW | M_DIFF | M_LUT: return launch_pair<M_HASW | M_DIFF | M_LUT>(P, A, st);
    case M_HASW | M_DIFF | M_LAMBDA: return launch_pair<M_HASW | M_DIFF | M_LAMBDA>(P, A, st);
    case M_DET: return launch_pair<0, true>(P, A, st);
    case M_DET | M_HASW: return launch_pair<M_HASW, true>(P, A, st);
    case M_DET | M_HASW | M_DIFF: return launch_pair<M_HASW | M_DIFF, true>(P, A, st);
#define RN_CASE(m) case m: return launch_pair<m>(P, A, st);
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
// (graph = true: the call stays on the product's default path, one launch of a cached CUDA graph that also holds two
// event-record nodes around the pair kernel; the call then waits for the stream and reads the elapsed time itself)
static struct { bool on = false; bool graph = false; int cap = 0; int n = 0; cudaEvent_t* ev = nullptr; float* ms = nullptr; } g_prof;

extern "C" int rn_profile_disable(void) {
  if (g_prof.ev) { for (int i = 0; i < 2 * g_prof.cap; ++i) cudaEventDestroy(g_prof.ev[i]); delete[] g_prof.ev; }
  delete[] g_prof.ms;
  g_prof.ev = nullptr; g_prof.ms = nullptr; g_prof.on = false; g_prof.graph = false; g_prof.cap = g_prof.n = 0;
  return RN_OK;
}
extern "C" int rn_profile_enable_ex(int32_t max_calls, int32_t in_graph);
extern "C" int rn_profile_enable(int32_t max_calls) {
  static const int in_graph = tune_int("RN_PROFILE_GRAPH", 1);
  return rn_profile_enable_ex(max_calls, in_graph);
}
extern "C" int rn_profile_enable_ex(int32_t max_calls, int32_t in_graph) {
  rn_profile_disable();
  if (max_calls <= 0 || max_calls > (1 << 20)) return RN_ERR_ARG;
  g_prof.ev = new cudaEvent_t[2 * max_calls];
  for (int i = 0; i < 2 * max_calls; ++i)
    if (cudaEventCreate(&g_prof.ev[i]) != cudaSuccess) return RN_ERR_LAUNCH;
  g_prof.ms = new float[max_calls];
  g_prof.cap = max_calls; g_prof.n = 0; g_prof.on = true;
  g_prof.graph = in_graph != 0;
  return RN_OK;
}
extern "C" int rn_profile_collect(float* ms_out_host, int32_t capacity, int32_t* n_out_host) {
  if (!ms_out_host || !n_out_host) return RN_ERR_ARG;
  int n = g_prof.n < capacity ? g_prof.n : capacity;
  for (int i = 0; i < n; ++i) {
    if (g_prof.ms[i] >= 0.f) { ms_out_host[i] = g_prof.ms[i]; continue; }          // (timed inside a graph launch)
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

extern "C" int rn_pairwise_scratch_init(void* scratch, size_t scratch_bytes, void* stream) {
  if (!scratch || !scratch_bytes) return RN_ERR_ARG;
  return cudaMemsetAsync(scratch, 0, scratch_bytes, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

// the counting path runs without k_init (persistent arena, one key column, contiguous rows, the whole pair space)
static bool counting_eligible(const rn_pairwise_args* a) {
  static const int off = tune_int("RN_SEG_COUNT", 1);
  // (batches beyond kMaxNibS I-blocks need the explicit unit records only the radix tail builds)
  // global mode (blocked rows): the rank blocks must be whole tiles of the count phase
  return off && a->scratch_persistent && a->K == 1 && (a->block_rows % kGTile) == 0 && (a->part_count == 1 || a->block_rows) &&
         !(a->block_rows && a->row_pairs) && (a->B + kIB - 1) / kIB <= (int64_t)kMaxNibS;
}

extern "C" int rn_pairwise_launch_count(int64_t B, int32_t K) {
  if (B <= 0 || K <= 0) return 0;
  return 3;        // k_init, k_seg<HeadsTail>, k_pair (2 with a persistent arena and one key column: no k_init)
}

static int validate_pairwise(const rn_pairwise_args* a, bool split = false) {
  if (!a || a->B <= 0 || a->B > (1ll << 28) || a->K <= 0 || a->K > 8) return RN_ERR_ARG;
  if (!a->keys || !a->logits || !a->labels || !a->loss || !a->n_pair_f32 || !a->n_pair || !a->dlogits) return RN_ERR_ARG;
  if (a->label_func != RN_LABEL_STEP && a->label_func != RN_LABEL_DIFF && a->label_func != RN_LABEL_GAIN2 &&
      a->label_func != RN_LABEL_LUT && a->label_func != RN_LABEL_LAMBDA) return RN_ERR_UNSUPPORTED;
  if ((a->label_func == RN_LABEL_LUT) != (a->weight_lut != nullptr)) return RN_ERR_ARG;
  if (a->label_func == RN_LABEL_LUT) {
    // the level table rides on the label-gain tiles of the one-GPU call; the score- / weight-dependent pair sets, the
    // deterministic instantiation and the blocked rows of the global mode have no table variant
    if (a->only_wrong || a->rw_neg || a->deterministic || a->block_rows || a->part_count != 1 || split) return RN_ERR_UNSUPPORTED;
  }
  if (a->label_func == RN_LABEL_LAMBDA) {
    // per-pair weights from the rows' score ranks: general tiles of the logistic loss on one GPU (the pre-pass lives in
    // the pair kernel; the negative-side weight column carries the rank discounts)
    if (a->only_wrong || a->rw_neg || a->deterministic || a->block_rows || a->part_count != 1 || split ||
        a->pair_loss != RN_LOSS_LOGISTIC || a->focal_weight != 0.f) return RN_ERR_UNSUPPORTED;
  }
  if (a->pair_loss != RN_LOSS_LOGISTIC && a->pair_loss != RN_LOSS_HINGE) return RN_ERR_UNSUPPORTED;
  if (a->pair_loss == RN_LOSS_HINGE && !(a->margin >= 0.f)) return RN_ERR_ARG;
  if (a->part_count < 1 || a->part_rank < 0 || a->part_rank >= a->part_count) return RN_ERR_ARG;
  if (a->scratch_rows && a->scratch_rows < a->B) return RN_ERR_ARG;
  if (a->block_rows) {
    if (a->block_rows < 0 || a->B % a->block_rows || a->block_stride <= 0 || (a->block_stride & 15) ||
        a->block_stride / 4 > 0xFFFFFFFFll) return RN_ERR_ARG;
    if (a->out_chunk && (a->out_chunk <= a->block_rows || a->out_chunk > 0x7FFFFFFFll)) return RN_ERR_ARG;
    if ((a->only_wrong || a->rw_neg) && !split) return RN_ERR_UNSUPPORTED;
    if (a->gather_dst) {
      const int64_t world = a->B / a->block_rows;
      if (world > 8 || check_align(a->gather_dst)) return RN_ERR_ARG;
      for (int64_t r = 0; r < world; ++r) if (!a->peer_blocks[r] || check_align(a->peer_blocks[r])) return RN_ERR_ARG;
    }
  } else if (a->out_chunk || a->gather_dst) return RN_ERR_ARG;
  const void* ptrs[] = {a->keys, a->logits, a->labels, a->row_ok, a->rw_pos, a->rw_neg, a->dlogits, a->row_pairs, a->weight_lut};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  return RN_OK;
}

extern "C" int rn_pairwise_fwd_bwd(const rn_pairwise_args* a, void* scratch, size_t scratch_bytes, void* stream) {
  RN_NVTX_RANGE("rn_pairwise_fwd_bwd");
  return rn::pairwise_call(a, scratch, scratch_bytes, stream, nullptr, 0);
}

// The call behind rn_pairwise_fwd_bwd.  split != nullptr (global mode, score- / weight-dependent pair set): stage 1 runs
// the kernels up to the per-row partial counts (published by original row in split->xcnt_out), stage 2 -- after the
// caller's cross-rank barrier -- launches k_fin_dyn, which sums the ranks' counts and finishes the call.
int rn::pairwise_call(const rn_pairwise_args* a, void* scratch, size_t scratch_bytes, void* stream, const DynSplit* split,
                      int stage) {
  int rc = validate_pairwise(a, split != nullptr);
  if (rc) return rc;
  const bool dyn = a->only_wrong || a->rw_neg;
  // partial (multi-GPU) evaluation needs globally consistent counts: position arithmetic gives them, the
  // score- / weight-dependent filters need the ranks' counts summed between counting and weighting (split)
  if ((a->part_count > 1 || a->block_rows) && dyn && !split) return RN_ERR_UNSUPPORTED;
  if (split && !dyn) return RN_ERR_ARG;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  const Layout L = make_layout(a->scratch_rows ? a->scratch_rows : a->B, a->K, a->B);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  if (a->focal_weight != 0.f && (dyn || a->block_rows || a->part_count > 1)) return RN_ERR_UNSUPPORTED;
  if (a->focal_weight != 0.f && ((a->focal_alpha != 0.f && !(a->focal_alpha > 0.f && a->focal_alpha < 1.f)) || a->focal_gamma < 0.f))
    return RN_ERR_ARG;                                  // (focal_loss.py:43-46)
  const bool det = a->deterministic != 0;
  if (det && (dyn || a->block_rows || a->part_count > 1)) return RN_ERR_UNSUPPORTED;
  const bool hinge = a->pair_loss == RN_LOSS_HINGE;
  if (det && hinge) return RN_ERR_UNSUPPORTED;
  const bool lut = a->label_func == RN_LABEL_LUT, lambda = a->label_func == RN_LABEL_LAMBDA;
  if (!split && !lambda && !(g_prof.on && g_prof.n < g_prof.cap)) {       // (LambdaRank weights: the general kernels at any size)
    // batches of up to 1024 rows: one launch of one CTA, everything in shared memory (small.cu)
    int smode = 0;
    const bool sdiff = a->label_func == RN_LABEL_DIFF || a->label_func == RN_LABEL_GAIN2 || lut;
    if (sdiff || a->rw_pos || a->rw_neg) smode |= M_HASW;
    if (sdiff) smode |= M_DIFF;
    if (a->rw_neg) smode |= M_RWN;
    if (a->only_wrong) smode |= M_WRONG;
    int src = RN_OK;
    if (small_pairwise(a, scratch, st, smode, hinge, &src)) return src;
  }
  const bool fast = counting_eligible(a) && !det;       // (the counting path places groups and rows in arrival order)
  PairParams P{};
  P.det = det ? 1 : 0;
  P.focal_w = a->focal_weight; P.focal_alpha = a->focal_alpha; P.focal_gamma = a->focal_gamma;
  P.focal_stop = a->focal_stop_weight_gradient;
  P.B = (u32)a->B; P.K = a->K; P.gbits = L.gbits;
  P.logits = a->logits; P.labels = a->labels; P.rw_pos = a->rw_pos; P.rw_neg = a->rw_neg;
  P.factor = a->factor; P.power = a->power; P.reduce_mean = a->reduce_mean; P.dyn_count = dyn ? 1 : 0;
  P.c_log2 = hinge ? a->factor : a->factor * 1.4426950408889634f;
  P.margin = a->margin; P.loss_unit = hinge ? 1.0 : 0.6931471805599453;
  P.gain2 = (a->label_func == RN_LABEL_GAIN2 || lambda) ? 1 : (lut ? 2 : 0);
  P.lambda = lambda ? 1 : 0;
  static const int pair_debug = tune_int("RN_PAIR_DEBUG", 0);
  P.debug = pair_debug;
  P.part_rank = a->part_rank; P.part_count = a->part_count;
  static const int ascending = tune_int("RN_PAIR_ASCENDING", 0);
  P.ascending = ascending;
  P.loss = a->loss; P.n_pair_f32 = a->n_pair_f32; P.n_pair = a->n_pair; P.dlogits = a->dlogits; P.row_pairs = a->row_pairs;
  P.rm = RowMap{0, 0, 0, 0}; P.out_chunk = 0;
  if (a->block_rows) {
    P.rm.Bl = (u32)a->block_rows; P.rm.s8 = (u32)(a->block_stride / 8); P.rm.s4 = (u32)(a->block_stride / 4);
    P.rm.s1 = (u32)a->block_stride; P.out_chunk = (u32)a->out_chunk;
  }
  HeadsTail H{};
  H.P = P;
  H.aj = at<uint2>(base, L.aj); H.ss = at<float>(base, L.ss); H.sy = at<float>(base, L.sy);
  H.swp = at<float>(base, L.swp); H.swn = at<float>(base, L.swn);
  H.gacc = at<float>(base, L.gacc); H.lossrow = at<float>(base, L.lossrow); H.cnt = at<u32>(base, L.cnt);
  H.perm = at<u32>(base, L.perm); H.sgrp = at<u32>(base, L.sgrp);
  H.blk = at<uint2>(base, L.blk); H.units = at<uint2>(base, L.units); H.nib = L.nib;
  H.cprim = at<u64>(base, L.cprim); H.pgid = at<u32>(base, L.slot1);
  H.target_units = target_units();
  H.pp = PrePart{};
  H.g64 = at<unsigned long long>(base, L.misc);
  SegInputs in{a->B, a->K, a->keys, a->labels, a->row_ok, true, true};
  in.rm = P.rm;
  if (a->gather_dst) {
    in.gather.world = (u32)(a->B / a->block_rows); in.gather.n16 = (u32)(a->block_stride / 16);
    in.gather.dst = static_cast<uint4*>(a->gather_dst);
    for (u32 r = 0; r < in.gather.world; ++r) in.gather.src[r] = static_cast<const uint4*>(a->peer_blocks[r]);
  }
  static const int allow_merged = tune_int("RN_SEG_MERGED", 1);
  in.fast = fast;
  in.allow_merged = allow_merged && a->part_count == 1 && !fast && !det;      // ranks of the global mode need identical ids; so does a deterministic call (table slots depend on insertion order)
  if (in.allow_merged && a->K == 1 && !P.rm.Bl) { P.gbits = seg_merged_gbits(L); H.P.gbits = P.gbits; }
  KpArgs A{};
  A.aj = H.aj; A.ss = H.ss; A.sy = H.sy; A.swp = (a->rw_pos || lambda) ? H.swp : nullptr; A.swn = H.swn;
  A.units = H.units; A.blk = H.blk; A.nib = L.nib; A.target_units = H.target_units;
  static const int cost_gen = tune_int("RN_PAIR_COST_GEN", 17);
  A.cost_gen = kCostUnit * (u32)(cost_gen < 1 ? 1 : (cost_gen > 64 ? 64 : cost_gen));            // (knobs are in eighths of a fast tile)
  static const int cost_switch = tune_int("RN_PAIR_COST_SWITCH", 4);
  A.cost_switch = kCostUnit * (u32)(cost_switch < 0 ? 0 : (cost_switch > 64 ? 64 : cost_switch));
  static const int cost_straddle = tune_int("RN_PAIR_COST_STRADDLE", 12);
  static const int cost_levels = tune_int("RN_PAIR_COST_LEVELS", 4);
  A.cost_straddle = kCostUnit * (u32)(cost_straddle < 0 ? 0 : (cost_straddle > 64 ? 64 : cost_straddle));
  A.cost_levels = (u32)(cost_levels < 0 ? 0 : (cost_levels > 64 ? 64 : cost_levels));
  A.gacc = H.gacc; A.lossrow = H.lossrow; A.cnt = H.cnt; A.perm = H.perm; A.ctl = at<Ctl>(base, L.ctl);
  A.sgrp = H.sgrp; A.cprim = H.cprim; A.lut = a->weight_lut;
  A.fast = fast ? 1 : 0; A.rec2_off = (fast && a->block_rows) ? (u32)((L.rec2 - L.rec) / sizeof(GRec)) : 0u; A.rec = at<GRec>(base, L.rec); A.glist = at<u32>(base, L.glist); A.gcount = at<u32>(base, L.gcount);
  A.ngt = (u32)((a->B + kGTile - 1) / kGTile); A.blk_w = H.blk;
  A.dbgbuf = at<u64>(base, L.gstat);
  A.g64 = H.g64; A.lpart = at<double>(base, L.lpart);
  if (split) {
    A.xcnt_out = split->xcnt_out; A.xworld = split->world;
    for (int r = 0; r < split->world; ++r) A.xcnt_peer[r] = split->xcnt_peer[r];
    A.xsum = at<u32>(base, L.misc);
  }
  int mode = 0;
  const bool diff = a->label_func == RN_LABEL_DIFF || a->label_func == RN_LABEL_GAIN2 || lut || lambda;
  if (diff || a->rw_pos || a->rw_neg) mode |= M_HASW;
  if (diff) mode |= M_DIFF;
  if (lut) mode |= M_LUT;
  if (lambda) mode |= M_LAMBDA;
  if (a->rw_neg) mode |= M_RWN;
  if (a->only_wrong) mode |= M_WRONG;
  if (det) mode |= M_DET;
  if (hinge) mode |= M_HINGE;
  if (fast) {
    // the pair kernel's partition is worked out by spare CTAs of k_seg beside the scatter phase (HeadsTail::partition)
    static const int pre_on = tune_int("RN_PAIR_PREPART", 1);
    int dev = 0; cudaGetDevice(&dev);
    const int bps = pair_blocks_per_sm(pair_func(mode), dev);
    const u32 pgrid = (u32)(device_sm_count() * (bps > 0 ? bps : 1));
    PrePart pp{};
    pp.bnd = at<uint4>(base, L.bnd); pp.jn = at<u32>(base, L.jn);
    pp.pair_grid = pgrid; pp.wpr = pgrid * kPairWarps;
    pp.cost_gen = A.cost_gen; pp.cost_switch = A.cost_switch; pp.cost_straddle = A.cost_straddle;
    pp.lv_cost = (mode & M_DIFF) ? A.cost_levels : 0u;
    pp.on = (pre_on && bps > 0 && pp.wpr + 1 <= kBndCap && 4 * (size_t)L.nib + 1 + kPairWarps + 2 <= (size_t)kSegSmemWords) ? 1 : 0;
    H.pp = pp;
    A.bnd = pp.bnd; A.jn = pp.jn; A.pre_on = pp.on;
  }
  if (split && stage == 2) {
    // second stage: the finishing kernel alone (cooperative, the pair kernel's grid)
    int dev = 0; cudaGetDevice(&dev);
    const int bps = pair_blocks_per_sm((const void*)k_fin_dyn, dev);
    if (bps < 1) return RN_ERR_LAUNCH;
    KpArgs A2 = A; A2.xcnt_out = nullptr;
    void* args[] = {&P, &A2};
    if (launch_coop((const void*)k_fin_dyn, device_sm_count() * bps, kPairThreads, args, st) != cudaSuccess) return RN_ERR_LAUNCH;
    return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
  }
  if (split) A.xworld = 0;          // (stage 1 tallies its own pairs; the ranks' sums are stage 2's)
  const bool prof = g_prof.on && g_prof.n < g_prof.cap;
  cudaEvent_t* tev = (prof && g_prof.graph) ? timed_events() : nullptr;
  const void* f_seg = (L.ipt == 2) ? (const void*)k_seg<2, HeadsTail> : (const void*)k_seg<8, HeadsTail>;
  // one launch of a cached CUDA graph; while the pair kernel is being timed: the timed variant of the graph (two
  // event-record nodes around the pair kernel), or plain launches with stream events (RN_PROFILE_GRAPH=0)
  GraphCall gc((fast && !a->gather_dst) ? nullptr : seg_init_func(), f_seg, pair_func(mode), st, !prof || tev != nullptr, tev != nullptr);
  const bool in_graph = gc.capturing() || gc.updating();
  // the three launches of the call (k_init, k_seg, k_pair) on stream s
  auto enqueue = [&](cudaStream_t s, bool graph) -> bool {
    if (seg_run(L, scratch, in, H, s) != cudaSuccess) return false;
    if (prof && !graph) cudaEventRecord(g_prof.ev[2 * g_prof.n], s);
    if (prof && graph && gc.capturing() && cudaEventRecordWithFlags(tev[0], s, cudaEventRecordExternal) != cudaSuccess) return false;
    if (dispatch_pair(mode, P, A, s) != cudaSuccess) return false;
    if (prof && !graph) cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], s);
    if (prof && graph && gc.capturing() && cudaEventRecordWithFlags(tev[1], s, cudaEventRecordExternal) != cudaSuccess) return false;
    return true;
  };
  const bool ok = enqueue(gc.run_stream, in_graph);
  bool timed_in_graph = prof && in_graph;
  if (gc.finish(ok) != cudaSuccess) {
    if (gc.mode == 0) return RN_ERR_LAUNCH;
    cudaGetLastError();
    timed_in_graph = false;
    if (!enqueue(st, false)) return RN_ERR_LAUNCH;       // (graphs are switched off for this thread from now on)
  }
  if (prof) {
    g_prof.ms[g_prof.n] = -1.f;
    if (timed_in_graph) {
      // measurement pass: wait for the call and read the pair kernel's time from the graph's two events
      float ms = 0.f;
      if (cudaStreamSynchronize(st) != cudaSuccess || cudaEventElapsedTime(&ms, tev[0], tev[1]) != cudaSuccess) return RN_ERR_LAUNCH;
      g_prof.ms[g_prof.n] = ms;
    }
    ++g_prof.n;
  }
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}


// ---- GAUC (gauc.cu): the pairwise segmentation, then the concordance kernel ---------------------------------------
extern "C" size_t rn_gauc_scratch_bytes(int64_t B, int32_t K) { return rn_pairwise_scratch_bytes(B, K); }

extern "C" int rn_gauc(const rn_gauc_args* g, void* scratch, size_t scratch_bytes, void* stream) {
  RN_NVTX_RANGE("rn_gauc");
  if (!g || g->B <= 0 || g->B > (1ll << 28) || g->K <= 0 || g->K > 8) return RN_ERR_ARG;
  if (!g->keys || !g->scores || !g->labels || !g->gauc || !g->n_valid_groups) return RN_ERR_ARG;
  const void* ptrs[] = {g->keys, g->scores, g->labels, g->row_ok};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  if (!scratch || check_align(scratch)) return scratch ? RN_ERR_ALIGN : RN_ERR_ARG;
  if ((g->B + kIB - 1) / kIB > (int64_t)kMaxNibS) return RN_ERR_UNSUPPORTED;       // (the explicit work list of very large batches is the pair kernel's)
  const Layout L = make_layout(g->B, g->K);
  if (scratch_bytes < L.total) return RN_ERR_SCRATCH;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(scratch);
  static const int count_on = tune_int("RN_SEG_COUNT", 1);
  const bool fast = count_on && g->scratch_persistent && g->K == 1;
  PairParams P{};
  P.B = (u32)g->B; P.K = g->K; P.gbits = L.gbits;
  P.logits = g->scores; P.labels = g->labels;
  P.factor = 1.f; P.power = 0.f; P.reduce_mean = 1; P.c_log2 = 1.4426950408889634f;
  P.part_rank = 0; P.part_count = 1;
  P.rm = RowMap{0, 0, 0, 0};
  HeadsTail H{};
  H.P = P;
  H.aj = at<uint2>(base, L.aj); H.ss = at<float>(base, L.ss); H.sy = at<float>(base, L.sy);
  H.swp = at<float>(base, L.swp); H.swn = at<float>(base, L.swn);
  H.gacc = at<float>(base, L.gacc); H.lossrow = at<float>(base, L.lossrow); H.cnt = at<u32>(base, L.cnt);
  H.perm = at<u32>(base, L.perm); H.sgrp = at<u32>(base, L.sgrp);
  H.blk = at<uint2>(base, L.blk); H.units = at<uint2>(base, L.units); H.nib = L.nib;
  H.cprim = at<u64>(base, L.cprim); H.pgid = at<u32>(base, L.slot1);
  H.target_units = target_units();
  H.pp = PrePart{};
  H.g64 = at<unsigned long long>(base, L.misc);
  SegInputs in{g->B, g->K, g->keys, g->labels, g->row_ok, true, true};
  in.fast = fast;
  if (seg_run(L, scratch, in, H, st) != cudaSuccess) return RN_ERR_LAUNCH;
  GaucArgs A{};
  A.B = (u32)g->B; A.nib = L.nib; A.aj = H.aj; A.ss = H.ss; A.blk = H.blk; A.blk_w = H.blk;
  A.acc2 = at<u64>(base, L.misc); A.npg = at<u64>(base, L.keyA); A.gsz = at<u32>(base, L.cnt);
  // (the sorted keys may live in keyA after an even number of radix passes: nothing reads them after k_seg)
  A.gauc = g->gauc; A.auc_mean = g->auc_mean; A.n_valid = g->n_valid_groups; A.n_pair = g->n_pair; A.conc2 = g->concordant2;
  A.fast = fast ? 1 : 0; A.rec = at<GRec>(base, L.rec); A.rec2_off = 0; A.glist = at<u32>(base, L.glist);
  A.gcount = at<u32>(base, L.gcount); A.ngt = (u32)((g->B + kGTile - 1) / kGTile);
  A.ctl = at<Ctl>(base, L.ctl);
  if (launch_gauc(A, st) != cudaSuccess) return RN_ERR_LAUNCH;
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
