// EXPERIMENT, NOT BUILT (kept for the record).  Segmentation of a <= 65536-row batch inside one
// 16-CTA cluster with distributed shared memory.  It is correct (all GPU parity tests passed with it) but SLOWER than
// the grid-wide kernel on B200: 72 us against 45 us at cfg3.  Measured with in-kernel stamps: random remote
// shared-memory accesses cost ~2-3 ns each per SM (hash inserts 5 us, group-id fetch 5 us for ~1500 accesses per CTA),
// the CTA-local aggregation 8.8 us, and each radix pass 8-13 us because sixteen SMs' shared-memory bandwidth
// (128 B/clk each) carries all of the per-warp histogram traffic.  The grid-wide kernel pays L2 latency and grid
// barriers instead, but spreads the same traffic over 128 SMs.
// K1c -- group segmentation + heads + work list for the pairwise path inside ONE thread-block cluster.
//
// Same job as k_seg<HeadsTail> (segment.cuh, pairwise.cu): it replaces the reference's (B,B) group-equality matrix
// and label-order mask (pairwise_loss_from_batch.py:33-37, 68-73, 187-190) by a sort of the batch by (group, label).
// A batch of <= 65536 rows (the north-star size) is ~1.5 MB: every step of the grid-wide kernel is a chain of L2 round
// trips (~0.7 us each) plus a grid barrier through global atomics (~2 us), seven times over.  Here the whole batch
// lives in the DISTRIBUTED SHARED MEMORY of one cluster of 8 or 16 CTAs (1024 threads each, <= 4096 rows per CTA):
// the hash table, both sort buffers and the histograms are shared memory, rows move between CTAs with remote
// shared-memory stores (~215 cycles), and the phases are separated by barrier.cluster (~380 cycles) instead of grid
// barriers.  Only 8-16 of the 148 SMs work, but the work is latency bound, not throughput bound.
//
//   P0  load keys / labels, insert the keys into the distributed open-addressing table (one insert per distinct key
//       and warp round: match.any); group id = table slot.  Rows that can form no pair (row_ok = 0, NaN label) go to
//       a trash id that sorts last.  OR-reduce the label bits (compact sort key).
//   P1  sort key = gid << labbits | varying label bits, payload = row.
//   Pk  stable LSD radix passes of <= 9 bits: per-warp match.any ranking, per-CTA digit histograms exchanged
//       through DSMEM, scatter with remote stores.
//   H   heads: group / level starts by max-scan (+ carry across CTAs), exact pair counts, gathers into sorted
//       order (global arrays read by k_pair), per-I-block J ranges.
//   U   occurrence weights folded into the positive-side weights, work-unit records.
//
// Used when K == 1 and B <= 65536; everything else takes the grid-wide kernel.  Both produce the same arrays.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace rn {

namespace cg = cooperative_groups;

constexpr int kCT = 1024;                 // threads per CTA of the cluster kernel
constexpr int kCW = kCT / 32;
constexpr u32 kCRowsMax = 4096;           // rows per CTA
constexpr int kCIpt = kCRowsMax / kCT;    // rows per thread

struct SegcParams {
  u32 B, R, ipt, r_log2;         // rows, rows per CTA (1024 / 2048 / 4096), rows per thread (R / 1024), log2(R)
  u32 idtop;                     // group ids are first-occurrence rows < B <= idtop; idtop = trash rows, idtop + 1 = padding
  u32 capmask, spc_log2;         // table capacity - 1, log2(slots per CTA)
  int gbits;                     // log2(idtop) + 1
  const int64_t* keys; const uint8_t* row_ok;
  Ctl* ctl;
  u32 smem_bytes;
};

// shared-memory carve-up (identical in every CTA of the cluster, so offsets map across ranks)
struct SegcSmem {
  u64* kb[2]; u32* vb[2]; u32* whist; u32* table; u32* pub; u32* gbase; u32* lbase; u32* misc;
  __device__ __forceinline__ SegcSmem(unsigned char* base, u32 R) {
    kb[0] = reinterpret_cast<u64*>(base); kb[1] = kb[0] + R;
    vb[0] = reinterpret_cast<u32*>(kb[1] + R); vb[1] = vb[0] + R;
    whist = vb[1] + R; table = whist;                      // the table is dead before the first pass
    pub = whist + kCW * kBins; gbase = pub + kBins; lbase = gbase + kBins; misc = lbase + kBins;
  }
};
inline u32 segc_smem_bytes(u32 R) { return R * 24u + (kCW * kBins + 3 * kBins + 256) * 4u; }

// Insert row i's key into the distributed table; the slot's value converges to the smallest row holding that key
// (= the group's first-occurrence row: a deterministic group id, identical on every rank of the global mode).
__device__ __forceinline__ u32 segc_insert(cg::cluster_group& cl, u32* table, const SegcParams& S, u64 key, u32 i) {
  u32 s = (u32)mix64(0x9E3779B97F4A7C15ull ^ key) & S.capmask;
  const u32 spcmask = (1u << S.spc_log2) - 1u;
  for (;;) {
    u32* p = cl.map_shared_rank(table, s >> S.spc_log2) + (s & spcmask);
    u32 cur = *reinterpret_cast<volatile u32*>(p);
    if (cur == kEmpty) {
      const u32 prev = atomicCAS(p, kEmpty, i);
      if (prev == kEmpty) return s;
      cur = prev;
    }
    if ((u64)S.keys[cur] == key) { if (i < cur) atomicMin(p, i); return s; }
    s = (s + 1) & S.capmask;
  }
}

// inclusive max-scan over the block of two values with a running carry (sm: [kCW][2] + carry[2])
__device__ __forceinline__ void segc_maxscan2(u32& x, u32& y, u32* sm, u32* carry) {
  const u32 ln = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 tx = __shfl_up_sync(0xFFFFFFFFu, x, o), ty = __shfl_up_sync(0xFFFFFFFFu, y, o);
    if (ln >= (u32)o) { x = max(x, tx); y = max(y, ty); }
  }
  if (ln == 31) { sm[2 * w] = x; sm[2 * w + 1] = y; }
  __syncthreads();
  u32 cx = carry[0], cy = carry[1];
  for (u32 k = 0; k < w; ++k) { cx = max(cx, sm[2 * k]); cy = max(cy, sm[2 * k + 1]); }
  x = max(x, cx); y = max(y, cy);
  __syncthreads();
  if (threadIdx.x == kCT - 1) { carry[0] = x; carry[1] = y; }
  __syncthreads();
}

// HeadsOut: the arrays k_pair reads (same meaning as in HeadsTail)
struct SegcOut {
  PairParams P;
  uint2* aj; float *ss, *sy, *swp, *swn, *gacc, *lossrow; u32 *cnt, *perm;
  uint2* units; u64* cprim; u32 target_units;
};

__global__ void __launch_bounds__(kCT, 1) k_segc_pair(SegcParams S, SegcOut O) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cl = cg::this_cluster();
  const u32 crank = cl.block_rank(), NC = cl.num_blocks();
  const u32 tid = threadIdx.x, w = tid >> 5, ln = tid & 31u;
  const u32 R = S.R, B = S.B, ipt = S.ipt;
  SegcSmem M(smem_raw, R);
  const PairParams& P = O.P;
  Ctl* ctl = S.ctl;

  // ---- init: control block (CTA 0), table, per-CTA words ----------------------------------------------------
  if (crank == 0) {
    u32* cw = reinterpret_cast<u32*>(ctl);
    for (u32 k = tid; k < sizeof(Ctl) / 4; k += kCT) cw[k] = 0;
  }
  for (u32 k = tid; k < (1u << S.spc_log2); k += kCT) M.table[k] = kEmpty;
  if (tid < 64) M.misc[tid] = 0;
  __threadfence();
  cl.sync();
  stamp(ctl, 0);

  // ---- P0: load + hash ----------------------------------------------------------------------------------
  // Two levels of key aggregation keep the remote traffic small: match.any inside the warp round, then a CTA-local
  // table (shared memory; raw keys staged in kb[0]); only one row per distinct key and CTA goes to the distributed
  // table.
  u32* ltab = reinterpret_cast<u32*>(M.kb[1]);            // [2R] CTA-local table: representative local index
  u32* lgid = M.vb[0];                                    // [2R] its global slot, then its group id (vb[0..1] are contiguous)
  const u32 lmask = 2u * R - 1u;                          // R is a power-of-two multiple of 1024 (1024, 2048, 4096)
  u32 lslot[kCIpt]; float lab[kCIpt]; u32 okbits = 0;
  u32 vor = 0, vnor = 0;
  const bool fold_occ = !P.dyn_count && P.power != 0.f;
  for (u32 k = tid; k < 2u * R; k += kCT) ltab[k] = kEmpty;
  u64 rkey[kCIpt];
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    lslot[r] = 0; lab[r] = 0.f; rkey[r] = 0;
    if ((u32)r < ipt) {                                   // uniform
      const u32 idx = r * kCT + tid, i = crank * R + idx;
      if (i < B) {
        rkey[r] = (u64)S.keys[i];
        const float y = P.labels[i];
        lab[r] = y;
        const bool ok = (S.row_ok ? S.row_ok[i] != 0 : true) && !(y != y);
        if (ok) { okbits |= 1u << r; const u32 e = enc_label(y); vor |= e; vnor |= ~e; }
        if (fold_occ) O.cprim[i] = 0;                      // per-group pair totals, indexed by group start position
      }
      M.kb[0][idx] = rkey[r];
    }
  }
  __syncthreads();
  stamp(ctl, 11);
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid;
      const bool ok = (okbits >> r) & 1u;
      const u64 key = rkey[r];
      const u32 okm = __ballot_sync(0xFFFFFFFFu, ok);
      const u32 m = __match_any_sync(0xFFFFFFFFu, key) & okm;
      const u32 leader = m ? (u32)(__ffs(m) - 1) : 0u;
      u32 ls = 0;
      if (ok && ln == leader) {
        ls = (u32)(mix64(0x9E3779B97F4A7C15ull ^ key) >> 40) & lmask;
        for (;;) {
          u32 cur = *reinterpret_cast<volatile u32*>(ltab + ls);
          if (cur == kEmpty) {
            const u32 prev = atomicCAS(ltab + ls, kEmpty, idx);
            if (prev == kEmpty) break;
            cur = prev;
          }
          if (M.kb[0][cur] == key) { if (idx < cur) atomicMin(ltab + ls, idx); break; }
          ls = (ls + 1) & lmask;
        }
      }
      lslot[r] = __shfl_sync(0xFFFFFFFFu, ls, leader);
    }
  }
  __syncthreads();
  stamp(ctl, 12);
  // representatives (smallest local row of each distinct key) insert into the distributed table
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid;
      if (((okbits >> r) & 1u) && ltab[lslot[r]] == idx) lgid[lslot[r]] = segc_insert(cl, M.table, S, rkey[r], crank * R + idx);
    }
  }
  vor = __reduce_or_sync(0xFFFFFFFFu, vor); vnor = __reduce_or_sync(0xFFFFFFFFu, vnor);
  if (ln == 0) { if (vor) atomicOr(&M.misc[0], vor); if (vnor) atomicOr(&M.misc[1], vnor); }
  stamp(ctl, 13);
  __threadfence();
  cl.sync();
  stamp(ctl, 1);
  // group id = the slot's final value (first-occurrence row); fetched once per distinct key and CTA
  const u32 spcmask = (1u << S.spc_log2) - 1u;
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid;
      if (((okbits >> r) & 1u) && ltab[lslot[r]] == idx) {
        const u32 sl = lgid[lslot[r]];
        lgid[lslot[r]] = cl.map_shared_rank(M.table, sl >> S.spc_log2)[sl & spcmask];
      }
    }
  }
  // label bit range over the whole cluster
  if (w == 0) {
    u32 a = 0, b = 0;
    if (ln < NC) { const u32* q = cl.map_shared_rank(M.misc, ln); a = q[0]; b = q[1]; }
    a = __reduce_or_sync(0xFFFFFFFFu, a); b = __reduce_or_sync(0xFFFFFFFFu, b);
    if (ln == 0) { M.misc[2] = a; M.misc[3] = b; }
  }
  __syncthreads();
  const Plan pl = make_plan(M.misc[2], M.misc[3], S.gbits, true);
  if (crank == 0 && tid == 0) { ctl->lab_or = M.misc[2]; ctl->lab_nor = M.misc[3]; }

  // ---- P1: sort keys (ids >= idtop cannot pair: trash rows and padding; they sort last) -------------------
  {
    const u32 labmask = pl.labbits >= 32 ? 0xFFFFFFFFu : ((1u << pl.labbits) - 1u);
    u32 gid[kCIpt];
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      gid[r] = S.idtop + 1u;
      if ((u32)r < ipt) {
        const u32 i = crank * R + r * kCT + tid;
        gid[r] = ((okbits >> r) & 1u) ? lgid[lslot[r]] : (i < B ? S.idtop : S.idtop + 1u);
      }
    }
    stamp(ctl, 14);
    cl.sync();                 // every remote table read is done: whist (aliases the table), kb and vb may be reused
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      if ((u32)r < ipt) {
        const u32 idx = r * kCT + tid;
        const u32 lb = gid[r] < S.idtop ? ((enc_label(lab[r]) >> pl.labshift) & labmask) : 0u;
        M.kb[0][idx] = ((u64)gid[r] << pl.labbits) | lb;
        M.vb[0][idx] = crank * R + idx;
      }
    }
  }
  __syncthreads();
  stamp(ctl, 15);

  // ---- radix passes -----------------------------------------------------------------------------------------
  const u32 RW = R / kCW;                 // rows per warp (multiple of 32)
  for (int pass = 0; pass < pl.npass; ++pass) {
    const int src = pass & 1, dst = src ^ 1;
    const int shift = pl.shift[pass];
    const u32 nb = 1u << pl.nbits[pass], dmask = nb - 1u;
    for (u32 k = tid; k < kCW * kBins; k += kCT) M.whist[k] = 0;
    __syncthreads();
    u64 key[kCIpt]; u32 val[kCIpt], rank[kCIpt];
    u32* wh = M.whist + w * kBins;
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      if ((u32)r < ipt) {
        const u32 idx = w * RW + r * 32 + ln;
        key[r] = M.kb[src][idx]; val[r] = M.vb[src][idx];
        const u32 d = (u32)(key[r] >> shift) & dmask;
        const u32 m = __match_any_sync(0xFFFFFFFFu, d);
        const u32 leader = __ffs(m) - 1;
        u32 old = 0;
        if (ln == leader) { old = wh[d]; wh[d] = old + __popc(m); }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rank[r] = old + __popc(m & lanemask_lt());
        __syncwarp();
      }
    }
    __syncthreads();
    if (pass == 0) stamp(ctl, 5);
    if (tid < kBins) {
      u32 sum = 0;
#pragma unroll 8
      for (int k = 0; k < kCW; ++k) { const u32 x = M.whist[k * kBins + tid]; M.whist[k * kBins + tid] = sum; sum += x; }
      M.pub[tid] = sum;
    }
    if (pass == 0) stamp(ctl, 6);
    cl.sync();
    if (pass == 0) stamp(ctl, 7);
    {
      u32 before = 0, all = 0, mine = 0;
      if (tid < nb) {
        for (u32 c2 = 0; c2 < NC; ++c2) {
          const u32 x = cl.map_shared_rank(M.pub, c2)[tid];
          all += x; if (c2 < crank) before += x; if (c2 == crank) mine = x;
        }
      }
      // exclusive scans over the bins (tid < 512: warps 0..15): global totals -> destination base, own counts -> local base
      u32 inc = all, linc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o), y = __shfl_up_sync(0xFFFFFFFFu, linc, o);
        if (ln >= (u32)o) { inc += x; linc += y; }
      }
      if (ln == 31 && w < kBins / 32) { M.misc[16 + w] = inc; M.misc[32 + w] = linc; }
      __syncthreads();
      if (tid < kBins) {
        u32 woff = 0, lwoff = 0;
        for (u32 k = 0; k < w; ++k) { woff += M.misc[16 + k]; lwoff += M.misc[32 + k]; }
        const u32 lb = lwoff + linc - mine;
        M.lbase[tid] = lb;
        M.gbase[tid] = woff + inc - all + before - lb;      // destination position = gbase[d] + local sorted index
      }
    }
    __syncthreads();
    if (pass == 0) stamp(ctl, 8);
    // local reorder by digit (the source buffers are free: every row is in registers), then rows that are neighbours
    // in the CTA's digit order are neighbours at the destination: the remote stores of a warp coalesce into runs
    u64* ks = M.kb[src]; u32* vs = M.vb[src];
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      if ((u32)r < ipt) {
        const u32 d = (u32)(key[r] >> shift) & dmask;
        const u32 li = M.lbase[d] + wh[d] + rank[r];
        ks[li] = key[r]; vs[li] = val[r];
      }
    }
    __syncthreads();
    if (pass == 0) stamp(ctl, 9);
    u64* kd = M.kb[dst]; u32* vd = M.vb[dst];
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      if ((u32)r < ipt) {
        const u32 li = r * kCT + tid;
        const u64 k = ks[li]; const u32 v = vs[li];
        const u32 pos = M.gbase[(u32)(k >> shift) & dmask] + li;
        const u32 dc = pos >> S.r_log2, off = pos & (R - 1u);
        cl.map_shared_rank(kd, dc)[off] = k;
        cl.map_shared_rank(vd, dc)[off] = v;
      }
    }
    if (pass == 0) stamp(ctl, 10);
    cl.sync();
    stamp(ctl, 2 + pass);
  }

  // ---- heads ------------------------------------------------------------------------------------------------
  const int fin = pl.npass & 1;
  const u64* key = M.kb[fin]; const u32* val = M.vb[fin];
  u32* sm_scan = M.misc + 64;              // [kCW][2]
  u32* sm_carry = M.misc + 16;             // [2] running carry; [2..3] published last head / level (+1) of this CTA
  u32* sm_wj = M.whist;                    // [kCIpt][kCW][2]  (the histograms are dead)
  uint2* sblk = reinterpret_cast<uint2*>(M.gbase);      // [R / 64] J range of this CTA's I-blocks
  u32* sm_red = M.pub;                     // scratch
  if (tid == 0) {
    sm_carry[0] = 0; sm_carry[1] = 0;
    // last key of the previous CTA (first-row head test)
    const u64 pk = crank ? cl.map_shared_rank(const_cast<u64*>(key), crank - 1)[R - 1] : ~0ull;
    M.misc[8] = (u32)pk; M.misc[9] = (u32)(pk >> 32);
  }
  __syncthreads();
  const u64 prev_cta_key = ((u64)M.misc[9] << 32) | M.misc[8];
  u32 xa[kCIpt], xl[kCIpt]; u64 kk[kCIpt];
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    xa[r] = xl[r] = 0; kk[r] = ~0ull;
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid, p = crank * R + idx;
      const u64 k = key[idx];
      const u64 kp = idx ? key[idx - 1] : prev_cta_key;
      kk[r] = k;
      const bool head = p == 0 || (k >> pl.labbits) != (kp >> pl.labbits);
      const bool lvl = head || k != kp;
      xa[r] = head ? p + 1 : 0; xl[r] = lvl ? p + 1 : 0;
      segc_maxscan2(xa[r], xl[r], sm_scan, sm_carry);
    }
  }
  if (tid == 0) { sm_carry[2] = sm_carry[0]; sm_carry[3] = sm_carry[1]; }
  cl.sync();
  // carry-in: the last head / level start of the CTAs before this one
  if (w == 0) {
    u32 a = 0, b = 0;
    if (ln < crank) { const u32* q = cl.map_shared_rank(sm_carry, ln); a = q[2]; b = q[3]; }
    a = warp_max(a); b = warp_max(b);
    if (ln == 0) { M.misc[10] = a; M.misc[11] = b; }
  }
  __syncthreads();
  const u32 cin_a = M.misc[10], cin_l = M.misc[11];
  u64 tot_pairs = 0; u32 my_tiles = 0;
  u32 a_[kCIpt], n_[kCIpt], row_[kCIpt]; float wp_[kCIpt], wn_[kCIpt], s_[kCIpt], y_[kCIpt];
  // gathers of all rounds first (read-only inputs: the loads are independent of the stores below)
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    a_[r] = 0; n_[r] = 0; row_[r] = 0; wp_[r] = 1.f; wn_[r] = 1.f; s_[r] = 0.f; y_[r] = 0.f;
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid, p = crank * R + idx;
      if (p < B) {
        const u32 row = val[idx];
        row_[r] = row;
        s_[r] = __ldg(P.logits + row); y_[r] = __ldg(P.labels + row);
        if (P.rw_pos) wp_[r] = __ldg(P.rw_pos + row);
        if (P.rw_neg) wn_[r] = __ldg(P.rw_neg + row);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kCIpt; ++r) {
    if ((u32)r < ipt) {
      const u32 idx = r * kCT + tid, p = crank * R + idx;
      const bool in = p < B;
      const u32 a = (xa[r] ? xa[r] : cin_a) - 1u, l = (xl[r] ? xl[r] : cin_l) - 1u;
      const bool real = in && (u32)(kk[r] >> pl.labbits) < S.idtop;
      u32 n = real ? l - a : 0u;
      if (P.rw_pos && !(wp_[r] > 0.f)) n = 0;                              // PW:193  C = W > 0
      if (in) {
        O.aj[p] = make_uint2(a, n);
        O.ss[p] = s_[r];
        O.sy[p] = y_[r];
        if (P.rw_pos && !fold_occ) O.swp[p] = wp_[r];
        if (P.rw_neg) O.swn[p] = wn_[r];
        O.gacc[p] = 0.f; O.perm[p] = row_[r];
        if (P.dyn_count) { O.lossrow[p] = 0.f; O.cnt[p] = 0; }
      }
      a_[r] = a; n_[r] = n;
      // J range of this warp's 32 rows; two warps make one I-block (R is a multiple of 64)
      const u32 jlo = warp_min(n ? a : 0xFFFFFFFFu), jhi = warp_max(n ? a + n : 0u);
      if (ln == 0) { sm_wj[2 * (r * kCW + w)] = jlo; sm_wj[2 * (r * kCW + w) + 1] = jhi; }
      u32 cn = 0;
      if (!P.dyn_count) {
        cn = n;
        if (in && P.row_pairs) P.row_pairs[row_[r]] = (int64_t)cn;
        if (fold_occ) {                                                  // per-group totals (PW:286-289), K = 1
          const u32 g = cn ? a : kEmpty;
          const u32 m = __match_any_sync(0xFFFFFFFFu, g);
          const u32 tot = __reduce_add_sync(m, cn);
          if (g != kEmpty && ln == (u32)(__ffs(m) - 1)) atomicAdd(O.cprim + g, (u64)tot);
        }
      }
      tot_pairs += cn;
    }
  }
  __syncthreads();
  if (tid < R / kIB) {                                   // I-block tid = rows [64 tid, 64 tid + 64) = round tid / 16, warps 2 (tid % 16), +1
    const u32 q = 2 * ((tid >> 4) * kCW + 2 * (tid & 15u));
    const u32 lo = min(sm_wj[q], sm_wj[q + 2]), hi = max(sm_wj[q + 1], sm_wj[q + 3]);
    sblk[tid] = make_uint2(lo, hi);
    my_tiles = hi > lo ? ((hi + 31) >> 5) - (lo >> 5) : 0u;
  }
  // CTA totals: pairs -> ctl->n_pair, tiles -> published for the work list
  tot_pairs = warp_sum(tot_pairs); my_tiles = warp_sum(my_tiles);
  u64* red64 = reinterpret_cast<u64*>(sm_red);
  if (ln == 0) { red64[w] = tot_pairs; sm_red[128 + w] = my_tiles; }
  __syncthreads();
  if (tid == 0) {
    u64 t = 0; u32 mt = 0;
    for (int q = 0; q < kCW; ++q) { t += red64[q]; mt += sm_red[128 + q]; }
    if (t) atomicAdd(&ctl->n_pair, t);
    M.misc[12] = mt;
  }
  __threadfence();
  cl.sync();
  stamp(ctl, 17);

  // ---- occurrence weights + work list --------------------------------------------------------------------------
  if (fold_occ) {
#pragma unroll
    for (int r = 0; r < kCIpt; ++r) {
      if ((u32)r < ipt) {
        const u32 p = crank * R + r * kCT + tid;
        if (p < B) {
          const u64 ch = n_[r] ? __ldcg(O.cprim + a_[r]) : 0ull;
          const float wocc = ch ? ((P.power == 1.0f) ? (float)ch : powf((float)ch, P.power)) : 0.f;   // PW:147-149
          O.swp[p] = wp_[r] * wocc;
        }
      }
    }
  }
  if (w == 0) {
    u32 mt = ln < NC ? cl.map_shared_rank(M.misc, ln)[12] : 0u;
    mt = warp_sum(mt);
    if (ln == 0) M.misc[13] = mt;
  }
  __syncthreads();
  const u32 Mt = M.misc[13];
  u32 C = (Mt + O.target_units - 1) / O.target_units;
  C = C < 1 ? 1 : (C > 64u ? 64u : C);
  const u32 nibc = R / kIB;                       // I-blocks of this CTA (<= 64)
  u32 nt = 0, v = 0, inc = 0;
  if (tid < 64) {
    if (tid < nibc) { const uint2 bj = sblk[tid]; nt = bj.y > bj.x ? ((bj.y + 31) >> 5) - (bj.x >> 5) : 0u; }
    v = (nt + C - 1) / C;
    inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
    if (ln == 31) M.misc[20 + w] = inc;
  }
  __syncthreads();
  if (tid < 64 && w == 1) inc += M.misc[20];
  if (tid == 63) M.misc[14] = inc;                 // units of this CTA
  cl.sync();
  if (w == 0) {
    u32 u = ln < NC ? cl.map_shared_rank(M.misc, ln)[14] : 0u;
    const u32 before = warp_sum(ln < crank ? u : 0u), total = warp_sum(u);
    if (ln == 0) { M.misc[15] = before; if (crank == 0) { ctl->n_units = total; ctl->unit_c = C; ctl->n_tiles = Mt; } }
  }
  __syncthreads();
  if (tid < 64) { M.lbase[3 * tid] = v; M.lbase[3 * tid + 1] = inc - v; M.lbase[3 * tid + 2] = nt; }
  __syncthreads();
  for (u32 ib = w; ib < nibc; ib += kCW) {           // a warp per I-block: its units are written 32 at a time
    const u32 vv = M.lbase[3 * ib], off = M.misc[15] + M.lbase[3 * ib + 1], ntt = M.lbase[3 * ib + 2];
    const u32 jfirst = sblk[ib].x >> 5;
    for (u32 q = ln; q < vv; q += 32) O.units[off + q] = make_uint2(crank * nibc + ib, (jfirst + q * C) | (min(C, ntt - q * C) << 24));
  }
  stamp(ctl, 19);
  cl.sync();                                       // no CTA may exit while its shared memory can still be read
}

}  // namespace rn
