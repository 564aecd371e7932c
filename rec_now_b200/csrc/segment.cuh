// K1 -- group segmentation as ONE persistent cooperative kernel (device code; instantiated per tail by
// pairwise.cu / listwise.cu / pairs.cu).
//
// Replaces the reference's (B,B) group-equality matrix (pairwise_loss_from_batch.py:33-37, 68-73) and
// tf.unique_with_counts (listwise_loss_from_batch.py:109, pairwise_loss_from_batch.py:146).
//
// The batch is ~1.5 MB (L2 resident), so every step is latency bound, not bandwidth bound: a chain of small
// launches costs more than the work.  k_seg therefore runs the whole chain in one launch, one CTA per SM, with
// device-side grid barriers (~1 us) between the phases:
//   hash    open-addressing insert of the (composite) int64 keys; the table slot ends up holding the MINIMUM row
//           of its group = the group's first-occurrence row id: a dense ceil(log2 B)-bit, deterministic group id
//           that is also the order the listwise path must report (LW:109).  Rows that can form no pair
//           (row_ok = 0, NaN label) become singleton groups (gid = own row) -- no compaction pass.
//   vkey    compact sort key (gid << labbits | varying label bits), row index payload, global digit histograms of
//           all passes, per-tile histogram of pass 0.
//   sort    stable LSD radix passes of 9 bits: per-warp match.any ranking, tile prefix = sum of the earlier
//           tiles' histograms (no look-back chain), direct scatter; the per-tile histogram of the NEXT pass is
//           accumulated during the scatter, so a pass needs exactly one grid barrier.
//   tail    consumer-specific: group / label-level starts + gathers + work list (pairwise), or group bounds +
//           gathers (listwise, pair materialisation).
#pragma once
#include "common.cuh"
#include "group_count.cuh"

namespace rn {

struct SegParams {
  u32 B; int K; int gbits; int use_label; int nan_trash;
  RowMap rm; u32 kpitch;          // input row layout; elements between key columns (B, or the block rows when blocked)
  const int64_t* keys; const uint8_t* row_ok; const float* labels;
  u32 capmask, ntiles;
  u32 *table, *table1, *slot, *slot1;
  u64 *keyA, *keyB; u32 *valA, *valB;
  u32 *hist, *tilehist, *th0;
  Ctl* ctl;
  // merged first phase (single key column, contiguous rows, no cross-rank determinism needed): the label bit range
  // comes from k_init's per-CTA partials, the group id is the table SLOT (known right after the insert), so the sort
  // keys and histograms are built in the hash phase and one phase + grid barrier disappear.  Rows that cannot pair
  // get the id `cap` (one group, one label level: no pairs), which sorts last.
  int merged; const u32* labpart; int npart;
  u64* dbgts;                     // RN_SEG_DEBUG=1: per-CTA arrival stamps [phase][cta] (dev tool), else nullptr
  // counting path (group_count.cuh; fast != 0: no k_init ran, the arena is in its clean state)
  int fast;
  GRec* rec; GRec2* rec2; u32 *glist, *gcount, *rslot, *rmeta;
  uint4 *init_zero, *init_ones; u32 init_zero16, init_ones16;    // regions the radix path needs initialised (fallback)
  __device__ __forceinline__ u32* th_buf(u32 k) const { return k ? tilehist + (size_t)(k - 1) * ntiles * kBins : th0; }
};

constexpr int kSegSmemWords = kSegWarps * kBins + kBins + 64;

// Open-addressing insert of row i's key tuple (nk columns); the slot's value converges to the smallest row
// holding that tuple (any representative has the same key, so probing stays consistent while it changes).
__device__ __forceinline__ u32 hash_insert(const SegParams& S, int nk, u32 i, u32* table) {
  const int64_t* keys = S.keys; const u32 capmask = S.capmask;
  const size_t oi = S.rm.i8(i);
  u64 h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < nk; ++k) h = mix64(h ^ (u64)keys[(size_t)k * S.kpitch + oi]);
  u32 s = (u32)h & capmask;
  for (;;) {
    u32 cur = ld_relaxed(table + s);
    if (cur == kEmpty) {
      const u32 prev = atomicCAS(table + s, kEmpty, i);
      if (prev == kEmpty) return s;
      cur = prev;
    }
    bool eq = true;
    const size_t oc = S.rm.i8(cur);
    for (int k = 0; k < nk; ++k) eq = eq && (keys[(size_t)k * S.kpitch + oc] == keys[(size_t)k * S.kpitch + oi]);
    if (eq) {
      if (i < cur) atomicMin(table + s, i);
      return s;
    }
    s = (s + 1) & capmask;
  }
}

// One hash pass over this CTA's 512-row chunks.  Rows are first grouped inside the CTA (shared-memory table keyed by
// the same tuple), so only ONE row per distinct key and chunk -- the smallest -- touches the global table: a hot
// key (Zipf head: >10% of the rows) costs one global atomic per CTA instead of one per row.
__device__ __forceinline__ void seg_hash_pass(const SegParams& S, int nk, u32* table, u32* slot_out, bool first,
                                              u32* smem, u32& vor, u32& vnor) {
  constexpr u32 kLoc = 2 * kSegThreads;               // local table size
  u32* sm_tab = smem;                                 // [kLoc] representative (local thread id)
  u32* sm_gslot = smem + kLoc;                        // [kLoc] its global slot
  const u32 tid = threadIdx.x;
  const u32 nchunks = (S.B + kSegThreads - 1) / kSegThreads;
  for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const u32 r0 = c * kSegThreads, i = r0 + tid;
    sm_tab[tid] = kEmpty; sm_tab[tid + kSegThreads] = kEmpty;
    __syncthreads();
    bool ok = false;
    u32 ls = 0;
    if (i < S.B) {
      const float y = S.labels ? S.labels[S.rm.i4(i)] : 0.f;
      ok = (S.row_ok ? S.row_ok[S.rm.i1(i)] != 0 : true) && !(S.nan_trash && (y != y));
      if (ok) {
        if (first) { const u32 e = enc_label(y); vor |= e; vnor |= ~e; }
        u64 h = 0x9E3779B97F4A7C15ull;
        const size_t oi = S.rm.i8(i);
        for (int k = 0; k < nk; ++k) h = mix64(h ^ (u64)S.keys[(size_t)k * S.kpitch + oi]);
        ls = (u32)(h >> 40) & (kLoc - 1);
        for (;;) {
          u32 cur = sm_tab[ls];
          if (cur == kEmpty) {
            const u32 prev = atomicCAS(&sm_tab[ls], kEmpty, tid);
            if (prev == kEmpty) break;
            cur = prev;
          }
          bool eq = true;
          const size_t oc = S.rm.i8(r0 + cur);
          for (int k = 0; k < nk; ++k) eq = eq && (S.keys[(size_t)k * S.kpitch + oc] == S.keys[(size_t)k * S.kpitch + oi]);
          if (eq) { if (tid < cur) atomicMin(&sm_tab[ls], tid); break; }
          ls = (ls + 1) & (kLoc - 1);
        }
      }
    }
    __syncthreads();
    if (ok && sm_tab[ls] == tid) sm_gslot[ls] = hash_insert(S, nk, i, table);
    __syncthreads();
    if (i < S.B) slot_out[i] = ok ? sm_gslot[ls] : kEmpty;
    __syncthreads();
  }
}

__device__ __forceinline__ void seg_hash(const SegParams& S, u32* smem) {
  const u32 gtid = blockIdx.x * kSegThreads + threadIdx.x, gthreads = gridDim.x * kSegThreads;
  u32 vor = 0, vnor = 0;
  seg_hash_pass(S, S.K, S.table, S.slot, true, smem, vor, vnor);
  if (S.K > 1) seg_hash_pass(S, 1, S.table1, S.slot1, false, smem, vor, vnor);
  vor = __reduce_or_sync(0xFFFFFFFFu, vor);
  vnor = __reduce_or_sync(0xFFFFFFFFu, vnor);
  if (lane_id() == 0 && S.use_label) {
    if (vor) atomicOr(&S.ctl->lab_or, vor);
    if (vnor) atomicOr(&S.ctl->lab_nor, vnor);
  }
  // tile-histogram buffer 1 is accumulated during pass 0: clear it here (buffer 0 is stored by vkey, buffer 2 is
  // cleared during pass 0)
  u32* th1 = S.th_buf(1);
  for (u32 k = gtid; k < S.ntiles * kBins; k += gthreads) th1[k] = 0;
}

// Merged first phase: hash + sort keys + histograms in one pass over this CTA's 512-row chunks (see SegParams::merged).
template <int IPT>
__device__ __forceinline__ void seg_hash_keys(const SegParams& S, const Plan& pl, u32* smem) {
  constexpr u32 kLoc = 2 * kSegThreads;
  constexpr u32 T = kSegThreads * IPT;
  u32* sm_tab = smem;                                 // [kLoc] representative (local thread id)
  u32* sm_gslot = smem + kLoc;                        // [kLoc] its global slot
  u32* sm_hist = smem + 2 * kLoc;                     // [npass][kBins]
  const u32 tid = threadIdx.x;
  const u32 gtid = blockIdx.x * kSegThreads + tid, gthreads = gridDim.x * kSegThreads;
  const u32 labmask = pl.labbits >= 32 ? 0xFFFFFFFFu : ((1u << pl.labbits) - 1u);
  const u32 trash = S.capmask + 1u;
  const u32 nchunks = (S.B + kSegThreads - 1) / kSegThreads;
  for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const u32 r0 = c * kSegThreads, i = r0 + tid;
    sm_tab[tid] = kEmpty; sm_tab[tid + kSegThreads] = kEmpty;
    for (u32 k = tid; k < (u32)pl.npass * kBins; k += kSegThreads) sm_hist[k] = 0;
    __syncthreads();
    bool ok = false;
    u32 ls = 0; float y = 0.f;
    if (i < S.B) {
      y = S.labels[i];
      ok = (S.row_ok ? S.row_ok[i] != 0 : true) && !(y != y);
      if (ok) {
        const u64 key = (u64)S.keys[i];
        ls = (u32)(mix64(0x9E3779B97F4A7C15ull ^ key) >> 40) & (kLoc - 1);
        for (;;) {
          u32 cur = sm_tab[ls];
          if (cur == kEmpty) {
            const u32 prev = atomicCAS(&sm_tab[ls], kEmpty, tid);
            if (prev == kEmpty) break;
            cur = prev;
          }
          if ((u64)S.keys[r0 + cur] == key) { if (tid < cur) atomicMin(&sm_tab[ls], tid); break; }
          ls = (ls + 1) & (kLoc - 1);
        }
      }
    }
    __syncthreads();
    if (ok && sm_tab[ls] == tid) sm_gslot[ls] = hash_insert(S, 1, i, S.table);
    __syncthreads();
    if (i < S.B) {
      const u32 gid = ok ? sm_gslot[ls] : trash;
      const u32 lab = ok ? ((enc_label(y) >> pl.labshift) & labmask) : 0u;
      const u64 key = ((u64)gid << pl.labbits) | lab;
      S.keyA[i] = key; S.valA[i] = i;
#pragma unroll 1
      for (int p = 0; p < pl.npass; ++p)
        atomicAdd(&sm_hist[p * kBins + (u32)((key >> pl.shift[p]) & ((1u << pl.nbits[p]) - 1u))], 1u);
    }
    __syncthreads();
    for (u32 k = tid; k < (u32)pl.npass * kBins; k += kSegThreads)
      if (sm_hist[k]) atomicAdd(S.hist + k, sm_hist[k]);
    if (sm_hist[tid]) atomicAdd(S.th0 + (size_t)(r0 / T) * kBins + tid, sm_hist[tid]);   // pass-0 histogram of the row's tile
    __syncthreads();
  }
  u32* th1 = S.th_buf(1);
  for (u32 k = gtid; k < S.ntiles * kBins; k += gthreads) th1[k] = 0;
}

template <int IPT>
__device__ __forceinline__ void seg_vkey(const SegParams& S, const Plan& pl, u32* smem) {
  constexpr u32 T = kSegThreads * IPT;
  const u32 tid = threadIdx.x;
  const u32 labmask = pl.labbits >= 32 ? 0xFFFFFFFFu : ((1u << pl.labbits) - 1u);
  for (u32 t = blockIdx.x; t < S.ntiles; t += gridDim.x) {
    for (u32 k = tid; k < (u32)pl.npass * kBins; k += kSegThreads) smem[k] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const u32 i = t * T + r * kSegThreads + tid;
      if (i < S.B) {
        const u32 s = S.slot[i];
        const u32 gid = (s == kEmpty) ? i : S.table[s];
        u32 lab = 0;
        if (S.use_label && s != kEmpty) lab = (enc_label(S.labels[S.rm.i4(i)]) >> pl.labshift) & labmask;
        const u64 key = ((u64)gid << pl.labbits) | lab;
        S.keyA[i] = key; S.valA[i] = i;
        if (S.K > 1) { const u32 s1 = S.slot1[i]; S.slot1[i] = (s1 == kEmpty) ? i : S.table1[s1]; }   // primary gid
#pragma unroll 1
        for (int p = 0; p < pl.npass; ++p)
          atomicAdd(&smem[p * kBins + (u32)((key >> pl.shift[p]) & ((1u << pl.nbits[p]) - 1u))], 1u);
      }
    }
    __syncthreads();
    for (u32 k = tid; k < (u32)pl.npass * kBins; k += kSegThreads)
      if (smem[k]) atomicAdd(S.hist + k, smem[k]);
    S.th0[(size_t)t * kBins + tid] = smem[tid];                // pass-0 histogram of this tile (buffer 0)
    __syncthreads();
  }
}

template <int IPT>
__device__ __forceinline__ void seg_sort_pass(const SegParams& S, const Plan& pl, int pass, u32* smem) {
  constexpr u32 T = kSegThreads * IPT;
  u32 (*whist)[kBins] = reinterpret_cast<u32 (*)[kBins]>(smem);
  u32* gbase = smem + kSegWarps * kBins;
  u32* wsum = gbase + kBins;
  const u64* ksrc = (pass & 1) ? S.keyB : S.keyA; const u32* vsrc = (pass & 1) ? S.valB : S.valA;
  u64* kdst = (pass & 1) ? S.keyA : S.keyB;       u32* vdst = (pass & 1) ? S.valA : S.valB;
  const int shift = pl.shift[pass];
  const u32 nb = 1u << pl.nbits[pass], dmask = nb - 1u;
  const bool more = pass + 1 < pl.npass;
  const int nshift = more ? pl.shift[pass + 1] : 0;
  const u32 ndmask = more ? ((1u << pl.nbits[pass + 1]) - 1u) : 0u;
  const u32* th_cur = S.th_buf(pass % 3);
  u32* th_next = S.th_buf((pass + 1) % 3);
  u32* th_zero = S.th_buf((pass + 2) % 3);
  const u32 tid = threadIdx.x, w = tid >> 5, ln = tid & 31u;
  for (u32 t = blockIdx.x; t < S.ntiles; t += gridDim.x) {
    for (u32 k = tid; k < kSegWarps * kBins; k += kSegThreads) smem[k] = 0;
    __syncthreads();
    const u32 base = t * T + w * (32 * IPT);
    u64 key[IPT]; u32 val[IPT]; u32 rank[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const u32 idx = base + r * 32 + ln;
      const bool v = idx < S.B;
      key[r] = v ? ksrc[idx] : ~0ull; val[r] = v ? vsrc[idx] : 0u;
      const u32 d = v ? (u32)((key[r] >> shift) & dmask) : 0xFFFFu;
      const u32 m = __match_any_sync(0xFFFFFFFFu, d);
      const u32 leader = __ffs(m) - 1;
      u32 old = 0;
      if (v && ln == leader) { old = whist[w][d]; whist[w][d] = old + __popc(m); }
      old = __shfl_sync(0xFFFFFFFFu, old, leader);
      rank[r] = old + __popc(m & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    // bin = tid: exclusive scan of the global digit histogram, prefix over the earlier tiles, prefix over warps
    const u32 hv = (tid < nb) ? S.hist[pass * kBins + tid] : 0u;
    u32 inc = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
    if (ln == 31) wsum[w] = inc;
    u32 tp = 0;
    if (tid < nb) {
      u32 t2 = 0;
      // all loads of a batch are independent: the deeper the batch, the fewer dependent L2 round trips (the prefix
      // over the earlier tiles is the longest chain of a pass)
      for (; t2 + 32 <= t; t2 += 32) {
        u32 x[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = th_cur[(size_t)(t2 + k) * kBins + tid];
#pragma unroll
        for (int k = 0; k < 32; ++k) tp += x[k];
      }
      if (t2 + 16 <= t) {
        u32 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = th_cur[(size_t)(t2 + k) * kBins + tid];
#pragma unroll
        for (int k = 0; k < 16; ++k) tp += x[k];
        t2 += 16;
      }
      if (t2 + 8 <= t) {
        u32 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = th_cur[(size_t)(t2 + k) * kBins + tid];
#pragma unroll
        for (int k = 0; k < 8; ++k) tp += x[k];
        t2 += 8;
      }
      for (; t2 < t; ++t2) tp += th_cur[(size_t)t2 * kBins + tid];
    }
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < kSegWarps; ++k) { const u32 x = whist[k][tid]; whist[k][tid] = sum; sum += x; }
    th_zero[(size_t)t * kBins + tid] = 0;
    __syncthreads();
    u32 woff = 0;
    for (u32 k = 0; k < w; ++k) woff += wsum[k];
    gbase[tid] = woff + inc - hv + tp;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const u32 idx = base + r * 32 + ln;
      if (idx < S.B) {
        const u32 d = (u32)((key[r] >> shift) & dmask);
        const u32 pos = gbase[d] + whist[w][d] + rank[r];
        kdst[pos] = key[r]; vdst[pos] = val[r];
        if (more) atomicAdd(&th_next[(size_t)(pos / T) * kBins + (u32)((key[r] >> nshift) & ndmask)], 1u);
      }
    }
    __syncthreads();
  }
}

// Warp-cooperative 32-ary lower bound over the sorted keys: first q in [0, t0] with (key[q] >> sh) >= (key[t0] >> sh).
__device__ __forceinline__ u32 coop_lower_bound(const u64* key, u32 t0, int sh) {
  const u32 ln = lane_id();
  const u64 g0 = key[t0] >> sh;
  u32 lo = 0, hi = t0;                   // answer in [lo, hi]; key[hi] >> sh >= g0
  while (hi > lo) {
    const u32 span = hi - lo, step = (span + 30) / 31;          // lane 31 always probes hi
    const u32 q = lo + min(ln * step, span);
    const bool ge = (key[q] >> sh) >= g0;
    const u32 bal = __ballot_sync(0xFFFFFFFFu, ge);
    const u32 f = __ffs(bal) - 1;        // first probe that is >= g0
    const u32 nhi = lo + min(f * step, span);
    const u32 nlo = f ? lo + min((f - 1) * step, span) + 1 : lo;
    hi = nhi; lo = min(nlo, nhi);
  }
  return lo;
}

// inclusive max-scan over the kSegThreads threads of the block of two values at once (sm: [kSegWarps][2])
__device__ __forceinline__ void block_maxscan2(u32& x, u32& y, u32* sm) {
  const u32 ln = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 tx = __shfl_up_sync(0xFFFFFFFFu, x, o), ty = __shfl_up_sync(0xFFFFFFFFu, y, o);
    if (ln >= (u32)o) { x = max(x, tx); y = max(y, ty); }
  }
  if (ln == 31) { sm[2 * w] = x; sm[2 * w + 1] = y; }
  __syncthreads();
  u32 cx = 0, cy = 0;
  for (u32 k = 0; k < w; ++k) { cx = max(cx, sm[2 * k]); cy = max(cy, sm[2 * k + 1]); }
  x = max(x, cx); y = max(y, cy);
}

// ---- tail: group bounds + gathers (listwise path and pair materialisation) -----------------------------
// astart[p] = first sorted position of p's group, gend[astart] = one past its last position, perm[p] = original
// row; up to 4 float columns are gathered into sorted order.
struct GatherCols { const float* src[4]; float* dst[4]; };

struct BoundsTail {
  static constexpr bool kFast = false;
  u32 *astart, *gend, *perm; GatherCols gc;
  __device__ __forceinline__ void run(const SegParams& S, const Plan& pl, const u64* key, const u32* val, u32* smem, u32& epoch) const {
    u32* sm_scan = smem;            // [kSegWarps][2]
    u32* sm_carry = smem + 2 * kSegWarps;
    const u32 ln = lane_id(), w = threadIdx.x >> 5;
    const u32 nchunks = (S.B + kSegThreads - 1) / kSegThreads;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const u32 t0 = c * kSegThreads, p = t0 + threadIdx.x;
      const bool in = p < S.B;
      const u64 k = in ? key[p] : ~0ull;
      const u64 gid = k >> pl.labbits;
      const u64 gprev = (in && p > 0) ? (key[p - 1] >> pl.labbits) : ~gid;
      const u64 gnext = (in && p + 1 < S.B) ? (key[p + 1] >> pl.labbits) : ~gid;
      if (w == 0) { const u32 lb = coop_lower_bound(key, t0, pl.labbits); if (ln == 0) sm_carry[0] = lb; }
      u32 x = (in && gid != gprev) ? p + 1 : 0u, dummy = 0;
      block_maxscan2(x, dummy, sm_scan);
      if (in) {
        const u32 a = x ? x - 1 : sm_carry[0];
        astart[p] = a;
        if (gid != gnext) gend[a] = p + 1;
        const u32 row = val[p];
        perm[p] = row;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (gc.src[q]) gc.dst[q][p] = gc.src[q][row];
      }
      __syncthreads();
    }
  }
};

// ---- the kernel ----------------------------------------------------------------------------------------
template <int IPT, class Tail>
__global__ void __launch_bounds__(kSegThreads, 1) k_seg(SegParams S, Tail T) {
  __shared__ __align__(16) u32 smem[kSegSmemWords];
  Ctl* ctl = S.ctl;
  u32 epoch = 0;
  grid_dep_wait();
  grid_dep_launch();
  stamp(ctl, 0);
  if constexpr (Tail::kFast) {
    if (S.fast) {
      // counting path; false = something outside its menu was seen: initialise what the radix path needs and go on
      if (T.count_run(S, smem, epoch)) { stamp(ctl, 19); return; }
      const u32 gtid = blockIdx.x * kSegThreads + threadIdx.x, gthreads = gridDim.x * kSegThreads;
      const uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(~0u, ~0u, ~0u, ~0u);
      for (u32 k = gtid; k < S.init_zero16; k += gthreads) S.init_zero[k] = z;
      for (u32 k = gtid; k < S.init_ones16; k += gthreads) S.init_ones[k] = f;
      grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->path = 2;
  Plan pl;
  if (S.merged) {
    // label bit range from k_init's partials
    u32 vor = 0, vnor = 0;
    for (int k = threadIdx.x; k < S.npart; k += kSegThreads) { vor |= S.labpart[2 * k]; vnor |= S.labpart[2 * k + 1]; }
    vor = __reduce_or_sync(0xFFFFFFFFu, vor); vnor = __reduce_or_sync(0xFFFFFFFFu, vnor);
    if ((threadIdx.x & 31u) == 0) { smem[2 * (threadIdx.x >> 5)] = vor; smem[2 * (threadIdx.x >> 5) + 1] = vnor; }
    __syncthreads();
    vor = 0; vnor = 0;
    for (int q = 0; q < kSegWarps; ++q) { vor |= smem[2 * q]; vnor |= smem[2 * q + 1]; }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->lab_or = vor; ctl->lab_nor = vnor; }
    pl = make_plan(vor, vnor, S.gbits, true);
    seg_hash_keys<IPT>(S, pl, smem);
    if (S.dbgts && threadIdx.x == 0) S.dbgts[blockIdx.x] = globaltimer();
    stamp(ctl, 3);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
  } else {
    seg_hash(S, smem);
    stamp(ctl, 1);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    stamp(ctl, 2);
    pl = make_plan(ld_relaxed(&ctl->lab_or), ld_relaxed(&ctl->lab_nor), S.gbits, S.use_label != 0);
    seg_vkey<IPT>(S, pl, smem);
    stamp(ctl, 3);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
  }
  stamp(ctl, 4);
  for (int p = 0; p < pl.npass; ++p) {
    seg_sort_pass<IPT>(S, pl, p, smem);
    if (S.dbgts && threadIdx.x == 0) S.dbgts[(size_t)(1 + p) * gridDim.x + blockIdx.x] = globaltimer();
    stamp(ctl, 5 + 2 * p);
    grid_sync(&ctl->bar_cnt, epoch, &ctl->err);
    stamp(ctl, 6 + 2 * p);
  }
  const u64* key = (pl.npass & 1) ? S.keyB : S.keyA;
  const u32* val = (pl.npass & 1) ? S.valB : S.valA;
  T.run(S, pl, key, val, smem, epoch);
  stamp(ctl, 19);
}

struct SegInputs {
  int64_t B; int K;
  const int64_t* keys; const float* labels; const uint8_t* row_ok;
  bool use_label;        // sort by (group, label, row) instead of (group, row)
  bool nan_label_is_trash;
  RowMap rm = RowMap{0, 0, 0, 0};
  bool allow_merged = false;   // the caller accepts table-slot group ids (not deterministic across calls / ranks)
  GatherArgs gather = GatherArgs{};   // peer-memory gather of the blocked input rows by k_init (world = 0: off)
  bool fast = false;                  // counting path (the tail supports it, the arena is persistent and clean): no k_init
};

inline SegParams make_seg_params(const Layout& L, void* scratch, const SegInputs& in) {
  char* base = static_cast<char*>(scratch);
  SegParams S{};
  S.B = (u32)in.B; S.K = in.K; S.gbits = L.gbits; S.use_label = in.use_label ? 1 : 0;
  S.nan_trash = in.nan_label_is_trash ? 1 : 0;
  S.keys = in.keys; S.row_ok = in.row_ok; S.labels = in.labels;
  S.rm = in.rm; S.kpitch = in.rm.Bl ? in.rm.Bl : (u32)in.B;
  S.capmask = L.cap - 1; S.ntiles = L.ntiles;
  S.table = at<u32>(base, L.table); S.table1 = at<u32>(base, L.table1);
  S.slot = at<u32>(base, L.slot); S.slot1 = at<u32>(base, L.slot1);
  S.keyA = at<u64>(base, L.keyA); S.keyB = at<u64>(base, L.keyB);
  S.valA = at<u32>(base, L.valA); S.valB = at<u32>(base, L.valB);
  S.hist = at<u32>(base, L.hist); S.tilehist = at<u32>(base, L.tilehist); S.th0 = at<u32>(base, L.th0);
  S.merged = 0; S.labpart = at<u32>(base, L.labpart); S.npart = 0;
  static const char* segdbg = getenv("RN_SEG_DEBUG");
  S.dbgts = (segdbg && *segdbg == '1') ? at<u64>(base, L.gstat) : nullptr;
  S.ctl = at<Ctl>(base, L.ctl);
  S.fast = 0;
  S.rec = at<GRec>(base, L.rec); S.rec2 = at<GRec2>(base, L.rec2); S.glist = at<u32>(base, L.glist); S.gcount = at<u32>(base, L.gcount);
  S.rslot = S.slot; S.rmeta = S.slot1;
  S.init_zero = at<uint4>(base, L.hist); S.init_zero16 = (u32)((L.zero_end - L.hist) / 16);
  S.init_ones = at<uint4>(base, L.ones_begin); S.init_ones16 = (u32)((L.ones_end - L.ones_begin) / 16);
  return S;
}

// group-id bits of the merged first phase: table slots < cap plus the id `cap` of the rows that cannot pair
inline int seg_merged_gbits(const Layout& L) { return bit_width_u64((uint64_t)L.cap); }

// Enqueue init + the segmentation kernel with the given tail (2 launches; 1 on the counting path).
template <class Tail>
cudaError_t seg_run(const Layout& L, void* scratch, const SegInputs& in, const Tail& tail, cudaStream_t st) {
  const bool fast = Tail::kFast && in.fast && in.K == 1 && in.use_label && in.labels;
  const bool merged = !fast && in.allow_merged && in.K == 1 && in.use_label && in.labels && !in.rm.Bl;
  int npart = 0;
  cudaError_t e = cudaSuccess;
  // (counting path: no initialisation kernel -- except in the global mode, whose first kernel is also the peer gather)
  if (!fast || in.gather.world)
    e = merged ? seg_init(L, scratch, st, in.labels, in.row_ok, &npart)
               : seg_init(L, scratch, st, nullptr, nullptr, nullptr, in.gather.world ? &in.gather : nullptr,
                          fast && in.gather.world != 0);
  if (e != cudaSuccess) return e;
  SegParams S = make_seg_params(L, scratch, in);
  if (merged) { S.merged = 1; S.npart = npart; S.gbits = seg_merged_gbits(L); }
  S.fast = fast ? 1 : 0;
  Tail T = tail;
  int grid = (int)((in.B + kSegThreads - 1) / kSegThreads);
  const int sms = device_sm_count();
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  if (fast) grid = sms;          // (CTAs without rows work out the pair kernel's partition beside the scatter phase)
  void* args[] = {&S, &T};
  const void* fn = (L.ipt == 2) ? (const void*)k_seg<2, Tail> : (const void*)k_seg<8, Tail>;
  return launch_coop(fn, grid, kSegThreads, args, st);
}

}  // namespace rn
