// K1 — group segmentation: hash grouping (composite int64 keys -> first-occurrence row id) followed by a
// stable LSD radix sort of the rows by (first-occurrence id, order-preserving label bits, row).
//
// Replaces the reference's (B,B) group-equality matrix (pairwise_loss_from_batch.py:33-37, 68-73) and
// tf.unique_with_counts (listwise_loss_from_batch.py:109, pairwise_loss_from_batch.py:146).
//
// Why hash first: the ids are sparse 62-bit values (8 radix passes); the first-occurrence row id is a dense
// ceil(log2 B)-bit value (2 passes at B = 65536), is deterministic, handles composite keys for free and is
// exactly the group order the listwise path must report (first occurrence, LW:109).  Rows that can form no
// pair (row_ok = 0, NaN label) become singleton groups (gid = own row), so no compaction pass is needed.
//
// Sort pass = one kernel ("onesweep" style): per-warp stable ranking with match.any, per-tile histogram,
// decoupled look-back over the tile status words for the cross-tile prefix, direct scatter.  Digits are
// chosen on the device from the varying bits of the label encoding (binary / graded labels need ONE label
// pass); the host launches the worst-case number of passes and void passes exit immediately.
#include "common.cuh"

namespace rn {

__global__ void __launch_bounds__(256) k_init(uint4* zero, size_t nzero16, uint4* ones, size_t nones16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(~0u, ~0u, ~0u, ~0u);
  for (size_t k = i; k < nzero16; k += stride) zero[k] = z;
  for (size_t k = i; k < nones16; k += stride) ones[k] = f;
}

__device__ __forceinline__ u32 ld_relaxed(const u32* p) {
  u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed(u32* p, u32 v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Open-addressing insert of row i's key tuple (nk columns) into `table`; slots hold a representative row.
__device__ __forceinline__ u32 hash_insert(const int64_t* __restrict__ keys, int64_t B, int nk, u32 i,
                                           u32* table, u32 capmask) {
  u64 h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < nk; ++k) h = mix64(h ^ (u64)keys[(size_t)k * B + i]);
  u32 s = (u32)h & capmask;
  for (;;) {
    u32 cur = ld_relaxed(table + s);
    if (cur == kEmpty) {
      u32 prev = atomicCAS(table + s, kEmpty, i);
      cur = (prev == kEmpty) ? i : prev;
    }
    bool eq = true;
    if (cur != i)
      for (int k = 0; k < nk; ++k) eq = eq && (keys[(size_t)k * B + cur] == keys[(size_t)k * B + i]);
    if (eq) return s;
    s = (s + 1) & capmask;
  }
}

__global__ void __launch_bounds__(256) k_hash(int64_t B, int K, const int64_t* __restrict__ keys,
                                              const uint8_t* __restrict__ row_ok, const float* __restrict__ labels,
                                              int nan_trash, u32* table, u32* first, u32* slot, u32* table1,
                                              u32* slot1, u32 capmask, Ctl* ctl) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool in = i < B;
  u32 e = 0;
  bool ok = false;
  if (in) {
    float y = labels ? labels[i] : 0.f;
    e = enc_label(y);
    ok = (row_ok ? row_ok[i] != 0 : true) && !(nan_trash && (y != y));
  }
  u32 vor = __reduce_or_sync(0xFFFFFFFFu, in ? e : 0u);
  u32 vnor = __reduce_or_sync(0xFFFFFFFFu, in ? ~e : 0u);
  if (lane_id() == 0) {
    if (vor & ~ld_relaxed(&ctl->lab_or)) atomicOr(&ctl->lab_or, vor);
    if (vnor & ~ld_relaxed(&ctl->lab_nor)) atomicOr(&ctl->lab_nor, vnor);
  }
  if (!in) return;
  if (!ok) { slot[i] = kEmpty; if (K > 1) slot1[i] = kEmpty; return; }
  u32 s = hash_insert(keys, B, K, (u32)i, table, capmask);
  slot[i] = s;
  if (ld_relaxed(first + s) > (u32)i) atomicMin(first + s, (u32)i);
  if (K > 1) slot1[i] = hash_insert(keys, B, 1, (u32)i, table1, capmask);
}

// Sort key of every row + the digit histograms of all passes.
__global__ void __launch_bounds__(256) k_vkey_hist(int64_t B, const float* __restrict__ labels,
                                                   const u32* __restrict__ slot, const u32* __restrict__ first,
                                                   u64* __restrict__ keyA, u32* __restrict__ valA, u32* hist,
                                                   const Ctl* ctl, int gbits, int use_label) {
  __shared__ u32 sh[kMaxPass * kBins];
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, gbits, use_label != 0);
  for (int k = threadIdx.x; k < pl.npass * kBins; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += stride) {
    u32 s = slot[i];
    u32 gid = (s == kEmpty) ? (u32)i : first[s];
    u32 lo = (use_label && labels) ? enc_label(labels[i]) : 0u;
    u64 key = ((u64)gid << 32) | lo;
    keyA[i] = key; valA[i] = (u32)i;
#pragma unroll 1
    for (int p = 0; p < pl.npass; ++p)
      atomicAdd(&sh[p * kBins + (u32)((key >> pl.shift[p]) & ((1u << pl.nbits[p]) - 1u))], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < pl.npass * kBins; k += blockDim.x)
    if (sh[k]) atomicAdd(hist + k, sh[k]);
}

constexpr u32 kFlagAgg = 1u << 30, kFlagPre = 2u << 30, kValMask = (1u << 30) - 1u;

__global__ void __launch_bounds__(kSortThreads) k_sort_pass(int pass, int64_t B, int gbits, int use_label,
                                                            u64* keyA, u32* valA, u64* keyB, u32* valB,
                                                            const u32* __restrict__ hist, u32* status, u32 ntiles,
                                                            Ctl* ctl) {
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, gbits, use_label != 0);
  if (pass >= pl.npass) return;
  const u64* ksrc = (pass & 1) ? keyB : keyA; const u32* vsrc = (pass & 1) ? valB : valA;
  u64* kdst = (pass & 1) ? keyA : keyB;       u32* vdst = (pass & 1) ? valA : valB;
  const int shift = pl.shift[pass];
  const u32 nb = 1u << pl.nbits[pass], dmask = nb - 1u;
  constexpr int W = kSortThreads / 32;
  __shared__ u32 whist[W][kBins];
  __shared__ u32 gbase[kBins];
  __shared__ u32 wsum[W];
  __shared__ u32 s_tile;
  const u32 tid = threadIdx.x, w = tid >> 5, ln = tid & 31u;
  if (tid == 0) s_tile = atomicAdd(&ctl->tickets[pass], 1u);
  for (u32 k = tid; k < W * kBins; k += kSortThreads) (&whist[0][0])[k] = 0;
  __syncthreads();
  const u32 tile = s_tile;
  const int64_t base = (int64_t)tile * kSortTile + (int64_t)w * (32 * kSortIpt);
  u64 key[kSortIpt]; u32 val[kSortIpt]; u32 rank[kSortIpt];
#pragma unroll
  for (int r = 0; r < kSortIpt; ++r) {
    int64_t idx = base + r * 32 + ln;
    bool v = idx < B;
    key[r] = v ? ksrc[idx] : ~0ull; val[r] = v ? vsrc[idx] : 0u;
    u32 d = v ? (u32)((key[r] >> shift) & dmask) : 0xFFFFu;
    u32 m = __match_any_sync(0xFFFFFFFFu, d);
    u32 leader = __ffs(m) - 1;
    u32 old = 0;
    if (v && ln == leader) { old = whist[w][d]; whist[w][d] = old + __popc(m); }
    old = __shfl_sync(0xFFFFFFFFu, old, leader);
    rank[r] = old + __popc(m & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan of the global digit histogram (bin bases) -- kBins == kSortThreads entries
  u32 hv = (tid < nb) ? hist[pass * kBins + tid] : 0u;
  u32 inc = hv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += t; }
  if (ln == 31) wsum[w] = inc;
  __syncthreads();
  u32 woff = 0;
  for (u32 k = 0; k < w; ++k) woff += wsum[k];
  const u32 binbase = woff + inc - hv;
  // per-bin: prefix over warps, publish, look back
  if (tid < nb) {
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) { u32 t = whist[k][tid]; whist[k][tid] = sum; sum += t; }
    u32* st = status + ((size_t)pass * ntiles + tile) * kBins + tid;
    u32 excl = 0;
    if (tile == 0) {
      st_relaxed(st, kFlagPre | sum);
    } else {
      st_relaxed(st, kFlagAgg | sum);
      int64_t t = (int64_t)tile - 1;
      u32 spins = 0;
      while (t >= 0) {
        u32 sv = ld_relaxed(status + ((size_t)pass * ntiles + t) * kBins + tid);
        if (sv == 0) {
          if (++spins > (1u << 24)) { atomicOr(&ctl->err, 1u); break; }
          __nanosleep(20);
          continue;
        }
        excl += sv & kValMask;
        if (sv & kFlagPre) break;
        --t;
      }
      st_relaxed(st, kFlagPre | (excl + sum));
    }
    gbase[tid] = binbase + excl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortIpt; ++r) {
    int64_t idx = base + r * 32 + ln;
    if (idx < B) {
      u32 d = (u32)((key[r] >> shift) & dmask);
      u32 pos = gbase[d] + whist[w][d] + rank[r];
      kdst[pos] = key[r]; vdst[pos] = val[r];
    }
  }
}

// ---- group bounds (used by the listwise path and by pair materialisation) ----------------------------
__global__ void __launch_bounds__(1024) k_bounds(u32 B, int gbits, int use_label, const u64* keyA, const u64* keyB,
                                                 const u32* valA, const u32* valB, u32* __restrict__ astart,
                                                 u32* __restrict__ gend, u32* __restrict__ perm, GatherCols gc,
                                                 Ctl* ctl) {
  const Plan pl = make_plan(ctl->lab_or, ctl->lab_nor, gbits, use_label != 0);
  const u64* __restrict__ key = (pl.npass & 1) ? keyB : keyA;
  const u32* __restrict__ val = (pl.npass & 1) ? valB : valA;
  __shared__ u32 wmax[32];
  __shared__ u32 carry;
  const u32 t0 = blockIdx.x * 1024u, p = t0 + threadIdx.x, ln = lane_id(), w = threadIdx.x >> 5;
  const bool in = p < B;
  const u32 gid = in ? (u32)(key[p] >> 32) : 0u;
  const u32 gprev = (in && p > 0) ? (u32)(key[p - 1] >> 32) : ~gid;
  const u32 gnext = (in && p + 1 < B) ? (u32)(key[p + 1] >> 32) : ~gid;
  if (w == 0) {
    // warp-cooperative 32-ary lower bound of the group id at the tile start over [0, t0]
    const u32 g0 = (u32)(key[t0] >> 32);
    u32 lo = 0, hi = t0;                 // answer in [lo, hi]; key[hi] has gid == g0
    while (hi - lo > 0) {
      const u32 span = hi - lo, step = (span + 30) / 31;     // lane 31 always probes hi
      const u32 q = lo + min(ln * step, span);
      const bool ge = (u32)(key[q] >> 32) >= g0;
      const u32 bal = __ballot_sync(0xFFFFFFFFu, ge);
      const u32 f = __ffs(bal) - 1;      // first probe that is >= g0 (exists: lane hitting hi or beyond)
      const u32 nhi = lo + min(f * step, span);
      const u32 nlo = f ? lo + min((f - 1) * step, span) + 1 : lo;
      hi = nhi; lo = min(nlo, nhi);
    }
    if (ln == 0) carry = lo;
  }
  u32 x = (in && gid != gprev) ? p + 1 : 0u;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, x, o); if (ln >= (u32)o) x = max(x, t); }
  if (ln == 31) wmax[w] = x;
  __syncthreads();
  u32 c = 0;
  for (u32 k = 0; k < w; ++k) c = max(c, wmax[k]);
  x = max(x, c);
  if (in) {
    const u32 a = x ? x - 1 : carry;
    astart[p] = a;
    if (gid != gnext) gend[a] = p + 1;
    const u32 row = val[p];
    perm[p] = row;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (gc.src[k]) gc.dst[k][p] = gc.src[k][row];
  }
}

cudaError_t seg_bounds(const Layout& L, void* scratch, int use_label, u32* astart, u32* gend, u32* perm,
                       const GatherCols& gc, cudaStream_t st) {
  char* base = static_cast<char*>(scratch);
  k_bounds<<<(unsigned)((L.B + 1023) / 1024), 1024, 0, st>>>((u32)L.B, L.gbits, use_label, at<u64>(base, L.keyA),
                                                            at<u64>(base, L.keyB), at<u32>(base, L.valA),
                                                            at<u32>(base, L.valB), astart, gend, perm, gc,
                                                            at<Ctl>(base, L.ctl));
  return cudaGetLastError();
}

int seg_launch_count(const Layout& L) { return 3 + max_label_passes() + group_passes(L.gbits); }

cudaError_t seg_run(const Layout& L, void* scratch, const SegInputs& in, cudaStream_t st) {
  char* base = static_cast<char*>(scratch);
  Ctl* ctl = at<Ctl>(base, L.ctl);
  const int64_t B = in.B;
  {
    size_t nz = (L.zero_end - L.zero_begin) / 16, no = (L.ones_end - L.ones_begin) / 16;
    int grid = (int)((nz + no + 255) / 256); if (grid > 148 * 8) grid = 148 * 8; if (grid < 1) grid = 1;
    k_init<<<grid, 256, 0, st>>>(at<uint4>(base, L.zero_begin), nz, at<uint4>(base, L.ones_begin), no);
  }
  int grid = (int)((B + 255) / 256);
  k_hash<<<grid, 256, 0, st>>>(B, in.K, in.keys, in.row_ok, in.labels, in.nan_label_is_trash ? 1 : 0,
                               at<u32>(base, L.table), at<u32>(base, L.first), at<u32>(base, L.slot),
                               at<u32>(base, L.table1), at<u32>(base, L.slot1), L.cap - 1, ctl);
  int g2 = (int)((B + 1023) / 1024); if (g2 > 148 * 4) g2 = 148 * 4;
  k_vkey_hist<<<g2, 256, 0, st>>>(B, in.labels, at<u32>(base, L.slot), at<u32>(base, L.first),
                                  at<u64>(base, L.keyA), at<u32>(base, L.valA), at<u32>(base, L.hist), ctl,
                                  L.gbits, in.use_label ? 1 : 0);
  int npass_max = (in.use_label ? max_label_passes() : 0) + group_passes(L.gbits);
  for (int p = 0; p < npass_max; ++p)
    k_sort_pass<<<L.ntiles, kSortThreads, 0, st>>>(p, B, L.gbits, in.use_label ? 1 : 0, at<u64>(base, L.keyA),
                                                   at<u32>(base, L.valA), at<u64>(base, L.keyB),
                                                   at<u32>(base, L.valB), at<u32>(base, L.hist),
                                                   at<u32>(base, L.status), L.ntiles, ctl);
  return cudaGetLastError();
}

}  // namespace rn
