// Shared device helpers, scratch layout and launch plumbing.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/recnow_b200.h"

namespace rn {

// NVTX range around a C-ABI entry point, so that a profiler can filter on the call (`ncu --nvtx --nvtx-include
// "rn_pairwise_fwd_bwd/"`) and a timeline names it.  Opt-in (RN_NVTX=1, read once): the default call path stays free of
// third-party callbacks -- the NVTX 3 globals are shared with the framework's copy in the same process.  (An A/B of the
// host-buffer front end with the ranges on and off showed no difference beyond the host's own noise.)
inline bool nvtx_enabled() {
  static const bool on = []() { const char* v = getenv("RN_NVTX"); return v && *v && *v != '0'; }();
  return on;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_enabled()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define RN_NVTX_RANGE(name) ::rn::NvtxRange rn_nvtx_range_(name)

typedef unsigned int u32;
typedef unsigned long long u64;

constexpr int kRadixBits = 9;                 // digit width of one LSD pass (512 bins = one per thread)
constexpr int kBins = 1 << kRadixBits;
constexpr int kMaxPass = 7;                   // ceil((28 group bits + 32 label bits) / 9)
constexpr int kSegThreads = 512;              // threads per CTA of the segmentation kernel (one CTA per SM)
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kIB = 64;                       // rows per I-block of the pair kernel (two per lane)
constexpr u32 kEmpty = 0xFFFFFFFFu;
constexpr int kInitMaxCtas = 1024;            // upper bound of k_init's grid (label-range partials)

// Blocked row layout of the input columns (multi-GPU global mode: an all-gather of packed per-rank blocks leaves the
// rows as B / Bl blocks of Bl rows, `stride` bytes apart).  Element i of a column lives at block i / Bl, offset
// i % Bl; the strides below are in ELEMENTS of 8-, 4- and 1-byte columns.  Bl = 0: contiguous columns.
struct RowMap {
  u32 Bl, s8, s4, s1;
  __host__ __device__ __forceinline__ size_t i8(u32 i) const { return Bl ? (size_t)(i / Bl) * s8 + (i % Bl) : i; }
  __host__ __device__ __forceinline__ size_t i4(u32 i) const { return Bl ? (size_t)(i / Bl) * s4 + (i % Bl) : i; }
  __host__ __device__ __forceinline__ size_t i1(u32 i) const { return Bl ? (size_t)(i / Bl) * s1 + (i % Bl) : i; }
};

// Device-side control block.  Zeroed by k_init on the launch paths that have one; on the PERSISTENT path (see
// rn_pairwise_args.scratch_persistent) the arena is zeroed once by the caller and the last CTA of every call puts the
// block back into its clean state (all zero except the barrier generations and the report of the finished call).
struct Ctl {
  u32 lab_or, lab_nor;          // OR of label bits / OR of ~label bits over pairable rows -> varying bit range
  u32 det_w, det_y;             // deterministic mode: max |row weight| and max |label| (float bits), bound of a pair weight
  u32 pad_a[2];
  u32 k2_ticket;                // dynamic work-unit ticket of the pair kernel
  u32 fin_done;                 // CTAs that finished the final reduction (the last one writes the scalars / cleans up)
  u32 n_units, unit_c;          // work list: number of units, J-blocks per unit
  u32 n_groups;                 // distinct groups (listwise)
  u32 n_valid;                  // valid lists (listwise)
  u32 err;                      // device-side error flags (0 ok; 1 barrier timeout, 2 pair capacity, 4 hash table full)
  u32 fallback;                 // counting path: a label outside the level menu / a non-positive row weight was seen -> radix path
  u32 cursor;                   // counting path: next free sorted position (group base allocation)
  u32 path;                     // which segmentation ran (1 counting, 2 radix); report only
  u64 n_pair;                   // exact kept-pair count
  u64 n_tiles;                  // total 32x32 micro-tiles in the work list
  double loss_sum;              // sum_p wocc * lossrow[p]   (log2 units)
  double acc_d[4];              // generic double accumulators (GAUC: weighted AUC sum, weight sum, plain AUC sum)
  u64 ts[24];                   // phase timestamps (%globaltimer, ns) written by CTA 0: measurement aid
  u64 dbg[8];                   // pair-kernel debug tallies (RN_PAIR_DEBUG=1): see k_pair
  // report of the last finished call (copied here before the working fields are reset; read by the rn_debug_* calls)
  u32 rep_err, rep_path, rep_n_units, rep_unit_c; u64 rep_n_tiles;
  // arrival counters of the grid barriers of the segmentation / pair kernel, each in a line of its own (they are
  // polled); zero at the start of a call, counting up across the barriers of the kernel
  __align__(128) u32 bar_cnt; u32 pad_g[31];
  __align__(128) u32 bar2_cnt; u32 pad_h[31];
};

// Sort plan.  Compact sort key = (gid << labbits) | ((enc_label >> labshift) & mask): only the varying bit
// range of the order-preserving label encoding is kept (binary / graded labels need 7-9 bits), then the key's
// gbits + labbits bits are split evenly over the fewest 9-bit passes.  Recomputed (cheaply) from the control
// block by every kernel that needs it.
struct Plan { int npass, labshift, labbits; int shift[kMaxPass]; int nbits[kMaxPass]; };

__host__ __device__ inline Plan make_plan(u32 lab_or, u32 lab_nor, int gbits, bool use_label) {
  Plan p; p.npass = 0; p.labshift = 0; p.labbits = 0;
  const u32 vary = use_label ? (lab_or & lab_nor) : 0u;
  if (vary) {
    int lo = 0; while (!((vary >> lo) & 1u)) ++lo;
    int hi = 31; while (!((vary >> hi) & 1u)) --hi;
    p.labshift = lo; p.labbits = hi - lo + 1;
  }
  const int total = gbits + p.labbits;
  const int np = (total + kRadixBits - 1) / kRadixBits;
  int s = 0, left = total;
  for (int k = 0; k < np; ++k) {
    const int nb = (left + (np - k) - 1) / (np - k);
    p.shift[p.npass] = s; p.nbits[p.npass] = nb; ++p.npass; s += nb; left -= nb;
  }
  return p;
}

inline int bit_width_u64(uint64_t v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }

// ---- scratch arena layout (host side) ---------------------------------------------------------------
struct Layout {
  int64_t B; int K; int gbits; int ipt; u32 cap; u32 tile; u32 ntiles; u32 nib;
  size_t zero_begin, zero_end, ones_begin, ones_end, total;
  // zero-initialised region
  size_t ctl, hist, cprim, th0;
  // 0xFF-initialised region
  size_t table, table1;
  // plain
  size_t labpart, slot, slot1, keyA, keyB, valA, valB, tilehist, aj, ss, sy, swp, swn, gacc, lossrow, cnt, perm, blk, units, misc, gstat;
  // counting path (group_count.cuh): group records (clean = zero between calls), created-group lists per 512-row tile,
  // sorted group index column
  size_t rec, rec2, glist, gcount, sgrp, bnd, jn, lpart;
  int64_t Bcap;
};

// Group record of the counting segmentation path (64 bytes, one per hash slot + one for the rows that cannot pair).
// The first 16 bytes are claimed with ONE 128-bit compare-and-swap (key + creator row), so a probe never needs a second
// dependent load to compare keys.  All zero = free.
struct __align__(16) GRec {
  u64 key; u32 rep1; u32 base;  // key, creator row + 1 (0 = free), first sorted position of the group
  u32 cnt[8];                   // rows per label level; rewritten by the offsets phase to the level starts relative to base
  float wocc; u32 own;          // occurrence weight c_h ^ power of the group; global mode: 0 = this rank scores its own rows of
                                // the group, 1 = ALL rows (it owns the whole small group), 2 = none (another rank does)
  u64 npair;                    // kept pairs of the group (PW:286-289)
};
// Second record of a group, global mode only: rows per label level among THIS rank's rows (cl) and among the other
// ranks' rows (cr); cr is rewritten by the offsets phase to the level starts of the remote part.  All zero = clean.
struct __align__(16) GRec2 { u32 cl[8]; u32 cr[8]; };
static_assert(sizeof(GRec2) == sizeof(GRec), "the second records are addressed through the first table's base pointer");
constexpr int kLevels = 8;       // label levels of the counting path: integer-valued labels -1 .. 6
constexpr int kGTile = 512;      // rows per tile of the counting path (= kSegThreads)
constexpr u32 kOwnRows = 2048;   // global mode: groups below this many rows are scored whole by ONE rank (hash of the key)
constexpr u32 kBndCap = 16384;   // piece boundaries of the pair kernel's partition the arena holds (warps of its grid + 1)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Work-list granularity of the pair kernel: the J ranges of the I-blocks are cut into about this many units
// (developer knob RN_TARGET_UNITS).
inline u32 target_units() {
  static u32 n = 0;
  if (!n) { const char* v = getenv("RN_TARGET_UNITS"); const long x = (v && *v) ? atol(v) : 0; n = (x > 0 && x <= (1 << 22)) ? (u32)x : 16384u; }
  return n;
}

// Offsets depend on the CAPACITY Bcap the arena was sized for (rn_pairwise_args.scratch_rows; = B when 0), the loop
// bounds (tiles, I-blocks, id bits) on the rows B of the call: a persistent arena keeps one layout for every B <= Bcap.
inline Layout make_layout(int64_t Bcap, int K, int64_t B = 0) {
  if (B <= 0) B = Bcap;
  Layout L; L.B = B; L.K = K; L.Bcap = Bcap;
  L.gbits = B > 1 ? bit_width_u64((uint64_t)(B - 1)) : 1;
  u32 cap = 1024; while ((int64_t)cap < 2 * Bcap) cap <<= 1;
  L.cap = cap;
  L.ipt = (B <= 148 * 1024) ? 2 : 8;                 // sort tile = 1024 rows (<= 148 tiles) or 4096 rows
  L.tile = (u32)(kSegThreads * L.ipt);
  L.ntiles = (u32)((B + L.tile - 1) / L.tile);
  L.nib = (u32)((B + kIB - 1) / kIB);
  const size_t ipt_cap = (Bcap <= 148 * 1024) ? 2 : 8;
  const size_t ntiles_cap = (size_t)((Bcap + 1023) / 1024);       // upper bound for either tile size
  (void)ipt_cap;
  const size_t nib_cap = (size_t)((Bcap + kIB - 1) / kIB);
  const size_t Bc = (size_t)Bcap;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  L.zero_begin = o;
  L.ctl = take(sizeof(Ctl));
  L.hist = take(sizeof(u32) * kMaxPass * kBins);
  L.cprim = take(sizeof(u64) * cap);                   // pair totals per group id (first-occurrence row < B, or table slot < cap)
  L.th0 = take(sizeof(u32) * ntiles_cap * kBins);      // tile histograms, buffer 0 (accumulated with atomics in the merged first phase)
  L.zero_end = o;
  L.ones_begin = o;
  L.table = take(sizeof(u32) * cap);
  L.table1 = K > 1 ? take(sizeof(u32) * cap) : L.table;
  L.ones_end = o;
  L.labpart = take(sizeof(u32) * 2 * kInitMaxCtas);
  L.slot = take(sizeof(u32) * Bc);
  L.slot1 = take(sizeof(u32) * Bc);
  L.keyA = take(sizeof(u64) * Bc);
  L.keyB = take(sizeof(u64) * Bc);
  L.valA = take(sizeof(u32) * Bc);
  L.valB = take(sizeof(u32) * Bc);
  L.tilehist = take(sizeof(u32) * 2 * ntiles_cap * kBins);    // buffers 1 and 2 (buffer 0 is L.th0)
  L.aj = take(sizeof(uint2) * Bc);
  L.ss = take(sizeof(float) * Bc);
  L.sy = take(sizeof(float) * Bc);
  L.swp = take(sizeof(float) * Bc);
  L.swn = take(sizeof(float) * Bc);
  L.gacc = take(sizeof(float) * Bc);
  L.lossrow = take(sizeof(float) * Bc);
  L.cnt = take(sizeof(u32) * Bc);
  L.perm = take(sizeof(u32) * Bc);
  L.blk = take(sizeof(uint2) * 2 * nib_cap);           // two J ranges per I-block
  L.units = take(sizeof(uint2) * (2 * nib_cap + target_units() + 1));
  L.misc = take(sizeof(u64) * (Bc + 1));
  L.gstat = take(sizeof(float) * 8 * Bc);   // listwise per-list records
  const size_t ngt = (Bc + kGTile - 1) / kGTile;
  L.rec = take(sizeof(GRec) * ((size_t)cap + 1));
  L.rec2 = take(sizeof(GRec2) * ((size_t)cap + 1));
  L.glist = take(sizeof(u32) * ngt * kGTile);
  L.gcount = take(sizeof(u32) * ngt);
  L.sgrp = take(sizeof(u32) * Bc);
  L.bnd = take(sizeof(uint4) * kBndCap);               // piece boundaries of the pair kernel's partition (PrePart)
  L.jn = take(sizeof(u32) * 2 * nib_cap);
  L.lpart = take(sizeof(double) * 1024);                // per-CTA loss partials (deterministic mode)
  L.total = o;
  return L;
}

template <typename T> inline T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }

// ---- device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ u32 ld_relaxed(const u32* p) {
  u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ u32 ld_acquire(const u32* p) {
  u32 v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// Grid-wide barrier of a cooperatively launched kernel (all CTAs co-resident).  The counter starts a call at zero
// (k_init, or the clean state the previous call left) and counts up across the barriers of the kernel: `epoch` is the
// per-thread running arrival target.  The waiters see the last arrival as soon as it lands in L2 -- no second store to
// wait for.  The fences publish this CTA's writes and invalidate its L1 so that plain loads after the barrier see other
// CTAs' data.  A bounded spin turns a scheduling failure (or an arena that was not clean) into ctl->err instead of a
// hung GPU.
__device__ __forceinline__ void st_relaxed(u32* p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void grid_sync(u32* cnt, u32& epoch, u32* err) {
  epoch += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(cnt, 1u);
    u32 spins = 0;
    while (ld_relaxed(cnt) < epoch) {
      if (++spins > (1u << 23)) { atomicOr(err, 1u); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

// Wait for the previous kernel of the stream (programmatic dependent launch, see launch_coop); a no-op when the kernel
// was launched without the attribute.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Programmatic dependent launch, primary side: the next kernel of the stream / graph may be scheduled from now on (its
// CTAs start on SMs as they become free and block in grid_dep_wait until this grid has completed and flushed).
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ u64 globaltimer() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// phase timestamp i (CTA 0, thread 0 only)
__device__ __forceinline__ void stamp(Ctl* ctl, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ts[i] = globaltimer();
}

// order-preserving float -> u32 (after -0.0 -> +0.0): a < b  <=>  enc(a) < enc(b) for non-NaN a, b
__device__ __forceinline__ u32 enc_label(float y) {
  u32 u = __float_as_uint(y + 0.0f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__device__ __forceinline__ u64 mix64(u64 x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31; return x;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ u32 warp_max(u32 v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ u32 warp_min(u32 v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// The CTA that finishes last files the report of the call and resets the working fields of the control block.
__device__ __forceinline__ void ctl_finish(Ctl* ctl) {
  ctl->rep_err = ctl->err; ctl->rep_path = ctl->path; ctl->rep_n_units = ctl->n_units; ctl->rep_unit_c = ctl->unit_c;
  ctl->rep_n_tiles = ctl->n_tiles;
  ctl->bar_cnt = 0; ctl->bar2_cnt = 0; ctl->det_w = 0; ctl->det_y = 0;
  ctl->lab_or = 0; ctl->lab_nor = 0; ctl->k2_ticket = 0; ctl->fin_done = 0; ctl->n_units = 0; ctl->unit_c = 0;
  ctl->n_groups = 0; ctl->n_valid = 0; ctl->err = 0; ctl->fallback = 0; ctl->cursor = 0; ctl->path = 0;
  ctl->n_pair = 0; ctl->n_tiles = 0; ctl->loss_sum = 0.0;
  ctl->acc_d[0] = 0.0; ctl->acc_d[1] = 0.0; ctl->acc_d[2] = 0.0; ctl->acc_d[3] = 0.0;
}

// J-block range of virtual block v (2b = range R1 of I-block b, 2b + 1 = range R2, see HeadsTail): first J-block and
// J-block count; R2 is trimmed so that no (I-block, J-block) tile is listed twice.
// (a range is stored as (~lo, hi): the counting path builds it with atomicMax from an all-zero = empty record)
__device__ __forceinline__ u32 vblock_tiles(const uint2* blk, u32 v, u32& jfirst) {
  const uint2 r = blk[v];
  u32 lo = 0, hi = 0;
  if (r.y > ~r.x) { lo = ~r.x >> 5; hi = (r.y + 31) >> 5; }
  if (v & 1) { const uint2 r1 = blk[v - 1]; if (r1.y > ~r1.x) lo = max(lo, (r1.y + 31) >> 5); }
  jfirst = lo;
  return hi > lo ? hi - lo : 0u;
}

// Zero every group record the count phase created (listed per 512-row tile): all CTAs of the grid, behind a grid barrier
// that nobody reads the records after.  A warp per tile, four lanes per 64-byte record.
__device__ __forceinline__ void clean_records_grid(GRec* rec, u32 rec2_off, const u32* glist, const u32* gcount, u32 ngt) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  const u32 ln = threadIdx.x & 31u, nw = blockDim.x >> 5;
  for (u32 t = blockIdx.x * nw + (threadIdx.x >> 5); t < ngt; t += gridDim.x * nw) {
    const u32 n = gcount[t];
    for (u32 k = ln; k < 4 * n; k += 32) {
      const u32 slot = glist[(size_t)t * kGTile + (k >> 2)];
      reinterpret_cast<uint4*>(rec + slot)[k & 3u] = z;
      if (rec2_off) reinterpret_cast<uint4*>(rec + rec2_off + slot)[k & 3u] = z;
    }
  }
}

// In-place exclusive prefix sum of a[0, n) in shared memory by the whole CTA (blockDim.x a multiple of 32, <= 1024).
// Every WARP owns a contiguous segment and walks it 32 elements at a time (coalesced, conflict-free accesses, a shuffle
// scan per step, the running sum in a register) -- no barrier inside the walk; then one scan over the warps' totals and
// a second walk that adds the warp's offset.  a[n] receives the total (sc: 34 words).  All threads must call it.
__device__ __forceinline__ void block_excl_scan(u32* a, u32 n, u32* sc) {
  const u32 ln = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (n <= 4u * blockDim.x && (n & 3u) == 0u) {
    // up to four elements per thread: one 16-byte load, the four sums in registers, ONE shuffle scan per warp and one over
    // the warps' totals (the walk below pays a dependent shuffle scan per 32 elements of a warp's segment)
    const u32 i4 = 4u * threadIdx.x;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i4 < n) v = *reinterpret_cast<const uint4*>(a + i4);
    const u32 tsum = v.x + v.y + v.z + v.w;
    u32 inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
    if (ln == 31) sc[w] = inc;
    __syncthreads();
    if (w == 0) {
      const u32 x = ln < nw ? sc[ln] : 0u;
      u32 xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (ln >= (u32)o) xi += y; }
      sc[ln] = xi - x;
      if (ln == 31) sc[32] = xi;
    }
    __syncthreads();
    const u32 base = sc[w] + inc - tsum;
    if (i4 < n) *reinterpret_cast<uint4*>(a + i4) = make_uint4(base, base + v.x, base + v.x + v.y, base + v.x + v.y + v.z);
    if (threadIdx.x == 0) a[n] = sc[32];
    __syncthreads();
    return;
  }
  const u32 seg = ((n + nw - 1) / nw + 31u) & ~31u;
  const u32 s0 = min(w * seg, n), s1 = min(s0 + seg, n);
  u32 carry = 0;
  for (u32 base = s0; base < s1; base += 32) {
    const u32 i = base + ln;
    const u32 v = i < s1 ? a[i] : 0u;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (ln >= (u32)o) inc += x; }
    if (i < s1) a[i] = carry + inc - v;
    carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
  }
  if (ln == 0) sc[w] = carry;
  __syncthreads();
  if (w == 0) {
    const u32 x = ln < nw ? sc[ln] : 0u;
    u32 xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (ln >= (u32)o) xi += y; }
    sc[ln] = xi - x;
    if (ln == 31) sc[32] = xi;
  }
  __syncthreads();
  const u32 off = sc[w];
  if (off) for (u32 i = s0 + ln; i < s1; i += 32) a[i] += off;
  if (threadIdx.x == 0) a[n] = sc[32];
  __syncthreads();
}

// Largest idx in [lo, hi) with arr[idx] <= key (arr nondecreasing, arr[lo] <= key); one thread, binary search.
__device__ __forceinline__ u32 last_le(const u32* arr, u32 lo, u32 hi, u32 key) {
  while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (arr[mid] <= key) lo = mid; else hi = mid; }
  return lo;
}

// GAUC kernel (gauc.cu) on the arrays the pairwise segmentation leaves; the host entry lives with the segmentation's
// instantiation in pairwise.cu.
struct GaucArgs {
  u32 B, nib;
  const uint2* aj; const float* ss; const uint2* blk; uint2* blk_w;
  u64* acc2; u64* npg; u32* gsz;        // per group (indexed by its first sorted position): 2 x concordant + ties, pairs, rows
  float* gauc; float* auc_mean; int32_t* n_valid; int64_t* n_pair; int64_t* conc2;
  int fast; GRec* rec; u32 rec2_off; const u32* glist; const u32* gcount; u32 ngt;
  Ctl* ctl;
};
cudaError_t launch_gauc(const GaucArgs& A, cudaStream_t st);

// Peer-memory gather done by k_init (see rn_pairwise_args.peer_blocks): `world` blocks of n16 16-byte words.
struct GatherArgs { const uint4* src[8]; uint4* dst; u32 n16; u32 world; };

// ---- host plumbing (segment.cu) ---------------------------------------------------------------------
// Zero / 0xFF-fill the initialised regions of the arena (one ordinary launch; it also resets the barriers).
// Optionally (labels != nullptr) the same launch OR-reduces the label bits of the pairable rows into per-CTA partials
// (Layout::labpart), which lets k_seg build the sort keys in its first phase; returns the grid size in *ncta.
cudaError_t seg_init(const Layout& L, void* scratch, cudaStream_t st, const float* labels = nullptr,
                     const uint8_t* row_ok = nullptr, int* ncta = nullptr, const GatherArgs* gather = nullptr,
                     bool gather_only = false);
int device_sm_count();
// Cooperative launch of `kernel` with `grid` CTAs of `threads` threads (grid must not exceed the co-resident limit).
cudaError_t launch_coop(const void* kernel, int grid, int threads, void** args, cudaStream_t st, size_t smem = 0);

// One pairwise call as ONE launch of a cached CUDA graph (see segment.cu).  Construct it before the launches of the
// call, enqueue them on run_stream, then finish().  Falls back to plain launches on the caller's stream when graphs
// are disabled (RN_GRAPH=0), the stream is being captured by the caller, or a graph API call has failed.
struct GraphSlot;
const void* seg_init_func();
long long graph_launch_count();
struct GraphCall {
  cudaStream_t st, run_stream; GraphSlot* slot = nullptr; int mode = 0;     // 0 direct, 1 node update, 2 capture
  // timed = true: the graph also holds two event-record nodes around the pair kernel (see timed_events)
  GraphCall(const void* f_init, const void* f_seg, const void* f_pair, cudaStream_t user_stream, bool allow,
            bool timed = false);
  cudaError_t finish(bool ok);
  bool capturing() const { return mode == 2; }
  bool updating() const { return mode == 1; }
};
// Calls enqueued from now on by this thread use the cached graphs of `lane` (0 = the default); see GraphSlot.
void set_graph_lane(int lane);
// The two events recorded by the timed graphs of this thread (created on first use; nullptr on failure).
cudaEvent_t* timed_events();

// Global mode with a score- / weight-dependent pair set: where stage 1 publishes this rank's partial per-row pair counts
// (by original row) and where stage 2 finds every rank's (peer-mapped), see pairwise_call.
struct DynSplit { u32* xcnt_out; const u32* xcnt_peer[8]; int world; };
int pairwise_call(const rn_pairwise_args* a, void* scratch, size_t scratch_bytes, void* stream, const DynSplit* split, int stage);

// Small batches (small.cu): the whole call as one launch of one CTA; true if it took the call (*rc = its status).
bool small_pairwise(const rn_pairwise_args* a, void* scratch, cudaStream_t st, int mode, bool hinge, int* rc);

inline int check_align(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) ? RN_ERR_ALIGN : RN_OK; }

}  // namespace rn
