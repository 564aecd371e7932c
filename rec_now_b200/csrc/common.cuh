// Shared device helpers, scratch layout and internal launcher declarations.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/recnow_b200.h"

namespace rn {

typedef unsigned int u32;
typedef unsigned long long u64;

constexpr int kRadixBits = 8;                 // digit width of one LSD pass
constexpr int kBins = 1 << kRadixBits;
constexpr int kMaxPass = 8;                   // <= 4 label passes + <= 4 group passes
constexpr int kSortThreads = 256;             // 8 warps
constexpr int kSortIpt = 8;                   // items per thread -> 2048-row tiles
constexpr int kSortTile = kSortThreads * kSortIpt;
constexpr u32 kEmpty = 0xFFFFFFFFu;

// Device-side control block (zero-initialised by k_init at the start of every call).
struct Ctl {
  u32 lab_or, lab_nor;          // OR of label bits / OR of ~label bits -> varying bit range
  u32 tickets[kMaxPass];        // tile tickets of the sort passes (look-back needs in-order start)
  u32 heads_done;               // last-block detection in k_heads
  u32 k2_ticket;                // dynamic work-unit ticket of the pair kernel
  u32 fin_done;                 // last-block detection in the final kernels
  u32 n_units, unit_c;          // work list: number of units, J-blocks per unit
  u32 n_groups;                 // distinct groups (listwise)
  u32 n_valid;                  // valid lists (listwise)
  u32 err;                      // device-side error flag (0 ok)
  u64 n_pair;                   // exact kept-pair count
  u64 n_tiles;                  // total 32x32 micro-tiles in the work list
  double loss_sum;              // sum_p wocc * lossrow[p]   (log2 units)
};

struct Plan { int npass; int shift[kMaxPass]; int nbits[kMaxPass]; };

// Sort plan: label digits over the varying bit range of the order-preserving label encoding, then group
// digits over gbits bits starting at bit 32.  Recomputed (cheaply) by every kernel that needs it.
__host__ __device__ inline Plan make_plan(u32 lab_or, u32 lab_nor, int gbits, bool use_label) {
  Plan p; p.npass = 0;
  u32 vary = use_label ? (lab_or & lab_nor) : 0u;
  if (vary) {
    int lo = 0; while (!((vary >> lo) & 1u)) ++lo;
    int hi = 31; while (!((vary >> hi) & 1u)) --hi;
    for (int s = lo; s <= hi; s += kRadixBits) {
      int nb = hi - s + 1; if (nb > kRadixBits) nb = kRadixBits;
      p.shift[p.npass] = s; p.nbits[p.npass] = nb; ++p.npass;
    }
  }
  // spread the group bits evenly over the fewest passes
  int ngp = (gbits + kRadixBits - 1) / kRadixBits;
  int s = 32, left = gbits;
  for (int k = 0; k < ngp; ++k) {
    int nb = (left + (ngp - k) - 1) / (ngp - k);
    p.shift[p.npass] = s; p.nbits[p.npass] = nb; ++p.npass; s += nb; left -= nb;
  }
  return p;
}

inline int bit_width_u64(uint64_t v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }
inline int max_label_passes() { return (32 + kRadixBits - 1) / kRadixBits; }
inline int group_passes(int gbits) { return (gbits + kRadixBits - 1) / kRadixBits; }

// ---- scratch arena layout (host side) ---------------------------------------------------------------
struct Layout {
  int64_t B; int K; int gbits; u32 cap; u32 ntiles; u32 nblk;
  size_t zero_begin, zero_end, ones_begin, ones_end, total;
  // zero-initialised region
  size_t ctl, hist, status, cprim, gacc, lossrow, cnt, gstat;
  // 0xFF-initialised region
  size_t table, first, table1;
  // plain
  size_t slot, slot1, keyA, keyB, valA, valB, aj, ss, sy, swp, swn, blk, ustart, misc;
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

inline Layout make_layout(int64_t B, int K) {
  Layout L; L.B = B; L.K = K;
  L.gbits = B > 1 ? bit_width_u64((uint64_t)(B - 1)) : 1;
  u32 cap = 1024; while ((int64_t)cap < 2 * B) cap <<= 1;
  L.cap = cap;
  L.ntiles = (u32)((B + kSortTile - 1) / kSortTile);
  L.nblk = (u32)((B + 31) / 32);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  L.zero_begin = o;
  L.ctl = take(sizeof(Ctl));
  L.hist = take(sizeof(u32) * kMaxPass * kBins);
  L.status = take(sizeof(u32) * (size_t)kMaxPass * L.ntiles * kBins);
  L.cprim = take(sizeof(u64) * cap);
  L.gacc = take(sizeof(float) * B);
  L.lossrow = take(sizeof(float) * B);
  L.cnt = take(sizeof(u32) * B);
  L.gstat = take(sizeof(float) * 8 * B);   // listwise per-group statistics / misc per-row accumulators
  L.zero_end = o;
  L.ones_begin = o;
  L.table = take(sizeof(u32) * cap);
  L.first = take(sizeof(u32) * cap);
  L.table1 = take(sizeof(u32) * cap);
  L.ones_end = o;
  L.slot = take(sizeof(u32) * B);
  L.slot1 = take(sizeof(u32) * B);
  L.keyA = take(sizeof(u64) * B);
  L.keyB = take(sizeof(u64) * B);
  L.valA = take(sizeof(u32) * B);
  L.valB = take(sizeof(u32) * B);
  L.aj = take(sizeof(uint2) * B);
  L.ss = take(sizeof(float) * B);
  L.sy = take(sizeof(float) * B);
  L.swp = take(sizeof(float) * B);
  L.swn = take(sizeof(float) * B);
  L.blk = take(sizeof(uint2) * L.nblk);
  L.ustart = take(sizeof(u32) * (L.nblk + 1));
  L.misc = take(sizeof(u64) * (B + 1));
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

// ---- internal launchers (segment.cu) ----------------------------------------------------------------
struct SegInputs {
  int64_t B; int K;
  const int64_t* keys; const float* labels; const uint8_t* row_ok;
  bool use_label;        // sort by (group, label, row) instead of (group, row)
  bool nan_label_is_trash;
};
// Enqueues init + hash grouping + radix sort.  Afterwards the sorted (key, row) arrays are in
// (keyA,valA) if plan.npass is even else (keyB,valB) -- consumers recompute the plan from ctl.
cudaError_t seg_run(const Layout& L, void* scratch, const SegInputs& in, cudaStream_t st);
int seg_launch_count(const Layout& L);
// Group bounds of the sorted batch: astart[p] = first sorted position of p's group, gend[astart] = one past
// its last position, perm[p] = original row; up to 4 float columns are gathered into sorted order.  One launch.
struct GatherCols { const float* src[4]; float* dst[4]; };
cudaError_t seg_bounds(const Layout& L, void* scratch, int use_label, u32* astart, u32* gend, u32* perm,
                       const GatherCols& gc, cudaStream_t st);

inline int check_align(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) ? RN_ERR_ALIGN : RN_OK; }

}  // namespace rn
