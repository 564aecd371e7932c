// GAUC -- group AUC, the evaluation metric the in-batch ranking losses are meant to move (the reference's README.md:5, 8
// quotes its online GAUC uplift; the reference ships no implementation of the metric, so the definition is stated here):
//   for every group g (same composite key, rows that can pair: row_ok, non-NaN label)
//     pairs_g = {(i, j) in g : y_i > y_j}                     -- the pair set of pairwise_loss_from_batch.py:189
//     AUC_g   = (#{s_i > s_j} + 1/2 #{s_i == s_j}) / |pairs_g|
//   GAUC = sum_g |g| AUC_g / sum_g |g| over the groups with |pairs_g| > 0 (weight = rows of the group); 0 if there is none.
// Binary labels give the usual per-user AUC; graded labels the pairwise accuracy over label-ordered pairs.
//
// Same K1 segmentation as the pairwise loss (k_seg<HeadsTail>: counting path or radix sort), then ONE cooperative kernel
// that walks the same staircase of (I-block x J-block) tiles with integer compares instead of the logistic loss: the
// concordance counts are exact integers (2 x concordant + ties, 64 bit).  HBM / issue bound; no SFU work at all.
#include <stdlib.h>
#include "common.cuh"

namespace rn {

constexpr int kGaucThreads = 512;

__global__ void __launch_bounds__(kGaucThreads) k_gauc(GaucArgs A) {
  extern __shared__ __align__(16) u32 dsm[];
  __shared__ double red_d[3][kGaucThreads / 32];
  __shared__ u64 red_u[2][kGaucThreads / 32];
  Ctl* ctl = A.ctl;
  const u32 B = A.B, ln = lane_id(), w = threadIdx.x >> 5;
  const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  u32 epoch = 0;
  grid_dep_wait();
  stamp(ctl, 20);
  // ---- phase 0: per-group accumulators (indexed by sorted position) ----------------------------------------------
  for (u32 p = gtid; p < B; p += gthreads) { A.acc2[p] = 0ull; A.npg[p] = 0ull; A.gsz[p] = 0u; }
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
  stamp(ctl, 11);
  // ---- phase 1a: rows and ordered label pairs per group (position arithmetic: n = rows below the row's level) ---------
  // (consecutive sorted rows share their group: one atomic per group and warp, not 8 000 on the counter of a big group)
  for (u32 p0 = 0; p0 < B; p0 += gthreads) {
    const u32 p = p0 + gtid;
    uint2 an = make_uint2(kEmpty, 0);
    if (p < B) an = A.aj[p];
    const u32 m = __match_any_sync(0xFFFFFFFFu, an.x);
    const u32 tot = __reduce_add_sync(m, an.y);
    if (an.x != kEmpty && ln == (u32)(__ffs(m) - 1)) {
      atomicAdd(A.gsz + an.x, (u32)__popc(m));
      if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(A.npg + an.x), (unsigned long long)tot);
    }
  }
  stamp(ctl, 12);
  // ---- phase 1b: the tiles.  Every CTA prefix-sums the tile counts of the virtual blocks (the two J ranges of every
  //      64-row I-block) in shared memory; every warp of the grid then takes the same number of consecutive tiles.
  const u32 nvb = 2 * A.nib;
  u32* s_pi = dsm;                          // [nvb + 1]
  u32* s_sc = dsm + nvb + 1;                // scan partials
  for (u32 v = threadIdx.x; v < nvb; v += kGaucThreads) { u32 jf; s_pi[v] = vblock_tiles(A.blk, v, jf); }
  __syncthreads();
  block_excl_scan(s_pi, nvb, s_sc);
  {
    const u32 T = s_pi[nvb];
    const u32 nwarps = gridDim.x * (kGaucThreads / 32);
    const u32 wg = w * gridDim.x + blockIdx.x;                     // (pieces dealt to the CTAs round-robin)
    u32 t0 = (u32)(((u64)T * wg) / nwarps);
    const u32 t1 = (u32)(((u64)T * (wg + 1)) / nwarps);
    u32 v = t0 < t1 ? last_le(s_pi, 0, nvb, t0) : nvb;
    while (t0 < t1) {
      while (s_pi[v + 1] <= t0) ++v;                               // (skips virtual blocks without tiles)
      u32 jf;
      const u32 nt = vblock_tiles(A.blk, v, jf);
      const u32 jb0 = jf + (t0 - s_pi[v]), jb1 = jf + min(nt, t1 - s_pi[v]);
      const u32 pi0 = (v >> 1) * kIB + ln, pi1 = pi0 + 32;
      uint2 an0 = make_uint2(0, 0), an1 = make_uint2(0, 0); float si0 = 0.f, si1 = 0.f;
      if (pi0 < B) { an0 = A.aj[pi0]; si0 = A.ss[pi0]; }
      if (pi1 < B) { an1 = A.aj[pi1]; si1 = A.ss[pi1]; }
      const u32 lo0 = an0.x, lo1 = an1.x;
      u32 c0 = 0, c1 = 0;
      float sjn = (jb0 * 32 + ln) < B ? A.ss[jb0 * 32 + ln] : 0.f;
      for (u32 jb = jb0; jb < jb1; ++jb) {
        const u32 pjm = jb * 32 + ln;
        const float sjm = sjn;
        if (jb + 1 < jb1) sjn = (pjm + 32) < B ? A.ss[pjm + 32] : 0.f;       // next J-block in flight while this one is compared
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
          const float sj = __shfl_xor_sync(0xFFFFFFFFu, sjm, t);
          const u32 pj = pjm ^ (u32)t;
          const u32 k0 = si0 > sj ? 2u : (si0 == sj ? 1u : 0u), k1 = si1 > sj ? 2u : (si1 == sj ? 1u : 0u);
          c0 += (pj - lo0 < an0.y) ? k0 : 0u;                 // (unsigned: lo <= pj < lo + n)
          c1 += (pj - lo1 < an1.y) ? k1 : 0u;
        }
      }
      // (the rows of an I-block mostly share one group: one atomic per group and warp)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const u32 key = h ? (c1 ? lo1 : kEmpty) : (c0 ? lo0 : kEmpty), cv = h ? c1 : c0;
        const u32 m = __match_any_sync(0xFFFFFFFFu, key);
        const u32 tot = __reduce_add_sync(m, cv);
        if (key != kEmpty && ln == (u32)(__ffs(m) - 1)) atomicAdd(reinterpret_cast<unsigned long long*>(A.acc2 + key), (unsigned long long)tot);
      }
      t0 = s_pi[v] + (jb1 - jf);
      ++v;
    }
  }
  stamp(ctl, 21);
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
  stamp(ctl, 22);
  // ---- phase 2: per-group AUC at the group heads, weighted by the rows of the group -----------------------------------
  double wsum = 0.0, den = 0.0, asum = 0.0; u64 np = 0, c2 = 0; u32 nv = 0;
  for (u32 p = gtid; p < B; p += gthreads) {
    if (A.aj[p].x != p) continue;                 // (a group's first sorted position)
    const u64 n = A.npg[p];
    if (!n) continue;
    const u64 a2 = A.acc2[p];
    const double auc = (double)a2 / (2.0 * (double)n);
    const double sz = (double)A.gsz[p];
    wsum += sz * auc; den += sz; asum += auc; np += n; c2 += a2; ++nv;
  }
  wsum = warp_sum(wsum); den = warp_sum(den); asum = warp_sum(asum); np = warp_sum(np); c2 = warp_sum(c2);
  nv = __reduce_add_sync(0xFFFFFFFFu, nv);
  if (ln == 0) { red_d[0][w] = wsum; red_d[1][w] = den; red_d[2][w] = asum; red_u[0][w] = np; red_u[1][w] = c2; }
  __syncthreads();
  const uint2 z2 = make_uint2(0, 0);
  for (u32 q = gtid; q < nvb; q += gthreads) A.blk_w[q] = z2;            // leave the arena clean (see k_pair)
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cc = 0; u64 x = 0, y = 0;
    for (int q = 0; q < kGaucThreads / 32; ++q) { a += red_d[0][q]; b += red_d[1][q]; cc += red_d[2][q]; x += red_u[0][q]; y += red_u[1][q]; }
    if (b != 0.0) { atomicAdd(&ctl->acc_d[0], a); atomicAdd(&ctl->acc_d[1], b); atomicAdd(&ctl->acc_d[2], cc); }
    (void)x;          // (the pair total is already in ctl->n_pair: the segmentation kernel counts it by position arithmetic)
    if (y) atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->n_tiles), (unsigned long long)y);
  }
  if (ln == 0 && nv) atomicAdd(&ctl->n_valid, nv);
  grid_sync(&ctl->bar2_cnt, epoch, &ctl->err);
  // ---- behind the last barrier: the scalars (CTA 0), clean arena (all CTAs) ----------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double a = *reinterpret_cast<volatile double*>(&ctl->acc_d[0]), b = *reinterpret_cast<volatile double*>(&ctl->acc_d[1]);
    const double cc = *reinterpret_cast<volatile double*>(&ctl->acc_d[2]);
    const u32 nvt = ld_relaxed(&ctl->n_valid);
    *A.gauc = b > 0.0 ? (float)(a / b) : 0.f;
    if (A.auc_mean) *A.auc_mean = nvt ? (float)(cc / (double)nvt) : 0.f;
    *A.n_valid = (int32_t)nvt;
    if (A.n_pair) *A.n_pair = (int64_t)*reinterpret_cast<volatile u64*>(&ctl->n_pair);
    if (A.conc2) *A.conc2 = (int64_t)*reinterpret_cast<volatile u64*>(&ctl->n_tiles);
  }
  if (A.fast) clean_records_grid(A.rec, A.rec2_off, A.glist, A.gcount, A.ngt);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->fin_done, 1u) == gridDim.x - 1) { __threadfence(); ctl->ts[23] = globaltimer(); ctl_finish(ctl); }
  }
}

cudaError_t launch_gauc(const GaucArgs& A, cudaStream_t st) {
  static thread_local int bps_dev[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!bps_dev[dev]) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gauc, kGaucThreads, 70 * 1024) != cudaSuccess) return cudaGetLastError();
    bps_dev[dev] = nb < 1 ? 1 : (nb > 2 ? 2 : nb);
  }
  GaucArgs a = A;
  void* args[] = {&a};
  const size_t smem = sizeof(u32) * (2 * (size_t)A.nib + 1 + 40);          // tile prefix over the virtual blocks
  static thread_local size_t smem_set[64] = {0};
  if (smem > 48 * 1024 && smem > smem_set[dev]) {
    if (cudaFuncSetAttribute(k_gauc, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024) != cudaSuccess) return cudaGetLastError();
    smem_set[dev] = 72 * 1024;
  }
  return launch_coop((const void*)k_gauc, device_sm_count() * bps_dev[dev], kGaucThreads, args, st, smem);
}

}  // namespace rn
