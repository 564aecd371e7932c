// K1 (counting form) -- group segmentation WITHOUT a sort: device helpers shared by the consumers of the counting path.
//
// Replaces the same reference code as segment.cuh (the (B,B) group-equality matrix, pairwise_loss_from_batch.py:33-37,
// 68-73, and the label-order mask, :187-190).  What the pair kernel needs is not a sorted batch but a GROUPED one: the
// rows of a group contiguous, ordered by label level inside the group.  When the labels are small integers (binary
// clicks, graded relevance 0-4: the BASELINE configurations) that is a two-level counting sort whose histogram is the
// hash table itself:
//   count    every 512-row tile aggregates its keys in shared memory; ONE 128-bit compare-and-swap per distinct key and
//            tile claims / finds the group's record (key and creator row travel in the same word, so a probe never
//            needs a second dependent load), one atomicAdd per (group, label level, tile) reserves the tile's ranks
//            inside the level.  Created records are listed per tile.
//   offsets  one thread per created record: level counts -> level starts, exact pair total of the group (position
//            arithmetic: sum over levels of count x rows below), occurrence weight c_h^power, and the group's base by a
//            warp-aggregated atomicAdd on one cursor (groups land in allocation order: any order is a valid grouping).
//   scatter  every row: position = base + level start + rank; it writes its sorted columns, its negative range
//            (group start, rows below its level) and folds that range into the two J ranges of its I-block with
//            atomicMax.  No group heads to find, no max-scans, no searches: the record already holds them.
// Two grid barriers instead of five, no k_init: the records are zeroed again by the pair kernel's last phase (the
// arena is persistent, see rn_pairwise_args.scratch_persistent).  Anything outside the menu (non-integer labels, more
// than 8 levels, non-positive row weights) raises ctl->fallback in the count phase and the kernel continues with the
// radix path of segment.cuh.
#pragma once
#include "common.cuh"

namespace rn {

struct W128 { u64 lo, hi; };

// 128-bit compare-and-swap (sm_90+: ATOMG.E.CAS.128); returns the previous value of the word.
__device__ __forceinline__ W128 cas128(void* p, W128 cmp, W128 val) {
  W128 old;
  asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 v, {%5, %6};\n\t"
               "atom.relaxed.gpu.global.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
               : "=l"(old.lo), "=l"(old.hi) : "l"(p), "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi) : "memory");
  return old;
}

// Find or create the record of `key` (open addressing, linear probing; load factor <= 1/2).  Every probe is the CAS
// itself: it returns the word atomically, so there are no torn reads and no separate key compare.
__device__ __forceinline__ u32 grec_insert(GRec* rec, u32 capmask, u64 h, u64 key, u32 row, bool& created, u32* err) {
  u32 s = (u32)h & capmask;
  const W128 empty{0ull, 0ull}, mine{key, (u64)(row + 1u)};
  for (u32 probes = 0;; ++probes) {
    const W128 old = cas128(rec + s, empty, mine);
    if ((old.lo | old.hi) == 0ull) { created = true; return s; }
    if (old.lo == key && (u32)old.hi != 0u) { created = false; return s; }
    if (probes > capmask) { atomicOr(err, 4u); created = false; return s; }     // (arena not clean: never with a valid one)
    s = (s + 1u) & capmask;
  }
}

// Label level of the counting path: integer-valued labels -1 .. 6 -> 0 .. 7; anything else is outside the menu.
__device__ __forceinline__ bool label_level(float y, int& li) {
  const float yf = y + 1.0f;
  const int q = (int)yf;
  li = q;
  return (float)q == yf && q >= 0 && q < kLevels;
}

}  // namespace rn
