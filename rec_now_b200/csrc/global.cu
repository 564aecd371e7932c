// Global in-batch mode behind the C ABI (rn_global_*): ONE call per step and rank, no framework on the data path.
//
// Semantics: pairwise_loss (pairwise_loss_from_batch.py:228-279) evaluated on the concatenation of all ranks' rows
// (rank order = row order) -- SURVEY.md section 8e; the reference itself has no multi-GPU mode.
//
// One process per GPU.  The caller owns one SYMMETRIC buffer per rank (rn_global_buffer_bytes), maps every rank's
// buffer into every process (cudaIpc / CUDA VMM / torch symmetric memory: the plumbing is the caller's), zeroes its
// own once, and hands the mapped base pointers to every call.  A step enqueues, on the caller's stream:
//   k_pack        this rank's columns -> its packed row block inside its own buffer
//   k_xbar        device-side barrier over NVLink: every rank stores the step number into a flag word of every peer's
//                 buffer (st.release.sys) and waits for its own flag words (ld.acquire.sys)
//   k_init / k_seg / k_pair   (one cached CUDA graph) -- k_init gathers all ranks' blocks with peer loads, every rank
//                 segments the same global rows, the pair kernel scores this rank's share of the cost line and leaves
//                 its partial gradients, chunked per owner rank, in its own buffer
//   k_xbar
//   k_reduce_out  this rank's chunk summed over the peers' buffers (the reduce-scatter, read straight from peer
//                 memory) -> dlogits of the local rows and the global loss
// Input blocks, output chunks and flag words alternate between consecutive steps, so a rank that runs ahead never
// overwrites what a peer still reads.
#include "common.cuh"

namespace rn {

struct GlobalLayout { size_t flags, in[2], out[2], cnt[2], total; size_t stride, chunk; };

static size_t packed_stride(int64_t b_loc, int K, bool has_w, bool has_ok) {
  size_t o = (size_t)8 * K * b_loc + 4 * b_loc + 4 * b_loc;
  if (has_w) o += 4 * b_loc;
  if (has_ok) o += b_loc;
  return (o + 15) / 16 * 16;
}

static GlobalLayout global_layout(int64_t b_loc, int K, int world, bool has_w, bool has_ok) {
  GlobalLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  L.stride = packed_stride(b_loc, K, has_w, has_ok);
  L.chunk = (size_t)b_loc + 4;
  L.flags = take(sizeof(u32) * 2 * 4 * 8);              // [parity][channel][rank]
  for (int p = 0; p < 2; ++p) L.in[p] = take(L.stride);
  for (int p = 0; p < 2; ++p) L.out[p] = take(sizeof(float) * world * L.chunk);
  for (int p = 0; p < 2; ++p) L.cnt[p] = take(sizeof(u32) * (size_t)world * b_loc);     // per-row pair counts (score-dependent pair sets)
  L.total = o;
  return L;
}

struct XbarArgs { u32* peer_flags[8]; u32* my_flags; u32 world, rank, epoch; u32* err; };

// Device-side barrier across the ranks of one box.  Thread r tells rank r that this rank has reached `epoch` and waits
// until rank r has said the same.  Everything this rank enqueued before is complete (stream order: the previous kernel
// has finished and flushed), so a peer that sees the flag may read this rank's buffer.  Bounded spin -> *err.
__global__ void __launch_bounds__(32) k_xbar(XbarArgs X) {
  const u32 r = threadIdx.x;
  if (r < X.world) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(X.peer_flags[r] + X.rank), "r"(X.epoch) : "memory");
    u32 v = 0, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(X.my_flags + r) : "memory");
      if ((int)(v - X.epoch) >= 0) break;
      if (++spins > (1u << 24)) { if (X.err) atomicOr(X.err, 8u); break; }
    }
  }
}

struct ReduceOutArgs { const float4* src[8]; int world; u32 n4; u32 b_loc; float4* dst; float* loss; };
// dst[i] = sum over ranks of their chunk for this rank; element b_loc of the chunk is the loss.
__global__ void __launch_bounds__(256) k_reduce_out(ReduceOutArgs P) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n4; i += gridDim.x * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) if (r < P.world) v[r] = P.src[r][i];          // all peer loads in flight at once
    float4 a = v[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) if (r < P.world) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
    if (4 * i < P.b_loc) P.dst[i] = a;                                         // (b_loc is a multiple of 16)
    else if (4 * i == P.b_loc) *P.loss = a.x;
  }
}

}  // namespace rn

using namespace rn;

extern "C" size_t rn_global_buffer_bytes(int64_t B_loc, int32_t K, int32_t world, int32_t has_rw_pos, int32_t has_row_ok) {
  if (B_loc <= 0 || (B_loc & 15) || K <= 0 || K > 8 || world < 1 || world > 8) return 0;
  return global_layout(B_loc, K, world, has_rw_pos != 0, has_row_ok != 0).total;
}

extern "C" size_t rn_global_gather_bytes(int64_t B_loc, int32_t K, int32_t world, int32_t has_rw_pos, int32_t has_row_ok) {
  if (B_loc <= 0 || (B_loc & 15) || K <= 0 || K > 8 || world < 1 || world > 8) return 0;
  return (size_t)world * packed_stride(B_loc, K, has_rw_pos != 0, has_row_ok != 0);
}

extern "C" int rn_global_pairwise_fwd_bwd(const rn_global_args* g, void* scratch, size_t scratch_bytes, void* stream) {
  RN_NVTX_RANGE("rn_global_pairwise_fwd_bwd");
  if (!g || g->world < 1 || g->world > 8 || g->rank < 0 || g->rank >= g->world || !g->gather_buf || g->step < 0) return RN_ERR_ARG;
  const rn_pairwise_args& l = g->local;
  if (l.B <= 0 || (l.B & 15) || l.K <= 0 || l.K > 8) return RN_ERR_ARG;
  if (!l.keys || !l.logits || !l.labels || !l.loss || !l.n_pair_f32 || !l.n_pair || !l.dlogits) return RN_ERR_ARG;
  const bool dyn = l.rw_neg || l.only_wrong;        // score- / weight-dependent pair set: two stages with the ranks' counts summed between
  if (l.block_rows || l.out_chunk || l.gather_dst || l.part_count > 1) return RN_ERR_ARG;
  for (int r = 0; r < g->world; ++r) if (!g->peer_buf[r] || check_align(g->peer_buf[r])) return RN_ERR_ARG;
  const bool has_w = l.rw_pos != nullptr, has_ok = l.row_ok != nullptr;
  const GlobalLayout L = global_layout(l.B, l.K, g->world, has_w, has_ok);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int p = (int)(g->step & 1);
  const u32 epoch = (u32)(g->step / 2 + 1);                      // (each parity has its own flag words)
  char* mine = static_cast<char*>(g->peer_buf[g->rank]);
  // 1. pack
  int rc = rn_pack_row_block(l.keys, l.K, l.logits, l.labels, l.rw_pos, l.row_ok, l.B, mine + L.in[p], (int64_t)L.stride, stream);
  if (rc) return rc;
  // 2. barrier: every rank's block of this step is written
  auto xbar = [&](int channel) {
    XbarArgs X{};
    for (int r = 0; r < g->world; ++r)
      X.peer_flags[r] = reinterpret_cast<u32*>(static_cast<char*>(g->peer_buf[r]) + L.flags) + (p * 4 + channel) * 8;
    X.my_flags = reinterpret_cast<u32*>(mine + L.flags) + (p * 4 + channel) * 8;
    X.world = (u32)g->world; X.rank = (u32)g->rank; X.epoch = epoch;
    X.err = scratch ? &static_cast<Ctl*>(scratch)->rep_err : nullptr;
    k_xbar<<<1, 32, 0, st>>>(X);
  };
  xbar(0);
  // 3. the kernels, on the blocked global rows
  rn_pairwise_args a = l;
  const size_t koff = 0, soff = (size_t)8 * l.K * l.B, yoff = soff + 4 * l.B, woff = yoff + 4 * l.B;
  const size_t okoff = woff + (has_w ? 4 * l.B : 0);
  char* gb = static_cast<char*>(g->gather_buf);
  a.B = l.B * g->world;
  a.keys = reinterpret_cast<const int64_t*>(gb + koff); a.logits = reinterpret_cast<const float*>(gb + soff);
  a.labels = reinterpret_cast<const float*>(gb + yoff);
  a.rw_pos = has_w ? reinterpret_cast<const float*>(gb + woff) : nullptr;
  if (l.rw_neg) return RN_ERR_UNSUPPORTED;          // (a negative-side weight column is not part of the packed row block)
  a.row_ok = has_ok ? reinterpret_cast<const uint8_t*>(gb + okoff) : nullptr;
  a.part_rank = g->rank; a.part_count = g->world;
  a.block_rows = l.B; a.block_stride = (int64_t)L.stride; a.out_chunk = (int64_t)L.chunk;
  for (int r = 0; r < g->world; ++r) a.peer_blocks[r] = static_cast<char*>(g->peer_buf[r]) + L.in[p];
  a.gather_dst = gb;
  a.dlogits = reinterpret_cast<float*>(mine + L.out[p]);
  a.row_pairs = nullptr;
  // (this rank's partial loss goes to a spare word of the flag block -- channel 3 is unused; the SUM rides in the chunks)
  a.loss = reinterpret_cast<float*>(reinterpret_cast<u32*>(mine + L.flags) + (p * 4 + 3) * 8);
  // (the arena is persistent: the caller zeroed it once, every call leaves it clean -- the sort-free counting
  //  segmentation then runs on the gathered rows, each rank scoring the pairs whose negative row it owns)
  a.scratch_persistent = g->local.scratch_persistent; a.scratch_rows = 0;
  if (!dyn) {
    rc = rn_pairwise_fwd_bwd(&a, scratch, scratch_bytes, stream);
    if (rc) return rc;
  } else {
    // stage 1: the kernels up to this rank's partial per-row pair counts, published by original row in its buffer;
    // barrier; stage 2: sum of the ranks' counts -> exact n and c_h -> weights -> partial gradients (PW:197-203, 282-291)
    DynSplit sp{};
    sp.world = g->world;
    sp.xcnt_out = reinterpret_cast<u32*>(mine + L.cnt[p]);
    for (int r = 0; r < g->world; ++r) sp.xcnt_peer[r] = reinterpret_cast<const u32*>(static_cast<char*>(g->peer_buf[r]) + L.cnt[p]);
    rc = pairwise_call(&a, scratch, scratch_bytes, stream, &sp, 1);
    if (rc) return rc;
    xbar(2);
    rc = pairwise_call(&a, scratch, scratch_bytes, stream, &sp, 2);
    if (rc) return rc;
  }
  // 4. barrier: every rank's partial gradients are written
  xbar(1);
  // 5. this rank's chunk, summed over the peers' buffers
  ReduceOutArgs R{};
  R.world = g->world; R.n4 = (u32)(L.chunk / 4); R.b_loc = (u32)l.B;
  for (int r = 0; r < g->world; ++r)
    R.src[r] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(static_cast<char*>(g->peer_buf[r]) + L.out[p]) + (size_t)g->rank * L.chunk);
  R.dst = reinterpret_cast<float4*>(l.dlogits); R.loss = l.loss;
  int grid = (int)((R.n4 + 255) / 256);
  const int cap = device_sm_count() * 4;
  if (grid > cap) grid = cap;
  k_reduce_out<<<grid, 256, 0, st>>>(R);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}
