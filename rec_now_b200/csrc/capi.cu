// Small C-ABI entry points: version, error strings, float-id canonicalisation, SFU peak probe.
#include "common.cuh"

namespace rn {

template <typename F, typename I>
__global__ void __launch_bounds__(256) k_canon(const F* __restrict__ ids, int64_t B, int64_t* __restrict__ out,
                                               uint8_t* __restrict__ ok, int and_into) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  F v = ids[i];
  bool fin = (v - v) == (F)0;          // false for NaN and +-inf
  // (and_into bit 1: +-inf are ids like any other -- tf.unique, LW:109, compares with ==, and inf == inf; the pairwise
  //  path compares g_i - g_j with 0, PW:33-35, and inf - inf is NaN)
  if ((and_into & 2) && v == v) fin = true;
  v = fin ? v + (F)0 : (F)0;           // -0.0 -> +0.0
  double dv = (double)v;               // float32 ids widen exactly; equality is preserved
  out[i] = __double_as_longlong(dv);
  if (ok) ok[i] = (and_into & 1) ? (uint8_t)(ok[i] && fin) : (uint8_t)fin;
}

__global__ void __launch_bounds__(256) k_mufu(int iters, float* sink) {
  float a = 0.5f + 1e-3f * threadIdx.x, b = 1.5f, c = 0.25f, d = 0.75f;
  for (int i = 0; i < iters; ++i) {
    // four independent ex2 -> lg2 -> rcp chains: 12 MUFU per iteration, values stay in (0.4, 2.5)
    a = mufu_rcp(mufu_lg2(mufu_ex2(a) + 1.0f) + 0.5f);
    b = mufu_rcp(mufu_lg2(mufu_ex2(b) + 1.0f) + 0.5f);
    c = mufu_rcp(mufu_lg2(mufu_ex2(c) + 1.0f) + 0.5f);
    d = mufu_rcp(mufu_lg2(mufu_ex2(d) + 1.0f) + 0.5f);
  }
  if (a + b + c + d == -1.f) sink[0] = a;      // never true; keeps the chains alive
}

// One rank's row block of the global mode: the columns copied back to back as 16-byte words (see rn_pack_row_block).
struct PackCols { const uint4* src[12]; u32 n16[12]; int ncol; u32 total16, pad_from16; };
__global__ void __launch_bounds__(256) k_pack(PackCols P, uint4* __restrict__ dst) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < P.total16; i += gridDim.x * blockDim.x) {
    u32 o = i;
    uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll 1
    for (int c = 0; c < P.ncol; ++c) {
      if (o < P.n16[c]) { v = P.src[c][o]; break; }
      o -= P.n16[c];
    }
    dst[i] = v;
  }
}

struct PeerChunks { const float4* src[8]; int world; };
__global__ void __launch_bounds__(256) k_reduce_chunks(PeerChunks P, u32 n4, float4* __restrict__ dst) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) if (r < P.world) v[r] = P.src[r][i];          // all peer loads in flight at once
    float4 a = v[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) if (r < P.world) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
    dst[i] = a;
  }
}

}  // namespace rn

using namespace rn;

extern "C" int rn_reduce_peer_chunks(const void* const* peer_out, int32_t world, int32_t my_rank, int64_t chunk, float* dst,
                                     void* stream) {
  if (!peer_out || !dst || world < 1 || world > 8 || my_rank < 0 || my_rank >= world || chunk <= 0 || (chunk & 3)) return RN_ERR_ARG;
  if (check_align(dst)) return RN_ERR_ALIGN;
  PeerChunks P{};
  P.world = world;
  for (int r = 0; r < world; ++r) {
    if (!peer_out[r] || check_align(peer_out[r])) return RN_ERR_ARG;
    P.src[r] = reinterpret_cast<const float4*>(static_cast<const float*>(peer_out[r]) + (size_t)my_rank * chunk);
  }
  const u32 n4 = (u32)(chunk / 4);
  int grid = (int)((n4 + 255) / 256);
  const int cap = device_sm_count() * 4;
  if (grid > cap) grid = cap;
  k_reduce_chunks<<<grid, 256, 0, (cudaStream_t)stream>>>(P, n4, reinterpret_cast<float4*>(dst));
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_pack_row_block(const int64_t* keys, int32_t K, const float* logits, const float* labels,
                                 const float* rw_pos, const uint8_t* row_ok, int64_t B_loc, void* block_out,
                                 int64_t stride, void* stream) {
  if (!keys || !logits || !labels || !block_out || K <= 0 || K > 8 || B_loc <= 0 || (B_loc & 15) || (stride & 15)) return RN_ERR_ARG;
  const void* ptrs[] = {keys, logits, labels, rw_pos, row_ok, block_out};
  for (const void* p : ptrs) if (p && check_align(p)) return RN_ERR_ALIGN;
  PackCols P{};
  int c = 0;
  for (int k = 0; k < K; ++k) { P.src[c] = reinterpret_cast<const uint4*>(keys + (size_t)k * B_loc); P.n16[c++] = (u32)(B_loc / 2); }
  P.src[c] = reinterpret_cast<const uint4*>(logits); P.n16[c++] = (u32)(B_loc / 4);
  P.src[c] = reinterpret_cast<const uint4*>(labels); P.n16[c++] = (u32)(B_loc / 4);
  if (rw_pos) { P.src[c] = reinterpret_cast<const uint4*>(rw_pos); P.n16[c++] = (u32)(B_loc / 4); }
  if (row_ok) { P.src[c] = reinterpret_cast<const uint4*>(row_ok); P.n16[c++] = (u32)(B_loc / 16); }
  P.ncol = c;
  u64 used = 0;
  for (int k = 0; k < c; ++k) used += P.n16[k];
  if ((u64)stride / 16 < used || (u64)stride / 16 > 0xFFFFFFFFull) return RN_ERR_ARG;
  P.total16 = (u32)(stride / 16);
  int grid = (int)((P.total16 + 255) / 256);
  const int cap = device_sm_count() * 8;
  if (grid > cap) grid = cap;
  k_pack<<<grid, 256, 0, (cudaStream_t)stream>>>(P, static_cast<uint4*>(block_out));
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_version(void) { return RN_VERSION; }

extern "C" const char* rn_strerror(int code) {
  switch (code) {
    case RN_OK: return "ok";
    case RN_ERR_ARG: return "invalid argument";
    case RN_ERR_ALIGN: return "device pointer not 16-byte aligned";
    case RN_ERR_SCRATCH: return "scratch arena too small";
    case RN_ERR_LAUNCH: return "CUDA launch/runtime error";
    case RN_ERR_UNSUPPORTED: return "request outside the fused menu";
    case RN_ERR_NO_DEVICE: return "no sm_100 device";
    case RN_ERR_INTERNAL: return "device-side consistency check failed";
  }
  return "unknown error";
}

extern "C" int rn_canon_keys_f32(const float* ids, int64_t B, int64_t* keys_out, uint8_t* row_ok, int and_into,
                                 void* stream) {
  if (!ids || !keys_out || B <= 0) return RN_ERR_ARG;
  k_canon<float, int><<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, B, keys_out, row_ok, and_into);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_canon_keys_f64(const double* ids, int64_t B, int64_t* keys_out, uint8_t* row_ok, int and_into,
                                 void* stream) {
  if (!ids || !keys_out || B <= 0) return RN_ERR_ARG;
  k_canon<double, int><<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, B, keys_out, row_ok, and_into);
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_bench_mufu(int32_t iters, float* sink, int64_t* mufu_ops_out_host, void* stream) {
  if (iters <= 0 || !sink) return RN_ERR_ARG;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8;
  k_mufu<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
  if (mufu_ops_out_host) *mufu_ops_out_host = (int64_t)blocks * 256 * 12 * (int64_t)iters;
  return cudaGetLastError() == cudaSuccess ? RN_OK : RN_ERR_LAUNCH;
}

extern "C" int rn_last_device_error(void* scratch, int32_t* err_host, void* stream) {
  if (!scratch || !err_host) return RN_ERR_ARG;
  // the control block sits at offset 0 of every arena layout
  Ctl h;
  if (cudaMemcpyAsync(&h, scratch, sizeof(Ctl), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return RN_ERR_LAUNCH;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return RN_ERR_LAUNCH;
  *err_host = (int32_t)(h.err | h.rep_err);      // (rep_err: filed by the last CTA of a finished pairwise call)
  return RN_OK;
}

extern "C" int rn_debug_timestamps(void* scratch, uint64_t* ts_host, int32_t capacity, void* stream) {
  if (!scratch || !ts_host || capacity <= 0) return RN_ERR_ARG;
  Ctl h;
  if (cudaMemcpyAsync(&h, scratch, sizeof(Ctl), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return RN_ERR_LAUNCH;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return RN_ERR_LAUNCH;
  for (int i = 0; i < capacity && i < 24; ++i) ts_host[i] = h.ts[i];
  for (int i = 24; i < capacity && i < 32; ++i) ts_host[i] = h.dbg[i - 24];
  if (capacity > 32) ts_host[32] = ((uint64_t)(h.unit_c | h.rep_unit_c) << 32) | (h.n_units | h.rep_n_units);
  if (capacity > 33) ts_host[33] = h.n_tiles | h.rep_n_tiles;
  if (capacity > 35) ts_host[35] = h.path | h.rep_path;          // segmentation path of the call: 1 counting, 2 radix
  if (capacity > 34) ts_host[34] = (uint64_t)make_layout(1, 1).gstat;   // (layout probe for debug tools: see scripts/)
  return RN_OK;
}

extern "C" int64_t rn_debug_arena_offset(int64_t B, int32_t K, int32_t which) {
  if (B <= 0 || K <= 0) return -1;
  const Layout L = make_layout(B, K);
  switch (which) { case 0: return (int64_t)L.gstat; case 1: return (int64_t)L.rec; case 2: return (int64_t)L.blk; case 3: return (int64_t)L.bnd; }
  return -1;
}

extern "C" int64_t rn_debug_graph_launches(void) { return (int64_t)graph_launch_count(); }
