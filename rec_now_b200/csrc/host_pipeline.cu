// Host-buffer front end of rn_pairwise_fwd_bwd (rn_host_pairwise_*, see recnow_b200.h).
//
// The reference's pairwise_loss (pairwise_loss_from_batch.py:228-279) is called with tensors that live wherever the
// framework put them; a TF2 CPU placement, a data loader or a parameter server hands over HOST buffers.  This front
// end owns everything between those buffers and the kernels: `depth` slots of device memory (input columns, outputs,
// the scratch arena of the call), one stream for the host->device copies, one for the kernels and one for the
// device->host copies.  A submit enqueues copy-in -> rn_pairwise_fwd_bwd -> copy-out for one batch and returns at
// once; with depth >= 2 the copies of one batch overlap the kernels of its neighbours.  Nothing is allocated after
// create, and the only host synchronisation is the wait for a slot's results.
#include <new>
#include <string.h>
#include <chrono>
#include <stdio.h>
#include "common.cuh"

namespace {

struct Slot {
  char* dev = nullptr;                 // one allocation: [keys][logits][labels][rw_pos][rw_neg][row_ok][outs][dlogits][row_pairs][weight_lut][scratch]
  float* outs_host = nullptr;          // pinned staging of the three scalars: [loss f32][n_pair f32][n_pair i64]
  float* loss = nullptr; float* n_pair_f32 = nullptr; int64_t* n_pair = nullptr;      // where the caller wants them
  cudaEvent_t in_done = nullptr, cmp_done = nullptr, out_done = nullptr;
  // the slot's own compute stream: the launches of consecutive batches are processed side by side, so the first kernel of
  // batch k+1 is ready to start the moment the last kernel of batch k leaves the SMs (on one stream the ~7 us between two
  // graph launches were idle GPU time); each stream has its own instance of the library's cached graphs (set_graph_lane)
  cudaStream_t s_cmp = nullptr;
  bool busy = false;
  // Whole-step graph (opt-in, RN_HOST_STEP_GRAPH): copy-in -> kernels -> copy-out of this slot as ONE cudaGraphLaunch on the
  // slot's stream -- a caller that submits the same pinned buffers step after step (a loader refilling its staging
  // buffers: HostPairwise.bind) pays three driver calls per step instead of fourteen.  Built when a slot sees the same
  // arguments a second time; any other submit takes the eager path below.
  rn_pairwise_args seen{};            // the host-side arguments of the slot's last submit
  bool seen_valid = false;
  bool seen_pageable = false;         // ... and they name pageable memory: no graph for them (asked once, not per submit)
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
  rn_pairwise_args graph_args{};      // ... the arguments the graph was captured for
};

// true if [ptr, ptr + bytes) is page-locked host memory (a copy from pageable memory cannot be a graph node here)
bool pinned(const void* ptr) {
  if (!ptr) return true;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

}  // namespace

struct rn_host_pairwise {
  int64_t B_max = 0; int32_t K = 0; int32_t depth = 0; int device = 0;
  size_t o_keys = 0, o_logits = 0, o_labels = 0, o_rwp = 0, o_rwn = 0, o_ok = 0, o_outs = 0, o_dl = 0, o_rp = 0, o_scr = 0, o_lut = 0;
  size_t scratch_bytes = 0, slot_bytes = 0;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  Slot* slots = nullptr;
  int32_t next = 0;
  bool graph_broken = false;           // a whole-step graph could not be built: eager submits from now on
  int64_t graph_steps = 0;             // submits that went out as one launch of a whole-step graph
  // RN_HOST_STEP_GRAPH (read at create): unset / 0 never (the default), 1 always, 2 automatic -- whole-step graphs only
  // once the eager submits have shown themselves to be host bound (their mean host time above graph_threshold_us, default
  // 25).  Opt-in because the measurements do not favour it: on a quiet host an eager submit costs 14 us, the pipeline is
  // device bound at 46.4 us per step and the graphs are SLOWER (48.6-50 us in one session, 87 us in another, where the
  // eager path of the same process still measured 46.4); it is meant for hosts whose eager submits are the bottleneck
  // (60-76 us per step seen on a loaded box).
  int graph_mode = 0;
  double graph_threshold_us = 25.0, eager_avg_us = 0.0; int64_t eager_n = 0;
};

static int enqueue_step(rn_host_pairwise* p, Slot& s, const rn_pairwise_args* h, cudaStream_t s_in, cudaStream_t cmp,
                        cudaStream_t s_out, bool events);

extern "C" int rn_host_pairwise_destroy(rn_host_pairwise* p) {
  if (!p) return RN_OK;
  cudaSetDevice(p->device);
  if (p->slots) {
    for (int q = 0; q < p->depth; ++q) {
      Slot& s = p->slots[q];
      if (s.out_done) cudaEventSynchronize(s.out_done);
      if (s.exec) cudaGraphExecDestroy(s.exec);
      if (s.graph) cudaGraphDestroy(s.graph);
      if (s.dev) cudaFree(s.dev);
      if (s.outs_host) cudaFreeHost(s.outs_host);
      if (s.in_done) cudaEventDestroy(s.in_done);
      if (s.cmp_done) cudaEventDestroy(s.cmp_done);
      if (s.out_done) cudaEventDestroy(s.out_done);
      if (s.s_cmp) cudaStreamDestroy(s.s_cmp);
    }
    delete[] p->slots;
  }
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_cmp) cudaStreamDestroy(p->s_cmp);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  return RN_OK;
}

extern "C" int rn_host_pairwise_create(int64_t B_max, int32_t K, int32_t depth, rn_host_pairwise** out) {
  if (!out || B_max <= 0 || B_max > (1ll << 28) || K <= 0 || K > 8 || depth < 1 || depth > 8) return RN_ERR_ARG;
  *out = nullptr;
  rn_host_pairwise* p = new (std::nothrow) rn_host_pairwise;
  if (!p) return RN_ERR_LAUNCH;
  p->B_max = B_max; p->K = K; p->depth = depth;
  if (const char* v = getenv("RN_HOST_STEP_GRAPH")) { if (*v == '1') p->graph_mode = 1; else if (*v == '2') p->graph_mode = 2; }
  if (const char* v = getenv("RN_HOST_STEP_GRAPH_US")) { if (*v) p->graph_threshold_us = atof(v); }
  if (cudaGetDevice(&p->device) != cudaSuccess) { delete p; return RN_ERR_NO_DEVICE; }
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = rn::align_up(o + bytes); return r; };
  const size_t B = (size_t)B_max;
  p->o_keys = take(8 * B * (size_t)K); p->o_logits = take(4 * B); p->o_labels = take(4 * B);
  p->o_rwp = take(4 * B); p->o_rwn = take(4 * B); p->o_ok = take(B);
  p->o_outs = take(32); p->o_dl = take(4 * B); p->o_rp = take(8 * B);
  p->o_lut = take(64 * sizeof(float));                  // the level weight table of RN_LABEL_LUT calls
  p->scratch_bytes = rn_pairwise_scratch_bytes(B_max, K);
  p->o_scr = take(p->scratch_bytes);
  p->slot_bytes = o;
  p->slots = new (std::nothrow) Slot[depth];
  bool ok = p->slots != nullptr;
  ok = ok && cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&p->s_cmp, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) == cudaSuccess;
  for (int q = 0; ok && q < depth; ++q) {
    Slot& s = p->slots[q];
    ok = ok && cudaMalloc(reinterpret_cast<void**>(&s.dev), p->slot_bytes) == cudaSuccess;
    // the slot's scratch arena is persistent (rn_pairwise_args.scratch_persistent): zeroed once, here
    ok = ok && cudaMemset(s.dev + p->o_scr, 0, p->scratch_bytes) == cudaSuccess;
    ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&s.outs_host), 32, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.cmp_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&s.s_cmp, cudaStreamNonBlocking) == cudaSuccess;
  }
  if (!ok) { cudaGetLastError(); rn_host_pairwise_destroy(p); return RN_ERR_LAUNCH; }
  *out = p;
  return RN_OK;
}

extern "C" int64_t rn_host_pairwise_graph_steps(const rn_host_pairwise* p) { return p ? p->graph_steps : 0; }

extern "C" int rn_host_pairwise_wait(rn_host_pairwise* p, int32_t ticket) {
  RN_NVTX_RANGE("rn_host_pairwise_wait");
  if (!p || ticket < 0 || ticket >= p->depth) return RN_ERR_ARG;
  Slot& s = p->slots[ticket];
  if (!s.busy) return RN_OK;
  if (cudaEventSynchronize(s.out_done) != cudaSuccess) return RN_ERR_LAUNCH;
  *s.loss = s.outs_host[0]; *s.n_pair_f32 = s.outs_host[1]; *s.n_pair = *reinterpret_cast<const int64_t*>(s.outs_host + 2);
  s.busy = false;
  return RN_OK;
}

extern "C" int rn_host_pairwise_submit(rn_host_pairwise* p, const rn_pairwise_args* h, int32_t* ticket) {
  RN_NVTX_RANGE("rn_host_pairwise_submit");
  if (!p || !h || !ticket) return RN_ERR_ARG;
  if (h->B <= 0 || h->B > p->B_max || h->K != p->K) return RN_ERR_ARG;
  if (!h->keys || !h->logits || !h->labels || !h->loss || !h->n_pair_f32 || !h->n_pair || !h->dlogits) return RN_ERR_ARG;
  if (h->block_rows || h->out_chunk || h->gather_dst || h->part_count != 1 || h->part_rank != 0) return RN_ERR_UNSUPPORTED;
  const int32_t q = p->next;
  int rc = rn_host_pairwise_wait(p, q);                 // the slot's previous batch has left the device
  if (rc) return rc;
  Slot& s = p->slots[q];
  static const bool one_stream = []() { const char* v = getenv("RN_HOST_ONE_STREAM"); return v && *v == '1'; }();
  const bool step_graph = p->graph_mode == 1 ||
                          (p->graph_mode == 2 && p->eager_n >= 8 && p->eager_avg_us > p->graph_threshold_us);
  const auto t_begin = std::chrono::steady_clock::now();
  // ---- whole-step graph: the slot has a graph for exactly these arguments -> one launch
  const bool same_as_last = s.seen_valid && memcmp(&s.seen, h, sizeof(*h)) == 0;
  s.seen = *h; s.seen_valid = true;
  if (!same_as_last) s.seen_pageable = false;
  if (step_graph && !one_stream && !p->graph_broken) {
    if (s.exec && memcmp(&s.graph_args, h, sizeof(*h)) == 0) {
      if (cudaGraphLaunch(s.exec, s.s_cmp) != cudaSuccess || cudaEventRecord(s.out_done, s.s_cmp) != cudaSuccess) {
        cudaGetLastError(); return RN_ERR_LAUNCH;
      }
      s.loss = h->loss; s.n_pair_f32 = h->n_pair_f32; s.n_pair = h->n_pair;
      s.busy = true; *ticket = q; p->next = (q + 1) % p->depth; ++p->graph_steps;
      return RN_OK;
    }
    if (same_as_last && !s.seen_pageable)
      s.seen_pageable = !(pinned(h->keys) && pinned(h->logits) && pinned(h->labels) && pinned(h->rw_pos) && pinned(h->rw_neg) &&
                          pinned(h->row_ok) && pinned(h->weight_lut) && pinned(h->dlogits) && pinned(h->row_pairs));
    if (same_as_last && !s.seen_pageable) {
      // second submit of the same buffers: capture this step on the slot's stream, instantiate, launch
      if (s.exec) { cudaGraphExecDestroy(s.exec); s.exec = nullptr; }
      if (s.graph) { cudaGraphDestroy(s.graph); s.graph = nullptr; }
      int grc = RN_OK;
      if (cudaStreamBeginCapture(s.s_cmp, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        grc = enqueue_step(p, s, h, s.s_cmp, s.s_cmp, s.s_cmp, false);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s.s_cmp, &g);
        if (e == cudaSuccess && grc == RN_OK && g && cudaGraphInstantiate(&s.exec, g, 0) == cudaSuccess) {
          s.graph = g; s.graph_args = *h;
          if (cudaGraphLaunch(s.exec, s.s_cmp) != cudaSuccess || cudaEventRecord(s.out_done, s.s_cmp) != cudaSuccess) {
            cudaGetLastError(); return RN_ERR_LAUNCH;
          }
          s.loss = h->loss; s.n_pair_f32 = h->n_pair_f32; s.n_pair = h->n_pair;
          s.busy = true; *ticket = q; p->next = (q + 1) % p->depth; ++p->graph_steps;
          return RN_OK;
        }
        if (g) cudaGraphDestroy(g);
        s.exec = nullptr;
      }
      if (getenv("RN_GRAPH_DEBUG")) fprintf(stderr, "[recnow] host front end: whole-step graph failed (%s), eager submits from now on\n", cudaGetErrorString(cudaPeekAtLastError()));
      cudaGetLastError();
      p->graph_broken = true;          // (nothing was executed: the eager path below runs this step, and every later one)
      if (grc != RN_OK && grc != RN_ERR_LAUNCH) return grc;       // (an argument error shows the same way on either path)
    }
  }
  rc = enqueue_step(p, s, h, p->s_in, one_stream ? p->s_cmp : s.s_cmp, p->s_out, true);
  if (rc) return rc;
  {
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
    ++p->eager_n;
    p->eager_avg_us += (us - p->eager_avg_us) / (double)(p->eager_n < 16 ? p->eager_n : 16);
  }
  s.loss = h->loss; s.n_pair_f32 = h->n_pair_f32; s.n_pair = h->n_pair;
  s.busy = true;
  *ticket = q;
  p->next = (q + 1) % p->depth;
  return RN_OK;
}

// One step of slot `s` enqueued on the given streams: copy-in on s_in, kernels on cmp, copy-out on s_out (events between them
// when the streams differ; under a stream capture all three are the capturing stream and `events` is false).
static int enqueue_step(rn_host_pairwise* p, Slot& s, const rn_pairwise_args* h, cudaStream_t s_in, cudaStream_t cmp,
                        cudaStream_t s_out, bool events) {
  const size_t B = (size_t)h->B;
  char* d = s.dev;
  bool ok = true;
  const int32_t q = (int32_t)(&s - p->slots);
  auto h2d = [&](size_t off, const void* src, size_t bytes) {
    ok = ok && cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, s_in) == cudaSuccess;
  };
  // ---- copy-in (key column k of the device block starts B int64 after column k-1, as rn_pairwise_args wants it)
  h2d(p->o_keys, h->keys, 8 * B * (size_t)p->K);
  h2d(p->o_logits, h->logits, 4 * B);
  h2d(p->o_labels, h->labels, 4 * B);
  if (h->rw_pos) h2d(p->o_rwp, h->rw_pos, 4 * B);
  if (h->rw_neg) h2d(p->o_rwn, h->rw_neg, 4 * B);
  if (h->row_ok) h2d(p->o_ok, h->row_ok, B);
  if (h->weight_lut) h2d(p->o_lut, h->weight_lut, 64 * sizeof(float));
  if (events) ok = ok && cudaEventRecord(s.in_done, s_in) == cudaSuccess;
  // ---- kernels
  if (events) ok = ok && cudaStreamWaitEvent(cmp, s.in_done, 0) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return RN_ERR_LAUNCH; }
  rn_pairwise_args a = *h;
  a.keys = reinterpret_cast<const int64_t*>(d + p->o_keys);
  a.logits = reinterpret_cast<const float*>(d + p->o_logits);
  a.labels = reinterpret_cast<const float*>(d + p->o_labels);
  a.rw_pos = h->rw_pos ? reinterpret_cast<const float*>(d + p->o_rwp) : nullptr;
  a.rw_neg = h->rw_neg ? reinterpret_cast<const float*>(d + p->o_rwn) : nullptr;
  a.row_ok = h->row_ok ? reinterpret_cast<const uint8_t*>(d + p->o_ok) : nullptr;
  a.weight_lut = h->weight_lut ? reinterpret_cast<const float*>(d + p->o_lut) : nullptr;
  float* outs = reinterpret_cast<float*>(d + p->o_outs);          // [loss f32][n_pair f32][n_pair i64]
  a.loss = outs; a.n_pair_f32 = outs + 1; a.n_pair = reinterpret_cast<int64_t*>(outs + 2);
  a.dlogits = reinterpret_cast<float*>(d + p->o_dl);
  a.row_pairs = h->row_pairs ? reinterpret_cast<int64_t*>(d + p->o_rp) : nullptr;
  a.scratch_persistent = 1; a.scratch_rows = p->B_max;     // (one layout for every batch size this object accepts)
  rn::set_graph_lane(cmp == p->s_cmp ? 0 : q + 1);
  const int rc = rn_pairwise_fwd_bwd(&a, d + p->o_scr, p->scratch_bytes, cmp);
  rn::set_graph_lane(0);
  if (rc) return rc;
  if (events) ok = ok && cudaEventRecord(s.cmp_done, cmp) == cudaSuccess;
  // ---- copy-out
  if (events) ok = ok && cudaStreamWaitEvent(s_out, s.cmp_done, 0) == cudaSuccess;
  auto d2h = [&](void* dst, size_t off, size_t bytes) {
    ok = ok && cudaMemcpyAsync(dst, d + off, bytes, cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
  };
  d2h(h->dlogits, p->o_dl, 4 * B);
  if (h->row_pairs) d2h(h->row_pairs, p->o_rp, 8 * B);
  d2h(s.outs_host, p->o_outs, 16);                      // (handed to the caller's three pointers by the wait)
  if (events) ok = ok && cudaEventRecord(s.out_done, s_out) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return RN_ERR_LAUNCH; }
  return RN_OK;
}
