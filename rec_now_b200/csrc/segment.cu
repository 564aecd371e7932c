// Host plumbing of the segmentation kernel: arena initialisation and cooperative launches.
// (The device code lives in segment.cuh and is instantiated by the three consumers.)
#include <stdio.h>
#include "common.cuh"

namespace rn {

__global__ void __launch_bounds__(256) k_init(uint4* zero, size_t nzero16, uint4* ones, size_t nones16,
                                              const float* labels, const uint8_t* row_ok, u32 B, u32* labpart,
                                              GatherArgs G) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  grid_dep_launch();
  const uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(~0u, ~0u, ~0u, ~0u);
  if (G.world) {
    // global mode: the all-gather -- every rank's packed row block, read over NVLink from its peer mapping (4 loads
    // in flight per thread: the round trip is a few microseconds)
    const size_t tot = (size_t)G.world * G.n16;
    for (size_t k0 = i; k0 < tot; k0 += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const size_t k = k0 + q * stride;
        if (k < tot) { const u32 r = (u32)(k / G.n16); v[q] = G.src[r][k - (size_t)r * G.n16]; }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { const size_t k = k0 + q * stride; if (k < tot) G.dst[k] = v[q]; }
    }
  }
  for (size_t k = i; k < nzero16; k += stride) zero[k] = z;
  for (size_t k = i; k < nones16; k += stride) ones[k] = f;
  if (labels) {
    // OR / OR-of-complement of the order-preserving label encoding over the rows that can pair: the varying bit
    // range of the labels (make_plan), one partial per CTA
    __shared__ u32 red[2][8];
    u32 vor = 0, vnor = 0;
    for (size_t k = i; k < B; k += stride) {
      const float y = labels[k];
      if ((row_ok ? row_ok[k] != 0 : true) && !(y != y)) { const u32 e = enc_label(y); vor |= e; vnor |= ~e; }
    }
    vor = __reduce_or_sync(0xFFFFFFFFu, vor); vnor = __reduce_or_sync(0xFFFFFFFFu, vnor);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = vor; red[1][threadIdx.x >> 5] = vnor; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < 8; ++q) { vor |= red[0][q]; vnor |= red[1][q]; }
      labpart[2 * blockIdx.x] = vor; labpart[2 * blockIdx.x + 1] = vnor;
    }
  }
}

// ---- launch plumbing --------------------------------------------------------------------------------------------
// A call of the pairwise path is three dependent launches (k_init, k_seg, k_pair) of a few tens of microseconds
// each, so the gaps between them count.  When a GraphCall is active on this thread the launches below do not go to
// the stream: they set the parameters of the three kernel nodes of a cached, instantiated CUDA graph, which the
// caller then launches once (smaller inter-kernel gaps, one driver call instead of three).
// (f[0] == nullptr: a call without k_init -- the counting path on a persistent arena -- is a graph of two kernel nodes)
// (lane: an executable graph runs one launch at a time -- a caller that keeps several calls in flight on different streams,
// as the host-buffer front end does with its slots, gives every stream a lane of its own and so an executable of its own)
struct GraphSlot { int dev; int lane; const void* f[3]; int nf; bool timed; cudaGraph_t graph; cudaGraphExec_t exec; cudaGraphNode_t node[3]; };
static thread_local int tl_lane = 0;
void set_graph_lane(int lane) { tl_lane = lane; }
static thread_local GraphSlot tl_slots[96];
static thread_local int tl_nslots = 0;
static thread_local GraphSlot* tl_update = nullptr;      // slot whose nodes receive the launches of this thread
static thread_local int tl_next = 0;
static thread_local bool tl_capturing = false;           // launches go to the private capture stream (graph build)
static thread_local bool tl_capture_timed = false;       // ... of a timed graph (event-record nodes between the kernels: no programmatic edges)
static thread_local bool tl_graph_broken = false;        // a graph API call failed: direct launches from now on
static thread_local cudaStream_t tl_cap_stream = nullptr;
static thread_local long long tl_graph_launches = 0;   // calls of this thread that went out as one graph launch
long long graph_launch_count() { return tl_graph_broken ? -tl_graph_launches - 1 : tl_graph_launches; }

static cudaError_t emit(const void* f, dim3 grid, dim3 block, size_t smem, void** args, bool coop, cudaStream_t st) {
  if (tl_update) {
    const int i = tl_next++;
    if (i >= tl_update->nf || tl_update->f[i] != f) { if (getenv("RN_GRAPH_DEBUG")) fprintf(stderr, "[recnow] graph node %d: unexpected kernel\n", i); return cudaErrorInvalidValue; }
    cudaKernelNodeParams p{};
    p.func = const_cast<void*>(f); p.gridDim = grid; p.blockDim = block; p.sharedMemBytes = (unsigned)smem;
    p.kernelParams = args; p.extra = nullptr;
    const cudaError_t e = cudaGraphExecKernelNodeSetParams(tl_update->exec, tl_update->node[i], &p);
    if (e != cudaSuccess && getenv("RN_GRAPH_DEBUG")) fprintf(stderr, "[recnow] graph node %d update: %s\n", i, cudaGetErrorString(e));
    return e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (coop) {
    attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na;
    // programmatic dependent launch: the launch is processed while the previous kernel of the stream still runs; the
    // kernel itself waits for that kernel's completion and memory flush with griddepcontrol.wait (grid_dep_wait)
    static const char* nopdl = getenv("RN_NO_PDL");
    static const char* graphpdl = getenv("RN_GRAPH_PDL");       // (RN_GRAPH_PDL=0: no programmatic edges inside the captured graph)
    if ((!tl_capturing || (!tl_capture_timed && !(graphpdl && *graphpdl == '0'))) && !(nopdl && *nopdl == '1')) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1; ++na;
    }
  }
  cfg.attrs = attr; cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelExC(&cfg, f, args);
}

const void* seg_init_func() { return (const void*)k_init; }

cudaEvent_t* timed_events() {
  static thread_local cudaEvent_t ev[2] = {nullptr, nullptr};
  static thread_local bool tried = false;
  if (!tried) {
    tried = true;
    if (cudaEventCreate(&ev[0]) != cudaSuccess || cudaEventCreate(&ev[1]) != cudaSuccess) { ev[0] = ev[1] = nullptr; cudaGetLastError(); }
  }
  return ev[0] && ev[1] ? ev : nullptr;
}

GraphCall::GraphCall(const void* f_init, const void* f_seg, const void* f_pair, cudaStream_t user_stream, bool allow,
                     bool timed)
    : st(user_stream), run_stream(user_stream) {
  static const char* off = getenv("RN_GRAPH");
  if (!allow || tl_graph_broken || (off && *off == '0')) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(user_stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  const void* fl[3]; int nf = 0;
  if (f_init) fl[nf++] = f_init;
  fl[nf++] = f_seg; fl[nf++] = f_pair;
  if (nf == 2) fl[2] = nullptr;
  for (int i = 0; i < tl_nslots; ++i) {
    GraphSlot& g = tl_slots[i];
    if (g.dev == dev && g.lane == tl_lane && g.nf == nf && g.f[0] == fl[0] && g.f[1] == fl[1] && g.f[2] == fl[2] && g.timed == timed) {
      slot = &g; tl_update = slot; tl_next = 0; mode = 1;
      return;
    }
  }
  if (tl_nslots >= (int)(sizeof(tl_slots) / sizeof(tl_slots[0]))) return;
  if (!tl_cap_stream && cudaStreamCreateWithFlags(&tl_cap_stream, cudaStreamNonBlocking) != cudaSuccess) { tl_graph_broken = true; return; }
  if (cudaStreamBeginCapture(tl_cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { tl_graph_broken = true; cudaGetLastError(); return; }
  slot = &tl_slots[tl_nslots];
  slot->dev = dev; slot->lane = tl_lane; slot->f[0] = fl[0]; slot->f[1] = fl[1]; slot->f[2] = fl[2]; slot->nf = nf; slot->timed = timed;
  run_stream = tl_cap_stream; tl_capturing = true; tl_capture_timed = timed; mode = 2;
}

cudaError_t GraphCall::finish(bool ok) {
  if (mode == 0) return ok ? cudaSuccess : cudaErrorUnknown;
  if (mode == 1) {
    const bool all = tl_next == slot->nf;
    tl_update = nullptr;
    if (!ok || !all) { tl_graph_broken = true; return cudaErrorUnknown; }
    ++tl_graph_launches;
    return cudaGraphLaunch(slot->exec, st);
  }
  // mode 2: end the capture, instantiate, find the three kernel nodes, launch
  tl_capturing = false;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(tl_cap_stream, &graph);
  if (e != cudaSuccess || !ok || !graph) { tl_graph_broken = true; if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return ok ? e : cudaErrorUnknown; }
  cudaGraphNode_t nodes[8]; size_t n = 8;
  bool found[3] = {false, false, false};
  if (cudaGraphGetNodes(graph, nodes, &n) == cudaSuccess && n == (size_t)slot->nf + (slot->timed ? 2u : 0u)) {
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType ty;
      if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
      cudaKernelNodeParams p{};
      if (cudaGraphKernelNodeGetParams(nodes[i], &p) != cudaSuccess) continue;
      for (int k = 0; k < slot->nf; ++k) if (!found[k] && p.func == slot->f[k]) { slot->node[k] = nodes[i]; found[k] = true; break; }
    }
  }
  if (slot->nf == 2) found[2] = true;
  if (!(found[0] && found[1] && found[2]) || cudaGraphInstantiate(&slot->exec, graph, 0) != cudaSuccess) {
    // (still run this call: the captured work was not executed)
    tl_graph_broken = true; cudaGetLastError();
    cudaGraphExec_t once = nullptr;
    e = cudaGraphInstantiate(&once, graph, 0);
    if (e == cudaSuccess) { e = cudaGraphLaunch(once, st); cudaGraphExecDestroy(once); }
    cudaGraphDestroy(graph);
    return e;
  }
  // (the graph stays alive: its node handles address the nodes of the executable graph in later updates)
  slot->graph = graph;
  ++tl_nslots; ++tl_graph_launches;
  return cudaGraphLaunch(slot->exec, st);
}

cudaError_t seg_init(const Layout& L, void* scratch, cudaStream_t st, const float* labels, const uint8_t* row_ok, int* ncta,
                     const GatherArgs* gather, bool gather_only) {
  char* base = static_cast<char*>(scratch);
  // (gather_only: the counting path on a persistent arena needs nothing initialised -- the launch is only the peer gather)
  const size_t nz = gather_only ? 0 : (L.zero_end - L.zero_begin) / 16, no = gather_only ? 0 : (L.ones_end - L.ones_begin) / 16;
  GatherArgs G{};
  if (gather) G = *gather;
  int grid = (int)((nz + no + (size_t)G.world * G.n16 + 255) / 256);
  int cap = device_sm_count() * 4;
  if (cap > kInitMaxCtas) cap = kInitMaxCtas;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (ncta) *ncta = grid;
  uint4* zp = at<uint4>(base, L.zero_begin); uint4* op = at<uint4>(base, L.ones_begin);
  size_t nzv = nz, nov = no; u32 Bv = (u32)L.B; u32* lp = at<u32>(base, L.labpart);
  void* args[] = {&zp, &nzv, &op, &nov, &labels, &row_ok, &Bv, &lp, &G};
  return emit((const void*)k_init, dim3((unsigned)grid), dim3(256), 0, args, false, st);
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

cudaError_t launch_coop(const void* kernel, int grid, int threads, void** args, cudaStream_t st, size_t smem) {
  return emit(kernel, dim3((unsigned)grid), dim3((unsigned)threads), smem, args, true, st);
}

}  // namespace rn
