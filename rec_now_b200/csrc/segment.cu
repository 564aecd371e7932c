// Host plumbing of the segmentation kernel: arena initialisation and cooperative launches.
// (The device code lives in segment.cuh and is instantiated by the three consumers.)
#include "common.cuh"

namespace rn {

__global__ void __launch_bounds__(256) k_init(uint4* zero, size_t nzero16, uint4* ones, size_t nones16,
                                              const float* labels, const uint8_t* row_ok, u32 B, u32* labpart) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(~0u, ~0u, ~0u, ~0u);
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

cudaError_t seg_init(const Layout& L, void* scratch, cudaStream_t st, const float* labels, const uint8_t* row_ok, int* ncta) {
  char* base = static_cast<char*>(scratch);
  const size_t nz = (L.zero_end - L.zero_begin) / 16, no = (L.ones_end - L.ones_begin) / 16;
  int grid = (int)((nz + no + 255) / 256);
  int cap = device_sm_count() * 4;
  if (cap > kInitMaxCtas) cap = kInitMaxCtas;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (ncta) *ncta = grid;
  k_init<<<grid, 256, 0, st>>>(at<uint4>(base, L.zero_begin), nz, at<uint4>(base, L.ones_begin), no,
                               labels, row_ok, (u32)L.B, at<u32>(base, L.labpart));
  return cudaGetLastError();
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
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

}  // namespace rn
