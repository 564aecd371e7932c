// Host plumbing of the segmentation kernel: arena initialisation and cooperative launches.
// (The device code lives in segment.cuh and is instantiated by the three consumers.)
#include "common.cuh"

namespace rn {

__global__ void __launch_bounds__(256) k_init(uint4* zero, size_t nzero16, uint4* ones, size_t nones16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(~0u, ~0u, ~0u, ~0u);
  for (size_t k = i; k < nzero16; k += stride) zero[k] = z;
  for (size_t k = i; k < nones16; k += stride) ones[k] = f;
}

cudaError_t seg_init(const Layout& L, void* scratch, cudaStream_t st) {
  char* base = static_cast<char*>(scratch);
  const size_t nz = (L.zero_end - L.zero_begin) / 16, no = (L.ones_end - L.ones_begin) / 16;
  int grid = (int)((nz + no + 255) / 256);
  const int cap = device_sm_count() * 4;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_init<<<grid, 256, 0, st>>>(at<uint4>(base, L.zero_begin), nz, at<uint4>(base, L.ones_begin), no);
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

cudaError_t launch_coop(const void* kernel, int grid, int threads, void** args, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

}  // namespace rn
