// Dev micro-benchmark (not part of the product library): cycles per fast tile as a function of resident warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo scripts/dev/tile_bench.cu -o scripts/dev/tile_bench
#include <cstdio>
#include <vector>
#include "../../rec_now_b200/csrc/pair_tiles.cuh"
using namespace rn;

template <int VAR>
__global__ void __launch_bounds__(VAR == 2 ? 512 : 1024) k_tb(int iters, float* out, unsigned long long* cyc) {
  const u32 ln = threadIdx.x & 31;
  float si0 = 0.01f * ln, si1 = -0.02f * ln, sjm = 0.03f * ln - 0.5f;
  float li0 = 0, li1 = 0, gi0 = 0, gi1 = 0, accj = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (VAR == 0) tile_fast<true>(si0, si1, 1.1f, 0.9f, sjm, 1.4427f, li0, li1, gi0, gi1, accj);
    if (VAR == 1) tile_fast<false>(si0, si1, 1.1f, 0.9f, sjm, 1.4427f, li0, li1, gi0, gi1, accj);
    if (VAR == 2) tile_fast<true, false, 8>(si0, si1, 1.1f, 0.9f, sjm, 1.4427f, li0, li1, gi0, gi1, accj);
    if (VAR == 3) tile_prod<false>(mufu_ex2(-si0), mufu_ex2(-si1), 1.1f, 0.9f, mufu_ex2(sjm), li0, li1, gi0, gi1, accj);
    if (VAR == 4) tile_prod<true>(mufu_ex2(-si0), mufu_ex2(-si1), 1.1f, 0.9f, mufu_ex2(sjm), li0, li1, gi0, gi1, accj, 0, 32);
    sjm += 1e-3f;
  }
  const long long t1 = clock64();
  if (ln == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = (unsigned long long)(t1 - t0);
  out[blockIdx.x * blockDim.x + threadIdx.x] = li0 + li1 + gi0 + gi1 + accj;
}

template <int VAR>
void run(const char* name) {
  float* out; unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  const int iters = 64;
  for (int nw : {1, 2, 4, 8, 12, 16, 24, 32}) {
    if (VAR == 2 && nw > 16) continue;
    for (int rep = 0; rep < 3; ++rep) k_tb<VAR><<<148, 32 * nw>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h(148 * nw);
    cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (auto v : h) s += (double)v;
    const double per_tile = s / h.size() / iters;
    // MUFU-bound time per tile per SMSP: 128 (tile_fast) or 64 (tile_prod) MUFU warp-instr x 8 cycles; nw/4 warps share an SMSP
    const double mufu_cyc = VAR >= 3 ? 512.0 : 1024.0;
    printf("%s nw=%2d  %8.0f cyc/tile/warp   mufu util %.2f  (%s)\n", name, nw, per_tile, (nw / 4.0 < 1 ? 1 : nw / 4.0) * mufu_cyc / per_tile,
           cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("fast<HASW>");
  run<1>("fast<noW> ");
  run<3>("prod          ");
  run<4>("prod<looped>  ");
  return 0;
}
