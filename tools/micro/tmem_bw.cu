// Microbenchmark: tcgen05.ld bandwidth per SM as a function of the number of warps issuing it.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vimo-clip_b200/csrc -I../../include tmem_bw.cu -o tmem_bw
#include <cstdio>
#include "common.cuh"
void vmc_set_error(const char*, ...) {}
using namespace vmc;

template <int MODE>  // 0: ld x32, 1: ld x16, 2: st x16
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tb + (i & 7) * 32, r);
      tmem_ld_wait();
      acc += r[0] + r[31];
    } else if (MODE == 1) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tb + (i & 15) * 16, r);
      tmem_ld_wait();
      acc += r[0] + r[15];
    } else {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = i + j;
      tmem_st_32x32b_x16(tb + (i & 15) * 16, r);
      tmem_st_wait();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0xdeadbeef) out[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      if (mode == 0) k<0><<<148, warps * 32>>>(d, iters);
      if (mode == 1) k<1><<<148, warps * 32>>>(d, iters);
      if (mode == 2) k<2><<<148, warps * 32>>>(d, iters);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * (mode == 0 ? 4096.0 : 2048.0);
      printf("%s warps=%2d: %lld cycles, %.1f B/cycle/SM, %.1f cycles per instr per warp (%s)\n",
             mode == 0 ? "ld.x32" : mode == 1 ? "ld.x16" : "st.x16", warps, h, bytes / h, (double)h / iters, cudaGetErrorString(e));
    }
  return 0;
}
