// Microbenchmark: MUFU throughput of scalar vs packed-half transcendental instructions on sm_100a.
// Question behind it (attention v5): the ViT attention softmax is MUFU.EX2-bound while both query tiles are in
// their softmax phase (208 ex2 per row, 16 lanes / clk / SM).  Does ex2.approx.f16x2 retire two exponentials per MUFU slot?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bw tools/micro/mufu_bw.cu && /tmp/mufu_bw
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("{.reg .f32 t; mov.b32 t, %1; fma.rn.f32 t, t, t, t; mov.b32 %0, t;}" : "=r"(y) : "r"(x));
  if (OP == 7) asm volatile("{.reg .f32 t; mov.b32 t, %1; rcp.approx.ftz.f32 t, t; mov.b32 %0, t;}" : "=r"(y) : "r"(x));
  return y;
}

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters, long long* cycles) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0x3c003c00u + threadIdx.x + i;  // harmless bit patterns for every format
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int elems) {
  const int blocks = 148 * 4, threads = 256, iters = 4096;
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  bench<OP><<<blocks, threads>>>(out, 16, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 4];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  // 4 blocks x 8 warps per SM, each warp issues iters * 8 instructions
  const double warp_instr_per_sm = 4.0 * 8 * iters * 8;
  printf("%-22s %8.3f ms  %7.2f clk per warp-instruction per SM  -> %6.1f lanes/clk/SM, %6.1f results/clk/SM  (%s)\n", name, ms,
         avg / warp_instr_per_sm, 32.0 * warp_instr_per_sm / avg, 32.0 * elems * warp_instr_per_sm / avg,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  run<5>("tanh.approx.bf16x2", 2);
  run<6>("fma.rn.f32", 1);
  run<7>("rcp.approx.ftz.f32", 1);
  return 0;
}
