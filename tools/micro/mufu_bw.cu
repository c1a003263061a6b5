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

// mma.sync.m16n8k16 (bf16 -> fp32) issue rate: NACC independent accumulator chains per warp, 8 warps per block, 4 blocks per SM
template <int NACC>
__global__ void __launch_bounds__(256) hmma_bench(float* out, int iters, long long* cycles) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c01u, 0x3c003c02u};
  uint32_t b[2] = {0x3c003c00u, 0x3c003c03u + threadIdx.x};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NACC>
void run_hmma(int blocks_per_sm) {
  const int blocks = 148 * blocks_per_sm, threads = 256, iters = 2048;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  hmma_bench<NACC><<<blocks, threads>>>(out, 16, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  hmma_bench<NACC><<<blocks, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 16 * 8 * 16 * (double)NACC * iters * 8.0 * blocks;
  long long h[148 * 4];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  printf("mma.sync m16n8k16 bf16: %d chains/warp, %d warps/SM: %8.3f ms  %7.1f TFLOP/s  %.2f clk per HMMA per SM (%s)\n", NACC,
         8 * blocks_per_sm, ms, flops / ms / 1e9, avg / ((double)NACC * iters * 8 * blocks_per_sm), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run_hmma<1>(1);
  run_hmma<4>(1);
  run_hmma<8>(1);
  run_hmma<8>(2);
  run_hmma<8>(4);
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
