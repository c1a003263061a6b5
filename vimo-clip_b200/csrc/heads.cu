// H1 / T1-head: per-clip fused head kernels (fp32 SIMT; these stages are ~0.002 % of the path's FLOPs and
// were launch-bound as separate cast / GEMM / mean / LayerNorm launches).
//
//   student_heads_kernel  (models/student_model.py:33-35, 90-96): for one clip, ResidualMLP
//       distill = emb + alpha * fc2(GELU_erf(fc1(emb)))  on every frame row,
//       pooled = mean_t emb (RAW embeddings),  logits = W2 relu(W1 pooled + b1) + b2.
//   tfam_head_kernel      (TFAM/models/AMO_CLIP.py:170, classifier :84): for one clip,
//       logits = Linear(GELU_erf(Linear(LayerNorm(mean over ALL T rows of x)))).
// One CTA per clip.  Weights are fp32, stored TRANSPOSED ([K, N]) by the host packer so that for a fixed
// k the threads of a warp read consecutive addresses; the clip's activations live in shared memory,
// transposed ([K][16 rows]) so one 16-byte broadcast load feeds four rows.
#include "common.cuh"
#include "vimoclip_b200.h"

namespace {

constexpr int RT = 16;  // frame rows per tile

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// acc[c][r] += sum_k xt[k][r] * wt[k][n0 + c*blockDim.x], r < RT, for NC columns per thread
template <int NC>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ xt, const float* __restrict__ wt, int K,
                                          int N, int n0, float (&acc)[NC][RT]) {
  // weights of KU k-steps are requested before the first of them is used: the loop was bound by the L2 latency of one weight
  // load per k-step (0.34 ms per clip CTA at ~10 % of the FMA rate)
  constexpr int KU = 8;
  int k = 0;
  for (; k + KU <= K; k += KU) {
    float w[KU][NC];
#pragma unroll
    for (int u = 0; u < KU; ++u)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int n = n0 + c * blockDim.x;
        w[u][c] = n < N ? __ldg(wt + (size_t)(k + u) * N + n) : 0.f;
      }
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const float4* xr = reinterpret_cast<const float4*>(xt + (k + u) * RT);
#pragma unroll
      for (int q = 0; q < RT / 4; ++q) {
        const float4 x4 = xr[q];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[c][4 * q + 0] = fmaf(x4.x, w[u][c], acc[c][4 * q + 0]);
          acc[c][4 * q + 1] = fmaf(x4.y, w[u][c], acc[c][4 * q + 1]);
          acc[c][4 * q + 2] = fmaf(x4.z, w[u][c], acc[c][4 * q + 2]);
          acc[c][4 * q + 3] = fmaf(x4.w, w[u][c], acc[c][4 * q + 3]);
        }
      }
    }
  }
  for (; k < K; ++k) {
    float w[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int n = n0 + c * blockDim.x;
      w[c] = n < N ? __ldg(wt + (size_t)k * N + n) : 0.f;
    }
    const float4* xr = reinterpret_cast<const float4*>(xt + k * RT);
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 x4 = xr[q];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        acc[c][4 * q + 0] = fmaf(x4.x, w[c], acc[c][4 * q + 0]);
        acc[c][4 * q + 1] = fmaf(x4.y, w[c], acc[c][4 * q + 1]);
        acc[c][4 * q + 2] = fmaf(x4.z, w[c], acc[c][4 * q + 2]);
        acc[c][4 * q + 3] = fmaf(x4.w, w[c], acc[c][4 * q + 3]);
      }
    }
  }
}

// D (embedding width) == 2 * blockDim.x columns per thread pair; H = hidden of the classifier.
__global__ void __launch_bounds__(256)
student_heads_kernel(const float* __restrict__ emb, const float* __restrict__ w1t, const float* __restrict__ b1,
                     const float* __restrict__ w2t, const float* __restrict__ b2, float alpha,
                     const float* __restrict__ wc1t, const float* __restrict__ bc1,
                     const float* __restrict__ wc2t, const float* __restrict__ bc2, float* __restrict__ distill,
                     float* __restrict__ logits, int T, int D, int H, int C) {
  extern __shared__ __align__(16) float sm[];
  float* xt = sm;            // [D][RT] current tile of embeddings, transposed
  float* ht = sm + D * RT;   // [D][RT] GELU(fc1) of the tile, transposed
  float* pooled = ht + D * RT;  // [D]
  float* hid = pooled + D;      // [H]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* e = emb + (size_t)b * T * D;
  for (int i = tid; i < D; i += blockDim.x) pooled[i] = 0.f;
  for (int t0 = 0; t0 < T; t0 += RT) {
    __syncthreads();
    for (int i = tid; i < RT * D; i += blockDim.x) {
      const int r = i / D, k = i - r * D;  // coalesced read of row r
      const float v = (t0 + r < T) ? e[(size_t)(t0 + r) * D + k] : 0.f;
      xt[k * RT + r] = v;
    }
    __syncthreads();
    // pooled sum of the RAW embeddings (student_model.py:93)
    for (int k = tid; k < D; k += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < RT; ++r) s += xt[k * RT + r];
      pooled[k] += s;
    }
    // h = GELU(fc1(x))
    for (int n0 = tid; n0 < D; n0 += 2 * blockDim.x) {
      float acc[2][RT] = {};
      tile_gemm<2>(xt, w1t, D, D, n0, acc);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n = n0 + c * blockDim.x;
        if (n < D) {
          const float bb = b1[n];
#pragma unroll
          for (int r = 0; r < RT; ++r) ht[n * RT + r] = gelu_erf(acc[c][r] + bb);
        }
      }
    }
    __syncthreads();
    // distill = x + alpha * (fc2(h) + b2)
    for (int n0 = tid; n0 < D; n0 += 2 * blockDim.x) {
      float acc[2][RT] = {};
      tile_gemm<2>(ht, w2t, D, D, n0, acc);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n = n0 + c * blockDim.x;
        if (n < D) {
          const float bb = b2[n];
#pragma unroll
          for (int r = 0; r < RT; ++r)
            if (t0 + r < T)
              distill[((size_t)b * T + t0 + r) * D + n] = fmaf(alpha, acc[c][r] + bb, xt[n * RT + r]);
        }
      }
    }
  }
  __syncthreads();
  const float inv_t = 1.0f / (float)T;
  for (int k = tid; k < D; k += blockDim.x) pooled[k] *= inv_t;
  __syncthreads();
  for (int n = tid; n < H; n += blockDim.x) {
    float s = bc1[n];
    for (int k = 0; k < D; ++k) s = fmaf(pooled[k], __ldg(wc1t + (size_t)k * H + n), s);
    hid[n] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int n = tid; n < C; n += blockDim.x) {
    float s = bc2[n];
    for (int k = 0; k < H; ++k) s = fmaf(hid[k], __ldg(wc2t + (size_t)k * C + n), s);
    logits[(size_t)b * C + n] = s;
  }
}

__global__ void __launch_bounds__(256)
tfam_head_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                 float eps, const float* __restrict__ w1t, const float* __restrict__ b1,
                 const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ logits, int T,
                 int D, int H, int C) {
  extern __shared__ __align__(16) float sm[];
  float* pooled = sm;      // [D]
  float* hid = sm + D;     // [H]
  float* red = hid + H;    // [32]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xb = x + (size_t)b * T * D;
  // mean over ALL T rows, padded ones included (AMO_CLIP.py:170)
  float lsum = 0.f;
  for (int k = tid; k < D; k += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += xb[(size_t)t * D + k];
    s /= (float)T;
    pooled[k] = s;
    lsum += s;
  }
  // LayerNorm statistics (two-pass, fp32)
  lsum = vmc::warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float mean = tot / (float)D;
  __syncthreads();
  float lq = 0.f;
  for (int k = tid; k < D; k += blockDim.x) {
    const float dlt = pooled[k] - mean;
    lq += dlt * dlt;
  }
  lq = vmc::warp_sum(lq);
  if (lane == 0) red[warp] = lq;
  __syncthreads();
  float var = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) var += red[i];
  const float rstd = rsqrtf(var / (float)D + eps);
  __syncthreads();
  for (int k = tid; k < D; k += blockDim.x) pooled[k] = (pooled[k] - mean) * rstd * ln_g[k] + ln_b[k];
  __syncthreads();
  for (int n = tid; n < H; n += blockDim.x) {
    float s = b1[n];
    for (int k = 0; k < D; ++k) s = fmaf(pooled[k], __ldg(w1t + (size_t)k * H + n), s);
    hid[n] = gelu_erf(s);
  }
  __syncthreads();
  for (int n = tid; n < C; n += blockDim.x) {
    float s = b2[n];
    for (int k = 0; k < H; ++k) s = fmaf(hid[k], __ldg(w2t + (size_t)k * C + n), s);
    logits[(size_t)b * C + n] = s;
  }
}

}  // namespace

extern "C" {

int vmc_student_heads(const float* emb, const float* w_fc1_t, const float* b_fc1, const float* w_fc2_t,
                      const float* b_fc2, float alpha, const float* w_c1_t, const float* b_c1,
                      const float* w_c2_t, const float* b_c2, float* distill, float* logits, int B, int T,
                      int D, int H, int C, void* stream) {
  VMC_CHECK_ARG(emb && w_fc1_t && b_fc1 && w_fc2_t && b_fc2 && w_c1_t && b_c1 && w_c2_t && b_c2 && distill && logits,
                VMC_ERR_ARG, "vmc_student_heads: null pointer");
  VMC_CHECK_ARG(B > 0 && T > 0 && D > 0 && H > 0 && C > 0 && D <= 1024 && H <= 1024, VMC_ERR_SHAPE,
                "vmc_student_heads: bad shape B=%d T=%d D=%d H=%d C=%d", B, T, D, H, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = ((size_t)2 * D * RT + D + H) * sizeof(float);
  VMC_CUDA(cudaFuncSetAttribute(student_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    VmcProfScope prof(VMC_K_OTHER, st, 4.0 * B * T * (double)D * D + 2.0 * B * ((double)D * H + (double)H * C), 0.0);
    student_heads_kernel<<<B, 256, smem, st>>>(emb, w_fc1_t, b_fc1, w_fc2_t, b_fc2, alpha, w_c1_t, b_c1, w_c2_t,
                                               b_c2, distill, logits, T, D, H, C);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_tfam_head(const float* x, const float* ln_g, const float* ln_b, float eps, const float* w1_t,
                  const float* b1, const float* w2_t, const float* b2, float* logits, int B, int T, int D, int H,
                  int C, void* stream) {
  VMC_CHECK_ARG(x && ln_g && ln_b && w1_t && b1 && w2_t && b2 && logits, VMC_ERR_ARG, "vmc_tfam_head: null pointer");
  VMC_CHECK_ARG(B > 0 && T > 0 && D > 0 && H > 0 && C > 0 && D <= 4096 && H <= 4096, VMC_ERR_SHAPE,
                "vmc_tfam_head: bad shape B=%d T=%d D=%d H=%d C=%d", B, T, D, H, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = ((size_t)D + H + 32) * sizeof(float);
  {
    VmcProfScope prof(VMC_K_OTHER, st, 2.0 * B * ((double)D * H + (double)H * C), 4.0 * B * T * (double)D);
    tfam_head_kernel<<<B, 256, smem, st>>>(x, ln_g, ln_b, eps, w1_t, b1, w2_t, b2, logits, T, D, H, C);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // extern "C"
