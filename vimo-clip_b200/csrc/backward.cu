// Backward-pass kernels of the TFAM block (the TFAM training step,
// TFAM/train_and_eval.py:66-101 -> loss.backward() through TFAM/models/AMO_CLIP.py:37-51,170).
//
// The training step is GEMM-shaped like the forward: every nn.Linear backward is two more tcgen05 GEMMs
// (dX = dY W, dW = dY^T X) on the split-bf16 operands, fed by the transposing cast below.  The remaining
// pieces are small fp32 SIMT kernels: LayerNorm backward (row statistics recomputed from the saved input),
// masked attention backward (one CTA per (clip, head), probabilities recomputed in shared memory), column
// sums for the bias / LayerNorm parameter gradients and a few element-wise ops (ReLU / GELU backward,
// dropout, add).  TFAM is 0.08 % of the path's FLOPs; these kernels are written for clarity, not peak.
#include "common.cuh"
#include "vimoclip_b200.h"

int vmc_get_option(int option);

namespace {

using namespace vmc;

constexpr int HD = 64;

// fp32 [R, C] -> bf16 [C, 3R] "split" operand of the TRANSPOSE: hi = bf16(x), lo = bf16(x - hi);
// form 0 (activation side) = [hi | lo | hi], form 1 (weight side) = [hi | hi | lo].
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                       long long ldy, int R, int C, int form) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? x[(size_t)r * ldx + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) {
      const float v = tile[tx][i];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* row = y + (size_t)c * ldy;
      row[r] = hi;
      row[(size_t)R + r] = form == 0 ? lo : hi;
      row[(size_t)2 * R + r] = form == 0 ? hi : lo;
    }
  }
}

// plain transposing cast: x [R, C] (fp32 or bf16) -> y bf16 [C, R] (ldy >= R); the operands of the bf16 dW GEMMs
// of the student's backward (dW = dY^T X: both operands are transposes of row-major activations)
template <typename T>
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const T* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? static_cast<float>(x[(size_t)r * ldx + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) y[(size_t)c * ldy + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}

__global__ void __launch_bounds__(256)
cast_f32_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ y, long long ldy, int rows, int d) {
  const size_t total = (size_t)rows * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / d;
    const int c = (int)(i - r * d);
    y[r * ldy + c] = __bfloat162float(x[r * ldx + c]);
  }
}

// out[c] (+)= sum_r x[r, c] * (y ? y[r, c] : 1): deterministic (one block owns 32 columns, fixed order)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y, long long ldy,
              float* __restrict__ out, int R, int C, int accumulate) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < C) {
    for (int r = ty; r < R; r += 8) {
      const float a = x[(size_t)r * ldx + c];
      s += y ? a * y[(size_t)r * ldy + c] : a;
    }
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// two-stage variant for tall matrices: grid (C / 32, slices); block (x, s) sums rows [s * rows_per, (s + 1) * rows_per) into
// part[s][c]; colsum_finish_kernel adds the slices in index order, so the result is still deterministic
__global__ void __launch_bounds__(256)
colsum_slice_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y, long long ldy,
                    float* __restrict__ part, int R, int C, int rows_per) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(R, r0 + rows_per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains: loads of four rows in flight per thread
  if (c < C) {
    int r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      const float a0 = x[(size_t)r * ldx + c], a1 = x[(size_t)(r + 8) * ldx + c];
      const float a2 = x[(size_t)(r + 16) * ldx + c], a3 = x[(size_t)(r + 24) * ldx + c];
      if (y) {
        s0 = fmaf(a0, y[(size_t)r * ldy + c], s0);
        s1 = fmaf(a1, y[(size_t)(r + 8) * ldy + c], s1);
        s2 = fmaf(a2, y[(size_t)(r + 16) * ldy + c], s2);
        s3 = fmaf(a3, y[(size_t)(r + 24) * ldy + c], s3);
      } else {
        s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      }
    }
    for (; r < r1; r += 8) s0 += y ? x[(size_t)r * ldx + c] * y[(size_t)r * ldy + c] : x[(size_t)r * ldx + c];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    part[(size_t)blockIdx.y * C + c] = t;
  }
}
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ part, float* __restrict__ out, int slices, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int s = 0; s < slices; ++s) t += part[(size_t)s * C + c];
  out[c] = accumulate ? out[c] + t : t;
}

// Fused pass of the bf16 Linear backward: v = x (or x * QuickGELU'(aux) when the layer behind is c_fc), y16 = bf16(v) = the operand
// of the dX / dW GEMMs, part[s][c] = column sums of the UNROUNDED v (= the bias gradient, finished by colsum_finish_kernel).
// One pass over the fp32 gradient instead of three (element-wise backward, cast, column sum).  grid (C / 128, slices); a thread owns
// four adjacent columns (16-byte loads, 8-byte bf16 stores: a warp row = 512 B in, 256 B out) and every 8th row of the slice.
template <bool QGELU>
__global__ void __launch_bounds__(256)
cast_colsum_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ aux, long long lda,
                   __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ part, int R, int C, int rows_per) {
  __shared__ float red[8][132];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + 4 * tx;  // four adjacent columns per thread: 16-byte loads, 8-byte bf16 stores
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(R, r0 + rows_per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    auto dq = [](float g, float u) {  // g * d/du [u sigmoid(1.702 u)]
      const float sg = 1.0f / (1.0f + __expf(-1.702f * u));
      return g * sg * (1.0f + 1.702f * u * (1.0f - sg));
    };
    auto one = [&](int r) {
      float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + c);
      if (QGELU) {
        const float4 u = *reinterpret_cast<const float4*>(aux + (size_t)r * lda + c);
        v.x = dq(v.x, u.x); v.y = dq(v.y, u.y); v.z = dq(v.z, u.z); v.w = dq(v.w, u.w);
      }
      *reinterpret_cast<uint2*>(y + (size_t)r * ldy + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    };
    int r = r0 + ty;
#pragma unroll 4
    for (; r < r1; r += 8) one(r);
  }
  red[ty][4 * tx] = s0; red[ty][4 * tx + 1] = s1; red[ty][4 * tx + 2] = s2; red[ty][4 * tx + 3] = s3;
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + (int)threadIdx.x < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    part[(size_t)blockIdx.y * C + blockIdx.x * 128 + threadIdx.x] = t;
  }
}

// LayerNorm backward, one warp per row.  z = the LayerNorm INPUT (saved by the forward), statistics
// recomputed.  dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;  xhat_out = xhat (for
// dgamma = colsum(dy * xhat), dbeta = colsum(dy)).
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ gamma, float eps,
                     const float* __restrict__ dy, long long lddy, float* __restrict__ dz, long long lddz,
                     float* __restrict__ xhat_out, long long ldxh, int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* zr = z + (size_t)row * ldz;
  const float* dyr = dy + (size_t)row * lddy;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) s += zr[j];
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float a = zr[j] - mean;
    q += a * a;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float xh = (zr[j] - mean) * rstd;
    const float g = dyr[j] * gamma[j];
    m1 += g;
    m2 += g * xh;
  }
  m1 = warp_sum(m1) / (float)d;
  m2 = warp_sum(m2) / (float)d;
  for (int j = lane; j < d; j += 32) {
    const float xh = (zr[j] - mean) * rstd;
    const float g = dyr[j] * gamma[j];
    dz[(size_t)row * lddz + j] = rstd * (g - m1 - xh * m2);
    if (xhat_out) xhat_out[(size_t)row * ldxh + j] = xh;
  }
}

// Fused LayerNorm backward for the residual stream (d = 32 NV <= 1024): dz (+ the gradient arriving over the residual branch),
// and per-block partial sums of dgamma = sum_r dy * xhat and dbeta = sum_r dy, accumulated in registers while a persistent
// block walks its rows -- xhat is never written, dy is not re-read by two column-sum passes, the residual add is not a pass of
// its own: 16 bytes per element instead of 40.  part[block][0..d) = dgamma, [d..2d) = dbeta; colsum_finish_kernel adds the
// blocks in index order (deterministic).
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_fused_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ gamma, float eps,
                           const float* __restrict__ dy, long long lddy, const float* __restrict__ add, long long ldadd,
                           float* __restrict__ dz, long long lddz, float* __restrict__ part, int rows) {
  constexpr int d = 32 * NV;
  extern __shared__ float s_acc[];  // [8 warps][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gm[i] = gamma[lane + 32 * i];
    ag[i] = 0.f;
    ab[i] = 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float* zr = z + (size_t)row * ldz;
    const float* dyr = dy + (size_t)row * lddy;
    float zv[NV], gv[NV], av[NV];
    float s = 0.f;
    // all loads of the row first (z, dy and the residual-branch gradient): issued inside the final loop the `add` loads were
    // 24 serialised DRAM round trips per row (0.30 ms instead of 0.06 ms per call)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      zv[i] = zr[lane + 32 * i];
      gv[i] = dyr[lane + 32 * i];
      av[i] = add != nullptr ? add[(size_t)row * ldadd + lane + 32 * i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) s += zv[i];
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      zv[i] -= mean;
      q += zv[i] * zv[i];
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      zv[i] *= rstd;  // xhat
      ag[i] = fmaf(gv[i], zv[i], ag[i]);
      ab[i] += gv[i];
      gv[i] *= gm[i];  // g = dy * gamma
      m1 += gv[i];
      m2 = fmaf(gv[i], zv[i], m2);
    }
    m1 = warp_sum(m1) / (float)d;
    m2 = warp_sum(m2) / (float)d;
    float* dzr = dz + (size_t)row * lddz;
#pragma unroll
    for (int i = 0; i < NV; ++i) dzr[lane + 32 * i] = fmaf(rstd, gv[i] - m1 - zv[i] * m2, av[i]);
  }
  // block partials: the 8 warps' accumulators are added in warp order
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) s_acc[warp * d + lane + 32 * i] = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_acc[w * d + c];
      part[(size_t)blockIdx.x * 2 * d + pass * d + c] = t;
    }
  }
}

// element-wise helpers of the backward pass
enum { ELT_MUL = 0, ELT_RELU_BWD = 1, ELT_GELU_BWD = 2, ELT_ADD = 3, ELT_SCALE = 4, ELT_QGELU_FWD = 5, ELT_QGELU_BWD = 6,
       ELT_AXPY = 7, ELT_GELU_FWD = 8 };
__global__ void __launch_bounds__(256)
eltwise_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b, float scale,
               float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = a[i];
    float r;
    switch (mode) {
      case ELT_MUL: r = x * b[i]; break;
      case ELT_RELU_BWD: r = b[i] > 0.f ? x : 0.f; break;
      case ELT_GELU_BWD: {  // b = pre-activation; d/db [0.5 b (1 + erf(b / sqrt 2))]
        const float u = b[i];
        r = x * (0.5f * (1.0f + erff(u * 0.70710678118654752440f)) + u * 0.3989422804014327f * expf(-0.5f * u * u));
        break;
      }
      case ELT_ADD: r = x + b[i]; break;
      case ELT_QGELU_FWD: r = x / (1.0f + __expf(-1.702f * x)); break;  // x * sigmoid(1.702 x), CLIP's QuickGELU
      case ELT_QGELU_BWD: {  // b = pre-activation: d/db [b sigmoid(1.702 b)] = s (1 + 1.702 b (1 - s))
        const float u = b[i];
        const float sg = 1.0f / (1.0f + __expf(-1.702f * u));
        r = x * sg * (1.0f + 1.702f * u * (1.0f - sg));
        break;
      }
      case ELT_AXPY: r = fmaf(scale, x, b[i]); break;  // scale * a + b
      case ELT_GELU_FWD: r = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); break;
      default: r = x * scale; break;
    }
    out[i] = r;
  }
}

// h16 = bf16(QuickGELU(h_pre)): the activation of the training forward without an fp32 intermediate
__global__ void __launch_bounds__(256)
qgelu_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    o.x = pack_bf16x2(v.x / (1.0f + __expf(-1.702f * v.x)), v.y / (1.0f + __expf(-1.702f * v.y)));
    o.y = pack_bf16x2(v.z / (1.0f + __expf(-1.702f * v.z)), v.w / (1.0f + __expf(-1.702f * v.w)));
    reinterpret_cast<uint2*>(y)[i] = o;
  }
}

// mean over T backward: out[b, t, :] = g[b, :] * scale
__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const float* __restrict__ g, float* __restrict__ out, int B, int T, int d, float scale) {
  const size_t n = (size_t)B * T * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % d);
    const size_t b = i / ((size_t)T * d);
    out[i] = g[b * d + j] * scale;
  }
}

// Masked attention backward for one (clip, head): probabilities recomputed in shared memory.
//   A = P (.) m  (m = attention-probability dropout mask, already scaled by 1/(1-p); NULL = ones)
//   O = A V;  dV = A^T dO;  dA = dO V^T;  dP = dA (.) m;  dS = P (.) (dP - rowsum(P (.) dP));
//   dQ = scale dS K;  dK = scale dS^T Q
__global__ void __launch_bounds__(256)
attention_masked_bwd_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                            const float* __restrict__ v, long long ldv, const uint8_t* __restrict__ key_valid,
                            const float* __restrict__ pmask, const float* __restrict__ dO, long long lddo,
                            float* __restrict__ dq, long long lddq, float* __restrict__ dk, long long lddk,
                            float* __restrict__ dv, long long lddv, int Tq, int Tk, int heads) {
  extern __shared__ float sm[];
  float* sq = sm;                      // [Tq][64]
  float* sdo = sq + (size_t)Tq * HD;   // [Tq][64]
  float* sk = sdo + (size_t)Tq * HD;   // [Tk][65]
  float* sv = sk + (size_t)Tk * 65;    // [Tk][65]
  float* sp = sv + (size_t)Tk * 65;    // [Tq][Tk + 1]
  const int ldp = Tk + 1;
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const uint8_t* kvb = key_valid ? key_valid + (size_t)b * Tk : nullptr;
  const float* pm = pmask ? pmask + ((size_t)b * heads + h) * Tq * Tk : nullptr;
  for (int i = tid; i < Tq * HD; i += blockDim.x) {
    const int r = i / HD, c = i % HD;
    sq[i] = q[((size_t)b * Tq + r) * ldq + h * HD + c];
    sdo[i] = dO[((size_t)b * Tq + r) * lddo + h * HD + c];
  }
  for (int i = tid; i < Tk * HD; i += blockDim.x) {
    const int r = i / HD, c = i % HD;
    sk[r * 65 + c] = k[((size_t)b * Tk + r) * ldk + h * HD + c];
    sv[r * 65 + c] = v[((size_t)b * Tk + r) * ldv + h * HD + c];
  }
  __syncthreads();
  // ---- P = softmax(scale q k^T + mask) ----
  for (int i = warp; i < Tq; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < Tk; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int c = 0; c < HD; ++c) s = fmaf(sq[i * HD + c], sk[j * 65 + c], s);
      s *= 0.125f;
      if (kvb && !kvb[j]) s = -INFINITY;
      sp[i * ldp + j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float s = sp[i * ldp + j];
      const float p = (s == -INFINITY) ? 0.f : __expf(s - mx);
      sp[i * ldp + j] = p;
      sum += p;
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < Tk; j += 32) sp[i * ldp + j] *= inv;
  }
  __syncthreads();
  // ---- dV_j = sum_i A_ij dO_i ----
  for (int j = warp; j < Tk; j += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < Tq; ++i) {
      float a = sp[i * ldp + j];
      if (pm) a *= pm[(size_t)i * Tk + j];
      a0 = fmaf(a, sdo[i * HD + lane], a0);
      a1 = fmaf(a, sdo[i * HD + lane + 32], a1);
    }
    float* o = dv + ((size_t)b * Tk + j) * lddv + h * HD;
    o[lane] = a0;
    o[lane + 32] = a1;
  }
  __syncthreads();
  // ---- dS = P (.) (dP - rowsum(P (.) dP)), in place over P ----
  for (int i = warp; i < Tq; i += nw) {
    float acc = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      float dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < HD; ++c) dp = fmaf(sdo[i * HD + c], sv[j * 65 + c], dp);
      if (pm) dp *= pm[(size_t)i * Tk + j];
      acc = fmaf(sp[i * ldp + j], dp, acc);
    }
    acc = warp_sum(acc);
    for (int j = lane; j < Tk; j += 32) {  // dP recomputed: cheaper than a second Tq x Tk buffer
      float dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < HD; ++c) dp = fmaf(sdo[i * HD + c], sv[j * 65 + c], dp);
      if (pm) dp *= pm[(size_t)i * Tk + j];
      sp[i * ldp + j] = sp[i * ldp + j] * (dp - acc);
    }
  }
  __syncthreads();
  // ---- dQ_i = scale sum_j dS_ij K_j ----
  for (int i = warp; i < Tq; i += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < Tk; ++j) {
      const float s = sp[i * ldp + j];
      a0 = fmaf(s, sk[j * 65 + lane], a0);
      a1 = fmaf(s, sk[j * 65 + lane + 32], a1);
    }
    float* o = dq + ((size_t)b * Tq + i) * lddq + h * HD;
    o[lane] = a0 * 0.125f;
    o[lane + 32] = a1 * 0.125f;
  }
  // ---- dK_j = scale sum_i dS_ij Q_i ----
  for (int j = warp; j < Tk; j += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < Tq; ++i) {
      const float s = sp[i * ldp + j];
      a0 = fmaf(s, sq[i * HD + lane], a0);
      a1 = fmaf(s, sq[i * HD + lane + 32], a1);
    }
    float* o = dk + ((size_t)b * Tk + j) * lddk + h * HD;
    o[lane] = a0 * 0.125f;
    o[lane + 32] = a1 * 0.125f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Masked attention backward for SHORT sequences (Tq, Tk <= 64: the 50-token ViT-B/32 student, 16-frame TFAM clips).
// The kernel above spends two shared-memory loads per FMA (ncu launch list of a student training step: 1.3 ms per layer,
// 27 % of the step).  Here the five products of the backward are small register-tiled GEMMs: 256 threads as a 16 x 16
// grid, each thread owns a 4 x 4 block of the output (rows ty + 16 a, columns tx + 16 b or head dims 4 tx .. 4 tx + 3),
// operands are read as float4 -- 8 LDS.128 per 64 FMAs in the K-contiguous products.
//   S = Q K^T / 8 -> P = softmax(S + mask);   dP = (dO V^T) (.) m;   dS = P (.) (dP - rowsum(P (.) dP))
//   dV = (P (.) m)^T dO;   dQ = dS K / 8;   dK = dS^T Q / 8
// ---------------------------------------------------------------------------------------------------------
constexpr int SB_LD = 68;   // row stride of the [T][64] operands: 16-byte aligned, conflict-free for quarter-warp LDS.128
constexpr int SB_LDP = 68;  // row stride of the [Tq][Tk] matrices (Tk <= 64)

// C[i][j] = sum_c A[i][c] B[j][c] over 64 columns; rows i = ty + 16 a, j = tx + 16 b
__device__ __forceinline__ void nt_tile(const float* __restrict__ A, const float* __restrict__ B, int ty, int tx,
                                        float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
  for (int c = 0; c < HD; c += 4) {
    float4 av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(A + (ty + 16 * a) * SB_LD + c);
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = *reinterpret_cast<const float4*>(B + (tx + 16 * b) * SB_LD + c);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        acc[a][b] = fmaf(av[a].x, bv[b].x, fmaf(av[a].y, bv[b].y, fmaf(av[a].z, bv[b].z, fmaf(av[a].w, bv[b].w, acc[a][b]))));
  }
}
// C[r][d] = sum_t A(t, r) B[t][d], r = ty + 16 a, d = 4 tx .. 4 tx + 3; TRANS: A(t, r) = A[t][r] else A[r][t]
template <bool TRANS>
__device__ __forceinline__ void tn_tile(const float* __restrict__ A, const float* __restrict__ B, int T, int ty, int tx,
                                        float4 (&acc)[4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float4 bv = *reinterpret_cast<const float4*>(B + t * SB_LD + 4 * tx);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float av = TRANS ? A[t * SB_LDP + ty + 16 * a] : A[(ty + 16 * a) * SB_LDP + t];
      acc[a].x = fmaf(av, bv.x, acc[a].x);
      acc[a].y = fmaf(av, bv.y, acc[a].y);
      acc[a].z = fmaf(av, bv.z, acc[a].z);
      acc[a].w = fmaf(av, bv.w, acc[a].w);
    }
  }
}

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

template <typename TQ>  // float (TFAM) or __nv_bfloat16 (the ViT's packed qkv buffer, read without a cast pass)
__global__ void __launch_bounds__(256)
attention_bwd_short_kernel(const TQ* __restrict__ q, long long ldq, const TQ* __restrict__ k, long long ldk,
                           const TQ* __restrict__ v, long long ldv, const uint8_t* __restrict__ key_valid,
                           const float* __restrict__ pmask, const float* __restrict__ dO, long long lddo,
                           float* __restrict__ dq, long long lddq, float* __restrict__ dk, long long lddk,
                           float* __restrict__ dv, long long lddv, int Tq, int Tk, int heads) {
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                  // [64][SB_LD]  (rows >= Tq zero)
  float* sdo = sq + 64 * SB_LD;
  float* sk = sdo + 64 * SB_LD;    // rows >= Tk zero
  float* sv = sk + 64 * SB_LD;
  float* sp = sv + 64 * SB_LD;     // [64][SB_LDP]  P, then A = P (.) m
  float* sds = sp + 64 * SB_LDP;   // dP (.) m, then dS
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
  const uint8_t* kvb = key_valid ? key_valid + (size_t)b * Tk : nullptr;
  const float* pm = pmask ? pmask + ((size_t)b * heads + h) * Tq * Tk : nullptr;
  for (int i = tid; i < 64 * 16; i += 256) {  // float4 granules
    const int r = i >> 4, c = (i & 15) * 4;
    float4 zq = make_float4(0.f, 0.f, 0.f, 0.f), zo = zq, zk = zq, zv = zq;
    if (r < Tq) {
      zq = load4(q + ((size_t)b * Tq + r) * ldq + h * HD + c);
      zo = load4(dO + ((size_t)b * Tq + r) * lddo + h * HD + c);
    }
    if (r < Tk) {
      zk = load4(k + ((size_t)b * Tk + r) * ldk + h * HD + c);
      zv = load4(v + ((size_t)b * Tk + r) * ldv + h * HD + c);
    }
    *reinterpret_cast<float4*>(sq + r * SB_LD + c) = zq;
    *reinterpret_cast<float4*>(sdo + r * SB_LD + c) = zo;
    *reinterpret_cast<float4*>(sk + r * SB_LD + c) = zk;
    *reinterpret_cast<float4*>(sv + r * SB_LD + c) = zv;
  }
  __syncthreads();
  float acc[4][4];
  nt_tile(sq, sk, ty, tx, acc);  // scores
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int j = tx + 16 * bb;
      const bool ok = j < Tk && !(kvb && !kvb[j]);
      sp[(ty + 16 * a) * SB_LDP + j] = ok ? acc[a][bb] * 0.125f : -INFINITY;
    }
  nt_tile(sdo, sv, ty, tx, acc);  // dP
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int i = ty + 16 * a, j = tx + 16 * bb;
      float dp = acc[a][bb];
      if (pm && i < Tq && j < Tk) dp *= pm[(size_t)i * Tk + j];
      sds[i * SB_LDP + j] = dp;
    }
  __syncthreads();
  // row pass: softmax, D_i, dS; A = P (.) m for dV
  for (int i = warp; i < 64; i += 8) {
    const float s0 = sp[i * SB_LDP + lane], s1 = sp[i * SB_LDP + lane + 32];
    const float mx = warp_max(fmaxf(s0, s1));
    float p0 = (s0 == -INFINITY) ? 0.f : __expf(s0 - mx), p1 = (s1 == -INFINITY) ? 0.f : __expf(s1 - mx);
    const float inv = 1.0f / warp_sum(p0 + p1);
    p0 *= inv;
    p1 *= inv;
    const float d0 = sds[i * SB_LDP + lane], d1 = sds[i * SB_LDP + lane + 32];
    const float D = warp_sum(p0 * d0 + p1 * d1);
    sds[i * SB_LDP + lane] = (i < Tq) ? p0 * (d0 - D) : 0.f;
    sds[i * SB_LDP + lane + 32] = (i < Tq) ? p1 * (d1 - D) : 0.f;
    float a0 = p0, a1 = p1;
    if (pm && i < Tq) {
      if (lane < Tk) a0 *= pm[(size_t)i * Tk + lane];
      if (lane + 32 < Tk) a1 *= pm[(size_t)i * Tk + lane + 32];
    }
    sp[i * SB_LDP + lane] = (i < Tq) ? a0 : 0.f;
    sp[i * SB_LDP + lane + 32] = (i < Tq) ? a1 : 0.f;
  }
  __syncthreads();
  float4 o[4];
  tn_tile<true>(sp, sdo, Tq, ty, tx, o);  // dV[j][d] = sum_i A[i][j] dO[i][d]
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = ty + 16 * a;
    if (j < Tk) *reinterpret_cast<float4*>(dv + ((size_t)b * Tk + j) * lddv + h * HD + 4 * tx) = o[a];
  }
  tn_tile<true>(sds, sq, Tq, ty, tx, o);  // dK[j][d] = sum_i dS[i][j] Q[i][d] / 8
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = ty + 16 * a;
    if (j < Tk)
      *reinterpret_cast<float4*>(dk + ((size_t)b * Tk + j) * lddk + h * HD + 4 * tx) =
          make_float4(o[a].x * 0.125f, o[a].y * 0.125f, o[a].z * 0.125f, o[a].w * 0.125f);
  }
  tn_tile<false>(sds, sk, Tk, ty, tx, o);  // dQ[i][d] = sum_j dS[i][j] K[j][d] / 8
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = ty + 16 * a;
    if (i < Tq)
      *reinterpret_cast<float4*>(dq + ((size_t)b * Tq + i) * lddq + h * HD + 4 * tx) =
          make_float4(o[a].x * 0.125f, o[a].y * 0.125f, o[a].z * 0.125f, o[a].w * 0.125f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled masked-attention backward for any Tq / Tk (the kernel above keeps the whole Tq x Tk probability matrix in
// shared memory: ~128 x 128).  Flash-attention style, fp32 SIMT, 32 x 32 tiles:
//   pass A (one block per 32 queries): stream the key tiles twice -- (1) online log-sum-exp and
//           D_i = sum_j P_ij m_ij (dO_i . v_j); (2) dS_ij = P_ij (m_ij dO_i . v_j - D_i), dQ_i += dS_ij k_j / 8.
//           lse and D go to global memory for pass B.
//   pass B (one block per 32 keys): stream the query tiles: P_ij = exp(s_ij - lse_i), dV_j += P_ij m_ij dO_i,
//           dK_j += dS_ij q_i / 8.
// Thread roles per tile: warp w computes the 8 x 32 block of scores of query rows 8 w .. 8 w + 7 (lane = key), then
// thread t accumulates 16 of the 64 head dims of row t / 4.
// ---------------------------------------------------------------------------------------------------------
constexpr int TB = 32;

__device__ __forceinline__ float dot64(const float* __restrict__ a, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll 16
  for (int c = 0; c < HD; ++c) s = fmaf(a[c], b[c], s);
  return s;
}

__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                   const float* __restrict__ v, long long ldv, const uint8_t* __restrict__ key_valid,
                   const float* __restrict__ pmask, const float* __restrict__ dO, long long lddo,
                   float* __restrict__ dq, long long lddq, float* __restrict__ lse_out, float* __restrict__ d_out,
                   int Tq, int Tk, int heads) {
  __shared__ float sq[TB][HD], sdo[TB][HD], sk[TB][HD + 1], sv[TB][HD + 1], sds[TB][TB + 1];
  __shared__ float s_lse[TB], s_d[TB];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* kvb = key_valid ? key_valid + (size_t)b * Tk : nullptr;
  const float* pm = pmask ? pmask + ((size_t)b * heads + h) * Tq * Tk : nullptr;
  for (int i = tid; i < TB * HD; i += 128) {
    const int r = i / HD, c = i % HD;
    const bool ok = q0 + r < Tq;
    sq[r][c] = ok ? q[((size_t)b * Tq + q0 + r) * ldq + h * HD + c] : 0.f;
    sdo[r][c] = ok ? dO[((size_t)b * Tq + q0 + r) * lddo + h * HD + c] : 0.f;
  }
  float m_run[8], l_run[8], d_run[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; d_run[i] = 0.f; }
  auto load_kv = [&](int k0) {
    __syncthreads();
    for (int i = tid; i < TB * HD; i += 128) {
      const int r = i / HD, c = i % HD;
      const bool ok = k0 + r < Tk;
      sk[r][c] = ok ? k[((size_t)b * Tk + k0 + r) * ldk + h * HD + c] : 0.f;
      sv[r][c] = ok ? v[((size_t)b * Tk + k0 + r) * ldv + h * HD + c] : 0.f;
    }
    __syncthreads();
  };
  // ---- pass 1: log-sum-exp and D, online over the key tiles ----
  for (int k0 = 0; k0 < Tk; k0 += TB) {
    load_kv(k0);
    const int j = k0 + lane;
    const bool jv = j < Tk && !(kvb && !kvb[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const float s = jv ? dot64(sq[r], sk[lane]) * 0.125f : -INFINITY;
      float dp = dot64(sdo[r], sv[lane]);
      if (pm && q0 + r < Tq && j < Tk) dp *= pm[(size_t)(q0 + r) * Tk + j];
      const float m_new = fmaxf(m_run[i], warp_max(s));
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run[i] - m_new);
      const float p = (s == -INFINITY) ? 0.f : __expf(s - m_new);
      l_run[i] = l_run[i] * corr + warp_sum(p);
      d_run[i] = d_run[i] * corr + warp_sum(p * dp);
      m_run[i] = m_new;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      s_lse[r] = m_run[i] + __logf(l_run[i]);
      s_d[r] = d_run[i] / l_run[i];
      if (q0 + r < Tq) {
        lse_out[((size_t)b * heads + h) * Tq + q0 + r] = s_lse[r];
        d_out[((size_t)b * heads + h) * Tq + q0 + r] = s_d[r];
      }
    }
  }
  // ---- pass 2: dQ ----
  const int orow = tid >> 2, oc = (tid & 3) * 16;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int k0 = 0; k0 < Tk; k0 += TB) {
    load_kv(k0);  // (also orders the s_lse / s_d writes before their first use)
    const int j = k0 + lane;
    const bool jv = j < Tk && !(kvb && !kvb[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      float ds = 0.f;
      if (jv) {
        const float p = __expf(dot64(sq[r], sk[lane]) * 0.125f - s_lse[r]);
        float dp = dot64(sdo[r], sv[lane]);
        if (pm && q0 + r < Tq) dp *= pm[(size_t)(q0 + r) * Tk + j];
        ds = p * (dp - s_d[r]);
      }
      sds[r][lane] = ds;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < TB; ++jj) {
      const float ds = sds[orow][jj];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fmaf(ds, sk[jj][oc + c], acc[c]);
    }
  }
  if (q0 + orow < Tq) {
    float* o = dq + ((size_t)b * Tq + q0 + orow) * lddq + h * HD + oc;
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = acc[c] * 0.125f;
  }
}

__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                    const float* __restrict__ v, long long ldv, const uint8_t* __restrict__ key_valid,
                    const float* __restrict__ pmask, const float* __restrict__ dO, long long lddo,
                    const float* __restrict__ lse, const float* __restrict__ dvec, float* __restrict__ dk,
                    long long lddk, float* __restrict__ dv, long long lddv, int Tq, int Tk, int heads) {
  __shared__ float sq[TB][HD + 1], sdo[TB][HD + 1], sk[TB][HD], sv[TB][HD], sp[TB][TB + 1], sds[TB][TB + 1];
  __shared__ float s_lse[TB], s_d[TB];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * TB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* kvb = key_valid ? key_valid + (size_t)b * Tk : nullptr;
  const float* pm = pmask ? pmask + ((size_t)b * heads + h) * Tq * Tk : nullptr;
  for (int i = tid; i < TB * HD; i += 128) {
    const int r = i / HD, c = i % HD;
    const bool ok = k0 + r < Tk;
    sk[r][c] = ok ? k[((size_t)b * Tk + k0 + r) * ldk + h * HD + c] : 0.f;
    sv[r][c] = ok ? v[((size_t)b * Tk + k0 + r) * ldv + h * HD + c] : 0.f;
  }
  const int orow = tid >> 2, oc = (tid & 3) * 16;  // this thread accumulates key row orow, head dims oc .. oc + 15
  float akk[16], avv[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { akk[c] = 0.f; avv[c] = 0.f; }
  for (int q0 = 0; q0 < Tq; q0 += TB) {
    __syncthreads();
    for (int i = tid; i < TB * HD; i += 128) {
      const int r = i / HD, c = i % HD;
      const bool ok = q0 + r < Tq;
      sq[r][c] = ok ? q[((size_t)b * Tq + q0 + r) * ldq + h * HD + c] : 0.f;
      sdo[r][c] = ok ? dO[((size_t)b * Tq + q0 + r) * lddo + h * HD + c] : 0.f;
    }
    if (tid < TB) {
      const bool ok = q0 + tid < Tq;
      s_lse[tid] = ok ? lse[((size_t)b * heads + h) * Tq + q0 + tid] : 0.f;
      s_d[tid] = ok ? dvec[((size_t)b * heads + h) * Tq + q0 + tid] : 0.f;
    }
    __syncthreads();
    // warp w: keys 8 w .. 8 w + 7 of the tile, lane = query row
    const int i = q0 + lane;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int r = warp * 8 + jj, j = k0 + r;
      float pmv = 0.f, ds = 0.f;
      if (i < Tq && j < Tk && !(kvb && !kvb[j])) {
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int c = 0; c < HD; ++c) {
          s = fmaf(sq[lane][c], sk[r][c], s);
          dp = fmaf(sdo[lane][c], sv[r][c], dp);
        }
        const float p = __expf(s * 0.125f - s_lse[lane]);
        const float mk = pm ? pm[(size_t)i * Tk + j] : 1.f;
        pmv = p * mk;
        ds = p * (dp * mk - s_d[lane]);
      }
      sp[r][lane] = pmv;
      sds[r][lane] = ds;
    }
    __syncthreads();
#pragma unroll 4
    for (int ii = 0; ii < TB; ++ii) {
      const float pv = sp[orow][ii], ds = sds[orow][ii];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        avv[c] = fmaf(pv, sdo[ii][oc + c], avv[c]);
        akk[c] = fmaf(ds, sq[ii][oc + c], akk[c]);
      }
    }
  }
  if (k0 + orow < Tk) {
    float* ok_ = dk + ((size_t)b * Tk + k0 + orow) * lddk + h * HD + oc;
    float* ov_ = dv + ((size_t)b * Tk + k0 + orow) * lddv + h * HD + oc;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      ok_[c] = akk[c] * 0.125f;
      ov_[c] = avv[c];
    }
  }
}

int grid_cap(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)vmc_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int vmc_transpose_split(const float* x, long long ldx, void* y, long long ldy, int R, int C, int form,
                        void* stream) {
  VMC_CHECK_ARG(x && y, VMC_ERR_ARG, "vmc_transpose_split: null pointer");
  VMC_CHECK_ARG(R > 0 && C > 0 && ldx >= C && ldy >= 3LL * R && (form == 0 || form == 1), VMC_ERR_SHAPE,
                "vmc_transpose_split: bad shape R=%d C=%d ldx=%lld ldy=%lld", R, C, ldx, ldy);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  transpose_split_kernel<<<grid, 256, 0, st>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, R, C, form);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_transpose_cast(const void* x, int src_bf16, long long ldx, void* y, long long ldy, int R, int C, void* stream) {
  VMC_CHECK_ARG(x && y, VMC_ERR_ARG, "vmc_transpose_cast: null pointer");
  VMC_CHECK_ARG(R > 0 && C > 0 && ldx >= C && ldy >= R, VMC_ERR_SHAPE, "vmc_transpose_cast: bad shape R=%d C=%d ldx=%lld ldy=%lld",
                R, C, ldx, ldy);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  if (src_bf16)
    transpose_cast_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                                               reinterpret_cast<__nv_bfloat16*>(y), ldy, R, C);
  else
    transpose_cast_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), ldx,
                                                       reinterpret_cast<__nv_bfloat16*>(y), ldy, R, C);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_cast_f32(const void* x, long long ldx, float* y, long long ldy, int rows, int d, void* stream) {
  VMC_CHECK_ARG(x && y && rows > 0 && d > 0 && ldx >= d && ldy >= d, VMC_ERR_ARG, "vmc_cast_f32: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  cast_f32_kernel<<<grid_cap((size_t)rows * d), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, y, ldy, rows, d);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_colsum_slices(int R, int C) {
  if (R <= 0 || C <= 0) return -1;
  if (R < 2048) return 1;  // single-stage kernel
  const int col_blocks = (C + 31) / 32;
  int slices = (4 * vmc_num_sms() + col_blocks - 1) / col_blocks;  // ~4 blocks per SM in total
  const int max_slices = (R + 255) / 256;
  slices = slices < 1 ? 1 : (slices > max_slices ? max_slices : slices);
  return slices > 64 ? 64 : slices;
}

int vmc_colsum(const float* x, long long ldx, const float* y, long long ldy, float* out, int R, int C,
               int accumulate, float* workspace, void* stream) {
  VMC_CHECK_ARG(x && out, VMC_ERR_ARG, "vmc_colsum: null pointer");
  VMC_CHECK_ARG(R > 0 && C > 0, VMC_ERR_SHAPE, "vmc_colsum: bad shape R=%d C=%d", R, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, (double)R * C * (y ? 8.0 : 4.0));
  const int slices = vmc_colsum_slices(R, C);
  if (workspace != nullptr && slices > 1) {  // workspace: slices * C floats
    const int rows_per = ((R + slices - 1) / slices + 7) / 8 * 8;
    colsum_slice_kernel<<<dim3((C + 31) / 32, slices), 256, 0, st>>>(x, ldx, y, ldy, workspace, R, C, rows_per);
    colsum_finish_kernel<<<(C + 255) / 256, 256, 0, st>>>(workspace, out, slices, C, accumulate);
    VMC_LAUNCH_CHECK();
    vmc_count_launch(2);
    return VMC_OK;
  }
  colsum_kernel<<<(C + 31) / 32, 256, 0, st>>>(x, ldx, y, ldy, out, R, C, accumulate);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_cast_colsum_slices(int R, int C) {
  if (R <= 0 || C <= 0) return -1;
  const int col_blocks = (C + 127) / 128;
  int slices = (8 * vmc_num_sms() + col_blocks - 1) / col_blocks;  // ~8 blocks of 256 threads per SM: enough loads in flight
  const int max_slices = (R + 63) / 64;
  slices = slices < 1 ? 1 : (slices > max_slices ? max_slices : slices);
  return slices > 128 ? 128 : slices;
}

int vmc_cast_colsum(const float* x, long long ldx, const float* aux, long long ldaux, void* y16, long long ldy, float* colsum,
                    int R, int C, float* workspace, void* stream) {
  VMC_CHECK_ARG(x && y16 && colsum && workspace, VMC_ERR_ARG, "vmc_cast_colsum: null pointer");
  VMC_CHECK_ARG(R > 0 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && (aux == nullptr || (ldaux % 4) == 0), VMC_ERR_SHAPE,
                "vmc_cast_colsum: C and the row strides must be multiples of 4 (R=%d C=%d)", R, C);
  VMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(aux)) & 15) == 0 && (reinterpret_cast<uintptr_t>(y16) & 7) == 0,
                VMC_ERR_ALIGN, "vmc_cast_colsum: misaligned pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, (double)R * C * (aux ? 10.0 : 6.0));
  const int slices = vmc_cast_colsum_slices(R, C);
  const int rows_per = ((R + slices - 1) / slices + 7) / 8 * 8;
  const dim3 grid((C + 127) / 128, slices);
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(y16);
  if (aux) cast_colsum_kernel<true><<<grid, 256, 0, st>>>(x, ldx, aux, ldaux, y, ldy, workspace, R, C, rows_per);
  else cast_colsum_kernel<false><<<grid, 256, 0, st>>>(x, ldx, nullptr, 0, y, ldy, workspace, R, C, rows_per);
  colsum_kernel<<<(C + 31) / 32, 256, 0, st>>>(workspace, C, nullptr, 0, colsum, slices, C, 0);  // slices in row order: deterministic
  VMC_LAUNCH_CHECK();
  vmc_count_launch(2);
  return VMC_OK;
}

int vmc_layernorm_bwd(const float* z, long long ldz, const float* gamma, float eps, const float* dy,
                      long long lddy, float* dz, long long lddz, float* xhat, long long ldxh, int rows, int d,
                      void* stream) {
  VMC_CHECK_ARG(z && gamma && dy && dz, VMC_ERR_ARG, "vmc_layernorm_bwd: null pointer");
  VMC_CHECK_ARG(rows > 0 && d > 0, VMC_ERR_SHAPE, "vmc_layernorm_bwd: bad shape rows=%d d=%d", rows, d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_LAYERNORM, st, 0.0, (double)rows * d * 16.0);
  layernorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(z, ldz, gamma, eps, dy, lddy, dz, lddz, xhat, ldxh, rows, d);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_layernorm_bwd_fused_blocks(int rows) {
  const int want = (rows + 7) / 8, cap = vmc_num_sms() * 2;
  return want < cap ? want : cap;
}

int vmc_layernorm_bwd_fused(const float* z, long long ldz, const float* gamma, float eps, const float* dy, long long lddy,
                            const float* add, long long ldadd, float* dz, long long lddz, float* dgamma_dbeta, float* workspace,
                            int rows, int d, void* stream) {
  VMC_CHECK_ARG(z && gamma && dy && dz && dgamma_dbeta && workspace, VMC_ERR_ARG, "vmc_layernorm_bwd_fused: null pointer");
  VMC_CHECK_ARG(rows > 0 && (d == 512 || d == 768 || d == 1024), VMC_ERR_SHAPE, "vmc_layernorm_bwd_fused: d must be 512, 768 or 1024 (d=%d)", d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = vmc_layernorm_bwd_fused_blocks(rows);
  const size_t smem = (size_t)8 * d * sizeof(float);
  {
    VmcProfScope prof(VMC_K_LAYERNORM, st, 0.0, (double)rows * d * (add ? 16.0 : 12.0));
    if (d == 512) {
      layernorm_bwd_fused_kernel<16><<<blocks, 256, smem, st>>>(z, ldz, gamma, eps, dy, lddy, add, ldadd, dz, lddz, workspace, rows);
    } else if (d == 768) {
      layernorm_bwd_fused_kernel<24><<<blocks, 256, smem, st>>>(z, ldz, gamma, eps, dy, lddy, add, ldadd, dz, lddz, workspace, rows);
    } else {
      layernorm_bwd_fused_kernel<32><<<blocks, 256, smem, st>>>(z, ldz, gamma, eps, dy, lddy, add, ldadd, dz, lddz, workspace, rows);
    }
  }
  // the per-block partials [blocks, 2 d] are summed by the wide column-sum kernel (one block per 32 columns, 8 row lanes, four
  // loads in flight per thread): colsum_finish_kernel's one-thread-per-column loop over hundreds of blocks took 0.25 ms
  colsum_kernel<<<(2 * d + 31) / 32, 256, 0, st>>>(workspace, 2 * d, nullptr, 0, dgamma_dbeta, blocks, 2 * d, 0);
  VMC_LAUNCH_CHECK();
  vmc_count_launch(2);
  return VMC_OK;
}

int vmc_eltwise(int mode, const float* a, const float* b, float scale, float* out, long long n, void* stream) {
  VMC_CHECK_ARG(a && out && n > 0 && mode >= ELT_MUL && mode <= ELT_GELU_FWD, VMC_ERR_ARG, "vmc_eltwise: bad argument");
  VMC_CHECK_ARG(b != nullptr || mode == ELT_SCALE || mode == ELT_QGELU_FWD || mode == ELT_GELU_FWD, VMC_ERR_ARG,
                "vmc_eltwise: mode %d needs a second operand", mode);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  eltwise_kernel<<<grid_cap((size_t)n), 256, 0, st>>>(mode, a, b, scale, out, (size_t)n);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_qgelu_cast(const float* x, void* y, long long n, void* stream) {
  VMC_CHECK_ARG(x && y && n > 0 && (n % 4) == 0, VMC_ERR_ARG, "vmc_qgelu_cast: n must be a positive multiple of 4");
  VMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(y) & 7)) == 0, VMC_ERR_ALIGN,
                "vmc_qgelu_cast: misaligned pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 6.0 * n);
  qgelu_cast_kernel<<<grid_cap((size_t)n / 4), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y), (size_t)n / 4);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_broadcast_rows(const float* g, float* out, int B, int T, int d, float scale, void* stream) {
  VMC_CHECK_ARG(g && out && B > 0 && T > 0 && d > 0, VMC_ERR_ARG, "vmc_broadcast_rows: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  broadcast_rows_kernel<<<grid_cap((size_t)B * T * d), 256, 0, st>>>(g, out, B, T, d, scale);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// ViT self-attention backward for short sequences (L <= 64: ViT-B/32, the reference's training default) on the warp-level
// tensor path (ldmatrix + mma.sync.m16n8k16, bf16 operands, fp32 accumulation).  The register-tiled CUDA-core kernel above
// runs the five 64 x 64 x 64 products of an item at 19 TFLOP/s (6.2 ms of a 30 ms training step).  One CTA of 4 warps per
// (frame, head); Q, K, V (bf16, from the saved qkv buffer) and dO (fp32 -> bf16) are staged as 64 x 64 tiles with 128-byte
// rows whose 16-byte chunks are XOR-swizzled with (row & 7), so every ldmatrix is conflict free.
//   phase 1, warp w = queries 16 w .. 16 w + 15:  S = Q K^T / 8 and dP = dO V^T (A fragments of Q / dO, B fragments of K / V
//     straight from the token-major tiles), row softmax P and delta = sum_j P dP in registers (a row lives in one lane quad),
//     dS = P (dP - delta) / 8, then dQ = dS K with dS re-packed from accumulator to A-fragment layout and K through
//     ldmatrix.trans; P and dS (bf16) go to shared memory;
//   phase 2, warp w = keys 16 w .. 16 w + 15:  dV = P^T dO and dK = dS^T Q, the transposed A fragments through ldmatrix.trans.
// Every sum runs in a fixed order: deterministic.  Operands are rounded to bf16 once (P, dS, dO), like every GEMM of the step.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int chunk) {
  return base + (uint32_t)row * 128u + (((uint32_t)chunk ^ ((uint32_t)row & 7u)) << 4);
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128)
attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ dO, long long lddo,
                         float* __restrict__ dqkv, int L, int heads) {
  extern __shared__ __align__(1024) uint8_t bw_smem[];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bw_smem);
  const uint32_t sQ = sb, sK = sb + 8192u, sV = sb + 16384u, sD = sb + 24576u, sP = sb + 32768u, sS = sb + 40960u;
  const int head = blockIdx.x, frame = blockIdx.y;
  const int d = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const size_t row0 = (size_t)frame * L;

  // ---- stage Q, K, V (bf16) and dO (fp32 -> bf16); rows >= L are zero.  ALL global loads are issued before the first shared
  // store: written as load -> store per chunk, the 16 loads of a thread were serialised round trips (ncu: 70 % of the samples
  // on the first STS; ~12 us per item) ----
  {
    uint4 v[12];
    float4 lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + 128 * j;
      const int mat = i >> 9, r = (i >> 3) & 63, c = i & 7;
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (r < L) v[j] = __ldg(reinterpret_cast<const uint4*>(qkv + (row0 + r) * 3 * d + (size_t)mat * d + head * HD) + c);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + 128 * j;
      const int r = i >> 3, c = i & 7;
      lo[j] = hi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < L) {
        const float4* src = reinterpret_cast<const float4*>(dO + (row0 + r) * lddo + head * HD + c * 8);
        lo[j] = __ldg(src);
        hi[j] = __ldg(src + 1);
      }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + 128 * j;
      const int mat = i >> 9, r = (i >> 3) & 63, c = i & 7;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(swz(sb + (uint32_t)mat * 8192u, r, c)), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w) : "memory");
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + 128 * j;
      const int r = i >> 3, c = i & 7;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(swz(sD, r, c)), "r"(pack_bf16x2(lo[j].x, lo[j].y)), "r"(pack_bf16x2(lo[j].z, lo[j].w)),
                   "r"(pack_bf16x2(hi[j].x, hi[j].y)), "r"(pack_bf16x2(hi[j].z, hi[j].w)) : "memory");
    }
  }
  __syncthreads();

  const int lm = lane >> 3, lr = lane & 7;  // ldmatrix: this lane supplies row lr of matrix lm
  // ================= phase 1: this warp's 16 queries =================
  {
    const int q0 = 16 * warp;
    uint32_t aQ[4][4], aD[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm4(swz(sQ, q0 + (lm & 1) * 8 + lr, 2 * ks + (lm >> 1)), aQ[ks]);
      ldsm4(swz(sD, q0 + (lm & 1) * 8 + lr, 2 * ks + (lm >> 1)), aD[ks]);
    }
    float sa[8][4], pa[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) sa[nb][e] = pa[nb][e] = 0.f;
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // key blocks 2 np, 2 np + 1
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4], bv[4];
        ldsm4(swz(sK, 8 * (2 * np + (lm >> 1)) + lr, 2 * ks + (lm & 1)), bk);
        ldsm4(swz(sV, 8 * (2 * np + (lm >> 1)) + lr, 2 * ks + (lm & 1)), bv);
        mma16816(sa[2 * np], aQ[ks], bk[0], bk[1]);
        mma16816(sa[2 * np + 1], aQ[ks], bk[2], bk[3]);
        mma16816(pa[2 * np], aD[ks], bv[0], bv[1]);
        mma16816(pa[2 * np + 1], aD[ks], bv[2], bv[3]);
      }
    }
    // softmax over the keys (rows g and g + 8 of the warp's block; a row is spread over the 4 lanes of a quad)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = 8 * nb + 2 * tig + e < L;
        sa[nb][e] = ok ? sa[nb][e] * 0.125f : -INFINITY;
        sa[nb][2 + e] = ok ? sa[nb][2 + e] * 0.125f : -INFINITY;
        mx0 = fmaxf(mx0, sa[nb][e]);
        mx1 = fmaxf(mx1, sa[nb][2 + e]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sa[nb][e] = __expf(sa[nb][e] - mx0);
        sa[nb][2 + e] = __expf(sa[nb][2 + e] - mx1);
        sm0 += sa[nb][e];
        sm1 += sa[nb][2 + e];
      }
    sm0 += __shfl_xor_sync(0xffffffffu, sm0, 1); sm0 += __shfl_xor_sync(0xffffffffu, sm0, 2);
    sm1 += __shfl_xor_sync(0xffffffffu, sm1, 1); sm1 += __shfl_xor_sync(0xffffffffu, sm1, 2);
    const float i0 = 1.0f / sm0, i1 = 1.0f / sm1;
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sa[nb][e] *= i0;
        sa[nb][2 + e] *= i1;
        dl0 = fmaf(sa[nb][e], pa[nb][e], dl0);
        dl1 = fmaf(sa[nb][2 + e], pa[nb][2 + e], dl1);
      }
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1); dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1); dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
    // P and dS = P (dP - delta) / 8 as bf16: to shared memory for phase 2, dS also as the A operand of dQ = dS K
    uint32_t pds[8][2];  // dS pairs of rows g / g + 8
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float s00 = sa[nb][0] * (pa[nb][0] - dl0) * 0.125f, s01 = sa[nb][1] * (pa[nb][1] - dl0) * 0.125f;
      const float s10 = sa[nb][2] * (pa[nb][2] - dl1) * 0.125f, s11 = sa[nb][3] * (pa[nb][3] - dl1) * 0.125f;
      pds[nb][0] = pack_bf16x2(s00, s01);
      pds[nb][1] = pack_bf16x2(s10, s11);
      const uint32_t o = 4u * (uint32_t)tig;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sP, q0 + g, nb) + o), "r"(pack_bf16x2(sa[nb][0], sa[nb][1])) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sP, q0 + g + 8, nb) + o), "r"(pack_bf16x2(sa[nb][2], sa[nb][3])) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sS, q0 + g, nb) + o), "r"(pds[nb][0]) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sS, q0 + g + 8, nb) + o), "r"(pds[nb][1]) : "memory");
    }
    // dQ = dS K: A = dS (k-step kk = key blocks 2 kk, 2 kk + 1), B[k = key][n = dim] through ldmatrix.trans of the K tile
    float dq[8][4];
#pragma unroll
    for (int db = 0; db < 8; ++db) dq[db][0] = dq[db][1] = dq[db][2] = dq[db][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a[4] = {pds[2 * kk][0], pds[2 * kk][1], pds[2 * kk + 1][0], pds[2 * kk + 1][1]};
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        uint32_t b[4];
        ldsm4t(swz(sK, 16 * kk + (lm & 1) * 8 + lr, 2 * cp + (lm >> 1)), b);
        mma16816(dq[2 * cp], a, b[0], b[1]);
        mma16816(dq[2 * cp + 1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int tok = q0 + g + 8 * h;
      if (tok < L) {
        float* o = dqkv + (row0 + tok) * 3 * d + head * HD + 2 * tig;
#pragma unroll
        for (int db = 0; db < 8; ++db) *reinterpret_cast<float2*>(o + 8 * db) = make_float2(dq[db][2 * h], dq[db][2 * h + 1]);
      }
    }
  }
  __syncthreads();
  // ================= phase 2: this warp's 16 keys: dV = P^T dO, dK = dS^T Q =================
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const uint32_t sA = which == 0 ? sP : sS, sB = which == 0 ? sD : sQ;
    float acc[8][4];
#pragma unroll
    for (int db = 0; db < 8; ++db) acc[db][0] = acc[db][1] = acc[db][2] = acc[db][3] = 0.f;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {  // 16 queries per k-step
      uint32_t a[4];  // A[m = key][k = query] = transposed 8 x 8 blocks of the [query][key] tile
      ldsm4t(swz(sA, 16 * kq + (lm >> 1) * 8 + lr, 2 * warp + (lm & 1)), a);
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        uint32_t b[4];  // B[k = query][n = dim]
        ldsm4t(swz(sB, 16 * kq + (lm & 1) * 8 + lr, 2 * cp + (lm >> 1)), b);
        mma16816(acc[2 * cp], a, b[0], b[1]);
        mma16816(acc[2 * cp + 1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int tok = 16 * warp + g + 8 * h;
      if (tok < L) {
        float* o = dqkv + (row0 + tok) * 3 * d + (size_t)(which == 0 ? 2 : 1) * d + head * HD + 2 * tig;
#pragma unroll
        for (int db = 0; db < 8; ++db) *reinterpret_cast<float2*>(o + 8 * db) = make_float2(acc[db][2 * h], acc[db][2 * h + 1]);
      }
    }
  }
}

// Forward for short sequences on the same warp-level tensor path (VMC_OPT_ATTN_IMPL = 8): one CTA of 4 warps per (frame, head),
// 24 KB of shared memory (Q, K, V tiles) -> nine CTAs per SM hide the staging latency that the persistent tcgen05 kernel (v7: two
// tiles in flight per SM, bound by its S -> softmax -> PV -> drain chain) cannot.  A warp owns 16 queries: S = Q K^T (ldmatrix +
// mma.sync), single-pass row softmax inside a lane quad, P kept in registers and re-packed from accumulator to A-fragment layout
// (bf16, the normaliser sums the ROUNDED values), O = P V with V through ldmatrix.trans, rows staged through the warp's own
// (dead) Q rows so that they leave as full 128-byte lines.
extern "C++" {
// HM = true: what-if input layout [F, heads, 3, L, 64] (head-major: an item's Q, K, V are contiguous runs of L * 128 bytes) instead of
// the packed [F * L, 3 d] rows the qkv GEMM writes -- tools/kernel_bench.py impl 81, to test whether the access pattern is the ceiling
template <bool HM>
__global__ void __launch_bounds__(128)
attention_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int L, int heads) {
  extern __shared__ __align__(1024) uint8_t fw_smem[];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(fw_smem);
  const uint32_t sQ = sb, sK = sb + 8192u, sV = sb + 16384u;
  const int head = blockIdx.x, frame = blockIdx.y;
  const int d = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const size_t row0 = (size_t)frame * L;
  {
    uint4 v[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + 128 * j;
      const int mat = i >> 9, r = (i >> 3) & 63, c = i & 7;
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (r < L)
        v[j] = HM ? __ldg(reinterpret_cast<const uint4*>(qkv + ((((size_t)frame * heads + head) * 3 + mat) * L + r) * HD) + c)
                  : __ldg(reinterpret_cast<const uint4*>(qkv + (row0 + r) * 3 * d + (size_t)mat * d + head * HD) + c);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int i = tid + 128 * j;
      const int mat = i >> 9, r = (i >> 3) & 63, c = i & 7;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(swz(sb + (uint32_t)mat * 8192u, r, c)), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w) : "memory");
    }
  }
  __syncthreads();
  const int q0 = 16 * warp;
  if (q0 >= L) return;  // warp-uniform: no query of this warp exists (no further block-wide barrier below)
  const int lm = lane >> 3, lr = lane & 7;
  uint32_t aQ[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm4(swz(sQ, q0 + (lm & 1) * 8 + lr, 2 * ks + (lm >> 1)), aQ[ks]);
  float sa[8][4];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) sa[nb][0] = sa[nb][1] = sa[nb][2] = sa[nb][3] = 0.f;
#pragma unroll
  for (int np = 0; np < 4; ++np)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bk[4];
      ldsm4(swz(sK, 8 * (2 * np + (lm >> 1)) + lr, 2 * ks + (lm & 1)), bk);
      mma16816(sa[2 * np], aQ[ks], bk[0], bk[1]);
      mma16816(sa[2 * np + 1], aQ[ks], bk[2], bk[3]);
    }
  const float sc = 0.125f * 1.4426950408889634f;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = 8 * nb + 2 * tig + e < L;
      sa[nb][e] = ok ? sa[nb][e] * sc : -INFINITY;
      sa[nb][2 + e] = ok ? sa[nb][2 + e] * sc : -INFINITY;
      mx0 = fmaxf(mx0, sa[nb][e]);
      mx1 = fmaxf(mx1, sa[nb][2 + e]);
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  uint32_t pp[8][2];
  float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const __nv_bfloat162 r0 = __floats2bfloat162_rn(exp2f(sa[nb][0] - mx0), exp2f(sa[nb][1] - mx0));
    const __nv_bfloat162 r1 = __floats2bfloat162_rn(exp2f(sa[nb][2] - mx1), exp2f(sa[nb][3] - mx1));
    sm0 += __low2float(r0) + __high2float(r0);
    sm1 += __low2float(r1) + __high2float(r1);
    pp[nb][0] = *reinterpret_cast<const uint32_t*>(&r0);
    pp[nb][1] = *reinterpret_cast<const uint32_t*>(&r1);
  }
  sm0 += __shfl_xor_sync(0xffffffffu, sm0, 1); sm0 += __shfl_xor_sync(0xffffffffu, sm0, 2);
  sm1 += __shfl_xor_sync(0xffffffffu, sm1, 1); sm1 += __shfl_xor_sync(0xffffffffu, sm1, 2);
  float oa[8][4];
#pragma unroll
  for (int db = 0; db < 8; ++db) oa[db][0] = oa[db][1] = oa[db][2] = oa[db][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a[4] = {pp[2 * kk][0], pp[2 * kk][1], pp[2 * kk + 1][0], pp[2 * kk + 1][1]};
#pragma unroll
    for (int cp = 0; cp < 4; ++cp) {
      uint32_t b[4];
      ldsm4t(swz(sV, 16 * kk + (lm & 1) * 8 + lr, 2 * cp + (lm >> 1)), b);
      mma16816(oa[2 * cp], a, b[0], b[1]);
      mma16816(oa[2 * cp + 1], a, b[2], b[3]);
    }
  }
  // rows g / g + 8 -> the warp's own Q rows (dead: the A fragments are in registers), then out as full 128-byte lines
  const float i0 = 1.0f / sm0, i1 = 1.0f / sm1;
  __syncwarp();
#pragma unroll
  for (int db = 0; db < 8; ++db) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sQ, q0 + g, db) + 4u * (uint32_t)tig), "r"(pack_bf16x2(oa[db][0] * i0, oa[db][1] * i0)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(swz(sQ, q0 + g + 8, db) + 4u * (uint32_t)tig), "r"(pack_bf16x2(oa[db][2] * i1, oa[db][3] * i1)) : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + 4 * i + (lane >> 3), c = lane & 7;
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(swz(sQ, r, c)));
    if (r < L) *reinterpret_cast<uint4*>(out + (row0 + r) * d + head * HD + c * 8) = v;
  }
}


}  // extern "C++"

int vmc_attention_vit_short_mma(const void* qkv, void* out, int F, int L, int heads, void* stream) {
  VMC_CHECK_ARG(qkv && out, VMC_ERR_ARG, "vmc_attention_vit_short_mma: null pointer");
  VMC_CHECK_ARG(F > 0 && F <= 65535 && heads > 0 && L > 0 && L <= 64, VMC_ERR_SHAPE, "vmc_attention_vit_short_mma: need 0 < L <= 64 (L=%d)", L);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int d = heads * HD;
  {
    VmcProfScope prof(VMC_K_ATTN_VIT, st, 4.0 * F * heads * (double)L * L * HD, 8.0 * F * L * d);
    if (vmc_get_option(VMC_OPT_ATTN_PREFETCH) == 81)  // what-if: head-major input layout (timing experiment only)
      attention_fwd_mma_kernel<true><<<dim3(heads, F), 128, 24576, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                        reinterpret_cast<__nv_bfloat16*>(out), L, heads);
    else
      attention_fwd_mma_kernel<false><<<dim3(heads, F), 128, 24576, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                         reinterpret_cast<__nv_bfloat16*>(out), L, heads);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_attention_vit_bwd_short(const void* qkv, const float* dO, long long lddo, float* dqkv, int F, int L, int heads,
                                void* stream) {
  if (vmc_get_option(VMC_OPT_ATTN_BWD_IMPL) == 0 && F > 0 && F <= 65535 && heads > 0 && L > 0 && L <= 64 && qkv && dO && dqkv &&
      (lddo % 4) == 0 && (reinterpret_cast<uintptr_t>(dO) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dqkv) & 7) == 0) {
    // default: warp-level tensor-core kernel (VMC_OPT_ATTN_BWD_IMPL = 2 selects the register-tiled CUDA-core kernel, 1 the first one)
    cudaStream_t stm = reinterpret_cast<cudaStream_t>(stream);
    {
      VmcProfScope prof(VMC_K_ATTN_SMALL, stm, 10.0 * F * heads * (double)L * L * HD, 0.0);
      attention_bwd_mma_kernel<<<dim3(heads, F), 128, 49152, stm>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), dO, lddo, dqkv, L, heads);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  VMC_CHECK_ARG(qkv && dO && dqkv, VMC_ERR_ARG, "vmc_attention_vit_bwd_short: null pointer");
  VMC_CHECK_ARG(F > 0 && heads > 0 && L > 0 && L <= 64 && F <= 65535 && (lddo % 4) == 0, VMC_ERR_SHAPE,
                "vmc_attention_vit_bwd_short: need 0 < L <= 64 tokens (L=%d); longer towers go through vmc_attention_masked_bwd", L);
  const int d = heads * HD;
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(qkv);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem_s = (size_t)(4 * 64 * SB_LD + 2 * 64 * SB_LDP) * sizeof(float);
  VMC_CUDA(cudaFuncSetAttribute(attention_bwd_short_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem_s));
  {
    VmcProfScope prof(VMC_K_ATTN_SMALL, st, 10.0 * F * heads * (double)L * L * HD, 0.0);
    attention_bwd_short_kernel<__nv_bfloat16><<<dim3(heads, F), 256, smem_s, st>>>(
        p, 3LL * d, p + d, 3LL * d, p + 2 * d, 3LL * d, nullptr, nullptr, dO, lddo, dqkv, 3LL * d, dqkv + d, 3LL * d,
        dqkv + 2 * d, 3LL * d, L, L, heads);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_attention_masked_bwd(const float* q, long long ldq, const float* k, long long ldk, const float* v,
                             long long ldv, const uint8_t* key_valid, const float* prob_mask, const float* dO,
                             long long lddo, float* dq, long long lddq, float* dk, long long lddk, float* dv,
                             long long lddv, int B, int Tq, int Tk, int heads, float* workspace, void* stream) {
  VMC_CHECK_ARG(q && k && v && dO && dq && dk && dv, VMC_ERR_ARG, "vmc_attention_masked_bwd: null pointer");
  VMC_CHECK_ARG(B > 0 && heads > 0 && Tq > 0 && Tk > 0 && B <= 65535, VMC_ERR_SHAPE,
                "vmc_attention_masked_bwd: bad shape B=%d Tq=%d Tk=%d heads=%d", B, Tq, Tk, heads);
  const size_t smem = ((size_t)2 * Tq * HD + (size_t)2 * Tk * 65 + (size_t)Tq * (Tk + 1)) * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned16 = ((ldq | ldk | ldv | lddo | lddq | lddk | lddv) % 4) == 0 &&
                         ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                           reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                           reinterpret_cast<uintptr_t>(dv)) & 15) == 0;
  if (workspace == nullptr && Tq <= 64 && Tk <= 64 && aligned16 && vmc_get_option(VMC_OPT_ATTN_BWD_IMPL) != 1) {
    // short sequences: register-tiled kernel (VMC_OPT_ATTN_BWD_IMPL = 1 selects the first-generation kernel for cross-checks)
    const size_t smem_s = (size_t)(4 * 64 * SB_LD + 2 * 64 * SB_LDP) * sizeof(float);
    VMC_CUDA(cudaFuncSetAttribute(attention_bwd_short_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    {
      VmcProfScope prof(VMC_K_ATTN_SMALL, st, 10.0 * B * heads * (double)Tq * Tk * HD, 0.0);
      attention_bwd_short_kernel<float><<<dim3(heads, B), 256, smem_s, st>>>(q, ldq, k, ldk, v, ldv, key_valid, prob_mask, dO,
                                                                             lddo, dq, lddq, dk, lddk, dv, lddv, Tq, Tk, heads);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  if (smem > 200 * 1024 || workspace != nullptr) {
    // tiled path: any Tq / Tk; workspace = 2 * B * heads * Tq floats (log-sum-exp and D per query row)
    VMC_CHECK_ARG(workspace != nullptr, VMC_ERR_WORKSPACE,
                  "vmc_attention_masked_bwd: Tq=%d, Tk=%d need the tiled path: pass a workspace of 2*B*heads*Tq floats", Tq, Tk);
    VMC_CHECK_ARG(heads <= 65535, VMC_ERR_SHAPE, "vmc_attention_masked_bwd: too many heads");
    float* lse = workspace;
    float* dvec = workspace + (size_t)B * heads * Tq;
    VmcProfScope prof(VMC_K_ATTN_SMALL, st, 14.0 * B * heads * (double)Tq * Tk * HD, 0.0);
    attn_bwd_dq_kernel<<<dim3((Tq + TB - 1) / TB, heads, B), 128, 0, st>>>(q, ldq, k, ldk, v, ldv, key_valid, prob_mask, dO,
                                                                            lddo, dq, lddq, lse, dvec, Tq, Tk, heads);
    attn_bwd_dkv_kernel<<<dim3((Tk + TB - 1) / TB, heads, B), 128, 0, st>>>(q, ldq, k, ldk, v, ldv, key_valid, prob_mask, dO,
                                                                             lddo, lse, dvec, dk, lddk, dv, lddv, Tq, Tk, heads);
    VMC_LAUNCH_CHECK();
    vmc_count_launch(2);
    return VMC_OK;
  }
  VMC_CUDA(cudaFuncSetAttribute(attention_masked_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  dim3 grid(heads, B);
  {
    VmcProfScope prof(VMC_K_ATTN_SMALL, st, 10.0 * B * heads * (double)Tq * Tk * HD, 0.0);
    attention_masked_bwd_kernel<<<grid, 256, smem, st>>>(q, ldq, k, ldk, v, ldv, key_valid, prob_mask, dO, lddo, dq,
                                                         lddq, dk, lddk, dv, lddv, Tq, Tk, heads);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // extern "C"
