// HBM-bound kernels of the path: the uint8 prologue (P1), frame difference (a1),
// LayerNorm (E1), temporal mean, casts and the cosine distillation loss (E2).
// All are plain coalesced / 16-byte vectorised kernels: there is no data reuse, so
// no shared-memory staging; grids are sized from the data, in 256-thread blocks.
#include "common.cuh"
#include "vimoclip_b200.h"

namespace {

using namespace vmc;

// CLIP normalisation constants (clip/clip.py::_transform; HF CLIPImageProcessor defaults)
__constant__ float c_mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
__constant__ float c_std[3] = {0.26862954f, 0.26130258f, 0.27577711f};

// RN(1/std_c) and RN(1/255) as fp32 bit patterns (numpy float32 division, correctly rounded)
__constant__ float c_rstd[3] = {3.7226030826568604f /*0x406e3f0e*/, 3.8269808292388916f /*0x4074ed41*/,
                                3.6261167526245117f /*0x4068124c*/};
#define VMC_R255 0.0039215688593685627f /* 0x3b808081 */

// Correctly rounded a / b from the correctly rounded reciprocal rb = RN(1/b):
// q0 = a*rb; r = fma(-q0, b, a); q = fma(r, rb, q0).  Verified EXHAUSTIVELY (exact rational
// arithmetic) to equal the IEEE division for every operand pair this file feeds it: u/255 for the
// 256 byte values and (u/255 - mean_c)/std_c for the 768 (u, c) pairs.  Three FMA-pipe instructions
// instead of the ~10-instruction IEEE division sequence, which made the first prologue compute-bound.
__device__ __forceinline__ float div_by_const(float a, float b, float rb) {
  const float q0 = __fmul_rn(a, rb);
  const float r = __fmaf_rn(-q0, b, a);
  return __fmaf_rn(r, rb, q0);
}

// (float32(u8) / 255 - mean) / std, bit-for-bit torchvision ToTensor + Normalize
// (models/student_model.py:78 via clip _transform): fp32, true division semantics, no contraction.
__device__ __forceinline__ float normalise_px(uint32_t u8, int c) {
  const float t = div_by_const(static_cast<float>(u8), 255.0f, VMC_R255);
  return div_by_const(__fsub_rn(t, c_mean[c]), c_std[c], c_rstd[c]);
}

// to_pil_image(float CHW) = pic.mul(255).byte(): fp32 multiply, then float -> uint8 through
// int64 truncation, keeping the low 8 bits.
__device__ __forceinline__ uint32_t wrap_f32(float x) {
  const float y = __fmul_rn(x, 255.0f);
  const long long i = static_cast<long long>(y);  // cvt.rzi.s64.f32 (truncation)
  return static_cast<uint32_t>(i) & 255u;
}

// Store 16 horizontally consecutive values (channel c, row y, columns x0..x0+15) of frame f,
// either as fp32 NCHW or as bf16 in the patchified GEMM-operand layout.
__device__ __forceinline__ void store_f16px(void* dst, int dst_kind, int f, int c, int y, int x0,
                                            const float (&v)[16], int H, int W, int p, int ld) {
  if (dst_kind == VMC_DST_F32_NCHW) {
    float* d = reinterpret_cast<float*>(dst) + (((size_t)f * 3 + c) * H + y) * W + x0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(d)[i] =
          make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
  }
  // VMC_DST_BF16_PATCH: row = f*n + (y/p)*(W/p) + x/p ; col = c*p*p + (y%p)*p + x%p
  const int gw = W / p;
  const int n = (H / p) * gw;
  const int py = y / p, iy = y - py * p;
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(dst);
  if ((p & 15) == 0) {
    const int px = x0 / p, ix = x0 - px * p;
    __nv_bfloat16* d =
        base + ((size_t)f * n + (size_t)py * gw + px) * ld + (size_t)c * p * p + iy * p + ix;
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);
    o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);
    o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);
    o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]);
    o1.w = pack_bf16x2(v[14], v[15]);
    reinterpret_cast<uint4*>(d)[0] = o0;
    reinterpret_cast<uint4*>(d)[1] = o1;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int x = x0 + i;
      const int px = x / p, ix = x - px * p;
      base[((size_t)f * n + (size_t)py * gw + px) * ld + (size_t)c * p * p + iy * p + ix] =
          __float2bfloat16_rn(v[i]);
    }
  }
}

// Store 16 horizontally consecutive uint8 pixels per dst_kind (wrapped u8 / normalised fp32 / bf16 patches).
__device__ __forceinline__ void store_px16(void* dst, int dst_kind, int f, int c, int y, int x0,
                                           const uint32_t (&u)[16], int H, int W, int p, int ld) {
  if (dst_kind == VMC_DST_U8) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = u[4 * i] | (u[4 * i + 1] << 8) | (u[4 * i + 2] << 16) | (u[4 * i + 3] << 24);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + (((size_t)f * 3 + c) * H + y) * W + x0;
    *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
    return;
  }
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = normalise_px(u[i], c);
  store_f16px(dst, dst_kind, f, c, y, x0, v, H, W, p, ld);
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
prologue_kernel(const void* __restrict__ src, int src_kind, void* __restrict__ dst, int dst_kind,
                int F, int H, int W, int p, int ld) {
  const int wv = W / 16;
  const size_t total = (size_t)F * 3 * H * wv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool f32_src = src_kind == VMC_SRC_F32_WRAP || src_kind == VMC_SRC_F32_NORM;
  for (size_t idx0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += UNROLL * stride) {
    // issue all UNROLL 16-byte loads first: ~4x the bytes in flight per thread (HBM latency hiding)
    uint4 q[UNROLL];
    if (!f32_src) {
#pragma unroll
      for (int k = 0; k < UNROLL; ++k) {
        const size_t idx = idx0 + k * stride;
        if (idx < total)
          q[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + idx * 16));
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const size_t idx = idx0 + k * stride;
      if (idx >= total) break;
      // 32-bit index arithmetic (the host checks total < 2^31): 64-bit div/mod cost more than the pixels
      const uint32_t i32 = (uint32_t)idx;
      const uint32_t row = i32 / (uint32_t)wv;
      const int xv = (int)(i32 - row * (uint32_t)wv);
      const uint32_t plane = row / (uint32_t)H;
      const int y = (int)(row - plane * (uint32_t)H);
      const int f = (int)(plane / 3u);
      const int c = (int)(plane - 3u * (uint32_t)f);
      const size_t off = idx * 16;  // == (((f*3 + c)*H + y)*W + xv*16): the source is contiguous
      uint32_t u[16];
      if (src_kind == VMC_SRC_F32_NORM) {
        // already-normalised pixel_values (HF get_image_features input): layout change + bf16 only
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = __ldg(s + i);
          v[4 * i + 0] = t.x;
          v[4 * i + 1] = t.y;
          v[4 * i + 2] = t.z;
          v[4 * i + 3] = t.w;
        }
        store_f16px(dst, dst_kind, f, c, y, xv * 16, v, H, W, p, ld);
        continue;
      }
      if (src_kind == VMC_SRC_F32_WRAP) {
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = __ldg(s + i);
          u[4 * i + 0] = wrap_f32(t.x);
          u[4 * i + 1] = wrap_f32(t.y);
          u[4 * i + 2] = wrap_f32(t.z);
          u[4 * i + 3] = wrap_f32(t.w);
        }
      } else {
        const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 255u;
          // regime A: uint8 -> float 0..255 -> *255 -> int64 -> low 8 bits == (-x) mod 256
          if (src_kind == VMC_SRC_U8_WRAP) b = (0u - b) & 255u;
          u[i] = b;
        }
      }
      store_px16(dst, dst_kind, f, c, y, xv * 16, u, H, W, p, ld);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Patch-matrix prologue, coalesced on both sides.  One block iteration = one "band": the p image
// rows of one (frame, channel, patch-row), which are p*W CONTIGUOUS source bytes.  The band is
// loaded with 16-byte coalesced loads, normalised, scattered into shared memory in patch-major
// order (element (iy, x) -> patch x/p, offset iy*p + x%p), and written out as W/p contiguous
// p*p-element runs (512 B at p = 16, 2 KB at p = 32) of the bf16 patch matrix.
// The first version (prologue_kernel) stored 32 bytes per thread 1.5 KB apart: 30 % of HBM peak.
// ---------------------------------------------------------------------------------------
template <typename LoadPx>
__device__ __forceinline__ void band_to_patches(LoadPx load16, __nv_bfloat16* __restrict__ tile,
                                                __nv_bfloat16* __restrict__ dst, int f, int c, int py,
                                                int H, int W, int p, int ld) {
  const int gw = W / p;
  const int n = (H / p) * gw;
  const int pp = p * p;
  const int chunks_in = p * (W / 16);  // 16-pixel chunks of the band
  for (int q = threadIdx.x; q < chunks_in; q += blockDim.x) {
    const int iy = q / (W / 16);
    const int x0 = (q - iy * (W / 16)) * 16;
    float v[16];
    load16(iy, x0, v);
    if ((p & 15) == 0) {
      const int px = x0 / p, ix = x0 - px * p;
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]);
      o0.y = pack_bf16x2(v[2], v[3]);
      o0.z = pack_bf16x2(v[4], v[5]);
      o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]);
      o1.y = pack_bf16x2(v[10], v[11]);
      o1.z = pack_bf16x2(v[12], v[13]);
      o1.w = pack_bf16x2(v[14], v[15]);
      uint4* t = reinterpret_cast<uint4*>(tile + px * pp + iy * p + ix);
      t[0] = o0;
      t[1] = o1;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int x = x0 + i;
        const int px = x / p, ix = x - px * p;
        tile[px * pp + iy * p + ix] = __float2bfloat16_rn(v[i]);
      }
    }
  }
  __syncthreads();
  __nv_bfloat16* drow = dst + ((size_t)f * n + (size_t)py * gw) * ld + (size_t)c * pp;
  if ((pp & 7) == 0) {
    const int per_patch = pp / 8;  // 16-byte chunks per patch-channel run
    for (int q = threadIdx.x; q < gw * per_patch; q += blockDim.x) {
      const int px = q / per_patch, w = q - px * per_patch;
      reinterpret_cast<uint4*>(drow + (size_t)px * ld)[w] = reinterpret_cast<const uint4*>(tile + px * pp)[w];
    }
  } else {
    const int per_patch = pp / 4;  // 8-byte chunks (p = 14: 196 elements per run)
    for (int q = threadIdx.x; q < gw * per_patch; q += blockDim.x) {
      const int px = q / per_patch, w = q - px * per_patch;
      reinterpret_cast<uint2*>(drow + (size_t)px * ld)[w] = reinterpret_cast<const uint2*>(tile + px * pp)[w];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
prologue_patch_kernel(const void* __restrict__ src, int src_kind, __nv_bfloat16* __restrict__ dst,
                      int F, int H, int W, int p, int ld) {
  extern __shared__ __align__(16) uint8_t smem_tile[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_tile);
  const int gh = H / p;
  const int bands = F * 3 * gh;
  for (int b = blockIdx.x; b < bands; b += gridDim.x) {
    const int py = b % gh;
    const int c = (b / gh) % 3;
    const int f = b / (3 * gh);
    const size_t off = (((size_t)f * 3 + c) * H + (size_t)py * p) * W;
    if (src_kind == VMC_SRC_U8 || src_kind == VMC_SRC_U8_WRAP) {
      const uint8_t* s = reinterpret_cast<const uint8_t*>(src) + off;
      const bool wrap = src_kind == VMC_SRC_U8_WRAP;
      band_to_patches(
          [&](int iy, int x0, float (&v)[16]) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(s + (size_t)iy * W + x0));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              uint32_t u = (w[i >> 2] >> (8 * (i & 3))) & 255u;
              if (wrap) u = (0u - u) & 255u;
              v[i] = normalise_px(u, c);
            }
          },
          tile, dst, f, c, py, H, W, p, ld);
    } else {
      const float* s = reinterpret_cast<const float*>(src) + off;
      const bool norm = src_kind == VMC_SRC_F32_NORM;
      band_to_patches(
          [&](int iy, int x0, float (&v)[16]) {
            const float4* q4 = reinterpret_cast<const float4*>(s + (size_t)iy * W + x0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 q = __ldg(q4 + i);
              const float in[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) v[4 * i + k] = norm ? in[k] : normalise_px(wrap_f32(in[k]), c);
            }
          },
          tile, dst, f, c, py, H, W, p, ld);
    }
  }
}

// OpenCV 8-bit BGR2GRAY: (B*3735 + G*19235 + R*9798 + (1<<14)) >> 15
// (utils/generate_frame_diff_video.py:37,46; exhaustively verified)
__device__ __forceinline__ uint32_t bgr_gray(uint32_t b, uint32_t g, uint32_t r) {
  return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256)
frame_diff_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ diff_u8,
                  void* __restrict__ dst, int dst_kind, int clips, int T, int H, int W, int p,
                  int ld) {
  const int wv = W / 16;
  const size_t total = (size_t)clips * T * H * wv;
  const size_t frame_bytes = (size_t)H * W * 3;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i32 = (uint32_t)idx;
    const uint32_t row = i32 / (uint32_t)wv;
    const int xv = (int)(i32 - row * (uint32_t)wv);
    const uint32_t fr = row / (uint32_t)H;
    const int y = (int)(row - fr * (uint32_t)H);
    const int clip = (int)(fr / (uint32_t)T);
    const int t = (int)(fr - (uint32_t)clip * (uint32_t)T);
    const size_t f0 = (size_t)clip * (T + 1) + t;  // previous frame; current = f0 + 1
    const size_t off = f0 * frame_bytes + ((size_t)y * W + (size_t)xv * 16) * 3;
    const uint4* s0 = reinterpret_cast<const uint4*>(bgr + off);
    const uint4* s1 = reinterpret_cast<const uint4*>(bgr + off + frame_bytes);
    uint32_t w0[12], w1[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 a = __ldg(s0 + i);
      const uint4 b = __ldg(s1 + i);
      w0[4 * i] = a.x; w0[4 * i + 1] = a.y; w0[4 * i + 2] = a.z; w0[4 * i + 3] = a.w;
      w1[4 * i] = b.x; w1[4 * i + 1] = b.y; w1[4 * i + 2] = b.z; w1[4 * i + 3] = b.w;
    }
    uint32_t d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t c0[3], c1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int byte = 3 * i + k;
        c0[k] = (w0[byte >> 2] >> (8 * (byte & 3))) & 255u;
        c1[k] = (w1[byte >> 2] >> (8 * (byte & 3))) & 255u;
      }
      const int g0 = (int)bgr_gray(c0[0], c0[1], c0[2]);
      const int g1 = (int)bgr_gray(c1[0], c1[1], c1[2]);
      d[i] = (uint32_t)(g1 > g0 ? g1 - g0 : g0 - g1);  // cv2.absdiff(curr, prev)
    }
    const size_t fo = (size_t)clip * T + t;
    if (diff_u8 != nullptr) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w[i] = d[4 * i] | (d[4 * i + 1] << 8) | (d[4 * i + 2] << 16) | (d[4 * i + 3] << 24);
      *reinterpret_cast<uint4*>(diff_u8 + (fo * H + y) * W + (size_t)xv * 16) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (dst != nullptr) {
      // the student sees uint8 frames with three identical channels (regime A wrap)
      uint32_t u[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) u[i] = (0u - d[i]) & 255u;
#pragma unroll
      for (int c = 0; c < 3; ++c) store_px16(dst, dst_kind, (int)fo, c, y, xv * 16, u, H, W, p, ld);
    }
  }
}

// Frame difference -> student patch matrix, band-wise (see band_to_patches).  The grey difference of
// the band is computed once into shared memory (uint8), then emitted for the 3 identical channels.
__global__ void __launch_bounds__(256)
frame_diff_patch_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ diff_u8,
                        __nv_bfloat16* __restrict__ dst, int clips, int T, int H, int W, int p, int ld) {
  extern __shared__ __align__(16) uint8_t smem_tile[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_tile);
  uint8_t* dband = smem_tile + (size_t)p * W * 2;  // p*W wrapped-diff bytes
  const int gh = H / p;
  const int bands = clips * T * gh;
  const size_t frame_bytes = (size_t)H * W * 3;
  for (int b = blockIdx.x; b < bands; b += gridDim.x) {
    const int py = b % gh;
    const int t = (b / gh) % T;
    const int clip = b / (gh * T);
    const size_t f0 = (size_t)clip * (T + 1) + t;
    const size_t fo = (size_t)clip * T + t;
    const uint8_t* s0 = bgr + f0 * frame_bytes + (size_t)py * p * W * 3;
    const uint8_t* s1 = s0 + frame_bytes;
    for (int q = threadIdx.x; q < p * (W / 16); q += blockDim.x) {
      const uint4* a4 = reinterpret_cast<const uint4*>(s0 + (size_t)q * 48);
      const uint4* b4 = reinterpret_cast<const uint4*>(s1 + (size_t)q * 48);
      uint32_t w0[12], w1[12];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 a = __ldg(a4 + i);
        const uint4 bb = __ldg(b4 + i);
        w0[4 * i] = a.x; w0[4 * i + 1] = a.y; w0[4 * i + 2] = a.z; w0[4 * i + 3] = a.w;
        w1[4 * i] = bb.x; w1[4 * i + 1] = bb.y; w1[4 * i + 2] = bb.z; w1[4 * i + 3] = bb.w;
      }
      uint32_t d[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t c0[3], c1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int byte = 3 * i + k;
          c0[k] = (w0[byte >> 2] >> (8 * (byte & 3))) & 255u;
          c1[k] = (w1[byte >> 2] >> (8 * (byte & 3))) & 255u;
        }
        const int g0 = (int)bgr_gray(c0[0], c0[1], c0[2]);
        const int g1 = (int)bgr_gray(c1[0], c1[1], c1[2]);
        d[i] = (uint32_t)(g1 > g0 ? g1 - g0 : g0 - g1);
      }
      uint32_t wd[4], ww[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        wd[i] = d[4 * i] | (d[4 * i + 1] << 8) | (d[4 * i + 2] << 16) | (d[4 * i + 3] << 24);
        ww[i] = ((0u - d[4 * i]) & 255u) | (((0u - d[4 * i + 1]) & 255u) << 8) |
                (((0u - d[4 * i + 2]) & 255u) << 16) | (((0u - d[4 * i + 3]) & 255u) << 24);
      }
      if (diff_u8 != nullptr)
        *reinterpret_cast<uint4*>(diff_u8 + (fo * H + (size_t)py * p) * W + (size_t)q * 16) =
            make_uint4(wd[0], wd[1], wd[2], wd[3]);
      *reinterpret_cast<uint4*>(dband + (size_t)q * 16) = make_uint4(ww[0], ww[1], ww[2], ww[3]);
    }
    __syncthreads();
    for (int c = 0; c < 3; ++c) {
      band_to_patches(
          [&](int iy, int x0, float (&v)[16]) {
            const uint4 q = *reinterpret_cast<const uint4*>(dband + (size_t)iy * W + x0);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = normalise_px((w[i >> 2] >> (8 * (i & 3))) & 255u, c);
          },
          tile, dst, (int)fo, c, py, H, W, p, ld);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Patch-matrix prologue in GATHER form (default for p % 8 == 0): threads are laid out over the OUTPUT.
// One block = one band (frame f, patch-row py); work item q = (patch px, 16-byte output chunk j) with j
// fastest, so a warp writes 512 contiguous bytes of one patch row (4 full 128-byte lines) where the
// direct kernel writes 32 half-sectors 1.5 KB apart.  The matching 8 source pixels are an 8-byte load
// (32 bytes for float sources); neighbouring patches of the band pick up the other half of each source
// sector from L1/L2, so DRAM traffic stays at the algorithmic 1 + 2 bytes per pixel.  All loads of a
// thread are issued before the first pixel is converted.
// ---------------------------------------------------------------------------------------
// bf16 patch output only: bf16(RN_fp32((u/255 - mean_c)/std_c)) == bf16(fma(float(u), K_c, B_c)) with
// K_c = RN(1/(255 std_c)), B_c = RN(-mean_c/std_c) for ALL 768 (u, c) pairs (checked exhaustively on the host
// in exact arithmetic, tests/test_host_cpu.py::test_fast_normalise_constants, and on the GPU against the
// exact chain of the direct kernel).  One FMA instead of two verified divisions: the exact chain costs
// 16 instructions per pixel, which kept the patch prologue at ~60 % of HBM peak.
__constant__ uint32_t c_fastK[3] = {0x3c6f2e3du, 0x3c75e324u, 0x3c68fb48u};
__constant__ uint32_t c_fastB[3] = {0xbfe568dcu, 0xbfe044b8u, 0xbfbd77d8u};

// per-byte (-x) mod 256 of four packed uint8 (regime A wrap), no cross-byte carries
__device__ __forceinline__ uint32_t neg4_u8(uint32_t w) {
  const uint32_t x = ~w;
  return ((x & 0x7f7f7f7fu) + 0x01010101u) ^ (x & 0x80808080u);
}
// byte i of w as an exact float: PRMT builds the bit pattern of 2^23 + b, one FADD removes the offset
template <int I>
__device__ __forceinline__ float byte_to_float(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7550 + I)) - 8388608.0f;
}
__device__ __forceinline__ uint4 normalise_pack8_fast(uint32_t w0, uint32_t w1, int c) {
  const float K = __uint_as_float(c_fastK[c]), B = __uint_as_float(c_fastB[c]);
  uint4 o;
  o.x = pack_bf16x2(__fmaf_rn(byte_to_float<0>(w0), K, B), __fmaf_rn(byte_to_float<1>(w0), K, B));
  o.y = pack_bf16x2(__fmaf_rn(byte_to_float<2>(w0), K, B), __fmaf_rn(byte_to_float<3>(w0), K, B));
  o.z = pack_bf16x2(__fmaf_rn(byte_to_float<0>(w1), K, B), __fmaf_rn(byte_to_float<1>(w1), K, B));
  o.w = pack_bf16x2(__fmaf_rn(byte_to_float<2>(w1), K, B), __fmaf_rn(byte_to_float<3>(w1), K, B));
  return o;
}

template <int P, int ITERS, bool F32SRC>
__global__ void __launch_bounds__(384)
prologue_gather_kernel(const void* __restrict__ src, int src_kind, __nv_bfloat16* __restrict__ dst,
                       int H, int W, int ld) {
  constexpr int PP = P * P;
  constexpr int CPR = 3 * PP / 8;  // 16-byte chunks per patch row
  constexpr int CPC = PP / 8;      // chunks per channel
  constexpr int CPL = P / 8;       // chunks per image line of a patch
  const int gw = W / P, gh = H / P;
  const int f = blockIdx.x / gh, py = blockIdx.x - f * gh;
  const int items = gw * CPR;
  const size_t plane = (size_t)H * W;
  const size_t band0 = (size_t)f * 3 * plane + (size_t)py * P * W;
  __nv_bfloat16* drow = dst + ((size_t)f * gh * gw + (size_t)py * gw) * ld;

  uint2 raw[F32SRC ? 1 : ITERS];
  float4 rawf[F32SRC ? ITERS : 1][2];
  if constexpr (!F32SRC) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int q = it * blockDim.x + threadIdx.x;
      if (q < items) {
        const int px = q / CPR, j = q - px * CPR;
        const int c = j / CPC, r = j - c * CPC;
        const int iy = r / CPL, h = r - iy * CPL;
        raw[it] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(src) + band0 +
                                                       c * plane + (size_t)iy * W + px * P + h * 8));
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int q = it * blockDim.x + threadIdx.x;
      if (q < items) {
        const int px = q / CPR, j = q - px * CPR;
        const int c = j / CPC, r = j - c * CPC;
        const int iy = r / CPL, h = r - iy * CPL;
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + band0 +
                                                          c * plane + (size_t)iy * W + px * P + h * 8);
        rawf[it][0] = __ldg(s);
        rawf[it][1] = __ldg(s + 1);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    if (q >= items) break;
    const int px = q / CPR, j = q - px * CPR;
    const int c = j / CPC;
    uint4 o;
    if constexpr (F32SRC) {
      const float4 lo = rawf[it][0], hi = rawf[it][1];
      if (src_kind == VMC_SRC_F32_NORM) {  // already-normalised pixel_values: layout change + bf16 only
        o = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y),
                       pack_bf16x2(hi.z, hi.w));
      } else {
        const uint32_t w0 = wrap_f32(lo.x) | (wrap_f32(lo.y) << 8) | (wrap_f32(lo.z) << 16) | (wrap_f32(lo.w) << 24);
        const uint32_t w1 = wrap_f32(hi.x) | (wrap_f32(hi.y) << 8) | (wrap_f32(hi.z) << 16) | (wrap_f32(hi.w) << 24);
        o = normalise_pack8_fast(w0, w1, c);
      }
    } else {
      uint32_t w0 = raw[it].x, w1 = raw[it].y;
      if (src_kind == VMC_SRC_U8_WRAP) {  // regime A: (-x) mod 256
        w0 = neg4_u8(w0);
        w1 = neg4_u8(w1);
      }
      o = normalise_pack8_fast(w0, w1, c);
    }
    reinterpret_cast<uint4*>(drow + (size_t)px * ld)[j] = o;
  }
}

// uint8 sources, 16 pixels per work item (one 16-byte load -> two adjacent 16-byte output chunks = 32 contiguous bytes per thread,
// 1 KB per warp): same arithmetic as prologue_gather_kernel, but the index math, the wrap and the load / store instructions are
// amortised over twice as many pixels (~5.8 instead of 7 instructions per pixel).  Inside the full step the SM clock sits at
// ~1.4 GHz under the power cap, where the 8-pixel form becomes issue-bound (0.61-0.68 of the HBM peak in-step vs 0.89-0.92 alone).
template <int P, int ITERS>
__global__ void __launch_bounds__(384)
prologue_gather16_kernel(const uint8_t* __restrict__ src, int src_kind, __nv_bfloat16* __restrict__ dst, int H, int W, int ld) {
  constexpr int PP = P * P;
  constexpr int CPR = 3 * PP / 16;  // 16-pixel items per patch row
  constexpr int CPC = PP / 16;      // per channel
  constexpr int CPL = P / 16;       // per image line of a patch
  const int gw = W / P, gh = H / P;
  const int f = blockIdx.x / gh, py = blockIdx.x - f * gh;
  const int items = gw * CPR;
  const size_t plane = (size_t)H * W;
  const size_t band0 = (size_t)f * 3 * plane + (size_t)py * P * W;
  __nv_bfloat16* drow = dst + ((size_t)f * gh * gw + (size_t)py * gw) * ld;
  uint4 raw[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    if (q < items) {
      const int px = q / CPR, j = q - px * CPR;
      const int c = j / CPC, r = j - c * CPC;
      const int iy = r / CPL, h = r - iy * CPL;
      raw[it] = __ldg(reinterpret_cast<const uint4*>(src + band0 + c * plane + (size_t)iy * W + px * P + h * 16));
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    if (q >= items) break;
    const int px = q / CPR, j = q - px * CPR;
    const int c = j / CPC;
    uint32_t w0 = raw[it].x, w1 = raw[it].y, w2 = raw[it].z, w3 = raw[it].w;
    if (src_kind == VMC_SRC_U8_WRAP) {
      w0 = neg4_u8(w0);
      w1 = neg4_u8(w1);
      w2 = neg4_u8(w2);
      w3 = neg4_u8(w3);
    }
    uint4* o = reinterpret_cast<uint4*>(drow + (size_t)px * ld) + 2 * j;
    o[0] = normalise_pack8_fast(w0, w1, c);
    o[1] = normalise_pack8_fast(w2, w3, c);
  }
}

// Gather form for patch sizes whose lines are not 8-pixel multiples (ViT-L/14: p = 14, patch row 588 elements, ld = 592):
// same output-ordered work split as prologue_gather_kernel (one block per band, work item = (patch, 16-byte output
// chunk), a warp writes 512 contiguous bytes), but a chunk's 8 source pixels straddle image lines / channels, so they
// are fetched as single elements (L1 hits: the band's lines are read completely by this block) with the position
// advanced incrementally.  The pad columns [3 p^2, ld) are written here (zeros): no separate memset pass.
template <bool F32SRC, int ITERS>
__global__ void __launch_bounds__(256)
prologue_gather_any_kernel(const void* __restrict__ src, int src_kind, __nv_bfloat16* __restrict__ dst,
                           int H, int W, int P, int ld) {
  const int PP = P * P;
  const int CPR = ld / 8;  // 16-byte chunks per patch row, padding included
  const int gw = W / P, gh = H / P;
  const int f = blockIdx.x / gh, py = blockIdx.x - f * gh;
  const int items = gw * CPR;
  const size_t plane = (size_t)H * W;
  const size_t band0 = (size_t)f * 3 * plane + (size_t)py * P * W;
  __nv_bfloat16* drow = dst + ((size_t)f * gh * gw + (size_t)py * gw) * ld;
  const float K0 = __uint_as_float(c_fastK[0]), K1 = __uint_as_float(c_fastK[1]), K2 = __uint_as_float(c_fastK[2]);
  const float B0 = __uint_as_float(c_fastB[0]), B1 = __uint_as_float(c_fastB[1]), B2 = __uint_as_float(c_fastB[2]);
  for (int q0 = 0; q0 < items; q0 += ITERS * (int)blockDim.x) {
    float x[ITERS][8];       // raw source values (bytes as exact floats)
    uint32_t ccode[ITERS];   // 2 bits per element: channel, 3 = padding
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int q = q0 + it * (int)blockDim.x + (int)threadIdx.x;
      ccode[it] = 0xFFFFu;
      if (q < items) {
        const int px = q / CPR, j = q - px * CPR;
        const int e = 8 * j;
        int c = e / PP;
        const int r = e - c * PP;
        int iy = r / P, ix = r - iy * P;
        size_t off = band0 + (size_t)c * plane + (size_t)iy * W + (size_t)px * P + ix;
        uint32_t code = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = c < 3;
          if constexpr (F32SRC) x[it][i] = ok ? __ldg(reinterpret_cast<const float*>(src) + off) : 0.f;
          else x[it][i] = ok ? (float)__ldg(reinterpret_cast<const uint8_t*>(src) + off) : 0.f;
          code |= (uint32_t)(ok ? c : 3) << (2 * i);
          ++ix;
          ++off;
          if (ix == P) {
            ix = 0;
            ++iy;
            off += (size_t)(W - P);
            if (iy == P) {
              iy = 0;
              ++c;
              off += plane - (size_t)P * W;
            }
          }
        }
        ccode[it] = code;
      }
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int q = q0 + it * (int)blockDim.x + (int)threadIdx.x;
      if (q >= items) break;
      const int px = q / CPR, j = q - px * CPR;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t c = (ccode[it] >> (2 * i)) & 3u;
        float u = x[it][i];
        if (F32SRC && src_kind == VMC_SRC_F32_NORM) {
          v[i] = c == 3u ? 0.f : u;
        } else {
          if constexpr (F32SRC) u = (float)wrap_f32(u);
          else if (src_kind == VMC_SRC_U8_WRAP) u = (float)((256u - (uint32_t)u) & 255u);
          const float K = c == 0u ? K0 : (c == 1u ? K1 : K2), B = c == 0u ? B0 : (c == 1u ? B1 : B2);
          v[i] = c == 3u ? 0.f : __fmaf_rn(u, K, B);
        }
      }
      reinterpret_cast<uint4*>(drow + (size_t)px * ld)[j] =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// ViT-L/14 (p = 14), uint8 sources: the element-wise gather above is bound by L1 wavefronts (a warp's byte loads touch
// ~19 image lines), and a block that loads its band, synchronises, computes, synchronises and stores keeps too few
// bytes in flight (3.3 TB/s).  This kernel is persistent and moves whole bands with 1-D bulk async copies (TMA) on
// both sides, double buffered:
//   1. load: the P lines of a channel are one contiguous run of P * W bytes -> three cp.async.bulk per band into
//      stage k & 1, completion on an mbarrier, issued one band ahead;
//   2. compute: work item = (patch, line l = c * P + iy): P pixels of ONE channel -> P/2 aligned pixel pairs
//      (LDS.U16; P even, so a pair never straddles a line), the single-FMA normalise with one (K, B) pair per item,
//      P/2 32-bit stores into the staged patch rows (row stride + 16 B: 2-way instead of 8-way bank conflicts);
//   3. store: one cp.async.bulk per patch row (ld bf16, pad columns [3 p^2, ld) = zeros written once), read out of
//      the other output stage while the next band is computed (cp.async.bulk.wait_group.read).
template <int P>
__global__ void __launch_bounds__(256)
prologue_band_gather_u8_kernel(const uint8_t* __restrict__ src, int src_kind, __nv_bfloat16* __restrict__ dst,
                               int H, int W, int ld, int n_bands) {
  extern __shared__ __align__(128) uint8_t s_raw[];  // [2 x (3 P W) pixel bytes | 2 x gw staged rows | 2 mbarriers]
  const int gw = W / P, gh = H / P;
  const size_t plane = (size_t)H * W;
  const uint32_t in_bytes = 3u * P * (uint32_t)W;
  const uint32_t rs = (uint32_t)ld * 2u + 16u;
  const uint32_t out_bytes = (uint32_t)gw * rs;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(s_raw);
  const uint32_t so = sb + 2u * in_bytes;
  const uint32_t bar = so + 2u * out_bytes;
  const int tid = threadIdx.x;
  const int n_my = (n_bands - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue_load = [&](int k) {  // thread 0
    const int b = blockIdx.x + k * gridDim.x;
    const int f = b / gh, py = b - f * gh;
    const uint8_t* band = src + (size_t)f * 3 * plane + (size_t)py * P * W;
    const uint32_t st = sb + (uint32_t)(k & 1) * in_bytes, mb = bar + 8u * (k & 1);
    mbar_arrive_expect_tx(mb, in_bytes);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       st + (uint32_t)c * P * W),
                   "l"(band + (size_t)c * plane), "r"((uint32_t)(P * W)), "r"(mb)
                   : "memory");
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  // zero both output stages once: the pad columns stay zero, everything else is overwritten per band
  for (uint32_t i = tid; i < 2u * out_bytes / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(so + 16u * i), "r"(0u) : "memory");
  __syncthreads();
  if (tid == 0 && n_my > 0) issue_load(0);
  const bool wrap = src_kind == VMC_SRC_U8_WRAP;
  const int items = gw * 3 * P;
  for (int k = 0; k < n_my; ++k) {
    const int s = k & 1;
    // output stage s was handed to the bulk stores of band k - 2: wait until they have READ it
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();  // ... and every thread is done reading input stage s ^ 1 (band k - 1)
    if (tid == 0 && k + 1 < n_my) issue_load(k + 1);
    mbar_wait(bar + 8u * s, ((uint32_t)k >> 1) & 1u);
    const uint32_t ib = sb + (uint32_t)s * in_bytes, ob = so + (uint32_t)s * out_bytes;
    for (int q = tid; q < items; q += blockDim.x) {
      const int l = q / gw, px = q - l * gw;  // patches fastest
      const int c = l / P;
      const float K = __uint_as_float(c_fastK[c]), B = __uint_as_float(c_fastB[c]);
      const uint32_t ia = ib + (uint32_t)(l * W + px * P);
      const uint32_t oa = ob + (uint32_t)px * rs + (uint32_t)(l * P) * 2u;
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        uint32_t w;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w) : "r"(ia + 2u * i));
        if (wrap) w = neg4_u8(w) & 0xFFFFu;
        const uint32_t o = pack_bf16x2(__fmaf_rn(byte_to_float<0>(w), K, B), __fmaf_rn(byte_to_float<1>(w), K, B));
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(oa + 4u * i), "r"(o) : "memory");
      }
    }
    fence_proxy_async_smem();  // the staged rows are read by the async proxy
    __syncthreads();
    if (tid == 0) {
      const int b = blockIdx.x + k * gridDim.x;  // band b = (frame, patch row): its gw patch rows are rows b * gw ..
      __nv_bfloat16* drow = dst + (size_t)b * gw * ld;
      for (int px = 0; px < gw; ++px)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(drow + (size_t)px * ld),
                     "r"(ob + (uint32_t)px * rs), "r"((uint32_t)ld * 2u)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Frame difference -> student patch matrix, gather form: work item = (patch px, line iy, 8-pixel group h);
// the grey |difference| of the 8 pixels is computed once from 2 x 24 BGR bytes and emitted for the three
// identical channels (each a 16-byte chunk; a warp writes 512 contiguous bytes per channel).
template <int P, int ITERS>
__global__ void __launch_bounds__(128)
frame_diff_gather_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ diff_u8,
                         __nv_bfloat16* __restrict__ dst, int T, int H, int W, int ld) {
  constexpr int PP = P * P;
  constexpr int CPC = PP / 8;
  constexpr int CPL = P / 8;
  const int gw = W / P, gh = H / P;
  const int fo = blockIdx.x / gh, py = blockIdx.x - fo * gh;  // output frame = clip * T + t
  const int clip = fo / T, t = fo - clip * T;
  const size_t frame_bytes = (size_t)H * W * 3;
  const uint8_t* s0 = bgr + ((size_t)clip * (T + 1) + t) * frame_bytes + (size_t)py * P * W * 3;
  const int items = gw * CPC;
  __nv_bfloat16* drow = dst + ((size_t)fo * gh * gw + (size_t)py * gw) * ld;

  uint2 a[ITERS][3], b[ITERS][3];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    if (q < items) {
      const int px = q / CPC, r = q - px * CPC;
      const int iy = r / CPL, h = r - iy * CPL;
      const uint2* p0 = reinterpret_cast<const uint2*>(s0 + ((size_t)iy * W + px * P + h * 8) * 3);
      const uint2* p1 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p0) + frame_bytes);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a[it][i] = __ldg(p0 + i);
        b[it][i] = __ldg(p1 + i);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    if (q >= items) break;
    const int px = q / CPC, r = q - px * CPC;
    const int iy = r / CPL, h = r - iy * CPL;
    const uint32_t w0[6] = {a[it][0].x, a[it][0].y, a[it][1].x, a[it][1].y, a[it][2].x, a[it][2].y};
    const uint32_t w1[6] = {b[it][0].x, b[it][0].y, b[it][1].x, b[it][1].y, b[it][2].x, b[it][2].y};
    uint32_t d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t c0[3], c1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int byte = 3 * i + k;
        c0[k] = (w0[byte >> 2] >> (8 * (byte & 3))) & 255u;
        c1[k] = (w1[byte >> 2] >> (8 * (byte & 3))) & 255u;
      }
      const int g0 = (int)bgr_gray(c0[0], c0[1], c0[2]);
      const int g1 = (int)bgr_gray(c1[0], c1[1], c1[2]);
      d[i] = (uint32_t)(g1 > g0 ? g1 - g0 : g0 - g1);  // cv2.absdiff(curr, prev)
    }
    uint2 wd;
    wd.x = d[0] | (d[1] << 8) | (d[2] << 16) | (d[3] << 24);
    wd.y = d[4] | (d[5] << 8) | (d[6] << 16) | (d[7] << 24);
    if (diff_u8 != nullptr) {
      *reinterpret_cast<uint2*>(diff_u8 + ((size_t)fo * H + (size_t)py * P + iy) * W + px * P + h * 8) = wd;
    }
    uint4* orow = reinterpret_cast<uint4*>(drow + (size_t)px * ld);
#pragma unroll
    for (int c = 0; c < 3; ++c) orow[c * CPC + r] = normalise_pack8_fast(neg4_u8(wd.x), neg4_u8(wd.y), c);  // regime A wrap
  }
}

// bf16 "split" operand for near-fp32 GEMMs on the bf16 tensor cores: x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi).  Rows are written as [hi | lo | hi] (3*d columns) and multiplied against weights
// packed as [Whi | Whi | Wlo], i.e. x W^T ~= hi Whi^T + lo Whi^T + hi Wlo^T in ONE K = 3d GEMM.
__device__ __forceinline__ void store_split4(__nv_bfloat16* row, int d, int j, float4 o) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(o.x, o.y), h23 = __floats2bfloat162_rn(o.z, o.w);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(o.x - __low2float(h01), o.y - __high2float(h01));
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(o.z - __low2float(h23), o.w - __high2float(h23));
  uint2 hi, lo;
  hi.x = *reinterpret_cast<const uint32_t*>(&h01);
  hi.y = *reinterpret_cast<const uint32_t*>(&h23);
  lo.x = *reinterpret_cast<const uint32_t*>(&l01);
  lo.y = *reinterpret_cast<const uint32_t*>(&l23);
  reinterpret_cast<uint2*>(row)[j] = hi;
  reinterpret_cast<uint2*>(row + d)[j] = lo;
  reinterpret_cast<uint2*>(row + 2 * d)[j] = hi;
}

// ---------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row cached in registers, fp32 two-pass statistics.
// ---------------------------------------------------------------------------------------
// XB16: the input rows are bf16 (the bf16 residual stream of the ViT tower; cls_every must be 0).
// stats_rounded: stats_out holds the statistics of the bf16-ROUNDED output row (what a GEMM reading y16 multiplies).
template <int NV, bool XB16>  // float4 per lane; supports d <= NV * 128
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x_, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float* y32, long long ld32,
                 __nv_bfloat16* y16, long long ld16, int y16_split, int rows, int d,
                 const float* __restrict__ cls_row, int cls_every, float2* __restrict__ stats_out, int stats_rounded) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = XB16 ? nullptr : reinterpret_cast<const float*>(x_) + (size_t)row * ldx;
  const __nv_bfloat16* xr16 = XB16 ? reinterpret_cast<const __nv_bfloat16*>(x_) + (size_t)row * ldx : nullptr;
  if (!XB16 && cls_every > 0 && (row % cls_every) == 0) xr = cls_row;
  const int nv = d >> 2;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      if constexpr (XB16) {
        const uint2 u = reinterpret_cast<const uint2*>(xr16)[j];
        v[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                           __uint_as_float(u.y & 0xFFFF0000u));
      } else {
        v[i] = reinterpret_cast<const float4*>(xr)[j];
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  float so = 0.f, qo = 0.f;  // (sum, sum of squares) of the OUTPUT row, for a LayerNorm folded into the next GEMM
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + j);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (stats_rounded) {
        const float r0 = __bfloat162float(__float2bfloat16_rn(o.x)), r1 = __bfloat162float(__float2bfloat16_rn(o.y));
        const float r2 = __bfloat162float(__float2bfloat16_rn(o.z)), r3 = __bfloat162float(__float2bfloat16_rn(o.w));
        so += (r0 + r1) + (r2 + r3);
        qo += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
      } else {
        so += (o.x + o.y) + (o.z + o.w);
        qo += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
      }
      if (y32 != nullptr) reinterpret_cast<float4*>(y32 + (size_t)row * ld32)[j] = o;
      if (y16 != nullptr) {
        if (y16_split) {
          store_split4(y16 + (size_t)row * ld16, d, j, o);
        } else {
          uint2 pk;
          pk.x = pack_bf16x2(o.x, o.y);
          pk.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(y16 + (size_t)row * ld16)[j] = pk;
        }
      }
    }
  }
  if (stats_out != nullptr) {
    so = warp_sum(so);
    qo = warp_sum(qo);
    if (lane == 0) stats_out[row] = make_float2(so, qo);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                 long long ldy, int rows, int d, int split) {
  const int nv = d >> 2;
  const size_t total = (size_t)rows * nv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int j = idx % nv;
    const size_t r = idx / nv;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + j);
    if (split) {
      store_split4(y + r * ldy, d, j, v);
    } else {
      uint2 pk;
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(y + r * ldy)[j] = pk;
    }
  }
}

// mean over T rows: thread per (b, 4 columns)
__global__ void __launch_bounds__(256)
mean_rows_kernel(const float* __restrict__ x, float* y32, __nv_bfloat16* y16, int B, int T, int d) {
  const int nv = d >> 2;
  const size_t total = (size_t)B * nv;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = idx % nv;
  const size_t b = idx / nv;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (b * T + t) * d) + j);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float inv = 1.0f / (float)T;
  s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
  if (y32 != nullptr) reinterpret_cast<float4*>(y32 + b * d)[j] = s;
  if (y16 != nullptr) {
    uint2 pk;
    pk.x = pack_bf16x2(s.x, s.y);
    pk.y = pack_bf16x2(s.z, s.w);
    reinterpret_cast<uint2*>(y16 + b * d)[j] = pk;
  }
}

// losses.py:27-40 -- one warp per row, atomicAdd of (1 - cos) / rows
__global__ void __launch_bounds__(256)
cosine_loss_kernel(const float* __restrict__ s, const float* __restrict__ t, int rows, int d,
                   float* out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float ss = 0.f, tt = 0.f, st = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float a = s[(size_t)row * d + j], b = t[(size_t)row * d + j];
    ss += a * a;
    tt += b * b;
    st += a * b;
  }
  ss = warp_sum(ss);
  tt = warp_sum(tt);
  st = warp_sum(st);
  if (lane == 0) {
    const float eps = 1e-5f;
    const float ns = fmaxf(sqrtf(ss), eps), nt = fmaxf(sqrtf(tt), eps);
    float c = st / (ns * nt);
    c = fminf(fmaxf(c, -1.0f + eps), 1.0f - eps);
    atomicAdd(out, (1.0f - c) / (float)rows);
  }
}

// ---------------------------------------------------------------------------------------
// Pillow bicubic resize + centre crop for non-224x224 frames (Resize(224, BICUBIC) -> CenterCrop(224)
// of clip's _transform, reached from models/student_model.py:77-78).  Pillow's 8-bit resampler
// (src/libImaging/Resample.c): 22-bit fixed-point coefficients, horizontal pass then vertical pass,
// each accumulated from 2^21 and clipped to uint8.  Integer arithmetic: bit-exact.
// The to_pil_image wrap of the student is applied to the source pixel on load (it precedes the resize).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_wrapped(const void* src, int src_kind, size_t idx) {
  if (src_kind == VMC_SRC_F32_WRAP) return wrap_f32(reinterpret_cast<const float*>(src)[idx]);
  const uint32_t u = reinterpret_cast<const uint8_t*>(src)[idx];
  return src_kind == VMC_SRC_U8_WRAP ? ((0u - u) & 255u) : u;
}
__device__ __forceinline__ uint8_t clip8_fixed(int acc) {
  const int v = acc >> 22;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[f][c][y][xo] = horizontal pass at resized column (left + xo)
__global__ void __launch_bounds__(256)
resize_h_kernel(const void* __restrict__ src, int src_kind, uint8_t* __restrict__ tmp, int planes, int H,
                int W, int size, int left, const int* __restrict__ bounds, const int* __restrict__ coef,
                int ksize) {
  const size_t total = (size_t)planes * H * size;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % size);
    const size_t row = idx / size;  // plane * H + y
    const int xr = left + xo;
    const int xmin = bounds[2 * xr], n = bounds[2 * xr + 1];
    const int* k = coef + (size_t)xr * ksize;
    int acc = 1 << 21;
    for (int i = 0; i < n; ++i) acc += (int)load_wrapped(src, src_kind, row * W + xmin + i) * k[i];
    tmp[idx] = clip8_fixed(acc);
  }
}
// out[f][c][yo][xo] = vertical pass at resized row (top + yo) over tmp [planes, H, size]
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out, int planes, int H, int size,
                int top, const int* __restrict__ bounds, const int* __restrict__ coef, int ksize) {
  const size_t total = (size_t)planes * size * size;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % size);
    const int yo = (int)((idx / size) % size);
    const size_t plane = idx / ((size_t)size * size);
    const int yr = top + yo;
    const int ymin = bounds[2 * yr], n = bounds[2 * yr + 1];
    const int* k = coef + (size_t)yr * ksize;
    const uint8_t* col = tmp + (plane * H + ymin) * size + xo;
    int acc = 1 << 21;
    for (int i = 0; i < n; ++i) acc += (int)col[(size_t)i * size] * k[i];
    out[idx] = clip8_fixed(acc);
  }
}

}  // namespace
int vmc_get_option(int option);
namespace {

int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)vmc_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int check_geometry(const char* who, int H, int W, int patch, int ld_patch, int dst_kind) {
  VMC_CHECK_ARG(H > 0 && W > 0 && (W % 16) == 0, VMC_ERR_SHAPE,
                "%s: W must be a positive multiple of 16 (H=%d W=%d)", who, H, W);
  VMC_CHECK_ARG(dst_kind >= VMC_DST_U8 && dst_kind <= VMC_DST_BF16_PATCH, VMC_ERR_ARG,
                "%s: unknown dst_kind %d", who, dst_kind);
  if (dst_kind == VMC_DST_BF16_PATCH) {
    VMC_CHECK_ARG(patch > 0 && (H % patch) == 0 && (W % patch) == 0, VMC_ERR_SHAPE,
                  "%s: H,W must be multiples of patch %d (no resize in this path: the reference's "
                  "Resize/CenterCrop is the identity only at the model resolution)",
                  who, patch);
    VMC_CHECK_ARG(ld_patch >= 3 * patch * patch && (ld_patch % 8) == 0, VMC_ERR_ALIGN,
                  "%s: ld_patch must be >= 3*p*p and a multiple of 8", who);
  }
  return VMC_OK;
}

}  // namespace

extern "C" {

int vmc_prologue(const void* frames, int src_kind, void* dst, int dst_kind, int F, int H, int W,
                 int patch, int ld_patch, void* stream) {
  VMC_CHECK_ARG(frames && dst, VMC_ERR_ARG, "vmc_prologue: null pointer");
  VMC_CHECK_ARG(src_kind >= VMC_SRC_U8 && src_kind <= VMC_SRC_F32_NORM, VMC_ERR_ARG,
                "vmc_prologue: unknown src_kind %d", src_kind);
  VMC_CHECK_ARG(src_kind != VMC_SRC_F32_NORM || dst_kind == VMC_DST_BF16_PATCH, VMC_ERR_ARG,
                "vmc_prologue: VMC_SRC_F32_NORM only feeds VMC_DST_BF16_PATCH");
  VMC_CHECK_ARG(F > 0, VMC_ERR_SHAPE, "vmc_prologue: F must be positive");
  VMC_TRY(check_geometry("vmc_prologue", H, W, patch, ld_patch, dst_kind));
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(frames) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                VMC_ERR_ALIGN, "vmc_prologue: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)F * 3 * H * (W / 16);
  VMC_CHECK_ARG(total < (1ull << 31), VMC_ERR_SHAPE, "vmc_prologue: too many frames in one call (F=%d)", F);
  const int impl = vmc_get_option(VMC_OPT_PROLOGUE_IMPL);
  const bool f32_src = src_kind == VMC_SRC_F32_WRAP || src_kind == VMC_SRC_F32_NORM;
  // patch sizes other than 16 / 32 (ViT-L/14): element-wise gather kernel, writes the pad columns itself
  const bool gather_any = dst_kind == VMC_DST_BF16_PATCH && impl != 1 && impl != 2 && patch != 16 && patch != 32 &&
                          (ld_patch % 8) == 0 && ld_patch >= 3 * patch * patch;
  if (dst_kind == VMC_DST_BF16_PATCH && ld_patch != 3 * patch * patch && !gather_any) {
    const size_t n = (size_t)F * (H / patch) * (W / patch);
    VMC_CUDA(cudaMemsetAsync(dst, 0, n * ld_patch * 2, st));
  }
  if (gather_any) {
    const double px = (double)F * 3 * H * W;
    const int bands = F * (H / patch);
    __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dst);
    {
      VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * ((f32_src ? 4.0 : 1.0) + 2.0));
      const size_t smem14 = 2 * ((size_t)3 * 14 * W + (size_t)(W / 14) * (ld_patch * 2 + 16)) + 16;
      if (!f32_src && patch == 14 && (W % 16) == 0 && (((size_t)H * W) % 16) == 0 && smem14 <= 100 * 1024) {
        auto k14 = prologue_band_gather_u8_kernel<14>;
        VMC_CUDA(cudaFuncSetAttribute(k14, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem14));
        const int per_sm = (int)((220 * 1024) / (smem14 + 1024));
        const int grid14 = bands < vmc_num_sms() * per_sm ? bands : vmc_num_sms() * per_sm;
        k14<<<grid14, 256, smem14, st>>>(reinterpret_cast<const uint8_t*>(frames), src_kind, d16, H, W, ld_patch, bands);
      }
      else if (f32_src) prologue_gather_any_kernel<true, 5><<<bands, 256, 0, st>>>(frames, src_kind, d16, H, W, patch, ld_patch);
      else prologue_gather_any_kernel<false, 5><<<bands, 256, 0, st>>>(frames, src_kind, d16, H, W, patch, ld_patch);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  // gather kernel: 7 work items per thread, block = items / 7 (192 threads at p = 16, 384 at p = 32, W = 224)
  const int g_items = dst_kind == VMC_DST_BF16_PATCH ? (W / patch) * (3 * patch * patch / 8) : 0;
  const int g_block = ((g_items + 6) / 7 + 31) / 32 * 32;
  if (dst_kind == VMC_DST_BF16_PATCH && impl != 1 && impl != 2 && (patch == 16 || patch == 32) && (W % 8) == 0 &&
      g_block <= 384) {
    const double px = (double)F * 3 * H * W;
    const int bands = F * (H / patch);
    __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dst);
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * ((f32_src ? 4.0 : 1.0) + 2.0));
    // uint8 sources: 16-pixel items -- opt-in only (VMC_OPT_PROLOGUE_IMPL = 4): measured SLOWER inside the step (1.07-1.12 vs 0.82 ms)
    const int g16_items = g_items / 2;
    const int g16_block = ((g16_items + 6) / 7 + 31) / 32 * 32;
    if (!f32_src && impl == 4 && (W % 16) == 0 && (((size_t)H * W) % 16) == 0 && g16_block <= 384) {
      const uint8_t* s8 = reinterpret_cast<const uint8_t*>(frames);
      if (patch == 16) prologue_gather16_kernel<16, 7><<<bands, g16_block, 0, st>>>(s8, src_kind, d16, H, W, ld_patch);
      else prologue_gather16_kernel<32, 7><<<bands, g16_block, 0, st>>>(s8, src_kind, d16, H, W, ld_patch);
    } else if (patch == 16 && !f32_src)
      prologue_gather_kernel<16, 7, false><<<bands, g_block, 0, st>>>(frames, src_kind, d16, H, W, ld_patch);
    else if (patch == 16)
      prologue_gather_kernel<16, 7, true><<<bands, g_block, 0, st>>>(frames, src_kind, d16, H, W, ld_patch);
    else if (!f32_src)
      prologue_gather_kernel<32, 7, false><<<bands, g_block, 0, st>>>(frames, src_kind, d16, H, W, ld_patch);
    else
      prologue_gather_kernel<32, 7, true><<<bands, g_block, 0, st>>>(frames, src_kind, d16, H, W, ld_patch);
  } else if (dst_kind == VMC_DST_BF16_PATCH && impl == 2) {
    const double px = (double)F * 3 * H * W;
    const double in_b = f32_src ? 4.0 : 1.0;
    const int bands = F * 3 * (H / patch);
    const int grid = bands < vmc_num_sms() * 8 ? bands : vmc_num_sms() * 8;
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * (in_b + 2.0));
    prologue_patch_kernel<<<grid, 256, (size_t)patch * W * 2, st>>>(
        frames, src_kind, reinterpret_cast<__nv_bfloat16*>(dst), F, H, W, patch, ld_patch);
  } else {
    const double px = (double)F * 3 * H * W;
    const double in_b = (src_kind == VMC_SRC_F32_WRAP || src_kind == VMC_SRC_F32_NORM) ? 4.0 : 1.0;
    const double out_b = dst_kind == VMC_DST_U8 ? 1.0 : (dst_kind == VMC_DST_F32_NCHW ? 4.0 : 2.0);
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * (in_b + out_b));
    prologue_kernel<4><<<grid_for((total + 3) / 4, 256), 256, 0, st>>>(frames, src_kind, dst, dst_kind,
                                                                       F, H, W, patch, ld_patch);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_frame_diff_prologue(const uint8_t* bgr, uint8_t* diff_u8, void* dst, int dst_kind,
                            int clips, int T, int H, int W, int patch, int ld_patch, void* stream) {
  VMC_CHECK_ARG(bgr && (diff_u8 || dst), VMC_ERR_ARG, "vmc_frame_diff_prologue: null pointer");
  VMC_CHECK_ARG(clips > 0 && T > 0, VMC_ERR_SHAPE, "vmc_frame_diff_prologue: empty input");
  if (dst) VMC_TRY(check_geometry("vmc_frame_diff_prologue", H, W, patch, ld_patch, dst_kind));
  VMC_CHECK_ARG(H > 0 && W > 0 && (W % 16) == 0, VMC_ERR_SHAPE,
                "vmc_frame_diff_prologue: W must be a positive multiple of 16");
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(bgr) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(diff_u8) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                VMC_ERR_ALIGN, "vmc_frame_diff_prologue: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dst && dst_kind == VMC_DST_BF16_PATCH && ld_patch != 3 * patch * patch) {
    const size_t n = (size_t)clips * T * (H / patch) * (W / patch);
    VMC_CUDA(cudaMemsetAsync(dst, 0, n * ld_patch * 2, st));
  }
  const size_t total = (size_t)clips * T * H * (W / 16);
  VMC_CHECK_ARG(total < (1ull << 31), VMC_ERR_SHAPE, "vmc_frame_diff_prologue: too many frames in one call");
  const int impl = vmc_get_option(VMC_OPT_PROLOGUE_IMPL);
  const int g_items = (dst && dst_kind == VMC_DST_BF16_PATCH) ? (W / patch) * (patch * patch / 8) : 0;
  const int g_block = ((g_items + 6) / 7 + 31) / 32 * 32;
  if (dst && dst_kind == VMC_DST_BF16_PATCH && impl != 1 && impl != 2 && (patch == 16 || patch == 32) &&
      (W % 8) == 0 && g_block <= 128) {
    const double px = (double)clips * T * H * W;
    const int bands = clips * T * (H / patch);
    __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dst);
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * (6.0 + (diff_u8 ? 1.0 : 0.0) + 6.0));
    if (patch == 16)
      frame_diff_gather_kernel<16, 7><<<bands, g_block, 0, st>>>(bgr, diff_u8, d16, T, H, W, ld_patch);
    else
      frame_diff_gather_kernel<32, 7><<<bands, g_block, 0, st>>>(bgr, diff_u8, d16, T, H, W, ld_patch);
  } else if (dst && dst_kind == VMC_DST_BF16_PATCH && impl == 2) {
    const double px = (double)clips * T * H * W;
    const int bands = clips * T * (H / patch);
    const int grid = bands < vmc_num_sms() * 8 ? bands : vmc_num_sms() * 8;
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * (6.0 + (diff_u8 ? 1.0 : 0.0) + 6.0));
    frame_diff_patch_kernel<<<grid, 256, (size_t)patch * W * 3, st>>>(
        bgr, diff_u8, reinterpret_cast<__nv_bfloat16*>(dst), clips, T, H, W, patch, ld_patch);
  } else {
    const double px = (double)clips * T * H * W;
    const double out_b = !dst ? 0.0 : (dst_kind == VMC_DST_U8 ? 3.0 : (dst_kind == VMC_DST_F32_NCHW ? 12.0 : 6.0));
    // each BGR frame is read twice (as "previous" and as "current") except at clip ends
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, px * (6.0 + (diff_u8 ? 1.0 : 0.0) + out_b));
    frame_diff_kernel<<<grid_for(total, 256), 256, 0, st>>>(bgr, diff_u8, dst, dst_kind, clips, T, H,
                                                            W, patch, ld_patch);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_layernorm(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                  float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows,
                  int d, const float* cls_row, int cls_every, void* stream) {
  return vmc_layernorm_stats(x, ldx, gamma, beta, eps, y32, ld32, y16, ld16, y16_split, rows, d, cls_row,
                             cls_every, nullptr, stream);
}

int vmc_layernorm_stats(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                        float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows,
                        int d, const float* cls_row, int cls_every, float* stats_out, void* stream) {
  return vmc_layernorm_ex(x, 0, ldx, gamma, beta, eps, y32, ld32, y16, ld16, y16_split, rows, d, cls_row, cls_every,
                          stats_out, 0, stream);
}

int vmc_layernorm_ex(const void* x, int x_bf16, long long ldx, const float* gamma, const float* beta, float eps,
                     float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows, int d,
                     const float* cls_row, int cls_every, float* stats_out, int stats_rounded, void* stream) {
  VMC_CHECK_ARG(x && gamma && beta && (y32 || y16), VMC_ERR_ARG, "vmc_layernorm: null pointer");
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(stats_out) & 7) == 0, VMC_ERR_ALIGN,
                "vmc_layernorm_stats: stats_out must be 8-byte aligned");
  VMC_CHECK_ARG(rows > 0 && d > 0 && (d % 4) == 0 && d <= 4096, VMC_ERR_SHAPE,
                "vmc_layernorm: need d %% 4 == 0 and d <= 4096 (rows=%d d=%d)", rows, d);
  VMC_CHECK_ARG((ldx % 4) == 0 && (!y32 || (ld32 % 4) == 0) && (!y16 || (ld16 % 4) == 0),
                VMC_ERR_ALIGN, "vmc_layernorm: row strides must be multiples of 4 elements");
  VMC_CHECK_ARG(cls_every <= 0 || (cls_row != nullptr && !x_bf16), VMC_ERR_ARG,
                "vmc_layernorm: cls_every needs cls_row and an fp32 input");
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & (x_bf16 ? 7 : 15)) == 0, VMC_ERR_ALIGN,
                "vmc_layernorm: input rows must be 16-byte (fp32) / 8-byte (bf16) aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  __nv_bfloat16* y16b = reinterpret_cast<__nv_bfloat16*>(y16);
#define LN_LAUNCH(NV)                                                                                              \
  do {                                                                                                             \
    if (x_bf16)                                                                                                    \
      layernorm_kernel<NV, true><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, eps, y32, ld32, y16b, ld16, y16_split, \
                                                       rows, d, cls_row, cls_every,                                \
                                                       reinterpret_cast<float2*>(stats_out), stats_rounded);       \
    else                                                                                                           \
      layernorm_kernel<NV, false><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, eps, y32, ld32, y16b, ld16, y16_split, \
                                                        rows, d, cls_row, cls_every,                               \
                                                        reinterpret_cast<float2*>(stats_out), stats_rounded);      \
  } while (0)
  VmcProfScope prof(VMC_K_LAYERNORM, st, 0.0,
                    (double)rows * d * ((x_bf16 ? 2.0 : 4.0) + (y32 ? 4.0 : 0.0) + (y16 ? 2.0 : 0.0)));
  if (d <= 512) LN_LAUNCH(4);
  else if (d <= 768) LN_LAUNCH(6);
  else if (d <= 1024) LN_LAUNCH(8);
  else if (d <= 2048) LN_LAUNCH(16);
  else LN_LAUNCH(32);
#undef LN_LAUNCH
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_cast_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int d,
                  int split, void* stream) {
  VMC_CHECK_ARG(x && y, VMC_ERR_ARG, "vmc_cast_bf16: null pointer");
  VMC_CHECK_ARG(rows > 0 && d > 0 && (d % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0,
                VMC_ERR_SHAPE, "vmc_cast_bf16: d and strides must be multiples of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)rows * (d / 4);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  cast_bf16_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(y),
                                                         ldy, rows, d, split);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_mean_rows(const float* x, float* y32, void* y16, int B, int T, int d, void* stream) {
  VMC_CHECK_ARG(x && (y32 || y16), VMC_ERR_ARG, "vmc_mean_rows: null pointer");
  VMC_CHECK_ARG(B > 0 && T > 0 && d > 0 && (d % 4) == 0, VMC_ERR_SHAPE,
                "vmc_mean_rows: bad shape B=%d T=%d d=%d", B, T, d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)B * (d / 4);
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  mean_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      x, y32, reinterpret_cast<__nv_bfloat16*>(y16), B, T, d);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_cosine_distill_loss(const float* s, const float* t, int rows, int d, float* out,
                            void* stream) {
  VMC_CHECK_ARG(s && t && out, VMC_ERR_ARG, "vmc_cosine_distill_loss: null pointer");
  VMC_CHECK_ARG(rows > 0 && d > 0, VMC_ERR_SHAPE, "vmc_cosine_distill_loss: empty input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VMC_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  VmcProfScope prof(VMC_K_OTHER, st, 0.0, 0.0);
  cosine_loss_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s, t, rows, d, out);
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}


// ---- Pillow coefficient tables (host, double precision exactly as Resample.c) ----
}  // extern "C"
#include <math.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
namespace {
double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
struct ResizeTable {
  int* bounds = nullptr;  // device [out, 2]
  int* coef = nullptr;    // device [out, ksize]
  int ksize = 0;
};
int build_table(int in_size, int out_size, ResizeTable* t) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<int> bounds(2 * (size_t)out_size), kk((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  VMC_CUDA(cudaMalloc(&t->bounds, bounds.size() * sizeof(int)));
  VMC_CUDA(cudaMalloc(&t->coef, kk.size() * sizeof(int)));
  VMC_CUDA(cudaMemcpy(t->bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
  VMC_CUDA(cudaMemcpy(t->coef, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
  t->ksize = ksize;
  return VMC_OK;
}
// tables are owned by the library, one per (device, in, out), built on first use
int get_table(int in_size, int out_size, ResizeTable* out) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, ResizeTable> cache;
  int dev = 0;
  VMC_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(dev, in_size, out_size);
  auto it = cache.find(key);
  if (it == cache.end()) {
    ResizeTable t;
    VMC_TRY(build_table(in_size, out_size, &t));
    it = cache.emplace(key, t).first;
  }
  *out = it->second;
  return VMC_OK;
}
}  // namespace
extern "C" {

int vmc_resize_geometry(int H, int W, int size, int* new_h, int* new_w, int* top, int* left) {
  return vmc_resize_geometry_ex(H, W, size, 0, new_h, new_w, top, left);
}

int vmc_resize_geometry_ex(int H, int W, int size, int crop_floor, int* new_h, int* new_w, int* top, int* left) {
  VMC_CHECK_ARG(H > 0 && W > 0 && size > 0, VMC_ERR_SHAPE, "vmc_resize_geometry: bad geometry");
  // torchvision Resize(int) / HF get_resize_output_image_size: short side -> size, long side -> int(size * long / short)
  // (frames smaller than `size` are scaled UP, so the centre crop never has to pad)
  int nh, nw;
  if (W <= H) {
    nw = size;
    nh = (int)((double)size * H / W);
  } else {
    nh = size;
    nw = (int)((double)size * W / H);
  }
  // torchvision CenterCrop: int(round((dim - size) / 2.0)) with Python's round-half-to-even;
  // HF image_transforms.center_crop (extract_embeddings.py:91): (dim - size) // 2
  auto crop = [&](int dim) { return crop_floor ? (dim - size) / 2 : (int)nearbyint((dim - size) / 2.0); };
  if (new_h) *new_h = nh;
  if (new_w) *new_w = nw;
  if (top) *top = crop(nh);
  if (left) *left = crop(nw);
  return VMC_OK;
}

int vmc_resize_center_crop(const void* frames, int src_kind, uint8_t* out, uint8_t* tmp, int F, int H,
                           int W, int size, void* stream) {
  return vmc_resize_center_crop_ex(frames, src_kind, out, tmp, F, H, W, size, 0, stream);
}

int vmc_resize_center_crop_ex(const void* frames, int src_kind, uint8_t* out, uint8_t* tmp, int F, int H,
                              int W, int size, int crop_floor, void* stream) {
  VMC_CHECK_ARG(frames && out && tmp, VMC_ERR_ARG, "vmc_resize_center_crop: null pointer");
  VMC_CHECK_ARG(src_kind >= VMC_SRC_U8 && src_kind <= VMC_SRC_F32_WRAP, VMC_ERR_ARG,
                "vmc_resize_center_crop: unknown src_kind %d", src_kind);
  VMC_CHECK_ARG(F > 0 && H > 0 && W > 0 && size > 0, VMC_ERR_SHAPE, "vmc_resize_center_crop: bad geometry %dx%d -> %d", H, W, size);
  int nh, nw, top, left;
  VMC_TRY(vmc_resize_geometry_ex(H, W, size, crop_floor, &nh, &nw, &top, &left));
  ResizeTable th, tv;
  VMC_TRY(get_table(W, nw, &th));
  VMC_TRY(get_table(H, nh, &tv));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int planes = F * 3;
  {
    const size_t total = (size_t)planes * H * size;
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, (double)planes * H * (W + size));
    resize_h_kernel<<<grid_for(total, 256), 256, 0, st>>>(frames, src_kind, tmp, planes, H, W, size, left,
                                                          th.bounds, th.coef, th.ksize);
  }
  VMC_LAUNCH_CHECK();
  {
    const size_t total = (size_t)planes * size * size;
    VmcProfScope prof(VMC_K_PROLOGUE, st, 0.0, (double)planes * size * (H + size));
    resize_v_kernel<<<grid_for(total, 256), 256, 0, st>>>(tmp, out, planes, H, size, top, tv.bounds, tv.coef,
                                                          tv.ksize);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch(2);
  return VMC_OK;
}
}  // extern "C"
