// A1: ViT self-attention on tcgen05 (one CTA per (frame, head)), and the small
// masked attention of the TFAM block (SIMT, fp32).
//
// ViT kernel (head_dim 64, no mask, L <= 272 tokens):
//   TMA loads Q, K, V head slices of the packed qkv buffer ([F, L, 3d] bf16, 128B swizzle;
//   rows >= L are zero-filled by the tensor map).  Per 128-row query tile:
//     S = Q K^T          tcgen05.mma 128 x Lk16 x 16 (x4), fp32 in TMEM
//     softmax            one thread per row: tcgen05.ld of its own TMEM lane, so the row
//                        max / sum are thread-local (no shuffles); P (bf16) is written to
//                        shared memory in the K-major 128B-swizzled UMMA layout
//     O = P V            tcgen05.mma 128 x 64 x 16, V consumed MN-major straight from the
//                        token-major TMA tile (no transpose)
//     O / rowsum -> bf16 -> global
// Replaces nn.MultiheadAttention in the OpenAI ResidualAttentionBlock (clip/model.py) and
// HF CLIPAttention reached from models/student_model.py:84 / extract_embeddings.py:94.
#include "common.cuh"
#include "vimoclip_b200.h"

namespace {

using namespace vmc;

constexpr int HD = 64;          // head dim
constexpr uint32_t TILE = 16384;  // 128 rows x 128 B

struct AttnArgs {
  int F, L, heads, d;
  int n_mt;     // ceil(L / 128)
  int lk16;     // ceil(L / 16) * 16
  int n_pkb;    // 16 KB blocks of P in smem (0 when P lives in TMEM)
  int o_col;    // TMEM column of the O accumulator
  uint32_t tmem_cols;
  __nv_bfloat16* out;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kPTmem = true : P (bf16) is written back into TMEM over the dead S columns and the PV MMA takes
//                 its A operand from TMEM (tcgen05.mma ... [d], [a], b-desc): 96 KB of smem and 256
//                 TMEM columns per CTA at L = 197, so two CTAs share an SM and overlap each other's
//                 load / MMA / softmax phases.
// kPTmem = false: P goes through shared memory in the K-major 128B-swizzled layout (first version,
//                 kept as a cross-check of the TMEM-operand path).
template <bool kPTmem>
__global__ void __launch_bounds__(128)
attention_vit_kernel(const __grid_constant__ CUtensorMap tm, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + a.n_mt * TILE;
  const uint32_t sV = sK + a.n_mt * TILE;
  const uint32_t sP = sV + a.n_mt * TILE;
  const uint32_t bar_base = sP + a.n_pkb * TILE;
  const uint32_t bar_load = bar_base, bar_mma = bar_base + 8, tmem_ptr_addr = bar_base + 16;
  uint8_t* sP_generic = smem_raw + (sP - raw_addr);
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int head = blockIdx.x % a.heads;
  const int frame = blockIdx.x / a.heads;

  if (tid == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_addr, a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, 3u * a.n_mt * TILE);
    for (int mt = 0; mt < a.n_mt; ++mt) {
      tma_load_3d(sQ + mt * TILE, &tm, bar_load, head * HD, mt * 128, frame);
      tma_load_3d(sK + mt * TILE, &tm, bar_load, a.d + head * HD, mt * 128, frame);
      tma_load_3d(sV + mt * TILE, &tm, bar_load, 2 * a.d + head * HD, mt * 128, frame);
    }
  }
  mbar_wait(bar_load, 0);

  const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const uint32_t lane_addr = uint32_t(warp * 32) << 16;
  uint32_t mma_phase = 0;
  const int n_chunks = (a.lk16 + 31) / 32;

  for (int mt = 0; mt < a.n_mt; ++mt) {
    // ---- S = Q_mt K^T ----
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dq = umma_desc_sw128(sQ + mt * TILE);
      const int n1 = a.lk16 > 256 ? 256 : a.lk16;
      const uint32_t idesc1 = umma_idesc_bf16(128, n1, 0, 0);
      const uint64_t dk = umma_desc_sw128(sK);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_ss(tmem_base, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc1, k != 0);
      if (a.lk16 > 256) {
        const uint32_t idesc2 = umma_idesc_bf16(128, a.lk16 - 256, 0, 0);
        const uint64_t dk2 = umma_desc_sw128(sK + 256 * 128);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tmem_base + 256, dq + uint64_t(2 * k), dk2 + uint64_t(2 * k), idesc2, k != 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    // ---- softmax over this thread's row ----
    const int row = mt * 128 + tid;
    const bool warp_active = (mt * 128 + warp * 32) < a.L;  // warp-uniform
    float row_sum = 1.0f;
    if (warp_active) {
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + uint32_t(c * 32), r);
        tmem_ld_wait();
        const int c0 = c * 32;
        if (c0 + 32 <= a.L) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < a.L) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
      const float mxs = mx * sc;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + uint32_t(c * 32), r);
        tmem_ld_wait();
        const int c0 = c * 32;
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(r[j]), sc, -mxs));
          p[j] = (c0 + j < a.L) ? e : 0.f;
        }
        // round to bf16 first so the normaliser matches what the PV MMA actually sums
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
          sum += __low2float(h2) + __high2float(h2);
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if constexpr (kPTmem) {
          // P[row, 32c .. 32c+31] as 16 packed columns at TMEM cols [16c, 16c+16): these alias S
          // columns this thread has already consumed (chunk c/2 <= c), lane-private.
          uint32_t lo8[8], hi8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            lo8[j] = pk[j];
            hi8[j] = pk[8 + j];
          }
          tmem_st_32x32b_x8(tmem_base + lane_addr + uint32_t(c * 16), lo8);
          tmem_st_32x32b_x8(tmem_base + lane_addr + uint32_t(c * 16 + 8), hi8);
        } else {
          // K-major SW128 layout: block kb = c/2 (64 columns), row tid, 16-byte chunk ((c&1)*4 + i)
          uint8_t* prow = sP_generic + (c >> 1) * TILE + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int chunk = ((c & 1) * 4 + i) ^ (tid & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) =
                make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        }
      }
      row_sum = sum;
    }
    if constexpr (kPTmem) {
      tmem_st_wait();
    } else {
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();

    // ---- O = P V ----
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);  // B (V) is MN-major
      const int nkk = a.lk16 / 16;
      for (int kk = 0; kk < nkk; ++kk) {
        const uint64_t dv = umma_desc_sw128(sV + kk * 2048);
        if constexpr (kPTmem) {
          // A from TMEM: 16 bf16 of K = 8 packed 32-bit columns per step
          umma_ts(tmem_base + uint32_t(a.o_col), tmem_base + uint32_t(kk * 8), dv, idesc_pv, kk != 0);
        } else {
          const uint64_t dp = umma_desc_sw128(sP + (kk >> 2) * TILE) + uint64_t(2 * (kk & 3));
          umma_ss(tmem_base + uint32_t(a.o_col), dp, dv, idesc_pv, kk != 0);
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    if (warp_active) {
      const float inv = 1.0f / row_sum;
      __nv_bfloat16* orow =
          a.out + ((size_t)frame * a.L + row) * a.d + (size_t)head * HD;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + uint32_t(a.o_col + half * 32), r);
        tmem_ld_wait();
        if (row < a.L) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(orow + half * 32)[i] = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM S/O region and P smem are reused by the next query tile
  }

  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------
// TFAM attention: fp32, boolean key-padding mask (1 = attend), online softmax over 64-key tiles.
// grid = (ceil(Tq/16), heads, B); 4 warps, each owning 4 query rows.
// ---------------------------------------------------------------------------------------
constexpr int QB = 16;
constexpr int KB = 64;

__global__ void __launch_bounds__(128)
attention_masked_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
                        long long ldk, const float* __restrict__ v, long long ldv,
                        const uint8_t* __restrict__ key_valid, void* __restrict__ out_v,
                        int out_f32, long long ldo, int Tq, int Tk) {
  __shared__ float sq[QB][HD];
  __shared__ float sk[KB][HD + 1];
  __shared__ float sv[KB][HD];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* kvb = key_valid ? key_valid + (size_t)b * Tk : nullptr;

  for (int i = tid; i < QB * HD; i += 128) {
    const int r = i / HD, c = i % HD;
    sq[r][c] = (q0 + r < Tq) ? q[((size_t)b * Tq + q0 + r) * ldq + h * HD + c] * 0.125f : 0.f;
  }
  float m[4], l[4], o0[4], o1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
    o0[i] = 0.f;
    o1[i] = 0.f;
  }
  for (int k0 = 0; k0 < Tk; k0 += KB) {
    __syncthreads();
    for (int i = tid; i < KB * HD; i += 128) {
      const int r = i / HD, c = i % HD;
      const bool ok = k0 + r < Tk;
      sk[r][c] = ok ? k[((size_t)b * Tk + k0 + r) * ldk + h * HD + c] : 0.f;
      sv[r][c] = ok ? v[((size_t)b * Tk + k0 + r) * ldv + h * HD + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp * 4 + i;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
      for (int c = 0; c < HD; ++c) {
        const float qv = sq[r][c];
        s0 = fmaf(qv, sk[lane][c], s0);
        s1 = fmaf(qv, sk[lane + 32][c], s1);
      }
      const int j0 = k0 + lane, j1 = k0 + lane + 32;
      if (j0 >= Tk || (kvb && !kvb[j0])) s0 = -INFINITY;
      if (j1 >= Tk || (kvb && !kvb[j1])) s1 = -INFINITY;
      const float tile_max = warp_max(fmaxf(s0, s1));
      const float m_new = fmaxf(m[i], tile_max);
      // a fully masked prefix keeps m = -inf; guard the (-inf) - (-inf) case
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m[i] - m_new);
      const float p0 = (s0 == -INFINITY) ? 0.f : __expf(s0 - m_new);
      const float p1 = (s1 == -INFINITY) ? 0.f : __expf(s1 - m_new);
      l[i] = l[i] * corr + warp_sum(p0 + p1);
      float a0 = o0[i] * corr, a1 = o1[i] * corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj0 = __shfl_sync(0xffffffffu, p0, j);
        const float pj1 = __shfl_sync(0xffffffffu, p1, j);
        a0 = fmaf(pj0, sv[j][lane], a0);
        a1 = fmaf(pj0, sv[j][lane + 32], a1);
        a0 = fmaf(pj1, sv[j + 32][lane], a0);
        a1 = fmaf(pj1, sv[j + 32][lane + 32], a1);
      }
      o0[i] = a0;
      o1[i] = a1;
      m[i] = m_new;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + warp * 4 + i;
    if (r < Tq) {
      // all keys masked: torch's softmax over an all -inf row yields NaN; keep that behaviour
      const float inv = 1.0f / l[i];
      const size_t off = ((size_t)b * Tq + r) * ldo + h * HD;
      if (out_f32) {
        float* orow = reinterpret_cast<float*>(out_v) + off;
        orow[lane] = o0[i] * inv;
        orow[lane + 32] = o1[i] * inv;
      } else {
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(out_v) + off;
        orow[lane] = __float2bfloat16_rn(o0[i] * inv);
        orow[lane + 32] = __float2bfloat16_rn(o1[i] * inv);
      }
    }
  }
}

uint32_t pow2_at_least(uint32_t v) {
  uint32_t p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

int vmc_get_option(int option);

extern "C" {

int vmc_attention_vit_impl(const void* qkv, void* out, int F, int L, int heads, int impl,
                           void* stream) {
  VMC_CHECK_ARG(qkv && out, VMC_ERR_ARG, "vmc_attention_vit: null pointer");
  VMC_CHECK_ARG(F > 0 && heads > 0 && L > 0 && L <= 272, VMC_ERR_SHAPE,
                "vmc_attention_vit: need 0 < L <= 272 tokens (L=%d)", L);
  VMC_CHECK_ARG(impl == 1 || impl == 2, VMC_ERR_ARG, "vmc_attention_vit: impl must be 1 or 2");
  const int d = heads * HD;
  const bool p_tmem = impl == 2;
  AttnArgs a;
  a.F = F;
  a.L = L;
  a.heads = heads;
  a.d = d;
  a.n_mt = (L + 127) / 128;
  a.lk16 = ((L + 15) / 16) * 16;
  const int lk32 = ((a.lk16 + 31) / 32) * 32;
  if (p_tmem) {
    a.n_pkb = 0;
    a.o_col = ((lk32 / 2 + 31) / 32) * 32;  // first 32-aligned column past the packed P
    const int need = lk32 > a.o_col + HD ? lk32 : a.o_col + HD;
    a.tmem_cols = pow2_at_least(need);
  } else {
    a.n_pkb = (lk32 + 63) / 64;
    a.o_col = 0;
    a.tmem_cols = pow2_at_least(lk32 > 64 ? lk32 : 64);
  }
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  CUtensorMap tm;
  const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)F};
  const uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
  const uint32_t box[3] = {HD, 128, 1};
  VMC_TRY(vmc_encode_tmap_bf16(&tm, qkv, 3, dims, strides, box));
  const uint32_t smem = (3 * a.n_mt + a.n_pkb) * TILE + 64 + 1024;
  VMC_CHECK_ARG(smem <= 227 * 1024, VMC_ERR_SHAPE, "vmc_attention_vit: L=%d needs %u B of smem", L,
                smem);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double fl = 4.0 * F * heads * (double)L * L * HD;
  const unsigned grid = (unsigned)((long long)F * heads);
  if (p_tmem) {
    VMC_CUDA(cudaFuncSetAttribute(attention_vit_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VmcProfScope prof(VMC_K_ATTN_VIT, st, fl, 8.0 * F * L * d);
    attention_vit_kernel<true><<<grid, 128, smem, st>>>(tm, a);
  } else {
    VMC_CUDA(cudaFuncSetAttribute(attention_vit_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VmcProfScope prof(VMC_K_ATTN_VIT, st, fl, 8.0 * F * L * d);
    attention_vit_kernel<false><<<grid, 128, smem, st>>>(tm, a);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_attention_vit(const void* qkv, void* out, int F, int L, int heads, void* stream) {
  return vmc_attention_vit_impl(qkv, out, F, L, heads, vmc_get_option(VMC_OPT_ATTN_IMPL) == 1 ? 1 : 2,
                                stream);
}

int vmc_attention_masked(const float* q, long long ldq, const float* k, long long ldk,
                         const float* v, long long ldv, const uint8_t* key_valid, void* out,
                         int out_f32, long long ldo, int B, int Tq, int Tk, int heads,
                         void* stream) {
  VMC_CHECK_ARG(q && k && v && out, VMC_ERR_ARG, "vmc_attention_masked: null pointer");
  VMC_CHECK_ARG(B > 0 && Tq > 0 && Tk > 0 && heads > 0 && B <= 65535 && heads <= 65535,
                VMC_ERR_SHAPE, "vmc_attention_masked: bad shape B=%d Tq=%d Tk=%d heads=%d", B, Tq,
                Tk, heads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((Tq + QB - 1) / QB, heads, B);
  {
    VmcProfScope prof(VMC_K_ATTN_SMALL, st, 4.0 * B * heads * (double)Tq * Tk * HD,
                      4.0 * B * heads * HD * (Tq + 2.0 * Tk) + 2.0 * B * Tq * heads * HD);
    attention_masked_kernel<<<grid, 128, 0, st>>>(q, ldq, k, ldk, v, ldv, key_valid,
                                                  out, out_f32, ldo, Tq, Tk);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // extern "C"
