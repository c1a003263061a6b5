// G-family: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[orow, n] = alpha * act(sum_k A[m,k] * W[n,k] + bias[n]) + resid[rrow, n]
//
// A [M,K] and W [N,K] are bf16, K-major (W is exactly nn.Linear.weight), so both
// operands are TMA-loaded as 128B-swizzled [rows x 64] boxes and consumed by
// tcgen05.mma.kind::f16 (UMMA 128 x BN x 16) with the fp32 accumulator in TMEM.
//
// Warp roles (320 threads, one CTA per SM, grid = min(#SM, #tiles)):
//   warp 0      TMA producer  (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue: tcgen05.ld 32 lanes x 32 columns -> smem transpose -> bias/act/residual
//               -> coalesced global stores (two warps per TMEM lane quarter, half the columns each)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer full/empty
// (MMA <-> epilogue) so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Reference call sites this replaces: the cuBLAS/cuDNN GEMMs under nn.Conv2d patch
// embed, nn.MultiheadAttention in/out projections, mlp.c_fc / c_proj and `@ proj`
// reached from models/student_model.py:84 and extract_embeddings.py:94, and every
// nn.Linear of TFAM/models/AMO_CLIP.py and of the student heads.
#include "common.cuh"
#include "vimoclip_b200.h"

namespace {

using namespace vmc;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps

struct GemmArgs {
  int M, N, K;
  int tiles_m, tiles_n;
  vmc_gemm_epilogue epi;
};

template <int BN>
struct Cfg {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t STG_BYTES = 8 * 4096;  // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // double-buffered accumulator (power of two)
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case VMC_ACT_QUICKGELU:
      // x * sigmoid(1.702 x)  (OpenAI clip/model.py QuickGELU; HF hidden_act="quick_gelu")
      return __fdividef(v, 1.0f + __expf(-1.702f * v));
    case VMC_ACT_GELU_ERF:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case VMC_ACT_RELU:
      return fmaxf(v, 0.0f);
    default:
      return v;
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * C::STAGE_BYTES + C::STG_BYTES;
  // barrier layout (8 B each): full[STAGES] | empty[STAGES] | tfull[2] | tempty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (g.K + BK - 1) / BK;
  const int num_tiles = g.tiles_m * g.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_blk = t % g.tiles_n;
        const int m_blk = t / g.tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM);
          tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4) units
            umma_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc,
                    (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (8) =====================
    // Warps 2..9: TMEM lane quarter = warp & 3 (hardware rule), column half = (warp - 2) >> 2.
    // tcgen05.ld hands each thread one ROW of the accumulator (32 consecutive columns).  Writing
    // that straight to global memory makes every warp store touch 32 different 128-byte lines
    // (measured: the epilogue, not the MMA, bounded the K = 768 GEMMs).  So each warp transposes
    // its 32 x 32 fp32 chunk through a private 4 KB, XOR-swizzled shared-memory tile and then
    // reads/writes global memory with 8 lanes per 128-byte row segment (4 full lines per warp
    // instruction): residual loads, bias and the stores are all coalesced.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_N = BN / 2;
    constexpr int CH = HALF_N / 32;  // 32-column chunks per warp
    const vmc_gemm_epilogue& e = g.epi;
    uint8_t* stg = smem_raw + (base + STAGES * C::STAGE_BYTES - raw_addr) + ew * 4096;
    const int lr = lane >> 3;  // row within a group of 4 rows
    const int lc = lane & 7;   // 16-byte column chunk within the 128-byte row segment
    const bool has_res = e.resid != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % g.tiles_n;
      const int m_blk = t / g.tiles_n;
      const int row0 = m_blk * BM + quarter * 32;
      const int n_base = n_blk * BN + half * HALF_N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_acc =
          tmem_base + uint32_t(acc * BN + half * HALF_N) + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < CH; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 >= g.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
        // coalesced-layout operands of this chunk, issued before the TMEM load is waited for
        const int col = n0 + lc * 4;
        const bool col_full = col + 4 <= g.N;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias != nullptr && col_full) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
        float4 res[8];
        long long ooff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = row0 + i * 4 + lr;
          long long orow = m, rrow = m;
          if (e.row_group > 0) {
            const int f = m / e.row_group;
            orow = (long long)m + f + 1;
            rrow = m - f * e.row_group + 1;
          }
          ooff[i] = (m < g.M) ? orow * e.ldo + col : -1;
          res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_res && m < g.M && col_full)
            res[i] = *reinterpret_cast<const float4*>(e.resid + rrow * e.ldr + col);
        }
        tmem_ld_wait();
        // transpose through the swizzled staging tile: row = lane, 16-byte chunk j at (j ^ (lane & 7))
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        float4 w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + lr;
          w[i] = *reinterpret_cast<const float4*>(stg + rl * 128 + ((lc ^ (rl & 7)) << 4));
        }
        __syncwarp();
        if (col_full) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (ooff[i] < 0) continue;
            float4 v = w[i];
            v.x += b4.x;
            v.y += b4.y;
            v.z += b4.z;
            v.w += b4.w;
            if (e.act != VMC_ACT_NONE) {
              v.x = apply_act(v.x, e.act);
              v.y = apply_act(v.y, e.act);
              v.z = apply_act(v.z, e.act);
              v.w = apply_act(v.w, e.act);
            }
            v.x = fmaf(v.x, e.alpha, res[i].x);
            v.y = fmaf(v.y, e.alpha, res[i].y);
            v.z = fmaf(v.z, e.alpha, res[i].z);
            v.w = fmaf(v.w, e.alpha, res[i].w);
            if (e.out_bf16) {
              uint2 o;
              o.x = pack_bf16x2(v.x, v.y);
              o.y = pack_bf16x2(v.z, v.w);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + ooff[i]) = o;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + ooff[i]) = v;
            }
          }
        } else if (col < g.N) {
          // ragged last columns (N not a multiple of 4 or of the chunk): scalar path
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (ooff[i] < 0) continue;
            const float wv[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
            const int m = row0 + i * 4 + lr;
            long long rrow = m;
            if (e.row_group > 0) rrow = m - (m / e.row_group) * e.row_group + 1;
            for (int q = 0; q < 4; ++q) {
              if (col + q < g.N) {
                float x = wv[q];
                if (e.bias != nullptr) x += __ldg(e.bias + col + q);
                x = apply_act(x, e.act) * e.alpha;
                if (has_res) x += e.resid[rrow * e.ldr + col + q];
                if (e.out_bf16)
                  reinterpret_cast<__nv_bfloat16*>(e.out)[ooff[i] + q] = __float2bfloat16_rn(x);
                else
                  reinterpret_cast<float*>(e.out)[ooff[i] + q] = x;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN>
int launch_gemm(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                const vmc_gemm_epilogue* epi, cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {BK, BM};
    VMC_TRY(vmc_encode_tmap_bf16(&tmA, A, 2, dims, strides, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {BK, BN};
    VMC_TRY(vmc_encode_tmap_bf16(&tmB, W, 2, dims, strides, box));
  }
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.tiles_m = (M + BM - 1) / BM;
  g.tiles_n = (N + BN - 1) / BN;
  g.epi = *epi;
  // per device/context attribute; cheap enough to set on every launch (DataParallel: several devices)
  VMC_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = g.tiles_m * g.tiles_n;
  const int grid = tiles < vmc_num_sms() ? tiles : vmc_num_sms();
  {
    const double out_b = (double)M * N * (epi->out_bf16 ? 2 : 4);
    VmcProfScope prof(VMC_K_GEMM, stream, 2.0 * M * N * K,
                      2.0 * ((double)M * K + (double)N * K) + out_b + (epi->resid ? 4.0 * M * N : 0.0));
    gemm_bf16_tcgen05_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, g);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // namespace

int vmc_gemm2_dispatch(const void* A, long long lda, const void* W, long long ldw, int M, int N,
                       int K, const vmc_gemm_epilogue* epi, cudaStream_t stream, int a_mn, int b_mn);
int vmc_get_option(int option);

extern "C" int vmc_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M,
                             int N, int K, const vmc_gemm_epilogue* epi, void* stream) {
  return vmc_gemm_bf16_ex(A, lda, 0, W, ldw, 0, M, N, K, epi, stream);
}

extern "C" int vmc_gemm_bf16_ex(const void* A, long long lda, int a_transposed, const void* W, long long ldw,
                                int w_transposed, int M, int N, int K, const vmc_gemm_epilogue* epi, void* stream) {
  VMC_CHECK_ARG(A && W && epi && epi->out, VMC_ERR_ARG, "vmc_gemm_bf16: null pointer");
  VMC_CHECK_ARG(M > 0 && N > 0 && K > 0, VMC_ERR_SHAPE, "vmc_gemm_bf16: bad shape M=%d N=%d K=%d",
                M, N, K);
  VMC_CHECK_ARG(lda >= (a_transposed ? M : K) && ldw >= (w_transposed ? N : K) && (lda % 8) == 0 && (ldw % 8) == 0,
                VMC_ERR_ALIGN,
                "vmc_gemm_bf16: lda/ldw must cover the operand's row length and be multiples of 8 elements (lda=%lld ldw=%lld K=%d)",
                lda, ldw, K);
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, VMC_ERR_ALIGN,
                "vmc_gemm_bf16: operands must be 16-byte aligned");
  const int oalign = epi->out_bf16 ? 8 : 4;
  VMC_CHECK_ARG(epi->ldo >= N && (epi->ldo % oalign) == 0 &&
                    (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0,
                VMC_ERR_ALIGN, "vmc_gemm_bf16: out must be 16-byte aligned with ldo %% %d == 0",
                oalign);
  if (epi->resid)
    VMC_CHECK_ARG((epi->ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(epi->resid) & (epi->resid_bf16 ? 7 : 15)) == 0,
                  VMC_ERR_ALIGN, "vmc_gemm_bf16: resid must be 16-byte (fp32) / 8-byte (bf16) aligned with ldr %% 4 == 0");
  if (epi->bias)
    VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(epi->bias) & 15) == 0, VMC_ERR_ALIGN,
                  "vmc_gemm_bf16: bias must be 16-byte aligned");
  VMC_CHECK_ARG(epi->act >= VMC_ACT_NONE && epi->act <= VMC_ACT_RELU, VMC_ERR_ARG,
                "vmc_gemm_bf16: unknown activation %d", epi->act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (vmc_get_option(VMC_OPT_GEMM_IMPL) != 1)  // default: CTA-pair kernel (gemm2.cu)
    return vmc_gemm2_dispatch(A, lda, W, ldw, M, N, K, epi, st, a_transposed ? 1 : 0, w_transposed ? 1 : 0);
  VMC_CHECK_ARG(!a_transposed && !w_transposed, VMC_ERR_ARG,
                "vmc_gemm_bf16_ex: transposed (MN-major) operands exist only in the CTA-pair kernel");
  VMC_CHECK_ARG(epi->raw16_out == nullptr && epi->stats_out == nullptr &&
                    epi->stats_in == nullptr && !epi->resid_bf16,
                VMC_ERR_ARG, "vmc_gemm_bf16: the fused / folded LayerNorm epilogues exist only in the CTA-pair kernel");
  // single-CTA kernel: 128x256 tiles when there are enough to fill the machine twice; 128x128 otherwise.
  const long long tiles256 = (long long)((M + BM - 1) / BM) * ((N + 255) / 256);
  if (N > 128 && tiles256 >= 2LL * vmc_num_sms())
    return launch_gemm<256>(A, lda, W, ldw, M, N, K, epi, st);
  return launch_gemm<128>(A, lda, W, ldw, M, N, K, epi, st);
}
