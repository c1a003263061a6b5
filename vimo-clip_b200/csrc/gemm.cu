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
//   warps 2..9  epilogue: tcgen05.ld 32 lanes x 32 columns -> bias/act/residual -> global
//               (two warps per TMEM lane quarter, each owning half of the tile's columns)
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
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr uint32_t BIAS_BYTES = BN * 4 * 4;  // 8 epilogue warps x BN/2 floats
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
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
  const uint32_t bar_base = base + STAGES * C::STAGE_BYTES;
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
    // Per tile each warp (a) stages its half of the bias vector in shared memory BEFORE the
    // accumulator is ready, (b) software-prefetches the residual of chunk c+1 while chunk c is
    // processed, so no global-load latency sits between tcgen05.ld and the stores.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_N = BN / 2;
    constexpr int CH = HALF_N / 32;  // 32-column chunks per warp
    const vmc_gemm_epilogue& e = g.epi;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + C::BAR_BYTES - raw_addr)) + ew * HALF_N;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % g.tiles_n;
      const int m_blk = t / g.tiles_n;
      const int m = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = m < g.M;
      long long orow = m, rrow = m;
      if (e.row_group > 0) {
        const int f = m / e.row_group;
        orow = (long long)m + f + 1;
        rrow = m - f * e.row_group + 1;
      }
      const int n_base = n_blk * BN + half * HALF_N;
      // (a) bias -> smem (zero beyond N), one float4 (or one float for BN = 128... HALF_N/32 floats) per lane
      __syncwarp();
#pragma unroll
      for (int j = lane * 4; j < HALF_N; j += 128) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias != nullptr) {
          if (n_base + j + 4 <= g.N) {
            b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n_base + j));
          } else {
            if (n_base + j + 0 < g.N) b4.x = __ldg(e.bias + n_base + j + 0);
            if (n_base + j + 1 < g.N) b4.y = __ldg(e.bias + n_base + j + 1);
            if (n_base + j + 2 < g.N) b4.z = __ldg(e.bias + n_base + j + 2);
          }
        }
        *reinterpret_cast<float4*>(bias_s + j) = b4;
      }
      __syncwarp();
      const float* rbase = (e.resid != nullptr && row_ok) ? e.resid + rrow * e.ldr + n_base : nullptr;
      float4 res[2][8];
      if (rbase != nullptr && n_base + 32 <= g.N) {
#pragma unroll
        for (int j = 0; j < 8; ++j) res[0][j] = reinterpret_cast<const float4*>(rbase)[j];
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_acc =
          tmem_base + uint32_t(acc * BN + half * HALF_N) + (uint32_t(quarter * 32) << 16);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 < g.N) {  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
          // (b) prefetch the next chunk's residual while this one is in flight
          if (c + 1 < CH && rbase != nullptr && n0 + 64 <= g.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              res[(c + 1) & 1][j] = reinterpret_cast<const float4*>(rbase + (c + 1) * 32)[j];
          }
          tmem_ld_wait();
          if (row_ok) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + 4 * j);
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
            }
            if (e.act != VMC_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act);
            }
            if (e.alpha != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
            }
            if (n0 + 32 <= g.N) {
              if (e.resid != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = res[c & 1][j];
                  v[4 * j + 0] += b.x;
                  v[4 * j + 1] += b.y;
                  v[4 * j + 2] += b.z;
                  v[4 * j + 3] += b.w;
                }
              }
              if (e.out_bf16) {
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) +
                                                     orow * e.ldo + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 o;
                  o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                  o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                  o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                  o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                  op[j] = o;
                }
              } else {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) +
                                                       orow * e.ldo + n0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  op[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            } else {
              // ragged last column chunk: scalar path (residual read directly)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int n = n0 + j;
                if (n < g.N) {
                  float x = v[j];
                  if (e.resid != nullptr) x += e.resid[rrow * e.ldr + n];
                  if (e.out_bf16)
                    reinterpret_cast<__nv_bfloat16*>(e.out)[orow * e.ldo + n] = __float2bfloat16_rn(x);
                  else
                    reinterpret_cast<float*>(e.out)[orow * e.ldo + n] = x;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
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

extern "C" int vmc_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M,
                             int N, int K, const vmc_gemm_epilogue* epi, void* stream) {
  VMC_CHECK_ARG(A && W && epi && epi->out, VMC_ERR_ARG, "vmc_gemm_bf16: null pointer");
  VMC_CHECK_ARG(M > 0 && N > 0 && K > 0, VMC_ERR_SHAPE, "vmc_gemm_bf16: bad shape M=%d N=%d K=%d",
                M, N, K);
  VMC_CHECK_ARG(lda >= K && ldw >= K && (lda % 8) == 0 && (ldw % 8) == 0, VMC_ERR_ALIGN,
                "vmc_gemm_bf16: lda/ldw must be >= K and multiples of 8 elements (lda=%lld ldw=%lld K=%d)",
                lda, ldw, K);
  const int oalign = epi->out_bf16 ? 8 : 4;
  VMC_CHECK_ARG(epi->ldo >= N && (epi->ldo % oalign) == 0 &&
                    (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0,
                VMC_ERR_ALIGN, "vmc_gemm_bf16: out must be 16-byte aligned with ldo %% %d == 0",
                oalign);
  if (epi->resid)
    VMC_CHECK_ARG((epi->ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(epi->resid) & 15) == 0,
                  VMC_ERR_ALIGN, "vmc_gemm_bf16: resid must be 16-byte aligned with ldr %% 4 == 0");
  if (epi->bias)
    VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(epi->bias) & 15) == 0, VMC_ERR_ALIGN,
                  "vmc_gemm_bf16: bias must be 16-byte aligned");
  VMC_CHECK_ARG(epi->act >= VMC_ACT_NONE && epi->act <= VMC_ACT_RELU, VMC_ERR_ARG,
                "vmc_gemm_bf16: unknown activation %d", epi->act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // 128x256 tiles when there are enough of them to fill the machine twice; 128x128 otherwise.
  const long long tiles256 = (long long)((M + BM - 1) / BM) * ((N + 255) / 256);
  if (N > 128 && tiles256 >= 2LL * vmc_num_sms())
    return launch_gemm<256>(A, lda, W, ldw, M, N, K, epi, st);
  return launch_gemm<128>(A, lda, W, ldw, M, N, K, epi, st);
}
