// G-family, second generation: CTA-pair (cta_group::2) tcgen05 GEMM.
//
// Why a pair: with one CTA per 128 x 256 tile, every k-step moves A (4 KB) + B (8 KB) into shared
// memory by TMA and out again into the tensor core, 192 B/cycle/SM at the full MMA rate against a
// 128 B/cycle shared-memory pipe -- measured: the single-CTA kernel (gemm.cu) saturates at ~65 % of
// the MMA rate with the tensor pipe waiting on operands.  Two CTAs on the two SMs of a TPC share one
// 256 x 256 tile: each holds its own 128 rows of A and HALF of B (128 rows), the tensor cores read
// the other half from the peer's shared memory, so per SM the traffic is 128 B/cycle.
//
//   cluster (2,1,1); CTA rank r owns rows [m0 + 128 r, +128) of the pair tile
//   warp 0      TMA producer in BOTH CTAs (own A rows, own half of B); all transaction bytes are
//               signalled on the LEADER's (rank 0) full barrier
//   warp 1      TMEM allocator (both CTAs, cta_group::2); MMA issuer in the leader only:
//               tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16), commits multicast to both CTAs
//   warps 2..9  epilogue in both CTAs on their own 128 TMEM lanes; accumulator-free signal goes to
//               the leader's tempty barrier (remote mbarrier arrive)
// Epilogue: tcgen05.ld gives a thread one accumulator ROW.  The tower's modes (bf16 output with bias [+ QuickGELU]
// [+ folded LayerNorm]; bf16 residual stream + row statistics) do all arithmetic in that row layout and move whole
// 32 x 32 bf16 boxes between a private XOR-swizzled staging tile and global memory by TMA (epilogue_rowmajor,
// epilogue_rowmajor_resid); the fp32-residual and generic modes transpose the fp32 chunk through the staging tile so
// that global loads (residual, bias) and stores are issued with 8 lanes per 128-byte row segment (epilogue_fast).
// Pipeline: 4 stages x 32 KB at BN = 256 (the MMA issuer is never short of operands: tools/gemm_timeline.py), 12 KB of
// epilogue staging per warp.
#include <type_traits>

#include "common.cuh"
#include "vimoclip_b200.h"

long long vmc_get_option64(int option);
int vmc_get_option(int option);

namespace {

using namespace vmc;

constexpr int BM = 128;  // rows per CTA (pair tile: 256)
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;

struct Gemm2Args {
  int M, N, K;
  int tiles_m, tiles_n;  // pair tiles
  int a_mn, b_mn;        // operand given TRANSPOSED in memory ([K, M] / [K, N] row-major): MN-major UMMA operand
  vmc_gemm_epilogue epi;
  int store_tma;         // bf16 epilogues without a residual: the output leaves through tmC (TMA store)
  long long* dbg;        // VMC_OPT_DEBUG_PTR: clock64 stamps of CTA 0's first 32 tiles, 8 slots each (tools/gemm_timeline.py)
};
#define VMC_DBG2(i_, slot)                                                                              \
  do {                                                                                                  \
    if (g.dbg != nullptr && blockIdx.x == 0 && (i_) < 32) g.dbg[(i_) * 8 + (slot)] = clock64();         \
  } while (0)

// MN-major SW128 operand: TMA boxes of 64 (MN, contiguous) x 64 (K) elements land as 64 rows of 128 bytes = the canonical
// layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) with SBO = 1024 B between 8-row K groups and LBO = 8192 B between the
// 64-element MN atoms (one box each).  A 16-deep K step advances the start address by two K groups (2048 B).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(8192 >> 4) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

template <int BN>
struct Cfg2 {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of B
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // Pipeline depth at BN = 256: 6, 5 and 4 stages measured the same on every tower shape (profiles/r02_gemm_stages.txt; the MMA
  // issuer never waits for operands), so 4 it is, and the 64 KB go to the epilogue: 12 KB of staging per warp, which is
  // what lets the bf16-residual epilogue keep a whole tile part of residual rows in flight (epilogue_rowmajor_resid).
#ifndef VMC_G2_STAGES
#define VMC_G2_STAGES 4
#endif
  static constexpr int STAGES = (BN == 256) ? VMC_G2_STAGES : 8;
  static constexpr uint32_t STG_PER_WARP = (BN == 256) ? 12288 : 4096;
  static constexpr uint32_t STG_BYTES = 8 * STG_PER_WARP;
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the barrier only orders TMEM reuse (tcgen05.wait::ld + tcgen05.fence before it); a
  // release here makes ptxas emit MEMBAR.ALL.GPU, which waited for every epilogue store to drain
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* m,
                                                uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive(1) on the barrier at this offset in BOTH CTAs of the pair when all prior MMAs retire
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

// TMA store of one shared-memory box (bulk async-group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__device__ __forceinline__ float act2(float v, int act) {
  switch (act) {
    case VMC_ACT_QUICKGELU:
      return __fdividef(v, 1.0f + __expf(-1.702f * v));
    case VMC_ACT_GELU_ERF:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case VMC_ACT_RELU:
      return fmaxf(v, 0.0f);
    default:
      return v;
  }
}

// x * sigmoid(1.702 x) with sigmoid(z) = 0.5 + 0.5 tanh(z / 2): one MUFU per element instead of two
// the same from h = x / 2: h + h tanh(1.702 h)
__device__ __forceinline__ float quickgelu_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(1.702f * h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float quickgelu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

// Specialised epilogues for the shapes that carry the ViT (row_group == 0, bias present, alpha == 1,
// N a multiple of 32): all row addressing is hoisted to one pointer + stride per tile.
//   MODE 1: bf16 out = acc + bias            (qkv)
//   MODE 2: bf16 out = QuickGELU(acc + bias) (mlp.c_fc)
//   MODE 3: fp32 out = resid + acc + bias    (attn.out_proj, mlp.c_proj; out may alias resid)
// LNF   (consumer of a folded LayerNorm, MODE 1 / 2): acc <- rstd[m] * acc - mean[m] * rstd[m] * colsum[n]
//       before the bias; mean / rstd come from the producer's partial row sums (statistics over K columns).
// STATS (producer, MODE 3): also write the bf16 copy of the output rows and this warp's partial
//       (sum, sum of squares) of each row over its HALF_N columns.
// R16   (MODE 3 + STATS): the residual stream itself is bf16 -- resid is read as bf16, out is written as bf16 (it IS the A
//       operand of the next GEMM: no separate raw16 copy) and the statistics are those of the ROUNDED values.
// row statistics of a folded LayerNorm for row m: (rstd, -mean * rstd) from the producer's partial (sum, sum of squares)
__device__ __forceinline__ void ln_row_stats(const vmc_gemm_epilogue& e, int m, int M, int K, float& rstd, float& shift) {
  rstd = 1.f;
  shift = 0.f;
  if (m < M) {
    float s = 0.f, q = 0.f;
    for (int p = 0; p < e.stats_parts; ++p) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(e.stats_in) + (long long)p * e.stats_ld + m);
      s += v.x;
      q += v.y;
    }
    const float mean = s / (float)K;
    const float var = fmaxf(q / (float)K - mean * mean, 0.f);
    rstd = rsqrtf(var + e.ln_eps);
    shift = -mean * rstd;
  }
}

// The residual rows of a WHOLE tile part (32 rows x HALF_N columns of bf16 per warp: 8 x HALF_N / 32 uint2 per lane), requested
// BEFORE the warp waits for the accumulator.  Round-2 timeline of attn.out_proj (K = 768): the epilogue took 9.4 - 10.7 k
// cycles per tile against 8.5 k for the MMA main loop -- 2.5 k cycles per 32-column chunk, the latency of the residual loads
// under load, which a one-chunk look-ahead does not cover.  The loads do not depend on the accumulator, so their latency now
// overlaps the wait for it.
template <int HALF_N>
__device__ __forceinline__ void prefetch_resid16(const vmc_gemm_epilogue& e, int M, int N, int row0, int n_base, int lane,
                                                 uint2 (&pre)[HALF_N / 32][8]) {
  const int lr = lane >> 3, lc = lane & 7;
  const int m_first = row0 + lr;
  int nvalid = (M - m_first + 3) >> 2;
  nvalid = nvalid < 0 ? 0 : (nvalid > 8 ? 8 : nvalid);
  const char* rptr = reinterpret_cast<const char*>(e.resid) + ((long long)m_first * e.ldr + n_base + lc * 4) * 2;
  const long long rstride = 4 * e.ldr * 2;
#pragma unroll
  for (int c = 0; c < HALF_N / 32; ++c) {
    const char* rp = rptr + c * 64;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pre[c][i] = make_uint2(0u, 0u);
      if (i < nvalid && n_base + c * 32 < N) pre[c][i] = *reinterpret_cast<const uint2*>(rp);
      rp += rstride;
    }
  }
}

template <int MODE, int HALF_N, bool LNF = false, bool STATS = false, bool R16 = false>
__device__ __forceinline__ void epilogue_fast(const vmc_gemm_epilogue& e, int M, int N, int K, int row0,
                                              int n_base, uint32_t t_acc, uint8_t* stg, int lane, float rstd = 1.f,
                                              float shift = 0.f, const uint2 (*pre)[8] = nullptr) {
  const int lr = lane >> 3, lc = lane & 7;
  const int m_first = row0 + lr;
  int nvalid = (M - m_first + 3) >> 2;  // rows m_first + 4 i, i < nvalid, are inside the matrix
  nvalid = nvalid < 0 ? 0 : (nvalid > 8 ? 8 : nvalid);
  constexpr int ESZ = (MODE == 3 && !R16) ? 4 : 2;
  constexpr int RSZ = R16 ? 2 : 4;  // bytes per residual element
  char* optr = reinterpret_cast<char*>(e.out) + ((long long)m_first * e.ldo + n_base + lc * 4) * ESZ;
  const long long ostride = 4 * e.ldo * ESZ;
  const char* rptr = nullptr;
  long long rstride = 0;
  if constexpr (MODE == 3) {
    rptr = reinterpret_cast<const char*>(e.resid) + ((long long)m_first * e.ldr + n_base + lc * 4) * RSZ;
    rstride = 4 * e.ldr * RSZ;
  }
  const float* bptr = e.bias + n_base + lc * 4;
  // folded LayerNorm: the caller derived rstd / -mean * rstd of ONE row per thread (row0 + lane) from the producer's partial
  // sums BEFORE waiting for the accumulator (ln_row_stats: the loads are off the tile's critical path); after the
  // transpose a lane holds rows 4 i + lr, so the 8 (rstd, -mean * rstd) pairs it needs come by shuffle
  float ln_rs[8], ln_sh[8];
  if constexpr (LNF) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ln_rs[i] = __shfl_sync(0xffffffffu, rstd, i * 4 + lr);
      ln_sh[i] = __shfl_sync(0xffffffffu, shift, i * 4 + lr);
    }
  }
  float psum[8], psq[8];
  char* r16ptr = nullptr;
  long long r16stride = 0;
  if constexpr (STATS) {
#pragma unroll
    for (int i = 0; i < 8; ++i) psum[i] = psq[i] = 0.f;
    if constexpr (!R16) {
      r16ptr = reinterpret_cast<char*>(e.raw16_out) + ((long long)m_first * e.raw16_ld + n_base + lc * 4) * 2;
      r16stride = 4 * e.raw16_ld * 2;
    }
  }
  // residual rows of one 32-column chunk: 4 consecutive elements per lane and row (float4, or 4 bf16 kept packed in a uint2)
  using ResT = typename std::conditional<R16, uint2, float4>::type;
  auto res_f4 = [](const ResT& u) -> float4 {
    if constexpr (R16) {
      return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                         __uint_as_float(u.y & 0xFFFF0000u));
    } else {
      return u;
    }
  };
  ResT res_next[8];
  if constexpr (MODE == 3 && !R16) {
    const char* rp = rptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nvalid && n_base < N) res_next[i] = *reinterpret_cast<const ResT*>(rp);
      rp += rstride;
    }
  }
  // bias / column sums of the NEXT chunk are requested while this one is processed (ncu: the FMAs that consume them were
  // 13 % of the c_fc kernel's stall samples, waiting on these two loads)
  float4 b4n = make_float4(0.f, 0.f, 0.f, 0.f), cs4n = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n_base < N) {
    b4n = __ldg(reinterpret_cast<const float4*>(bptr));
    if constexpr (LNF) cs4n = __ldg(reinterpret_cast<const float4*>(e.colsum + n_base + lc * 4));
  }
#pragma unroll(R16 ? HALF_N / 32 : 1)  // R16: unrolled, the prefetched residual registers are indexed by the chunk
  for (int c = 0; c < HALF_N / 32; ++c) {
    if (n_base + c * 32 >= N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
    const float4 b4 = b4n;
    const float4 cs4 = cs4n;
    if (c + 1 < HALF_N / 32 && n_base + (c + 1) * 32 < N) {
      b4n = __ldg(reinterpret_cast<const float4*>(bptr + (c + 1) * 32));
      if constexpr (LNF) cs4n = __ldg(reinterpret_cast<const float4*>(e.colsum + n_base + (c + 1) * 32 + lc * 4));
    }
    ResT res[8];
    if constexpr (MODE == 3 && R16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = pre[c][i];
    } else if constexpr (MODE == 3) {
      // the residual rows of the NEXT chunk are requested while this one is processed: the epilogue of the
      // K = 768 residual GEMMs is bound by bytes in flight (8 warps x 4 KB per SM), not by HBM bandwidth
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = res_next[i];
      if (c + 1 < HALF_N / 32 && n_base + (c + 1) * 32 < N) {
        const char* rp = rptr + (c + 1) * 32 * RSZ;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nvalid) res_next[i] = *reinterpret_cast<const ResT*>(rp);
          rp += rstride;
        }
      }
    }
    tmem_ld_wait();
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
    char* op = optr + c * 32 * ESZ;
    char* r16p = (STATS && !R16) ? r16ptr + c * 64 : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = w[i];
      if constexpr (LNF) {  // rstd * acc - mean * rstd * colsum[n] + bias'[n]
        v.x = fmaf(v.x, ln_rs[i], fmaf(ln_sh[i], cs4.x, b4.x));
        v.y = fmaf(v.y, ln_rs[i], fmaf(ln_sh[i], cs4.y, b4.y));
        v.z = fmaf(v.z, ln_rs[i], fmaf(ln_sh[i], cs4.z, b4.z));
        v.w = fmaf(v.w, ln_rs[i], fmaf(ln_sh[i], cs4.w, b4.w));
      } else {
        v.x += b4.x;
        v.y += b4.y;
        v.z += b4.z;
        v.w += b4.w;
      }
      if constexpr (MODE == 2) {
        v.x = quickgelu_fast(v.x);
        v.y = quickgelu_fast(v.y);
        v.z = quickgelu_fast(v.z);
        v.w = quickgelu_fast(v.w);
      }
      if constexpr (MODE == 3) {
        const float4 rr = res_f4(res[i]);
        v.x += rr.x;
        v.y += rr.y;
        v.z += rr.z;
        v.w += rr.w;
        if constexpr (R16) {
          uint2 o;
          o.x = pack_bf16x2(v.x, v.y);
          o.y = pack_bf16x2(v.z, v.w);
          if (i < nvalid) *reinterpret_cast<uint2*>(op) = o;
          if constexpr (STATS) {  // of the values the next GEMM will read
            const float r0 = __uint_as_float(o.x << 16), r1 = __uint_as_float(o.x & 0xFFFF0000u);
            const float r2 = __uint_as_float(o.y << 16), r3 = __uint_as_float(o.y & 0xFFFF0000u);
            psum[i] += (r0 + r1) + (r2 + r3);
            psq[i] += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
          }
        } else {
          if (i < nvalid) *reinterpret_cast<float4*>(op) = v;
          if constexpr (STATS) {
            psum[i] += (v.x + v.y) + (v.z + v.w);
            psq[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            uint2 o;
            o.x = pack_bf16x2(v.x, v.y);
            o.y = pack_bf16x2(v.z, v.w);
            if (i < nvalid) *reinterpret_cast<uint2*>(r16p) = o;
            r16p += r16stride;
          }
        }
      } else {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        if (i < nvalid) *reinterpret_cast<uint2*>(op) = o;
      }
      op += ostride;
    }
  }
  if constexpr (STATS) {
    // the 8 lanes lc = 0..7 of a row hold its HALF_N columns between them
    float2* sp = reinterpret_cast<float2*>(e.stats_out) + (long long)(n_base / HALF_N) * e.stats_ld + m_first;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = psum[i], q = psq[i];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lc == 0 && i < nvalid) sp[4 * i] = make_float2(s, q);
    }
  }
}

// bf16-output epilogues without a residual (MODE 1 / 2, LayerNorm-folded 6 / 7: qkv, c_fc -- 58 % of the tower's GEMM FLOPs).
// ncu / cycle accounting of round 2: at K = 768 a pair tile is 6 144 tensor cycles during which the UMMA operand reads already take
// the whole 128 B/clk of shared-memory bandwidth; epilogue_fast stages the fp32 accumulator through shared memory to transpose it
// (128 KB written + 128 KB read per tile and CTA = 2 048 more cycles of that pipe) -- which is where the K = 768 GEMMs lose their
// ~25 % against K = 3072.  Here ALL the arithmetic runs in the TMEM-native layout (a thread owns one accumulator row: its rstd /
// shift are per-thread scalars, bias / colsum are warp-uniform and come as broadcast reads of a 1 KB per-warp shared copy), and
// only the PACKED bf16 result is transposed: half the staging traffic, 4 + 4 shared-memory and 4 global instructions per
// 32-column chunk instead of 8 + 8 + 8.
//   stg (4 KB per warp): [0, 2048) swizzled 32 x 64-byte tile; [2048, 2560) bias of the warp's HALF_N columns; [2560, 3072) colsum
// tmc != nullptr: the staged 32 x 32 bf16 tile leaves through ONE TMA store (the XOR pattern of the staging tile is the
// tensor map's 64-byte swizzle) instead of 4 LDS.128 + 4 STG.128 per lane.  The tensor core's operand reads, the TMA writes
// and every LSU access share the SM's shared-memory / L1 data pipe: the round-2 tile timeline shows the main loop stretching by
// about one cycle per epilogue wavefront on that pipe (profiles/r02_gemm_tile_timeline.txt), and the bulk store reads the tile
// once (16 wavefronts per chunk) where the LDS + STG pair moved it twice.  Out-of-range rows / columns are clipped by the map.
template <int MODE, int HALF_N, bool LNF>
__device__ __forceinline__ void epilogue_rowmajor(const vmc_gemm_epilogue& e, int M, int N, int row0, int n_base, uint32_t t_acc,
                                                  uint8_t* stg, int lane, float rstd, float shift, const CUtensorMap* tmc) {
  float* sb = reinterpret_cast<float*>(stg + 2048);
  float* scs = sb + 128;
  // QuickGELU(v) = v sigmoid(1.702 v) = h + h tanh(1.702 h) with h = v / 2: the halving is folded into the per-row scalars and
  // the shared bias copy (exact: powers of two), which leaves FMUL + MUFU.TANH + FFMA per element
  constexpr float HS = MODE == 2 ? 0.5f : 1.0f;
  if (4 * lane < HALF_N && n_base + 4 * lane < N) {
    float4 bv = __ldg(reinterpret_cast<const float4*>(e.bias + n_base) + lane);
    bv.x *= HS, bv.y *= HS, bv.z *= HS, bv.w *= HS;
    *reinterpret_cast<float4*>(sb + 4 * lane) = bv;
    if constexpr (LNF) *reinterpret_cast<float4*>(scs + 4 * lane) = __ldg(reinterpret_cast<const float4*>(e.colsum + n_base) + lane);
  }
  rstd *= HS;
  shift *= HS;
  __syncwarp();
  const int sw = (lane >> 1) & 3;           // write side: this thread's row = lane
  const int rrow = lane >> 2, rch = lane & 3;  // read side: rows 8 i + rrow, 16-byte chunk rch
  __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(e.out) + (long long)(row0 + rrow) * e.ldo + n_base + rch * 8;
#pragma unroll 1
  for (int c = 0; c < HALF_N / 32; ++c) {
    if (n_base + c * 32 >= N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 32 + 4 * q);  // same address in every lane: broadcast
      float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                             __uint_as_float(r[4 * q + 3]));
      if constexpr (LNF) {  // rstd * acc - mean * rstd * colsum[n] + bias'[n]
        const float4 cs4 = *reinterpret_cast<const float4*>(scs + c * 32 + 4 * q);
        v.x = fmaf(v.x, rstd, fmaf(shift, cs4.x, b4.x));
        v.y = fmaf(v.y, rstd, fmaf(shift, cs4.y, b4.y));
        v.z = fmaf(v.z, rstd, fmaf(shift, cs4.z, b4.z));
        v.w = fmaf(v.w, rstd, fmaf(shift, cs4.w, b4.w));
      } else if constexpr (MODE == 2) {
        v.x = fmaf(v.x, HS, b4.x);
        v.y = fmaf(v.y, HS, b4.y);
        v.z = fmaf(v.z, HS, b4.z);
        v.w = fmaf(v.w, HS, b4.w);
      } else {
        v.x += b4.x;
        v.y += b4.y;
        v.z += b4.z;
        v.w += b4.w;
      }
      if constexpr (MODE == 2) {
        v.x = quickgelu_half(v.x);
        v.y = quickgelu_half(v.y);
        v.z = quickgelu_half(v.z);
        v.w = quickgelu_half(v.w);
      }
      pk[2 * q] = pack_bf16x2(v.x, v.y);
      pk[2 * q + 1] = pack_bf16x2(v.z, v.w);
    }
    if (tmc != nullptr) {
      if (lane == 0) bulk_wait_read0();  // the previous chunk's store has finished reading the tile
      __syncwarp();
    }
    // row `lane` of the 32 x 32 bf16 tile: four 16-byte chunks, chunk index XOR ((row >> 1) & 3): conflict-free both ways
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((ch ^ sw) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    if (tmc != nullptr) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tma_store_2d(tmc, smem_u32(stg), n_base + c * 32, row0);
      continue;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = 8 * i + rrow;
      const uint4 o = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((rch ^ ((rl >> 1) & 3)) << 4));
      if (row0 + rl < M) *reinterpret_cast<uint4*>(obase + (long long)(8 * i) * e.ldo + c * 32) = o;  // 4 lanes = one 64-byte row segment
    }
    __syncwarp();
  }
}

// bf16 residual stream (MODE 8) in the TMEM-native row layout, BN = 256.  A thread owns one accumulator row, so its row
// statistics are two private registers (no shuffles) and nothing fp32 is staged.  The residual rows of the warp's whole tile
// part arrive by FOUR TMA loads (32 x 32 bf16 boxes, 64-byte swizzle) issued by lane 0 before the warp waits for the
// accumulator -- their latency overlaps that wait -- and complete on the warp's own mbarrier; each thread reads its row of a box
// with four conflict-free LDS.128.  The result rows go out through a TMA store from one of two alternating 2 KB tiles.
//   stg (12 KB per warp): [0, 8192) residual boxes of chunks 0..3; [8192, 12288) two output tiles
// In-place use (out == resid) is safe: the warp's residual boxes have all landed before its first store is issued.
template <int HALF_N>
__device__ __forceinline__ void resid_rowmajor_prefetch(const CUtensorMap* tmr, uint32_t stg_addr, uint32_t rbar, int N, int row0,
                                                        int n_base, int lane) {
  if (lane == 0) {
    int nb = 0;
#pragma unroll
    for (int c = 0; c < HALF_N / 32; ++c) nb += (n_base + c * 32 < N) ? 1 : 0;
    mbar_arrive_expect_tx(rbar, (uint32_t)nb * 2048u);
#pragma unroll
    for (int c = 0; c < HALF_N / 32; ++c)
      if (n_base + c * 32 < N) tma_load_2d(stg_addr + c * 2048u, tmr, rbar, n_base + c * 32, row0);
  }
}

template <int HALF_N>
__device__ __forceinline__ void epilogue_rowmajor_resid(const vmc_gemm_epilogue& e, int M, int N, int row0, int n_base, uint32_t t_acc,
                                                        uint8_t* stg, uint32_t rbar, uint32_t rphase, int lane, const CUtensorMap* tmc) {
  const int sw = (lane >> 1) & 3;
  float psum = 0.f, psq = 0.f;
  mbar_wait(rbar, rphase);
#pragma unroll 1
  for (int c = 0; c < HALF_N / 32; ++c) {
    if (n_base + c * 32 >= N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
    uint4 rs[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) rs[ch] = *reinterpret_cast<const uint4*>(stg + c * 2048 + lane * 64 + ((ch ^ sw) << 4));
    tmem_ld_wait();
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(rs);
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n_base + c * 32) + q);  // same address in every lane
      const uint32_t w0 = rw[2 * q], w1 = rw[2 * q + 1];
      const float v0 = __uint_as_float(r[4 * q]) + b4.x + __uint_as_float(w0 << 16);
      const float v1 = __uint_as_float(r[4 * q + 1]) + b4.y + __uint_as_float(w0 & 0xFFFF0000u);
      const float v2 = __uint_as_float(r[4 * q + 2]) + b4.z + __uint_as_float(w1 << 16);
      const float v3 = __uint_as_float(r[4 * q + 3]) + b4.w + __uint_as_float(w1 & 0xFFFF0000u);
      const uint32_t o0 = pack_bf16x2(v0, v1), o1 = pack_bf16x2(v2, v3);
      pk[2 * q] = o0;
      pk[2 * q + 1] = o1;
      // statistics of the values the next GEMM will read
      const float r0 = __uint_as_float(o0 << 16), r1 = __uint_as_float(o0 & 0xFFFF0000u);
      const float r2 = __uint_as_float(o1 << 16), r3 = __uint_as_float(o1 & 0xFFFF0000u);
      psum += (r0 + r1) + (r2 + r3);
      psq += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
    }
    uint8_t* ot = stg + 8192 + (c & 1) * 2048;
    if (lane == 0) bulk_wait_read1();  // the store issued two chunks ago has finished reading this tile
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      *reinterpret_cast<uint4*>(ot + lane * 64 + ((ch ^ sw) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) tma_store_2d(tmc, smem_u32(ot), n_base + c * 32, row0);
  }
  if (row0 + lane < M && n_base < N)
    reinterpret_cast<float2*>(e.stats_out)[(long long)(n_base / HALF_N) * e.stats_ld + row0 + lane] = make_float2(psum, psq);
}

template <int BN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                          const __grid_constant__ CUtensorMap tmR, const Gemm2Args g) {
  using C = Cfg2<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t stg_base = base + STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = stg_base + C::STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  auto resid_bar = [&](int ew) { return bar_base + 8u * (2 * STAGES + 5 + ew); };  // one per epilogue warp (MODE 8, BN = 256)
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_kb = (g.K + BK - 1) / BK;
  const int num_tiles = g.tiles_m * g.tiles_n;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // Tile order: tile t = pair + i * num_pairs with N fastest, so the pairs running together share one A block.
  const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
  auto tile_of = [&](int i, int& m_blk, int& n_blk) {
    const int t = pair + i * num_pairs;
    n_blk = t % g.tiles_n;
    m_blk = t / g.tiles_n;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.store_tma) tma_prefetch_desc(&tmC);
    if (g.store_tma && MODE == 8) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // leader's producer arrive (+ both CTAs' transaction bytes)
      mbar_init(empty_bar(s), 1);  // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);    // multicast tcgen05.commit
      mbar_init(tempty_bar(a), 16);  // 8 epilogue warps x 2 CTAs (used in the leader only)
    }
    for (int w = 0; w < 8; ++w) mbar_init(resid_bar(w), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr_addr, C::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int m_blk, n_blk;
        tile_of(i, m_blk, n_blk);
        const int m_row = m_blk * (2 * BM) + (int)rank * BM;
        const int n_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint32_t full_leader = mapa_rank(full_bar(stage), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * C::STAGE_BYTES);
          if (g.a_mn) {
            tma_load_2d_cg2(sa, &tmA, full_leader, m_row, kb * BK);
            tma_load_2d_cg2(sa + 8192, &tmA, full_leader, m_row + 64, kb * BK);
          } else {
            tma_load_2d_cg2(sa, &tmA, full_leader, kb * BK, m_row);
          }
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_cg2(sb + j * 8192, &tmB, full_leader, n_row + 64 * j, kb * BK);
          } else {
            tma_load_2d_cg2(sb, &tmB, full_leader, kb * BK, n_row);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, g.a_mn, g.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        VMC_DBG2(i, 0);
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        long long starved = 0;  // debug timeline: cycles this tile's issuer waited for operands that had not landed yet
        for (int kb = 0; kb < num_kb; ++kb) {
          if (g.dbg != nullptr && blockIdx.x == 0 && !mbar_test_wait(full_bar(stage), phase)) {
            const long long w0 = clock64();
            mbar_wait(full_bar(stage), phase);
            starved += clock64() - w0;
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = g.a_mn ? umma_desc_sw128_mn(sa) : umma_desc_sw128(sa);
          const uint64_t db = g.b_mn ? umma_desc_sw128_mn(sb) : umma_desc_sw128(sb);
          const uint64_t sta = g.a_mn ? 128 : 2, stb = g.b_mn ? 128 : 2;  // start-address step per UMMA_K (16-byte units)
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_ss_cg2(d_tmem, da + sta * k, db + stb * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_mc2(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_mc2(tfull_bar(acc));
        VMC_DBG2(i, 1);
        if (g.dbg != nullptr && blockIdx.x == 0 && i < 32) g.dbg[i * 8 + 6] = starved;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (8, both CTAs) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_N = BN / 2;
    const vmc_gemm_epilogue& e = g.epi;
    uint8_t* stg = smem_raw + (stg_base - raw_addr) + ew * C::STG_PER_WARP;
    const int lr = lane >> 3;
    const int lc = lane & 7;
    const bool has_res = e.resid != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int m_blk, n_blk;
      tile_of(i, m_blk, n_blk);
      const int row0 = m_blk * (2 * BM) + (int)rank * BM + quarter * 32;
      const int n_base = n_blk * BN + half * HALF_N;
      float ln_rstd = 1.f, ln_shift = 0.f;
      if constexpr (MODE == 6 || MODE == 7) ln_row_stats(e, row0 + lane, g.M, g.K, ln_rstd, ln_shift);  // before the wait
      constexpr bool RESID_TMA = MODE == 8 && BN == 256;  // row-layout epilogue, residual by TMA (needs the 12 KB staging)
      uint2 rpre[(MODE == 8 && !RESID_TMA) ? HALF_N / 32 : 1][8];
      if constexpr (RESID_TMA) {
        if (g.store_tma) resid_rowmajor_prefetch<HALF_N>(&tmR, smem_u32(stg), resid_bar(ew), g.N, row0, n_base, lane);
      }
      if constexpr (MODE == 8 && !RESID_TMA) prefetch_resid16<HALF_N>(e, g.M, g.N, row0, n_base, lane, rpre);
      if (ew == 0 && lane == 0) VMC_DBG2(i, 2);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (ew == 0 && lane == 0) VMC_DBG2(i, 3);
      const uint32_t t_acc =
          tmem_base + uint32_t(acc * BN + half * HALF_N) + (uint32_t(quarter * 32) << 16);
      if constexpr (RESID_TMA) {
        epilogue_rowmajor_resid<HALF_N>(e, g.M, g.N, row0, n_base, t_acc, stg, resid_bar(ew), (uint32_t)i & 1u, lane, &tmC);
      } else if constexpr (MODE == 8) {
        epilogue_fast<3, HALF_N, false, true, true>(e, g.M, g.N, g.K, row0, n_base, t_acc, stg, lane, 1.f, 0.f, rpre);
      } else if constexpr (MODE == 5) {
        epilogue_fast<3, HALF_N, false, true>(e, g.M, g.N, g.K, row0, n_base, t_acc, stg, lane);
      } else if constexpr (MODE == 6 || MODE == 7) {
        epilogue_rowmajor<MODE - 5, HALF_N, true>(e, g.M, g.N, row0, n_base, t_acc, stg, lane, ln_rstd, ln_shift, g.store_tma ? &tmC : nullptr);
      } else if constexpr (MODE == 1 || MODE == 2) {
        epilogue_rowmajor<MODE, HALF_N, false>(e, g.M, g.N, row0, n_base, t_acc, stg, lane, 1.f, 0.f, g.store_tma ? &tmC : nullptr);
      } else if constexpr (MODE != 0) {
        epilogue_fast<MODE, HALF_N>(e, g.M, g.N, g.K, row0, n_base, t_acc, stg, lane);
      } else {
        // ---- generic path: any activation / alpha / ragged N / patch-embed row remap ----
        // [r2] The rows of a lane are the same for every 32-column chunk of the tile: their (remapped) output / residual row
        // offsets are computed once per tile (the runtime division by row_group was repeated 8 x per chunk), and the residual of
        // the NEXT chunk is requested while this one is processed.  Timeline of the patch embedding before: epilogue 17 - 23 k
        // cycles per tile against 8.7 k for the main loop.
        long long obase[8], rbase[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = row0 + i * 4 + lr;
          long long orow = m, rrow = m;
          if (e.row_group > 0) {
            const int f = m / e.row_group;
            orow = (long long)m + f + 1;
            rrow = m - f * e.row_group + 1;
          }
          obase[i] = (m < g.M) ? orow * e.ldo : -1;
          rbase[i] = rrow * e.ldr;
        }
        auto load_res = [&](int col, bool col_full, float4 (&dst)[8]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && obase[i] >= 0 && col_full) {
              if (e.resid_bf16) {
                const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.resid) + rbase[i] + col);
                dst[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
              } else {
                dst[i] = *reinterpret_cast<const float4*>(e.resid + rbase[i] + col);
              }
            }
          }
        };
        float4 res_next[8];
        if (n_base < g.N) load_res(n_base + lc * 4, n_base + lc * 4 + 4 <= g.N, res_next);
#pragma unroll 1
        for (int c = 0; c < HALF_N / 32; ++c) {
          const int n0 = n_base + c * 32;
          if (n0 >= g.N) break;
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_acc + uint32_t(c * 32), r);
          const int col = n0 + lc * 4;
          const bool col_full = col + 4 <= g.N;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e.bias != nullptr && col_full) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
          float4 res[8];
          long long ooff[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            res[i] = res_next[i];
            ooff[i] = obase[i] >= 0 ? obase[i] + col : -1;
          }
          if (c + 1 < HALF_N / 32 && n0 + 32 < g.N) load_res(col + 32, col + 36 <= g.N, res_next);
          tmem_ld_wait();
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
                v.x = act2(v.x, e.act);
                v.y = act2(v.y, e.act);
                v.z = act2(v.z, e.act);
                v.w = act2(v.w, e.act);
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
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (ooff[i] < 0) continue;
              const float wv[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
              for (int q = 0; q < 4; ++q) {
                if (col + q < g.N) {
                  float x = wv[q];
                  if (e.bias != nullptr) x += __ldg(e.bias + col + q);
                  x = act2(x, e.act) * e.alpha;
                  if (has_res)
                    x += e.resid_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.resid)[rbase[i] + col + q])
                                      : e.resid[rbase[i] + col + q];
                  if (e.out_bf16)
                    reinterpret_cast<__nv_bfloat16*>(e.out)[ooff[i] + q] = __float2bfloat16_rn(x);
                  else
                    reinterpret_cast<float*>(e.out)[ooff[i] + q] = x;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));
      if (lane == 0 && (ew == 0 || ew == 7)) VMC_DBG2(i, ew == 0 ? 4 : 5);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  if (g.store_tma && warp >= 2 && lane == 0) bulk_wait_read0();  // the staging tiles stay valid until the last store has read them
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int MODE>
int launch_gemm2(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                 const vmc_gemm_epilogue* epi, cudaStream_t stream, int a_mn = 0, int b_mn = 0) {
  using C = Cfg2<BN>;
  CUtensorMap tmA, tmB;
  if (a_mn) {  // A given as A^T [K, M] row-major
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, BK};
    VMC_TRY(vmc_encode_tmap_bf16(&tmA, A, 2, dims, strides, box));
  } else {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {BK, BM};
    VMC_TRY(vmc_encode_tmap_bf16(&tmA, A, 2, dims, strides, box));
  }
  if (b_mn) {  // W given as W^T [K, N] row-major
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {64, BK};
    VMC_TRY(vmc_encode_tmap_bf16(&tmB, W, 2, dims, strides, box));
  } else {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {BK, BN / 2};
    VMC_TRY(vmc_encode_tmap_bf16(&tmB, W, 2, dims, strides, box));
  }
  Gemm2Args g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.a_mn = a_mn;
  g.b_mn = b_mn;
  CUtensorMap tmC = tmA, tmR = tmA;  // placeholders when unused
  g.store_tma = 0;
  if ((MODE == 1 || MODE == 2 || MODE == 6 || MODE == 7) && (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0 && (epi->ldo % 8) == 0 &&
      vmc_get_option(VMC_OPT_GEMM_IMPL) != 3) {  // option value 3: LDS + STG epilogue (A/B runs)
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)epi->ldo * 2};
    const uint32_t box[2] = {32, 32};
    VMC_TRY(vmc_encode_tmap_bf16_sw(&tmC, epi->out, 2, dims, strides, box, 64));
    g.store_tma = 1;
  }
  if (MODE == 8 && BN == 256) {  // row-layout bf16-residual epilogue: residual in, result out by TMA
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint32_t box[2] = {32, 32};
    const uint64_t so[1] = {(uint64_t)epi->ldo * 2}, sr[1] = {(uint64_t)epi->ldr * 2};
    VMC_TRY(vmc_encode_tmap_bf16_sw(&tmC, epi->out, 2, dims, so, box, 64));
    VMC_TRY(vmc_encode_tmap_bf16_sw(&tmR, epi->resid, 2, dims, sr, box, 64));
    g.store_tma = 1;
  }
  g.tiles_m = (M + 2 * BM - 1) / (2 * BM);
  g.tiles_n = (N + BN - 1) / BN;
  g.epi = *epi;
  g.dbg = reinterpret_cast<long long*>((uintptr_t)(unsigned long long)vmc_get_option64(VMC_OPT_DEBUG_PTR));
  VMC_CUDA(cudaFuncSetAttribute(gemm2_bf16_tcgen05_kernel<BN, MODE>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = g.tiles_m * g.tiles_n;
  const int max_pairs = vmc_num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  {
    const double out_b = (double)M * N * (epi->out_bf16 ? 2 : 4);
    VmcProfScope prof(VMC_K_GEMM, stream, 2.0 * M * N * K,
                      2.0 * ((double)M * K + (double)N * K) + out_b + (epi->resid ? (epi->resid_bf16 ? 2.0 : 4.0) * M * N : 0.0) +
                          (epi->raw16_out ? 2.0 * M * N : 0.0));
    gemm2_bf16_tcgen05_kernel<BN, MODE><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, tmC, tmR, g);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // namespace

// Column slices per row a producer GEMM (stats_out) of this shape writes: one per epilogue warp column half.
extern "C" int vmc_gemm_stats_parts(int M, int N) {
  if (M <= 0 || N <= 0) return -1;
  const long long tiles256 = (long long)((M + 255) / 256) * ((N + 255) / 256);
  const bool big = N > 128 && tiles256 >= (long long)vmc_num_sms();
  return (N + (big ? 127 : 63)) / (big ? 128 : 64);
}

// Called by vmc_gemm_bf16 (gemm.cu) after argument validation.
int vmc_gemm2_dispatch(const void* A, long long lda, const void* W, long long ldw, int M, int N,
                       int K, const vmc_gemm_epilogue* epi, cudaStream_t stream, int a_mn, int b_mn) {
  const long long tiles256 = (long long)((M + 255) / 256) * ((N + 255) / 256);
  const bool big = N > 128 && tiles256 >= (long long)vmc_num_sms();
  // specialised epilogues (see epilogue_fast): the four GEMMs of every ViT block
  const bool fast_ok = epi->row_group == 0 && epi->bias != nullptr && epi->alpha == 1.0f && (N % 32) == 0;
  int mode = 0;
  if (fast_ok) {
    if (epi->out_bf16 && !epi->resid && epi->act == VMC_ACT_NONE) mode = 1;
    else if (epi->out_bf16 && !epi->resid && epi->act == VMC_ACT_QUICKGELU) mode = 2;
    else if (!epi->out_bf16 && epi->resid && !epi->resid_bf16 && epi->act == VMC_ACT_NONE) mode = 3;
  }
  if (epi->resid && epi->resid_bf16 && epi->out_bf16 && epi->stats_out != nullptr) {
    // bf16 residual stream: out = bf16(acc + bias + resid), row statistics of the rounded values
    VMC_CHECK_ARG(fast_ok && epi->act == VMC_ACT_NONE && epi->raw16_out == nullptr &&
                      epi->stats_in == nullptr && (N % (big ? 128 : 64)) == 0 && epi->stats_ld >= M &&
                      (epi->ldr % 4) == 0 && (epi->ldo % 4) == 0 && (reinterpret_cast<uintptr_t>(epi->resid) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(epi->out) & 7) == 0 && (reinterpret_cast<uintptr_t>(epi->stats_out) & 7) == 0,
                  VMC_ERR_ARG,
                  "vmc_gemm_bf16: the bf16 residual-stream epilogue needs bias, alpha 1, no activation, N a multiple of the "
                  "column slice (%d), stats_ld >= M and 8-byte aligned rows", big ? 128 : 64);
    if (big) {
      VMC_CHECK_ARG((epi->ldr % 8) == 0 && (epi->ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(epi->resid) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0,
                    VMC_ERR_ALIGN, "vmc_gemm_bf16: the bf16 residual-stream epilogue moves its rows by TMA: 16-byte aligned resid / out rows");
      return launch_gemm2<256, 8>(A, lda, W, ldw, M, N, K, epi, stream, a_mn, b_mn);
    }
    return launch_gemm2<128, 8>(A, lda, W, ldw, M, N, K, epi, stream, a_mn, b_mn);
  }
  if (epi->raw16_out != nullptr || epi->stats_out != nullptr) {
    VMC_CHECK_ARG(mode == 3 && epi->raw16_out && epi->stats_out && (N % (big ? 128 : 64)) == 0 &&
                      (epi->raw16_ld % 4) == 0 && epi->raw16_ld >= N && epi->stats_ld >= M &&
                      (reinterpret_cast<uintptr_t>(epi->raw16_out) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(epi->stats_out) & 7) == 0,
                  VMC_ERR_ARG,
                  "vmc_gemm_bf16: raw16_out / stats_out need the fp32 bias+residual epilogue, both pointers, N a multiple "
                  "of the column slice (%d) and stats_ld >= M", big ? 128 : 64);
    mode = 5;
  }
  if (epi->stats_in != nullptr) {
    VMC_CHECK_ARG((mode == 1 || mode == 2) && epi->colsum && epi->stats_parts > 0 && epi->stats_ld >= M &&
                      (reinterpret_cast<uintptr_t>(epi->colsum) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(epi->stats_in) & 7) == 0,
                  VMC_ERR_ARG,
                  "vmc_gemm_bf16: a folded LayerNorm (stats_in) needs the bf16 bias [+ QuickGELU] epilogue, colsum and "
                  "stats_parts > 0");
    mode += 5;
  }
#define VMC_G2(BN_, MODE_) return launch_gemm2<BN_, MODE_>(A, lda, W, ldw, M, N, K, epi, stream, a_mn, b_mn)
  if (big) {
    switch (mode) {
      case 1: VMC_G2(256, 1);
      case 2: VMC_G2(256, 2);
      case 3: VMC_G2(256, 3);
      case 5: VMC_G2(256, 5);
      case 6: VMC_G2(256, 6);
      case 7: VMC_G2(256, 7);
      default: VMC_G2(256, 0);
    }
  }
  switch (mode) {
    case 1: VMC_G2(128, 1);
    case 2: VMC_G2(128, 2);
    case 3: VMC_G2(128, 3);
    case 5: VMC_G2(128, 5);
    case 6: VMC_G2(128, 6);
    case 7: VMC_G2(128, 7);
    default: VMC_G2(128, 0);
  }
#undef VMC_G2
}
