// TFAM fusion block as ONE kernel (north star piece 4): self-attention -> cross-attention -> FFN (post-LN) x layers ->
// temporal mean over ALL rows -> LayerNorm -> Linear -> GELU(erf) -> Linear, i.e. the whole of
// TFAM/models/AMO_CLIP.py:37-51 (AttentionLayer.forward), :146-150 (layer loop), :170 (pool + classifier).
//
// Mapping.  All 22 configurations of the reference (TFAM/cfg_AK/*.yaml) use d_model 512, nhead 8, dim_feedforward 2048:
// the kernel is specialised for that geometry (others take the batched tcgen05 path of tfam.py).  A clip has 16 rows, a
// UMMA tile 128, and the block's 16.8 M weights (33.6 MB as fp16) must stream through whoever computes a clip -- so a
// clip (or a group of two) is given to a CLUSTER of 8 CTAs, one per attention head:
//   * every projection is split over the cluster along its OUTPUT columns (CTA c computes head c's q / k / v, 64 columns
//     of out_proj, 256 hidden units of the FFN), so each CTA streams 1/8 of the weights, straight from L2 into
//     mma.sync B fragments: the host packs the weights once in "fragment-major" order -- per (CTA, warp) ONE contiguous
//     stream of 512-byte blocks (32 lanes x 16 B = the B fragments of one 8-column tile for two 16-deep k steps) in
//     exactly the order the warp consumes them, across GEMMs and layers -- and each warp keeps 16 blocks (8 KB) in
//     flight in a register ring that runs ahead across GEMM boundaries and cluster barriers;
//   * activations never leave the SMs: the fp16 A operand [rows, 512] is replicated in every CTA's shared memory, the
//     fp32 residual stream is column-sharded (64 columns per CTA); slices are exchanged through distributed shared
//     memory: attention output all-gather, FFN2 partial-sum reduce-scatter (fixed order: deterministic), LayerNorm row
//     statistics, normalised rows all-gather.  Every exchange is st.async (SASS STAS) with mbarrier complete_tx on the
//     RECEIVER's barrier: a CTA waits on its own mbarrier for the bytes of all eight senders, no cluster-wide barrier
//     and no fence.  (The first version used barrier.cluster.arrive.release / wait.acquire between phases: the release
//     compiles to MEMBAR.ALL.GPU, which also waits for the warp's 16 in-flight weight loads -- ncu: 27 % of the stall
//     samples on the 38 barriers of a pass, profiles/r02_tfam_fused_v1_ncu_summary.txt.)  Buffer reuse is safe without
//     extra handshakes because consecutive uses of a buffer are always separated by another all-to-all exchange: a
//     CTA can only be one exchange ahead of its slowest peer;
//   * the attention of head c (<= 32 keys) runs in fp32 on the CUDA cores of CTA c out of shared memory.
// Operands are fp16 (not bf16): with one rounding per operand the config-1 logits are within 2.4e-3 of fp32 (bf16:
// 1.2e-2, over the 1e-2 bar; tools/emulate_tfam_precision.py), fp32 accumulation, fp32 residual / LayerNorm / softmax.
//
// Latency regime (B = 2, the reference's evaluation batches): 2 clusters = 16 SMs, 4.2 MB of weights per CTA.
// Throughput regime (B = 256): 16 co-resident clusters, two clips (32 rows, two M tiles) per pass so the weight
// stream is read from L2 once per two clips.
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "vimoclip_b200.h"

namespace {

using namespace vmc;

constexpr int TF_D = 512, TF_HD = 64, TF_FF = 2048;
constexpr int TF_NC = 8;  // CTAs per cluster = heads
constexpr int TF_NW = 8;  // warps per CTA
constexpr int TF_THREADS = TF_NW * 32;
constexpr int TF_PD = 16;                    // weight blocks in flight per warp (register ring)
constexpr int TF_LDA = TF_D + 8;             // halfs per row of the replicated fp16 operands (16-byte row skew: conflict-free ldmatrix)
constexpr int TF_FFC = TF_FF / TF_NC;        // hidden units per CTA (256)
constexpr int TF_LDH = TF_FFC + 8;           // halfs per row of the local hidden operand
constexpr int TF_LDK = TF_HD + 1;            // floats per row of K (lane = key: odd stride)
constexpr int TF_MAXG = 2;                   // clips per group (pooled / head buffers)
// 512-byte weight blocks per warp and GEMM (n-tiles of the warp x k-pairs), in stream order
constexpr int NB_SIN = 48, NB_SOUT = 16, NB_CQ = 16, NB_CKV = 32, NB_COUT = 16, NB_F1 = 64, NB_F2 = 64;
constexpr int NB_LAYER = NB_SIN + NB_SOUT + NB_CQ + NB_CKV + NB_COUT + NB_F1 + NB_F2;  // 256
static_assert(NB_SIN % TF_PD == 0 && NB_SOUT % TF_PD == 0 && NB_CQ % TF_PD == 0 && NB_CKV % TF_PD == 0 &&
                  NB_COUT % TF_PD == 0 && NB_F1 % TF_PD == 0 && NB_F2 % TF_PD == 0,
              "every GEMM must consume a whole number of ring revolutions (static ring indexing)");

struct TfLayer {
  const float *b_sin, *b_sout, *b_cin, *b_cout, *b_f1, *b_f2;
  const float *ns_g, *ns_b, *nc_g, *nc_b, *nf_g, *nf_b;
  float ns_eps, nc_eps, nf_eps;
};
struct TfArgs {
  const float* x;          // [B, T, 512]
  const float* mot;        // [B, Tm, 512] (cross-attention source) or nullptr
  const uint8_t* valid_x;  // [B, T] 1 = real frame, or nullptr
  const uint8_t* valid_m;  // [B, Tm] or nullptr
  float* logits;           // [B, C]
  const uint4* wstream;    // [8 CTAs][8 warps][layers * 256 blocks][32 lanes] uint4
  int B, T, Tm, C, hidden, layers, G, act;
  const float *hd_g, *hd_b, *w1t, *b1, *w2t, *b2;  // classifier: LayerNorm, Linear^T [512, hidden], Linear^T [hidden, C]
  float hd_eps;
  TfLayer layer[VMC_TFAM_MAX_LAYERS];
};

// ---- cluster / DSMEM helpers ----
__device__ __forceinline__ uint32_t tf_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tf_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tf_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// st.async: the store and its byte count travel together; the receiver's mbarrier phase completes when the expected
// bytes of all senders have landed (and its own arrive.expect_tx has been posted)
__device__ __forceinline__ void tf_sta_u32(uint32_t addr, uint32_t v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tf_sta_f32(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tf_sta_f32x2(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a),
               "f"(b), "r"(bar)
               : "memory");
}
enum { XB_OH = 0, XB_STATS = 1, XB_XH = 2, XB_RECV = 3, XB_POOL = 4, XB_UH = 5, XB_COUNT = 6 };
// receiver side of an exchange: thread 0 posts the expected byte count (one arrival), everybody waits for the phase
__device__ __forceinline__ void tf_xchg_wait(uint32_t bar_base, uint32_t& parity_bits, int which, uint32_t bytes, int tid) {
  const uint32_t bar = bar_base + 8u * which;
  if (tid == 0) mbar_arrive_expect_tx(bar, bytes);
  mbar_wait(bar, (parity_bits >> which) & 1u);
  parity_bits ^= 1u << which;
}
__device__ __forceinline__ uint4 tf_ldg_w(const uint4* p) {  // weights: read once per pass, keep them out of L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t tf_pack_h2(float a, float b) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);  // saturate instead of overflowing to inf
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void tf_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void tf_mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[mt][i] (+)= A[16 mt .., K] * W_i^T for the warp's NT column tiles.  A: fp16 rows in shared memory (row stride lda
// halfs).  The warp's weight blocks arrive through `ring` (block b of this GEMM sits in slot b % PD; on entry the ring
// holds blocks 0 .. PD-1); every consumed slot is refilled with block b + PD, which past the end of this GEMM is block
// (b + PD - NB) of the NEXT GEMM the warp will run (`nxt`), so the stream never drains between GEMMs.
// Block (p, i) = k-pair p (k = 32 p .. 32 p + 31), column tile i; stream order p-major.  Lane (g = lane / 4, t = lane % 4)
// holds {W_i[g][32p + 2t, +1], W_i[g][32p + 2t + 8, +9], W_i[g][32p + 16 + 2t, +1], W_i[g][32p + 24 + 2t, +1]} = (b0, b1) of
// k-step 2p and of k-step 2p + 1 of mma.m16n8k16.
template <int MT, int NT, int KP>
__device__ __forceinline__ void tf_gemm(uint32_t a_smem, int lda, const uint4* cur, const uint4* nxt, uint4 (&ring)[TF_PD],
                                        float (&acc)[MT][NT][4], int lane) {
  constexpr int NB = NT * KP;
  static_assert(NB % TF_PD == 0, "ring revolutions");
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][i][q] = 0.f;
  // ldmatrix.x4 source rows: lanes 0-15 -> rows 0-15 at column k0, lanes 16-31 -> rows 0-15 at column k0 + 8
  const uint32_t a_lane = a_smem + (uint32_t)((lane & 15) * lda + (lane >> 4) * 8) * 2u;
  uint32_t afr[MT][2][4];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int p = b / NT, i = b % NT;
    const uint4 w = ring[b % TF_PD];
    ring[b % TF_PD] = tf_ldg_w(b + TF_PD < NB ? cur + (size_t)(b + TF_PD) * 32 : nxt + (size_t)(b + TF_PD - NB) * 32);
    if (i == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        tf_ldmatrix_x4(a_lane + (uint32_t)(mt * 16 * lda + p * 32) * 2u, afr[mt][0]);
        tf_ldmatrix_x4(a_lane + (uint32_t)(mt * 16 * lda + p * 32 + 16) * 2u, afr[mt][1]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      tf_mma_f16(acc[mt][i], afr[mt][0], w.x, w.y);
      tf_mma_f16(acc[mt][i], afr[mt][1], w.z, w.w);
    }
  }
}

template <int MT>
struct TfSmem {
  static constexpr int RM = 16 * MT;
  static constexpr uint32_t XH = 0;                                   // half [RM][LDA]   LayerNorm output / layer input (A operand)
  static constexpr uint32_t MH = XH + RM * TF_LDA * 2;                // half [RM][LDA]   motion rows (cross-attention source)
  static constexpr uint32_t OH = MH + RM * TF_LDA * 2;                // half [RM][LDA]   attention output, all heads
  static constexpr uint32_t HH = OH + RM * TF_LDA * 2;                // half [RM][LDH]   this CTA's 256 hidden units
  static constexpr uint32_t X32 = HH + RM * TF_LDH * 2;               // float [RM][64]   residual stream, this CTA's columns
  static constexpr uint32_t QS = X32 + RM * 64 * 4;                   // float [RM][64]
  static constexpr uint32_t KS = QS + RM * 64 * 4;                    // float [RM][65]
  static constexpr uint32_t VS = KS + RM * TF_LDK * 4;                // float [RM][64]
  static constexpr uint32_t RECV = (VS + RM * 64 * 4 + 15) & ~15u;    // float [8][RM][64] FFN2 partial sums from every CTA
  static constexpr uint32_t STATS = RECV + 8 * RM * 64 * 4;           // float [8][RM][2]  LayerNorm partial (sum, sum of squares)
  static constexpr uint32_t POOL = STATS + 8 * RM * 2 * 4;            // float [MAXG][512] temporal mean
  static constexpr uint32_t UH = POOL + TF_MAXG * TF_D * 4;           // float [MAXG][256] classifier hidden
  static constexpr uint32_t RED = UH + TF_MAXG * 256 * 4;             // float [8][32]     head reduction scratch
  static constexpr uint32_t PB = RED + 8 * 32 * 4;                    // float [8 warps][32] attention probabilities of a row
  static constexpr uint32_t BAR = PB + 8 * 32 * 4;                    // mbarrier [XB_COUNT]: one per exchange buffer
  static constexpr uint32_t TOTAL = BAR + 64;
};

// y (already in the x32 slice) -> LayerNorm over the 512 columns held by the 8 CTAs -> x32 slice (fp32, next residual) and
// the fp16 rows of every CTA's xh.  Two exchanges: row statistics, normalised-row all-gather.
template <int MT>
__device__ __forceinline__ void tf_layernorm_exchange(uint8_t* sm, uint32_t sm_addr, int R, uint32_t c, const float* gamma,
                                                      const float* beta, float eps, int warp, int lane, int tid,
                                                      uint32_t& parity_bits) {
  using S = TfSmem<MT>;
  float* x32 = reinterpret_cast<float*>(sm + S::X32);
  const float* stats = reinterpret_cast<const float*>(sm + S::STATS);
  __syncthreads();  // the slice is complete (written by all warps)
  for (int r = warp; r < R; r += TF_NW) {
    const float2 v = *reinterpret_cast<const float2*>(x32 + r * 64 + 2 * lane);
    const float s = warp_sum(v.x + v.y);
    const float q = warp_sum(v.x * v.x + v.y * v.y);
    if (lane < TF_NC)  // lane = destination CTA
      tf_sta_f32x2(tf_mapa(sm_addr + S::STATS + (uint32_t)((c * S::RM + r) * 2) * 4u, (uint32_t)lane), s, q,
                   tf_mapa(sm_addr + S::BAR + 8u * XB_STATS, (uint32_t)lane));
  }
  tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_STATS, (uint32_t)R * 64u, tid);
  const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + c * 64) + lane);
  const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + c * 64) + lane);
  uint32_t rbar[TF_NC];
#pragma unroll
  for (int j = 0; j < TF_NC; ++j) rbar[j] = tf_mapa(sm_addr + S::BAR + 8u * XB_XH, (uint32_t)j);
  for (int r = warp; r < R; r += TF_NW) {
    float s = 0.f, q = 0.f;
    if (lane < TF_NC) {
      const float2 p = *reinterpret_cast<const float2*>(stats + (lane * S::RM + r) * 2);
      s = p.x;
      q = p.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {  // fixed order over the 8 CTAs' partials (lanes 0..7)
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    q = __shfl_sync(0xffffffffu, q, 0);
    const float mean = s * (1.0f / TF_D);
    const float var = fmaxf(q * (1.0f / TF_D) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float2 v = *reinterpret_cast<const float2*>(x32 + r * 64 + 2 * lane);
    v.x = (v.x - mean) * rstd * gm.x + bt.x;
    v.y = (v.y - mean) * rstd * gm.y + bt.y;
    *reinterpret_cast<float2*>(x32 + r * 64 + 2 * lane) = v;
    const uint32_t h2 = tf_pack_h2(v.x, v.y);
    const uint32_t dst = sm_addr + S::XH + (uint32_t)(r * TF_LDA + c * 64 + 2 * lane) * 2u;
#pragma unroll
    for (int j = 0; j < TF_NC; ++j) tf_sta_u32(tf_mapa(dst, (uint32_t)j), h2, rbar[j]);
  }
  tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_XH, (uint32_t)R * 1024u, tid);
}

// attention of head c for the rows of this group (fp32, CUDA cores): one warp per query row, lane = key.
// keys of clip gi are rows [gi * Tk, gi * Tk + Tk) of ks / vs.  Output: fp16 columns [64 c, 64 c + 64) of EVERY CTA's oh.
template <int MT>
__device__ __forceinline__ void tf_attention(uint8_t* sm, uint32_t sm_addr, int R, int Tq, int Tk, const uint8_t* valid,
                                             int clip0, uint32_t c, int warp, int lane) {
  using S = TfSmem<MT>;
  const float* qs = reinterpret_cast<const float*>(sm + S::QS);
  const float* ks = reinterpret_cast<const float*>(sm + S::KS);
  const float* vs = reinterpret_cast<const float*>(sm + S::VS);
  float* pb = reinterpret_cast<float*>(sm + S::PB) + warp * 32;
  uint32_t rbar[TF_NC];
#pragma unroll
  for (int j = 0; j < TF_NC; ++j) rbar[j] = tf_mapa(sm_addr + S::BAR + 8u * XB_OH, (uint32_t)j);
  for (int r = warp; r < R; r += TF_NW) {
    const int gi = r / Tq;
    const bool in = lane < Tk;
    float s = -INFINITY;
    if (in) {
      // 64-deep dot product as four independent chains (the first version's single chain was 8 % of the kernel's stall samples)
      const float4* qr = reinterpret_cast<const float4*>(qs + r * 64);
      const float* kr = ks + (gi * Tk + lane) * TF_LDK;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int d = 0; d < TF_HD / 4; ++d) {
        const float4 qv = qr[d];
        s0 = fmaf(qv.x, kr[4 * d], s0);
        s1 = fmaf(qv.y, kr[4 * d + 1], s1);
        s2 = fmaf(qv.z, kr[4 * d + 2], s2);
        s3 = fmaf(qv.w, kr[4 * d + 3], s3);
      }
      if (valid == nullptr || valid[(size_t)(clip0 + gi) * Tk + lane]) s = ((s0 + s1) + (s2 + s3)) * 0.125f;
    }
    const float m = warp_max(s);
    const float p = (s == -INFINITY) ? 0.f : __expf(s - m);
    const float l = warp_sum(p);
    __syncwarp();
    pb[lane] = p;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    const float* vrow = vs + (gi * Tk) * 64 + 2 * lane;
    int j = 0;
    for (; j + 1 < Tk; j += 2) {  // p broadcast from shared memory, two independent accumulator pairs
      const float pa = pb[j], pc = pb[j + 1];
      const float2 va = *reinterpret_cast<const float2*>(vrow + j * 64);
      const float2 vc = *reinterpret_cast<const float2*>(vrow + (j + 1) * 64);
      o0 = fmaf(pa, va.x, o0);
      o1 = fmaf(pa, va.y, o1);
      o2 = fmaf(pc, vc.x, o2);
      o3 = fmaf(pc, vc.y, o3);
    }
    if (j < Tk) {
      const float pa = pb[j];
      const float2 va = *reinterpret_cast<const float2*>(vrow + j * 64);
      o0 = fmaf(pa, va.x, o0);
      o1 = fmaf(pa, va.y, o1);
    }
    // all keys masked: l = 0 -> 0 * inf = NaN, as torch's softmax over an all -inf row (and vmc_attention_masked)
    const float inv = 1.0f / l;
    const uint32_t h2 = tf_pack_h2((o0 + o2) * inv, (o1 + o3) * inv);
    const uint32_t dst = sm_addr + S::OH + (uint32_t)(r * TF_LDA + c * 64 + 2 * lane) * 2u;
#pragma unroll
    for (int jj = 0; jj < TF_NC; ++jj) tf_sta_u32(tf_mapa(dst, (uint32_t)jj), h2, rbar[jj]);
  }
}

template <int MT>
__global__ void __cluster_dims__(TF_NC, 1, 1) __launch_bounds__(TF_THREADS, 1)
tfam_fused_kernel(const __grid_constant__ TfArgs a) {
  using S = TfSmem<MT>;
  constexpr int RM = S::RM;
  extern __shared__ __align__(16) uint8_t sm[];
  const uint32_t sm_addr = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t c = tf_cta_rank();
  const int cluster_id = blockIdx.x / TF_NC, num_clusters = gridDim.x / TF_NC;
  const int n_groups = (a.B + a.G - 1) / a.G;
  const bool cross = a.mot != nullptr;

  __half* xh = reinterpret_cast<__half*>(sm + S::XH);
  __half* mh = reinterpret_cast<__half*>(sm + S::MH);
  __half* hh = reinterpret_cast<__half*>(sm + S::HH);
  float* x32 = reinterpret_cast<float*>(sm + S::X32);
  float* qs = reinterpret_cast<float*>(sm + S::QS);
  float* ks = reinterpret_cast<float*>(sm + S::KS);
  float* vs = reinterpret_cast<float*>(sm + S::VS);
  const float* recv = reinterpret_cast<const float*>(sm + S::RECV);
  float* pool = reinterpret_cast<float*>(sm + S::POOL);
  float* uh = reinterpret_cast<float*>(sm + S::UH);
  float* red = reinterpret_cast<float*>(sm + S::RED);

  // rows past the group's last one are multiplied too (M tiles of 16): keep every operand row finite
  for (uint32_t i = tid; i < S::X32 / 16; i += TF_THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < XB_COUNT; ++i) mbar_init(sm_addr + S::BAR + 8u * i, 1);  // one arrival: the local arrive.expect_tx
    fence_mbar_init();
  }
  uint32_t parity_bits = 0;  // phase parity of each exchange barrier (tracked identically by every thread)
  // the warp's weight stream: [CTA][warp][layers * 256 blocks][32 lanes]
  const uint4* wbase = a.wstream + ((size_t)(c * TF_NW + warp) * a.layers * NB_LAYER) * 32 + lane;
  uint4 ring[TF_PD];
#pragma unroll
  for (int j = 0; j < TF_PD; ++j) ring[j] = tf_ldg_w(wbase + (size_t)j * 32);
  __syncthreads();
  tf_cluster_sync();  // every CTA of the cluster is resident and its barriers initialised before the first remote store

  for (int grp = cluster_id; grp < n_groups; grp += num_clusters) {
    const int clip0 = grp * a.G;
    const int nclip = min(a.G, a.B - clip0);
    const int R = nclip * a.T, Rm = nclip * a.Tm;
    // ---- group input: fp32 rows -> fp16 operand (all 512 columns) + this CTA's fp32 residual columns ----
    {
      const float4* src = reinterpret_cast<const float4*>(a.x + (size_t)clip0 * a.T * TF_D);
      for (int i = tid; i < R * (TF_D / 4); i += TF_THREADS) {
        const int r = i >> 7, j = i & 127;
        const float4 v = __ldg(src + i);
        uint2 h;
        h.x = tf_pack_h2(v.x, v.y);
        h.y = tf_pack_h2(v.z, v.w);
        *reinterpret_cast<uint2*>(xh + r * TF_LDA + 4 * j) = h;
        if ((uint32_t)(j >> 4) == c) *reinterpret_cast<float4*>(x32 + r * 64 + 4 * (j & 15)) = v;
      }
      if (cross) {
        const float4* msrc = reinterpret_cast<const float4*>(a.mot + (size_t)clip0 * a.Tm * TF_D);
        for (int i = tid; i < Rm * (TF_D / 4); i += TF_THREADS) {
          const int r = i >> 7, j = i & 127;
          const float4 v = __ldg(msrc + i);
          uint2 h;
          h.x = tf_pack_h2(v.x, v.y);
          h.y = tf_pack_h2(v.z, v.w);
          *reinterpret_cast<uint2*>(mh + r * TF_LDA + 4 * j) = h;
        }
      }
    }
    __syncthreads();

    for (int l = 0; l < a.layers; ++l) {
      const TfLayer& ly = a.layer[l];
      const uint4* L = wbase + (size_t)l * NB_LAYER * 32;
      const uint4* w_sin = L;
      const uint4* w_sout = w_sin + NB_SIN * 32;
      const uint4* w_cq = w_sout + NB_SOUT * 32;
      const uint4* w_ckv = w_cq + NB_CQ * 32;
      const uint4* w_cout = w_ckv + NB_CKV * 32;
      const uint4* w_f1 = w_cout + NB_COUT * 32;
      const uint4* w_f2 = w_f1 + NB_F1 * 32;
      // after the last layer the stream wraps to layer 0: the next group (if any) starts with a primed ring
      const uint4* w_next_layer = (l + 1 < a.layers) ? L + (size_t)NB_LAYER * 32 : wbase;

      // ================= self-attention: x = LN(x + out_proj(MHA(x))) =================
      {  // q | k | v of head c: warp w computes columns 8 w .. 8 w + 7 of each
        float acc[MT][3][4];
        tf_gemm<MT, 3, 16>(sm_addr + S::XH, TF_LDA, w_sin, w_sout, ring, acc, lane);
        const int col = 8 * warp + 2 * t;
        const float2 bq = __ldg(reinterpret_cast<const float2*>(ly.b_sin + c * 64 + col));
        const float2 bk = __ldg(reinterpret_cast<const float2*>(ly.b_sin + TF_D + c * 64 + col));
        const float2 bv = __ldg(reinterpret_cast<const float2*>(ly.b_sin + 2 * TF_D + c * 64 + col));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int r = mt * 16 + g + 8 * hf;
            qs[r * 64 + col] = acc[mt][0][2 * hf] + bq.x;
            qs[r * 64 + col + 1] = acc[mt][0][2 * hf + 1] + bq.y;
            ks[r * TF_LDK + col] = acc[mt][1][2 * hf] + bk.x;
            ks[r * TF_LDK + col + 1] = acc[mt][1][2 * hf + 1] + bk.y;
            vs[r * 64 + col] = acc[mt][2][2 * hf] + bv.x;
            vs[r * 64 + col + 1] = acc[mt][2][2 * hf + 1] + bv.y;
          }
      }
      __syncthreads();
      tf_attention<MT>(sm, sm_addr, R, a.T, a.T, a.valid_x, clip0, c, warp, lane);
      tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_OH, (uint32_t)R * 1024u, tid);  // all eight heads' output rows are here
      {  // this CTA's 64 columns of out_proj + residual
        float acc[MT][1][4];
        tf_gemm<MT, 1, 16>(sm_addr + S::OH, TF_LDA, w_sout, cross ? w_cq : w_f1, ring, acc, lane);
        const int col = 8 * warp + 2 * t;
        const float2 bo = __ldg(reinterpret_cast<const float2*>(ly.b_sout + c * 64 + col));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int r = mt * 16 + g + 8 * hf;
            x32[r * 64 + col] += acc[mt][0][2 * hf] + bo.x;
            x32[r * 64 + col + 1] += acc[mt][0][2 * hf + 1] + bo.y;
          }
      }
      tf_layernorm_exchange<MT>(sm, sm_addr, R, c, ly.ns_g, ly.ns_b, ly.ns_eps, warp, lane, tid, parity_bits);

      // ================= cross-attention: x = LN(x + out_proj(MHA(q = x, k = v = motion))) =================
      if (cross) {
        {
          float acc[MT][1][4];
          tf_gemm<MT, 1, 16>(sm_addr + S::XH, TF_LDA, w_cq, w_ckv, ring, acc, lane);
          const int col = 8 * warp + 2 * t;
          const float2 bq = __ldg(reinterpret_cast<const float2*>(ly.b_cin + c * 64 + col));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = mt * 16 + g + 8 * hf;
              qs[r * 64 + col] = acc[mt][0][2 * hf] + bq.x;
              qs[r * 64 + col + 1] = acc[mt][0][2 * hf + 1] + bq.y;
            }
        }
        {
          float acc[MT][2][4];
          tf_gemm<MT, 2, 16>(sm_addr + S::MH, TF_LDA, w_ckv, w_cout, ring, acc, lane);
          const int col = 8 * warp + 2 * t;
          const float2 bk = __ldg(reinterpret_cast<const float2*>(ly.b_cin + TF_D + c * 64 + col));
          const float2 bv = __ldg(reinterpret_cast<const float2*>(ly.b_cin + 2 * TF_D + c * 64 + col));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = mt * 16 + g + 8 * hf;
              ks[r * TF_LDK + col] = acc[mt][0][2 * hf] + bk.x;
              ks[r * TF_LDK + col + 1] = acc[mt][0][2 * hf + 1] + bk.y;
              vs[r * 64 + col] = acc[mt][1][2 * hf] + bv.x;
              vs[r * 64 + col + 1] = acc[mt][1][2 * hf + 1] + bv.y;
            }
        }
        __syncthreads();
        tf_attention<MT>(sm, sm_addr, R, a.T, a.Tm, a.valid_m, clip0, c, warp, lane);
        tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_OH, (uint32_t)R * 1024u, tid);
        {
          float acc[MT][1][4];
          tf_gemm<MT, 1, 16>(sm_addr + S::OH, TF_LDA, w_cout, w_f1, ring, acc, lane);
          const int col = 8 * warp + 2 * t;
          const float2 bo = __ldg(reinterpret_cast<const float2*>(ly.b_cout + c * 64 + col));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = mt * 16 + g + 8 * hf;
              x32[r * 64 + col] += acc[mt][0][2 * hf] + bo.x;
              x32[r * 64 + col + 1] += acc[mt][0][2 * hf + 1] + bo.y;
            }
        }
        tf_layernorm_exchange<MT>(sm, sm_addr, R, c, ly.nc_g, ly.nc_b, ly.nc_eps, warp, lane, tid, parity_bits);
      }

      // ================= feed-forward: x = LN(x + W2 act(W1 x + b1) + b2) =================
      {  // hidden units [256 c, 256 c + 256): warp w computes 32 of them
        float acc[MT][4][4];
        tf_gemm<MT, 4, 16>(sm_addr + S::XH, TF_LDA, w_f1, w_f2, ring, acc, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = 32 * warp + 8 * i + 2 * t;
          const float2 b1 = __ldg(reinterpret_cast<const float2*>(ly.b_f1 + c * TF_FFC + col));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = mt * 16 + g + 8 * hf;
              float h0 = acc[mt][i][2 * hf] + b1.x, h1 = acc[mt][i][2 * hf + 1] + b1.y;
              if (a.act == VMC_ACT_GELU_ERF) {
                h0 = 0.5f * h0 * (1.0f + erff(h0 * 0.70710678118654752440f));
                h1 = 0.5f * h1 * (1.0f + erff(h1 * 0.70710678118654752440f));
              } else {
                h0 = fmaxf(h0, 0.f);
                h1 = fmaxf(h1, 0.f);
              }
              *reinterpret_cast<uint32_t*>(hh + r * TF_LDH + col) = tf_pack_h2(h0, h1);
            }
        }
      }
      __syncthreads();
      {  // partial sums over this CTA's 256 hidden units for ALL 512 outputs: warp w computes the 64 columns CTA w owns
        float acc[MT][8][4];
        tf_gemm<MT, 8, 8>(sm_addr + S::HH, TF_LDH, w_f2, w_next_layer, ring, acc, lane);
        const uint32_t dst_cta = tf_mapa(sm_addr + S::RECV, (uint32_t)warp);
        const uint32_t dst_bar = tf_mapa(sm_addr + S::BAR + 8u * XB_RECV, (uint32_t)warp);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = mt * 16 + g + 8 * hf;
              if (r < R)  // the receiver expects exactly R rows from every CTA
                tf_sta_f32x2(dst_cta + (uint32_t)((c * RM + r) * 64 + 8 * i + 2 * t) * 4u, acc[mt][i][2 * hf],
                             acc[mt][i][2 * hf + 1], dst_bar);
            }
      }
      tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_RECV, (uint32_t)R * 2048u, tid);  // every CTA's partial sums have arrived
      {
        const float2 b2 = __ldg(reinterpret_cast<const float2*>(ly.b_f2 + c * 64) + lane);
        for (int r = warp; r < R; r += TF_NW) {
          float2 v = *reinterpret_cast<const float2*>(x32 + r * 64 + 2 * lane);
          v.x += b2.x;
          v.y += b2.y;
#pragma unroll
          for (int s = 0; s < TF_NC; ++s) {  // fixed order: deterministic
            const float2 p = *reinterpret_cast<const float2*>(recv + (s * RM + r) * 64 + 2 * lane);
            v.x += p.x;
            v.y += p.y;
          }
          *reinterpret_cast<float2*>(x32 + r * 64 + 2 * lane) = v;
        }
      }
      tf_layernorm_exchange<MT>(sm, sm_addr, R, c, ly.nf_g, ly.nf_b, ly.nf_eps, warp, lane, tid, parity_bits);
    }

    // ================= head: mean over ALL T rows -> LayerNorm -> Linear -> GELU(erf) -> Linear =================
    if (tid < nclip * 64) {
      const int gi = tid >> 6, col = tid & 63;
      float s = 0.f;
      for (int r = gi * a.T; r < (gi + 1) * a.T; ++r) s += x32[r * 64 + col];
      s /= (float)a.T;
      const uint32_t dst = sm_addr + S::POOL + (uint32_t)(gi * TF_D + c * 64 + col) * 4u;
#pragma unroll
      for (int j = 0; j < TF_NC; ++j)
        tf_sta_f32(tf_mapa(dst, (uint32_t)j), s, tf_mapa(sm_addr + S::BAR + 8u * XB_POOL, (uint32_t)j));
    }
    tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_POOL, (uint32_t)nclip * 2048u, tid);
    if (warp < nclip) {  // LayerNorm of the pooled row (every CTA redundantly: 512 values)
      float* pr = pool + warp * TF_D;
      float v[16];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = pr[lane + 32 * i];
        s += v[i];
      }
      const float mean = warp_sum(s) * (1.0f / TF_D);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) q += (v[i] - mean) * (v[i] - mean);
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / TF_D) + a.hd_eps);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = lane + 32 * i;
        pr[k] = (v[i] - mean) * rstd * __ldg(a.hd_g + k) + __ldg(a.hd_b + k);
      }
    }
    __syncthreads();
    const int o = tid & 31, kp = tid >> 5;  // thread = (k slice, output unit): weight reads coalesced over the units
    {  // Linear 512 -> hidden: this CTA computes hidden / 8 units (32 for hidden = 256)
      const int per = a.hidden / TF_NC;  // <= 32
      for (int gi = 0; gi < nclip; ++gi) {
        float s0 = 0.f, s1 = 0.f;
        if (o < per) {
          const float* w = a.w1t + (size_t)(kp * 64) * a.hidden + c * per + o;
          const float* pr = pool + gi * TF_D + kp * 64;
#pragma unroll 16
          for (int k = 0; k < 64; k += 2) {
            s0 = fmaf(pr[k], __ldg(w + (size_t)k * a.hidden), s0);
            s1 = fmaf(pr[k + 1], __ldg(w + (size_t)(k + 1) * a.hidden), s1);
          }
        }
        red[kp * 32 + o] = s0 + s1;
        __syncthreads();
        if (tid < per) {
          float u = __ldg(a.b1 + c * per + tid);
#pragma unroll
          for (int k = 0; k < 8; ++k) u += red[k * 32 + tid];
          u = 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f));
          const uint32_t dst = sm_addr + S::UH + (uint32_t)(gi * 256 + c * per + tid) * 4u;
#pragma unroll
          for (int j = 0; j < TF_NC; ++j)
            tf_sta_f32(tf_mapa(dst, (uint32_t)j), u, tf_mapa(sm_addr + S::BAR + 8u * XB_UH, (uint32_t)j));
        }
        __syncthreads();
      }
    }
    tf_xchg_wait(sm_addr + S::BAR, parity_bits, XB_UH, (uint32_t)(nclip * a.hidden) * 4u, tid);
    {  // Linear hidden -> C: this CTA computes classes [c * cper, (c + 1) * cper)
      const int cper = (a.C + TF_NC - 1) / TF_NC;
      const int ksl = a.hidden / 8;  // k slice of one warp
      for (int gi = 0; gi < nclip; ++gi)
        for (int o0 = 0; o0 < cper; o0 += 32) {
          const int cls = (int)c * cper + o0 + o;
          const bool ok = (o0 + o) < cper && cls < a.C;
          float s0 = 0.f, s1 = 0.f;
          if (ok) {
            const float* w = a.w2t + (size_t)(kp * ksl) * a.C + cls;
            const float* ur = uh + gi * 256 + kp * ksl;
#pragma unroll 8
            for (int k = 0; k + 1 < ksl; k += 2) {
              s0 = fmaf(ur[k], __ldg(w + (size_t)k * a.C), s0);
              s1 = fmaf(ur[k + 1], __ldg(w + (size_t)(k + 1) * a.C), s1);
            }
            if (ksl & 1) s0 = fmaf(ur[ksl - 1], __ldg(w + (size_t)(ksl - 1) * a.C), s0);
          }
          red[kp * 32 + o] = s0 + s1;
          __syncthreads();
          if (tid < 32 && ok) {
            float u = __ldg(a.b2 + cls);
#pragma unroll
            for (int k = 0; k < 8; ++k) u += red[k * 32 + tid];
            a.logits[(size_t)(clip0 + gi) * a.C + cls] = u;
          }
          __syncthreads();
        }
    }
  }
  tf_cluster_sync();  // no CTA exits while a peer may still store into its shared memory
}

int tf_max_clusters(int mt, size_t smem) {
  static std::mutex mu;
  static int cached[64][2];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 16;
  std::lock_guard<std::mutex> lock(mu);
  int& slot = cached[dev][mt - 1];
  if (slot > 0) return slot;
  // co-resident clusters of 8 CTAs (one CTA per SM: ~220 KB of shared memory).  More clusters than that would only queue
  // (clusters never wait for one another), so an over-estimate costs balance, not correctness.
  int n = 0;
  cudaError_t e = mt == 1 ? cudaFuncSetAttribute(tfam_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(tfam_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TF_NC * 32, 1, 1);
    cfg.blockDim = dim3(TF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = TF_NC;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    e = mt == 1 ? cudaOccupancyMaxActiveClusters(&n, tfam_fused_kernel<1>, &cfg)
                : cudaOccupancyMaxActiveClusters(&n, tfam_fused_kernel<2>, &cfg);
  }
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    n = vmc_num_sms() / TF_NC - 2;  // 148 SMs in GPCs of 16-20: two clusters per GPC
    if (n < 1) n = 1;
  }
  slot = n;
  return n;
}

}  // namespace

extern "C" {

long long vmc_tfam_wstream_bytes(int layers) {
  if (layers <= 0 || layers > VMC_TFAM_MAX_LAYERS) return -1;
  return (long long)TF_NC * TF_NW * layers * NB_LAYER * 512;
}

int vmc_tfam_fused_supported(const vmc_tfam_model* m, int B, int T, int Tm) {
  if (!m || m->d_model != TF_D || m->nhead != TF_NC || m->dim_ff != TF_FF || m->layers <= 0 ||
      m->layers > VMC_TFAM_MAX_LAYERS || m->hidden <= 0 || m->hidden > 256 || (m->hidden % TF_NC) != 0 ||
      m->num_classes <= 0)
    return 0;
  if (B <= 0 || T <= 0 || T > 32 || Tm < 0 || Tm > 32) return 0;
  return 1;
}

int vmc_tfam_forward(const vmc_tfam_model* m, const float* x, const float* motion, const uint8_t* valid_x,
                     const uint8_t* valid_m, float* logits, int B, int T, int Tm, void* stream) {
  VMC_CHECK_ARG(m && x && logits && m->wstream && m->layer, VMC_ERR_ARG, "vmc_tfam_forward: null pointer");
  VMC_CHECK_ARG(vmc_tfam_fused_supported(m, B, T, motion ? Tm : 0), VMC_ERR_SHAPE,
                "vmc_tfam_forward: the fused kernel covers d_model 512, nhead 8, dim_feedforward 2048, <= %d layers, "
                "classifier hidden <= 256 and at most 32 frames per clip (got d %d, heads %d, ff %d, layers %d, T %d, Tm %d)",
                VMC_TFAM_MAX_LAYERS, m->d_model, m->nhead, m->dim_ff, m->layers, T, Tm);
  VMC_CHECK_ARG(motion == nullptr || Tm > 0, VMC_ERR_SHAPE, "vmc_tfam_forward: motion given with Tm = %d", Tm);
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(motion) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(m->wstream) & 15) == 0,
                VMC_ERR_ALIGN, "vmc_tfam_forward: x / motion / wstream must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TfArgs a = {};
  a.x = x;
  a.mot = motion;
  a.valid_x = valid_x;
  a.valid_m = motion ? valid_m : nullptr;
  a.logits = logits;
  a.wstream = reinterpret_cast<const uint4*>(m->wstream);
  a.B = B;
  a.T = T;
  a.Tm = motion ? Tm : 0;
  a.C = m->num_classes;
  a.hidden = m->hidden;
  a.layers = m->layers;
  a.act = m->act;
  a.hd_g = m->cls_ln_g;
  a.hd_b = m->cls_ln_b;
  a.hd_eps = m->cls_ln_eps;
  a.w1t = m->w1t;
  a.b1 = m->b1;
  a.w2t = m->w2t;
  a.b2 = m->b2;
  for (int l = 0; l < m->layers; ++l) {
    const vmc_tfam_layer& s = m->layer[l];
    VMC_CHECK_ARG(s.b_sin && s.b_sout && s.b_f1 && s.b_f2 && s.ns_g && s.ns_b && s.nf_g && s.nf_b &&
                      (!motion || (s.b_cin && s.b_cout && s.nc_g && s.nc_b)),
                  VMC_ERR_ARG, "vmc_tfam_forward: layer %d has null parameters", l);
    TfLayer& d = a.layer[l];
    d.b_sin = s.b_sin; d.b_sout = s.b_sout; d.b_cin = s.b_cin; d.b_cout = s.b_cout; d.b_f1 = s.b_f1; d.b_f2 = s.b_f2;
    d.ns_g = s.ns_g; d.ns_b = s.ns_b; d.nc_g = s.nc_g; d.nc_b = s.nc_b; d.nf_g = s.nf_g; d.nf_b = s.nf_b;
    d.ns_eps = s.ns_eps; d.nc_eps = s.nc_eps; d.nf_eps = s.nf_eps;
  }
  const int tmax = T > a.Tm ? T : a.Tm;
  // Clips per pass: two (32 rows, two M tiles: the weight stream is read once per two clips) when the batch is large
  // enough to keep every cluster busy anyway; one clip per cluster in the latency regime.
  const int clusters1 = tf_max_clusters(1, TfSmem<1>::TOTAL);
  int G = 1, mt = tmax <= 16 ? 1 : 2;
  if (tmax <= 16 && B > 2 * clusters1) {
    G = 2;
    mt = 2;
  }
  a.G = G;
  const size_t smem = mt == 1 ? TfSmem<1>::TOTAL : TfSmem<2>::TOTAL;
  const int n_groups = (B + G - 1) / G;
  const int maxc = tf_max_clusters(mt, smem);
  const int nclusters = n_groups < maxc ? n_groups : maxc;
  {
    // algorithmic FLOPs: 2 * rows * params of the projections; bytes: the weight stream once per pass of every cluster
    const double params = (double)m->layers * ((motion ? 2.0 : 1.0) * 4.0 * TF_D * TF_D + 2.0 * TF_D * TF_FF);
    VmcProfScope prof(VMC_K_ATTN_SMALL, st, 2.0 * B * T * params, (double)n_groups * params * 2.0);
    if (mt == 1) {
      VMC_CUDA(cudaFuncSetAttribute(tfam_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tfam_fused_kernel<1><<<nclusters * TF_NC, TF_THREADS, smem, st>>>(a);
    } else {
      VMC_CUDA(cudaFuncSetAttribute(tfam_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tfam_fused_kernel<2><<<nclusters * TF_NC, TF_THREADS, smem, st>>>(a);
    }
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // extern "C"
