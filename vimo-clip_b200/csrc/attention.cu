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
// (Generations 3 and 4 -- the first persistent pipeline with 8 / 16 softmax warps marching in lockstep -- were measured
// against v5 in round 1 and removed from the library in round 2.)
// ---------------------------------------------------------------------------------------
// max over one 32-column chunk of a score row (columns >= valid are padding)
template <bool MASK>
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], float mx, int valid) {
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (!MASK || j < valid) mx = fmaxf(mx, __uint_as_float(r[j]));
  return mx;
}
// ---------------------------------------------------------------------------------------
// A1, fifth generation (129 <= L <= 224): the two query tiles run OUT OF PHASE and the softmax warps do
// nothing but softmax.
//
// What the v3 timeline showed: both tiles march in lockstep through  S MMA -> softmax -> PV -> O drain ->
// next S MMA, because one MMA thread serves them in a fixed order, and the ~3k-cycle tail (PV part B,
// accumulator drain, next S MMA) of BOTH tiles sits idle on the MUFU pipe at the same time (XU 54 % busy).
// TMEM (512 columns) cannot hold a third score tile, so the bubble is hidden by phase instead:
//   * warps 13 / 14 are the MMA issuers, one per query tile, each blocking on its own tile's barriers;
//   * warps 8-11 are EPILOGUE warps (one per TMEM lane quarter, both tiles): they drain O, release the
//     tile's TMEM and write the output rows, so a softmax warp goes straight from its last chunk to the
//     next item's scores;
//   * the softmax normaliser comes from the tensor core: the PV MMA runs with N = 80, columns 64..79 of
//     its B operand being a constant all-ones tile reached through the descriptor's leading-dimension
//     offset, so column 64 of the accumulator is the row sum of exactly the bf16 probabilities the MMA
//     consumed.  This removes the rounded-sum bookkeeping (2 LOP + 2 FADD of 13 instructions per pair) and
//     the softmax -> epilogue hand-over of the row sums.  The rows of that tile which belong to padding keys
//     (>= L) are ZERO, like the zero-filled V rows, so whatever P holds for them contributes nothing;
//   * [r2] the score MMA of an item is issued in two key ranges, the first one ahead of time (see the MMA
//     issuers).
// TMEM plan of one query tile (base = 256 t, n_chunks = ceil(Lk16 / 32) in 5..7):
//   S_a  [0, 64)   S_b [64, Lk16)   fp32 scores of keys 0..63 / 64..
//   P_a  [16 c, 16 c + 16)      bf16 pairs of chunk c < 5 (keys 0..159), over S columns already consumed
//   O    [80, 144) sum [144,160) fp32 accumulator of the N = 80 PV MMA, over consumed S columns
//   P_b  [32 c, 32 c + 16)      chunk c >= 5, in place over its own scores
// ---------------------------------------------------------------------------------------
// VAR: 0 = production; what-if timing variants (WRONG results; make WHATIF=1, tools/kernel_bench.py only):
//      1 = softmax warps only signal (pipeline floor), 2 = TMEM load + store without the math, 3 = FMA instead of MUFU,
//      4 = no loads after the first items, 5 = 1 + 4, 6 = no output stores
template <bool MASK, int VAR = 0>
__device__ __forceinline__ void chunk_exp_store5(const uint32_t (&r)[32], float sc, float mxs, int valid,
                                                 uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float e0 = 0.f, e1 = 0.f;
    if (VAR == 2) {
      e0 = __uint_as_float(r[2 * j]);
      e1 = __uint_as_float(r[2 * j + 1]);
    } else if (!MASK || 2 * j < valid) {  // warp-uniform: padding columns cost no MUFU work
      if (VAR == 3) {
        e0 = fmaf(fminf(fmaf(__uint_as_float(r[2 * j]), sc, -mxs), 120.f), 0.001f, 1.0f);
        e1 = fmaf(fminf(fmaf(__uint_as_float(r[2 * j + 1]), sc, -mxs), 120.f), 0.001f, 1.0f);
      } else {
        e0 = ex2_approx(fminf(fmaf(__uint_as_float(r[2 * j]), sc, -mxs), 120.f));
        e1 = ex2_approx(fminf(fmaf(__uint_as_float(r[2 * j + 1]), sc, -mxs), 120.f));
      }
      if (MASK && 2 * j + 1 >= valid) e1 = 0.f;
    }
    // one F2FP per pair (round to nearest even).  Round 1 rounded in the integer ALU (2 IADD + PRMT per pair) to keep the
    // conversion off the MUFU pipe; with the softmax warps bound by their own instruction stream rather than by that pipe
    // (round-2 timeline) the shorter sequence is 1 - 3 % faster.
    pk[j] = pack_bf16x2(e0, e1);
  }
  tmem_st_32x32b_x16(taddr, pk);
}

struct Attn5Args {
  int L, heads, d, lk16, n_items;
  int qk_stages;  // depth of the Q / K ring: 3 where shared memory allows (<= 208 keys), else 2
  __nv_bfloat16* out;
  long long* dbg;  // timeline of CTA 0 (DBG instantiation only; tools/attn_timeline.py)
  int cq, ck, cv, chead;  // column of head h's Q / K / V slice = c{q,k,v} + h * chead
};
#define VMC_DBG5(k_, slot)                                                                     \
  do {                                                                                         \
    if (DBG && a.dbg != nullptr && blockIdx.x == 0 && (k_) < 16) a.dbg[((k_) * 32) + (slot)] = clock64(); \
  } while (0)

constexpr int A5_PA_CHUNKS = 5;  // chunks (32 keys) whose P is packed into [0, 80): part A of the PV MMA
constexpr int A5_OCOL = 80;      // accumulator columns [80, 160): 64 of O + 16 copies of the row sum
constexpr int A5B_THREADS = 480;  // 8 softmax + 4 epilogue + TMA + 2 MMA issuer warps (one per query tile)

template <int VAR, bool DBG = false>
__global__ void __launch_bounds__(A5B_THREADS, 1)
attention_vit5_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm1,
                      const Attn5Args a) {
  // Shared memory: two rings, released at different times.  The timeline of the first v5 (one Q/K/V stage
  // per item, freed when the item's last PV MMA retires) showed the NEXT-NEXT item's load being issued
  // ~6k cycles (its own latency under load) before tile 0 needed it: the kernel was bound by the latency
  // of one 96 KB load per SM.  Q and K are dead as soon as both score MMAs of the item have retired, i.e. a
  // whole softmax earlier, so they get their own ring; the second tiles are loaded with (Lk16 - 128)-row
  // boxes instead of 128-row ones (rows >= L are zero-filled by TMA either way).
  // [r2] The all-ones tile shrank from one 2 KB block per 16-key step to two blocks (all keys real / the last, partly
  // padded step), which leaves room for a THIRD Q / K stage up to 208 keys.  With the score MMA issued early the timeline
  // shows a load landing 8 - 9.5 k cycles after its issue (every SM pulls its share of ~4.5 TB/s) and the issuer waiting
  // ~1.1 k cycles per item for it -- yet the deeper ring measured the same (0.281 vs 0.279 ms per 1024 frames): the wait
  // moves to the next link of the chain.  The depth is a kernel argument: 3 where it fits, else 2 (both are tested).
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const int r1 = a.lk16 - 128;                          // rows of the second Q / K / V box (16..96)
  const uint32_t mat_bytes = (uint32_t)(128 + r1) * 128u;  // one of Q, K, V for one item
  const uint32_t qk_base = base;                                      // ring of qk_stages x [Q | K]
  const uint32_t v_base = base + 2u * a.qk_stages * mat_bytes;        // ring of 2 x V
  const uint32_t ones_base = v_base + 2 * mat_bytes;                  // block 0: ones; block 1: ones for keys < L of the last step
  const int nkk = a.lk16 / 16;  // 16-key steps of the PV MMA
  const uint32_t bar_base = ones_base + 4096u;
  auto qk_full = [&](int s) { return bar_base + 8u * s; };            // 3
  auto qk_empty = [&](int s) { return bar_base + 24u + 8u * s; };     // 3
  auto s_full = [&](int t) { return bar_base + 48u + 8u * t; };       // scores of keys 0..63
  auto sb_full = [&](int t) { return bar_base + 64u + 8u * t; };      // scores of keys 64..
  auto pa_full = [&](int t) { return bar_base + 80u + 8u * t; };
  auto pb_full = [&](int t) { return bar_base + 96u + 8u * t; };
  auto o_full = [&](int t) { return bar_base + 112u + 8u * t; };
  auto s_empty = [&](int t) { return bar_base + 128u + 8u * t; };
  auto v_full = [&](int s) { return bar_base + 144u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 160u + 8u * s; };
  const uint32_t tmem_ptr_addr = bar_base + 176u;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_my = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA

  // the "second MN atom" of the PV B operand: bf16 1.0 for every real key.  One 128-byte row per key (the swizzle only
  // permutes the 16-byte chunks inside a row, and every element of a row is the same), 16 rows per block.
  {
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (ones_base - raw_addr));
    const int last_valid = a.L - (nkk - 1) * 16;  // real keys of the last step (1..16)
    for (uint32_t i = tid; i < 4096u / 16; i += A5B_THREADS) {
      const int row = (int)(i >> 3) & 15;
      const uint32_t one2 = (i < 128u || row < last_valid) ? 0x3F803F80u : 0u;
      ones[i] = make_uint4(one2, one2, one2, one2);
    }
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(qk_full(i), 1);
      mbar_init(qk_empty(i), 2);  // one tcgen05.commit per tile's MMA issuer
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(v_full(i), 1);
      mbar_init(v_empty(i), 2);
      mbar_init(s_full(i), 1);
      mbar_init(sb_full(i), 1);
      mbar_init(pa_full(i), 4);  // one arrive per softmax warp of the tile
      mbar_init(pb_full(i), 4);
      mbar_init(o_full(i), 1);
      mbar_init(s_empty(i), 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;
  const int n_chunks = (a.lk16 + 31) / 32;  // 5..7

  if (warp == 12) {
    // ===================== TMA producer =====================
    // The rings are released in a fixed order -- Q/K of item k (both tiles' score MMAs retired), then V of item k (both PV
    // MMAs), then Q/K of item k + 1 (its S_b needs the drained accumulators of item k) -- so the producer issues in that
    // order and BLOCKS on the next release.  (Round 1 polled both barriers; measured the same: 0.294 vs 0.296 ms.)
    if (lane == 0) {
      int kq = 0, kv = 0;          // next Q/K and V loads
      int sq = 0;                  // Q/K ring slot and its use count parity
      uint32_t pq = 0;
      while (kq < n_my || kv < n_my) {
        const bool do_qk = kq < n_my && (kv >= n_my || kq < kv + a.qk_stages - 1);  // Q/K runs (stages - 1) items ahead of V
        if (do_qk) {
          mbar_wait(qk_empty(sq), pq ^ 1u);
          const int item = blockIdx.x + kq * gridDim.x;
          const int head = item % a.heads, frame = item / a.heads;
          const uint32_t q = qk_base + sq * 2 * mat_bytes, kk_ = q + mat_bytes;
          VMC_DBG5(kq, 24);
          if (VAR >= 4 && VAR <= 5 && kq >= a.qk_stages) {  // what-if: ring contents reused, no HBM traffic
            mbar_arrive(qk_full(sq));
          } else {
            mbar_arrive_expect_tx(qk_full(sq), 2 * mat_bytes);
            tma_load_3d(q, &tm, qk_full(sq), a.cq + head * a.chead, 0, frame);
            tma_load_3d(q + TILE, &tm1, qk_full(sq), a.cq + head * a.chead, 128, frame);
            tma_load_3d(kk_, &tm, qk_full(sq), a.ck + head * a.chead, 0, frame);
            tma_load_3d(kk_ + TILE, &tm1, qk_full(sq), a.ck + head * a.chead, 128, frame);
          }
          ++kq;
          if (++sq == a.qk_stages) {
            sq = 0;
            pq ^= 1u;
          }
        } else {
          const int s = kv & 1;
          mbar_wait(v_empty(s), (((uint32_t)kv >> 1) & 1u) ^ 1u);
          const int item = blockIdx.x + kv * gridDim.x;
          const int head = item % a.heads, frame = item / a.heads;
          const uint32_t v = v_base + s * mat_bytes;
          VMC_DBG5(kv, 26);
          if (VAR >= 4 && VAR <= 5 && kv >= 2) {
            mbar_arrive(v_full(s));
          } else {
            mbar_arrive_expect_tx(v_full(s), mat_bytes);
            tma_load_3d(v, &tm, v_full(s), a.cv + head * a.chead, 0, frame);
            tma_load_3d(v + TILE, &tm1, v_full(s), a.cv + head * a.chead, 128, frame);
          }
          ++kv;
        }
      }
    }
    __syncwarp();
  } else if (warp == 13 || warp == 14) {
    // ===================== MMA issuers: one warp per query tile =====================
    // Round 1 had ONE thread serve both tiles by polling their barriers round-robin with mbarrier.test_wait (~150 cycles
    // per probe, two to four probes per turn): every hand-over on a tile's S -> softmax -> PV -> drain chain waited for the
    // poller to come round, and an MMA burst for one tile delayed the other.  Each tile now has its own issuer that BLOCKS
    // on that tile's barriers (try_wait: the thread sleeps in hardware and wakes ~60 cycles after the arrive).  The shared
    // rings are released by both: qk_empty / v_empty count two arrivals, each issuer's tcgen05.commit covering its own MMAs.
    //
    // The chain  S MMA -> softmax -> PV -> accumulator drain -> next S MMA  still left the tile's four softmax warps idle
    // for ~2 k cycles between their last probability chunk and the next item's scores (timeline: PV part B 0.4 k, drain
    // 0.1 k, wait + issue 0.3 k, score MMA 0.7 - 1.3 k: its operands come from shared memory at ~100 B/clk next to the TMA
    // writes).  The score MMA is therefore issued in two key ranges:
    //   S_a (keys 0..63 -> columns [0, 64))    right after PV part A of the PREVIOUS item: those columns then hold only
    //       P_a, which that MMA -- issued by this same thread, so ordered before S_a -- is the last to read;
    //   S_b (keys 64..  -> columns [64, Lk16)) once the accumulator (columns 80..159) has been drained, as before.
    // The softmax warps go from the last chunk of item k straight to chunks 0 and 1 of item k + 1, and PV part B, the
    // drain and S_b run underneath those two chunks.
    const int t = warp - 13;
    if (lane == 0) {
      const uint32_t idesc_sa = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_sb = umma_idesc_bf16(128, a.lk16 - 64, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 80, 0, 1);  // B = [V | ones], MN-major
      const int nka = nkk < 2 * A5_PA_CHUNKS ? nkk : 2 * A5_PA_CHUNKS;
      const bool last_partial = (a.L & 15) != 0;
      const uint32_t tcol = tmem_base + uint32_t(t * 256);
      int sq = 0;  // Q/K ring slot / parity of the next S_a
      uint32_t pq = 0;
      int sq_b = 0;  // ... of the next S_b (one item behind S_a in the steady state)
      auto issue_sa = [&](int k) {
        mbar_wait(qk_full(sq), pq);
        tc_fence_after();
        if (t == 0) VMC_DBG5(k, 25);
        VMC_DBG5(k, 0 + t);
        const uint32_t qst = qk_base + sq * 2 * mat_bytes;
        const uint64_t dq = umma_desc_sw128(qst + t * TILE);
        const uint64_t dk = umma_desc_sw128(qst + mat_bytes);
#pragma unroll
        for (int kq = 0; kq < HD / 16; ++kq)
          umma_ss(tcol, dq + uint64_t(2 * kq), dk + uint64_t(2 * kq), idesc_sa, kq != 0);
        umma_commit(s_full(t));
        if (++sq == a.qk_stages) {
          sq = 0;
          pq ^= 1u;
        }
      };
      auto issue_sb = [&]() {
        const uint32_t qst = qk_base + sq_b * 2 * mat_bytes;
        const uint64_t dq = umma_desc_sw128(qst + t * TILE);
        const uint64_t dk = umma_desc_sw128(qst + mat_bytes + 64u * 128u);  // K rows 64..: 8 KB further (swizzle-atom aligned)
#pragma unroll
        for (int kq = 0; kq < HD / 16; ++kq)
          umma_ss(tcol + 64u, dq + uint64_t(2 * kq), dk + uint64_t(2 * kq), idesc_sb, kq != 0);
        umma_commit(sb_full(t));
        umma_commit(qk_empty(sq_b));  // second arrival (the other tile's issuer) frees Q and K
        if (++sq_b == a.qk_stages) sq_b = 0;
      };
      if (n_my > 0) {
        issue_sa(0);
        issue_sb();
      }
      for (int k = 0; k < n_my; ++k) {
        const int s = k & 1;
        const uint32_t par = (uint32_t)k & 1u;
        const uint32_t ring_par = ((uint32_t)k >> 1) & 1u;
        const uint32_t vst = v_base + s * mat_bytes;
        // V descriptor of 16-key step kk: atom 0 = rows 16 kk.. of the TMA tile (64 head dims x 16 keys), atom 1 (N = 64..79)
        // = the leading-dimension offset further = one of the two all-ones blocks
        const uint64_t dv_base = umma_desc_sw128(vst) & ~(uint64_t(0x3FFF) << 16);
        auto dv = [&](int kk) {
          const uint32_t blk = ones_base + ((last_partial && kk == nkk - 1) ? 2048u : 0u);
          return (dv_base + uint64_t(kk * 128)) | (uint64_t((blk - (vst + uint32_t(kk) * 2048u)) >> 4) << 16);
        };
        // ---- O = P V, keys of part A (their probabilities are stored while the softmax still works on part B) ----
        mbar_wait(pa_full(t), par);
        mbar_wait(v_full(s), ring_par);
        tc_fence_after();
        VMC_DBG5(k, 2 + t);
        for (int kk = 0; kk < nka; ++kk) umma_ts(tcol + A5_OCOL, tcol + uint32_t(kk * 8), dv(kk), idesc_pv, kk != 0);
        if (k + 1 < n_my) issue_sa(k + 1);
        // ---- part B ----
        mbar_wait(pb_full(t), par);
        tc_fence_after();
        VMC_DBG5(k, 4 + t);
        for (int kk = nka; kk < nkk; ++kk)
          umma_ts(tcol + A5_OCOL, tcol + uint32_t((kk >> 1) * 32 + (kk & 1) * 8), dv(kk), idesc_pv, 1u);
        umma_commit(o_full(t));
        umma_commit(v_empty(s));  // second arrival frees V
        if (k + 1 < n_my) {
          mbar_wait(s_empty(t), par);  // accumulator of item k drained
          tc_fence_after();
          issue_sb();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ===================== epilogue warps: drain O, release the tile, write the rows =====================
    const int q = warp - 8;
    for (int k = 0; k < n_my; ++k) {
      const int item = blockIdx.x + k * gridDim.x;
      const int head = item % a.heads;
      const int frame = item / a.heads;
      const uint32_t par = (uint32_t)k & 1u;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {  // tile 0 leads tile 1 by half a period, so this is the completion order
        const bool warp_active = (t * 128 + q * 32) < a.L;  // warp-uniform
        const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
        mbar_wait(o_full(t), par);
        tc_fence_after();
        if (q == 0 && lane == 0) VMC_DBG5(k, 11 + 8 * t);
        uint32_t o0[32], o1[32];
        uint32_t rs = 0x3F800000u;
        if (warp_active) {
          tmem_ld_32x32b_x32(tb + A5_OCOL, o0);
          tmem_ld_32x32b_x32(tb + A5_OCOL + 32u, o1);
          rs = tmem_ld_32x32b_x1(tb + A5_OCOL + 64u);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(s_empty(t));
        if (q == 0 && lane == 0) VMC_DBG5(k, 12 + 8 * t);
        if (warp_active && (t * 128 + q * 32 + lane) < a.L && VAR != 6) {
          // (A transposing epilogue -- 32 x 64 block through a swizzled smem tile so that a warp store covers 4 full 128-byte
          // lines instead of 32 x 16 bytes -- was measured SLOWER, 0.395 vs 0.381 ms: the 12 % that what-if variant 6 (no
          // stores) saves is the 0.31 GB of output traffic itself, not the store pattern.)
          const int row = t * 128 + q * 32 + lane;
          const float inv = 1.0f / __uint_as_float(rs);
          __nv_bfloat16* orow = a.out + ((size_t)frame * a.L + row) * a.d + (size_t)head * HD;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o0[8 * i + 0]) * inv, __uint_as_float(o0[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(orow)[i] = o;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o1[8 * i + 0]) * inv, __uint_as_float(o1[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(orow + 32)[i] = o;
          }
        }
      }
    }
  } else {
    // ===================== softmax warps =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    const bool warp_active = (t * 128 + q * 32) < a.L;  // warp-uniform
    const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;
    const int n_full = a.L >> 5;           // chunks with 32 valid columns
    const int tail = a.L - (n_full << 5);  // valid columns of the last, partial chunk (0: none)
    auto pcol = [&](int c) { return tb + uint32_t(c < A5_PA_CHUNKS ? c * 16 : c * 32); };
    constexpr int VX = VAR >= 4 ? 0 : VAR;
    for (int k = 0; k < n_my; ++k) {
      const uint32_t par = (uint32_t)k & 1u;
      mbar_wait(s_full(t), par);
      tc_fence_after();
      if (q == 0 && lane == 0) VMC_DBG5(k, 8 + 8 * t);
      if (warp_active && (VAR == 1 || VAR == 5)) {
        if (lane == 0) mbar_arrive_relaxed(pa_full(t));
      } else if (warp_active) {
        // single pass over the score row, 32-column chunks double-buffered in registers; stabiliser = max of
        // the first 32 keys (<= row max, so the row sum is >= 1; argument clamped at +120), see v3.
        // Chunks 0 and 1 (keys 0..63) come from S_a, which was computed while the previous item was being finished; the
        // rest of the row (S_b) needs the drained accumulator and is waited for only after them.
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tb, r0);
        tmem_ld_32x32b_x32(tb + 32u, r1);
        tmem_ld_wait();
        const float mxs = chunk_max<false>(r0, -INFINITY, 32) * sc;
        chunk_exp_store5<false, VX>(r0, sc, mxs, 32, pcol(0));
        chunk_exp_store5<false, VX>(r1, sc, mxs, 32, pcol(1));
        if (DBG && t == 0 && q == 0 && lane == 0) VMC_DBG5(k, 7);
        mbar_wait(sb_full(t), par);
        tc_fence_after();
        if (DBG && t == 0 && q == 0 && lane == 0) VMC_DBG5(k, 6);
        tmem_ld_32x32b_x32(tb + 64u, r0);
        for (int c = 2; c < n_chunks; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n_chunks) tmem_ld_32x32b_x32(tb + uint32_t((c + 1) * 32), r1);
          if (c < n_full) chunk_exp_store5<false, VX>(r0, sc, mxs, 32, pcol(c));
          else chunk_exp_store5<true, VX>(r0, sc, mxs, tail, pcol(c));
          if (DBG && t == 0 && q == 0 && lane == 0) VMC_DBG5(k, 25 + c);  // chunks 2, 4, 6: slots 27, 29, 31
          if (c == A5_PA_CHUNKS - 1) {
            // chunks 0..4 (keys 0..159) are stored: release part A of the PV MMA.  The load of chunk 5
            // issued above reads columns >= 160, disjoint from P_a and the accumulator.
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(pa_full(t));
            if (q == 0 && lane == 0) VMC_DBG5(k, 9 + 8 * t);
          }
          if (c + 1 < n_chunks) {
            tmem_ld_wait();
            if (c + 2 < n_chunks) tmem_ld_32x32b_x32(tb + uint32_t((c + 2) * 32), r0);
            if (c + 1 < n_full) chunk_exp_store5<false, VX>(r1, sc, mxs, 32, pcol(c + 1));
            else chunk_exp_store5<true, VX>(r1, sc, mxs, tail, pcol(c + 1));
            if (DBG && t == 0 && q == 0 && lane == 0) VMC_DBG5(k, 26 + c);  // chunks 3, 5: slots 28, 30
          }
        }
        tmem_st_wait();
      } else {
        if (lane == 0) mbar_arrive_relaxed(pa_full(t));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(pb_full(t));
      if (q == 0 && lane == 0) VMC_DBG5(k, 10 + 8 * t);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// Attention v6 (224 < L <= 257; ViT-L/14: 257 tokens = CLS + 256 patches).  Two 128-row query tiles with the whole
// score row in TMEM cover at most 256 keys (2 x 256 of the 512 TMEM columns), so v5 stops at 256 tokens and 257
// would need a third, one-row query tile.  v6 runs the v5 pipeline on the PATCH tokens only (queries and keys
// 1 .. L-1: the TMA boxes simply start at token 1) and handles the CLS token on the CUDA cores, in two extra warps
// that work out of the same Q / K / V tiles in shared memory while the tensor core is busy:
//   * warp 14, "CLS key": s_cls[r] = q_r . k_cls for the <= 256 patch queries (swizzled LDS.128 of the Q rows).  The
//     softmax thread of row r turns it into p_cls = 2^(s_cls sc - m_r) with the row's own stabiliser; the epilogue
//     adds p_cls v_cls to the accumulator and p_cls to the tensor-core row sum before normalising;
//   * warp 15, "CLS query": the whole attention row of q_cls in fp32 (scores over the K tile + k_cls, softmax,
//     P V over the V tile + v_cls) -> output row 0.
// Q / K are released by three arrivals (score MMAs retired + both CLS warps), V by two.  The all-ones tile of the
// N = 80 PV MMA (row sums from the tensor core) is ONE 2 KB block here: each 16-key step gets its own descriptor
// whose leading-dimension offset points back at it (v5 keeps one copy per step: 32 KB at 256 keys).
// ---------------------------------------------------------------------------------------
struct Attn6Args {
  int L, heads, d, lk16, n_items;  // lk16 = 16-multiple >= L - 1 (patch keys)
  __nv_bfloat16* out;
  const __nv_bfloat16* qkv;
  int var;  // what-if timing variants (WRONG results; tools/kernel_bench.py impl 61..): bit 0 warp 14 idle, 1 warp 15 no scores,
            // 2 warp 15 no P V, 3 epilogue without the CLS key, 4 softmax without the CLS key
};
constexpr int A6_THREADS = 512;  // 8 softmax + 4 epilogue + TMA + MMA + 2 CLS warps

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// A 64-element bf16 row in global memory as the B operand of four mma.m16n8k16 k-steps (every column of B = the row, so
// every column of C is the same dot product): lane l holds k = 2 (l % 4), +1 (b0) and k + 8, +9 (b1) of each 16-deep step.
__device__ __forceinline__ void load_bfrag64(const __nv_bfloat16* row, int lane, uint32_t (&b)[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    b[2 * ks] = __ldg(w + 8 * ks + (lane & 3));
    b[2 * ks + 1] = __ldg(w + 8 * ks + 4 + (lane & 3));
  }
}
// dst[r] = tile row r . vector for the first 16 * n_blocks rows of a 128B-swizzled [rows x 64] bf16 tile (contiguous 128-row
// tiles, base 1024-byte aligned), on the warp-level tensor path: per 16-row block four ldmatrix.x4 (the swizzled 16-byte
// chunk of row r is chunk ^ (r & 7): the lanes supply the row addresses) + four mma.sync.m16n8k16 with fp32 accumulation --
// 8 instructions per 16 rows instead of ~2200 CUDA-core instructions.  bf16 x bf16 products are exact in fp32.
__device__ __forceinline__ void tile_rows_dot_mma(uint32_t tile_base, int n_blocks, const uint32_t (&b)[8], float* dst, int lane) {
  const uint32_t m = (uint32_t)lane >> 3, r = (uint32_t)lane & 7u;
  const uint32_t rowoff = ((m & 1u) * 8u + r) * 128u;
  uint32_t coff[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) coff[ks] = ((uint32_t(2 * ks) + (m >> 1)) ^ r) << 4;
#pragma unroll 2
  for (int rb = 0; rb < n_blocks; ++rb) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    const uint32_t ra = tile_base + (uint32_t)rb * 2048u + rowoff;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a0, a1, a2, a3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                   : "r"(ra + coff[ks]));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b[2 * ks]), "r"(b[2 * ks + 1]));
    }
    if ((lane & 3) == 0) {  // column 0 of C: rows lane / 4 (c0) and lane / 4 + 8 (c2)
      dst[rb * 16 + (lane >> 2)] = c0;
      dst[rb * 16 + 8 + (lane >> 2)] = c2;
    }
    (void)c1;
    (void)c3;
  }
}

template <int VAR>  // 0 = production; see Attn6Args::var (what-if timing variants, wrong results)
__global__ void __launch_bounds__(A6_THREADS, 1)
attention_vit6_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm1,
                      const Attn6Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const int Lp = a.L - 1;                                  // patch tokens = MMA queries and keys
  const int r1 = a.lk16 - 128;                             // rows of the second Q / K / V box
  const uint32_t mat_bytes = (uint32_t)(128 + r1) * 128u;  // one of Q, K, V for one item
  const uint32_t qk_base = base;                           // ring of 2 x [Q | K]
  const uint32_t v_base = base + 4 * mat_bytes;            // ring of 2 x V
  const uint32_t ones_base = base + 6 * mat_bytes;         // 2 KB of bf16 1.0
  const uint32_t scls_base = ones_base + 2048u;            // fp32 [2][256]: q_r . k_cls of the patch queries
  const uint32_t pcls_base = scls_base + 2048u;            // fp32 [2][256]: p_cls of the patch queries
  const uint32_t pbuf_base = pcls_base + 2048u;            // fp32 [256]: probabilities of the CLS query row
  const uint32_t pb16_base = pbuf_base + 1024u;            // bf16 [256]: the same, rounded, as the A operand of the P V mma
  const uint32_t bar_base = pb16_base + 512u;
  const int nkk = a.lk16 / 16;  // 16-key steps of the PV MMA
  auto qk_full = [&](int s) { return bar_base + 8u * s; };
  auto qk_empty = [&](int s) { return bar_base + 16u + 8u * s; };
  auto s_full = [&](int t) { return bar_base + 32u + 8u * t; };
  auto pa_full = [&](int t) { return bar_base + 48u + 8u * t; };
  auto pb_full = [&](int t) { return bar_base + 64u + 8u * t; };
  auto o_full = [&](int t) { return bar_base + 80u + 8u * t; };
  auto s_empty = [&](int t) { return bar_base + 96u + 8u * t; };
  auto v_full = [&](int s) { return bar_base + 112u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 128u + 8u * s; };
  auto c_full = [&](int s) { return bar_base + 144u + 8u * s; };
  auto c_empty = [&](int s) { return bar_base + 160u + 8u * s; };
  const uint32_t tmem_ptr_addr = bar_base + 176u;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));
  float* scls = reinterpret_cast<float*>(smem_raw + (scls_base - raw_addr));
  float* pcls = reinterpret_cast<float*>(smem_raw + (pcls_base - raw_addr));
  float* pbuf = reinterpret_cast<float*>(smem_raw + (pbuf_base - raw_addr));
  __nv_bfloat16* pb16 = reinterpret_cast<__nv_bfloat16*>(smem_raw + (pb16_base - raw_addr));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_my = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA

  {
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (ones_base - raw_addr));
    for (uint32_t i = tid; i < 2048u / 16; i += A6_THREADS)
      ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(qk_full(i), 1);
      mbar_init(qk_empty(i), 3);  // score MMAs retired + CLS-key warp + CLS-query warp
      mbar_init(v_full(i), 1);
      mbar_init(v_empty(i), 2);   // PV MMAs retired + CLS-query warp
      mbar_init(s_full(i), 1);
      mbar_init(pa_full(i), 4);
      mbar_init(pb_full(i), 4);
      mbar_init(o_full(i), 1);
      mbar_init(s_empty(i), 4);
      mbar_init(c_full(i), 1);
      mbar_init(c_empty(i), 8);  // one arrive per softmax warp
    }
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;
  const int n_chunks = (a.lk16 + 31) / 32;  // 5..8

  if (warp == 12) {
    // ===================== TMA producer (patch tokens: rows 1 .. of the frame) =====================
    if (lane == 0) {
      int kq = 0, kv = 0;
      while (kq < n_my || kv < n_my) {
        if (kq < n_my && mbar_test_wait(qk_empty(kq & 1), (((uint32_t)kq >> 1) & 1u) ^ 1u)) {
          const int item = blockIdx.x + kq * gridDim.x;
          const int head = item % a.heads, frame = item / a.heads;
          const int s = kq & 1;
          const uint32_t q = qk_base + s * 2 * mat_bytes, kk_ = q + mat_bytes;
          mbar_arrive_expect_tx(qk_full(s), 2 * mat_bytes);
          tma_load_3d(q, &tm, qk_full(s), head * HD, 1, frame);
          tma_load_3d(q + TILE, &tm1, qk_full(s), head * HD, 129, frame);
          tma_load_3d(kk_, &tm, qk_full(s), a.d + head * HD, 1, frame);
          tma_load_3d(kk_ + TILE, &tm1, qk_full(s), a.d + head * HD, 129, frame);
          ++kq;
        }
        if (kv < n_my && mbar_test_wait(v_empty(kv & 1), (((uint32_t)kv >> 1) & 1u) ^ 1u)) {
          const int item = blockIdx.x + kv * gridDim.x;
          const int head = item % a.heads, frame = item / a.heads;
          const int s = kv & 1;
          const uint32_t v = v_base + s * mat_bytes;
          mbar_arrive_expect_tx(v_full(s), mat_bytes);
          tma_load_3d(v, &tm, v_full(s), 2 * a.d + head * HD, 1, frame);
          tma_load_3d(v + TILE, &tm1, v_full(s), 2 * a.d + head * HD, 129, frame);
          ++kv;
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ===================== MMA issuer (event driven: one thread polls both tiles' barriers) =====================
    // (v5 / v7 moved to one BLOCKING issuer warp per tile in round 2: 0.381 -> 0.30 ms and 0.087 -> 0.056 ms.  The same change
    // here needs a 17th warp, which caps the kernel at 120 registers: measured 0.387 vs 0.354 ms, so v6 keeps the poller.)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, a.lk16, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 80, 0, 1);  // B = [V | ones], MN-major
      const int nka = nkk < 2 * A5_PA_CHUNKS ? nkk : 2 * A5_PA_CHUNKS;
      int kt[2] = {0, 0};
      int ph[2] = {0, 0};
      int sdone[2] = {0, 0};
      int fin[2] = {0, 0};
      while (kt[0] < n_my || kt[1] < n_my) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (kt[t] >= n_my) continue;
          const int k = kt[t];
          const int s = k & 1;
          const uint32_t par = (uint32_t)k & 1u;
          const uint32_t ring_par = ((uint32_t)k >> 1) & 1u;
          const uint32_t vst = v_base + s * mat_bytes;
          // V descriptor of 16-key step kk: atom 0 = the TMA tile rows [16 kk, 16 kk + 16) (64 head dims), atom 1
          // (N = 64..79) = LBO further = the ones block
          auto dv = [&](int kk) {
            const uint32_t va = vst + (uint32_t)kk * 2048u;
            return (umma_desc_sw128(va) & ~(uint64_t(0x3FFF) << 16)) | (uint64_t((ones_base - va) >> 4) << 16);
          };
          const uint32_t tcol = tmem_base + uint32_t(t * 256);
          if (ph[t] == 0) {
            if (t == 1 && k == 0 && kt[0] == 0 && ph[0] < 2) continue;
            if (!mbar_test_wait(qk_full(s), ring_par)) continue;
            if (!mbar_test_wait(s_empty(t), par ^ 1u)) continue;
            tc_fence_after();
            const uint32_t qst = qk_base + s * 2 * mat_bytes;
            const uint64_t dq = umma_desc_sw128(qst + t * TILE);
            const uint64_t dk = umma_desc_sw128(qst + mat_bytes);
#pragma unroll
            for (int kq = 0; kq < HD / 16; ++kq)
              umma_ss(tcol, dq + uint64_t(2 * kq), dk + uint64_t(2 * kq), idesc_s, kq != 0);
            umma_commit(s_full(t));
            if (++sdone[s] == 2) {
              sdone[s] = 0;
              umma_commit(qk_empty(s));
            }
            ph[t] = 1;
          } else if (ph[t] == 1) {
            if (!mbar_test_wait(pa_full(t), par)) continue;
            if (!mbar_test_wait(v_full(s), ring_par)) continue;
            tc_fence_after();
            for (int kk = 0; kk < nka; ++kk)
              umma_ts(tcol + A5_OCOL, tcol + uint32_t(kk * 8), dv(kk), idesc_pv, kk != 0);
            ph[t] = 2;
          } else {
            if (!mbar_test_wait(pb_full(t), par)) continue;
            tc_fence_after();
            for (int kk = nka; kk < nkk; ++kk)
              umma_ts(tcol + A5_OCOL, tcol + uint32_t((kk >> 1) * 32 + (kk & 1) * 8), dv(kk), idesc_pv, 1u);
            umma_commit(o_full(t));
            if (++fin[s] == 2) {
              fin[s] = 0;
              umma_commit(v_empty(s));
            }
            kt[t] = k + 1;
            ph[t] = 0;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ===================== CLS key: s_cls[r] = q_r . k_cls for the patch queries =====================
    for (int k = 0; k < n_my; ++k) {
      const int item = blockIdx.x + k * gridDim.x;
      const int head = item % a.heads, frame = item / a.heads;
      const int s = k & 1;
      uint32_t bk[8];
      load_bfrag64(a.qkv + (size_t)frame * a.L * 3 * a.d + a.d + head * HD, lane, bk);  // k of token 0 as the B operand
      if (lane == 0 && k + 1 < n_my) {  // next item's row into L1: the load above is otherwise an exposed L2 round trip per item
        const int nit = item + gridDim.x;
        prefetch_l1(a.qkv + (size_t)(nit / a.heads) * a.L * 3 * a.d + a.d + (nit % a.heads) * HD);
      }
      mbar_wait(qk_full(s), ((uint32_t)k >> 1) & 1u);
      mbar_wait(c_empty(s), (((uint32_t)k >> 1) & 1u) ^ 1u);  // slot s of s_cls: read by all softmax warps of item k - 2
      if (!(VAR & 1)) tile_rows_dot_mma(qk_base + s * 2 * mat_bytes, a.lk16 / 16, bk, scls + s * 256, lane);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(c_full(s));
        mbar_arrive(qk_empty(s));
      }
    }
  } else if (warp == 15) {
    // ===================== CLS query: the attention row of token 0 =====================
    // (A software-pipelined order -- scores(k), release Q / K, P V(k - 1), softmax(k) -- was measured SLOWER, 0.51 vs 0.41 ms:
    // it delays the V release instead.  What the what-if variants showed is that the ~1100 CUDA-core instructions of the score
    // phase cost 0.1 ms; they now run on the warp-level tensor path, see tile_rows_dot_mma.)
    for (int k = 0; k < n_my; ++k) {
      const int item = blockIdx.x + k * gridDim.x;
      const int head = item % a.heads, frame = item / a.heads;
      const int s = k & 1;
      const uint32_t ring_par = ((uint32_t)k >> 1) & 1u;
      const __nv_bfloat16* row0 = a.qkv + (size_t)frame * a.L * 3 * a.d + head * HD;
      uint32_t bq[8];
      load_bfrag64(row0, lane, bq);  // q of token 0 as the B operand
      const uint32_t kc2 = __ldg(reinterpret_cast<const uint32_t*>(row0 + a.d) + lane);
      uint32_t vcw[8];  // v_cls dims 8 nb + 2 (lane & 3), +1: the accumulator layout of lanes 0..3
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) vcw[nb] = __ldg(reinterpret_cast<const uint32_t*>(row0 + 2 * a.d) + 4 * nb + (lane & 3));
      if (lane < 3 && k + 1 < n_my) {  // next item's q / k / v rows of token 0 into L1
        const int nit = item + gridDim.x;
        prefetch_l1(a.qkv + (size_t)(nit / a.heads) * a.L * 3 * a.d + lane * a.d + (nit % a.heads) * HD);
      }
      // q_cls . k_cls: lane l holds dims 2l, 2l+1 of both rows
      const uint32_t qc2 = __ldg(reinterpret_cast<const uint32_t*>(row0) + lane);
      const float sc0 = warp_sum(fmaf(bf_lo(qc2), bf_lo(kc2), bf_hi(qc2) * bf_hi(kc2))) * 0.125f;
      mbar_wait(qk_full(s), ring_par);
      if (!(VAR & 2)) tile_rows_dot_mma(qk_base + s * 2 * mat_bytes + mat_bytes, a.lk16 / 16, bq, pbuf, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(qk_empty(s));
      float sj[8];
      float mx = sc0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = lane + 32 * i;
        sj[i] = (j < Lp && !(VAR & 2)) ? pbuf[j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, sj[i]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // exp(-inf) = 0 for the padding keys; rounded to bf16 like the probabilities of every other row, and the
        // normaliser sums the ROUNDED values
        const __nv_bfloat16 pb = __float2bfloat16_rn(__expf(sj[i] - mx));
        pb16[lane + 32 * i] = pb;
        sum += __bfloat162float(pb);
      }
      const float pc = __expf(sc0 - mx);
      sum = warp_sum(sum) + pc;
      __syncwarp();
      mbar_wait(v_full(s), ring_par);
      // O[64] = sum_j p_j V[j][:] on the warp-level tensor path: A = P (row 0 of the 16 x 16 fragment = 16 probabilities, the
      // other rows zero: only lanes 0..3 hold non-zero A registers), B = V[16 keys x 8 dims] through ldmatrix.trans from
      // the token-major swizzled tile (one x4 = two 8-dim blocks), fp32 accumulators: lane t < 4 ends up with dims
      // 8 nb + 2 t, +1 of every 8-dim block nb.
      const uint32_t vst = v_base + s * mat_bytes;
      float acc[8][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
      {
        const uint32_t m = (uint32_t)lane >> 3, r = (uint32_t)lane & 7u;
        const uint32_t rowoff = ((m & 1u) * 8u + r) * 128u;
        const uint32_t pa = pb16_base + ((uint32_t)lane & 3u) * 4u;
        for (int ks = 0; ks < ((VAR & 4) ? 0 : a.lk16 / 16); ++ks) {
          uint32_t a0 = 0u, a2 = 0u;
          if (lane < 4) {
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(pa + (uint32_t)ks * 32u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a2) : "r"(pa + (uint32_t)ks * 32u + 16u));
          }
          const uint32_t ra = vst + (uint32_t)ks * 2048u + rowoff;
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                         : "r"(ra + (((uint32_t(2 * np) + (m >> 1)) ^ r) << 4)));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[2 * np][0]), "+f"(acc[2 * np][1]), "+f"(acc[2 * np][2]), "+f"(acc[2 * np][3])
                         : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[2 * np + 1][0]), "+f"(acc[2 * np + 1][1]), "+f"(acc[2 * np + 1][2]), "+f"(acc[2 * np + 1][3])
                         : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b2), "r"(b3));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty(s));
      if (lane < 4) {
        const float inv = 1.0f / sum;
        uint32_t* orow = reinterpret_cast<uint32_t*>(a.out + (size_t)frame * a.L * a.d + (size_t)head * HD);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
          orow[4 * nb + lane] = pack_bf16x2(fmaf(pc, bf_lo(vcw[nb]), acc[nb][0]) * inv, fmaf(pc, bf_hi(vcw[nb]), acc[nb][1]) * inv);
      }
      __syncwarp();  // pbuf is rewritten by the next item's scores
    }
  } else if (warp >= 8) {
    // ===================== epilogue warps: drain O, add the CLS key, write the rows =====================
    const int q = warp - 8;
    for (int k = 0; k < n_my; ++k) {
      const int item = blockIdx.x + k * gridDim.x;
      const int head = item % a.heads;
      const int frame = item / a.heads;
      const uint32_t par = (uint32_t)k & 1u;
      // v of token 0: the same 128 bytes for every row of the item (L2 hit for the first reader, L1 after that)
      const uint4* vc4 = reinterpret_cast<const uint4*>(a.qkv + (size_t)frame * a.L * 3 * a.d + 2 * a.d + head * HD);
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int row = t * 128 + q * 32 + lane;           // patch index; token row + 1
        const bool warp_active = (t * 128 + q * 32) < Lp;  // warp-uniform
        const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
        mbar_wait(o_full(t), par);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        uint32_t rs = 0x3F800000u;
        if (warp_active) {
          tmem_ld_32x32b_x32(tb + A5_OCOL, o0);
          tmem_ld_32x32b_x32(tb + A5_OCOL + 32u, o1);
          rs = tmem_ld_32x32b_x1(tb + A5_OCOL + 64u);
          tmem_ld_wait();
        }
        const float pc = pcls[(k & 1) * 256 + row];
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(s_empty(t));
        if (warp_active && row < Lp && (VAR & 8)) {
          const float inv = 1.0f / __uint_as_float(rs);
          __nv_bfloat16* orow = a.out + ((size_t)frame * a.L + row + 1) * a.d + (size_t)head * HD;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<uint4*>(orow)[i] = make_uint4(pack_bf16x2(__uint_as_float(o0[4 * i]) * inv, __uint_as_float(o0[4 * i + 1]) * inv), pack_bf16x2(__uint_as_float(o0[4 * i + 2]) * inv, __uint_as_float(o0[4 * i + 3]) * inv), pack_bf16x2(__uint_as_float(o1[4 * i]) * inv, __uint_as_float(o1[4 * i + 1]) * inv), pack_bf16x2(__uint_as_float(o1[4 * i + 2]) * inv, __uint_as_float(o1[4 * i + 3]) * inv));
        } else if (warp_active && row < Lp) {
          const float inv = 1.0f / (__uint_as_float(rs) + pc);
          __nv_bfloat16* orow = a.out + ((size_t)frame * a.L + row + 1) * a.d + (size_t)head * HD;
          uint32_t vc[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = __ldg(vc4 + c);
            vc[4 * c] = v.x; vc[4 * c + 1] = v.y; vc[4 * c + 2] = v.z; vc[4 * c + 3] = v.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(fmaf(pc, bf_lo(vc[4 * i + 0]), __uint_as_float(o0[8 * i + 0])) * inv,
                              fmaf(pc, bf_hi(vc[4 * i + 0]), __uint_as_float(o0[8 * i + 1])) * inv);
            o.y = pack_bf16x2(fmaf(pc, bf_lo(vc[4 * i + 1]), __uint_as_float(o0[8 * i + 2])) * inv,
                              fmaf(pc, bf_hi(vc[4 * i + 1]), __uint_as_float(o0[8 * i + 3])) * inv);
            o.z = pack_bf16x2(fmaf(pc, bf_lo(vc[4 * i + 2]), __uint_as_float(o0[8 * i + 4])) * inv,
                              fmaf(pc, bf_hi(vc[4 * i + 2]), __uint_as_float(o0[8 * i + 5])) * inv);
            o.w = pack_bf16x2(fmaf(pc, bf_lo(vc[4 * i + 3]), __uint_as_float(o0[8 * i + 6])) * inv,
                              fmaf(pc, bf_hi(vc[4 * i + 3]), __uint_as_float(o0[8 * i + 7])) * inv);
            reinterpret_cast<uint4*>(orow)[i] = o;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(fmaf(pc, bf_lo(vc[16 + 4 * i + 0]), __uint_as_float(o1[8 * i + 0])) * inv,
                              fmaf(pc, bf_hi(vc[16 + 4 * i + 0]), __uint_as_float(o1[8 * i + 1])) * inv);
            o.y = pack_bf16x2(fmaf(pc, bf_lo(vc[16 + 4 * i + 1]), __uint_as_float(o1[8 * i + 2])) * inv,
                              fmaf(pc, bf_hi(vc[16 + 4 * i + 1]), __uint_as_float(o1[8 * i + 3])) * inv);
            o.z = pack_bf16x2(fmaf(pc, bf_lo(vc[16 + 4 * i + 2]), __uint_as_float(o1[8 * i + 4])) * inv,
                              fmaf(pc, bf_hi(vc[16 + 4 * i + 2]), __uint_as_float(o1[8 * i + 5])) * inv);
            o.w = pack_bf16x2(fmaf(pc, bf_lo(vc[16 + 4 * i + 3]), __uint_as_float(o1[8 * i + 6])) * inv,
                              fmaf(pc, bf_hi(vc[16 + 4 * i + 3]), __uint_as_float(o1[8 * i + 7])) * inv);
            reinterpret_cast<uint4*>(orow + 32)[i] = o;
          }
        }
      }
    }
  } else {
    // ===================== softmax warps (as v5, plus the CLS key's probability) =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    const bool warp_active = (t * 128 + q * 32) < Lp;  // warp-uniform
    const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;
    const int n_full = Lp >> 5;
    const int tail = Lp - (n_full << 5);
    const int row = t * 128 + q * 32 + lane;
    auto pcol = [&](int c) { return tb + uint32_t(c < A5_PA_CHUNKS ? c * 16 : c * 32); };
    for (int k = 0; k < n_my; ++k) {
      const uint32_t par = (uint32_t)k & 1u;
      mbar_wait(s_full(t), par);
      tc_fence_after();
      float mxs = 0.f;
      if (warp_active) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tb, r0);
        tmem_ld_wait();
        mxs = chunk_max<false>(r0, -INFINITY, 32) * sc;
        for (int c = 0; c < n_chunks; c += 2) {
          if (c != 0) tmem_ld_wait();
          if (c + 1 < n_chunks) tmem_ld_32x32b_x32(tb + uint32_t((c + 1) * 32), r1);
          if (c < n_full) chunk_exp_store5<false, 0>(r0, sc, mxs, 32, pcol(c));
          else chunk_exp_store5<true, 0>(r0, sc, mxs, tail, pcol(c));
          if (c == A5_PA_CHUNKS - 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(pa_full(t));
          }
          if (c + 1 < n_chunks) {
            tmem_ld_wait();
            if (c + 2 < n_chunks) tmem_ld_32x32b_x32(tb + uint32_t((c + 2) * 32), r0);
            if (c + 1 < n_full) chunk_exp_store5<false, 0>(r1, sc, mxs, 32, pcol(c + 1));
            else chunk_exp_store5<true, 0>(r1, sc, mxs, tail, pcol(c + 1));
          }
        }
        tmem_st_wait();
      } else {
        if (lane == 0) mbar_arrive_relaxed(pa_full(t));
      }
      // the CLS key, scored by warp 14: same stabiliser as the row's other keys; consumed by the epilogue warps
      if (!(VAR & 16)) mbar_wait(c_full(k & 1), ((uint32_t)k >> 1) & 1u);
      if (!(VAR & 16)) pcls[(k & 1) * 256 + row] = ex2_approx(fminf(fmaf(scls[(k & 1) * 256 + row], sc, -mxs), 120.f));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c_empty(k & 1));
      if (lane == 0) mbar_arrive(pb_full(t));  // release: publishes p_cls (via the MMA commit) to the epilogue warps
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// Attention v7 (L <= 64; ViT-B/32: 50 tokens): the v5 pipeline with TWO (frame, head) items packed into each 128-row
// query tile.  v2 gives a 50-token item a whole CTA and a whole 128-row tile (39 % of the rows and of the softmax threads
// busy, one serial load -> MMA -> softmax -> MMA -> store chain per CTA).  Here rows 0..63 of a tile are item A and rows
// 64..127 item B; the K and V tiles are stacked the same way, S = [Q_A; Q_B] [K_A; K_B]^T is one 128 x 128 UMMA whose
// off-diagonal blocks are discarded: a softmax thread exponentiates the 64 columns of its own item and writes ZERO
// probabilities for the other item's keys, so O = P [V_A; V_B] (N = 80: 64 dims + the all-ones row-sum columns) is
// block diagonal too.  One CTA iteration = 4 items (two tiles in flight); everything else -- Q/K and V rings released
// separately, event-driven MMA issuer, epilogue warps, single-pass softmax with the first-chunk stabiliser -- is v5.
// TMEM plan of a tile (base 256 t): S [0,128); P chunk c at [16 c, 16 c + 16); O [80,144) + row sum [144,160), written
// only after every S column has been consumed (one PV phase).
// ---------------------------------------------------------------------------------------
struct Attn7Args {
  int L, heads, d, n_items;
  int pf;  // L2 prefetch distance in CTA iterations (0 = off)
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(A5B_THREADS, 1)
attention_vit7_kernel(const __grid_constant__ CUtensorMap tm, const Attn7Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t qk_base = base;                // ring of 2 x [Q t0 | Q t1 | K t0 | K t1], 16 KB tiles
  const uint32_t v_base = base + 8 * TILE;      // ring of 2 x [V t0 | V t1]
  const uint32_t ones_base = base + 12 * TILE;  // 2 KB of bf16 1.0
  const uint32_t bar_base = ones_base + 2048u;
  auto qk_full = [&](int s) { return bar_base + 8u * s; };
  auto qk_empty = [&](int s) { return bar_base + 16u + 8u * s; };
  auto s_full = [&](int t) { return bar_base + 32u + 8u * t; };
  auto p_full = [&](int t) { return bar_base + 48u + 8u * t; };
  auto o_full = [&](int t) { return bar_base + 64u + 8u * t; };
  auto s_empty = [&](int t) { return bar_base + 80u + 8u * t; };
  auto v_full = [&](int s) { return bar_base + 96u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 112u + 8u * s; };
  const uint32_t tmem_ptr_addr = bar_base + 128u;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw_addr));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_groups = (a.n_items + 3) / 4;  // 4 items per CTA iteration
  const int n_my = (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  {
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (ones_base - raw_addr));
    for (uint32_t i = tid; i < 2048u / 16; i += A5B_THREADS)
      ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    // rows of a half tile that no TMA box ever writes (a group's missing items) must not hold NaN patterns: K / V
    // garbage would reach valid rows through 0 * NaN.  Zero the rings once.
    uint4* ring = reinterpret_cast<uint4*>(smem_raw + (base - raw_addr));
    for (uint32_t i = tid; i < 12u * TILE / 16; i += A5B_THREADS) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    tma_prefetch_desc(&tm);
    for (int i = 0; i < 2; ++i) {
      mbar_init(qk_full(i), 1);
      mbar_init(qk_empty(i), 2);  // one tcgen05.commit per tile's MMA issuer
      mbar_init(v_full(i), 1);
      mbar_init(v_empty(i), 2);
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 4);
      mbar_init(o_full(i), 1);
      mbar_init(s_empty(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 12) {
    // ===================== TMA producer: 64-row boxes (rows >= L zero-filled), item j of the group -> half tile j =====
    if (lane == 0) {
      int kq = 0, kv = 0;
      while (kq < n_my || kv < n_my) {
        if (kq < n_my && mbar_test_wait(qk_empty(kq & 1), (((uint32_t)kq >> 1) & 1u) ^ 1u)) {
          const int g = blockIdx.x + kq * gridDim.x;
          const int s = kq & 1;
          const uint32_t q = qk_base + s * 4 * TILE, kk_ = q + 2 * TILE;
          const int n_it = a.n_items - 4 * g < 4 ? a.n_items - 4 * g : 4;
          mbar_arrive_expect_tx(qk_full(s), (uint32_t)n_it * TILE);
          for (int j = 0; j < n_it; ++j) {
            const int item = 4 * g + j;
            const int head = item % a.heads, frame = item / a.heads;
            tma_load_3d(q + j * (TILE / 2), &tm, qk_full(s), head * HD, 0, frame);
            tma_load_3d(kk_ + j * (TILE / 2), &tm, qk_full(s), a.d + head * HD, 0, frame);
          }
          if (a.pf > 0 && kq + a.pf < n_my) {
            // experiment (VMC_OPT_ATTN_PREFETCH): pull the boxes of a later iteration into L2 now.  Measured slightly
            // SLOWER at every distance: the kernel is bound by the per-tile S -> softmax -> PV -> drain chain (ncu: every warp
            // parked on an mbarrier, tensor pipe 17 %, DRAM 41 %), not by the load latency.
            const int gp = blockIdx.x + (kq + a.pf) * gridDim.x;
            const int n_p = a.n_items - 4 * gp < 4 ? a.n_items - 4 * gp : 4;
            for (int j = 0; j < n_p; ++j) {
              const int item = 4 * gp + j;
              const int head = item % a.heads, frame = item / a.heads;
              tma_prefetch_3d(&tm, head * HD, 0, frame);
              tma_prefetch_3d(&tm, a.d + head * HD, 0, frame);
              tma_prefetch_3d(&tm, 2 * a.d + head * HD, 0, frame);
            }
          }
          ++kq;
        }
        if (kv < n_my && mbar_test_wait(v_empty(kv & 1), (((uint32_t)kv >> 1) & 1u) ^ 1u)) {
          const int g = blockIdx.x + kv * gridDim.x;
          const int s = kv & 1;
          const uint32_t v = v_base + s * 2 * TILE;
          const int n_it = a.n_items - 4 * g < 4 ? a.n_items - 4 * g : 4;
          mbar_arrive_expect_tx(v_full(s), (uint32_t)n_it * (TILE / 2));
          for (int j = 0; j < n_it; ++j) {
            const int item = 4 * g + j;
            const int head = item % a.heads, frame = item / a.heads;
            tma_load_3d(v + j * (TILE / 2), &tm, v_full(s), 2 * a.d + head * HD, 0, frame);
          }
          ++kv;
        }
      }
    }
    __syncwarp();
  } else if (warp == 13 || warp == 14) {
    // ===================== MMA issuers: one blocking issuer per query tile (see attention_vit5_kernel) =====================
    const int t = warp - 13;
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 80, 0, 1);  // B = [V | ones], MN-major
      const uint32_t tcol = tmem_base + uint32_t(t * 256);
      for (int k = 0; k < n_my; ++k) {
        const int s = k & 1;
        const uint32_t par = (uint32_t)k & 1u;
        const uint32_t ring_par = ((uint32_t)k >> 1) & 1u;
        mbar_wait(qk_full(s), ring_par);
        mbar_wait(s_empty(t), par ^ 1u);
        tc_fence_after();
        const uint32_t qst = qk_base + s * 4 * TILE;
        const uint64_t dq = umma_desc_sw128(qst + t * TILE);
        const uint64_t dk = umma_desc_sw128(qst + 2 * TILE + t * TILE);
#pragma unroll
        for (int kq = 0; kq < HD / 16; ++kq)
          umma_ss(tcol, dq + uint64_t(2 * kq), dk + uint64_t(2 * kq), idesc_s, kq != 0);
        umma_commit(s_full(t));
        umma_commit(qk_empty(s));  // two arrivals (one per tile) free Q and K
        mbar_wait(p_full(t), par);
        mbar_wait(v_full(s), ring_par);
        tc_fence_after();
        const uint32_t vst = v_base + s * 2 * TILE + t * TILE;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t va = vst + (uint32_t)kk * 2048u;
          const uint64_t dv = (umma_desc_sw128(va) & ~(uint64_t(0x3FFF) << 16)) | (uint64_t((ones_base - va) >> 4) << 16);
          umma_ts(tcol + A5_OCOL, tcol + uint32_t(kk * 8), dv, idesc_pv, kk != 0);
        }
        umma_commit(o_full(t));
        umma_commit(v_empty(s));
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ===================== epilogue warps =====================
    const int q = warp - 8;
    for (int k = 0; k < n_my; ++k) {
      const int g = blockIdx.x + k * gridDim.x;
      const uint32_t par = (uint32_t)k & 1u;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int item = 4 * g + 2 * t + (q >> 1);  // warp-uniform: lane quarters 0, 1 = item A, 2, 3 = item B of the tile
        const int tok = (q & 1) * 32 + lane;
        const int head = item % a.heads, frame = item / a.heads;
        const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
        mbar_wait(o_full(t), par);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(tb + A5_OCOL, o0);
        tmem_ld_32x32b_x32(tb + A5_OCOL + 32u, o1);
        const uint32_t rs = tmem_ld_32x32b_x1(tb + A5_OCOL + 64u);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(s_empty(t));
        if (item < a.n_items && tok < a.L) {
          const float inv = 1.0f / __uint_as_float(rs);
          __nv_bfloat16* orow = a.out + ((size_t)frame * a.L + tok) * a.d + (size_t)head * HD;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o0[8 * i + 0]) * inv, __uint_as_float(o0[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(orow)[i] = o;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o1[8 * i + 0]) * inv, __uint_as_float(o1[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(orow + 32)[i] = o;
          }
        }
      }
    }
  } else {
    // ===================== softmax warps: own item's 64 key columns, zeros for the other item's =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    const int blk = q >> 1;  // 0: item A (rows 0..63, keys 0..63), 1: item B
    const uint32_t tb = tmem_base + uint32_t(t * 256) + (uint32_t(q * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;
    const int v0 = a.L < 32 ? a.L : 32;  // valid keys of the item's first / second 32-key chunk
    const int v1 = a.L - 32 > 0 ? a.L - 32 : 0;
    for (int k = 0; k < n_my; ++k) {
      const uint32_t par = (uint32_t)k & 1u;
      mbar_wait(s_full(t), par);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tb + uint32_t(blk * 64), r0);
      tmem_ld_32x32b_x32(tb + uint32_t(blk * 64 + 32), r1);
      tmem_ld_wait();
      // (P overlays S columns [0, 64), but a thread only ever touches its own TMEM lane: its scores are in registers now)
      const float mxs = chunk_max<true>(r0, -INFINITY, v0) * sc;
      if (v0 == 32) chunk_exp_store5<false, 0>(r0, sc, mxs, 32, tb + uint32_t((2 * blk) * 16));
      else chunk_exp_store5<true, 0>(r0, sc, mxs, v0, tb + uint32_t((2 * blk) * 16));
      chunk_exp_store5<true, 0>(r1, sc, mxs, v1, tb + uint32_t((2 * blk + 1) * 16));
      {
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        tmem_st_32x32b_x16(tb + uint32_t((2 * (blk ^ 1)) * 16), z);
        tmem_st_32x32b_x16(tb + uint32_t((2 * (blk ^ 1) + 1) * 16), z);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(p_full(t));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
                        int out_f32, long long ldo, int Tq, int Tk, const float* __restrict__ pmask) {
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
      float pd0 = p0, pd1 = p1;
      if (pmask != nullptr && q0 + r < Tq) {
        // training: dropout on the attention probabilities (nn.MultiheadAttention(dropout=p)); the mask holds
        // 0 or 1/(1-p) and touches only the PV product, the normaliser sums the un-dropped probabilities
        const float* pmr = pmask + (((size_t)b * gridDim.y + h) * Tq + q0 + r) * Tk;
        if (j0 < Tk) pd0 *= pmr[j0];
        if (j1 < Tk) pd1 *= pmr[j1];
      }
      float a0 = o0[i] * corr, a1 = o1[i] * corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj0 = __shfl_sync(0xffffffffu, pd0, j);
        const float pj1 = __shfl_sync(0xffffffffu, pd1, j);
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

// ---------------------------------------------------------------------------------------
// CLS-query attention for the LAST transformer block (opt-in, VMC_OPT_LAST_BLOCK_CLS): the tower's output is
// ln_post(x[:, 0]) @ proj, so in the last block only the CLS row of the attention / MLP output is ever read.  One
// CTA per (frame, head): q = the CLS query [64], K / V = all L tokens (bf16, packed [F*L, 2d] = [k | v]).
// Memory-bound (K and V are read once); scores by one warp per key (coalesced 128-byte rows, shuffle reduction).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_cls_kernel(const __nv_bfloat16* __restrict__ qcls, const __nv_bfloat16* __restrict__ kv,
                     __nv_bfloat16* __restrict__ out, int L, int heads) {
  // One CTA per (frame, head); the CLS query against all L keys.  [r2] 16-byte loads: 8 lanes cover one 128-byte head row,
  // so a warp instruction fetches FOUR key (or value) rows and a score costs 3 shuffles instead of 5 -- the first version read
  // one row per warp instruction (4 bytes per lane) and ran at a third of the HBM rate (ncu launch list: 2.07 TB/s).
  extern __shared__ float s_sc[];  // [L] scores -> probabilities
  __shared__ float s_red[4];
  __shared__ float s_o[16][HD];
  const int f = blockIdx.y, h = blockIdx.x;
  const int d = heads * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = tid & 7;   // 16-byte chunk (8 head dims) of a row
  const int grp = tid >> 3;  // row within a 16-row step
  float q[8];
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(qcls + (size_t)f * d + h * HD + sub * 8);
    const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[2 * i] = __uint_as_float(w[i] << 16);
      q[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  const __nv_bfloat16* kbase = kv + (size_t)f * L * 2 * d + h * HD + sub * 8;
  const __nv_bfloat16* vbase = kbase + d;
#pragma unroll 4
  for (int j0 = 0; j0 < L; j0 += 16) {
    const int j = j0 + grp;
    float s = 0.f;
    if (j < L) {
      const uint4 kv4 = __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)j * 2 * d));
      const uint32_t w[4] = {kv4.x, kv4.y, kv4.z, kv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) s = fmaf(q[2 * i], __uint_as_float(w[i] << 16), fmaf(q[2 * i + 1], __uint_as_float(w[i] & 0xFFFF0000u), s));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0 && j < L) s_sc[j] = s * 0.125f;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < L; j += 128) mx = fmaxf(mx, s_sc[j]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < L; j += 128) {
    const float p = __expf(s_sc[j] - mx);
    s_sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3]));
  // O[dim] = sum_j p_j v_j[dim]: 16 row groups x 8 chunks of 8 dims
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int j = grp; j < L; j += 16) {
    const float p = s_sc[j];
    const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)j * 2 * d));
    const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(p, __uint_as_float(w[i] << 16), acc[2 * i]);
      acc[2 * i + 1] = fmaf(p, __uint_as_float(w[i] & 0xFFFF0000u), acc[2 * i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_o[grp][sub * 8 + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) o += s_o[g][tid];
    out[(size_t)f * d + h * HD + tid] = __float2bfloat16_rn(o * inv);
  }
}

uint32_t pow2_at_least(uint32_t v) {
  uint32_t p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

int vmc_get_option(int option);
long long vmc_get_option64(int option);
extern "C" int vmc_attention_vit_short_mma(const void* qkv, void* out, int F, int L, int heads, void* stream);

extern "C" {

int vmc_attention_vit_impl(const void* qkv, void* out, int F, int L, int heads, int impl,
                           void* stream) {
  VMC_CHECK_ARG(qkv && out, VMC_ERR_ARG, "vmc_attention_vit: null pointer");
  VMC_CHECK_ARG(F > 0 && heads > 0 && L > 0 && L <= 272, VMC_ERR_SHAPE,
                "vmc_attention_vit: need 0 < L <= 272 tokens (L=%d)", L);
  int var6 = 0;
#ifdef VMC_WHATIF  // timing experiments that produce WRONG results: never part of the shipped library (make WHATIF=1)
  if (impl >= 100 && impl < 132) {  // 100 + bits: what-if timing variants of v6
    var6 = impl - 100;
    impl = 6;
  }
  VMC_CHECK_ARG(impl == 2 || (impl >= 5 && impl <= 8) || (impl >= 51 && impl <= 58), VMC_ERR_ARG,
                "vmc_attention_vit: impl must be 2, 5, 6, 7 or 8");
#else
  VMC_CHECK_ARG(impl == 2 || (impl >= 5 && impl <= 8), VMC_ERR_ARG, "vmc_attention_vit: impl must be 2, 5, 6, 7 or 8");
#endif
  const int d = heads * HD;
  if (impl == 8) {  // warp-level tensor path for short sequences (backward.cu)
    // (The same flash-attention-2 style kernel for 64 < L <= 272 -- scores / P / O in registers, keys in chunks of 64, two
    // CTAs per SM -- was built and measured in round 2: 0.730 ms per 1024 ViT-B/16 frames against 0.381 ms for v5.  Each of
    // the 13 sixteen-row tiles re-reads the item's K and V fragments from shared memory: 0.7 MB per item, ~5.4 k cycles at
    // 128 B/clk, more than v5's whole item.  Removed; profiles/r02_kernel_bench_attention.txt.)
    if (L <= 64) return vmc_attention_vit_short_mma(qkv, out, F, L, heads, stream);
    impl = 5;
  }
  // v7 = two items packed per query tile in the v5 pipeline: default for short sequences (ViT-B/32: 50 tokens)
  if (impl == 5 && L <= 64) impl = 7;
  if (impl == 7 && L > 64) impl = 5;
  if (impl == 7) {
    Attn7Args a7;
    a7.L = L;
    a7.heads = heads;
    a7.d = d;
    a7.n_items = F * heads;
    {
      const int pf = vmc_get_option(VMC_OPT_ATTN_PREFETCH);
      a7.pf = pf > 0 ? pf : 0;  // measured: 0.086 ms (off) vs 0.090-0.093 ms (distance 1..4) per 1024 ViT-B/32 frames -- off by default
    }
    a7.out = reinterpret_cast<__nv_bfloat16*>(out);
    CUtensorMap tm7;
    const uint64_t dims7[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)F};
    const uint64_t strides7[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
    const uint32_t box7[3] = {HD, 64, 1};
    VMC_TRY(vmc_encode_tmap_bf16(&tm7, qkv, 3, dims7, strides7, box7));
    const uint32_t smem7 = 12u * TILE + 2048u + 256u + 1024u;
    cudaStream_t st7 = reinterpret_cast<cudaStream_t>(stream);
    const int groups7 = (a7.n_items + 3) / 4;
    const int grid7 = groups7 < vmc_num_sms() ? groups7 : vmc_num_sms();
    VMC_CUDA(cudaFuncSetAttribute(attention_vit7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem7));
    {
      VmcProfScope prof(VMC_K_ATTN_VIT, st7, 4.0 * F * heads * (double)L * L * HD, 8.0 * F * L * d);
      attention_vit7_kernel<<<grid7, A5B_THREADS, smem7, st7>>>(tm7, a7);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  // v6 = the v5 pipeline on the patch tokens + the CLS token on the CUDA cores: default above v5's 224 tokens (ViT-L/14: 257)
  if (impl == 5 && L > 224 && L <= 257) impl = 6;
  if (impl == 6 && (L < 145 || L > 257)) impl = 5;
  if (impl == 6) {
    Attn6Args a6;
    a6.L = L;
    a6.heads = heads;
    a6.d = d;
    a6.lk16 = ((L - 1 + 15) / 16) * 16;
    a6.n_items = F * heads;
    a6.out = reinterpret_cast<__nv_bfloat16*>(out);
    a6.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
    a6.var = var6;
    CUtensorMap tm6, tm6b;
    const uint64_t dims6[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)F};
    const uint64_t strides6[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
    const uint32_t box6[3] = {HD, 128, 1};
    const uint32_t box6b[3] = {HD, (uint32_t)(a6.lk16 - 128), 1};
    VMC_TRY(vmc_encode_tmap_bf16(&tm6, qkv, 3, dims6, strides6, box6));
    VMC_TRY(vmc_encode_tmap_bf16(&tm6b, qkv, 3, dims6, strides6, box6b));
    const uint32_t smem6 = 6u * (uint32_t)a6.lk16 * 128u + 2048u + 5u * 1024u + 512u + 256u + 1024u;
    cudaStream_t st6 = reinterpret_cast<cudaStream_t>(stream);
    const int grid6 = a6.n_items < vmc_num_sms() ? a6.n_items : vmc_num_sms();
#ifdef VMC_WHATIF
    auto kern6 = var6 == 0 ? attention_vit6_kernel<0>
                 : var6 == 1 ? attention_vit6_kernel<1>
                 : var6 == 2 ? attention_vit6_kernel<2>
                 : var6 == 4 ? attention_vit6_kernel<4>
                 : var6 == 7 ? attention_vit6_kernel<7> : attention_vit6_kernel<31>;
#else
    auto kern6 = attention_vit6_kernel<0>;
#endif
    VMC_CUDA(cudaFuncSetAttribute(kern6, cudaFuncAttributeMaxDynamicSharedMemorySize, smem6));
    {
      VmcProfScope prof(VMC_K_ATTN_VIT, st6, 4.0 * F * heads * (double)L * L * HD, 8.0 * F * L * d);
      kern6<<<grid6, A6_THREADS, smem6, st6>>>(tm6, tm6b, a6);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  if (impl >= 5 && (L <= 128 || L > 224)) impl = 2;  // v5 covers 129..224 tokens (two Q/K and two V slots + the ones tile in smem)
  if (impl == 5 || impl >= 51) {
    Attn5Args a5;
    a5.L = L;
    a5.heads = heads;
    a5.d = d;
    a5.lk16 = ((L + 15) / 16) * 16;
    a5.n_items = F * heads;
    a5.out = reinterpret_cast<__nv_bfloat16*>(out);
    a5.dbg = reinterpret_cast<long long*>((uintptr_t)(unsigned long long)vmc_get_option64(VMC_OPT_DEBUG_PTR));
    // packed [q | k | v] column blocks.  (A per-head interleaved layout [q_h | k_h | v_h], which would make
    // the three slices of an item adjacent in DRAM, was timed and makes no difference: 0.372 ms either way.)
    a5.cq = 0; a5.ck = d; a5.cv = 2 * d; a5.chead = HD;
    CUtensorMap tm5;
    const uint64_t dims5[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)F};
    const uint64_t strides5[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
    const uint32_t box5[3] = {HD, 128, 1};
    VMC_TRY(vmc_encode_tmap_bf16(&tm5, qkv, 3, dims5, strides5, box5));
    CUtensorMap tm5b;  // second Q / K / V tile: rows 128 .. Lk16 - 1
    const uint32_t box5b[3] = {HD, (uint32_t)(a5.lk16 - 128), 1};
    VMC_TRY(vmc_encode_tmap_bf16(&tm5b, qkv, 3, dims5, strides5, box5b));
    // shared memory: qk_stages x [Q | K] + 2 x V + two all-ones blocks + barriers + alignment slack
    auto smem5_for = [&](int stages) { return (uint32_t)(2 * stages + 2) * (uint32_t)a5.lk16 * 128u + 4096u + 256u + 1024u; };
    a5.qk_stages = smem5_for(3) <= 227u * 1024u ? 3 : 2;
    const uint32_t smem5 = smem5_for(a5.qk_stages);
    cudaStream_t st5 = reinterpret_cast<cudaStream_t>(stream);
    const int grid5 = a5.n_items < vmc_num_sms() ? a5.n_items : vmc_num_sms();
#ifdef VMC_WHATIF
    auto kern5 = impl == 5 ? attention_vit5_kernel<0>
                 : impl == 56 ? attention_vit5_kernel<4>
                 : impl == 57 ? attention_vit5_kernel<5>
                 : impl == 58 ? attention_vit5_kernel<6>
                 : impl == 51 ? attention_vit5_kernel<1>
                 : impl == 52 ? attention_vit5_kernel<2>
                 : impl == 53 ? attention_vit5_kernel<3>
                 : impl == 54 ? attention_vit5_kernel<1, true> : attention_vit5_kernel<0, true>;  // 54 / 55: timeline stamps
#else
    auto kern5 = attention_vit5_kernel<0>;
#endif
    VMC_CUDA(cudaFuncSetAttribute(kern5, cudaFuncAttributeMaxDynamicSharedMemorySize, smem5));
    {
      VmcProfScope prof(VMC_K_ATTN_VIT, st5, 4.0 * F * heads * (double)L * L * HD, 8.0 * F * L * d);
      kern5<<<grid5, A5B_THREADS, smem5, st5>>>(tm5, tm5b, a5);
    }
    VMC_LAUNCH_CHECK();
    vmc_count_launch();
    return VMC_OK;
  }
  AttnArgs a;
  a.F = F;
  a.L = L;
  a.heads = heads;
  a.d = d;
  a.n_mt = (L + 127) / 128;
  a.lk16 = ((L + 15) / 16) * 16;
  const int lk32 = ((a.lk16 + 31) / 32) * 32;
  a.n_pkb = 0;
  a.o_col = ((lk32 / 2 + 31) / 32) * 32;  // first 32-aligned column past the packed P
  const int need = lk32 > a.o_col + HD ? lk32 : a.o_col + HD;
  a.tmem_cols = pow2_at_least(need);
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
  {
    VMC_CUDA(cudaFuncSetAttribute(attention_vit_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VmcProfScope prof(VMC_K_ATTN_VIT, st, fl, 8.0 * F * L * d);
    attention_vit_kernel<true><<<grid, 128, smem, st>>>(tm, a);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_attention_cls(const void* q_cls, const void* kv, void* out, int F, int L, int heads, void* stream) {
  VMC_CHECK_ARG(q_cls && kv && out, VMC_ERR_ARG, "vmc_attention_cls: null pointer");
  VMC_CHECK_ARG(F > 0 && F <= 65535 && heads > 0 && L > 0 && L <= 8192, VMC_ERR_SHAPE, "vmc_attention_cls: bad shape F=%d L=%d", F, L);
  VMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(q_cls) | reinterpret_cast<uintptr_t>(kv)) & 15) == 0, VMC_ERR_ALIGN,
                "vmc_attention_cls: q_cls and kv must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    VmcProfScope prof(VMC_K_ATTN_VIT, st, 4.0 * F * heads * (double)L * HD, 4.0 * F * (double)L * heads * HD);
    attention_cls_kernel<<<dim3(heads, F), 128, (size_t)L * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(q_cls), reinterpret_cast<const __nv_bfloat16*>(kv),
        reinterpret_cast<__nv_bfloat16*>(out), L, heads);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

int vmc_attention_vit(const void* qkv, void* out, int F, int L, int heads, void* stream) {
  const int opt = vmc_get_option(VMC_OPT_ATTN_IMPL);
  return vmc_attention_vit_impl(qkv, out, F, L, heads, (opt == 2 || (opt >= 5 && opt <= 8)) ? opt : 5, stream);
}

int vmc_attention_masked(const float* q, long long ldq, const float* k, long long ldk,
                         const float* v, long long ldv, const uint8_t* key_valid, void* out,
                         int out_f32, long long ldo, int B, int Tq, int Tk, int heads,
                         void* stream) {
  return vmc_attention_masked_train(q, ldq, k, ldk, v, ldv, key_valid, nullptr, out, out_f32, ldo, B, Tq, Tk, heads,
                                    stream);
}

int vmc_attention_masked_train(const float* q, long long ldq, const float* k, long long ldk,
                               const float* v, long long ldv, const uint8_t* key_valid, const float* prob_mask,
                               void* out, int out_f32, long long ldo, int B, int Tq, int Tk, int heads,
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
                                                  out, out_f32, ldo, Tq, Tk, prob_mask);
  }
  VMC_LAUNCH_CHECK();
  vmc_count_launch();
  return VMC_OK;
}

}  // extern "C"
