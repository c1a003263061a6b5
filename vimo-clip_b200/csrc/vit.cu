// Whole CLIP vision tower (OpenAI VisionTransformer == HF CLIPVisionTransformer +
// visual_projection under the weight mapping of SURVEY.md Appendix A), orchestrated on
// one stream from the kernels of this library.  No host synchronisation.
//
// Data layout in HBM for F frames in flight, L = (image/patch)^2 + 1 tokens, d = width:
//   x     fp32 [F*L, d]      residual stream (fp32 so 12-24 pre-LN blocks stay within tolerance)
//   xn    bf16 [F*L, d]      ln_1 output / attention output (A operand of the next GEMM)
//   xn2   bf16 [F*L, d]      ln_2 output
//   big   bf16 [F*L, 4d]     qkv ([F*L, 3d]) and MLP hidden ([F*L, 4d]) share this buffer
//   cls   bf16 [F, d]        ln_post(CLS rows)
// Reference: models/student_model.py:84 (self.visual_encoder(x)), extract_embeddings.py:94
// (clip_model.get_image_features(pixel_values)).
#include "common.cuh"
#include "vimoclip_b200.h"

int vmc_get_option(int option);

namespace {

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* x;
  void* xn;
  void* xn2;
  void* big;
  void* cls;
  float* stats;  // float2 [8][rows]: partial row statistics for the folded LayerNorms
  long long total;
};

Workspace carve(const vmc_vit_model* m, int F, void* base) {
  const int g = m->image / m->patch;
  const long long L = (long long)g * g + 1;
  const long long rows = (long long)F * L;
  const long long d = m->width;
  Workspace w;
  long long off = 0;
  char* p = reinterpret_cast<char*>(base);
  w.x = reinterpret_cast<float*>(p + off);
  off += align_up(rows * d * 4, 1024);
  w.xn = p + off;
  off += align_up(rows * d * 2, 1024);
  w.xn2 = p + off;  // ln_2 output: the out_proj GEMM reads xn (attention output) while its epilogue writes this
  off += align_up(rows * d * 2, 1024);
  w.big = p + off;
  off += align_up(rows * 4 * d * 2, 1024);
  w.cls = p + off;
  off += align_up((long long)F * d * 2, 1024);
  w.stats = reinterpret_cast<float*>(p + off);
  off += align_up(rows * 8 * 8, 1024);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

long long vmc_vit_workspace_bytes(const vmc_vit_model* m, int F) {
  if (!m || F <= 0 || m->patch <= 0) return -1;
  return carve(m, F, nullptr).total;
}

int vmc_vit_forward(const vmc_vit_model* m, const void* patches, float* out, int F,
                    void* workspace, long long workspace_bytes, void* stream) {
  VMC_CHECK_ARG(m && patches && out && workspace, VMC_ERR_ARG, "vmc_vit_forward: null pointer");
  VMC_CHECK_ARG(F > 0, VMC_ERR_SHAPE, "vmc_vit_forward: F must be positive");
  VMC_CHECK_ARG(m->patch > 0 && m->image % m->patch == 0 && m->width == m->heads * 64 &&
                    m->layers > 0 && m->layer != nullptr,
                VMC_ERR_SHAPE, "vmc_vit_forward: unsupported model geometry (head_dim must be 64)");
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, VMC_ERR_ALIGN,
                "vmc_vit_forward: workspace must be 1024-byte aligned");
  const Workspace w = carve(m, F, workspace);
  VMC_CHECK_ARG(workspace_bytes >= w.total, VMC_ERR_WORKSPACE,
                "vmc_vit_forward: workspace too small (%lld < %lld bytes)", workspace_bytes,
                w.total);
  const int g = m->image / m->patch;
  const int n = g * g;
  const int L = n + 1;
  const int d = m->width;
  const int rows = F * L;
  const int kpatch = 3 * m->patch * m->patch;

  // conv1 as a GEMM over patchified frames; epilogue adds positional_embedding[1 + token] and
  // scatters token rows past each frame's CLS row.
  {
    vmc_gemm_epilogue e = {};
    e.resid = m->pos;
    e.ldr = d;
    e.out = w.x;
    e.ldo = d;
    e.out_bf16 = 0;
    e.act = VMC_ACT_NONE;
    e.alpha = 1.0f;
    e.row_group = n;
    VMC_TRY(vmc_gemm_bf16(patches, m->ld_patch, m->w_patch, m->ld_patch, F * n, d, kpatch, &e,
                          stream));
  }
  // LayerNorm FOLDING (VMC_OPT_LN_FUSE = 3 or 5): ln_1 / ln_2 never run as kernels.  The GEMMs
  // that write the residual stream (ln_pre here, then out_proj and c_proj) also emit the bf16 copy of their rows (xb)
  // and per-row partial (sum, sum of squares); the qkv / c_fc GEMMs run on xb with gamma folded into the weights and
  // apply mean / rstd in their epilogue.  Removes 2 x layers LayerNorm passes (6 bytes per element each).
  const int ln_opt = vmc_get_option(VMC_OPT_LN_FUSE);
  // Measured per 1024 ViT-B/16 frames (tools/kernel_bench.py): the consumer epilogues cost +0.03 ms (qkv) and +0.07 ms
  // (c_fc, already the heaviest epilogue: QuickGELU); the producer outputs are free behind the compute-bound c_proj but
  // cost +0.09..0.14 ms in the HBM-bound out_proj, against 0.14 ms per stand-alone LayerNorm.  A/B inside one process on
  // the 256-clip step (tools/ab_ln.py): separate kernels 215.7 ms, ln_1 folded (5) 213.3 ms, both folded (3) 212.0 ms.
  // The 1.7 % is real (fewer HBM bytes under the power cap) but moves LayerNorm work into the GEMM epilogues (GEMM class
  // 1126 -> 949 TFLOP/s), so the DEFAULT stays with separate LayerNorm kernels and the fold is opt-in.
  bool fold = (ln_opt == 3 || ln_opt == 5) && vmc_get_option(VMC_OPT_GEMM_IMPL) != 1 && (d % 128) == 0;
  const bool fold2 = ln_opt == 3;
  for (int i = 0; i < m->layers && fold; ++i)
    fold = m->layer[i].w_qkv_f && m->layer[i].b_qkv_f && m->layer[i].cs_qkv && m->layer[i].w_fc1_f &&
           m->layer[i].b_fc1_f && m->layer[i].cs_fc1;
  const int res_parts = vmc_gemm_stats_parts(rows, d);  // column slices written by the out_proj / c_proj GEMMs
  fold = fold && res_parts > 0 && res_parts <= 8 && (d % (d / res_parts)) == 0;
  // ln_pre in place on the residual stream; CLS rows are sourced from class_embedding + pos[0].
  VMC_TRY(vmc_layernorm_stats(w.x, d, m->ln_pre_g, m->ln_pre_b, 1e-5f, w.x, d, fold ? w.xn2 : nullptr, d, 0, rows,
                              d, m->cls_pos0, L, fold ? w.stats : nullptr, stream));
  if (fold) {
    int parts = 1;  // ln_pre wrote one plane; the residual GEMMs write res_parts
    for (int i = 0; i < m->layers; ++i) {
      const vmc_vit_layer& ly = m->layer[i];
      {  // qkv = LN1(x) Wqkv^T + b, on the raw rows
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_qkv_f;
        e.out = w.big;
        e.ldo = 3 * d;
        e.out_bf16 = 1;
        e.alpha = 1.0f;
        e.stats_in = w.stats;
        e.stats_parts = parts;
        e.stats_ld = rows;
        e.colsum = ly.cs_qkv;
        e.ln_eps = 1e-5f;
        VMC_TRY(vmc_gemm_bf16(w.xn2, d, ly.w_qkv_f, d, rows, 3 * d, d, &e, stream));
      }
      VMC_TRY(vmc_attention_vit(w.big, w.xn, F, L, m->heads, stream));
      {  // x += out_proj(attn); with ln_2 folded it also emits xb + statistics
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_out;
        e.resid = w.x;
        e.ldr = d;
        e.out = w.x;
        e.ldo = d;
        e.alpha = 1.0f;
        if (fold2) {
          e.raw16_out = w.xn2;
          e.raw16_ld = d;
          e.stats_out = w.stats;
          e.stats_ld = rows;
        }
        VMC_TRY(vmc_gemm_bf16(w.xn, d, ly.w_out, d, rows, d, d, &e, stream));
      }
      parts = res_parts;
      if (!fold2)
        VMC_TRY(vmc_layernorm(w.x, d, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr, 0, w.xn2, d, 0, rows, d, nullptr, 0, stream));
      {  // h = QuickGELU(LN2(x) Wfc1^T + b)
        vmc_gemm_epilogue e = {};
        e.bias = fold2 ? ly.b_fc1_f : ly.b_fc1;
        e.out = w.big;
        e.ldo = 4 * d;
        e.out_bf16 = 1;
        e.act = VMC_ACT_QUICKGELU;
        e.alpha = 1.0f;
        if (fold2) {
          e.stats_in = w.stats;
          e.stats_parts = parts;
          e.stats_ld = rows;
          e.colsum = ly.cs_fc1;
          e.ln_eps = 1e-5f;
        }
        VMC_TRY(vmc_gemm_bf16(w.xn2, d, fold2 ? ly.w_fc1_f : ly.w_fc1, d, rows, 4 * d, d, &e, stream));
      }
      {  // x += c_proj(h); emits xb + statistics for the next layer's ln_1
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_fc2;
        e.resid = w.x;
        e.ldr = d;
        e.out = w.x;
        e.ldo = d;
        e.alpha = 1.0f;
        e.raw16_out = w.xn2;
        e.raw16_ld = d;
        e.stats_out = w.stats;
        e.stats_ld = rows;
        VMC_TRY(vmc_gemm_bf16(w.big, 4 * d, ly.w_fc2, 4 * d, rows, d, 4 * d, &e, stream));
      }
    }
  }

  // Optional LayerNorm fusion (VMC_OPT_LN_FUSE = 1: both, 2: only c_proj -> next ln_1; default 0 = off):
  // ln_2 rides in the epilogue of the out_proj GEMM and the NEXT layer's ln_1 in the epilogue of the c_proj
  // GEMM (the CTA pair that owns a row block normalises it out of L2).  Bit-identical to the separate
  // kernels, but MEASURED SLOWER on the 256-clip step (233 ms vs 215 ms).  Split by experiment: the
  // row-owner tile order alone costs +12 ms of GEMM time (the 74 concurrently live A blocks, 116 MB at
  // K = 3072, no longer fit the L2 that the default order shares between pairs), and the in-epilogue
  // LayerNorm passes cost +22 ms (8 warps per SM walking rows serially are latency-bound) against the
  // 19 ms the stand-alone kernels take at 85-100 % of HBM peak.  Kept as a selectable, tested variant.
  const long long pair_tiles = (long long)((rows + 255) / 256) * ((d + 255) / 256);
  const bool fuse_ln = vmc_get_option(VMC_OPT_GEMM_IMPL) != 1 && (ln_opt == 1 || ln_opt == 2) &&
                       d <= 1024 && (d % 128) == 0 && pair_tiles >= vmc_num_sms();
  const bool fuse_ln2 = fuse_ln && ln_opt == 1;
  // Opt-in (VMC_OPT_LAST_BLOCK_CLS = 1): the tower returns ln_post(x[:, 0]) @ proj, so the LAST block's attention / MLP
  // output is only ever read at the CLS row.  K and V still come from all tokens; the query, out_proj, ln_2, c_fc and
  // c_proj run on the F CLS rows only (20 of the block's 24 L d^2 GEMM FLOPs and its attention disappear; the embeddings
  // are the same numbers).  Off by default: the bench measures the reference's full per-token work.
  const bool cls_only = vmc_get_option(VMC_OPT_LAST_BLOCK_CLS) == 1 && !fold && !fuse_ln;
  float* xcls = nullptr;  // [F, d] fp32: the CLS rows after the last block
  for (int i = 0; i < m->layers && !fold; ++i) {
    const vmc_vit_layer& ly = m->layer[i];
    if (cls_only && i == m->layers - 1) {
      // scratch carved from xn2 (only F rows of it are needed from here on)
      char* sp = reinterpret_cast<char*>(w.xn2);
      void* q_cls = sp;                                                   // bf16 [F, d]
      void* a_cls = sp + (size_t)F * d * 2;                               // bf16 [F, d]
      void* n_cls = sp + (size_t)F * d * 4;                               // bf16 [F, d]
      xcls = reinterpret_cast<float*>(sp + (size_t)F * d * 6);            // fp32 [F, d]
      void* h_cls = sp + (size_t)F * d * 10;                              // bf16 [F, 4d]  (18 F d bytes <= 2 F L d)
      const char* wq = reinterpret_cast<const char*>(ly.w_qkv);
      VMC_TRY(vmc_layernorm(w.x, d, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr, 0, w.xn, d, 0, rows, d, nullptr, 0, stream));
      {  // k, v of every token
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_qkv + d;
        e.out = w.big;
        e.ldo = 2 * d;
        e.out_bf16 = 1;
        e.alpha = 1.0f;
        VMC_TRY(vmc_gemm_bf16(w.xn, d, wq + (size_t)d * d * 2, d, rows, 2 * d, d, &e, stream));
      }
      {  // q of the CLS rows (row stride L * d)
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_qkv;
        e.out = q_cls;
        e.ldo = d;
        e.out_bf16 = 1;
        e.alpha = 1.0f;
        VMC_TRY(vmc_gemm_bf16(w.xn, (long long)L * d, wq, d, F, d, d, &e, stream));
      }
      VMC_TRY(vmc_attention_cls(q_cls, w.big, a_cls, F, L, m->heads, stream));
      {
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_out;
        e.resid = w.x;
        e.ldr = (long long)L * d;
        e.out = xcls;
        e.ldo = d;
        e.alpha = 1.0f;
        VMC_TRY(vmc_gemm_bf16(a_cls, d, ly.w_out, d, F, d, d, &e, stream));
      }
      VMC_TRY(vmc_layernorm(xcls, d, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr, 0, n_cls, d, 0, F, d, nullptr, 0, stream));
      {
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_fc1;
        e.out = h_cls;
        e.ldo = 4 * d;
        e.out_bf16 = 1;
        e.act = VMC_ACT_QUICKGELU;
        e.alpha = 1.0f;
        VMC_TRY(vmc_gemm_bf16(n_cls, d, ly.w_fc1, d, F, 4 * d, d, &e, stream));
      }
      {
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_fc2;
        e.resid = xcls;
        e.ldr = d;
        e.out = xcls;
        e.ldo = d;
        e.alpha = 1.0f;
        VMC_TRY(vmc_gemm_bf16(h_cls, 4 * d, ly.w_fc2, 4 * d, F, d, 4 * d, &e, stream));
      }
      break;
    }
    // x = x + out_proj(attn(ln_1(x)))
    if (i == 0 || !fuse_ln)
      VMC_TRY(vmc_layernorm(w.x, d, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr, 0, w.xn, d, 0, rows, d, nullptr,
                            0, stream));
    {
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_qkv;
      e.out = w.big;
      e.ldo = 3 * d;
      e.out_bf16 = 1;
      e.alpha = 1.0f;
      VMC_TRY(vmc_gemm_bf16(w.xn, d, ly.w_qkv, d, rows, 3 * d, d, &e, stream));
    }
    VMC_TRY(vmc_attention_vit(w.big, w.xn, F, L, m->heads, stream));
    {
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_out;
      e.resid = w.x;
      e.ldr = d;
      e.out = w.x;
      e.ldo = d;
      e.out_bf16 = 0;
      e.alpha = 1.0f;
      if (fuse_ln2) {  // ln_2: the attention output in xn has been consumed by this very GEMM
        e.ln_gamma = ly.ln2_g;
        e.ln_beta = ly.ln2_b;
        e.ln_out = w.xn2;
        e.ln_ldo = d;
        e.ln_eps = 1e-5f;
      }
      VMC_TRY(vmc_gemm_bf16(w.xn, d, ly.w_out, d, rows, d, d, &e, stream));
    }
    // x = x + c_proj(QuickGELU(c_fc(ln_2(x))))
    if (!fuse_ln2)
      VMC_TRY(vmc_layernorm(w.x, d, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr, 0, w.xn2, d, 0, rows, d, nullptr,
                            0, stream));
    {
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_fc1;
      e.out = w.big;
      e.ldo = 4 * d;
      e.out_bf16 = 1;
      e.act = VMC_ACT_QUICKGELU;
      e.alpha = 1.0f;
      VMC_TRY(vmc_gemm_bf16(w.xn2, d, ly.w_fc1, d, rows, 4 * d, d, &e, stream));
    }
    {
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_fc2;
      e.resid = w.x;
      e.ldr = d;
      e.out = w.x;
      e.ldo = d;
      e.out_bf16 = 0;
      e.alpha = 1.0f;
      if (fuse_ln && i + 1 < m->layers) {  // next layer's ln_1
        e.ln_gamma = m->layer[i + 1].ln1_g;
        e.ln_beta = m->layer[i + 1].ln1_b;
        e.ln_out = w.xn;
        e.ln_ldo = d;
        e.ln_eps = 1e-5f;
      }
      VMC_TRY(vmc_gemm_bf16(w.big, 4 * d, ly.w_fc2, 4 * d, rows, d, 4 * d, &e, stream));
    }
  }
  // ln_post on the CLS rows (row stride L*d), then @ proj (no bias), fp32 out.
  VMC_TRY(vmc_layernorm(xcls ? xcls : w.x, xcls ? (long long)d : (long long)L * d, m->ln_post_g, m->ln_post_b, 1e-5f,
                        nullptr, 0, w.cls, d, 0, F, d, nullptr, 0, stream));
  {
    vmc_gemm_epilogue e = {};
    e.out = out;
    e.ldo = m->out_dim;
    e.out_bf16 = 0;
    e.alpha = 1.0f;
    VMC_TRY(vmc_gemm_bf16(w.cls, d, m->w_proj, d, F, m->out_dim, d, &e, stream));
  }
  return VMC_OK;
}

}  // extern "C"
