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
  float* stats;  // float2 [16][rows]: partial row statistics for the folded LayerNorms (one plane per 64- / 128-column slice)
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
  off += align_up(rows * 16 * 8, 1024);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

long long vmc_vit_workspace_bytes(const vmc_vit_model* m, int F) {
  if (!m || F <= 0 || m->patch <= 0) return -1;
  return carve(m, F, nullptr).total;
}

// Tower variants (vmc_vit_model.ln_mode, 0 = the process default of vmc_set_option(VMC_OPT_LN_FUSE), 0 there = 6):
//   6  bf16 residual stream, both LayerNorms folded into the consuming GEMMs (DEFAULT)
//   3  fp32 residual stream + bf16 copy, both LayerNorms folded        5  fp32 stream, only ln_1 folded
//   4  fp32 residual stream, separate LayerNorm kernels (the round-1 default; cross-check in the tests)
static int tower_ln_mode(const vmc_vit_model* m) {
  int v = m->ln_mode != 0 ? m->ln_mode : vmc_get_option(VMC_OPT_LN_FUSE);
  return v == 0 ? 6 : v;
}
// 1 = the last block computes only the CLS row of its output (default), 2 = full last block
static int tower_last_block_cls(const vmc_vit_model* m) {
  int v = m->last_block_cls != 0 ? m->last_block_cls : vmc_get_option(VMC_OPT_LAST_BLOCK_CLS);
  return v == 0 ? 1 : v;
}

int vmc_vit_forward(const vmc_vit_model* m, const void* patches, float* out, int F, void* workspace,
                    long long workspace_bytes, void* stream) {
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
  const int attn_impl = m->attn_impl;  // 0 = by sequence length (vmc_attention_vit)
  auto attention = [&](const void* qkv, void* o) {
    return attn_impl ? vmc_attention_vit_impl(qkv, o, F, L, m->heads, attn_impl, stream)
                     : vmc_attention_vit(qkv, o, F, L, m->heads, stream);
  };

  // conv1 as a GEMM over patchified frames; epilogue adds positional_embedding[1 + token] and
  // scatters token rows past each frame's CLS row (fp32: this one buffer is read once, by ln_pre).
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
  int ln_mode = tower_ln_mode(m);
  VMC_CHECK_ARG(ln_mode >= 3 && ln_mode <= 6, VMC_ERR_ARG, "vmc_vit_forward: unknown ln_mode %d", ln_mode);
  // LayerNorm FOLDING: ln_1 / ln_2 never run as kernels.  The GEMMs that write the residual stream (ln_pre here, then
  // out_proj and c_proj) emit per-row partial (sum, sum of squares); the qkv / c_fc GEMMs run on the raw rows with gamma
  // folded into the weights and apply mean / rstd in their epilogue.
  bool have_folded = vmc_get_option(VMC_OPT_GEMM_IMPL) != 1 && (d % 128) == 0;
  for (int i = 0; i < m->layers && have_folded; ++i)
    have_folded = m->layer[i].w_qkv_f && m->layer[i].b_qkv_f && m->layer[i].cs_qkv && m->layer[i].w_fc1_f &&
                  m->layer[i].b_fc1_f && m->layer[i].cs_fc1;
  const int res_parts = vmc_gemm_stats_parts(rows, d);  // column slices written by the out_proj / c_proj GEMMs
  have_folded = have_folded && res_parts > 0 && res_parts <= 16 && (d % (d / res_parts)) == 0;
  if (!have_folded) ln_mode = 4;
  const bool cls_only = tower_last_block_cls(m) == 1 && (ln_mode == 6 || ln_mode == 4);
  float* xcls = nullptr;  // [F, d] fp32: the CLS rows after the last block

  // The LAST block, CLS rows only: the tower returns ln_post(x[:, 0]) @ proj, so only the CLS row of the last block's
  // attention / MLP output is ever read.  K and V still come from all tokens (kv bf16 [rows, 2d] in w.big is written by
  // the caller); the query, out_proj, ln_2, c_fc and c_proj run on the F CLS rows (20 of the block's 24 L d^2 GEMM FLOPs
  // and its L x L attention disappear; the embeddings are the same numbers).  xsrc: the residual stream (fp32 or bf16).
  auto last_block_cls_rows = [&](const vmc_vit_layer& ly, const void* xsrc, int x_bf16, char* sp) -> int {
    void* q_cls = sp;                                          // bf16 [F, d]
    void* a_cls = sp + (size_t)F * d * 2;                      // bf16 [F, d]
    void* n_cls = sp + (size_t)F * d * 4;                      // bf16 [F, d]
    xcls = reinterpret_cast<float*>(sp + (size_t)F * d * 6);   // fp32 [F, d]
    void* h_cls = sp + (size_t)F * d * 10;                     // bf16 [F, 4d]  (18 F d bytes in all)
    VMC_TRY(vmc_layernorm_ex(xsrc, x_bf16, (long long)L * d, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr, 0, n_cls, d, 0, F, d,
                             nullptr, 0, nullptr, 0, stream));
    {  // q of the CLS rows
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_qkv;
      e.out = q_cls;
      e.ldo = d;
      e.out_bf16 = 1;
      e.alpha = 1.0f;
      VMC_TRY(vmc_gemm_bf16(n_cls, d, ly.w_qkv, d, F, d, d, &e, stream));
    }
    VMC_TRY(vmc_attention_cls(q_cls, w.big, a_cls, F, L, m->heads, stream));
    {
      vmc_gemm_epilogue e = {};
      e.bias = ly.b_out;
      e.resid = reinterpret_cast<const float*>(xsrc);
      e.resid_bf16 = x_bf16;
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
    return VMC_OK;
  };

  if (ln_mode == 6) {
    // ---------------- DEFAULT: bf16 residual stream ----------------
    // xs (bf16 [rows, d], in w.xn2) is the residual stream AND the A operand of the qkv / c_fc GEMMs: out_proj and c_proj
    // read it as the residual and write bf16(acc + bias + xs) back in place with the row statistics of the rounded
    // values.  Per element of the stream and layer: 4 + 4 bytes instead of 10 + 10 (fp32 read + fp32 write + bf16 copy).
    // Precision (tools/emulate_residual_precision.py, tests): embedding cosine 0.99995 vs 0.999994 with an fp32 stream
    // (bar 0.9995), TFAM logit error unchanged (dominated by the bf16 GEMM operands either way).
    void* xs = w.xn2;
    VMC_TRY(vmc_layernorm_ex(w.x, 0, d, m->ln_pre_g, m->ln_pre_b, 1e-5f, nullptr, 0, xs, d, 0, rows, d, m->cls_pos0, L,
                             w.stats, 1, stream));
    int parts = 1;  // ln_pre wrote one plane; the residual GEMMs write res_parts
    auto fold_consumer = [&](vmc_gemm_epilogue& e, const float* bias, const float* colsum) {
      e.bias = bias;
      e.out = w.big;
      e.out_bf16 = 1;
      e.alpha = 1.0f;
      e.stats_in = w.stats;
      e.stats_parts = parts;
      e.stats_ld = rows;
      e.colsum = colsum;
      e.ln_eps = 1e-5f;
    };
    auto resid_producer = [&](vmc_gemm_epilogue& e, const float* bias) {
      e.bias = bias;
      e.resid = reinterpret_cast<const float*>(xs);
      e.resid_bf16 = 1;
      e.ldr = d;
      e.out = xs;
      e.ldo = d;
      e.out_bf16 = 1;
      e.alpha = 1.0f;
      e.stats_out = w.stats;
      e.stats_ld = rows;
    };
    for (int i = 0; i < m->layers; ++i) {
      const vmc_vit_layer& ly = m->layer[i];
      if (cls_only && i == m->layers - 1) {
        {  // k, v of every token: LN1 folded, weight rows [d, 3d)
          vmc_gemm_epilogue e = {};
          fold_consumer(e, ly.b_qkv_f + d, ly.cs_qkv + d);
          e.ldo = 2 * d;
          VMC_TRY(vmc_gemm_bf16(xs, d, reinterpret_cast<const char*>(ly.w_qkv_f) + (size_t)d * d * 2, d, rows, 2 * d, d, &e,
                                stream));
        }
        // scratch: the fp32 patch-embedding buffer is dead after ln_pre (18 F d <= 4 F L d bytes)
        VMC_TRY(last_block_cls_rows(ly, xs, 1, reinterpret_cast<char*>(w.x)));
        break;
      }
      {  // qkv = LN1(x) Wqkv^T + b, on the raw rows
        vmc_gemm_epilogue e = {};
        fold_consumer(e, ly.b_qkv_f, ly.cs_qkv);
        e.ldo = 3 * d;
        VMC_TRY(vmc_gemm_bf16(xs, d, ly.w_qkv_f, d, rows, 3 * d, d, &e, stream));
      }
      VMC_TRY(attention(w.big, w.xn));
      {  // x += out_proj(attn)
        vmc_gemm_epilogue e = {};
        resid_producer(e, ly.b_out);
        VMC_TRY(vmc_gemm_bf16(w.xn, d, ly.w_out, d, rows, d, d, &e, stream));
      }
      parts = res_parts;
      {  // h = QuickGELU(LN2(x) Wfc1^T + b)
        vmc_gemm_epilogue e = {};
        fold_consumer(e, ly.b_fc1_f, ly.cs_fc1);
        e.ldo = 4 * d;
        e.act = VMC_ACT_QUICKGELU;
        VMC_TRY(vmc_gemm_bf16(xs, d, ly.w_fc1_f, d, rows, 4 * d, d, &e, stream));
      }
      {  // x += c_proj(h)
        vmc_gemm_epilogue e = {};
        resid_producer(e, ly.b_fc2);
        VMC_TRY(vmc_gemm_bf16(w.big, 4 * d, ly.w_fc2, 4 * d, rows, d, 4 * d, &e, stream));
      }
    }
    if (xcls != nullptr)
      VMC_TRY(vmc_layernorm(xcls, d, m->ln_post_g, m->ln_post_b, 1e-5f, nullptr, 0, w.cls, d, 0, F, d, nullptr, 0, stream));
    else
      VMC_TRY(vmc_layernorm_ex(xs, 1, (long long)L * d, m->ln_post_g, m->ln_post_b, 1e-5f, nullptr, 0, w.cls, d, 0, F, d,
                               nullptr, 0, nullptr, 0, stream));
  } else {
    // ---------------- fp32 residual stream (selectable; the round-1 paths, kept as cross-checks) ----------------
    const bool fold = ln_mode == 3 || ln_mode == 5;
    const bool fold2 = ln_mode == 3;
    // ln_pre in place on the residual stream; CLS rows are sourced from class_embedding + pos[0].
    VMC_TRY(vmc_layernorm_stats(w.x, d, m->ln_pre_g, m->ln_pre_b, 1e-5f, w.x, d, fold ? w.xn2 : nullptr, d, 0, rows,
                                d, m->cls_pos0, L, fold ? w.stats : nullptr, stream));
    int parts = 1;
    for (int i = 0; i < m->layers; ++i) {
      const vmc_vit_layer& ly = m->layer[i];
      if (cls_only && i == m->layers - 1) {  // (ln_mode 4 only)
        VMC_TRY(vmc_layernorm(w.x, d, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr, 0, w.xn, d, 0, rows, d, nullptr, 0, stream));
        {  // k, v of every token
          vmc_gemm_epilogue e = {};
          e.bias = ly.b_qkv + d;
          e.out = w.big;
          e.ldo = 2 * d;
          e.out_bf16 = 1;
          e.alpha = 1.0f;
          VMC_TRY(vmc_gemm_bf16(w.xn, d, reinterpret_cast<const char*>(ly.w_qkv) + (size_t)d * d * 2, d, rows, 2 * d, d, &e,
                                stream));
        }
        // scratch carved from xn2 (unused without folding)
        VMC_TRY(last_block_cls_rows(ly, w.x, 0, reinterpret_cast<char*>(w.xn2)));
        break;
      }
      // x = x + out_proj(attn(ln_1(x)))
      if (!fold)
        VMC_TRY(vmc_layernorm(w.x, d, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr, 0, w.xn, d, 0, rows, d, nullptr, 0, stream));
      {
        vmc_gemm_epilogue e = {};
        e.bias = fold ? ly.b_qkv_f : ly.b_qkv;
        e.out = w.big;
        e.ldo = 3 * d;
        e.out_bf16 = 1;
        e.alpha = 1.0f;
        if (fold) {
          e.stats_in = w.stats;
          e.stats_parts = parts;
          e.stats_ld = rows;
          e.colsum = ly.cs_qkv;
          e.ln_eps = 1e-5f;
        }
        VMC_TRY(vmc_gemm_bf16(fold ? w.xn2 : w.xn, d, fold ? ly.w_qkv_f : ly.w_qkv, d, rows, 3 * d, d, &e, stream));
      }
      VMC_TRY(attention(w.big, w.xn));
      {  // x += out_proj(attn); with ln_2 folded it also emits the bf16 rows + statistics
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
      // x = x + c_proj(QuickGELU(c_fc(ln_2(x))))
      if (!fold2)
        VMC_TRY(vmc_layernorm(w.x, d, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr, 0, w.xn2, d, 0, rows, d, nullptr, 0, stream));
      {
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
      {  // x += c_proj(h); when ln_1 is folded it emits the bf16 rows + statistics for the next layer
        vmc_gemm_epilogue e = {};
        e.bias = ly.b_fc2;
        e.resid = w.x;
        e.ldr = d;
        e.out = w.x;
        e.ldo = d;
        e.alpha = 1.0f;
        if (fold) {
          e.raw16_out = w.xn2;
          e.raw16_ld = d;
          e.stats_out = w.stats;
          e.stats_ld = rows;
        }
        VMC_TRY(vmc_gemm_bf16(w.big, 4 * d, ly.w_fc2, 4 * d, rows, d, 4 * d, &e, stream));
      }
    }
    // ln_post on the CLS rows (row stride L*d)
    VMC_TRY(vmc_layernorm(xcls ? xcls : w.x, xcls ? (long long)d : (long long)L * d, m->ln_post_g, m->ln_post_b, 1e-5f,
                          nullptr, 0, w.cls, d, 0, F, d, nullptr, 0, stream));
  }
  {  // @ proj (no bias), fp32 out
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
