/*
 * vimoclip_b200 -- C-ABI of the B200-native ViMoCLIP per-frame encoding hot path.
 *
 * The reference (MarcosRodrigoT/VIMO-CLIP) is pure Python: its "plugin API" for this
 * path is the nn.Module call surface.  This header is the boundary a
 * maintainer binds from Python (ctypes): plain pointers and
 * sizes, no torch types.  Each entry cites the reference interface it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every entry returns int: 0 ok, <0 argument/shape/alignment error,
 *     >0 a cudaError_t; vmc_last_error() gives the thread-local message;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation; the caller owns all buffers including the workspace;
 *   - there is NO CPU fallback: host pointers are undefined behaviour.
 */
#ifndef VIMOCLIP_B200_H_
#define VIMOCLIP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMC_ABI_VERSION 3

/* ---- runtime ---------------------------------------------------------------- */
const char* vmc_last_error(void);
int vmc_abi_version(void);
int vmc_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Implementation selectors (0 = default).  VMC_OPT_GEMM_IMPL: 0/2 = CTA-pair cta_group::2 kernel, 1 = single-CTA kernel
 * (kept as an independent cross-check for the tests), 3 = the CTA-pair kernel with the LDS + STG form of the bf16 epilogues
 * instead of the TMA store (A/B runs).  VMC_OPT_ATTN_IMPL: 0 = by sequence length (7 for L <= 64, 5 for
 * 129..224, 6 for 225..257, else 2); 5 = persistent, split Q/K and V rings, one MMA issuer per query tile, score MMA in two key
 * ranges (the first issued one item ahead), epilogue warps, row sums from the tensor core; 6 = v5 (round-1 pipeline) on the
 * patch tokens + the CLS token on mma.sync warps; 7 = two items per query tile;
 * 8 = warp-level mma.sync kernel for L <= 64; 2 = one CTA per (frame, head) with P in TMEM (any L <= 272). */
/* vmc_set_option sets PROCESS-WIDE defaults (relaxed atomics; meant for tests, A/B runs and tools).  The per-model
 * selectors of vmc_vit_model take precedence, so two models in one process can run different variants concurrently. */
enum { VMC_OPT_ATTN_BWD_IMPL = 4 /* ViT attention backward, L <= 64: 0 = warp-level tensor-core kernel (ldmatrix + mma.sync), 2 = register-tiled fp32 kernel, 1 = first-generation shared-memory kernel (cross-checks) */,
       VMC_OPT_LAST_BLOCK_CLS = 5 /* ViT tower: 0 / 1 = in the LAST block compute only what the output reads (the CLS row): K / V of
                                     all tokens, but query, out_proj, ln_2 and the MLP on the F CLS rows only (default; the
                                     embeddings are the same numbers); 2 = the full last block */,
       VMC_OPT_ATTN_PREFETCH = 6 /* ViT attention: experiments, L <= 64 kernels: 1..8 = L2 prefetch distance of the v7 TMA producer in CTA iterations (measured slower; 0 = off, the default); 81 = impl 8 reads a head-major [F, heads, 3, L, 64] buffer (timing what-if only) */,
       VMC_OPT_DEBUG_PTR = 7 /* device pointer of a clock64 timeline buffer (tools/attn_timeline.py with a WHATIF build, tools/gemm_timeline.py), 0 = off */,
       VMC_OPT_GEMM_IMPL = 0, VMC_OPT_ATTN_IMPL = 1, VMC_OPT_PROLOGUE_IMPL = 2 /* patch-matrix prologue: 0 = gather kernel (output-ordered, default), 1 = direct (input-ordered), 2 = band (smem-staged), 4 = gather with 16 pixels per item (uint8 sources; bit-identical, measured slower: experiment) */,
       VMC_OPT_LN_FUSE = 3 /* ViT tower variant (vmc_vit_model.ln_mode): 0 / 6 = bf16 residual stream, ln_1 and ln_2 FOLDED into
                              the qkv / c_fc GEMMs (default); 3 = fp32 residual stream + bf16 copy, both folded; 5 = fp32
                              stream, only ln_1 folded; 4 = fp32 stream, separate LayerNorm kernels (round-1 default) */ };
int vmc_set_option(int option, long long value);
/* kernels launched by this library since the last reset (bench.py "gpu_launches") */
long long vmc_launch_count(void);
void vmc_reset_launch_count(void);
/* Per-kernel-class timing for the roofline lines of bench.py: while a profile is open every launch is
 * bracketed by CUDA events on its own stream.  Classes: 0 prologue, 1 tcgen05 GEMM, 2 ViT attention,
 * 3 LayerNorm, 4 small (TFAM) attention, 5 other.  vmc_profile_end synchronises the device and fills
 * per-class milliseconds, algorithmic FLOPs, algorithmic bytes and launch counts (arrays of >= 6). */
void vmc_profile_begin(void);
int vmc_profile_end(double* ms, double* flops, double* bytes, long long* launches, int ncat);

/* ---- P1: uint8 prologue -------------------------------------------------------
 * Replaces the per-frame PIL loop of models/student_model.py:74-81 (to_pil_image ->
 * Resize/CenterCrop (identity at 224x224) -> ToTensor -> Normalize) and the
 * cv2.cvtColor + cv2.absdiff of utils/generate_frame_diff_video.py:37,46,49.
 */
enum {
  VMC_SRC_U8 = 0,      /* uint8 pixels 0..255 taken as-is (HF processor path, extract_embeddings.py:91) */
  VMC_SRC_U8_WRAP = 1, /* regime A: uint8 -> .float() -> to_pil_image: u8 = (-x) mod 256 (student_model.py:74,78) */
  VMC_SRC_F32_WRAP = 2, /* regimes B/C: float input, u8 = (int64)trunc(x*255f) mod 256 */
  VMC_SRC_F32_NORM = 3  /* already-normalised fp32 pixel_values (get_image_features input): patchify + bf16 only */
};
enum {
  VMC_DST_U8 = 0,         /* wrapped uint8 [F,3,H,W] (bit-exact integer check) */
  VMC_DST_F32_NCHW = 1,   /* (u8/255 - mean)/std fp32 [F,3,H,W]: what the reference feeds the ViT */
  VMC_DST_BF16_PATCH = 2  /* same value rounded to bf16, patchified [F*(H/p)*(W/p), ld] (GEMM A operand) */
};
/* frames: [F,3,H,W] planar (src_kind u8 or f32).  dst per dst_kind.  ld_patch >= 3*p*p, multiple of 8;
 * pad columns [3*p*p, ld_patch) are zero-filled. */
int vmc_prologue(const void* frames, int src_kind, void* dst, int dst_kind, int F, int H, int W,
                 int patch, int ld_patch, void* stream);
/* BGR uint8 clips [clips, T+1, H, W, 3] (OpenCV layout) -> gray (9798 R + 19235 G + 3735 B + 16384) >> 15
 * -> |gray[t+1]-gray[t]| -> diff_u8 [clips, T, H, W] (may be NULL) and/or the student prologue on the
 * diff replicated to 3 channels (regime A wrap): dst per dst_kind as above with F = clips*T. */
int vmc_frame_diff_prologue(const uint8_t* bgr, uint8_t* diff_u8, void* dst, int dst_kind, int clips,
                            int T, int H, int W, int patch, int ld_patch, void* stream);

/* Non-224x224 frames: Resize(224, BICUBIC) -> CenterCrop(224) of clip's _transform (models/student_model.py:77-78),
 * i.e. Pillow's 8-bit bicubic resampler (22-bit fixed-point coefficients, horizontal then vertical pass, each
 * clipped to uint8) with torchvision's size / crop rules; bit-exact.  The student's to_pil_image wrap (src_kind)
 * is applied to the source pixels first, as in the reference.  frames [F,3,H,W] (u8 or f32), out uint8
 * [F,3,size,size], tmp uint8 [F,3,H,size] scratch.  Feed `out` to vmc_prologue with VMC_SRC_U8. */
int vmc_resize_geometry(int H, int W, int size, int* new_h, int* new_w, int* top, int* left);
int vmc_resize_center_crop(const void* frames, int src_kind, uint8_t* out, uint8_t* tmp, int F, int H, int W,
                           int size, void* stream);
/* crop_floor = 1: the crop offset of HF CLIPImageProcessor, (dim - size) // 2 (extract_embeddings.py:18,91), instead of
 * torchvision's round-half-to-even; they differ by one pixel when dim - size is 3 mod 4.  Frames smaller than `size` are
 * scaled up by the same resampler (Resize scales the SHORT side to `size`, so the crop never pads). */
int vmc_resize_geometry_ex(int H, int W, int size, int crop_floor, int* new_h, int* new_w, int* top, int* left);
int vmc_resize_center_crop_ex(const void* frames, int src_kind, uint8_t* out, uint8_t* tmp, int F, int H, int W,
                              int size, int crop_floor, void* stream);

/* ---- G: tcgen05/TMEM GEMM fed by TMA ------------------------------------------
 * out[orow, n] = alpha * act(sum_k A[m,k] * W[n,k] + bias[n]) + resid[rrow, n]
 * A [M,K] bf16 row-major (lda), W [N,K] bf16 row-major (ldw) == nn.Linear.weight layout.
 * Replaces nn.Conv2d patch embed / nn.Linear / in_proj / out_proj / c_fc / c_proj / @proj
 * reached from models/student_model.py:84 and extract_embeddings.py:94.
 */
enum { VMC_ACT_NONE = 0, VMC_ACT_QUICKGELU = 1, VMC_ACT_GELU_ERF = 2, VMC_ACT_RELU = 3 };
typedef struct vmc_gemm_epilogue {
  const float* bias;  /* [N] fp32 or NULL */
  const float* resid; /* fp32 matrix added after activation, or NULL; may alias out */
  long long ldr;      /* row stride of resid in elements */
  void* out;          /* bf16 or fp32 */
  long long ldo;      /* row stride of out in elements */
  int out_bf16;       /* 1: out is bf16, 0: fp32 */
  int act;            /* VMC_ACT_* */
  float alpha;
  int row_group;      /* 0: orow = rrow = m.  g > 0 (patch embed): f = m / g, orow = m + f + 1,
                         rrow = m - f*g + 1 (token rows of frame f skip the CLS row; resid = pos-emb) */
  float ln_eps;         /* epsilon of the folded LayerNorm below (consumer side) */
  /* LayerNorm FOLDING (no LayerNorm pass at all; CTA-pair kernel only).  For y = LayerNorm(x) * gamma + beta,
   *   y W^T + b = rstd * (x W'^T - mean * colsum) + b',   W' = gamma (.) W,  colsum[n] = sum_k W'[n,k],  b' = b + W beta,
   * so the GEMM that consumes LayerNorm(x) runs on the RAW rows x (bf16) with W', and its epilogue applies the per-row
   * mean / rstd.  Producer side (fp32 out + bias + residual epilogue, i.e. the GEMM that writes the residual stream):
   * raw16_out receives the bf16 copy of the output rows and stats_out per-row partial (sum, sum of squares), one
   * float2 plane of stats_ld rows per vmc_gemm_stats_parts(M, N) column slices.  Consumer side (bf16 out + bias
   * [+ QuickGELU] epilogue): stats_in / stats_parts / colsum describe the rows of A; statistics are over K columns. */
  void* raw16_out;       /* producer: bf16 [M, raw16_ld], or NULL */
  long long raw16_ld;
  float* stats_out;      /* producer: float2 [parts][stats_ld], or NULL */
  const float* stats_in; /* consumer: float2 [stats_parts][stats_ld], or NULL */
  int stats_parts;
  long long stats_ld;
  const float* colsum;   /* consumer: [N] fp32 */
  /* bf16 RESIDUAL STREAM (ViT tower default): resid_bf16 = 1 means `resid` points to bf16 rows.  With out_bf16 = 1, bias,
   * no activation and stats_out set, the epilogue writes out = bf16(acc + bias + resid) -- the residual stream AND the A
   * operand of the next (LayerNorm-folded) GEMM in one buffer, may alias resid -- plus the per-row partial statistics of
   * the ROUNDED values (raw16_out must be NULL); for N >= 256 the residual and result rows move by TMA: resid / out 16-byte
   * aligned with ldr, ldo multiples of 8 elements.  With an fp32 out it is only a dtype switch of the residual read.
   * (All bf16 outputs: out 16-byte aligned, ldo % 8 == 0.) */
  int resid_bf16;
} vmc_gemm_epilogue;
/* number of column slices (partials per row) a producer GEMM of this shape writes to stats_out */
int vmc_gemm_stats_parts(int M, int N);
int vmc_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                  const vmc_gemm_epilogue* epi, void* stream);
/* Same product with operands given TRANSPOSED in memory: a_transposed: A points to A^T [K, M] row-major (lda >= M);
 * w_transposed: W points to W^T [K, N] row-major (ldw >= N).  They become MN-major UMMA operands (TMA boxes of
 * 64 x 64, no transposing pass): the backward GEMMs dW = dY^T X (both transposed) and dX = dY W (W transposed) read
 * the row-major activations / weights as they are (train.py:95-107, TFAM/train_and_eval.py:80-84). */
int vmc_gemm_bf16_ex(const void* A, long long lda, int a_transposed, const void* W, long long ldw, int w_transposed,
                     int M, int N, int K, const vmc_gemm_epilogue* epi, void* stream);

/* ---- E1: LayerNorm (fp32 statistics, eps inside sqrt) ---------------------------
 * y = (x - mean) / sqrt(var + eps) * gamma + beta per row of width d (d % 4 == 0, d <= 4096).
 * x fp32 rows at stride ldx (elements).  Outputs (either may be NULL): y32 fp32 (ld32), y16 bf16 (ld16).
 * If cls_every > 0, rows r with r % cls_every == 0 take their input from cls_row[d] instead of x
 * (CLS token = class_embedding + positional_embedding[0], OpenAI VisionTransformer.forward).
 * y16_split = 1 writes the bf16 output as the 3*d-column "split" operand [hi | lo | hi] (hi = bf16(y),
 * lo = bf16(y - hi)) that, multiplied against weights packed [Whi | Whi | Wlo], gives a near-fp32 GEMM
 * on the bf16 tensor cores (used by the TFAM block to hold logit max-abs <= 1e-2).
 * Replaces ln_pre / ln_1 / ln_2 / ln_post and TFAM norm_self / norm_cross / norm_ffn / classifier.0.
 */
int vmc_layernorm(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                  float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows, int d,
                  const float* cls_row, int cls_every, void* stream);
/* Same, and stats_out[row] = float2(sum, sum of squares) of the fp32 OUTPUT row (NULL = off): the row statistics a
 * LayerNorm folded into the next GEMM needs (vmc_gemm_epilogue.stats_in with stats_parts = 1). */
int vmc_layernorm_stats(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                        float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows,
                        int d, const float* cls_row, int cls_every, float* stats_out, void* stream);

/* General form: x_bf16 = 1 reads bf16 input rows (cls_every must be 0); stats_rounded = 1 makes stats_out the statistics
 * of the bf16-ROUNDED output row, i.e. of the values a LayerNorm-folded GEMM reading y16 multiplies. */
int vmc_layernorm_ex(const void* x, int x_bf16, long long ldx, const float* gamma, const float* beta, float eps,
                     float* y32, long long ld32, void* y16, long long ld16, int y16_split, int rows, int d,
                     const float* cls_row, int cls_every, float* stats_out, int stats_rounded, void* stream);

/* ---- A1: ViT self-attention (no mask), head_dim 64 ------------------------------
 * qkv bf16 [F*L, 3*d] (q | k | v, heads contiguous inside each), out bf16 [F*L, d].
 * softmax(q k^T / 8) v per (frame, head); QK^T and PV on tcgen05, S/O in TMEM.
 * Replaces nn.MultiheadAttention inside OpenAI ResidualAttentionBlock / HF CLIPAttention.
 */
int vmc_attention_vit(const void* qkv, void* out, int F, int L, int heads, void* stream);
/* short sequences (L <= 64) on the warp-level tensor path (ldmatrix + mma.sync), one CTA per (frame, head); = vmc_attention_vit_impl(..., 8, ...) */
int vmc_attention_vit_short_mma(const void* qkv, void* out, int F, int L, int heads, void* stream);
/* CLS-query attention of the last block (VMC_OPT_LAST_BLOCK_CLS): q_cls bf16 [F, d], kv bf16 [F*L, 2d] = [k | v] -> out bf16 [F, d] */
int vmc_attention_cls(const void* q_cls, const void* kv, void* out, int F, int L, int heads, void* stream);
/* same, selecting the implementation (see VMC_OPT_ATTN_IMPL above); kernels that do not cover L fall back to one that does */
int vmc_attention_vit_impl(const void* qkv, void* out, int F, int L, int heads, int impl, void* stream);

/* ---- small masked attention (TFAM), fp32 in, bf16 out ---------------------------
 * q [B*Tq, ldq], k/v [B*Tk, ldk/ldv] fp32 (heads*64 columns used starting at the pointer),
 * key_valid uint8/bool [B, Tk]: 1 = real frame, 0 = padding (the mask_rgb / mask_flow of
 * collate_fn_pad, TFAM/data/dataset.py:86-103, i.e. the inverse of the key_padding_mask built at
 * TFAM/models/AMO_CLIP.py:125-126) or NULL for no mask.
 * out bf16 (out_f32 = 0) or fp32 (out_f32 = 1) [B*Tq, ldo].  Replaces self_attn / cross_attn of AttentionLayer (AMO_CLIP.py:39,44).
 */
int vmc_attention_masked(const float* q, long long ldq, const float* k, long long ldk, const float* v,
                         long long ldv, const uint8_t* key_valid, void* out, int out_f32, long long ldo,
                         int B, int Tq, int Tk, int heads, void* stream);

/* ---- utilities -------------------------------------------------------------- */
/* fp32 [rows, d] (ldx) -> bf16 [rows, d] (ldy), or with split = 1 -> bf16 [rows, 3*d] = [hi | lo | hi] */
int vmc_cast_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int d, int split,
                  void* stream);
/* temporal mean: x fp32 [B, T, d] -> y32 fp32 [B, d] and/or y16 bf16 [B, d] (mean over ALL T rows,
 * models/student_model.py:93, AMO_CLIP.py:170) */
int vmc_mean_rows(const float* x, float* y32, void* y16, int B, int T, int d, void* stream);
/* cosine distillation loss of losses.py:27-40: mean over rows of 1 - clamp(cos(s,t)); out: 1 fp32 */
int vmc_cosine_distill_loss(const float* s, const float* t, int rows, int d, float* out, void* stream);

/* ---- per-clip fused heads (fp32; weights TRANSPOSED [K, N] fp32) ---------------------------------
 * vmc_student_heads: models/student_model.py:33-35,90-96 for one clip per CTA: distill [B,T,D] =
 *   emb + alpha * fc2(GELU_erf(fc1(emb))); logits [B,C] = W2 relu(W1 mean_t(emb) + b1) + b2 (RAW embeddings pooled).
 * vmc_tfam_head: TFAM/models/AMO_CLIP.py:170 for one clip per CTA: logits [B,C] =
 *   Linear(GELU_erf(Linear(LayerNorm(mean over ALL T rows of x [B,T,D])))). */
int vmc_student_heads(const float* emb, const float* w_fc1_t, const float* b_fc1, const float* w_fc2_t,
                      const float* b_fc2, float alpha, const float* w_c1_t, const float* b_c1, const float* w_c2_t,
                      const float* b_c2, float* distill, float* logits, int B, int T, int D, int H, int C, void* stream);
int vmc_tfam_head(const float* x, const float* ln_g, const float* ln_b, float eps, const float* w1_t, const float* b1,
                  const float* w2_t, const float* b2, float* logits, int B, int T, int D, int H, int C, void* stream);

/* ---- TFAM block as ONE fused kernel (north star piece 4) ------------------------------------------
 * Replaces the whole of AMO_CLIP.forward after the mode selection: the AttentionLayer stack (TFAM/models/AMO_CLIP.py:37-51,
 * :146-150 -- self-attention, optional cross-attention, FFN, post-LN) and `classifier(x.mean(dim=1))` (:170).  One cluster
 * of 8 CTAs (one per head) per clip, or per pair of clips in large batches; weights streamed from L2 in a host-packed
 * fragment-major fp16 order; activations stay in (distributed) shared memory; fp32 accumulation, residual stream,
 * LayerNorm and softmax.  Geometry: d_model 512, nhead 8, dim_feedforward 2048 (every configuration of the reference,
 * TFAM/cfg_AK), at most 32 frames per clip and stream; vmc_tfam_fused_supported() tells, other shapes take the batched
 * GEMM path of the host code.
 * wstream: vmc_tfam_wstream_bytes(layers) bytes = [8 CTAs][8 warps][layers][256 blocks][32 lanes][8 fp16].  Per (CTA c,
 * warp w, layer) the blocks are, in order, (k-pair p major, column tile i minor), for
 *   self in_proj   (16 k-pairs x 3 tiles): rows {0, 512, 1024} + 64 c + 8 w + [0, 8) of self_attn.in_proj_weight
 *   self out_proj  (16 x 1): rows 64 c + 8 w + [0, 8) of self_attn.out_proj.weight
 *   cross q        (16 x 1): rows 64 c + 8 w + [0, 8) of cross_attn.in_proj_weight
 *   cross k | v    (16 x 2): rows {512, 1024} + 64 c + 8 w + [0, 8) of cross_attn.in_proj_weight
 *   cross out_proj (16 x 1): rows 64 c + 8 w + [0, 8) of cross_attn.out_proj.weight
 *   ffn.0          (16 x 4): rows 256 c + 32 w + 8 i + [0, 8) of ffn.0.weight
 *   ffn.3          ( 8 x 8): rows 64 w + 8 i + [0, 8), columns 256 c + [0, 256) of ffn.3.weight
 * and block (p, i) holds for lane (g = lane / 4, t = lane % 4) the 8 values W_i[g][32 p + {2t, 2t+1, 2t+8, 2t+9, 16+2t,
 * 17+2t, 24+2t, 25+2t}] (the mma.m16n8k16 B fragments of two k steps).  vimoclip_b200.tfam packs it. */
#define VMC_TFAM_MAX_LAYERS 8
typedef struct vmc_tfam_layer {
  const float *b_sin, *b_sout, *b_cin, *b_cout, *b_f1, *b_f2; /* in_proj_bias [1536], out_proj.bias [512] (self, cross), ffn biases */
  const float *ns_g, *ns_b, *nc_g, *nc_b, *nf_g, *nf_b;        /* norm_self / norm_cross / norm_ffn weight, bias */
  float ns_eps, nc_eps, nf_eps;
} vmc_tfam_layer;
typedef struct vmc_tfam_model {
  int d_model, nhead, dim_ff, layers, num_classes, hidden; /* hidden = classifier.1 out_features (d_model / 2) */
  int act;                        /* FFN activation: VMC_ACT_RELU (reference default) or VMC_ACT_GELU_ERF */
  const void* wstream;            /* device, fp16, layout above */
  const vmc_tfam_layer* layer;    /* HOST array of `layers` entries */
  const float *cls_ln_g, *cls_ln_b; /* classifier.0 */
  float cls_ln_eps;
  const float *w1t, *b1;          /* classifier.1: weight TRANSPOSED [512, hidden] fp32, bias */
  const float *w2t, *b2;          /* classifier.4: weight TRANSPOSED [hidden, C] fp32, bias */
} vmc_tfam_model;
long long vmc_tfam_wstream_bytes(int layers);
int vmc_tfam_fused_supported(const vmc_tfam_model* m, int B, int T, int Tm);
/* x [B, T, 512] fp32 (layer-0 input: rgb / flow / concatenated / projected rows per AMO_CLIP.py:136-167), motion [B, Tm, 512]
 * fp32 = cross-attention source or NULL (self-attention-only modes), valid_x [B, T] / valid_m [B, Tm] uint8 (1 = real frame)
 * or NULL, logits [B, C] fp32.  ONE kernel launch. */
int vmc_tfam_forward(const vmc_tfam_model* m, const float* x, const float* motion, const uint8_t* valid_x,
                     const uint8_t* valid_m, float* logits, int B, int T, int Tm, void* stream);

/* ---- TFAM training step: backward kernels -------------------------------
 * TFAM/train_and_eval.py:66-101 (`loss.backward()` through TFAM/models/AMO_CLIP.py:37-51,99-171).  Every nn.Linear
 * backward is two more vmc_gemm_bf16 calls on split-bf16 operands (dX = dY W, dW = dY^T X) fed by
 * vmc_transpose_split; the rest are fp32 kernels.  All gradients are fp32. */
/* x fp32 [R, C] -> y bf16 [C, 3R] (ldy >= 3R, multiple of 8): split operand of x^T; form 0 = [hi | lo | hi]
 * (A side of vmc_gemm_bf16), form 1 = [hi | hi | lo] (W side). */
int vmc_transpose_split(const float* x, long long ldx, void* y, long long ldy, int R, int C, int form,
                        void* stream);
/* plain transposing cast x [R, C] (fp32, or bf16 when src_bf16) -> y bf16 [C, R] (ldy >= R, multiple of 8): operands of the bf16
 * dW GEMMs of the student's backward (train.py:95-107); vmc_cast_f32: bf16 [rows, d] -> fp32 */
int vmc_transpose_cast(const void* x, int src_bf16, long long ldx, void* y, long long ldy, int R, int C, void* stream);
int vmc_cast_f32(const void* x, long long ldx, float* y, long long ldy, int rows, int d, void* stream);
/* out[c] (+)= sum_r x[r,c] * (y ? y[r,c] : 1): bias / LayerNorm parameter gradients; deterministic order.
 * workspace: vmc_colsum_slices(R, C) * C floats (row slices summed in parallel, then added in index order), or NULL for
 * the single-stage kernel (one block per 32 columns: slow for tall matrices). */
int vmc_colsum_slices(int R, int C);
int vmc_colsum(const float* x, long long ldx, const float* y, long long ldy, float* out, int R, int C,
               int accumulate, float* workspace, void* stream);
/* LayerNorm backward per row from the saved LayerNorm INPUT z: dz, and xhat (optional) for dgamma = colsum(dy*xhat) */
int vmc_layernorm_bwd(const float* z, long long ldz, const float* gamma, float eps, const float* dy,
                      long long lddy, float* dz, long long lddz, float* xhat, long long ldxh, int rows, int d,
                      void* stream);
/* mode 0: out = a*b (dropout mask), 1: ReLU backward (b = activation output or pre-activation), 2: GELU(erf)
 * backward (b = pre-activation), 3: out = a+b, 4: out = a*scale, 5: QuickGELU forward a*sigmoid(1.702a), 6: QuickGELU
 * backward (b = pre-activation), 7: out = scale*a + b, 8: GELU(erf) forward of a */
int vmc_eltwise(int mode, const float* a, const float* b, float scale, float* out, long long n, void* stream);
/* y = bf16(QuickGELU(x)) over n contiguous fp32 elements (n % 4 == 0): the c_fc activation of the student's training forward */
int vmc_qgelu_cast(const float* x, void* y, long long n, void* stream);
/* Fused LayerNorm backward of the residual stream (d = 512 / 768 / 1024): dz = LN'(z; gamma)[dy] (+ add, the gradient arriving over
 * the residual branch), dgamma_dbeta[0..d) = sum_r dy * xhat, [d..2d) = sum_r dy.  workspace: vmc_layernorm_bwd_fused_blocks(rows) * 2 * d
 * floats (per-block partials, added in block order: deterministic). */
int vmc_layernorm_bwd_fused_blocks(int rows);
int vmc_layernorm_bwd_fused(const float* z, long long ldz, const float* gamma, float eps, const float* dy, long long lddy,
                            const float* add, long long ldadd, float* dz, long long lddz, float* dgamma_dbeta, float* workspace,
                            int rows, int d, void* stream);
/* Fused pass of the bf16 Linear backward (train.py:104 loss.backward(), the nn.Linear grad of dY): y16 = bf16(v), colsum[c] = sum_r v[r,c]
 * with v = x, or v = x * QuickGELU'(aux) when aux (the c_fc pre-activation) is given -- replaces an element-wise backward, a cast and a
 * column sum (three passes over the fp32 gradient) by one.  workspace: vmc_cast_colsum_slices(R, C) * C floats. */
int vmc_cast_colsum_slices(int R, int C);
int vmc_cast_colsum(const float* x, long long ldx, const float* aux, long long ldaux, void* y16, long long ldy, float* colsum,
                    int R, int C, float* workspace, void* stream);
/* ViT self-attention backward for towers of at most 64 tokens (ViT-B/32), straight from the packed bf16 qkv buffer
 * [F*L, 3*heads*64]: dqkv fp32 [F*L, 3*heads*64] from dO fp32 [F*L, heads*64] (row stride lddo) */
int vmc_attention_vit_bwd_short(const void* qkv, const float* dO, long long lddo, float* dqkv, int F, int L, int heads,
                                void* stream);
/* out[b,t,:] = g[b,:] * scale (backward of the temporal mean, AMO_CLIP.py:170) */
int vmc_broadcast_rows(const float* g, float* out, int B, int T, int d, float scale, void* stream);
/* vmc_attention_masked with dropout on the attention probabilities: prob_mask fp32 [B, heads, Tq, Tk] holds 0 or
 * 1/(1-p) (NULL = no dropout), as nn.MultiheadAttention(dropout=p) does in training mode */
int vmc_attention_masked_train(const float* q, long long ldq, const float* k, long long ldk, const float* v,
                               long long ldv, const uint8_t* key_valid, const float* prob_mask, void* out,
                               int out_f32, long long ldo, int B, int Tq, int Tk, int heads, void* stream);
/* backward of the above: dq [B*Tq, .], dk, dv [B*Tk, .] written at column head*64.  workspace NULL: one CTA per (clip,
 * head) with the probabilities in shared memory (Tq, Tk <= ~128); workspace = 2*B*heads*Tq floats: tiled two-pass kernels
 * (log-sum-exp + dQ, then dK / dV) for any Tq, Tk. */
int vmc_attention_masked_bwd(const float* q, long long ldq, const float* k, long long ldk, const float* v,
                             long long ldv, const uint8_t* key_valid, const float* prob_mask, const float* dO,
                             long long lddo, float* dq, long long lddq, float* dk, long long lddk, float* dv,
                             long long lddv, int B, int Tq, int Tk, int heads, float* workspace, void* stream);

/* ---- whole ViT tower -----------------------------------------------------------
 * Replaces self.visual_encoder(x) (models/student_model.py:84) and
 * clip_model.get_image_features(pixel_values) (extract_embeddings.py:94).
 * All weights are device pointers packed by the host side (bf16 [N,K] row-major for GEMM
 * operands, fp32 for LN / bias / embeddings).
 */
typedef struct vmc_vit_layer {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void* w_qkv;  const float* b_qkv;   /* [3d, d] bf16, [3d] */
  const void* w_out;  const float* b_out;   /* [d, d], [d] */
  const void* w_fc1;  const float* b_fc1;   /* [4d, d], [4d] */
  const void* w_fc2;  const float* b_fc2;   /* [d, 4d], [d] */
  /* LayerNorm folding (VMC_OPT_LN_FUSE = 3 / 5): ln_1 folded into the qkv GEMM and
   * ln_2 into c_fc.  w_*_f = bf16(gamma (.) W), b_*_f = b + W beta, cs_* [n] = sum_k float(w_*_f[n,k]).  NULL = absent. */
  const void* w_qkv_f; const float* b_qkv_f; const float* cs_qkv;
  const void* w_fc1_f; const float* b_fc1_f; const float* cs_fc1;
} vmc_vit_layer;
typedef struct vmc_vit_model {
  int image, patch, width, layers, heads, out_dim;
  int ld_patch;             /* row stride of the patchified A operand / conv weight (>= 3*p*p, %8) */
  const void* w_patch;      /* [width, ld_patch] bf16 = conv1.weight.reshape(width, 3*p*p) */
  const float* cls_pos0;    /* [width] = class_embedding + positional_embedding[0] */
  const float* pos;         /* [L, width] fp32 positional_embedding */
  const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
  const void* w_proj;       /* [out_dim, width] bf16 = proj.T */
  const vmc_vit_layer* layer; /* HOST array of `layers` entries */
  /* Per-model variant selectors (0 = process default, see vmc_set_option).
   * ln_mode: 6 = bf16 residual stream with both LayerNorms folded into the consuming GEMMs (the default), 3 = fp32 residual
   *   stream + bf16 copy with both LayerNorms folded, 5 = fp32 stream with ln_1 folded, 4 = fp32 stream with separate
   *   LayerNorm kernels.  last_block_cls: 1 = the last block computes only what the output reads, the CLS row (default:
   *   identical embeddings), 2 = full last block.  attn_impl: see vmc_attention_vit_impl (0 = by sequence length). */
  int ln_mode, last_block_cls, attn_impl;
} vmc_vit_model;
/* bytes of workspace needed for F frames in flight */
long long vmc_vit_workspace_bytes(const vmc_vit_model* m, int F);
/* patches: bf16 [F*n, ld_patch] from vmc_prologue / vmc_frame_diff_prologue; out fp32 [F, out_dim] */
int vmc_vit_forward(const vmc_vit_model* m, const void* patches, float* out, int F, void* workspace,
                    long long workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIMOCLIP_B200_H_ */
