/* csam.h — C-ABI of libcsam_sm100.so: hand-written sm_100a kernels for the Crowd-SAM
 * inference hot path.
 *
 * The reference (FelixCaae/CrowdSAM) has NO plugin/FFI interface: it is pure PyTorch
 * (SURVEY.md §8b).  Each entry point below therefore replaces a chain of ATen calls in the
 * reference; the chain is cited as file:line relative to the reference root.  The Python
 * binding a reference maintainer would add is a ctypes stub, shown in INTEGRATION.md and
 * implemented in crowdsam_b200/lib.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the library allocates nothing the
 *     caller must free (scratch is passed in);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, non-zero = error, text via csam_last_error() (thread-local);
 *   - "h16 pair": an activation/weight stored as fp16 `hi` plus optional fp16 `lo` with
 *     value = hi + lo (error-compensated split; lo == NULL selects single-pass fp16).
 *     With both operands split a GEMM issues 3 tensor-core MMAs (hi*hi + lo*hi + hi*lo) and is
 *     fp32-accurate to ~1e-6 relative; this is what keeps the path inside the 1e-3 parity bound;
 *   - no CPU fallback exists anywhere in this library.
 */
#ifndef CSAM_H_
#define CSAM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSAM_ABI_VERSION 10
#if defined(__GNUC__)
#define CSAM_API __attribute__((visibility("default")))
#else
#define CSAM_API
#endif

CSAM_API const char* csam_last_error(void);
CSAM_API int csam_abi_version(void);
/* number of kernel launches issued by this library since load (for bench.py "gpu_launches") */
CSAM_API long long csam_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K-GEMM  out = epilogue(A[M,K] * W[N,K]^T)          (tcgen05 + TMA, persistent, warp-specialised)
 * replaces nn.Linear / 1x1 conv / ConvTranspose2d(k2,s2) / patch-embed conv call sites:
 *   image_encoder.py:212-213,235-239 (qkv, proj), common.py:21-26 (MLP), :88-104 (neck),
 *   :387-395 (patch embed); dinov2/layers/attention.py:58,66, layers/mlp.py:34-40;
 *   transformer.py:218-221,230-232,251 (decoder projections), mask_decoder.py:56-62 (ConvT),
 *   :72-74,175-198 (heads, dino_proj).
 * epilogue: v = acc * row_scale[r] + bias[c]; v = act(v); v = v * col_scale[c] (LayerScale);
 *           v += residual[rr*ldr + c]  (rr = out row, or out row % res_mod when res_mod > 0);
 *           out row = row_map ? row_map[r] : r  (negative = row dropped: window un-partition).
 * ------------------------------------------------------------------------------------------ */
enum { CSAM_ACT_NONE = 0, CSAM_ACT_GELU = 1, CSAM_ACT_RELU = 2 };
/* TCGEN05 = tensor-core kernels, tile shape / CTA-pair mode chosen by the library (big split-operand problems run on
 * CTA pairs with 256x128 pair tiles, the rest on the single-CTA kernel); SIMT = slow validation kernel;
 * TC_PAIR / TC_PAIR128 = force the 2-CTA (tcgen05 cta_group::2) kernel with 256x256 / 256x128 pair tiles, error if the
 * problem does not qualify; TC_SINGLE = force the single-CTA kernel (tests, A/B measurements) */
enum { CSAM_GEMM_TCGEN05 = 0, CSAM_GEMM_SIMT = 1, CSAM_GEMM_TC_PAIR = 2, CSAM_GEMM_TC_SINGLE = 3, CSAM_GEMM_TC_PAIR128 = 4 };
/* fused epilogues of the mask decoder (tcgen05 only):
 *  CSAM_EPI_LN   N == 256: y = LayerNorm(acc + bias + residual) * gamma + beta over the whole row;
 *                outputs (each optional): out_f32 = y, h16 pair (out_hi/lo) = y, h16 pair (out2_hi/lo) =
 *                y + pe[row % pe_mod].  transformer.py:184-190 (image->token out_proj + norm4).
 *  CSAM_EPI_UP1  N == 256 (col = (dy*2+dx)*64 + c), M = P*4096: ConvTranspose2d#1 + LayerNorm2d(64) + GELU,
 *                written pixel-shuffled as h16 pair [P*16384, 64].  mask_decoder.py:56-60.
 *  CSAM_EPI_UP2  N == 128 (col = (dy*2+dx)*32 + c), M = P*16384: ConvTranspose2d#2 + GELU + dot with
 *                hyper[P,4,32] -> masks fp32 [P,4,256,256].  mask_decoder.py:61-62,175-181. */
enum { CSAM_EPI_STD = 0, CSAM_EPI_LN = 1, CSAM_EPI_UP1 = 2, CSAM_EPI_UP2 = 3 };

typedef struct {
  const void* a_hi; const void* a_lo;     /* [M,K] fp16, row stride lda (elements, multiple of 8) */
  const void* w_hi; const void* w_lo;     /* [N,K] fp16, row stride ldw */
  int M, N, K, lda, ldw;
  const float* bias;                      /* [N] or NULL */
  const float* row_scale;                 /* [M] or NULL */
  const float* col_scale;                 /* [N] or NULL */
  int act;
  const float* residual; int ldr; int res_mod;
  const int* row_map;                     /* [M] or NULL */
  float* out_f32; int ldo;                /* optional fp32 output */
  void* out_hi; void* out_lo; int ldh;    /* optional h16-pair output */
  int impl;                               /* CSAM_GEMM_TCGEN05 / _SIMT / _TC_PAIR / _TC_SINGLE / _TC_PAIR128 */
  int b_mn_major;                         /* 1: W given as [K,N] row-major (ldw = N stride) */
  int epi;                                /* CSAM_EPI_* */
  const float* gamma; const float* beta; float eps;          /* EPI_LN / EPI_UP1 */
  const float* pe; int ldpe; int pe_mod; void* out2_hi; void* out2_lo;   /* EPI_LN second output */
  const float* hyper; float* masks;                          /* EPI_UP2 */
  const void* res_hi; const void* res_lo; int ldrh;          /* EPI_LN: residual as an h16 pair (instead of fp32) */
} csam_gemm_args;
CSAM_API int csam_gemm(const csam_gemm_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Image -> patch matrix.  sam.py:163-173 (normalise + zero pad) fused with the im2col of the
 * 16x16/s16 patch-embed conv (image_encoder.py:391-395), and, for DINOv2, with the bilinear
 * 1024->1022 resize (predictor.py:104) and the 14x14/s14 im2col (dinov2 layers/patch_embed.py).
 * img: uint8 [3,h,w] planar, or [h,w,3] interleaved when hwc != 0 (after ResizeLongestSide);
 * out: h16 pair [n_side*n_side, kpad],
 * column = c*patch*patch + py*patch + px (conv weight order), zero beyond 3*patch*patch.
 * ------------------------------------------------------------------------------------------ */
CSAM_API int csam_patchify(const uint8_t* img, int h, int w, int hwc, int patch, int n_side, int resize_to,
                  void* out_hi, void* out_lo, int kpad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row LayerNorm (+ fused residual add, gather, cast).  y = LN(x[src] + add[src % add_mod]) * g + b
 * replaces nn.LayerNorm / LayerNorm2d (channels-last rows): image_encoder.py:167,180,88-104,
 * common.py:38-43, transformer.py:166-191, dinov2 block.py:92-95, vision_transformer.py:261.
 * row_map (optional, [rows_out]): source row per output row, negative -> output zeros
 * (window partition with zero padding AFTER norm1, image_encoder.py:168-172,256-264).
 * normalize = 0 turns it into a pure cast/gather.  Outputs (each optional): fp32, h16 pair,
 * and a second h16 pair of y + pe[row % pe_mod] (keys + key_pe for the decoder).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* x; int ldx; int rows_in;
  const float* add; int ldadd; int add_mod;
  const int* row_map; int rows_out; int cols;
  const float* gamma; const float* beta; float eps; int normalize;
  float* out_f32; int ldo;
  void* out_hi; void* out_lo; int ldh;
  const float* pe; int ldpe; int pe_mod; void* out2_hi; void* out2_lo;
  int act;                                /* activation applied after the affine (GELU for upscaling) */
} csam_ln_args;
CSAM_API int csam_layernorm(const csam_ln_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * ViT attention (windowed or global) with decomposed relative-position bias.
 * replaces image_encoder.py:224-240,292-361 and dinov2/layers/attention.py:56-69.
 * qkv: h16 pair [groups*tokens, 3*heads*hd]  (q | k | v blocks of heads*hd columns).
 * rel_h/rel_w: fp32 [2S-1, hd] tables or NULL (DINOv2).  S*S == tokens when given.
 * bias uses the UNSCALED q (image_encoder.py:231-234); scores use q*scale.
 * out: h16 pair [groups*tokens, heads*hd].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* qkv_hi; const void* qkv_lo; int ld_qkv;
  int groups, tokens, heads, hd; float scale;
  const float* rel_h; const float* rel_w; int S;
  void* out_hi; void* out_lo; int ld_out;
  float* scratch; long long scratch_bytes;   /* >= csam_vit_attention_scratch_bytes() */
  int impl;
  int p_split;   /* tcgen05 path: 1 = softmax probabilities also as hi+lo pair (3 MMAs for P*V); 0 = P as one
                    fp16 (2 MMAs, relative error 2^-12 per probability); -1 = P and V each as one fp16 (1 MMA, no
                    V_lo load; 2^-12 per probability and per value) */
} csam_attn_args;
CSAM_API long long csam_vit_attention_scratch_bytes(int groups, int tokens, int heads, int hd, int S);
CSAM_API int csam_vit_attention(const csam_attn_args* a, void* stream);

/* neck 3x3 conv im2col (image_encoder.py:96-102): x h16 pair [64*64, C] channels-last ->
 * [64*64, 9*C], column = (ky*3+kx)*C + c, zero padding 1. */
CSAM_API int csam_im2col3x3(const void* x_hi, const void* x_lo, int side, int C,
                   void* out_hi, void* out_lo, void* stream);

/* fp32 [rows, cols] -> fp32 [cols, rows] (features to NCHW, sam.py / predictor.py:101) */
CSAM_API int csam_transpose_f32(const float* in, int rows, int cols, float* out, void* stream);

/* Generic bilinear resize, align_corners=False (ATen upsample_bilinear2d semantics), planes
 * layout [n, hin, win] -> [n, hout, wout] fp32; or channels-last when chlast != 0
 * ([hin,win,n] -> [hout,wout,n]).  predictor.py:104,120; mask_decoder.py:188; model.py:202. */
CSAM_API int csam_bilinear(const float* in, int n, int hin, int win, float* out, int hout, int wout,
                  int chlast, void* stream);

/* ------------------------------------------------------------------------------------------
 * Prompt tokens: prompt_encoder.py:75-93,189-218 + mask_decoder.py:153-155.
 * coords01: fp32 [P,2] = (point + 0.5)/1024 computed in fp64 on the host then cast (the
 * reference does that arithmetic in float64).  tokens fp32 [P,7,256] =
 * [iou_token, mask_tokens(4), PE(point)+label_embed, not_a_point_embed].
 * ------------------------------------------------------------------------------------------ */
CSAM_API int csam_prompt_tokens(const float* coords01, const int* labels, int P, const float* gauss /*[2,128]*/,
                       const float* out_tokens5 /*[5,256]*/, const float* point_emb /*[2,256] neg,pos*/,
                       const float* not_a_point /*[256]*/, float* tokens, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder attentions (transformer.py:228-254 after the q/k/v projections), fp32 math.
 *  few_keys : every query attends to nk <= 8 keys   (image->token, token self-attention)
 *  few_queries: nq <= 8 queries attend to nk keys     (token->image)
 * q [Bq,nq,C], k/v [Bk,nk,C] fp32 with C = heads*hd; B* == 1 broadcasts over the batch.
 * out: fp32 and/or h16 pair [B,nq,C].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* q; int Bq; const float* k; const float* v; int Bk;
  int B, nq, nk, heads, hd;
  float* out_f32; void* out_hi; void* out_lo;
  int ldq, ldk, ldv;   /* row strides in floats (multiples of 4); 0 = dense rows of C floats.  Lets the attention
                          read column slices of one fused projection output [rows, (k|v|q)]. */
} csam_dec_attn_args;
CSAM_API int csam_attn_few_keys(const csam_dec_attn_args* a, void* stream);
CSAM_API int csam_attn_few_queries(const csam_dec_attn_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-I2T  the image->token half of a two-way layer as ONE kernel (transformer.py:184-190 with
 * Attention.forward :228-254):  x' = LayerNorm(x + out_proj(softmax(q_proj(x + pe) k_t^T / 4) v_t)).
 * The 7 prompt tokens are folded into per-prompt operands first (csam_dec_fold_i2t):
 *   kt, vt: fp32 [P,7,128] = k_proj(tokens + pe), v_proj(tokens) of the layer's image->token attention;
 *   wq fp32 [128,256] (q_proj.weight), wo fp32 [256,128] (out_proj.weight);
 *   b1: h16 pair [P*64, 384], row h*8+j: (log2e/4) * (Wq_h^T kt[j,h] | block-diagonal kt[j,h]);
 *   b2: h16 pair [P*256, 64], column h*8+j: Wo_h vt[j,h].
 * csam_dec_i2t_layer: x h16 pair [P*4096,256] (or [4096,256] when x_shared: layer 0, where every prompt
 * starts from the same image embedding), peq h16 pair [4096,128] = pe Wq^T + b_q, bias = out_proj.bias,
 * gamma/beta/eps = norm4 -> out h16 pair [P*4096,256].  Reads x once and writes x' once; the q and
 * attention-output streams of the unfused path never exist.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x_hi; const void* x_lo; int x_shared;
  const void* peq_hi; const void* peq_lo;
  const void* b1_hi; const void* b1_lo;
  const void* b2_hi; const void* b2_lo;
  int P;
  const float* bias; const float* gamma; const float* beta; float eps;
  void* out_hi; void* out_lo;
} csam_i2t_layer_args;
CSAM_API int csam_dec_fold_i2t(const float* kt, const float* vt, int P, const float* wq, const float* wo,
                      const float* bo /* out_proj.bias or NULL: folded into b2 (then pass bias = NULL to the layer) */,
                      void* b1_hi, void* b1_lo, void* b2_hi, void* b2_lo, void* stream);
CSAM_API int csam_dec_i2t_layer(const csam_i2t_layer_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-T2I  token->image cross attention with k_proj / v_proj folded away (transformer.py:171-176 and the final
 * attention :104-112, Attention.forward :228-254): the 7 tokens of a prompt attend over its 4096 image tokens.
 *   csam_dec_fold_t2i: qt fp32 [P,7,128] = q_proj(tokens + pe), wk fp32 [128,256] (k_proj.weight)
 *                      -> b1 h16 pair [P*64, 384] (same layout as csam_dec_fold_i2t's B1);
 *   csam_dec_t2i: x h16 pair [P*4096,256] (or [4096,256] when x_shared), pek h16 pair [4096,128] = pe Wk^T + b_k,
 *     xbar fp32 scratch [P,64,256] (softmax-pooled keys per head and token), wv_t fp32 [256,128] = v_proj.weight^T,
 *     bv = v_proj.bias -> out fp32 and/or h16 pair [P,7,128] = the attention output BEFORE out_proj.
 * Reads the keys once; the k | v projection stream of the unfused path never exists.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x_hi; const void* x_lo; int x_shared;
  const void* pek_hi; const void* pek_lo;
  const void* b1_hi; const void* b1_lo;
  int P;
  float* xbar;
  const float* wv_t; const float* bv;
  float* out_f32; void* out_hi; void* out_lo;
} csam_t2i_args;
CSAM_API int csam_dec_fold_t2i(const float* qt, int P, const float* wk, void* b1_hi, void* b1_lo, void* stream);
CSAM_API int csam_dec_t2i(const csam_t2i_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask upscaling tail (mask_decoder.py:56-62,173-181).
 * shuffle_ln_gelu: ConvT1 GEMM output [P*4096, 4*64] (col = (dy*2+dx)*64 + c) -> LN2d over the
 *   64 channels + GELU -> h16 pair [P*16384, 64] with row = p*16384 + (2y+dy)*128 + (2x+dx).
 * hyper_masks: ConvT2 GEMM output [P*16384, 4*32] (+bias applied) -> GELU -> dot with
 *   hyper_in[P,4,32] -> masks fp32 [P,4,256,256].
 * ------------------------------------------------------------------------------------------ */
CSAM_API int csam_upscale_shuffle_ln_gelu(const float* y1, int P, const float* gamma, const float* beta, float eps,
                                 void* out_hi, void* out_lo, void* stream);
CSAM_API int csam_upscale_hyper_masks(const float* y2, int P, const float* hyper_in, float* masks, void* stream);

/* ------------------------------------------------------------------------------------------
 * PWD-Net pooling weights (mask_decoder.py:189): per row of masks [R, n] (n = 65536):
 * e = exp(x - max) * 2^14 as h16 pair, inv_sum[r] = 1 / (2^14 * sum exp(x - max)), so that
 * pooled = (E * dmap^T) * inv_sum is the softmax-weighted average (GEMM row_scale).
 * ------------------------------------------------------------------------------------------ */
CSAM_API int csam_softmax_weights(const float* x, int R, int n, void* e_hi, void* e_lo, float* inv_sum, void* stream);

/* PWD score + candidate selection (model.py:351-358,318-331):
 * score = clamp(iou,0) * sigmoid(cls); sel = argmax over the 4 candidates (first max);
 * mode 0 = max_iou.  cat = argmax over classes of the selected candidate. */
CSAM_API int csam_select_candidates(const float* iou /*[P,4]*/, const float* cls /*[P,4,ncls]*/, int P, int ncls,
                           float* score /*[P]*/, int* sel /*[P]*/, int* cat /*[P]*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-POST  (sam.py:132-161 + model.py:372-384 + amg.py:156-176,303-346), HBM-bound.
 * low: fp32 [P,4,256,256]; sel[P] picks the candidate plane.  The two bilinear resizes
 * (256 -> 1024, crop to (in_h,in_w), -> (out_h,out_w)) are evaluated on the fly.
 *  stats : per prompt  counts[p] = {#(m > thr+off), #(m > thr-off), #(m > thr)}, box[p] =
 *          inclusive XYXY of (m > thr) or zeros; nothing else is written.
 *  write : for each i < n_keep: masks[i] = (m[keep[i]] > thr) as uint8 [out_h,out_w]; an entry keep[i] < 0 leaves
 *          slot i untouched (a keep list compacted on the device without a host round trip ends in -1s);
 *          optionally logits fp32 [n_keep,out_h,out_w].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* low; int P; const int* sel;      /* sel NULL: low is [P,256,256] */
  int planes;                                   /* 4, or 1 when sel == NULL */
  int in_h, in_w, out_h, out_w;
  float thr, off;
  int* counts;  /* [P,3] */
  int* boxes;   /* [P,4] */
  const int* keep; int n_keep;
  uint8_t* masks; float* logits;
} csam_post_args;
CSAM_API int csam_mask_post_stats(const csam_post_args* a, void* stream);
CSAM_API int csam_mask_post_write(const csam_post_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-NMS  torchvision.ops.nms semantics (model.py:171,257,429): stable descending score order,
 * suppress when inter/(a+b-inter) > thr in fp32, areas without +1, NaN never suppresses.
 * keep_out[0..*n_keep) = kept original indices in stable descending-score order (bit-exact).
 * scratch >= csam_box_nms_scratch_bytes(n).
 * ------------------------------------------------------------------------------------------ */
CSAM_API long long csam_box_nms_scratch_bytes(int n);
CSAM_API int csam_box_nms(const float* boxes /*[n,4]*/, const float* scores /*[n]*/, int n, float thr,
                 int* keep_out /*[n]*/, int* n_keep /*[1]*/, void* scratch, long long scratch_bytes,
                 void* stream);

/* K-MIOU (crowdsam/utils.py:422-479, dead code in the reference): pairwise mask overlap on
 * nearest-resized 150x150 bitmaps. inter fp32?  -> int32 [n,n] intersections and area [n]. */
CSAM_API int csam_mask_overlap(const uint8_t* masks, int n, int h, int w, int* inter /*[n,n]*/, int* area /*[n]*/,
                      void* scratch, long long scratch_bytes, void* stream);
CSAM_API long long csam_mask_overlap_scratch_bytes(int n);

/* EPS occupancy test (model.py:229-246): occ[i] = OR_j masks[sel_j][py_i,px_i] for the masks
 * with flag[j] != 0. */
CSAM_API int csam_points_occupied(const uint8_t* masks, int n_masks, int h, int w, const uint8_t* flag,
                         const int* pts_xy, int n_pts, uint8_t* occ, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-CC  remove_small_regions on the device (amg.py:267-291 = cv2.connectedComponentsWithStats(working, 8) + area
 * filter, called per mask on the host from crowdsam/model.py:395-443).  masks uint8 [n,h,w] (0/1), edited in place.
 *   mode 0 "holes":   background components smaller than area_thresh are filled;
 *   mode 1 "islands": foreground components smaller than area_thresh are dropped; if all are smaller the largest
 *                     is kept (ties: the component OpenCV labels first).
 * changed[i] = 1 when mask i had any component below the threshold (the reference's second return value).
 * Integer-exact against OpenCV.  scratch >= csam_small_regions_scratch_bytes(n,h,w).
 * ------------------------------------------------------------------------------------------ */
CSAM_API long long csam_small_regions_scratch_bytes(int n, int h, int w);
CSAM_API int csam_remove_small_regions(uint8_t* masks, int n, int h, int w, int area_thresh, int mode, uint8_t* changed,
                              void* scratch, long long scratch_bytes, void* stream);

/* Column-major run-length encoding of bool masks (amg.py:107-135), two passes with a host decision in between (the
 * caller allocates `runs` from the counts):
 * count: n_runs[i] = number of runs of mask i (the first run counts zeros; 0 if the mask starts with 1); leaves the
 *        per-(column, row segment) change offsets of every mask in scratch (>= csam_rle_scratch_bytes(n,h,w)).
 * fill : runs[offsets[i] .. offsets[i]+n_runs[i]) = the run lengths (offsets = exclusive prefix sum of n_runs, computed
 *        by the caller; max_runs = max n_runs[i]); `pos` = int32 scratch of the same size as `runs` (change positions),
 *        `scratch` and `n_runs` as left by csam_rle_count on the same masks.
 * A mask is split into w * 8 work items, so a single kept mask occupies the whole GPU. */
CSAM_API long long csam_rle_scratch_bytes(int n, int h, int w);
CSAM_API int csam_rle_count(const uint8_t* masks, int n, int h, int w, int* n_runs, void* scratch, long long scratch_bytes,
                            void* stream);
CSAM_API int csam_rle_fill(const uint8_t* masks, int n, int h, int w, const long long* offsets, const int* n_runs,
                           int max_runs, int* pos, int* runs, const void* scratch, void* stream);

/* HOST function (no device work): COCO compressed RLE strings of n_masks run-length lists, replacing the per-mask
 * `coco_encode_rle` -> pycocotools `frPyObjects` call of amg.py:294-300 / model.py:184-185 (pycocotools' rleToString:
 * counts beyond the third are delta coded against counts[i-2]; 5-bit groups, little endian, bit 5 = continuation,
 * + 48).  counts = all run lengths concatenated (host pointer, what csam_rle_fill produced, read back), offsets[n_masks+1]
 * = start of each mask's runs; out receives the strings back to back (capacity cap bytes; 7 bytes per run always
 * suffice), out_offsets[n_masks+1] their starts.  Returns 0, or 1 when cap is too small. */
CSAM_API int csam_coco_rle_strings(const int* counts, const long long* offsets, int n_masks, char* out, long long cap,
                                   long long* out_offsets);

#ifdef __cplusplus
}
#endif
#endif /* CSAM_H_ */
