/*
 * gd_b200.h — C ABI of libgd_b200.so, the sm_100a kernel library behind the guided-diffusion sampling path.
 *
 * The reference (ErezYosef/guided-diffusion-clip, /root/reference) has no FFI of its own: its hot path is
 * PyTorch calls.  Every entry point below therefore names the reference *call site* it replaces
 * (file:line relative to /root/reference/guided_diffusion/).  Conventions:
 *   - extern "C", plain pointers and sizes only (no torch types); device pointers unless stated.
 *   - returns 0 on success, <0 on error; gd_last_error() gives a thread-local message.  Never throws.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - no allocation, no synchronisation inside; safe to capture into a CUDA graph.
 *   - activations are NHWC fp16 "channel views": (pointer, C, ld) where ld is the per-pixel stride in
 *     elements, so a view may be a channel slice of a wider concat buffer (unet.py:661 th.cat is free).
 */
#ifndef GD_B200_H_
#define GD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD_B200_ABI_VERSION 1

/* residual modes of the conv epilogue */
enum { GD_RES_NONE = 0, GD_RES_SAME = 1, GD_RES_UPSAMPLE2 = 2, GD_RES_AVGPOOL2 = 3 };
/* output modes of the conv epilogue */
enum { GD_OUT_NHWC_F16 = 0, GD_OUT_NCHW_F32 = 1 };
/* spatial modes of the fused GroupNorm apply */
enum { GD_GN_SAME = 0, GD_GN_AVGPOOL2 = 1, GD_GN_UPSAMPLE2 = 2 };
/* QKV channel order: unet.py:347 (legacy: [head][q,k,v][d]) vs unet.py:380-388 (new: [q,k,v][head][d]) */
enum { GD_QKV_LEGACY = 0, GD_QKV_NEW = 1 };

const char* gd_last_error(void);
int gd_version(void);
/* Number of kernels launched by this library on this thread since the last gd_launch_count_reset(). */
int64_t gd_launch_count(void);
void gd_launch_count_reset(void);
/* (Measurement hooks used by profiles/ and by one parity test live in include/gd_b200_devtools.h; they are compiled
 * into the library only with -DGD_B200_DEVTOOLS and are not part of the product ABI.) */

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 (3x3 pad 1 stride 1, or 1x1), fp16 operands, fp32 accumulate.
 * Replaces nn.Conv2d / nn.Conv1d(k=1): unet.py:185,211 (ResBlock convs), :222 (1x1 skip), :286,294
 * (qkv / proj_out), :616 (out head); the first conv :483 (C_in 3 or 6) is gd_im2col3x3_small_cin followed by a
 * K=64 GEMM through this entry point.  With gd_pack-ed flipped weights the same entry point is the conv backward-data of the guidance gradient
 * (scripts/classifier_sample.py:54-61 autograd through unet.py:872-895).
 *   out[n,y,x,co] = out_scale * ( bias[co] + sum_{tap,c} a0[n,y+dy,x+dx,c] * W[co][tap*C0+c]
 *                                 + sum_c a1[n,y,x,c] * W[co][taps*C0+c]  + residual )
 * ---------------------------------------------------------------------------------------------- */
typedef struct gd_conv_desc {
  const void* a0; /* fp16 NHWC view [n,h,w,c0], per-pixel stride ld0 */
  int32_t c0, ld0, taps; /* taps: 9 (3x3) or 1 (1x1) */
  const void* a1; /* optional second source, 1x1 (fused skip_connection); NULL if unused */
  int32_t c1, ld1;
  int32_t n, h, w;
  const void* wpack; /* fp16 [n_pad][k_total], k = tap*C0 + c then C1 channels (see gd pack helpers) */
  int32_t k_total, n_pad;
  const float* bias; /* fp32 [cout] or NULL */
  int32_t cout;
  const void* res; /* fp16 NHWC view with cout channels (resolution per res_mode) or NULL */
  int32_t ld_res, res_mode;
  void* out;
  int32_t ld_out, out_mode;
  int32_t bn; /* N tile, 0 = auto */
  float out_scale; /* 0 is treated as 1 */
  /* Optional fused GroupNorm statistics of the stored output (consumed by gd_groupnorm_finalize_partials):
   * fp32 [gd_conv_stats_rows(n,h,w)][n_pad/4][2] = per (row block, 4-channel chunk) sum and sum of squares; a row block
   * is a 128-pixel tile when the tile lies inside one image, else a 32-pixel quarter of it.
   * Requires fp16 NHWC output, cout % 64 == 0 and h*w >= 32 per image; NULL = not produced. */
  float* stats_out;
  /* Optional GroupNorm32 (+FiLM) (+SiLU) applied to the MAIN operand a0 on its way into the tensor cores
   * (ResBlock in_layers / out_layers: GN -> SiLU -> conv, unet.py:184-185, 205-211, 248-252): a0 is then the RAW tensor
   * and the conv computes conv3x3(pad0(act(GN(a0)))) — the zero padding applies to the NORMALISED tensor, exactly like
   * gd_groupnorm_apply followed by this conv, bit for bit, without the normalised tensor ever touching HBM.
   * gn_silu must be 1 (every GroupNorm in front of a 3x3 conv of the reference is followed by SiLU).
   * gn_mode 0 = off; GD_CONV_GN_SAME: a0 is [n,h,w,c0]; GD_CONV_GN_UPSAMPLE2: a0 is [n,h/2,w/2,c0] and is nearest-
   * upsampled x2 after the activation (h_upd of an "up" ResBlock, unet.py:191-195).  Requires gd_conv_gn_fusable(h,w)
   * and taps == 9.  gn_coef: the affine table of gd_groupnorm_coef / gd_groupnorm_finalize_partials, fp32 [n][c0/8][16].
   * The optional 1x1 source a1 stays raw. */
  int32_t gn_mode, gn_silu;
  const float* gn_coef;
  /* Optional split-K workspace (16-byte aligned device memory the launch may scribble on; NULL = never split).
   * OPT-IN: honoured only when the process runs with GD_B200_SPLITK=1 — the number of splits follows the batch, so a
   * sample's low-order bits then depend on the batch it is computed in (every other path is batch-invariant bit for bit).
   * A 3x3 conv whose (pixel tile, N tile) work items leave at least half of the SMs idle — the 8x8 / 16x16 layers of
   * unet.py:552-609 at small per-GPU batch, K up to 18 432 — is cut along K into up to 8 splits that accumulate into
   * fp32 slabs of this workspace; a second launch sums them in a fixed order and finishes the epilogue (bias,
   * residual GD_RES_SAME, fp16 output, stats_out).  gd_conv_splitk_ws_bytes(desc) = the most this conv can use
   * (0: it never splits); a smaller workspace just means fewer splits. */
  void* splitk_ws;
  int64_t splitk_ws_bytes;
} gd_conv_desc;
enum { GD_CONV_GN_OFF = 0, GD_CONV_GN_SAME = 1, GD_CONV_GN_UPSAMPLE2 = 2 };
/* 1 if a 3x3 conv over h x w images can normalise its operand on the fly (16 x 8 pixel tiles of one image). */
int gd_conv_gn_fusable(int32_t h, int32_t w);
int gd_conv_igemm(const gd_conv_desc* desc, void* stream);
/* Upper bound of the split-K workspace gd_conv_igemm(desc) can use on the current device (pointer fields other than
 * NULL-ness of res / bias alignment are not dereferenced); 0 if this conv never splits. */
int64_t gd_conv_splitk_ws_bytes(const gd_conv_desc* desc);
/* Geometry of the fused statistics: number of row blocks the conv writes (rows of stats_out), and how many
 * consecutive rows belong to one image (rows_per_image * n == rows).  Returns 0 rows if the geometry is ineligible. */
int64_t gd_conv_stats_rows(int32_t n, int32_t h, int32_t w, int32_t* rows_per_image);
/* mean / rstd of GroupNorm32 over a tensor whose channels come from one or two conv outputs (a skip concatenation,
 * unet.py:661) from their fused partials: c0 (+ c1) channels, 32 groups, biased variance.  p1 may be NULL. */
int gd_groupnorm_finalize_partials(const float* p0, int32_t c0, int32_t ld0, const float* p1, int32_t c1, int32_t ld1,
                                   int32_t rows_per_image, int32_t n, int32_t hw, float eps, float* mean_rstd,
                                   const float* gamma, const float* beta, const float* film, int32_t film_ld,
                                   float* coef_out, void* stream);
/* Per-channel affine of GroupNorm32 (+FiLM) as a table: coef_out fp32 [n][c/8][16] = for every 8-channel chunk
 * a[8] then b[8] with  y = x*a + b,  a = rstd*gamma[*(1+scale)],  b = (beta - mean*rstd*gamma)[*(1+scale) + shift]
 * (nn.py:17-19, unet.py:248-252; film rows are (scale[c], shift[c]) with stride film_ld, or NULL).  Consumed by
 * gd_conv_desc.gn_coef.  gd_groupnorm_finalize_partials writes the same table when coef_out != NULL (gamma / beta /
 * film are only read then). */
int gd_groupnorm_coef(const float* mean_rstd, const float* gamma, const float* beta, const float* film, int32_t film_ld,
                      int32_t n, int32_t c, float* coef_out, void* stream);

/* im2col for the first layer input_blocks.0.0 (unet.py:483,741), C_in = 3 or 6: fp32 NCHW [n,cin,h,w] -> fp16 NHWC
 * [n,h,w,64] with channel k = (ky*3+kx)*cin + ci, zero padded; the conv itself then runs on gd_conv_igemm with taps=1. */
int gd_im2col3x3_small_cin(const float* x, void* out, int32_t ld_out, int32_t n, int32_t cin, int32_t h, int32_t w,
                           void* stream);

/* The same first layer in ONE launch, straight from the fp32 NCHW network input (no im2col buffer): warp-level
 * tensor-core MMAs with register accumulators -- at 27 MACs per output the layer is bound by writing its fp16 NHWC
 * output, and on the tcgen05 path by reading accumulators out of TMEM.  wpack as for gd_im2col3x3_small_cin + gd_conv_igemm
 * (fp16 [cout][64], k = (ky*3+kx)*cin + ci, zero padded).  Requires cout % 64 == 0 and h*w % 128 == 0.
 * stats_out (optional): fused GroupNorm partials in gd_conv_igemm's format, one row per 128 pixels; rows beyond
 * h*w/128 of an image's gd_conv_stats_rows block are left untouched (the caller zero-fills the buffer once). */
typedef struct {
  const float* x;      /* fp32 NCHW [n,cin,h,w] */
  const void* wpack;   /* fp16 [cout][64] */
  const float* bias;   /* fp32 [cout] */
  void* out;           /* fp16 NHWC view, pixel stride ld_out */
  float* stats_out;    /* fp32 [gd_conv_stats_rows(n,h,w)][cout/4][2] or NULL */
  int32_t n, cin, h, w, cout, ld_out;
} gd_conv_in_desc;
int gd_conv_in3x3(const gd_conv_in_desc* desc, void* stream);

/* conv_resample layers and FiLM-less ResBlocks of models built with the factory defaults resblock_updown=False /
 * use_scale_shift_norm=False-style checkpoints (script_util.py:57-60).  fp16 NHWC views, c % 8 == 0.
 *   gd_im2col3x3_s2_nhwc: Downsample.op = conv3x3 stride 2 pad 1 (unet.py:125-136) as a gather to
 *     out[n][ho][wo][tap*c + ci] = x[n][2*yo+ky-1][2*xo+kx-1][ci] (0 outside), ho = (h-1)/2+1, followed by
 *     gd_conv_igemm with taps = 1 over K = 9*c (same packed weight order as a 3x3 conv); gd_col2im3x3_s2_nhwc is its
 *     transpose for the classifier's data-gradient;
 *   gd_upsample2_nhwc: F.interpolate(scale_factor=2, mode="nearest") of Upsample.forward (unet.py:100-110);
 *   gd_add_emb_nhwc: h += emb_out[n][c] in place (unet.py:253-254), emb fp32 with row stride ld_emb. */
int gd_im2col3x3_s2_nhwc(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t n, int32_t h, int32_t w, int32_t c,
                         void* stream);
int gd_upsample2_nhwc(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t n, int32_t h, int32_t w, int32_t c,
                      void* stream);
int gd_add_emb_nhwc(void* x, int32_t ld, const float* emb, int32_t ld_emb, int32_t n, int32_t hw, int32_t c, void* stream);
/* Data-gradient of Downsample.op: the transpose of gd_im2col3x3_s2_nhwc.  dcols = dY x W (gd_conv_igemm, taps = 1, packed
 * weights transposed to [9*c][c_out]) as fp16 [n][ho][wo][9*c]; dx[n][y][x][ci] sums the 1, 2 or 4 tap columns that read
 * input pixel (y, x) in the forward pass. */
int gd_col2im3x3_s2_nhwc(const void* dcols, int32_t ld, void* dx, int32_t ld_dx, int32_t n, int32_t h, int32_t w, int32_t c,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm32 (+SiLU) (+FiLM scale/shift) (+avgpool2 / nearest-upsample2), nn.py:17-19,93-100 with
 * unet.py:184,200-208,248-252 and the h_upd of unet.py:191-195.  Two launches: statistics, then apply.
 * stats: mean_rstd[n][32][2] fp32 (biased variance, eps 1e-5).
 * apply: y = act( (x-mean)*rstd*gamma+beta ) with optional *(1+scale)+shift before act;
 *        film points at [n][2*c] fp32, scale first (unet.py:250), film_ld = row stride in floats.
 * ---------------------------------------------------------------------------------------------- */
int gd_groupnorm_stats(const void* x, int32_t ld, int32_t n, int32_t hw, int32_t c, float eps, float* partial_ws,
                       float* mean_rstd, void* stream);
int64_t gd_groupnorm_ws_floats(int32_t n, int32_t hw, int32_t c);
/* aux_out (GD_GN_AVGPOOL2 only, may be NULL): fp16 [n,h/2,w/2,c] view receiving avgpool2(x) of the RAW input, i.e. the
 * x_upd(x) residual of a down ResBlock (unet.py:195,241). */
int gd_groupnorm_apply(const void* x, int32_t ld, const float* mean_rstd, const float* gamma, const float* beta,
                       const float* film, int32_t film_ld, void* out, int32_t ld_out, int32_t n, int32_t h, int32_t w,
                       int32_t c, int32_t silu, int32_t spatial_mode, void* aux_out, int32_t ld_aux, void* stream);
/* Backward of the fused op above w.r.t. x (no parameter gradients; the guidance gradient needs dX only).
 * dy is at the OUTPUT resolution of the forward op; dx (fp16 view, input resolution) = result (+ add if given). */
/* add_mode: GD_GN_SAME = add is at dx's resolution; GD_GN_AVGPOOL2 = add is the gradient of an avg-pooled copy of x
 * (x_upd of a down ResBlock, unet.py:195,241) living at half resolution: dx += add[y/2,x/2] / 4. */
int gd_groupnorm_bwd(const void* x, int32_t ld, const float* mean_rstd, const float* gamma, const float* beta,
                     const float* film, int32_t film_ld, const void* dy, int32_t ld_dy, const void* add, int32_t ld_add,
                     int32_t add_mode, void* dx, int32_t ld_dx, float* partial_ws, int32_t n, int32_t h, int32_t w,
                     int32_t c, int32_t silu, int32_t spatial_mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention, head dim 64: out[n,t,head*64+j] = sum_s softmax_s(q_t.k_s / 8) v_s  (fp32 softmax).
 * Replaces QKVAttentionLegacy.forward / QKVAttention.forward, unet.py:337-354,370-389.
 * qkv: fp16 [n,t,3*heads*64] view (ld), out: fp16 [n,t,heads*64] view. lse (optional): fp32 [n,heads,t].
 * ---------------------------------------------------------------------------------------------- */
int gd_attention_fwd(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n, int32_t t,
                     int32_t heads, int32_t order, void* stream);
/* dqkv from dout, using saved qkv, out and lse. delta_ws: fp32 [n,heads,t]. */
int gd_attention_bwd(const void* qkv, int32_t ld_qkv, const void* out, int32_t ld_out, const void* dout,
                     int32_t ld_dout, const float* lse, float* delta_ws, void* dqkv, int32_t ld_dqkv, int32_t n,
                     int32_t t, int32_t heads, int32_t order, void* stream);

/* Forward for head widths other than 64 (16..128 in steps of 16, 160, 192, 224, 256): the reference's factory default
 * num_heads=4 / num_head_channels=-1 (script_util.py:54-56) gives heads of ch/4 channels (unet.py:279-285).  Same
 * contract as gd_attention_fwd with channel count heads*head_dim, scale head_dim^-1/2 on the scores; any t >= 1
 * (ragged last tiles are masked).  Forward only. */
int gd_attention_fwd_hd(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n, int32_t t,
                        int32_t heads, int32_t head_dim, int32_t order, void* stream);

/* Same kernels for a sequence padded to t (multiple of 64) of which only the first t_valid tokens exist (ViT: 197 of
 * 256): keys >= t_valid get probability 0 in the forward and in both gradient kernels; padded query rows are computed
 * like any other row and are ignored by the caller. */
int gd_attention_fwd_masked(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n, int32_t t,
                            int32_t t_valid, int32_t heads, int32_t order, void* stream);
int gd_attention_bwd_masked(const void* qkv, int32_t ld_qkv, const void* out, int32_t ld_out, const void* dout,
                            int32_t ld_dout, const float* lse, float* delta_ws, void* dqkv, int32_t ld_dqkv, int32_t n,
                            int32_t t, int32_t t_valid, int32_t heads, int32_t order, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIP ViT image-encoder guidance (BASELINE configs[2]; SURVEY §8c spec — the reference repository has no CLIP code,
 * the architecture is openai/CLIP's as implemented by transformers.CLIPVisionModelWithProjection, quick_gelu).
 * Token tensors: fp16 [rows][c] with a row stride ld (rows = n * t_pad).  GEMMs of the encoder go through
 * gd_conv_igemm (taps = 1), attention through gd_attention_*_masked.
 * ---------------------------------------------------------------------------------------------- */
/* LayerNorm over c (<= 2048, multiple of 8) per row, eps as given (CLIP: 1e-5); mean_rstd [rows][2] optional. */
int gd_layernorm_fwd(const void* x, int32_t ld, const float* gamma, const float* beta, float eps, void* out,
                     int32_t ld_out, float* mean_rstd, int32_t rows, int32_t c, void* stream);
/* dx = dLN/dx^T dy (+ add): add is an optional fp16 tensor of dx's shape (the residual branch's gradient). */
int gd_layernorm_bwd(const void* x, int32_t ld, const float* mean_rstd, const float* gamma, const void* dy, int32_t ld_dy,
                     const void* add, int32_t ld_add, void* dx, int32_t ld_dx, int32_t rows, int32_t c, void* stream);
/* QuickGELU x * sigmoid(1.702 x) and its derivative applied to dy. */
int gd_quickgelu_fwd(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t rows, int32_t c, void* stream);
int gd_quickgelu_bwd(const void* x, int32_t ld, const void* dy, int32_t ld_dy, void* dx, int32_t ld_dx, int32_t rows,
                     int32_t c, void* stream);
/* x fp32 NCHW [n,3,hin,win] in [-1,1] -> (x+1)/2 -> bilinear resize to size x size (align_corners=False) -> CLIP
 * mean/std -> patches fp16 [n][t_pad][3*patch*patch] (row stride ld): token 1 + py*g + px, k = c*P*P + dy*P + dx;
 * token 0 (class slot) and tokens > g*g are written as zeros.  _bwd is the exact transpose, times out_scale. */
int gd_clip_preprocess_fwd(const float* x, void* patches, int32_t ld, int32_t n, int32_t hin, int32_t win, int32_t size,
                           int32_t patch, int32_t t_pad, void* stream);
int gd_clip_preprocess_bwd(const void* dpatches, int32_t ld, float* dx, int32_t n, int32_t hin, int32_t win, int32_t size,
                           int32_t patch, int32_t t_pad, float out_scale, void* stream);
/* Similarity head on the post-LayerNorm class token f (fp16 rows, stride ld_f): e = Wp f (Wp fp32 [p][h]),
 * sim[n] = scale * <e/|e|, text_n> (text fp32, row stride text_stride; 0 = one shared row), and
 * df = grad_scale * d sim / d f (fp16 rows, stride ld_df).  sim or df may be NULL. */
int gd_clip_head(const void* f, int32_t ld_f, const float* wproj, const float* text, int32_t text_stride, float scale,
                 float grad_scale, float* sim, void* df, int32_t ld_df, int32_t n, int32_t h, int32_t p, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Small fp32 pieces.
 * ---------------------------------------------------------------------------------------------- */
/* nn.py:103-121 timestep_embedding: out[b] = [cos(t*f_k) | sin(t*f_k)], f_k = exp(-ln(1e4) k/half). */
int gd_timestep_embedding(const float* t, float* out, int32_t n, int32_t dim, void* stream);
/* y[m][n] = act_out( sum_k act_in(x[m][k]) * W[n][k] + b[n] ) (+ add[m][n]); nn.Linear of time_embed,
 * label_emb MLP and every ResBlock.emb_layers (unet.py:199-205,472-476; unet_other.py:29-33). */
int gd_linear_f32(const float* x, int32_t ldx, const float* w, const float* b, const float* add, int32_t ld_add,
                  float* y, int32_t ldy, int32_t m, int32_t k, int32_t n, int32_t silu_in, int32_t silu_out,
                  void* stream);
/* out[b][:] = table[idx[b]][:]  (nn.Embedding label_emb, unet.py:479,653; the add happens in gd_linear_f32). */
int gd_embedding_gather(const float* table, const int64_t* idx, float* out, int32_t n, int32_t dim, int32_t num_rows,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention-pool head of the classifier (unet.py:22-51, 833-841) forward and dX backward, fp32.
 * h: fp16 NHWC [n,s,s,c] = SiLU(GN(h)) already applied.  See csrc/attnpool.cu for workspace layout.
 * ---------------------------------------------------------------------------------------------- */
int64_t gd_attnpool_ws_floats(int32_t n, int32_t tokens, int32_t c);
int gd_attnpool_fwd(const void* h, int32_t ld, const float* pos_emb, const float* w_qkv, const float* b_qkv,
                    const float* w_c, const float* b_c, float* logits, float* ws, int32_t n, int32_t hw, int32_t c,
                    int32_t heads, int32_t n_out, void* stream);
/* w_qkv_t: [c][3c] and w_c_t: [c][n_out] are host-side transposes of the forward weights. */
int gd_attnpool_bwd(const float* dlogits, const float* w_qkv_t, const float* w_c_t, float* ws, void* dh, int32_t ld_dh,
                    int32_t n, int32_t hw, int32_t c, int32_t heads, int32_t n_out, float out_scale, void* stream);
/* dlogits[b][j] = scale * (1[j==y_b] - softmax(logits[b])_j): gradient of sum_b log_softmax(logits)[b,y_b]
 * (scripts/classifier_sample.py:58-61). */
int gd_logsoftmax_select_bwd(const float* logits, const int64_t* y, float* dlogits, int32_t n, int32_t classes,
                             float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused posterior / noise update, one launch per sampling step (gaussian_diffusion.py:232-326, 356-393,
 * 395-439, 546-594).  All tensors fp32 NCHW [n,3,h,w] except model_out [n,6 or 3,h,w].
 * coef: device fp32 table [T'][GD_COEF_STRIDE]; t: device int64 [n], the (respaced) timestep index of each
 * sample selecting the row.  noise == NULL computes p_mean_variance only (mean/var/logvar/pred_xstart).
 * ---------------------------------------------------------------------------------------------- */
enum {
  GD_COEF_SQRT_RECIP_ACP = 0,   /* sqrt_recip_alphas_cumprod      gaussian_diffusion.py:150 */
  GD_COEF_SQRT_RECIPM1_ACP = 1, /* sqrt_recipm1_alphas_cumprod    :151 */
  GD_COEF_POST_MEAN1 = 2,       /* posterior_mean_coef1           :162 */
  GD_COEF_POST_MEAN2 = 3,       /* posterior_mean_coef2           :165 */
  GD_COEF_LOG_BETA = 4,         /* log(betas) (max_log)           :272 */
  GD_COEF_POST_LOGVAR = 5,      /* posterior_log_variance_clipped :159 */
  GD_COEF_FIXED_VAR = 6,        /* model_variance for FIXED_*     :278-291 */
  GD_COEF_FIXED_LOGVAR = 7,
  GD_COEF_ACP = 8,              /* alphas_cumprod                 :141 */
  GD_COEF_ACP_PREV = 9,         /* alphas_cumprod_prev            :142 */
  GD_COEF_NONZERO = 10,         /* 1.0 if t != 0 else 0.0         :431-433 */
  GD_COEF_ACP_NEXT = 11,        /* alphas_cumprod_next            :143 (ddim_reverse_sample :596-632) */
  GD_COEF_STRIDE = 12
};
enum { GD_VAR_LEARNED_RANGE = 0, GD_VAR_FIXED = 1, GD_VAR_LEARNED = 2 };
enum { GD_MEAN_EPSILON = 0, GD_MEAN_START_X = 1 };
typedef struct gd_posterior_desc {
  const float* x;         /* x_t */
  const float* model_out; /* eps (and v) */
  const float* grad;      /* cond_fn output (already scaled) or NULL */
  const float* noise;     /* z, or NULL */
  float* sample;          /* x_{t-1} (may alias x) */
  float* pred_xstart;     /* may be NULL */
  float* mean_out;        /* optional p_mean_variance outputs (pre-guidance), may be NULL */
  float* var_out;
  float* logvar_out;
  const float* coef;
  const int64_t* t;
  int32_t n, c, hw;
  int32_t var_type, mean_type, clip_denoised;
  int32_t ddim;           /* 0 ancestral p_sample, 1 ddim_sample, GD_DDIM_REVERSE ddim_reverse_sample (no noise / grad) */
  float eta;
  int32_t num_timesteps;  /* rows of `coef`; a t[b] outside [0, num_timesteps) yields NaN outputs for sample b
                           * (the reference raises IndexError in _extract_into_tensor, gaussian_diffusion.py:904-917) */
} gd_posterior_desc;
enum { GD_DDIM_REVERSE = 2 };
int gd_posterior_step(const gd_posterior_desc* desc, void* stream);

/* ((x+1)*127.5).clamp(0,255).to(uint8) NCHW -> NHWC (scripts/classifier_sample.py:87-89; truncation). */
int gd_to_uint8_nhwc(const float* x, uint8_t* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);

/* Second half of a 3x3 convolution with very few output channels (the UNet's 256->6 `out` head, unet.py:613-617, and
 * the classifier's 128->3 data-gradient conv): the conv is evaluated as ONE 1x1 GEMM with 9*cout output columns,
 * ytap[n][y][x][tap*cout+co] = sum_ci W[co][ci][ky][kx] * in[n][y][x][ci] (gd_conv_igemm, taps = 1, fp16 NHWC output with
 * pixel stride ld), and this kernel gathers the taps into fp32 NCHW:
 *   out[n][co][y][x] = out_scale * (bias[co] + sum_tap ytap[n][y+ky-1][x+kx-1][tap*cout+co]),  zero outside the image.
 * A tensor-core tile with 6 useful columns of 16 runs at 3 % of peak; the 54-of-64-column GEMM an order of magnitude
 * faster.  cout <= 7. */
int gd_tap_gather3x3(const void* ytap, int32_t ld, const float* bias, float* out, int32_t n, int32_t cout, int32_t h,
                     int32_t w, float out_scale, void* stream);

/* Layout helpers used at the API boundary and by tests. */
int gd_nchw_f32_to_nhwc_f16(const float* x, void* out, int32_t ld_out, int32_t n, int32_t c, int32_t h, int32_t w,
                            void* stream);
int gd_nhwc_f16_to_nchw_f32(const void* x, int32_t ld, float* out, int32_t n, int32_t c, int32_t h, int32_t w,
                            void* stream);
/* bilinear (align_corners=False) upsample of low_res fp32 NCHW into channels of an NCHW fp32 buffer
 * (SuperResModel.forward, unet.py:677-681). */
int gd_bilinear_upsample_nchw(const float* x, float* out, int32_t n, int32_t c, int32_t h_in, int32_t w_in,
                              int32_t h_out, int32_t w_out, int32_t out_c_total, int32_t out_c_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GD_B200_H_ */
