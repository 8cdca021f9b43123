/*
 * cvlm.h -- C ABI of the B200-native SAM-Adapter image encoder, its input pipeline and its mask decoder
 * (libcvlm_b200.so).
 *
 * The reference (intcomp/camouflaged-vlm) has no FFI or plugin layer on this path: its boundary is
 * the Python class models/mmseg/models/sam/image_encoder.py:25 `ImageEncoderViT`, called as
 * `self.image_encoder(input, interm=True)` (models/sam_maskdecoder_edge.py:278, 311, 337;
 * models/sam.py:406, 431, 456).  This header is the boundary a maintainer binds instead of that
 * module's torch ops (ctypes stub in INTEGRATION.md; the in-tree binding is
 * camouflaged_vlm_b200/_lib.py).  Everything is `extern "C"`, plain pointers and sizes, no torch or
 * CUDA types: device buffers are `void*` / `float*` device addresses, the stream is the
 * `cudaStream_t` handle passed as `void*`.
 *
 * Conventions
 *   - every function returns CVLM_OK (0) or a negative error code; cvlm_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - all launches are asynchronous on `stream`; nothing synchronises the device.
 *   - the library owns no device memory: weights, workspace and outputs are caller-allocated.
 *   - bf16 buffers are raw uint16 bit patterns (`__nv_bfloat16`).
 */
#ifndef CVLM_H_
#define CVLM_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CVLM_OK 0
#define CVLM_ERR_ARG (-1)          /* null / misaligned pointer, bad shape */
#define CVLM_ERR_CUDA (-2)         /* a CUDA runtime / driver call failed */
#define CVLM_ERR_UNSUPPORTED (-3)  /* shape outside what the kernels implement */
#define CVLM_ERR_MISSING (-4)      /* a required weight tensor was not registered or has the wrong size */
#define CVLM_ERR_WORKSPACE (-5)    /* workspace too small */

#define CVLM_MAX_DEPTH 64

typedef void* cvlm_stream_t; /* cudaStream_t */

const char* cvlm_last_error(void);
int cvlm_abi_version(void);

/* Test / tuning switches, process-wide, read by the launchers as one integer load (no environment lookups on the
 * launch path; the CVLM_* environment variables of the same meaning only provide the initial values):
 *   "gemm_mode"    0 automatic, 1 one CTA per 128x256 tile, 2 CTA pairs (cta_group::2) per 256x256 tile
 *   "dec_cluster"  0 automatic, 1/2/4/8 CTAs per image on the mask decoder's token side
 *   "attn_legacy"  1 forces the mma.sync attention kernel (A/B runs)
 *   "attn_poly"    how many of every 16 exponential pairs of the global attention run on the FMA pipe (0,2,4,6,8)
 *   "win_mode"     window-attention kernel generation (0 = default)
 *   "bwd_tc"       training, backward of the 64 x 64 global-attention blocks: 0 = mma.sync kernels, 1 = dK / dV on the
 *                  tcgen05 kernel, 2 = dQ as well (default; needs the tables the training forward leaves)
 * No reference analogue (the reference has no native layer on this path). */
int cvlm_set_option(const char* name, int value);
int cvlm_get_option(const char* name, int* value);

/* ------------------------------------------------------------------------------------------------
 * Encoder handle.  Mirrors ImageEncoderViT.__init__ (image_encoder.py:26-44) after the Python side
 * has resolved norm_layer / act_layer to LayerNorm(eps) / GELU.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cvlm_encoder_config {
    int img_size;     /* image side in pixels (square), multiple of 16 */
    int patch_size;   /* 16 */
    int in_chans;     /* 3 */
    int embed_dim;    /* D */
    int depth;        /* number of blocks, <= CVLM_MAX_DEPTH */
    int num_heads;    /* D / num_heads must be 32, 64 or 80 */
    int mlp_dim;      /* int(D * mlp_ratio) */
    int out_chans;    /* neck width, multiple of 8 */
    int window_size;  /* window side of the non-global blocks (0: every block global) */
    int is_global[CVLM_MAX_DEPTH]; /* 1 if block i uses global attention (global_attn_indexes) */
    int use_abs_pos;  /* pos_embed present */
    int rank;         /* adapter width D / scale_factor (image_encoder.py:252) */
    int rank_pad;     /* rank rounded up to a multiple of 8: one of 8,16,24,32,40 */
    float ln_eps;     /* eps of the block LayerNorms (callers bind 1e-6) */
    float ln2d_eps;   /* eps of LayerNorm2d (common.py:32, 1e-6) */
} cvlm_encoder_config;

typedef struct cvlm_encoder cvlm_encoder;

int cvlm_encoder_create(const cvlm_encoder_config* cfg, cvlm_encoder** out);
int cvlm_encoder_destroy(cvlm_encoder* enc);

/* Register one packed weight tensor (device pointer, 16-byte aligned).  Names and layouts are listed
 * in DESIGN.md section "Packed weights"; camouflaged_vlm_b200/packing.py derives them from the reference
 * state_dict.  Re-registering a name replaces it (load_state_dict). */
int cvlm_encoder_set_tensor(cvlm_encoder* enc, const char* name, const void* dev_ptr, long long nbytes);

/* Scratch bytes needed by cvlm_encoder_forward for `batch` images. */
long long cvlm_encoder_workspace_bytes(const cvlm_encoder* enc, int batch);

/* ImageEncoderViT.forward(x, interm) (image_encoder.py:132-155).
 *   images     fp32 NCHW [batch, in_chans, img, img], device
 *   highpass   NULL (computed here: hp.* tensors, csrc/highpass.cu), or fp32 NCHW of the same shape holding
 *              PromptGenerator.fft(images, 0.25) (:332-353), device
 *   out        fp32 NCHW [batch, out_chans, g, g], device
 *   interm     NULL, or one device pointer per global block (in block order), each fp32 NHWC
 *              [batch, g, g, D]: receives the residual stream after that block; the Python module
 *              hands callers .permute(0,3,1,2) views of these, exactly like the reference (:148)
 *   workspace  device scratch of at least cvlm_encoder_workspace_bytes(enc, batch), 256-byte aligned */
int cvlm_encoder_forward(cvlm_encoder* enc, const float* images, const float* highpass, int batch, float* out,
                         float* const* interm, void* workspace, long long workspace_bytes, cvlm_stream_t stream);

/* Same, but stops after block `stop_after_block` (0-based; -1 = after the stem, before block 0's
 * prompt add is NOT included: the stem output already holds x + pos_embed + prompt_0) and copies the
 * fp32 NHWC residual stream [batch*g*g, D] to `x_out`.  Parity-debug aid, not used by the product path. */
int cvlm_encoder_forward_partial(cvlm_encoder* enc, const float* images, const float* highpass, int batch,
                                 int stop_after_block, float* x_out, void* workspace, long long workspace_bytes,
                                 cvlm_stream_t stream);

/* Number of kernel launches the last cvlm_encoder_forward on this handle issued (for bench.py's
 * gpu_launches claim). */
/* The same forward for a batch that lives in HOST memory (the reference-facing call of a serving loop: images arrive on
 * the host, the decoder / caller wants the embeddings back; replaces `self.image_encoder(input.to(device), ...)`,
 * models/sam_maskdecoder_edge.py:203-208, 337).  images_host fp32 [batch,3,S,S] and out_host fp32 [batch,out,g,g] should be
 * pinned.  The batch is cut into chunks of `chunk` images: the upload of chunk i+1 and the download of the embeddings of
 * chunk i-1 run on `copy_stream` underneath the encoder of chunk i on `stream` (stage_dev: device, 2*chunk*3*S*S floats;
 * emb_dev: device, [batch,out,g,g]; interm_dev: NULL or one device pointer [batch,g,g,D] per global block; workspace for
 * `chunk` images).  Stream-ordered only: out_host is valid once `stream` has been synchronised.  Back-to-back calls on the
 * same handle, streams and stage_dev overlap: the first upload of a call runs underneath the last encoder of the previous
 * one (a staging half waits only for the encoder that last read it), so images_host must be complete when the call is made
 * -- it is host memory, no stream orders it -- and stay untouched until `copy_stream` has consumed it. */
int cvlm_encoder_forward_host(cvlm_encoder* enc, const float* images_host, int batch, int chunk, float* out_host,
                              float* emb_dev, float* const* interm_dev, float* stage_dev, void* workspace,
                              long long workspace_bytes, cvlm_stream_t copy_stream, cvlm_stream_t stream);

long long cvlm_encoder_last_launch_count(const cvlm_encoder* enc);

/* Per-kernel timing.  With profiling enabled every launch of cvlm_encoder_forward is bracketed by CUDA
 * events on `stream`; cvlm_encoder_get_profile waits for them, sums the elapsed milliseconds and the
 * launch counts per kind since the previous call, and resets.  Kinds, in order: qkv GEMM, proj GEMM,
 * lin1 GEMM, lin2 GEMM, other GEMMs (stem, adapter, neck), window attention, global attention,
 * LayerNorm, other row kernels. */
#define CVLM_PROF_KINDS 9
int cvlm_encoder_set_profiling(cvlm_encoder* enc, int enable);
int cvlm_encoder_get_profile(cvlm_encoder* enc, float* ms_by_kind, int* launches_by_kind);

/* ------------------------------------------------------------------------------------------------
 * Adapter-only training.  The reference trains `prompt_generator.*` with every other encoder
 * parameter frozen (train_ovcos_maskdecoder_edge.py:301-303); autograd carries the gradient through all blocks
 * (models/sam_maskdecoder_edge.py:429-460: loss.backward() through self.image_encoder(input, interm=True)).
 * These two calls replace that autograd graph: the training forward is the inference schedule that also keeps the
 * activations the backward pass needs in a caller-provided `saved` region (3.4 GB per ViT-H image), the backward pass
 * turns d(loss)/d(embedding) [and d(loss)/d(interm_k)] into the gradients of the adapter tensors.
 * Extra tensors to register before cvlm_encoder_backward (bf16, transposed copies of the frozen weights):
 *   blk.<i>.lin2.wt [4D, D], blk.<i>.lin1.wt [D, 4D] (norm2 gamma folded), blk.<i>.proj.wt [D, D],
 *   blk.<i>.qkv.wt [D, 3D] (transpose of blk.<i>.qkv.w), neck.conv1.wt [D, out], neck.conv2.wt [out, 9*out]
 *   (flipped taps), shared.wt [rank_pad, D], lw.wt [rank_pad, depth*rank_pad].
 * Gradient buffer (fp32, cvlm_encoder_grad_floats floats; cvlm_encoder_grad_layout gives the 5 offsets and jp):
 *   [0] shared  [D, jp]             cols < rank: d shared_mlp.weight[d, r];  col rank_pad: d shared_mlp.bias[d]
 *   [1] lw      [depth*rank_pad, jp] row i*rank_pad+s, col r: d lightweight_mlp_i.weight[s, r]; col rank_pad: bias
 *   [2] hand    [3*16*16, jp]        row k, col r: d prompt_generator.proj.weight[r, k]
 *   [3] dtsum   [rank_pad, jp]       col rank_pad, row r: d prompt_generator.proj.bias[r] = d embedding_generator.bias[r]
 *   [4] eg      [D, rank_pad]        d embedding_generator.weight[r, d] transposed
 * dout_nhwc: fp32 [batch*T, out_chans] (d loss / d embedding, channels last); dinterm_nhwc: NULL or one fp32
 * [batch*T, D] pointer (or NULL) per global block.
 * ---------------------------------------------------------------------------------------------- */
long long cvlm_encoder_saved_bytes(const cvlm_encoder* enc, int batch);
long long cvlm_encoder_backward_workspace_bytes(const cvlm_encoder* enc, int batch);
long long cvlm_encoder_grad_floats(const cvlm_encoder* enc);
int cvlm_encoder_grad_layout(const cvlm_encoder* enc, long long* offsets5, int* jp);
int cvlm_encoder_forward_train(cvlm_encoder* enc, const float* images, int batch, float* out, float* const* interm,
                               void* saved, long long saved_bytes, void* workspace, long long workspace_bytes,
                               cvlm_stream_t stream);
int cvlm_encoder_backward(cvlm_encoder* enc, int batch, const float* dout_nhwc, const float* const* dinterm_nhwc,
                          const void* saved, long long saved_bytes, float* grads, long long grads_floats, void* workspace,
                          long long workspace_bytes, cvlm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Op-level entry points (the kernels behind the encoder; used by the parity tests and profiling).
 * ---------------------------------------------------------------------------------------------- */

/* Backward of cvlm_op_attention: dqkv (bf16 [batch*G*G, 3*D]) from qkv, the forward output o and dout (bf16
 * [batch*G*G, D]); lse / dsum: fp32 [batch*G*G, heads] scratch (row log-sum-exp, rowsum(dout*o)).  mma.sync kernels. */
int cvlm_op_attention_backward(const void* qkv, const void* o, const void* dout, void* dqkv, float* lse, float* dsum,
                               const void* rel_pos_h, const void* rel_pos_w, int batch, int grid_side, int window,
                               int heads, int head_dim, cvlm_stream_t stream);
/* cvlm_op_attention that also leaves every token's row log-sum-exp (log2 domain, relative-position terms included) in
 * lse, fp32 [batch*G*G, heads] -- what the training forward keeps for the backward kernels.  CVLM_ERR_UNSUPPORTED for
 * shapes that run on the mma.sync fallback kernel, which does not produce it.  tw_tab / th_tab (both or neither; 64 x 64
 * global blocks only): fp32 [batch*G*G, heads, 64] each, every query's relative-position terms laid out for the dK / dV
 * kernel: tw[q][kx] = (q . rel_pos_w[qx - kx + 63]) log2(e), th[q][ky] likewise with rel_pos_h. */
int cvlm_op_attention_lse(const void* qkv, void* out, const void* rel_pos_h, const void* rel_pos_w, int batch,
                          int grid_side, int window, int heads, int head_dim, float* lse, float* tw_tab, float* th_tab,
                          cvlm_stream_t stream);
/* cvlm_op_attention_backward with lse already filled by cvlm_op_attention_lse: the kernels skip their own sweep over
 * S = q k^T that would compute it; with the tables as well (may be NULL) the dK / dV kernel of the global blocks does not
 * project the queries on the relative-position tables either and double-buffers its query tiles. */
int cvlm_op_attention_backward_lse(const void* qkv, const void* o, const void* dout, void* dqkv, const float* lse,
                                   float* dsum, const void* rel_pos_h, const void* rel_pos_w, int batch, int grid_side,
                                   int window, int heads, int head_dim, const float* tw_tab, const float* th_tab,
                                   cvlm_stream_t stream);
/* LayerNorm backward, one warp per row: dx = (residual ? residual : 0) + rstd * (g - mean(g) - xhat * mean(g * xhat)),
 * g = gin * gamma (gamma NULL: already folded).  xhat_kind 0: xsrc = raw bf16 x, stat = (rstd, -mean*rstd) rows;
 * 1: xsrc = bf16 xhat, stat = rstd rows; 2: xsrc = fp32 x, statistics recomputed with eps.  dxb (optional) = bf16(dx). */
int cvlm_op_ln_bwd(int xhat_kind, const float* gin, const float* gamma, const void* xsrc, const float* stat,
                   const float* residual, float* dx, void* dxb, long long rows, int cols, float eps, cvlm_stream_t stream);
/* Weight gradient out[i, j] = sum_m A[m, i] * B[m, j] (+ column J = sum_m A[m, i] when ones != 0): A bf16 [m, I], B bf16
 * [m, J]; out fp32 [I, jp] with jp = cvlm_op_wgrad_tn_jp(J, ones); scratch fp32 [cvlm_op_wgrad_tn_splits(m), I, jp]. */
int cvlm_op_wgrad_tn_jp(int j, int ones);
int cvlm_op_wgrad_tn_splits(long long m);
int cvlm_op_wgrad_tn(const void* a, long long lda, int i_cols, const void* b, long long ldb, int j_cols, int ones,
                     long long m, float* scratch, float* out, cvlm_stream_t stream);
/* cvlm_op_gemm_bf16 with the backward-of-GELU epilogue: out (bf16) = (A * B^T) * aux[row, col], aux bf16 [m, ldaux] =
 * gelu'(pre-activation) as the training forward stores it (copy_bf16 of a bf16 GEMM with act_gelu = 1). */
int cvlm_op_gemm_bf16_dgelu(const void* a, long long lda, const void* b, long long ldb, int k, int m, int n, int b_rows,
                            const void* aux, long long ldaux, void* out, long long ldo, cvlm_stream_t stream);

/* out = epilogue(A1 * B1^T [+ A2 * B2^T]); A* bf16 [m, k*] row stride lda*, B* bf16 [b_rows, k*] row
 * stride ldb* (weights, nn.Linear layout), rows >= b_rows read as zero.  Epilogue: + bias[n] (fp32),
 * optional exact-erf GELU, + residual fp32 [*, ldr] (row = row % res_row_mod when res_row_mod > 0),
 * store fp32 or bf16 with row stride ldo.  n % 8 == 0.  tcgen05 / TMEM / TMA kernel. */
int cvlm_op_gemm_bf16(const void* a1, long long lda1, const void* b1, long long ldb1, int k1, const void* a2,
                      long long lda2, const void* b2, long long ldb2, int k2, int m, int n, int b_rows,
                      const float* bias, const float* residual, long long ldr, int res_row_mod, void* out,
                      long long ldo, int out_bf16, int act_gelu, cvlm_stream_t stream);

/* Width of one column group of the row statistics below (one epilogue warp's share of an accumulator tile). */
#define CVLM_STAT_GROUP_COLS 128

/* The same GEMM with nn.LayerNorm folded around it (image_encoder.py:430-446 norm1 / norm2; DESIGN.md "LayerNorm fold").
 * Producer options (fp32 output only): copy_bf16 != NULL also stores bf16(out) with row stride ldc (must not alias
 * a1 / a2); stat_out != NULL receives fp32 [m, ceil(n/CVLM_STAT_GROUP_COLS), 2] = per row and column group (sum, sum of squares).
 * Consumer options (bf16 output only): ln_rows != NULL is fp32 [m, 2] = (rstd, -mean*rstd) of this GEMM's A rows
 * (cvlm_op_ln_finalize); the epilogue stores rstd * (acc - mean * ln_colsum[n]) + bias[n] (then GELU), i.e.
 * LayerNorm(x) W^T + b when B holds gamma-scaled weights, ln_colsum their row sums and bias = W beta + b. */
int cvlm_op_gemm_bf16_ln(const void* a1, long long lda1, const void* b1, long long ldb1, int k1, const void* a2,
                         long long lda2, const void* b2, long long ldb2, int k2, int m, int n, int b_rows,
                         const float* bias, const float* residual, long long ldr, int res_row_mod, void* out,
                         long long ldo, int out_bf16, int act_gelu, void* copy_bf16, long long ldc, float* stat_out,
                         const float* ln_rows, const float* ln_colsum, cvlm_stream_t stream);

/* partial fp32 [rows, groups, 2] (stat_out above) -> rows_out fp32 [rows, 2] = (rstd, -mean*rstd) of a LayerNorm of
 * `width` columns (biased variance, eps inside the square root). */
int cvlm_op_ln_finalize(const float* partial, int groups, int width, float eps, float* rows_out, long long rows,
                        cvlm_stream_t stream);

/* Window (window>0) or global (window==0) attention with decomposed rel-pos on the token-major qkv
 * buffer bf16 [batch*g*g, 3*heads*head_dim] (K and V stored without bias); out bf16 [batch*g*g, D]. */
int cvlm_op_attention(const void* qkv, void* out, const void* rel_pos_h, const void* rel_pos_w, int batch,
                      int grid_side, int window, int heads, int head_dim, cvlm_stream_t stream);

/* Plain attention softmax(q k^T / sqrt(head_dim)) v over `tokens` tokens per image, no relative-position terms
 * (Attention.forward of the Alpha-CLIP visual tower, alpha_clip_rw/model.py:224-256 with default arguments; the kernel
 * of SURVEY.md 8 f2, exposed ahead of that row).  Same tcgen05 pipeline as the global attention.
 * qkv bf16 [batch * tokens, 3 * heads * head_dim] laid out (q | k | v) x heads x head_dim, out bf16
 * [batch * tokens, heads * head_dim]; head_dim 64 or 80; any tokens >= 1. */
int cvlm_op_attention_plain(const void* qkv, void* out, int batch, int tokens, int heads, int head_dim,
                            cvlm_stream_t stream);

/* nn.LayerNorm over the last dim of fp32 [rows, cols]; out bf16 (out_bf16=1) or fp32. */
int cvlm_op_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_bf16,
                      long long rows, int cols, float eps, cvlm_stream_t stream);

/* LayerNorm2d of token-major fp32 [batch*tokens, chans], written as NCHW fp32 [batch, chans, tokens]. */
int cvlm_op_layernorm2d_nchw(const float* x, const float* gamma, const float* beta, float* out_nchw, int batch,
                             int tokens, int chans, float eps, cvlm_stream_t stream);

/* fp32 NCHW image -> bf16 patch matrix [batch*g*g, chans*256] for a 16x16 stride-16 conv. */
int cvlm_op_im2col_patch16(const float* img, void* out, int batch, int chans, int img_size, cvlm_stream_t stream);

/* bf16 NHWC [batch, g, g, chans] -> bf16 [batch*g*g, 9*chans] for a 3x3 pad-1 conv. */
int cvlm_op_im2col_3x3(const void* x_nhwc, void* out, int batch, int grid_side, int chans, cvlm_stream_t stream);

int cvlm_op_cast_bf16(const float* x, void* out, long long n, cvlm_stream_t stream);

/* PromptGenerator.fft(images, rate) (image_encoder.py:332-353) as two tcgen05 GEMMs with the low-pass projector
 * (csrc/highpass.cu): out = |images - Pr X Pr^T + Pi X Pi^T| per channel image, fp32 [batch, chans, img, img].
 * pcat = [Pr; Pi] bf16 [2*img, img]; npr = -Pr, pi = Pi bf16 [img, ceil(img/64)*64], zero padded (packing.py: hp.*);
 * scratch: device bytes >= cvlm_op_high_pass_scratch_bytes. */
long long cvlm_op_high_pass_scratch_bytes(int batch, int chans, int img_size);
int cvlm_op_high_pass(const float* images, float* out, int batch, int chans, int img_size, const void* pcat,
                      const void* npr, const void* pi, void* scratch, long long scratch_bytes, cvlm_stream_t stream);

/* The reference's image input pipeline (datasets/wrappers.py:22-27: transforms.Resize((S,S)) on a PIL image ->
 * ToTensor -> Normalize), bit-exact: img uint8 HWC [h, w, 3] (device) -> out fp32 [3, out_size, out_size] (device).
 * mean / std: 3 host floats each.  scratch: device bytes >= cvlm_op_preprocess_u8_scratch_bytes, 16-byte aligned. */
long long cvlm_op_preprocess_u8_scratch_bytes(int h, int w, int out_size);
int cvlm_op_preprocess_u8(const unsigned char* img, int h, int w, float* out, int out_size, const float* mean,
                          const float* std, void* scratch, long long scratch_bytes, cvlm_stream_t stream);

/* t = bf16(hand + cst + W_eg * scramble(x0)); see rowops.cu / DESIGN.md "adapter". */
int cvlm_op_adapter_embed(const float* x0, const float* hand, const float* w_t, const float* cst, void* t_out,
                          int batch, int tokens, int dim, int rank, int rank_pad, cvlm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Mask decoder: MaskDecoder_Edge.forward / predict_masks
 * (models/mmseg/models/sam/mask_decoder_edge.py:96-190) over TwoWayTransformer_MaskDecoder_Edge
 * (transformer_maskdecoder_edge.py:62-214), batched over images (image b is decoded with prompt b; the
 * reference itself only works at batch 1).  csrc/decoder.cu.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cvlm_decoder_config {
    int dim;              /* transformer_dim, 256 */
    int heads;            /* 8 */
    int mlp_dim;          /* 2048 */
    int depth;            /* TwoWayTransformer depth, 2 */
    int num_mask_tokens;  /* num_multimask_outputs + 1 = 4 (the edge token is added on top) */
    int grid;             /* side of the image embedding (64) */
    int cond_tokens;      /* prompt tokens per image in sparse_prompt_embeddings (2) */
    float ln_eps;         /* nn.LayerNorm of the transformer (1e-5) */
    float ln2d_eps;       /* LayerNorm2d of the up-scaling stacks (common.py:32, 1e-6) */
} cvlm_decoder_config;

typedef struct cvlm_decoder cvlm_decoder;

int cvlm_decoder_create(const cvlm_decoder_config* cfg, cvlm_decoder** out);
int cvlm_decoder_destroy(cvlm_decoder* dec);
/* fp32 parameters under the reference's state_dict names, plus the bf16 / re-laid-out copies listed in
 * camouflaged_vlm_b200/mask_decoder.py (pack_decoder_weights). */
int cvlm_decoder_set_tensor(cvlm_decoder* dec, const char* name, const void* dev_ptr, long long nbytes);
long long cvlm_decoder_workspace_bytes(const cvlm_decoder* dec, int batch);
long long cvlm_decoder_last_launches(const cvlm_decoder* dec);

/* predict_masks for `batch` images.
 *   image_embeddings fp32 NCHW [batch, 256, g, g]        image_pe fp32 NCHW [1, 256, g, g]
 *   sparse_prompt    fp32 [batch, cond_tokens, 256]       dense_prompt fp32 NCHW [dense_batch, 256, g, g], dense_batch 1 or batch
 *   masks  fp32 [batch, 4, 4g, 4g]  (all mask tokens; the caller slices like mask_decoder_edge.py:133-139)
 *   edge   fp32 [batch, 1, 4g, 4g]      iou_pred fp32 [batch, 4]
 *   hs     NULL or fp32 [batch, 6, 256]: the transformer's output tokens
 *   workspace: device, 256-byte aligned, >= cvlm_decoder_workspace_bytes(dec, batch) */
int cvlm_decoder_forward(cvlm_decoder* dec, const float* image_embeddings, const float* image_pe, const float* sparse_prompt,
                         const float* dense_prompt, int dense_batch, int batch, float* masks, float* edge, float* iou_pred,
                         float* hs, void* workspace, long long workspace_bytes, cvlm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Alpha-CLIP visual tower with MaPLe deep prompts: VisionTransformer_MaPLe.forward
 * (alpha_clip_rw/model.py:507-563), called as image_encoder(images, masks, shared_ctx, deep_prompts)
 * (cocotrainers/mapleAlphaCLIP.py:272, 286).  csrc/alpha_clip.cu.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cvlm_vit_config {
    int input_resolution; /* 336 */
    int patch_size;       /* 14 (even, divides the resolution) */
    int width;            /* 1024 (multiple of 128, <= 2048) */
    int layers;           /* 24 */
    int heads;            /* 16 (width / heads must be 64 or 80) */
    int output_dim;       /* 768 (multiple of 8) */
    int n_ctx;            /* MaPLe context tokens appended to the sequence (design_details['maple_length']) */
    float ln_eps;         /* 1e-5 */
} cvlm_vit_config;

typedef struct cvlm_vit cvlm_vit;

int cvlm_vit_create(const cvlm_vit_config* cfg, cvlm_vit** out);
int cvlm_vit_destroy(cvlm_vit* vit);
/* fp32 parameters under the reference's state_dict names, plus the bf16 copies listed in
 * camouflaged_vlm_b200/alpha_clip.py (pack_vit_weights). */
int cvlm_vit_set_tensor(cvlm_vit* vit, const char* name, const void* dev_ptr, long long nbytes);
long long cvlm_vit_workspace_bytes(const cvlm_vit* vit, int batch);
long long cvlm_vit_last_launches(const cvlm_vit* vit);

/*   images fp32 NCHW [batch, 3, R, R]      alpha fp32 [batch, 1, R, R]
 *   shared_ctx fp32 [n_ctx, width]         deep_prompts fp32 [n_deep, n_ctx, width] (NULL when n_deep == 0)
 *   out fp32 [batch, output_dim]           workspace: device, 256-byte aligned, >= cvlm_vit_workspace_bytes(vit, batch) */
int cvlm_vit_forward(cvlm_vit* vit, const float* images, const float* alpha, const float* shared_ctx, const float* deep_prompts,
                     int n_deep, int batch, float* out, void* workspace, long long workspace_bytes, cvlm_stream_t stream);

/* F.interpolate(mode="bilinear", align_corners=False) on `planes` fp32 images [in_h, in_w] -> [out_h, out_w]
 * (postprocess_masks, models/sam_maskdecoder_edge.py:359-388). */
int cvlm_op_bilinear_resize(const float* in, float* out, long long planes, int in_h, int in_w, int out_h, int out_w,
                            cvlm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CVLM_H_ */
