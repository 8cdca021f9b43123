// Encoder runtime: the C-ABI (include/cvlm.h) and the launch schedule of ImageEncoderViT.forward
// (image_encoder.py:132-155) over the kernels in gemm_tcgen05.cu / attention.cu / rowops.cu.
//
// Data layout in HBM for a batch of B images (M = B * T tokens, T = g*g, D = embed_dim):
//   x     fp32 [M, D]      residual stream (NHWC), updated in place by the proj / lin2 epilogues
//   a     bf16 [M, D]      LayerNorm output, then attention output, then LayerNorm output again
//   big   bf16 [M, 4D]     qkv (3D) / MLP hidden (4D) / im2col patches / neck im2col (they never coexist)
//   hand  fp32 [M, RP]     PatchEmbed2 of the high-passed image
//   t     bf16 [M, RP]     hand + embedding (adapter input)
//   gall  bf16 [M, depth*RP]  GELU(lightweight_mlp_i(t)) for all i; prompt_i = gall_i * W_shared^T + b
// The 32 prompt tensors the reference materialises (32 x [B,4096,1280] fp32) never exist: prompt_{i+1}
// is a second K segment (K = RP) of block i's lin2 GEMM.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "cvlm_internal.cuh"
#include "encoder_runtime.cuh"

namespace cvlm {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int device_sm_count() {
    static int sms[kMaxDevices] = {};       // 0 = not asked yet; written once per device with the same value
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < kMaxDevices;
    int n = cacheable ? __atomic_load_n(&sms[dev], __ATOMIC_RELAXED) : 0;
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        if (cacheable) __atomic_store_n(&sms[dev], n, __ATOMIC_RELAXED);
    }
    return n;
}

// Options (cvlm_set_option): test / tuning hooks.  Their initial values come from the environment ONCE, at the first
// use, so that `CVLM_GEMM_MODE=pair python tools/op_bench.py` keeps working; the launch path only loads an int.
static const char* const kOptionNames[OPT_COUNT] = {"gemm_mode", "dec_cluster", "attn_legacy", "attn_poly", "win_mode", "bwd_tc"};
static int g_options[OPT_COUNT];
static std::once_flag g_options_once;
static void options_from_env() {
    for (int& v : g_options) v = 0;
    g_options[OPT_ATTN_POLY] = 4;
    g_options[OPT_BWD_TC] = 2;            // backward of the global blocks on the tcgen05 kernels when the forward left its tables
    if (const char* e = getenv("CVLM_GEMM_MODE")) g_options[OPT_GEMM_MODE] = e[0] == '1' ? 1 : e[0] == 'p' ? 2 : 0;
    if (const char* e = getenv("CVLM_DEC_CLUSTER")) g_options[OPT_DEC_CLUSTER] = atoi(e);
    if (const char* e = getenv("CVLM_ATTN_LEGACY")) g_options[OPT_ATTN_LEGACY] = e[0] == '1';
    if (const char* e = getenv("CVLM_ATTN_POLY")) g_options[OPT_ATTN_POLY] = atoi(e);
    if (const char* e = getenv("CVLM_WIN_MODE")) g_options[OPT_WIN_MODE] = atoi(e);
}
int option(Option o) {
    std::call_once(g_options_once, options_from_env);
    return __atomic_load_n(&g_options[o], __ATOMIC_RELAXED);
}

}  // namespace cvlm

namespace cvlm {

struct Workspace {
    long long off_x, off_a, off_big, off_hand, off_t, off_gall, total;
    long long big_cols;
    long long off_xb, off_stat, off_rows, off_hp, off_hps, hps_bytes;
};

static Workspace plan_workspace(const cvlm_encoder_config& c, int batch) {
    const long long g = c.img_size / c.patch_size, T = g * g, M = T * batch;
    const long long D = c.embed_dim, OC = c.out_chans, RP = c.rank_pad;
    const long long kp = static_cast<long long>(c.in_chans) * c.patch_size * c.patch_size;
    Workspace w;
    long long big_cols = c.mlp_dim;
    if (3 * D > big_cols) big_cols = 3 * D;
    if (2 * kp > big_cols) big_cols = 2 * kp;
    if (9 * OC > big_cols) big_cols = 9 * OC;
    w.big_cols = big_cols;
    long long off = 0;
    const long long x_bytes = M * (D * 4 > 2 * OC * 4 ? D * 4 : 2 * OC * 4);
    const long long a_bytes = M * (D > OC ? D : OC) * 2;
    w.off_x = off;     off = align_up(off + x_bytes, 1024);
    w.off_a = off;     off = align_up(off + a_bytes, 1024);
    w.off_big = off;   off = align_up(off + M * big_cols * 2, 1024);
    w.off_hand = off;  off = align_up(off + M * RP * 4, 1024);
    w.off_t = off;     off = align_up(off + M * RP * 2, 1024);
    w.off_gall = off;  off = align_up(off + M * c.depth * RP * 2, 1024);
    w.off_xb = off;    off = align_up(off + M * D * 2, 1024);              // bf16 copy of the residual stream
    w.off_stat = off;  off = align_up(off + M * ((D + CVLM_STAT_GROUP_COLS - 1) / CVLM_STAT_GROUP_COLS) * 8, 1024);   // (sum, sum sq)
    w.off_rows = off;  off = align_up(off + M * 8, 1024);                  // (rstd, -mean * rstd) per row
    // high-pass filter: fp32 result + scratch (bf16 image copy and the half-transformed spectrum)
    const long long img_elems = static_cast<long long>(batch) * c.in_chans * c.img_size * c.img_size;
    w.off_hp = off;    off = align_up(off + img_elems * 4, 1024);
    w.off_hps = off;   off = align_up(off + high_pass_scratch_bytes(batch, c.in_chans, c.img_size), 1024);
    w.hps_bytes = high_pass_scratch_bytes(batch, c.in_chans, c.img_size);
    w.total = off;
    return w;
}

// Activations the training forward keeps for the backward pass (encoder_train.cu), carved out of one caller-provided
// region.  Per image (ViT-H): 105 MB per block (block input bf16 + LayerNorm rows, qkv, attention output, the
// normalised norm2 input, the pre-GELU hidden) + 45 MB of stem / adapter / neck state = 3.4 GB.
Saved plan_saved(const cvlm_encoder_config& c, int batch) {
    const long long g = c.img_size / c.patch_size, T = g * g, M = T * batch;
    const long long D = c.embed_dim, OC = c.out_chans, RP = c.rank_pad;
    const long long kp = static_cast<long long>(c.in_chans) * c.patch_size * c.patch_size;
    Saved s;
    long long off = 0;
    auto take = [&](long long bytes) { const long long o = off; off = align_up(off + bytes, 1024); return o; };
    s.off_x0b = take(M * D * 2);
    s.off_php = take(M * kp * 2);
    s.off_t = take(M * RP * 2);
    s.off_gall = take(M * c.depth * RP * 2);
    s.off_gpre = take(M * c.depth * RP * 2);
    s.off_n1 = take(M * OC * 4);
    s.off_n2 = take(M * OC * 4);
    s.off_blocks = off;
    off = 0;
    s.b_xb = take(M * D * 2);
    s.b_rows1 = take(M * 8);
    s.b_qkv = take(M * 3 * D * 2);
    s.b_o = take(M * D * 2);
    s.b_xh2 = take(M * D * 2);
    s.b_rstd2 = take(M * 4);
    s.b_hpre = take(M * c.mlp_dim * 2);
    s.b_lse = take(M * c.num_heads * 4);
    s.block_stride = off;
    s.off_gtab = s.off_blocks + s.block_stride * c.depth;
    s.gtab_half = align_up(M * c.num_heads * 64 * 4, 1024);
    s.gtab_stride = 2 * s.gtab_half;
    int n_global = 0;
    for (int i = 0; i < c.depth; ++i) n_global += (c.is_global[i] || c.window_size <= 0) ? 1 : 0;
    s.total = s.off_gtab + s.gtab_stride * n_global;
    return s;
}

int forward_impl(cvlm_encoder* enc, const float* images, const float* highpass, int batch, float* out,
                 float* const* interm, int stop_after_block, float* x_out, void* workspace,
                 long long workspace_bytes, void* saved, long long saved_bytes, cudaStream_t stream) {
    if (!enc || !images || !workspace) return set_error(CVLM_ERR_ARG, "forward: null argument");
    if (batch <= 0) return set_error(CVLM_ERR_ARG, "forward: batch must be positive");
    const cvlm_encoder_config& c = enc->cfg;
    const Workspace ws = plan_workspace(c, batch);
    if (workspace_bytes < ws.total)
        return set_error(CVLM_ERR_WORKSPACE, "workspace has %lld bytes, need %lld", workspace_bytes, ws.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error(CVLM_ERR_ARG, "workspace must be 256-byte aligned");

    const int g = c.img_size / c.patch_size, T = g * g, D = c.embed_dim, OC = c.out_chans, RP = c.rank_pad;
    const int hd = D / c.num_heads, depth = c.depth, kp = c.in_chans * c.patch_size * c.patch_size;
    const long long M = static_cast<long long>(T) * batch;
    if (M > 0x7fffffffLL / 2) return set_error(CVLM_ERR_UNSUPPORTED, "forward: batch too large for one call; split it");
    char* wsb = static_cast<char*>(workspace);
    float* x = reinterpret_cast<float*>(wsb + ws.off_x);
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_a);
    __nv_bfloat16* big = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_big);
    float* hand = reinterpret_cast<float*>(wsb + ws.off_hand);
    __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_t);
    __nv_bfloat16* gall = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_gall);
    const long long ldg = static_cast<long long>(depth) * RP;
    float* stat = reinterpret_cast<float*>(wsb + ws.off_stat);
    const int sgroups = (D + CVLM_STAT_GROUP_COLS - 1) / CVLM_STAT_GROUP_COLS;
    // Training forward (saved != NULL): every tensor the backward pass needs goes to its own slot of `saved` instead
    // of a re-used workspace buffer; the schedule and the arithmetic are the inference ones.
    const bool train = saved != nullptr;
    const Saved sv = plan_saved(c, batch);
    char* svb = static_cast<char*>(saved);
    if (train) {
        if (saved_bytes < sv.total) return set_error(CVLM_ERR_WORKSPACE, "saved-activation region has %lld bytes, need %lld", saved_bytes, sv.total);
        if (reinterpret_cast<uintptr_t>(saved) & 255) return set_error(CVLM_ERR_ARG, "saved-activation region must be 256-byte aligned");
        if (stop_after_block >= 0 || x_out) return set_error(CVLM_ERR_ARG, "forward_partial has no training mode");
        t = reinterpret_cast<__nv_bfloat16*>(svb + sv.off_t);
        gall = reinterpret_cast<__nv_bfloat16*>(svb + sv.off_gall);
    }
    auto blk_ptr = [&](int blk, long long off) { return svb + sv.off_blocks + sv.block_stride * blk + off; };
    // norm1 is folded around the GEMMs (cvlm_internal.cuh, GemmArgs): the GEMM that writes block `blk`'s input (lin2 of
    // the previous block, or the prompt add) also leaves bf16(x) and partial row statistics, which ln_finalize reduces
    // to (rstd, -mean * rstd) per row for that block's qkv GEMM
    auto xb_of = [&](int blk) {
        return reinterpret_cast<__nv_bfloat16*>(train ? blk_ptr(blk, sv.b_xb) : wsb + ws.off_xb);
    };
    auto rows_of = [&](int blk) { return reinterpret_cast<float*>(train ? blk_ptr(blk, sv.b_rows1) : wsb + ws.off_rows); };
    // shift_blk >= 0: the rows of block `shift_blk`'s input give the per-row shift of the copy (GemmArgs::shift_rows)
    auto feeds_norm = [&](GemmArgs& ga, int blk, int shift_blk) {
        ga.copy_bf16 = xb_of(blk); ga.ldc = D; ga.stat_out = stat; ga.stat_groups = sgroups;
        ga.shift_rows = shift_blk >= 0 ? rows_of(shift_blk) : nullptr;
    };
    auto normed = [&](GemmArgs& ga, const void* colsum, int blk) {
        ga.ln_rows = rows_of(blk); ga.ln_colsum = static_cast<const float*>(colsum);
    };

    Fwd f{enc, stream};
    const void *patch_w, *patch_b, *pos = nullptr, *hand_w, *hand_b, *emb_wt, *emb_cst, *lw_w, *lw_b, *sh_w, *sh_b;
    CVLM_TRY(f.get("patch.w", 2LL * D * kp, &patch_w));
    CVLM_TRY(f.get("patch.b", 4LL * D, &patch_b));
    if (c.use_abs_pos) CVLM_TRY(f.get("pos", 4LL * T * D, &pos));
    CVLM_TRY(f.get("hand.w", 2LL * c.rank * kp, &hand_w));
    CVLM_TRY(f.get("hand.b", 4LL * RP, &hand_b));
    CVLM_TRY(f.get("emb.wt", 4LL * D * RP, &emb_wt));
    CVLM_TRY(f.get("emb.cst", 4LL * T * RP, &emb_cst));
    CVLM_TRY(f.get("lw.w", 2LL * depth * RP * RP, &lw_w));
    CVLM_TRY(f.get("lw.b", 4LL * depth * RP, &lw_b));
    CVLM_TRY(f.get("shared.w", 2LL * D * RP, &sh_w));
    CVLM_TRY(f.get("shared.b", 4LL * D, &sh_b));

    // ------------------------------------------------------------------ stem
    if (!highpass) {   // PromptGenerator.fft (:332-353) as two GEMMs with the low-pass projector (highpass.cu)
        const void *pcat, *npr, *ppi;
        const long long n = c.img_size;
        const long long np = (n + 63) / 64 * 64;
        CVLM_TRY(f.get("hp.pcat", 4LL * n * n, &pcat));
        CVLM_TRY(f.get("hp.npr", 2LL * n * np, &npr));
        CVLM_TRY(f.get("hp.pi", 2LL * n * np, &ppi));
        float* hp = reinterpret_cast<float*>(wsb + ws.off_hp);
        f.launches += 2;    // cast + G1 + G2: three launches, one counted by run()
        CVLM_TRY(f.run(K_GEMM_OTHER, [&] {
            return high_pass_launch(images, hp, batch, c.in_chans, c.img_size, pcat, npr, ppi, wsb + ws.off_hps, ws.hps_bytes,
                                    stream);
        }));
        highpass = hp;
    }
    __nv_bfloat16* patches = big;
    __nv_bfloat16* patches_hp = train ? reinterpret_cast<__nv_bfloat16*>(svb + sv.off_php) : big + M * kp;
    CVLM_TRY(f.run(K_OTHER, [&] { return im2col_patch_launch(images, patches, batch, c.in_chans, c.img_size, c.patch_size, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return im2col_patch_launch(highpass, patches_hp, batch, c.in_chans, c.img_size, c.patch_size, stream); }));
    {   // PatchEmbed (:655-659) + pos_embed (:139-140): x = patches * Wp^T + bp + pos[row % T]
        GemmArgs ga;
        ga.a1 = patches; ga.lda1 = kp; ga.b1 = patch_w; ga.ldb1 = kp; ga.k1 = kp;
        ga.m = static_cast<int>(M); ga.n = D; ga.bias = static_cast<const float*>(patch_b);
        ga.residual = static_cast<const float*>(pos); ga.ldr = D; ga.res_row_mod = T;
        ga.out = x; ga.ldo = D;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    if (train) {   // the patch embedding WITHOUT pos_embed, bf16: input of embedding_generator, needed for its weight gradient
        GemmArgs ga;
        ga.a1 = patches; ga.lda1 = kp; ga.b1 = patch_w; ga.ldb1 = kp; ga.k1 = kp;
        ga.m = static_cast<int>(M); ga.n = D; ga.bias = static_cast<const float*>(patch_b);
        ga.out = svb + sv.off_x0b; ga.ldo = D; ga.out_bf16 = 1;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    {   // PatchEmbed2 of the high-passed image (:283-285): hand = patches_hp * Wh^T + bh   (fp32 [M, RP])
        GemmArgs ga;
        ga.a1 = patches_hp; ga.lda1 = kp; ga.b1 = hand_w; ga.ldb1 = kp; ga.k1 = kp;
        ga.m = static_cast<int>(M); ga.n = RP; ga.b_rows = c.rank; ga.bias = static_cast<const float*>(hand_b);
        ga.out = hand; ga.ldo = RP;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    CVLM_TRY(f.run(K_OTHER, [&] {
        return adapter_embed_launch(x, hand, static_cast<const float*>(emb_wt), static_cast<const float*>(emb_cst), t,
                                    batch, T, D, c.rank, RP, stream);
    }));
    {   // all 32 lightweight MLPs at once (:291-293): gall = GELU(t * Wlw^T + blw)   (bf16 [M, depth*RP])
        GemmArgs ga;
        ga.a1 = t; ga.lda1 = RP; ga.b1 = lw_w; ga.ldb1 = RP; ga.k1 = RP;
        ga.m = static_cast<int>(M); ga.n = static_cast<int>(ldg); ga.bias = static_cast<const float*>(lw_b);
        ga.out = gall; ga.ldo = ldg; ga.out_bf16 = 1; ga.act_gelu = 1;
        if (train) {   // pre-activation values for the backward of the GELU
            ga.copy_bf16 = svb + sv.off_gpre; ga.ldc = ldg;
        }
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    auto finalize_norm = [&](int blk) -> int {
        return f.run(K_LAYERNORM, [&] { return ln_finalize_launch(stat, sgroups, D, c.ln_eps, rows_of(blk), M, stream); });
    };
    auto prompt_add = [&](int blk, const float* src, float* dst, int shift_blk) -> int {
        // dst = src + gall_blk * Wshared^T + bshared    (shared_mlp :294, add :145); dst is the next norm1 input
        GemmArgs ga;
        ga.a1 = gall + static_cast<long long>(blk) * RP; ga.lda1 = ldg; ga.b1 = sh_w; ga.ldb1 = RP; ga.k1 = RP;
        ga.m = static_cast<int>(M); ga.n = D; ga.bias = static_cast<const float*>(sh_b);
        ga.residual = src; ga.ldr = D; ga.out = dst; ga.ldo = D;
        feeds_norm(ga, blk, shift_blk);
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
        return finalize_norm(blk);
    };
    CVLM_TRY(prompt_add(0, x, x, -1));

    auto finish_partial = [&](const float* src) -> int {
        CVLM_CUDA(cudaMemcpyAsync(x_out, src, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, stream));
        enc->last_launches = f.launches;
        return CVLM_OK;
    };
    if (x_out && stop_after_block < 0) return finish_partial(x);

    // ------------------------------------------------------------------ blocks (Block.forward :430-446)
    float* xcur = x;
    int interm_idx = 0;
    for (int i = 0; i < depth; ++i) {
        const int window = c.is_global[i] ? 0 : c.window_size;
        const int S = window > 0 ? window : g;
        const long long relb = 2LL * (2 * S - 1) * hd;
        int global_idx = 0;                  // this block's index among the global blocks (saved-table slot)
        for (int j = 0; j < i; ++j) global_idx += (c.is_global[j] || c.window_size <= 0) ? 1 : 0;
        const void *qkv_w, *qkv_b, *qkv_c, *relh, *relw, *proj_w, *proj_b, *ln2w, *ln2b, *l1w, *l1b, *l2w, *l2b,
            *l2bf = nullptr;
        CVLM_TRY(f.getb(i, "qkv.w", 2LL * 3 * D * D, &qkv_w));
        CVLM_TRY(f.getb(i, "qkv.b", 4LL * 3 * D, &qkv_b));
        CVLM_TRY(f.getb(i, "qkv.c", 4LL * 3 * D, &qkv_c));
        CVLM_TRY(f.getb(i, "relh", relb, &relh));
        CVLM_TRY(f.getb(i, "relw", relb, &relw));
        CVLM_TRY(f.getb(i, "proj.w", 2LL * D * D, &proj_w));
        CVLM_TRY(f.getb(i, "proj.b", 4LL * D, &proj_b));
        CVLM_TRY(f.getb(i, "ln2.w", 4LL * D, &ln2w));
        CVLM_TRY(f.getb(i, "ln2.b", 4LL * D, &ln2b));
        CVLM_TRY(f.getb(i, "lin1.w", 2LL * c.mlp_dim * D, &l1w));
        CVLM_TRY(f.getb(i, "lin1.b", 4LL * c.mlp_dim, &l1b));
        CVLM_TRY(f.getb(i, "lin2.w", 2LL * D * c.mlp_dim, &l2w));
        CVLM_TRY(f.getb(i, "lin2.b", 4LL * D, &l2b));
        if (i + 1 < depth) CVLM_TRY(f.getb(i, "lin2.bf", 4LL * D, &l2bf));

        __nv_bfloat16* qkv_buf = train ? reinterpret_cast<__nv_bfloat16*>(blk_ptr(i, sv.b_qkv)) : big;
        __nv_bfloat16* attn_out = train ? reinterpret_cast<__nv_bfloat16*>(blk_ptr(i, sv.b_o)) : a;
        {   // norm1 (:436) + qkv (:491) on the raw bf16 residual stream; bias = W beta1 + [b_q, 0, 0]
            GemmArgs ga;
            ga.a1 = xb_of(i); ga.lda1 = D; ga.b1 = qkv_w; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = 3 * D; ga.bias = static_cast<const float*>(qkv_b);
            ga.out = qkv_buf; ga.ldo = 3 * D; ga.out_bf16 = 1;
            normed(ga, qkv_c, i);
            CVLM_TRY(f.gemm(K_GEMM_QKV, ga));
        }
        CVLM_TRY(f.run(window > 0 ? K_ATTN_WINDOW : K_ATTN_GLOBAL, [&] {
            if (!train) return attention_launch(qkv_buf, attn_out, relh, relw, batch, g, window, c.num_heads, hd, stream);
            // the backward kernels need every row's log-sum-exp: the tcgen05 kernels leave it next to the activations
            bool wrote = false, tabs = false;
            float* tw_tab = nullptr;
            float* th_tab = nullptr;
            if (window == 0) {
                tw_tab = reinterpret_cast<float*>(svb + sv.off_gtab + sv.gtab_stride * global_idx);
                th_tab = reinterpret_cast<float*>(svb + sv.off_gtab + sv.gtab_stride * global_idx + sv.gtab_half);
            }
            const int rc = attention_launch(qkv_buf, attn_out, relh, relw, batch, g, window, c.num_heads, hd, stream,
                                            reinterpret_cast<float*>(blk_ptr(i, sv.b_lse)), &wrote, tw_tab, th_tab, &tabs);
            enc->lse_saved[i] = wrote ? 1 : 0;
            enc->tabs_saved[i] = tabs ? 1 : 0;
            return rc;
        }));
        {   // proj (:502) + shortcut (:443); bias = b_proj + W_proj * b_v
            GemmArgs ga;
            ga.a1 = attn_out; ga.lda1 = D; ga.b1 = proj_w; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = D; ga.bias = static_cast<const float*>(proj_b);
            ga.residual = xcur; ga.ldr = D; ga.out = xcur; ga.ldo = D;
            CVLM_TRY(f.gemm(K_GEMM_PROJ, ga));
        }
        // norm2 (:444) stays a kernel: proj (HBM-bound residual epilogue) and lin1 (GELU epilogue) are the two GEMMs
        // that cannot absorb the producer / consumer halves of the fold for free (measured: +5.2 ms and +6.6 ms per step)
        CVLM_TRY(f.run(K_LAYERNORM, [&] {
            if (train)   // also keeps the normalised input (bf16) and 1/std per row for the backward of norm2
                return layernorm_train_launch(xcur, static_cast<const float*>(ln2w), static_cast<const float*>(ln2b), a,
                                              blk_ptr(i, sv.b_xh2), reinterpret_cast<float*>(blk_ptr(i, sv.b_rstd2)), M, D,
                                              c.ln_eps, stream);
            return layernorm_launch(xcur, static_cast<const float*>(ln2w), static_cast<const float*>(ln2b), a, 1, M, D,
                                    c.ln_eps, stream);
        }));
        {   // lin1 + GELU (common.py:26)
            GemmArgs ga;
            ga.a1 = a; ga.lda1 = D; ga.b1 = l1w; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = c.mlp_dim; ga.bias = static_cast<const float*>(l1b);
            ga.out = big; ga.ldo = c.mlp_dim; ga.out_bf16 = 1; ga.act_gelu = 1;
            if (train) {   // pre-GELU hidden for the backward pass
                ga.copy_bf16 = blk_ptr(i, sv.b_hpre); ga.ldc = c.mlp_dim;
            }
            CVLM_TRY(f.gemm(K_GEMM_LIN1, ga));
        }
        const bool want_interm = interm != nullptr && c.is_global[i];
        const bool stop_here = x_out != nullptr && stop_after_block == i;
        const bool last = i + 1 == depth;
        const bool fold = !last && !want_interm && !stop_here;
        float* dst = want_interm ? interm[interm_idx] : xcur;
        {   // lin2 + shortcut (:444) [+ prompt_{i+1} as a second K segment when nothing needs x before it]
            GemmArgs ga;
            ga.a1 = big; ga.lda1 = c.mlp_dim; ga.b1 = l2w; ga.ldb1 = c.mlp_dim; ga.k1 = c.mlp_dim;
            if (fold) {
                ga.a2 = gall + static_cast<long long>(i + 1) * RP; ga.lda2 = ldg; ga.b2 = sh_w; ga.ldb2 = RP; ga.k2 = RP;
            }
            ga.m = static_cast<int>(M); ga.n = D; ga.bias = static_cast<const float*>(fold ? l2bf : l2b);
            ga.residual = xcur; ga.ldr = D; ga.out = dst; ga.ldo = D;
            if (fold) feeds_norm(ga, i + 1, i);   // otherwise prompt_add below writes the next norm1 input
            if (last) {   // the neck's 1x1 conv reads bf16(x): let this epilogue write it (unshifted producer copy; the
                          // statistics land in the scratch and are not used) instead of a separate cast pass
                ga.copy_bf16 = a; ga.ldc = D; ga.stat_out = stat; ga.stat_groups = sgroups;
            }
            CVLM_TRY(f.gemm(K_GEMM_LIN2, ga));
            if (fold) CVLM_TRY(finalize_norm(i + 1));
        }
        if (want_interm) ++interm_idx;
        if (stop_here) return finish_partial(dst);
        if (!fold && !last) {
            CVLM_TRY(prompt_add(i + 1, dst, x, i));
            xcur = x;
        } else {
            xcur = dst;
        }
    }

    // ------------------------------------------------------------------ neck (:98-114, :150)
    if (!out) return set_error(CVLM_ERR_ARG, "forward: null output");
    const void *c1w, *n1w, *n1b, *c2w, *n2w, *n2b;
    CVLM_TRY(f.get("neck.conv1.w", 2LL * OC * D, &c1w));
    CVLM_TRY(f.get("neck.ln1.w", 4LL * OC, &n1w));
    CVLM_TRY(f.get("neck.ln1.b", 4LL * OC, &n1b));
    CVLM_TRY(f.get("neck.conv2.w", 2LL * OC * 9 * OC, &c2w));
    CVLM_TRY(f.get("neck.ln2.w", 4LL * OC, &n2w));
    CVLM_TRY(f.get("neck.ln2.b", 4LL * OC, &n2b));
    // `a` holds bf16(x) already (written by the last lin2); n1 / n2 live in the x region, which is dead by now unless the
    // final residual stream is x itself -- and then nothing reads it any more either.
    float* n1 = train ? reinterpret_cast<float*>(svb + sv.off_n1) : x;
    float* n2 = train ? reinterpret_cast<float*>(svb + sv.off_n2) : x + M * OC;
    {   // 1x1 conv, no bias
        GemmArgs ga;
        ga.a1 = a; ga.lda1 = D; ga.b1 = c1w; ga.ldb1 = D; ga.k1 = D;
        ga.m = static_cast<int>(M); ga.n = OC; ga.out = n1; ga.ldo = OC;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    // LayerNorm2d over channels == LayerNorm over the last dim of NHWC
    CVLM_TRY(f.run(K_LAYERNORM, [&] {
        return layernorm_launch(n1, static_cast<const float*>(n1w), static_cast<const float*>(n1b), a, 1, M, OC,
                                c.ln2d_eps, stream);
    }));
    CVLM_TRY(f.run(K_OTHER, [&] { return im2col_3x3_launch(a, big, batch, g, OC, stream); }));
    {   // 3x3 conv as a GEMM over the 9 taps, no bias
        GemmArgs ga;
        ga.a1 = big; ga.lda1 = 9LL * OC; ga.b1 = c2w; ga.ldb1 = 9LL * OC; ga.k1 = 9 * OC;
        ga.m = static_cast<int>(M); ga.n = OC; ga.out = n2; ga.ldo = OC;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    CVLM_TRY(f.run(K_OTHER, [&] {
        return layernorm_nchw_out_launch(n2, static_cast<const float*>(n2w), static_cast<const float*>(n2b), out, batch,
                                         T, OC, c.ln2d_eps, stream);
    }));
    enc->last_launches = f.launches;
    return CVLM_OK;
}

}  // namespace cvlm

// ====================================================================================== C ABI
using namespace cvlm;

extern "C" {

const char* cvlm_last_error(void) { return g_err; }
int cvlm_abi_version(void) { return 1; }

int cvlm_set_option(const char* name, int value) {
    if (!name) return set_error(CVLM_ERR_ARG, "set_option: null name");
    option(OPT_GEMM_MODE);                               // make sure the environment defaults are in place first
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, kOptionNames[i])) {
            __atomic_store_n(&g_options[i], value, __ATOMIC_RELAXED);
            return CVLM_OK;
        }
    return set_error(CVLM_ERR_ARG, "set_option: unknown option '%s'", name);
}

int cvlm_get_option(const char* name, int* value) {
    if (!name || !value) return set_error(CVLM_ERR_ARG, "get_option: null argument");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, kOptionNames[i])) {
            *value = option(static_cast<Option>(i));
            return CVLM_OK;
        }
    return set_error(CVLM_ERR_ARG, "get_option: unknown option '%s'", name);
}

int cvlm_encoder_create(const cvlm_encoder_config* cfg, cvlm_encoder** out) {
    if (!cfg || !out) return set_error(CVLM_ERR_ARG, "create: null argument");
    const cvlm_encoder_config& c = *cfg;
    if (c.patch_size != 16 || c.img_size <= 0 || c.img_size % 16)
        return set_error(CVLM_ERR_UNSUPPORTED, "create: patch_size must be 16 and img_size a multiple of 16");
    if (c.depth <= 0 || c.depth > CVLM_MAX_DEPTH) return set_error(CVLM_ERR_UNSUPPORTED, "create: depth out of range");
    if (c.num_heads <= 0 || c.embed_dim % c.num_heads) return set_error(CVLM_ERR_ARG, "create: embed_dim %% num_heads != 0");
    const int hd = c.embed_dim / c.num_heads;
    if (hd != 32 && hd != 64 && hd != 80)
        return set_error(CVLM_ERR_UNSUPPORTED, "create: head_dim %d (supported: 32, 64, 80)", hd);
    if (c.embed_dim % 8 || c.mlp_dim % 8 || c.out_chans % 8)
        return set_error(CVLM_ERR_UNSUPPORTED, "create: embed_dim, mlp_dim, out_chans must be multiples of 8");
    if (c.rank <= 0 || c.rank_pad < c.rank || c.rank_pad % 8 || c.rank_pad > 40)
        return set_error(CVLM_ERR_UNSUPPORTED, "create: adapter rank %d / pad %d unsupported", c.rank, c.rank_pad);
    if (c.img_size / 16 > 64) return set_error(CVLM_ERR_UNSUPPORTED, "create: token grid larger than 64x64");
    if (c.window_size < 0) return set_error(CVLM_ERR_ARG, "create: negative window_size");
    cvlm_encoder* e = new cvlm_encoder();
    e->cfg = c;
    *out = e;
    return CVLM_OK;
}

int cvlm_encoder_destroy(cvlm_encoder* enc) {
    if (enc) {
        for (const ProfSpan& sp : enc->spans) {
            cudaEventDestroy(sp.start);
            cudaEventDestroy(sp.stop);
        }
        for (cudaEvent_t e : enc->event_pool) cudaEventDestroy(e);
        for (cudaEvent_t e : enc->host_events) cudaEventDestroy(e);
    }
    delete enc;
    return CVLM_OK;
}

int cvlm_encoder_set_tensor(cvlm_encoder* enc, const char* name, const void* dev_ptr, long long nbytes) {
    if (!enc || !name || !dev_ptr || nbytes <= 0) return set_error(CVLM_ERR_ARG, "set_tensor: bad argument");
    if (reinterpret_cast<uintptr_t>(dev_ptr) & 15) return set_error(CVLM_ERR_ARG, "set_tensor: '%s' is not 16-byte aligned", name);
    enc->tensors[name] = TensorRef{dev_ptr, nbytes};
    return CVLM_OK;
}

long long cvlm_encoder_workspace_bytes(const cvlm_encoder* enc, int batch) {
    if (!enc || batch <= 0) return set_error(CVLM_ERR_ARG, "workspace_bytes: bad argument");
    return plan_workspace(enc->cfg, batch).total;
}

int cvlm_encoder_forward(cvlm_encoder* enc, const float* images, const float* highpass, int batch, float* out,
                         float* const* interm, void* workspace, long long workspace_bytes, cvlm_stream_t stream) {
    return forward_impl(enc, images, highpass, batch, out, interm, -1, nullptr, workspace, workspace_bytes, nullptr, 0,
                        static_cast<cudaStream_t>(stream));
}

int cvlm_encoder_forward_partial(cvlm_encoder* enc, const float* images, const float* highpass, int batch,
                                 int stop_after_block, float* x_out, void* workspace, long long workspace_bytes,
                                 cvlm_stream_t stream) {
    if (!x_out) return set_error(CVLM_ERR_ARG, "forward_partial: null x_out");
    return forward_impl(enc, images, highpass, batch, nullptr, nullptr, stop_after_block, x_out, workspace,
                        workspace_bytes, nullptr, 0, static_cast<cudaStream_t>(stream));
}

// Host-buffer entry point: what a caller with a batch in (pinned) host memory runs.  The batch is cut into chunks; the
// host -> device copy of chunk i+1 and the device -> host copy of the embeddings of chunk i-1 run on `copy_stream`
// underneath the encoder of chunk i on `stream` (two device staging buffers).  Only stream-ordered waits, no host sync.
int cvlm_encoder_forward_host(cvlm_encoder* enc, const float* images_host, int batch, int chunk, float* out_host,
                              float* emb_dev, float* const* interm_dev, float* stage_dev, void* workspace,
                              long long workspace_bytes, cvlm_stream_t copy_stream_, cvlm_stream_t stream_) {
    if (!enc || !images_host || !out_host || !emb_dev || !stage_dev || !workspace)
        return set_error(CVLM_ERR_ARG, "forward_host: null argument");
    if (batch <= 0 || chunk <= 0) return set_error(CVLM_ERR_ARG, "forward_host: batch and chunk must be positive");
    cudaStream_t cs = static_cast<cudaStream_t>(copy_stream_), main = static_cast<cudaStream_t>(stream_);
    if (cs == main) return set_error(CVLM_ERR_ARG, "forward_host: the copy stream must differ from the compute stream");
    const cvlm_encoder_config& c = enc->cfg;
    const long long g = c.img_size / c.patch_size, T = g * g;
    const long long img_elems = static_cast<long long>(c.in_chans) * c.img_size * c.img_size;
    const long long emb_elems = static_cast<long long>(c.out_chans) * T;
    int n_glob = 0;
    for (int i = 0; i < c.depth; ++i) n_glob += c.is_global[i] ? 1 : 0;
    while (enc->host_events.size() < 5) {
        cudaEvent_t e = nullptr;
        CVLM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        enc->host_events.push_back(e);
    }
    cudaEvent_t ready[2] = {enc->host_events[0], enc->host_events[1]}, freed[2] = {enc->host_events[2], enc->host_events[3]};
    cudaEvent_t edge = enc->host_events[4];
    // The copy stream does NOT wait for what is already queued on `stream`: a staging half only has to be free, i.e. the
    // encoder that last read it must have run (freed[], kept across calls; waiting on a never-recorded event is a no-op).
    // Back-to-back calls therefore upload their first chunk underneath the previous call's last encoder.  images_host must
    // be complete when the call is made (it is host memory: no stream orders it).
    const int n_chunks = (batch + chunk - 1) / chunk;
    auto upload = [&](int i) -> int {
        const int s = i * chunk, e = s + chunk < batch ? s + chunk : batch;
        CVLM_CUDA(cudaStreamWaitEvent(cs, freed[i & 1], 0));
        CVLM_CUDA(cudaMemcpyAsync(stage_dev + static_cast<long long>(i & 1) * chunk * img_elems, images_host + s * img_elems,
                                  static_cast<size_t>(e - s) * img_elems * 4, cudaMemcpyHostToDevice, cs));
        CVLM_CUDA(cudaEventRecord(ready[i & 1], cs));
        return CVLM_OK;
    };
    CVLM_TRY(upload(0));
    long long launches = 0;
    std::vector<float*> islice(static_cast<size_t>(n_glob > 0 ? n_glob : 1), nullptr);
    for (int i = 0; i < n_chunks; ++i) {
        const int s = i * chunk, e = s + chunk < batch ? s + chunk : batch;
        if (i + 1 < n_chunks) CVLM_TRY(upload(i + 1));
        CVLM_CUDA(cudaStreamWaitEvent(main, ready[i & 1], 0));
        if (interm_dev)
            for (int k = 0; k < n_glob; ++k) islice[k] = interm_dev[k] + static_cast<long long>(s) * T * c.embed_dim;
        CVLM_TRY(forward_impl(enc, stage_dev + static_cast<long long>(i & 1) * chunk * img_elems, nullptr, e - s,
                              emb_dev + s * emb_elems, interm_dev ? islice.data() : nullptr, -1, nullptr, workspace,
                              workspace_bytes, nullptr, 0, main));
        launches += enc->last_launches;
        CVLM_CUDA(cudaEventRecord(freed[i & 1], main));
        // this chunk's embeddings leave under the next chunk's encoder; the last chunk's go out on `stream` itself, so that
        // the copy stream is idle again while the last encoder runs (the next call's first upload goes there)
        cudaStream_t ds = i + 1 < n_chunks ? cs : main;
        if (ds == cs) CVLM_CUDA(cudaStreamWaitEvent(cs, freed[i & 1], 0));
        CVLM_CUDA(cudaMemcpyAsync(out_host + s * emb_elems, emb_dev + s * emb_elems, static_cast<size_t>(e - s) * emb_elems * 4,
                                  cudaMemcpyDeviceToHost, ds));
    }
    CVLM_CUDA(cudaEventRecord(edge, cs));
    CVLM_CUDA(cudaStreamWaitEvent(main, edge, 0));          // out_host is valid once `stream` has been synchronised
    enc->last_launches = launches;
    return CVLM_OK;
}

long long cvlm_encoder_last_launch_count(const cvlm_encoder* enc) { return enc ? enc->last_launches : 0; }

int cvlm_encoder_set_profiling(cvlm_encoder* enc, int enable) {
    if (!enc) return set_error(CVLM_ERR_ARG, "set_profiling: null handle");
    enc->profiling = enable != 0;
    return CVLM_OK;
}

int cvlm_encoder_get_profile(cvlm_encoder* enc, float* ms_by_kind, int* launches_by_kind) {
    if (!enc || !ms_by_kind || !launches_by_kind) return set_error(CVLM_ERR_ARG, "get_profile: null argument");
    for (int k = 0; k < K_COUNT; ++k) {
        ms_by_kind[k] = 0.f;
        launches_by_kind[k] = 0;
    }
    for (const ProfSpan& sp : enc->spans) {
        CVLM_CUDA(cudaEventSynchronize(sp.stop));
        float ms = 0.f;
        CVLM_CUDA(cudaEventElapsedTime(&ms, sp.start, sp.stop));
        ms_by_kind[sp.kind] += ms;
        launches_by_kind[sp.kind] += 1;
        enc->event_pool.push_back(sp.start);
        enc->event_pool.push_back(sp.stop);
    }
    enc->spans.clear();
    return CVLM_OK;
}

int cvlm_op_gemm_bf16(const void* a1, long long lda1, const void* b1, long long ldb1, int k1, const void* a2,
                      long long lda2, const void* b2, long long ldb2, int k2, int m, int n, int b_rows,
                      const float* bias, const float* residual, long long ldr, int res_row_mod, void* out,
                      long long ldo, int out_bf16, int act_gelu, cvlm_stream_t stream) {
    GemmArgs ga;
    ga.a1 = a1; ga.lda1 = lda1; ga.b1 = b1; ga.ldb1 = ldb1; ga.k1 = k1;
    ga.a2 = a2; ga.lda2 = lda2; ga.b2 = b2; ga.ldb2 = ldb2; ga.k2 = k2;
    ga.m = m; ga.n = n; ga.b_rows = b_rows; ga.bias = bias; ga.residual = residual; ga.ldr = ldr;
    ga.res_row_mod = res_row_mod; ga.out = out; ga.ldo = ldo; ga.out_bf16 = out_bf16; ga.act_gelu = act_gelu;
    return gemm_launch(ga, static_cast<cudaStream_t>(stream));
}

int cvlm_op_gemm_bf16_ln(const void* a1, long long lda1, const void* b1, long long ldb1, int k1, const void* a2,
                         long long lda2, const void* b2, long long ldb2, int k2, int m, int n, int b_rows,
                         const float* bias, const float* residual, long long ldr, int res_row_mod, void* out,
                         long long ldo, int out_bf16, int act_gelu, void* copy_bf16, long long ldc, float* stat_out,
                         const float* ln_rows, const float* ln_colsum, cvlm_stream_t stream) {
    GemmArgs ga;
    ga.a1 = a1; ga.lda1 = lda1; ga.b1 = b1; ga.ldb1 = ldb1; ga.k1 = k1;
    ga.a2 = a2; ga.lda2 = lda2; ga.b2 = b2; ga.ldb2 = ldb2; ga.k2 = k2;
    ga.m = m; ga.n = n; ga.b_rows = b_rows; ga.bias = bias; ga.residual = residual; ga.ldr = ldr;
    ga.res_row_mod = res_row_mod; ga.out = out; ga.ldo = ldo; ga.out_bf16 = out_bf16; ga.act_gelu = act_gelu;
    ga.copy_bf16 = copy_bf16; ga.ldc = ldc; ga.stat_out = stat_out; ga.stat_groups = stat_out ? (n + CVLM_STAT_GROUP_COLS - 1) / CVLM_STAT_GROUP_COLS : 0;
    ga.ln_rows = ln_rows; ga.ln_colsum = ln_colsum;
    return gemm_launch(ga, static_cast<cudaStream_t>(stream));
}

int cvlm_op_ln_finalize(const float* partial, int groups, int width, float eps, float* rows_out, long long rows,
                        cvlm_stream_t stream) {
    return ln_finalize_launch(partial, groups, width, eps, rows_out, rows, static_cast<cudaStream_t>(stream));
}

long long cvlm_op_preprocess_u8_scratch_bytes(int h, int w, int out_size) {
    if (h <= 0 || w <= 0 || out_size <= 0) return set_error(CVLM_ERR_ARG, "preprocess_u8_scratch_bytes: bad argument");
    return preprocess_scratch_bytes(h, w, out_size);
}

int cvlm_op_preprocess_u8(const unsigned char* img, int h, int w, float* out, int out_size, const float* mean,
                          const float* std, void* scratch, long long scratch_bytes, cvlm_stream_t stream) {
    return preprocess_u8_launch(img, h, w, out, out_size, mean, std, scratch, scratch_bytes, static_cast<cudaStream_t>(stream));
}

long long cvlm_op_high_pass_scratch_bytes(int batch, int chans, int img_size) {
    if (batch <= 0 || chans <= 0 || img_size <= 0) return set_error(CVLM_ERR_ARG, "high_pass_scratch_bytes: bad argument");
    return high_pass_scratch_bytes(batch, chans, img_size);
}

int cvlm_op_high_pass(const float* images, float* out, int batch, int chans, int img_size, const void* pcat,
                      const void* npr, const void* pi, void* scratch, long long scratch_bytes, cvlm_stream_t stream) {
    return high_pass_launch(images, out, batch, chans, img_size, pcat, npr, pi, scratch, scratch_bytes,
                            static_cast<cudaStream_t>(stream));
}

int cvlm_op_attention(const void* qkv, void* out, const void* rel_pos_h, const void* rel_pos_w, int batch,
                      int grid_side, int window, int heads, int head_dim, cvlm_stream_t stream) {
    return attention_launch(qkv, out, rel_pos_h, rel_pos_w, batch, grid_side, window, heads, head_dim,
                            static_cast<cudaStream_t>(stream));
}

int cvlm_op_attention_plain(const void* qkv, void* out, int batch, int tokens, int heads, int head_dim,
                            cvlm_stream_t stream) {
    if (!qkv || !out) return set_error(CVLM_ERR_ARG, "attention_plain: null pointer");
    return attention_plain_tc2_launch(qkv, out, batch, tokens, heads, head_dim, static_cast<cudaStream_t>(stream));
}

int cvlm_op_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_bf16, long long rows,
                      int cols, float eps, cvlm_stream_t stream) {
    return layernorm_launch(x, gamma, beta, out, out_bf16, rows, cols, eps, static_cast<cudaStream_t>(stream));
}

int cvlm_op_layernorm2d_nchw(const float* x, const float* gamma, const float* beta, float* out_nchw, int batch,
                             int tokens, int chans, float eps, cvlm_stream_t stream) {
    return layernorm_nchw_out_launch(x, gamma, beta, out_nchw, batch, tokens, chans, eps,
                                     static_cast<cudaStream_t>(stream));
}

int cvlm_op_im2col_patch16(const float* img, void* out, int batch, int chans, int img_size, cvlm_stream_t stream) {
    return im2col_patch_launch(img, out, batch, chans, img_size, 16, static_cast<cudaStream_t>(stream));
}

int cvlm_op_im2col_3x3(const void* x_nhwc, void* out, int batch, int grid_side, int chans, cvlm_stream_t stream) {
    return im2col_3x3_launch(x_nhwc, out, batch, grid_side, chans, static_cast<cudaStream_t>(stream));
}

int cvlm_op_cast_bf16(const float* x, void* out, long long n, cvlm_stream_t stream) {
    return cast_bf16_launch(x, out, n, static_cast<cudaStream_t>(stream));
}

int cvlm_op_adapter_embed(const float* x0, const float* hand, const float* w_t, const float* cst, void* t_out,
                          int batch, int tokens, int dim, int rank, int rank_pad, cvlm_stream_t stream) {
    return adapter_embed_launch(x0, hand, w_t, cst, t_out, batch, tokens, dim, rank, rank_pad,
                                static_cast<cudaStream_t>(stream));
}

}  // extern "C"
