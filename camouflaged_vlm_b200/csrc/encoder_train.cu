// Adapter-only training of the encoder: the reference trains `prompt_generator.*` with the whole trunk
// frozen (train_ovcos_maskdecoder_edge.py:301-303), so the backward pass needs the gradient of the loss with respect to
// the input of EVERY block (prompt_i is added there, image_encoder.py:145) but no trunk weight gradient.
//
//   cvlm_encoder_forward_train   the inference schedule (encoder.cu) that also keeps, per block, what the backward pass
//                                needs (plan_saved): block input bf16 + LayerNorm rows, qkv, attention output, the
//                                normalised norm2 input, the pre-GELU hidden; plus stem / adapter / neck state.
//   cvlm_encoder_backward        d(out) [, d(interm_k)] -> gradients of the 6 adapter tensor groups.
//
// Per block (forward  u = LN1(x); qkv = u Wqkv; o = Attn(qkv); x1 = x + o Wp; v = LN2(x1); h = v W1; x2 = x1 + GELU(h) W2):
//   dh  = (dx2 W2) * gelu'(h)        tcgen05 GEMM with the gelu' epilogue          [M, 4D] bf16
//   dv  = dh (W1 gamma2)             tcgen05 GEMM, fp32 out (gamma folded into the transposed weight copy)
//   dx1 = dx2 + LN2'(dv)             ln_bwd (saved xhat, rstd)
//   do  = dx1 Wp                     tcgen05 GEMM, bf16 out
//   dqkv = Attn'(qkv, o, do)         attention_bwd.cu
//   du  = dqkv (Wqkv gamma1)         tcgen05 GEMM, fp32 out
//   dx  = dx1 + LN1'(du)             ln_bwd (saved bf16 x, (rstd, -mean rstd) rows)  = d(prompt_i) = d(output of block i-1)
// Adapter (image_encoder.py:278-296; prompt_i = shared_mlp(GELU(lightweight_mlp_i(hand + emb)))):
//   dW_shared += dx^T G_i (+ bias via a ones column)            wgrad_tn, per block
//   dpre_i = (dx W_shared) * gelu'(pre_i)                        tcgen05 GEMM (N = rank) with the gelu' epilogue
//   dt = sum_i dpre_i W_lw_i = dpre_all W_lw(stacked)            one tcgen05 GEMM after the loop
//   dW_lw = dpre_all^T t,  dW_hand = patches_hp^T dt,  db_hand = db_eg = colsum(dt)      wgrad_tn
//   dW_eg^T = sum_b X0_b dt_b  with X0_b the image's patch embedding re-read as [D, T] (the reference's scramble,
//             image_encoder.py:278-281): eg_wgrad (fp32 FMAs; 0.2 % of the step's FLOPs).
#include "cvlm_internal.cuh"
#include "encoder_runtime.cuh"

namespace cvlm {

namespace {

struct BwdWs {
    long long off_dx, off_dxb, off_g, off_big, off_do, off_lse, off_dsum, off_dpre, off_dt, off_p_eg, off_dn, off_dnb;
    long long off_p_shared, off_p_lw, off_p_hand, off_p_dt, total;
    int splits, jp;
};

BwdWs plan_bwd(const cvlm_encoder_config& c, int batch) {
    const long long g = c.img_size / c.patch_size, T = g * g, M = T * batch;
    const long long D = c.embed_dim, OC = c.out_chans, RP = c.rank_pad;
    const long long kp = static_cast<long long>(c.in_chans) * c.patch_size * c.patch_size;
    BwdWs w;
    long long off = 0;
    auto take = [&](long long bytes) { const long long o = off; off = align_up(off + bytes, 1024); return o; };
    long long big_cols = c.mlp_dim;
    if (3 * D > big_cols) big_cols = 3 * D;
    if (9 * OC > big_cols) big_cols = 9 * OC;
    w.splits = wgrad_splits(M);
    w.jp = wgrad_jp(static_cast<int>(RP), 1);
    w.off_dx = take(M * D * 4);
    w.off_dxb = take(M * D * 2);
    w.off_g = take(M * D * 4);
    w.off_big = take(M * big_cols * 2);
    w.off_do = take(M * D * 2);
    w.off_lse = take(M * c.num_heads * 4);
    w.off_dsum = take(M * c.num_heads * 4);
    w.off_dpre = take(M * c.depth * RP * 2);
    w.off_dt = take(M * RP * 2);
    w.off_p_eg = take(static_cast<long long>(batch) * D * RP * 4);
    w.off_dn = take(M * OC * 4);
    w.off_dnb = take(M * OC * 2);
    w.off_p_shared = take(static_cast<long long>(w.splits) * D * w.jp * 4);
    w.off_p_lw = take(static_cast<long long>(w.splits) * c.depth * RP * w.jp * 4);
    w.off_p_hand = take(static_cast<long long>(w.splits) * kp * w.jp * 4);
    w.off_p_dt = take(static_cast<long long>(w.splits) * RP * w.jp * 4);
    w.total = off;
    return w;
}

// fp32 gradient buffer handed back to the caller: [shared | lw | hand | dtsum | eg]
struct GradLayout {
    long long off_shared, off_lw, off_hand, off_dtsum, off_eg, total;   // in floats
    int jp;
};
GradLayout plan_grads(const cvlm_encoder_config& c) {
    const long long D = c.embed_dim, RP = c.rank_pad;
    const long long kp = static_cast<long long>(c.in_chans) * c.patch_size * c.patch_size;
    GradLayout g;
    g.jp = wgrad_jp(static_cast<int>(RP), 1);
    long long off = 0;
    g.off_shared = off; off += D * g.jp;                 // [D, jp]: cols < RP = dW_shared[d, r], col RP = db_shared[d]
    g.off_lw = off;     off += c.depth * RP * g.jp;      // [depth*RP, jp]: row i*RP+s, col r = dW_lw_i[s, r], col RP = db_lw_i[s]
    g.off_hand = off;   off += kp * g.jp;                // [kp, jp]: row k, col r = dW_hand[r, k]
    g.off_dtsum = off;  off += RP * g.jp;                // [RP, jp]: col RP = sum_m dt[m, r] = db_hand[r] = db_eg[r]
    g.off_eg = off;     off += D * RP;                   // [D, RP] = dW_eg^T
    g.total = off;
    return g;
}

int backward_impl(cvlm_encoder* enc, int batch, const float* dout_nhwc, const float* const* dinterm, const void* saved,
                  long long saved_bytes, float* grads, long long grads_floats, void* workspace, long long workspace_bytes,
                  cudaStream_t stream) {
    if (!enc || !dout_nhwc || !saved || !grads || !workspace) return set_error(CVLM_ERR_ARG, "backward: null argument");
    if (batch <= 0) return set_error(CVLM_ERR_ARG, "backward: batch must be positive");
    const cvlm_encoder_config& c = enc->cfg;
    const Saved sv = plan_saved(c, batch);
    const BwdWs ws = plan_bwd(c, batch);
    const GradLayout gl = plan_grads(c);
    if (saved_bytes < sv.total) return set_error(CVLM_ERR_WORKSPACE, "saved-activation region has %lld bytes, need %lld", saved_bytes, sv.total);
    if (workspace_bytes < ws.total) return set_error(CVLM_ERR_WORKSPACE, "backward workspace has %lld bytes, need %lld", workspace_bytes, ws.total);
    if (grads_floats < gl.total) return set_error(CVLM_ERR_WORKSPACE, "gradient buffer has %lld floats, need %lld", grads_floats, gl.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(saved) & 255) ||
        (reinterpret_cast<uintptr_t>(grads) & 15))
        return set_error(CVLM_ERR_ARG, "backward: workspace / saved region must be 256-byte aligned, gradients 16-byte");
    const int g = c.img_size / c.patch_size, T = g * g, D = c.embed_dim, OC = c.out_chans, RP = c.rank_pad;
    const int hd = D / c.num_heads, depth = c.depth, kp = c.in_chans * c.patch_size * c.patch_size;
    const long long M = static_cast<long long>(T) * batch;
    const long long ldg = static_cast<long long>(depth) * RP;
    const char* svb = static_cast<const char*>(saved);
    char* wsb = static_cast<char*>(workspace);
    float* dx = reinterpret_cast<float*>(wsb + ws.off_dx);
    __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_dxb);
    float* gbuf = reinterpret_cast<float*>(wsb + ws.off_g);
    __nv_bfloat16* big = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_big);
    __nv_bfloat16* dob = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_do);
    float* lse = reinterpret_cast<float*>(wsb + ws.off_lse);
    float* dsum = reinterpret_cast<float*>(wsb + ws.off_dsum);
    __nv_bfloat16* dpre = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_dpre);
    __nv_bfloat16* dt = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_dt);
    float* p_eg = reinterpret_cast<float*>(wsb + ws.off_p_eg);
    float* dn = reinterpret_cast<float*>(wsb + ws.off_dn);
    __nv_bfloat16* dnb = reinterpret_cast<__nv_bfloat16*>(wsb + ws.off_dnb);
    float* p_shared = reinterpret_cast<float*>(wsb + ws.off_p_shared);
    float* p_lw = reinterpret_cast<float*>(wsb + ws.off_p_lw);
    float* p_hand = reinterpret_cast<float*>(wsb + ws.off_p_hand);
    float* p_dt = reinterpret_cast<float*>(wsb + ws.off_p_dt);
    auto blk_ptr = [&](int blk, long long off) { return svb + sv.off_blocks + sv.block_stride * blk + off; };
    const __nv_bfloat16* gall = reinterpret_cast<const __nv_bfloat16*>(svb + sv.off_gall);
    const __nv_bfloat16* gpre = reinterpret_cast<const __nv_bfloat16*>(svb + sv.off_gpre);

    Fwd f{enc, stream};
    const void *c1wt, *c2wt, *n1w, *n2w, *sh_wt, *lw_wt;
    CVLM_TRY(f.get("neck.conv1.wt", 2LL * D * OC, &c1wt));
    CVLM_TRY(f.get("neck.conv2.wt", 2LL * OC * 9 * OC, &c2wt));
    CVLM_TRY(f.get("neck.ln1.w", 4LL * OC, &n1w));
    CVLM_TRY(f.get("neck.ln2.w", 4LL * OC, &n2w));
    CVLM_TRY(f.get("shared.wt", 2LL * RP * D, &sh_wt));
    CVLM_TRY(f.get("lw.wt", 2LL * RP * ldg, &lw_wt));

    // ------------------------------------------------------------------ neck (image_encoder.py:98-114), backwards
    CVLM_TRY(f.run(K_LAYERNORM, [&] {
        return ln_bwd_launch(2, dout_nhwc, static_cast<const float*>(n2w), svb + sv.off_n2, nullptr, nullptr, dn, dnb, M, OC,
                             c.ln2d_eps, stream);
    }));
    CVLM_TRY(f.run(K_OTHER, [&] { return im2col_3x3_launch(dnb, big, batch, g, OC, stream); }));
    {   // 3x3 conv backward = 3x3 conv with the flipped, transposed kernel (packed as neck.conv2.wt)
        GemmArgs ga;
        ga.a1 = big; ga.lda1 = 9LL * OC; ga.b1 = c2wt; ga.ldb1 = 9LL * OC; ga.k1 = 9 * OC;
        ga.m = static_cast<int>(M); ga.n = OC; ga.out = gbuf; ga.ldo = OC;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    CVLM_TRY(f.run(K_LAYERNORM, [&] {
        return ln_bwd_launch(2, gbuf, static_cast<const float*>(n1w), svb + sv.off_n1, nullptr, nullptr, dn, dnb, M, OC,
                             c.ln2d_eps, stream);
    }));
    {   // 1x1 conv backward: dx = dn Wc1  (+ bf16 copy for the first block GEMM)
        GemmArgs ga;
        ga.a1 = dnb; ga.lda1 = OC; ga.b1 = c1wt; ga.ldb1 = OC; ga.k1 = OC;
        ga.m = static_cast<int>(M); ga.n = D; ga.out = dx; ga.ldo = D; ga.copy_bf16 = dxb; ga.ldc = D;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }

    // ------------------------------------------------------------------ blocks, last to first
    int n_glob = 0;
    for (int i = 0; i < depth; ++i) n_glob += c.is_global[i] ? 1 : 0;
    int interm_idx = n_glob;
    for (int i = depth - 1; i >= 0; --i) {
        const int window = c.is_global[i] ? 0 : c.window_size;
        const int S = window > 0 ? window : g;
        const long long relb = 2LL * (2 * S - 1) * hd;
        if (c.is_global[i]) {
            --interm_idx;
            if (dinterm != nullptr && dinterm[interm_idx] != nullptr) {   // the block's output is also interm[k]
                CVLM_TRY(f.run(K_OTHER, [&] { return add_f32_launch(dx, dinterm[interm_idx], M * D, stream); }));
                CVLM_TRY(f.run(K_OTHER, [&] { return cast_bf16_launch(dx, dxb, M * D, stream); }));
            }
        }
        const void *l2wt, *l1wt, *pwt, *qwt, *relh, *relw;
        CVLM_TRY(f.getb(i, "lin2.wt", 2LL * c.mlp_dim * D, &l2wt));
        CVLM_TRY(f.getb(i, "lin1.wt", 2LL * D * c.mlp_dim, &l1wt));
        CVLM_TRY(f.getb(i, "proj.wt", 2LL * D * D, &pwt));
        CVLM_TRY(f.getb(i, "qkv.wt", 2LL * D * 3 * D, &qwt));
        CVLM_TRY(f.getb(i, "relh", relb, &relh));
        CVLM_TRY(f.getb(i, "relw", relb, &relw));
        {   // dh = (dx2 W2) * gelu'(h)
            GemmArgs ga;
            ga.a1 = dxb; ga.lda1 = D; ga.b1 = l2wt; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = c.mlp_dim; ga.out = big; ga.ldo = c.mlp_dim; ga.out_bf16 = 1;
            ga.aux = blk_ptr(i, sv.b_hpre); ga.ldaux = c.mlp_dim;
            CVLM_TRY(f.gemm(K_GEMM_LIN2, ga));
        }
        {   // dv * gamma2 = dh (W1 gamma2)
            GemmArgs ga;
            ga.a1 = big; ga.lda1 = c.mlp_dim; ga.b1 = l1wt; ga.ldb1 = c.mlp_dim; ga.k1 = c.mlp_dim;
            ga.m = static_cast<int>(M); ga.n = D; ga.out = gbuf; ga.ldo = D;
            CVLM_TRY(f.gemm(K_GEMM_LIN1, ga));
        }
        CVLM_TRY(f.run(K_LAYERNORM, [&] {
            return ln_bwd_launch(1, gbuf, nullptr, blk_ptr(i, sv.b_xh2), reinterpret_cast<const float*>(blk_ptr(i, sv.b_rstd2)),
                                 dx, dx, dxb, M, D, c.ln_eps, stream);
        }));
        {   // do = dx1 Wp
            GemmArgs ga;
            ga.a1 = dxb; ga.lda1 = D; ga.b1 = pwt; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = D; ga.out = dob; ga.ldo = D; ga.out_bf16 = 1;
            CVLM_TRY(f.gemm(K_GEMM_PROJ, ga));
        }
        f.launches += 1;     // two launches, one counted by run()
        CVLM_TRY(f.run(window > 0 ? K_ATTN_WINDOW : K_ATTN_GLOBAL, [&] {
            const bool ready = enc->lse_saved[i] != 0;
            float* lse_i = ready ? reinterpret_cast<float*>(const_cast<char*>(blk_ptr(i, sv.b_lse))) : lse;
            const float* tw_tab = nullptr;
            const float* th_tab = nullptr;
            if (ready && enc->tabs_saved[i]) {
                int gi = 0;                              // this block's slot among the global blocks
                for (int j = 0; j < i; ++j) gi += (c.is_global[j] || c.window_size <= 0) ? 1 : 0;
                tw_tab = reinterpret_cast<const float*>(svb + sv.off_gtab + sv.gtab_stride * gi);
                th_tab = reinterpret_cast<const float*>(svb + sv.off_gtab + sv.gtab_stride * gi + sv.gtab_half);
            }
            return attention_backward_launch(blk_ptr(i, sv.b_qkv), blk_ptr(i, sv.b_o), dob, big, lse_i, dsum, relh, relw,
                                             batch, g, window, c.num_heads, hd, stream, ready, tw_tab, th_tab);
        }));
        {   // du * gamma1 = dqkv (Wqkv gamma1)
            GemmArgs ga;
            ga.a1 = big; ga.lda1 = 3 * D; ga.b1 = qwt; ga.ldb1 = 3 * D; ga.k1 = 3 * D;
            ga.m = static_cast<int>(M); ga.n = D; ga.out = gbuf; ga.ldo = D;
            CVLM_TRY(f.gemm(K_GEMM_QKV, ga));
        }
        CVLM_TRY(f.run(K_LAYERNORM, [&] {
            return ln_bwd_launch(0, gbuf, nullptr, blk_ptr(i, sv.b_xb), reinterpret_cast<const float*>(blk_ptr(i, sv.b_rows1)),
                                 dx, dx, dxb, M, D, c.ln_eps, stream);
        }));
        // ---- dx is now d(loss)/d(prompt_i): the adapter's share
        CVLM_TRY(f.run(K_OTHER, [&] {
            return wgrad_tn_launch(dxb, D, D, gall + static_cast<long long>(i) * RP, ldg, RP, 1, M, p_shared,
                                   i == depth - 1 ? 0 : 1, stream);
        }));
        {   // dpre_i = (dx W_shared) * gelu'(pre_i)
            GemmArgs ga;
            ga.a1 = dxb; ga.lda1 = D; ga.b1 = sh_wt; ga.ldb1 = D; ga.k1 = D;
            ga.m = static_cast<int>(M); ga.n = RP; ga.out = dpre + static_cast<long long>(i) * RP; ga.ldo = ldg; ga.out_bf16 = 1;
            ga.aux = gpre + static_cast<long long>(i) * RP; ga.ldaux = ldg;
            CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
        }
    }

    // ------------------------------------------------------------------ adapter (image_encoder.py:278-296), backwards
    {   // dt = dpre_all W_lw(stacked)
        GemmArgs ga;
        ga.a1 = dpre; ga.lda1 = ldg; ga.b1 = lw_wt; ga.ldb1 = ldg; ga.k1 = static_cast<int>(ldg);
        ga.m = static_cast<int>(M); ga.n = RP; ga.out = dt; ga.ldo = RP; ga.out_bf16 = 1;
        CVLM_TRY(f.gemm(K_GEMM_OTHER, ga));
    }
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_tn_launch(dpre, ldg, static_cast<int>(ldg), svb + sv.off_t, RP, RP, 1, M, p_lw, 0, stream); }));
    // (ones column requested only so that every partial of this pass has the same width jp; its content is unused)
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_tn_launch(svb + sv.off_php, kp, kp, dt, RP, RP, 1, M, p_hand, 0, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_tn_launch(dt, RP, RP, dt, RP, RP, 1, M, p_dt, 0, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_reduce_launch(p_shared, ws.splits, static_cast<long long>(D) * gl.jp, grads + gl.off_shared, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_reduce_launch(p_lw, ws.splits, ldg * gl.jp, grads + gl.off_lw, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_reduce_launch(p_hand, ws.splits, static_cast<long long>(kp) * gl.jp, grads + gl.off_hand, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_reduce_launch(p_dt, ws.splits, static_cast<long long>(RP) * gl.jp, grads + gl.off_dtsum, stream); }));
    // dW_eg^T = sum_b X0_b dt_b  (X0_b: the image's patch embedding re-read as [D, T], the reference's scramble)
    CVLM_TRY(f.run(K_OTHER, [&] { return eg_wgrad_launch(svb + sv.off_x0b, dt, batch, T, D, RP, p_eg, stream); }));
    CVLM_TRY(f.run(K_OTHER, [&] { return wgrad_reduce_launch(p_eg, batch, static_cast<long long>(D) * RP, grads + gl.off_eg, stream); }));
    enc->last_launches = f.launches;
    return CVLM_OK;
}

}  // namespace

}  // namespace cvlm

using namespace cvlm;

extern "C" {

long long cvlm_encoder_saved_bytes(const cvlm_encoder* enc, int batch) {
    if (!enc || batch <= 0) return set_error(CVLM_ERR_ARG, "saved_bytes: bad argument");
    return plan_saved(enc->cfg, batch).total;
}

long long cvlm_encoder_backward_workspace_bytes(const cvlm_encoder* enc, int batch) {
    if (!enc || batch <= 0) return set_error(CVLM_ERR_ARG, "backward_workspace_bytes: bad argument");
    return plan_bwd(enc->cfg, batch).total;
}

long long cvlm_encoder_grad_floats(const cvlm_encoder* enc) {
    if (!enc) return set_error(CVLM_ERR_ARG, "grad_floats: null handle");
    return plan_grads(enc->cfg).total;
}

int cvlm_encoder_grad_layout(const cvlm_encoder* enc, long long* offsets5, int* jp) {
    if (!enc || !offsets5 || !jp) return set_error(CVLM_ERR_ARG, "grad_layout: null argument");
    const GradLayout g = plan_grads(enc->cfg);
    offsets5[0] = g.off_shared; offsets5[1] = g.off_lw; offsets5[2] = g.off_hand; offsets5[3] = g.off_dtsum; offsets5[4] = g.off_eg;
    *jp = g.jp;
    return CVLM_OK;
}

int cvlm_encoder_forward_train(cvlm_encoder* enc, const float* images, int batch, float* out, float* const* interm,
                               void* saved, long long saved_bytes, void* workspace, long long workspace_bytes,
                               cvlm_stream_t stream) {
    if (!saved) return set_error(CVLM_ERR_ARG, "forward_train: null saved-activation region");
    return forward_impl(enc, images, nullptr, batch, out, interm, -1, nullptr, workspace, workspace_bytes, saved,
                        saved_bytes, static_cast<cudaStream_t>(stream));
}

int cvlm_encoder_backward(cvlm_encoder* enc, int batch, const float* dout_nhwc, const float* const* dinterm_nhwc,
                          const void* saved, long long saved_bytes, float* grads, long long grads_floats, void* workspace,
                          long long workspace_bytes, cvlm_stream_t stream) {
    return backward_impl(enc, batch, dout_nhwc, dinterm_nhwc, saved, saved_bytes, grads, grads_floats, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream));
}

int cvlm_op_attention_backward(const void* qkv, const void* o, const void* dout, void* dqkv, float* lse, float* dsum,
                               const void* rel_pos_h, const void* rel_pos_w, int batch, int grid_side, int window,
                               int heads, int head_dim, cvlm_stream_t stream) {
    return attention_backward_launch(qkv, o, dout, dqkv, lse, dsum, rel_pos_h, rel_pos_w, batch, grid_side, window, heads,
                                     head_dim, static_cast<cudaStream_t>(stream));
}

int cvlm_op_attention_lse(const void* qkv, void* out, const void* rel_pos_h, const void* rel_pos_w, int batch,
                          int grid_side, int window, int heads, int head_dim, float* lse, float* tw_tab, float* th_tab,
                          cvlm_stream_t stream) {
    if (!lse) return set_error(CVLM_ERR_ARG, "attention_lse: null lse");
    if ((tw_tab == nullptr) != (th_tab == nullptr)) return set_error(CVLM_ERR_ARG, "attention_lse: tw_tab and th_tab go together");
    bool wrote = false, tabs = false;
    CVLM_TRY(attention_launch(qkv, out, rel_pos_h, rel_pos_w, batch, grid_side, window, heads, head_dim,
                              static_cast<cudaStream_t>(stream), lse, &wrote, tw_tab, th_tab, &tabs));
    if (!wrote) return set_error(CVLM_ERR_UNSUPPORTED, "attention_lse: this shape runs on the kernel that does not produce the log-sum-exp");
    if (tw_tab && !tabs) return set_error(CVLM_ERR_UNSUPPORTED, "attention_lse: the rel-pos tables are written for 64 x 64 global blocks only");
    return CVLM_OK;
}

int cvlm_op_attention_backward_lse(const void* qkv, const void* o, const void* dout, void* dqkv, const float* lse,
                                   float* dsum, const void* rel_pos_h, const void* rel_pos_w, int batch, int grid_side,
                                   int window, int heads, int head_dim, const float* tw_tab, const float* th_tab,
                                   cvlm_stream_t stream) {
    return attention_backward_launch(qkv, o, dout, dqkv, const_cast<float*>(lse), dsum, rel_pos_h, rel_pos_w, batch, grid_side,
                                     window, heads, head_dim, static_cast<cudaStream_t>(stream), true, tw_tab, th_tab);
}

int cvlm_op_ln_bwd(int xhat_kind, const float* gin, const float* gamma, const void* xsrc, const float* stat,
                   const float* residual, float* dx, void* dxb, long long rows, int cols, float eps, cvlm_stream_t stream) {
    return ln_bwd_launch(xhat_kind, gin, gamma, xsrc, stat, residual, dx, dxb, rows, cols, eps, static_cast<cudaStream_t>(stream));
}

int cvlm_op_wgrad_tn_jp(int j, int ones) { return wgrad_jp(j, ones); }
int cvlm_op_wgrad_tn_splits(long long m) { return wgrad_splits(m); }

int cvlm_op_wgrad_tn(const void* a, long long lda, int i_cols, const void* b, long long ldb, int j_cols, int ones,
                     long long m, float* scratch, float* out, cvlm_stream_t stream) {
    CVLM_TRY(wgrad_tn_launch(a, lda, i_cols, b, ldb, j_cols, ones, m, scratch, 0, static_cast<cudaStream_t>(stream)));
    return wgrad_reduce_launch(scratch, wgrad_splits(m), static_cast<long long>(i_cols) * wgrad_jp(j_cols, ones), out,
                               static_cast<cudaStream_t>(stream));
}

int cvlm_op_gemm_bf16_dgelu(const void* a, long long lda, const void* b, long long ldb, int k, int m, int n, int b_rows,
                            const void* aux, long long ldaux, void* out, long long ldo, cvlm_stream_t stream) {
    GemmArgs ga;
    ga.a1 = a; ga.lda1 = lda; ga.b1 = b; ga.ldb1 = ldb; ga.k1 = k; ga.m = m; ga.n = n; ga.b_rows = b_rows;
    ga.out = out; ga.ldo = ldo; ga.out_bf16 = 1; ga.aux = aux; ga.ldaux = ldaux;
    return gemm_launch(ga, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
