// Alpha-CLIP visual tower with MaPLe deep prompts: VisionTransformer_MaPLe.forward
// (alpha_clip_rw/model.py:507-563) over CustomTransformer (:478-504), CustomResidualAttentionBlock_MaPLe (:285-362),
// Attention (:176-256), LayerNorm / QuickGELU (:162-173).  Called by the cascade twice per image as
// image_encoder(images, masks, shared_ctx, deep_compound_prompts_vision) (cocotrainers/mapleAlphaCLIP.py:272, 286).
//
// Everything heavy is a kernel the encoder already has: the patch embedding of image (3 ch) and alpha map (1 ch) is ONE
// GEMM over a 4-channel im2col, every projection runs on the tcgen05 GEMM (QuickGELU in the c_fc epilogue, bias +
// residual in the out_proj / c_proj epilogues, fp32 residual stream updated in place), LayerNorm is rowops.cu's kernel,
// attention is the global-attention pipeline without relative-position terms (attention_tc2.cu, PLAIN).  The MaPLe part
// is bookkeeping on the token axis: n_ctx context tokens are appended after the patch tokens and overwritten by the next
// deep prompt before blocks 1 .. n_deep.
//
// Data layout in HBM for B images (G = resolution / patch, L = G*G + 1 + n_ctx tokens, M = B*L, W = width):
//   x    fp32 [M, W]     residual stream, token-major per image (class token, patches, context tokens)
//   a    bf16 [M, W]     LayerNorm output, then attention output
//   big  bf16 [M, 4W]    qkv (3W) / MLP hidden (4W)
//   pat  bf16 [B*G*G, 4*P*P]  im2col of (image, alpha);   pe fp32 [B*G*G, W] its embedding
#include <string.h>

#include <string>
#include <unordered_map>

#include "cvlm_internal.cuh"

struct cvlm_vit {
    cvlm_vit_config cfg;
    std::unordered_map<std::string, std::pair<const void*, long long>> tensors;
    long long last_launches = 0;
};

namespace cvlm {

namespace {

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// (image, alpha) patches -> bf16 [B*G*G, KP]; column (c*P + py)*P + px, c = 0..2 image, 3 alpha (the flattening of the
// two Conv2d weights side by side); columns >= 4*P*P are zero.  One thread per pair of columns (P is even).
__global__ void __launch_bounds__(256) vit_patch_im2col_kernel(const float* __restrict__ image, const float* __restrict__ alpha,
                                                               __nv_bfloat16* __restrict__ out, int R, int P, int G, int KP,
                                                               long long total_pairs) {
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= total_pairs) return;
    const int kp = static_cast<int>(idx % (KP / 2)) * 2;
    const long long row = idx / (KP / 2);
    const int g = static_cast<int>(row % (G * G));
    const long long b = row / (G * G);
    float2 v = make_float2(0.f, 0.f);
    if (kp < 4 * P * P) {
        const int c = kp / (P * P), py = (kp / P) % P, px = kp % P;
        const int y = (g / G) * P + py, x = (g % G) * P + px;
        const float* src = c < 3 ? image + ((b * 3 + c) * R + y) * static_cast<long long>(R) + x
                                 : alpha + (b * R + y) * static_cast<long long>(R) + x;
        v = make_float2(src[0], src[1]);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + row * KP + kp) = __floats2bfloat162_rn(v.x, v.y);
}

// x[b, 0] = class_embedding + pos[0];  x[b, 1 + p] = pe[b, p] + pos[1 + p];  x[b, 1 + G*G + i] = shared_ctx[i]
// (model.py:527-538).  One thread per float4.
__global__ void __launch_bounds__(256) vit_assemble_kernel(const float* __restrict__ pe, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, const float* __restrict__ ctx,
                                                           float* __restrict__ x, int L, int GG, int W, long long total4) {
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= total4) return;
    const int w4 = W / 4;
    const int c = static_cast<int>(idx % w4) * 4;
    const long long row = idx / w4;
    const int t = static_cast<int>(row % L);
    const long long b = row / L;
    float4 v;
    if (t > GG) {
        v = *reinterpret_cast<const float4*>(ctx + static_cast<long long>(t - GG - 1) * W + c);
    } else {
        const float4 p = *reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * W + c);
        const float4 e = t == 0 ? *reinterpret_cast<const float4*>(cls + c)
                                : *reinterpret_cast<const float4*>(pe + (b * GG + t - 1) * W + c);
        v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
    }
    *reinterpret_cast<float4*>(x + row * W + c) = v;
}

// x[b, L - n_ctx + i] = prompt[i]: the context tokens produced by the previous block are replaced (model.py:318-331)
__global__ void __launch_bounds__(256) vit_set_ctx_kernel(const float* __restrict__ prompt, float* __restrict__ x, int L,
                                                          int n_ctx, int W, long long total4) {
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= total4) return;
    const int w4 = W / 4;
    const int c = static_cast<int>(idx % w4) * 4;
    const int i = static_cast<int>((idx / w4) % n_ctx);
    const long long b = idx / (static_cast<long long>(w4) * n_ctx);
    *reinterpret_cast<float4*>(x + (b * L + L - n_ctx + i) * W + c) = *reinterpret_cast<const float4*>(prompt + static_cast<long long>(i) * W + c);
}

// cls[b] = x[b, 0]
__global__ void __launch_bounds__(256) vit_gather_cls_kernel(const float* __restrict__ x, float* __restrict__ cls, int L, int W,
                                                             long long total4) {
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= total4) return;
    const int w4 = W / 4;
    const int c = static_cast<int>(idx % w4) * 4;
    const long long b = idx / w4;
    *reinterpret_cast<float4*>(cls + b * W + c) = *reinterpret_cast<const float4*>(x + b * L * W + c);
}

struct VitWs {
    long long x, a, big, pat, pe, cls, clsb, total;
};

int vit_kp(const cvlm_vit_config& c) { return (4 * c.patch_size * c.patch_size + 7) / 8 * 8; }

VitWs plan(const cvlm_vit_config& c, int batch) {
    const long long G = c.input_resolution / c.patch_size, GG = G * G, L = GG + 1 + c.n_ctx, M = L * batch, W = c.width;
    VitWs w;
    long long off = 0;
    auto take = [&](long long bytes) {
        const long long o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    w.x = take(M * W * 4);
    w.a = take(M * W * 2);
    w.big = take(M * 4 * W * 2);
    w.pat = take(GG * batch * vit_kp(c) * 2);
    w.pe = take(GG * batch * W * 4);
    w.cls = take(static_cast<long long>(batch) * W * 4);
    w.clsb = take(static_cast<long long>(batch) * W * 2);
    w.total = off;
    return w;
}

int check_config(const cvlm_vit_config& c) {
    if (c.patch_size < 2 || c.patch_size % 2 || c.input_resolution < c.patch_size || c.input_resolution % c.patch_size)
        return set_error(CVLM_ERR_UNSUPPORTED, "vit: resolution %d / patch %d (patch must be even and divide the resolution)",
                         c.input_resolution, c.patch_size);
    if (c.width < 128 || c.width % 128 || c.width > 2048)
        return set_error(CVLM_ERR_UNSUPPORTED, "vit: width %d must be a multiple of 128 up to 2048", c.width);
    if (c.heads < 1 || c.width % c.heads || (c.width / c.heads != 64 && c.width / c.heads != 80))
        return set_error(CVLM_ERR_UNSUPPORTED, "vit: width %d / heads %d: head_dim must be 64 or 80", c.width, c.heads);
    if (c.layers < 1 || c.layers > 64) return set_error(CVLM_ERR_UNSUPPORTED, "vit: %d layers outside [1, 64]", c.layers);
    if (c.output_dim < 8 || c.output_dim % 8) return set_error(CVLM_ERR_UNSUPPORTED, "vit: output_dim %d must be a multiple of 8", c.output_dim);
    if (c.n_ctx < 1 || c.n_ctx > 64) return set_error(CVLM_ERR_UNSUPPORTED, "vit: %d context tokens outside [1, 64]", c.n_ctx);
    return CVLM_OK;
}

struct VitFwd {
    cvlm_vit* vit;
    cudaStream_t stream;
    long long launches = 0;
    char name[160];

    int get(const char* prefix, const char* suffix, long long bytes, const void** out) {
        snprintf(name, sizeof(name), "%s%s", prefix, suffix);
        auto it = vit->tensors.find(name);
        if (it == vit->tensors.end()) return set_error(CVLM_ERR_MISSING, "vit tensor '%s' was not registered", name);
        if (bytes > 0 && it->second.second != bytes)
            return set_error(CVLM_ERR_MISSING, "vit tensor '%s' has %lld bytes, expected %lld", name, it->second.second, bytes);
        *out = it->second.first;
        return CVLM_OK;
    }
    int getf(const char* prefix, const char* suffix, long long elems, const float** out) {
        const void* p = nullptr;
        CVLM_TRY(get(prefix, suffix, elems * 4, &p));
        *out = static_cast<const float*>(p);
        return CVLM_OK;
    }
    int gemm(const void* a, long long m, int k, long long lda, const void* w, int n, const float* bias, const float* residual,
             void* out, int out_bf16, int act) {
        GemmArgs g;
        g.a1 = a; g.lda1 = lda; g.b1 = w; g.ldb1 = lda; g.k1 = k;
        g.m = static_cast<int>(m); g.n = n;
        g.bias = bias; g.residual = residual; g.ldr = n;
        g.out = out; g.ldo = n; g.out_bf16 = out_bf16; g.act_gelu = act;
        ++launches;
        return gemm_launch(g, stream);
    }
    int ln(const float* x, const float* g, const float* b, void* out, int out_bf16, long long rows, int cols, float eps) {
        ++launches;
        return layernorm_launch(x, g, b, out, out_bf16, rows, cols, eps, stream);
    }
};

int vit_forward(cvlm_vit* vit, const float* images, const float* alpha, const float* shared_ctx, const float* deep, int n_deep,
                int batch, float* out, void* workspace, long long workspace_bytes, cudaStream_t stream) {
    const cvlm_vit_config& c = vit->cfg;
    if (!images || !alpha || !shared_ctx || !out || !workspace || (n_deep > 0 && !deep))
        return set_error(CVLM_ERR_ARG, "vit_forward: null pointer");
    if (batch < 1 || n_deep < 0) return set_error(CVLM_ERR_ARG, "vit_forward: batch %d, %d deep prompts", batch, n_deep);
    const VitWs ws = plan(c, batch);
    if (workspace_bytes < ws.total)
        return set_error(CVLM_ERR_WORKSPACE, "vit_forward: workspace %lld < %lld bytes", workspace_bytes, ws.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error(CVLM_ERR_ARG, "vit_forward: workspace must be 256-byte aligned");
    const int R = c.input_resolution, P = c.patch_size, G = R / P, GG = G * G, L = GG + 1 + c.n_ctx, W = c.width, KP = vit_kp(c);
    const long long M = static_cast<long long>(L) * batch, MP = static_cast<long long>(GG) * batch;
    if (M >= (1ll << 31) / 4) return set_error(CVLM_ERR_UNSUPPORTED, "vit_forward: batch %d too large for one call (chunk it)", batch);
    char* base = static_cast<char*>(workspace);
    float* x = reinterpret_cast<float*>(base + ws.x);
    void* a = base + ws.a;
    void* big = base + ws.big;
    VitFwd f{vit, stream};

    // ---- patch embedding of (image, alpha): conv1(x) + conv1_alpha(alpha) (model.py:522-524) as one GEMM
    const void* wpe;
    CVLM_TRY(f.get("patch.w", "", static_cast<long long>(W) * KP * 2, &wpe));
    {
        const long long pairs = MP * (KP / 2);
        ++f.launches;
        vit_patch_im2col_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, stream>>>(
            images, alpha, reinterpret_cast<__nv_bfloat16*>(base + ws.pat), R, P, G, KP, pairs);
        CVLM_CUDA(cudaGetLastError());
    }
    CVLM_TRY(f.gemm(base + ws.pat, MP, KP, KP, wpe, W, nullptr, nullptr, base + ws.pe, 0, 0));
    // ---- class token, positional embedding, shallow context tokens, ln_pre (:527-540)
    const float *cls, *pos, *g0, *b0;
    CVLM_TRY(f.getf("class_embedding", "", W, &cls));
    CVLM_TRY(f.getf("positional_embedding", "", static_cast<long long>(GG + 1) * W, &pos));
    CVLM_TRY(f.getf("ln_pre", ".weight", W, &g0));
    CVLM_TRY(f.getf("ln_pre", ".bias", W, &b0));
    {
        const long long total4 = M * (W / 4);
        ++f.launches;
        vit_assemble_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const float*>(base + ws.pe), cls, pos, shared_ctx, reinterpret_cast<float*>(big), L, GG, W, total4);
        CVLM_CUDA(cudaGetLastError());
    }
    CVLM_TRY(f.ln(reinterpret_cast<const float*>(big), g0, b0, x, 0, M, W, c.ln_eps));      // `big` (8W bytes per token) holds the fp32 rows
    // ---- transformer
    char pre[96];
    int counter = 0;
    for (int i = 0; i < c.layers; ++i) {
        snprintf(pre, sizeof(pre), "transformer.resblocks.%d", i);
        if (i > 0 && counter < n_deep) {           // CustomResidualAttentionBlock_MaPLe.forward :310-331
            const long long total4 = static_cast<long long>(batch) * c.n_ctx * (W / 4);
            ++f.launches;
            vit_set_ctx_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(
                deep + static_cast<long long>(counter) * c.n_ctx * W, x, L, c.n_ctx, W, total4);
            CVLM_CUDA(cudaGetLastError());
            ++counter;
        }
        const float *l1w, *l1b, *l2w, *l2b, *bqkv, *bo, *bfc, *bpr;
        const void *wqkv, *wo, *wfc, *wpr;
        CVLM_TRY(f.getf(pre, ".ln_1.weight", W, &l1w));
        CVLM_TRY(f.getf(pre, ".ln_1.bias", W, &l1b));
        CVLM_TRY(f.getf(pre, ".ln_2.weight", W, &l2w));
        CVLM_TRY(f.getf(pre, ".ln_2.bias", W, &l2b));
        CVLM_TRY(f.getf(pre, ".attn.in_proj.bias", 3ll * W, &bqkv));
        CVLM_TRY(f.getf(pre, ".attn.out_proj.bias", W, &bo));
        CVLM_TRY(f.getf(pre, ".mlp.c_fc.bias", 4ll * W, &bfc));
        CVLM_TRY(f.getf(pre, ".mlp.c_proj.bias", W, &bpr));
        CVLM_TRY(f.get(pre, ".attn.in_proj.weight.bf16", 3ll * W * W * 2, &wqkv));
        CVLM_TRY(f.get(pre, ".attn.out_proj.weight.bf16", static_cast<long long>(W) * W * 2, &wo));
        CVLM_TRY(f.get(pre, ".mlp.c_fc.weight.bf16", 4ll * W * W * 2, &wfc));
        CVLM_TRY(f.get(pre, ".mlp.c_proj.weight.bf16", 4ll * W * W * 2, &wpr));
        CVLM_TRY(f.ln(x, l1w, l1b, a, 1, M, W, c.ln_eps));
        CVLM_TRY(f.gemm(a, M, W, W, wqkv, 3 * W, bqkv, nullptr, big, 1, 0));
        ++f.launches;
        CVLM_TRY(attention_plain_tc2_launch(big, a, batch, L, c.heads, W / c.heads, stream));
        CVLM_TRY(f.gemm(a, M, W, W, wo, W, bo, x, x, 0, 0));
        CVLM_TRY(f.ln(x, l2w, l2b, a, 1, M, W, c.ln_eps));
        CVLM_TRY(f.gemm(a, M, W, W, wfc, 4 * W, bfc, nullptr, big, 1, 2));          // QuickGELU
        CVLM_TRY(f.gemm(big, M, 4 * W, 4ll * W, wpr, W, bpr, x, x, 0, 0));
    }
    // ---- ln_post on the class token, projection (:550-553)
    const float *gp, *bp;
    const void* wproj;
    CVLM_TRY(f.getf("ln_post", ".weight", W, &gp));
    CVLM_TRY(f.getf("ln_post", ".bias", W, &bp));
    CVLM_TRY(f.get("proj.t", "", static_cast<long long>(c.output_dim) * W * 2, &wproj));
    {
        const long long total4 = static_cast<long long>(batch) * (W / 4);
        ++f.launches;
        vit_gather_cls_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(
            x, reinterpret_cast<float*>(base + ws.cls), L, W, total4);
        CVLM_CUDA(cudaGetLastError());
    }
    CVLM_TRY(f.ln(reinterpret_cast<const float*>(base + ws.cls), gp, bp, base + ws.clsb, 1, batch, W, c.ln_eps));
    CVLM_TRY(f.gemm(base + ws.clsb, batch, W, W, wproj, c.output_dim, nullptr, nullptr, out, 0, 0));
    vit->last_launches = f.launches;
    return CVLM_OK;
}

}  // namespace

}  // namespace cvlm

extern "C" {

int cvlm_vit_create(const cvlm_vit_config* cfg, cvlm_vit** out) {
    if (!cfg || !out) return cvlm::set_error(CVLM_ERR_ARG, "vit_create: null argument");
    CVLM_TRY(cvlm::check_config(*cfg));
    cvlm_vit* v = new cvlm_vit();
    v->cfg = *cfg;
    *out = v;
    return CVLM_OK;
}

int cvlm_vit_destroy(cvlm_vit* vit) {
    delete vit;
    return CVLM_OK;
}

int cvlm_vit_set_tensor(cvlm_vit* vit, const char* name, const void* dev_ptr, long long nbytes) {
    if (!vit || !name || !dev_ptr || nbytes <= 0) return cvlm::set_error(CVLM_ERR_ARG, "vit_set_tensor: bad argument");
    if (reinterpret_cast<uintptr_t>(dev_ptr) & 15) return cvlm::set_error(CVLM_ERR_ARG, "vit_set_tensor: '%s' is not 16-byte aligned", name);
    vit->tensors[name] = {dev_ptr, nbytes};
    return CVLM_OK;
}

long long cvlm_vit_workspace_bytes(const cvlm_vit* vit, int batch) {
    if (!vit || batch < 1) return cvlm::set_error(CVLM_ERR_ARG, "vit_workspace_bytes: bad argument");
    return cvlm::plan(vit->cfg, batch).total;
}

long long cvlm_vit_last_launches(const cvlm_vit* vit) { return vit ? vit->last_launches : 0; }

int cvlm_vit_forward(cvlm_vit* vit, const float* images, const float* alpha, const float* shared_ctx, const float* deep_prompts,
                     int n_deep, int batch, float* out, void* workspace, long long workspace_bytes, cvlm_stream_t stream) {
    if (!vit) return cvlm::set_error(CVLM_ERR_ARG, "vit_forward: null handle");
    return cvlm::vit_forward(vit, images, alpha, shared_ctx, deep_prompts, n_deep, batch, out, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
