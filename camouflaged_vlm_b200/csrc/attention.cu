// Fused window / global attention with decomposed relative position bias.
//
// Replaces, for one block of the encoder, the reference chain
//   window_partition (image_encoder.py:507-528) -> Attention.forward core (:493-501) with
//   add_decomposed_rel_pos (:589-625) / get_rel_pos (:556-586) -> window_unpartition (:531-553)
// without materialising the padded windows, the [B*heads, N, N] score matrix or the rel-pos
// broadcast.  One kernel serves both block kinds: a "window" is an S x S patch of the G x G token
// grid (S = 14 windowed, S = G global).  Tokens of a window that fall outside the grid are the
// reference's zero padding: they are real softmax columns whose key and value are exactly zero
// (K/V are stored WITHOUT bias -- b_k shifts every logit of a row by the same q.b_k and cancels,
// b_v factors out of the convex combination and is folded into the proj bias on the host), so
// their logit is the rel-pos term alone.  Padded QUERIES are never stored (the reference crops them).
//
// Layout: qkv is the un-partitioned token-major buffer [B*G*G, 3*D] bf16 written by the QKV GEMM,
// column = which*D + head*HD + d  (the reference's reshape(B, HW, 3, heads, hd), :491).
// Output o is [B*G*G, D] bf16, column = head*HD + d  (:501).
//
// v1 math path: mma.sync m16n8k16 bf16 with fp32 accumulation, flash-style online softmax in the
// exp2 domain, K/V tiles double-buffered with cp.async.  (tcgen05 port: see DESIGN.md roadmap.)
#include <stdlib.h>

#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

constexpr int ATT_BQ = 64;       // queries per CTA (4 warps x 16 rows)
constexpr int ATT_BK = 64;       // keys per tile
constexpr int ATT_THREADS = 128;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnParams {
    const __nv_bfloat16* qkv;   // [B*G*G, 3*D]
    __nv_bfloat16* out;         // [B*G*G, D]
    const __nv_bfloat16* relh;  // [2S-1, HD]
    const __nv_bfloat16* relw;  // [2S-1, HD]
    int G;        // token grid side
    int S;        // window side (== G for a global block)
    int nW;       // windows per side = ceil(G / S)
    int heads;
    int D;        // heads * HD
    int TWP;      // padded rel-pos table width (multiple of 8, >= 2S-1)
    float scale_log2;  // HD^-0.5 * log2(e)
};

// Shared-memory plan.  Two modes:
//  generic     : Th, Tw = [64][TWP+1] fp32 live for the whole kernel; region U first stages both
//                rel-pos tables (bf16), then becomes the K/V double buffer.
//  row-aligned : S == 64, so a query tile is one grid row and a key tile is one grid row.  Th is
//                only [64][65] (rel_pos_h rows qt..qt+63), Tw is needed once (gathered into
//                registers) and shares region U with the staged tables; the K/V double buffer
//                then overlays all of U.  ~93 KB -> 2 CTAs per SM.
template <int HD>
struct AttnPlan {
    static constexpr int HDP = HD + 8;   // padded bf16 row: 176 B (HD 80) / 144 B (HD 64): conflict-free ldmatrix
    static constexpr int Q_BYTES = ATT_BQ * HDP * 2;
    static constexpr int KV_BYTES = 2 * 2 * ATT_BK * HDP * 2;
    int th_stride, tw_stride, th_bytes, tw_bytes, stage_bytes, u_bytes, total;
    int off_th, off_tw, off_stage, off_kv, off_kyx;
    __host__ __device__ AttnPlan(int TWP, bool aligned) {
        th_stride = aligned ? 65 : TWP + 1;
        tw_stride = TWP + 1;
        th_bytes = ATT_BQ * th_stride * 4;
        tw_bytes = ATT_BQ * tw_stride * 4;
        stage_bytes = (aligned ? (64 + TWP) : 2 * TWP) * HDP * 2;
        off_th = Q_BYTES;
        if (aligned) {
            off_stage = off_th + th_bytes;
            off_tw = off_stage + stage_bytes;
            off_kv = off_stage;
            u_bytes = stage_bytes + tw_bytes > KV_BYTES ? stage_bytes + tw_bytes : KV_BYTES;
            off_kyx = off_stage + u_bytes;
        } else {
            off_tw = off_th + th_bytes;
            off_stage = off_tw + tw_bytes;
            off_kv = off_stage;
            u_bytes = stage_bytes > KV_BYTES ? stage_bytes : KV_BYTES;
            off_kyx = off_stage + u_bytes;
        }
        total = off_kyx + 2 * ATT_BK * 4;
    }
};

template <int HD, bool kRowAligned>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
    using Plan = AttnPlan<HD>;
    constexpr int HDP = Plan::HDP;
    constexpr int KSTEPS = HD / 16;     // k-steps of QK^T
    constexpr int DT = HD / 8;          // n-tiles of the output (head dim)
    constexpr int CH = HD / 8;          // 16-byte chunks per row
    extern __shared__ __align__(16) unsigned char smem[];

    const int S = p.S, G = p.G, L = S * S, TW = 2 * S - 1, TWP = p.TWP;
    const Plan plan(TWP, kRowAligned);
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
    float* sTh = reinterpret_cast<float*>(smem + plan.off_th);
    float* sTw = reinterpret_cast<float*>(smem + plan.off_tw);
    __nv_bfloat16* sStage = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_stage);
    __nv_bfloat16* sKV = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_kv);   // [stage][K|V][64][HDP]
    short2* sKyx = reinterpret_cast<short2*>(smem + plan.off_kyx);               // [stage][64]
    const int THS = plan.th_stride, TWS = plan.tw_stride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, head = blockIdx.y;
    const int bw = blockIdx.z;
    const int b = bw / (p.nW * p.nW), win = bw % (p.nW * p.nW);
    const int wy = win / p.nW, wx = win % p.nW;
    const long long tok_base = static_cast<long long>(b) * G * G;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;

    // window-local token l -> (ly, lx) and global token (or -1 for zero padding / rows beyond L)
    auto global_token = [&](int l, int& ly, int& lx) -> int {
        if (l >= L) {
            ly = 0;
            lx = 0;
            return -1;
        }
        ly = l / S;
        lx = l - ly * S;
        const int gy = wy * S + ly, gx = wx * S + lx;
        return (gy < G && gx < G) ? gy * G + gx : -1;
    };

    // a query tile whose first row is already below the grid holds only padding (bottom-edge windows)
    if (wy * S + (qt * ATT_BQ) / S >= G) return;

    // ---- prologue loads: Q tile + rel-pos tables (staged in region U)
    for (int i = tid; i < ATT_BQ * CH; i += ATT_THREADS) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = global_token(qt * ATT_BQ + r, ly, lx);
        const __nv_bfloat16* src = qbase + (tok_base + (gt < 0 ? 0 : gt)) * ld + c * 8;
        cp_async_16(smem_u32(sQ + r * HDP + c * 8), src, gt >= 0);
    }
    // staged rows: generic = [relh 0..TWP) | relw 0..TWP) ; aligned = [relh qt..qt+64) | relw 0..TWP)
    const int hrows = kRowAligned ? 64 : TWP;
    const int hfirst = kRowAligned ? qt : 0;
    for (int i = tid; i < (hrows + TWP) * CH; i += ATT_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const bool is_w = r >= hrows;
        const int src_row = is_w ? r - hrows : hfirst + r;
        const __nv_bfloat16* tab = is_w ? p.relw : p.relh;
        cp_async_16(smem_u32(sStage + r * HDP + c * 8), tab + (src_row < TW ? src_row : 0) * HD + c * 8, src_row < TW);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- Q fragments (A operand), 16 rows per warp
    uint32_t qf[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
            ldmatrix_x4(smem_u32(sQ + row * HDP + ks * 16 + koff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    // ---- rel-pos projections of the UNSCALED q (image_encoder.py:618-619): T[q][d] = q . rel_pos[d],
    //      pre-multiplied by log2(e) because the softmax runs in the exp2 domain.
    {
        const int krow = (lane & 7) + ((lane >> 4) << 3);   // key row within a pair of 8-wide n-tiles
        const int kcol = ((lane >> 3) & 1) * 8;
        for (int which = 0; which < 2; ++which) {
            const __nv_bfloat16* sR = which ? sStage + hrows * HDP : sStage;
            float* sT = which ? sTw : sTh;
            const int ncols = which ? TWP : hrows;
            const int ts = which ? TWS : THS;
            for (int n0 = 0; n0 < ncols; n0 += 16) {
                float acc[2][4] = {};
                int r = n0 + krow;
                r = r < ncols ? r : ncols - 1;   // second n-tile of the last pair may lie beyond ncols
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(smem_u32(sR + r * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                    mma_bf16_16816(acc[0], qf[ks], b0, b1);
                    mma_bf16_16816(acc[1], qf[ks], b2, b3);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = n0 + h * 8 + 2 * t;
                    if (col < ncols) {
                        const int r0 = warp * 16 + g;
                        sT[r0 * ts + col] = acc[h][0] * kLog2e;
                        sT[r0 * ts + col + 1] = acc[h][1] * kLog2e;
                        sT[(r0 + 8) * ts + col] = acc[h][2] * kLog2e;
                        sT[(r0 + 8) * ts + col + 1] = acc[h][3] * kLog2e;
                    }
                }
            }
        }
    }
    __syncwarp();   // each warp only reads back its own 16 rows of Th / Tw

    // per-thread row bookkeeping: rows r0 = warp*16+g and r0+8
    int qyo[2], qxo[2];   // qy + S - 1, qx + S - 1
    const float* thr[2];
    const float* twr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = warp * 16 + g + h * 8;
        int ly, lx;
        global_token(qt * ATT_BQ + r, ly, lx);
        qyo[h] = ly + S - 1;
        qxo[h] = lx + S - 1;
        thr[h] = sTh + r * THS;
        twr[h] = sTw + r * TWS;
    }
    // Row-aligned fast path (S == 64 == ATT_BK): key tile kt is grid row kt, so the w-term is the same
    // for every key tile (kept in registers) and the h-term is one scalar per row and tile:
    // Th was projected on rel_pos_h rows qt..qt+63, and qy - ky + 63 = qt + (63 - kt).
    float twreg[kRowAligned ? 8 : 1][4];
    if constexpr (kRowAligned) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            twreg[nt][0] = twr[0][qxo[0] - col];
            twreg[nt][1] = twr[0][qxo[0] - col - 1];
            twreg[nt][2] = twr[1][qxo[1] - col];
            twreg[nt][3] = twr[1][qxo[1] - col - 1];
        }
    }
    __syncthreads();   // tables (and, when aligned, Tw) are dead: region U becomes the K/V double buffer

    auto load_kv = [&](int kt, int stage) {
        __nv_bfloat16* sK = sKV + stage * 2 * ATT_BK * HDP;
        __nv_bfloat16* sV = sK + ATT_BK * HDP;
        for (int i = tid; i < ATT_BK * CH; i += ATT_THREADS) {
            const int r = i / CH, c = i - r * CH;
            int ly, lx;
            const int gt = global_token(kt * ATT_BK + r, ly, lx);
            const long long row = (tok_base + (gt < 0 ? 0 : gt)) * ld + c * 8;
            cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, gt >= 0);
            cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, gt >= 0);
            if (c == 0) sKyx[stage * ATT_BK + r] = make_short2(static_cast<short>(ly), static_cast<short>(lx));
        }
    };
    const int nkt = (L + ATT_BK - 1) / ATT_BK;
    load_kv(0, 0);
    cp_async_commit();

    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int kt = 0; kt < nkt; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < nkt) load_kv(kt + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();      // tile kt landed (the prefetch of kt+1 may still be in flight)
        __syncthreads();
        const __nv_bfloat16* sK = sKV + stage * 2 * ATT_BK * HDP;
        const __nv_bfloat16* sV = sK + ATT_BK * HDP;
        const short2* kyx = sKyx + stage * ATT_BK;

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        {
            const int krow = (lane & 7) + ((lane >> 4) << 3);
            const int kcol = ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(smem_u32(sK + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
                    mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
                }
            }
        }
        // ---- logits (log2 domain) = scale*S + rel_h + rel_w ; keys beyond L masked out
        if constexpr (kRowAligned) {
            const float th0 = thr[0][63 - kt], th1 = thr[1][63 - kt];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = fmaf(s[nt][0], p.scale_log2, th0 + twreg[nt][0]);
                s[nt][1] = fmaf(s[nt][1], p.scale_log2, th0 + twreg[nt][1]);
                s[nt][2] = fmaf(s[nt][2], p.scale_log2, th1 + twreg[nt][2]);
                s[nt][3] = fmaf(s[nt][3], p.scale_log2, th1 + twreg[nt][3]);
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = nt * 8 + 2 * t + e;
                    const short2 yx = kyx[col];
                    const bool live = kt * ATT_BK + col < L;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float bias = thr[h][qyo[h] - yx.x] + twr[h][qxo[h] - yx.y];
                        const float v = fmaf(s[nt][2 * h + e], p.scale_log2, bias);
                        s[nt][2 * h + e] = live ? v : -INFINITY;
                    }
                }
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
        float alpha[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);   // finite: every tile holds >= 1 live key
            alpha[h] = fast_exp2(m_run[h] - m_new);
            m_run[h] = m_new;
            l_run[h] *= alpha[h];
        }
        // ---- P = exp2(logit - m), packed straight into A fragments for P V
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = fast_exp2(s[nt][0] - m_run[0]), p1 = fast_exp2(s[nt][1] - m_run[0]);
            const float p2 = fast_exp2(s[nt][2] - m_run[1]), p3 = fast_exp2(s[nt][3] - m_run[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            o[i][0] *= alpha[0];
            o[i][1] *= alpha[0];
            o[i][2] *= alpha[1];
            o[i][3] *= alpha[1];
        }
        // ---- O += P V   (V tile is [key][d]; ldmatrix.trans yields the col-major B fragments)
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {          // 16 keys per k-step
#pragma unroll
                for (int dp = 0; dp < DT / 2; ++dp) { // pairs of 8-wide d tiles
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sV + (kk * 16 + vrow) * HDP + dp * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * dp], pf[kk], b0, b1);
                    mma_bf16_16816(o[2 * dp + 1], pf[kk], b2, b3);
                }
            }
        }
        __syncthreads();   // everyone is done with this stage before it is refilled
    }

    // ---- epilogue: normalise, stage the warp's 16 rows in its (now dead) Q rows, 16-byte stores
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        l_run[h] = 1.f / l_run[h];
    }
    __nv_bfloat16* sO = sQ + warp * 16 * HDP;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(sO + g * HDP + i * 8 + 2 * t) = pack_bf16(o[i][0] * l_run[0], o[i][1] * l_run[0]);
        *reinterpret_cast<uint32_t*>(sO + (g + 8) * HDP + i * 8 + 2 * t) =
            pack_bf16(o[i][2] * l_run[1], o[i][3] * l_run[1]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = global_token(qt * ATT_BQ + warp * 16 + r, ly, lx);
        if (gt >= 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(sO + r * HDP + c * 8);
            *reinterpret_cast<uint4*>(p.out + (tok_base + gt) * p.D + head * HD + c * 8) = v;
        }
    }
}

template <int HD, bool kRowAligned>
static int launch_attention(const AttnParams& p, int batch, cudaStream_t stream) {
    const int L = p.S * p.S;
    const AttnPlan<HD> plan(p.TWP, kRowAligned);
    auto kern = attention_kernel<HD, kRowAligned>;
    CVLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total));
    dim3 grid((L + ATT_BQ - 1) / ATT_BQ, p.heads, batch * p.nW * p.nW);
    kern<<<grid, ATT_THREADS, plan.total, stream>>>(p);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

int attention_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int grid_side,
                     int window, int heads, int head_dim, cudaStream_t stream, float* lse_out, bool* lse_written,
                     float* tw_out, float* th_out, bool* tabs_written) {
    if (lse_written) *lse_written = false;
    if (tabs_written) *tabs_written = false;
    if (!qkv || !out || !relh || !relw) return set_error(CVLM_ERR_ARG, "attention: null pointer");
    if (batch <= 0 || grid_side <= 0 || heads <= 0) return set_error(CVLM_ERR_ARG, "attention: bad sizes");
    if (head_dim != 32 && head_dim != 64 && head_dim != 80)
        return set_error(CVLM_ERR_UNSUPPORTED, "attention: head_dim must be 32, 64 or 80");
    AttnParams p;
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.out = static_cast<__nv_bfloat16*>(out);
    p.relh = static_cast<const __nv_bfloat16*>(relh);
    p.relw = static_cast<const __nv_bfloat16*>(relw);
    p.G = grid_side;
    p.S = window > 0 ? window : grid_side;   // a window larger than the grid is one padded window
    p.nW = (p.G + p.S - 1) / p.S;
    p.heads = heads;
    p.D = heads * head_dim;
    p.TWP = ((2 * p.S - 1) + 7) / 8 * 8;
    p.scale_log2 = (1.0f / sqrtf(static_cast<float>(head_dim))) * kLog2e;
    if (p.S > 64) return set_error(CVLM_ERR_UNSUPPORTED, "attention: window / grid side > 64 tokens");
    if (static_cast<long long>(batch) * p.nW * p.nW > 65535)
        return set_error(CVLM_ERR_UNSUPPORTED, "attention: batch * windows exceeds grid.z; split the batch");
    // 64x64 global blocks and 14x14 windows run on the tcgen05 kernels (attention_tc2.cu, attention_tc_win2.cu);
    // cvlm_set_option("attn_legacy", 1) (or CVLM_ATTN_LEGACY=1 at start-up) forces the mma.sync kernel below (A/B comparisons, bisecting).
    const bool legacy = option(OPT_ATTN_LEGACY) != 0;
    if (!legacy && p.S == 64 && p.G == 64 && (head_dim == 64 || head_dim == 80)) {
        if (lse_written) *lse_written = lse_out != nullptr;
        if (tabs_written) *tabs_written = tw_out != nullptr && th_out != nullptr;
        return attention_global_tc2_launch(qkv, out, relh, relw, batch, heads, head_dim, stream, lse_out, tw_out, th_out);
    }
    if (!legacy && window == 14 && (head_dim == 64 || head_dim == 80)) {
        if (lse_written) *lse_written = lse_out != nullptr;
        return attention_window_tc2_launch(qkv, out, relh, relw, batch, grid_side, heads, head_dim, stream, lse_out);
    }
    const bool aligned = (p.S == ATT_BK);
    if (head_dim == 80)
        return aligned ? launch_attention<80, true>(p, batch, stream) : launch_attention<80, false>(p, batch, stream);
    if (head_dim == 64)
        return aligned ? launch_attention<64, true>(p, batch, stream) : launch_attention<64, false>(p, batch, stream);
    return aligned ? launch_attention<32, true>(p, batch, stream) : launch_attention<32, false>(p, batch, stream);
}

}  // namespace cvlm
