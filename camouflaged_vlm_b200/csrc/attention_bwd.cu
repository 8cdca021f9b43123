// Backward of the fused window / global attention with decomposed relative position bias (adapter-only training,
// SURVEY.md 8 f4): the gradient must cross every frozen block to reach the adapter's prompts
// (train_ovcos_maskdecoder_edge.py:301-303, models/sam_maskdecoder_edge.py:429-460).
//
// Forward (attention.cu; reference image_encoder.py:488-504, 589-625), per window / head, logits in the log2 domain:
//     l[q,k] = scale * q.k + q.Rh[qy - ky + S-1] + q.Rw[qx - kx + S-1],   P = softmax_k(l),   o = P v
// with bias-free K / V, zero K / V for the reference's padded tokens (they still count in the softmax) and padded
// queries never stored.  Given dO:
//     D[q]  = sum_d dO[q,d] o[q,d]            dP = dO V^T            dS = P * (dP - D)
//     dV    = P^T dO                          dK = scale * dS^T Q
//     dQ    = scale * dS K + dTh Rh + dTw Rw,   dTh[q,j] = sum_{k: qy-ky+S-1 == j} dS[q,k]  (dTw likewise)
// rel_pos_h / rel_pos_w belong to the frozen trunk: no gradient for them.  Padded keys produce no dK / dV (constants in
// the reference: b_k, b_v), padded queries have dO = 0.
//
// Two deterministic kernels (no atomics), mma.sync m16n8k16 bf16 with fp32 accumulation, same tiling as attention.cu:
//   attn_bwd_dq_kernel   one CTA per 64-query tile: pass 1 recomputes the row log-sum-exp (and stores it with D for the
//                        second kernel), pass 2 walks the key tiles again and accumulates dQ.  The per-row sums of dS
//                        over kx / ky that the rel-pos terms need are products with 0/1 indicator matrices built in
//                        registers (dSh = dS Eh, dSw = dS Ew), so they ride the tensor cores as well.
//   attn_bwd_dkv_kernel  one CTA per 64-key tile, loops over the query tiles: S^T = K Q^T, dP^T = V dO^T, dV += P^T dO,
//                        dK += dS^T Q.  The rel-pos projections T = q . rel_pos of each query tile are recomputed on the
//                        fly (cheaper than spilling [tokens, 2(2S-1)] floats per head).
// Both are generic in the window side S <= 64 (14x14 windows and the 64x64 global grid take the same path).
#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

namespace {

constexpr int BW_T = 64;            // queries / keys per tile
constexpr int BW_THREADS = 128;
constexpr float kLog2eB = 1.4426950408889634f;

struct BwdParams {
    const __nv_bfloat16* qkv;   // [B*G*G, 3*D] forward input (q with bias, bias-free k / v)
    const __nv_bfloat16* o;     // [B*G*G, D] forward output
    const __nv_bfloat16* dout;  // [B*G*G, D] gradient of the forward output
    __nv_bfloat16* dqkv;        // [B*G*G, 3*D] out: gradient of qkv
    float* lse;                 // [B*G*G, heads] row log-sum-exp in the log2 domain (written by dq, read by dkv)
    float* dsum;                // [B*G*G, heads] D = rowsum(dO * o)
    const __nv_bfloat16* relh;  // [2S-1, HD]
    const __nv_bfloat16* relw;
    int G, S, nW, heads, D, TWP, TWK;
    float scale, scale_log2;
    int lse_ready;              // lse was written by the forward kernel: the specialised kernels skip their own S sweep for it
    // 64 x 64 global blocks, written by the training forward (attention_tc2.cu), log2 domain, fp32 [B*G*G, heads, 64] each:
    // tw_tab[q][kx] = T_w[q][qx - kx + 63] - lse[q] and th_tab[q][ky] = T_h[q][qy - ky + 63]; NULL: the kernels project q
    // on the rel-pos tables themselves
    const float* tw_tab;
    const float* th_tab;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
struct BwdSmem {
    static constexpr int HDP = HD + 8;
    int ts;                       // fp32 row stride of the T tables
    int off_q, off_do, off_th, off_tw, off_rel, off_a, off_b, off_kyx, off_row, total;
    __host__ __device__ BwdSmem(int TWP, int TWK) {
        ts = TWP + 1;
        int off = 0;
        off_q = off;    off += BW_T * HDP * 2;
        off_do = off;   off += BW_T * HDP * 2;
        off_a = off;    off += 2 * BW_T * HDP * 2;          // dq: K/V stage 0 ; dkv: this CTA's K and V
        off_b = off;    off += 2 * BW_T * HDP * 2;          // dq: K/V stage 1
        off_rel = off;  off += 2 * TWK * HDP * 2;           // rel_pos_h | rel_pos_w, rows >= 2S-1 zero
        off = (off + 15) & ~15;
        off_th = off;   off += BW_T * ts * 4;
        off_tw = off;   off += BW_T * ts * 4;
        off_kyx = off;  off += 2 * BW_T * 4;                // (y, x) of the keys / queries of the tile in flight
        off_row = off;  off += 2 * BW_T * 4;                // lse | D of the query tile
        total = off;
    }
};

struct WinGeom {
    int S, G, L, wy, wx;
    long long tok_base;
    // window-local token -> (ly, lx) and global token (or -1: zero padding / beyond the window)
    __device__ __forceinline__ int token(int l, int& ly, int& lx) const {
        if (l >= L) {
            ly = 0;
            lx = 0;
            return -1;
        }
        ly = l / S;
        lx = l - ly * S;
        const int gy = wy * S + ly, gx = wx * S + lx;
        return (gy < G && gx < G) ? gy * G + gx : -1;
    }
};

// stage both rel-pos tables ([TWK rows each], zero beyond 2S-1) with cp.async
template <int HD>
__device__ __forceinline__ void stage_tables(__nv_bfloat16* sRel, const BwdParams& p, int tid) {
    constexpr int HDP = HD + 8, CH = HD / 8;
    const int TW = 2 * p.S - 1;
    for (int i = tid; i < 2 * p.TWK * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const bool is_w = r >= p.TWK;
        const int row = is_w ? r - p.TWK : r;
        const __nv_bfloat16* tab = is_w ? p.relw : p.relh;
        cp_async_16(smem_u32(sRel + r * HDP + c * 8), tab + (row < TW ? row : 0) * HD + c * 8, row < TW);
    }
}

// T[q][j] = log2(e) * q . rel_pos[j] for the 16 query rows of this warp (A fragments qf), both tables
template <int HD>
__device__ __forceinline__ void project_tables(const uint32_t (&qf)[HD / 16][4], const __nv_bfloat16* sRel, float* sTh,
                                               float* sTw, int TWP, int TWK, int ts, int warp, int lane) {
    constexpr int HDP = HD + 8, KSTEPS = HD / 16;
    const int g = lane >> 2, t = lane & 3;
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;
    for (int which = 0; which < 2; ++which) {
        const __nv_bfloat16* sR = sRel + which * TWK * HDP;
        float* sT = which ? sTw : sTh;
        for (int n0 = 0; n0 < TWP; n0 += 16) {
            float acc[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sR + (n0 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);   // rows < TWK: staged
                mma_bf16_16816(acc[0], qf[ks], b0, b1);
                mma_bf16_16816(acc[1], qf[ks], b2, b3);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + h * 8 + 2 * t;
                if (col < TWP) {
                    const int r0 = warp * 16 + g;
                    sT[r0 * ts + col] = acc[h][0] * kLog2eB;
                    sT[r0 * ts + col + 1] = acc[h][1] * kLog2eB;
                    sT[(r0 + 8) * ts + col] = acc[h][2] * kLog2eB;
                    sT[(r0 + 8) * ts + col + 1] = acc[h][3] * kLog2eB;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ dQ (+ lse, D)
// SN = n-tiles of the indicator products: 2 for S <= 16, 8 for S <= 64
template <int HD, int SN>
__global__ void __launch_bounds__(BW_THREADS) attn_bwd_dq_kernel(const BwdParams p) {
    using Plan = BwdSmem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    const Plan plan(p.TWP, p.TWK);
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_q);
    __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_do);
    __nv_bfloat16* sKV[2] = {reinterpret_cast<__nv_bfloat16*>(smem + plan.off_a),
                             reinterpret_cast<__nv_bfloat16*>(smem + plan.off_b)};
    __nv_bfloat16* sRel = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_rel);
    float* sTh = reinterpret_cast<float*>(smem + plan.off_th);
    float* sTw = reinterpret_cast<float*>(smem + plan.off_tw);
    short2* sKyx = reinterpret_cast<short2*>(smem + plan.off_kyx);
    float* sRow = reinterpret_cast<float*>(smem + plan.off_row);
    const int TS = plan.ts;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, head = blockIdx.y, bw = blockIdx.z;
    const int S = p.S, G = p.G;
    WinGeom w;
    w.S = S; w.G = G; w.L = S * S;
    const int win = bw % (p.nW * p.nW);
    w.wy = win / p.nW; w.wx = win % p.nW;
    w.tok_base = static_cast<long long>(bw / (p.nW * p.nW)) * G * G;
    const int L = w.L;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;
    if (w.wy * S + (qt * BW_T) / S >= G) return;            // query tile entirely below the grid: nothing to do

    // ---- prologue: Q and dO tiles, tables; D = rowsum(dO * o) straight from global memory
    for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = w.token(qt * BW_T + r, ly, lx);
        const long long tok = w.tok_base + (gt < 0 ? 0 : gt);
        cp_async_16(smem_u32(sQ + r * HDP + c * 8), qbase + tok * ld + c * 8, gt >= 0);
        cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, gt >= 0);
    }
    stage_tables<HD>(sRel, p, tid);
    cp_async_commit();
    {
        const int r = tid >> 1, half = tid & 1;
        int ly, lx;
        const int gt = w.token(qt * BW_T + r, ly, lx);
        float d = 0.f;
        if (gt >= 0) {
            const __nv_bfloat16* po = p.o + (w.tok_base + gt) * p.D + head * HD + half * (HD / 2);
            const __nv_bfloat16* pd = p.dout + (w.tok_base + gt) * p.D + head * HD + half * (HD / 2);
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                const uint4 a = *reinterpret_cast<const uint4*>(po + c * 8), b = *reinterpret_cast<const uint4*>(pd + c * 8);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    d = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(bw4[e] << 16), d);
                    d = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(bw4[e] & 0xffff0000u), d);
                }
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        if (half == 0) {
            sRow[BW_T + r] = d;
            if (gt >= 0) p.dsum[(w.tok_base + gt) * p.heads + head] = d;
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    uint32_t qf[KSTEPS][4], dof[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sQ + row * HDP + ks * 16 + koff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            ldmatrix_x4(smem_u32(sdO + row * HDP + ks * 16 + koff), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
        }
    }
    project_tables<HD>(qf, sRel, sTh, sTw, p.TWP, p.TWK, TS, warp, lane);
    __syncwarp();                                          // a warp only reads back its own 16 rows of Th / Tw

    int qyo[2], qxo[2];
    const float* thr[2];
    const float* twr[2];
    float dD[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = warp * 16 + g + h * 8;
        int ly, lx;
        w.token(qt * BW_T + r, ly, lx);
        qyo[h] = ly + S - 1;
        qxo[h] = lx + S - 1;
        thr[h] = sTh + r * TS;
        twr[h] = sTw + r * TS;
        dD[h] = sRow[BW_T + r];
    }

    auto load_kv = [&](int kt, int stage) {
        __nv_bfloat16* sK = sKV[stage];
        __nv_bfloat16* sV = sK + BW_T * HDP;
        for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            int ly, lx;
            const int gt = w.token(kt * BW_T + r, ly, lx);
            const long long row = (w.tok_base + (gt < 0 ? 0 : gt)) * ld + c * 8;
            cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, gt >= 0);
            cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, gt >= 0);
            if (c == 0) sKyx[stage * BW_T + r] = make_short2(static_cast<short>(ly), static_cast<short>(lx));
        }
    };
    const int nkt = (L + BW_T - 1) / BW_T;
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;

    // logits of one key tile (log2 domain) into s[][]; keys beyond L -> -inf
    auto logits = [&](const __nv_bfloat16* sK, const short2* kyx, int kt, float (&s)[8][4]) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
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
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = nt * 8 + 2 * t + e;
                const short2 yx = kyx[col];
                const bool live = kt * BW_T + col < L;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float bias = thr[h][qyo[h] - yx.x] + twr[h][qxo[h] - yx.y];
                    const float v = fmaf(s[nt][2 * h + e], p.scale_log2, bias);
                    s[nt][2 * h + e] = live ? v : -INFINITY;
                }
            }
        }
    };

    // ---- pass 1: row log-sum-exp (log2 domain)
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    load_kv(0, 0);
    cp_async_commit();
    for (int kt = 0; kt < nkt; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < nkt) load_kv(kt + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        float s[8][4];
        logits(sKV[stage], sKyx + stage * BW_T, kt, s);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);       // finite: every tile holds >= 1 live key
            l_run[h] *= ex2f(m_run[h] - m_new);
            m_run[h] = m_new;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            l_run[0] += ex2f(s[nt][0] - m_run[0]) + ex2f(s[nt][1] - m_run[0]);
            l_run[1] += ex2f(s[nt][2] - m_run[1]) + ex2f(s[nt][3] - m_run[1]);
        }
        __syncthreads();
    }
    float lse[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        lse[h] = m_run[h] + log2f(l_run[h]);
        if (t == 0) {
            int ly, lx;
            const int gt = w.token(qt * BW_T + warp * 16 + g + h * 8, ly, lx);
            if (gt >= 0) p.lse[(w.tok_base + gt) * p.heads + head] = lse[h];
        }
    }

    // ---- pass 2: dS and the three products with it
    float dq[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    float ah[SN][4], aw[SN][4];
#pragma unroll
    for (int i = 0; i < SN; ++i) {
        ah[i][0] = ah[i][1] = ah[i][2] = ah[i][3] = 0.f;
        aw[i][0] = aw[i][1] = aw[i][2] = aw[i][3] = 0.f;
    }
    load_kv(0, 0);
    cp_async_commit();
    for (int kt = 0; kt < nkt; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < nkt) load_kv(kt + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const __nv_bfloat16* sK = sKV[stage];
        const __nv_bfloat16* sV = sK + BW_T * HDP;
        const short2* kyx = sKyx + stage * BW_T;
        float s[8][4];
        logits(sK, kyx, kt, s);
        // dP = dO V^T
        float dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sV + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], dof[ks], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], dof[ks], b2, b3);
            }
        }
        // dS = P (dP - D), packed as A fragments
        uint32_t dsf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float d0 = ex2f(s[nt][0] - lse[0]) * (dp[nt][0] - dD[0]);
            const float d1 = ex2f(s[nt][1] - lse[0]) * (dp[nt][1] - dD[0]);
            const float d2 = ex2f(s[nt][2] - lse[1]) * (dp[nt][2] - dD[1]);
            const float d3 = ex2f(s[nt][3] - lse[1]) * (dp[nt][3] - dD[1]);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(d0, d1);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(d2, d3);
        }
        // dQ += dS K  (K tile is [key][d]: ldmatrix.trans gives the col-major B fragments)
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sK + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dq[2 * dpi], dsf[kk], b0, b1);
                    mma_bf16_16816(dq[2 * dpi + 1], dsf[kk], b2, b3);
                }
            }
        }
        // dSh += dS Eh, dSw += dS Ew with Eh[k][y] = (ky(k) == y), Ew[k][x] = (kx(k) == x) built in registers
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const short2 y0 = kyx[kk * 16 + 2 * t], y1 = kyx[kk * 16 + 2 * t + 1];
            const short2 y2 = kyx[kk * 16 + 2 * t + 8], y3 = kyx[kk * 16 + 2 * t + 9];
#pragma unroll
            for (int j = 0; j < SN; ++j) {
                const int n = j * 8 + g;
                const uint32_t h0 = (y0.x == n ? 0x3F80u : 0u) | (y1.x == n ? 0x3F800000u : 0u);
                const uint32_t h1 = (y2.x == n ? 0x3F80u : 0u) | (y3.x == n ? 0x3F800000u : 0u);
                const uint32_t w0 = (y0.y == n ? 0x3F80u : 0u) | (y1.y == n ? 0x3F800000u : 0u);
                const uint32_t w1 = (y2.y == n ? 0x3F80u : 0u) | (y3.y == n ? 0x3F800000u : 0u);
                mma_bf16_16816(ah[j], dsf[kk], h0, h1);
                mma_bf16_16816(aw[j], dsf[kk], w0, w1);
            }
        }
        __syncthreads();
    }

    // ---- rel-pos part of dQ:  dTh[q][j] = dSh[q][qy + S-1 - j],  dQ += dTh Rh + dTw Rw   (unscaled q in the forward)
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        dq[i][0] *= p.scale; dq[i][1] *= p.scale; dq[i][2] *= p.scale; dq[i][3] *= p.scale;
    }
    // park dSh / dSw in the (now dead) Th / Tw rows of this warp: [row][y], stride TS
#pragma unroll
    for (int j = 0; j < SN; ++j) {
        const int col = j * 8 + 2 * t;
        float* h0 = sTh + (warp * 16 + g) * TS;
        float* w0 = sTw + (warp * 16 + g) * TS;
        if (col + 1 < TS) {
            h0[col] = ah[j][0]; h0[col + 1] = ah[j][1];
            h0[8 * TS + col] = ah[j][2]; h0[8 * TS + col + 1] = ah[j][3];
            w0[col] = aw[j][0]; w0[col + 1] = aw[j][1];
            w0[8 * TS + col] = aw[j][2]; w0[8 * TS + col + 1] = aw[j][3];
        }
    }
    __syncwarp();
    {
        const int vrow = lane & 15;
        const int vcol = (lane >> 4) * 8;
        for (int which = 0; which < 2; ++which) {
            const float* base = which ? sTw : sTh;
            const __nv_bfloat16* sR = sRel + which * p.TWK * HDP;
            const int o0 = which ? qxo[0] : qyo[0], o1 = which ? qxo[1] : qyo[1];
            const float* r0 = base + (warp * 16 + g) * TS;
            const float* r1 = r0 + 8 * TS;
            auto pick = [&](const float* row, int off, int j) -> float {
                const int y = off - j;
                return (y >= 0 && y < S) ? row[y] : 0.f;
            };
            for (int kk = 0; kk < p.TWK / 16; ++kk) {
                const int j0 = kk * 16 + 2 * t;
                uint32_t af[4];
                af[0] = pack_bf16(pick(r0, o0, j0), pick(r0, o0, j0 + 1));
                af[1] = pack_bf16(pick(r1, o1, j0), pick(r1, o1, j0 + 1));
                af[2] = pack_bf16(pick(r0, o0, j0 + 8), pick(r0, o0, j0 + 9));
                af[3] = pack_bf16(pick(r1, o1, j0 + 8), pick(r1, o1, j0 + 9));
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sR + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dq[2 * dpi], af, b0, b1);
                    mma_bf16_16816(dq[2 * dpi + 1], af, b2, b3);
                }
            }
        }
    }

    // ---- store dQ (bf16) for the queries that exist, staged through the warp's own Q rows
    __nv_bfloat16* sO = sQ + warp * 16 * HDP;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(sO + g * HDP + i * 8 + 2 * t) = pack_bf16(dq[i][0], dq[i][1]);
        *reinterpret_cast<uint32_t*>(sO + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dq[i][2], dq[i][3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = w.token(qt * BW_T + warp * 16 + r, ly, lx);
        if (gt >= 0)
            *reinterpret_cast<uint4*>(p.dqkv + (w.tok_base + gt) * ld + head * HD + c * 8) =
                *reinterpret_cast<const uint4*>(sO + r * HDP + c * 8);
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int HD>
__global__ void __launch_bounds__(BW_THREADS) attn_bwd_dkv_kernel(const BwdParams p) {
    using Plan = BwdSmem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    const Plan plan(p.TWP, p.TWK);
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_q);
    __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_do);
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_a);
    __nv_bfloat16* sV = sK + BW_T * HDP;
    __nv_bfloat16* sRel = reinterpret_cast<__nv_bfloat16*>(smem + plan.off_rel);
    float* sTh = reinterpret_cast<float*>(smem + plan.off_th);
    float* sTw = reinterpret_cast<float*>(smem + plan.off_tw);
    short2* sQyx = reinterpret_cast<short2*>(smem + plan.off_kyx);
    float* sLse = reinterpret_cast<float*>(smem + plan.off_row);
    float* sD = sLse + BW_T;
    const int TS = plan.ts;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int kt = blockIdx.x, head = blockIdx.y, bw = blockIdx.z;
    const int S = p.S, G = p.G;
    WinGeom w;
    w.S = S; w.G = G; w.L = S * S;
    const int win = bw % (p.nW * p.nW);
    w.wy = win / p.nW; w.wx = win % p.nW;
    w.tok_base = static_cast<long long>(bw / (p.nW * p.nW)) * G * G;
    const int L = w.L;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;
    if (w.wy * S + (kt * BW_T) / S >= G) return;            // key tile entirely below the grid: constants, no gradient

    // ---- this CTA's keys and values, the tables
    for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = w.token(kt * BW_T + r, ly, lx);
        const long long row = (w.tok_base + (gt < 0 ? 0 : gt)) * ld + c * 8;
        cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, gt >= 0);
        cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, gt >= 0);
    }
    stage_tables<HD>(sRel, p, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[KSTEPS][4], vf[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sK + row * HDP + ks * 16 + koff), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
            ldmatrix_x4(smem_u32(sV + row * HDP + ks * 16 + koff), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
        }
    }
    int kyv[2], kxv[2];
    bool klive[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int l = kt * BW_T + warp * 16 + g + h * 8;
        w.token(l, kyv[h], kxv[h]);
        klive[h] = l < L;
    }
    float dk[DT][4], dv[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;
    const int nqt = (L + BW_T - 1) / BW_T;

    for (int qt = 0; qt < nqt; ++qt) {
        if (w.wy * S + (qt * BW_T) / S >= G) break;         // the remaining query tiles are padding only
        __syncthreads();                                     // previous tile fully consumed
        for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            int ly, lx;
            const int gt = w.token(qt * BW_T + r, ly, lx);
            const long long tok = w.tok_base + (gt < 0 ? 0 : gt);
            cp_async_16(smem_u32(sQ + r * HDP + c * 8), qbase + tok * ld + c * 8, gt >= 0);
            cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, gt >= 0);
            if (c == 0) {
                sQyx[r] = make_short2(static_cast<short>(ly + S - 1), static_cast<short>(lx + S - 1));
                // queries that do not exist contribute nothing: P = exp2(l - inf) = 0
                sLse[r] = gt >= 0 ? p.lse[tok * p.heads + head] : INFINITY;
                sD[r] = gt >= 0 ? p.dsum[tok * p.heads + head] : 0.f;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        {   // T of this query tile: each warp projects 16 query rows
            uint32_t qf[KSTEPS][4];
            const int row = warp * 16 + (lane & 15);
            const int koff = (lane >> 4) * 8;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
                ldmatrix_x4(smem_u32(sQ + row * HDP + ks * 16 + koff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            project_tables<HD>(qf, sRel, sTh, sTw, p.TWP, p.TWK, TS, warp, lane);
        }
        __syncthreads();
        // S^T = K Q^T and dP^T = V dO^T: rows = this warp's 16 keys, columns = the 64 queries
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sQ + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], kf[ks], b0, b1);
                mma_bf16_16816(s[2 * np + 1], kf[ks], b2, b3);
                ldmatrix_x4(smem_u32(sdO + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], vf[ks], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], vf[ks], b2, b3);
            }
        }
        uint32_t pf[4][4], dsf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = nt * 8 + 2 * t + e;
                const short2 qo = sQyx[q];
                const float lse = sLse[q], dd = sD[q];
                const float* th = sTh + q * TS;
                const float* tw = sTw + q * TS;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float bias = th[qo.x - kyv[h]] + tw[qo.y - kxv[h]];
                    const float l2 = fmaf(s[nt][2 * h + e], p.scale_log2, bias);
                    const float pr = klive[h] ? ex2f(l2 - lse) : 0.f;
                    pv[2 * h + e] = pr;
                    dsv[2 * h + e] = pr * (dp[nt][2 * h + e] - dd);
                }
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        // dV += P^T dO,  dK += dS^T Q   (tiles are [query][d]: ldmatrix.trans)
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sdO + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dv[2 * dpi], pf[kk], b0, b1);
                    mma_bf16_16816(dv[2 * dpi + 1], pf[kk], b2, b3);
                    ldmatrix_x4_trans(smem_u32(sQ + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dk[2 * dpi], dsf[kk], b0, b1);
                    mma_bf16_16816(dk[2 * dpi + 1], dsf[kk], b2, b3);
                }
            }
        }
    }

    // ---- store dK (scaled) and dV for the keys that exist, staged through this warp's K / V rows
    __syncthreads();
    __nv_bfloat16* oK = sK + warp * 16 * HDP;
    __nv_bfloat16* oV = sV + warp * 16 * HDP;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(oK + g * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][0] * p.scale, dk[i][1] * p.scale);
        *reinterpret_cast<uint32_t*>(oK + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][2] * p.scale, dk[i][3] * p.scale);
        *reinterpret_cast<uint32_t*>(oV + g * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][0], dv[i][1]);
        *reinterpret_cast<uint32_t*>(oV + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][2], dv[i][3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        int ly, lx;
        const int gt = w.token(kt * BW_T + warp * 16 + r, ly, lx);
        if (gt >= 0) {
            __nv_bfloat16* dst = p.dqkv + (w.tok_base + gt) * ld + head * HD + c * 8;
            *reinterpret_cast<uint4*>(dst + p.D) = *reinterpret_cast<const uint4*>(oK + r * HDP + c * 8);
            *reinterpret_cast<uint4*>(dst + 2 * p.D) = *reinterpret_cast<const uint4*>(oV + r * HDP + c * 8);
        }
    }
}

// ================================================================================================
// S == 64 on the 64 x 64 grid (the four global blocks of ViT-B/L/H): a query tile is grid row qy, a key tile is grid row
// ky, every token exists.  The generic kernels above spend half of their mma.sync on the rel-pos bookkeeping there
// (indicator products in dq, the full 2 x 127-column T recompute per tile pair in dkv) and their 178 KB of shared memory
// allow one CTA of four warps per SM (measured: 45.7 ms per block for 8 images, 102 TFLOP/s).  Specialised:
//   * rel_h depends on (qy - ky) only: ONE column of T_h per tile pair.  dq keeps T_h on table rows qy..qy+63 ([64][65]
//     floats) and dS's row sum over a key tile IS dSh[q][ky]; dkv projects a single table row per query tile (5 mma).
//   * rel_w: the 64 values of a row are the same for every key tile: gathered into registers once (dq); its gradient is
//     a per-column register accumulator.
//   * shared memory is re-used along the kernel (tables <-> K/V double buffer, Q/dO tiles <-> T_h / dSh, T_w <-> dSw):
//     ~101 KB (dq) / ~104 KB (dkv) -> two CTAs per SM.
template <int HD>
struct Bwd64Smem {
    static constexpr int HDP = HD + 8;
    static constexpr int KV = 4 * BW_T * HDP * 2;                 // 2 stages x (K | V); also: tables (192 rows) + dQ staging
    static constexpr int QO = 2 * BW_T * HDP * 2;                 // Q | dO tiles; later T_h / dSh [64][65] fp32
    static constexpr int TH = BW_T * 65 * 4;
    static constexpr int TW = BW_T * 129 * 4;                     // T_w [64][129] fp32; later dSw [64][65]
    static constexpr int DQ_QO = QO > TH ? QO : TH;
    static constexpr int dq_off_kv = 0, dq_off_qo = KV, dq_off_tw = KV + DQ_QO, dq_off_row = dq_off_tw + TW;
    static constexpr int dq_total = dq_off_row + 2 * BW_T * 4;
    // dkv: this CTA's K | V, the query tile's Q | dO, rel_pos_w (128 rows), 16 rows for the one rel_pos_h row, T_w, rows
    static constexpr int dkv_off_kv = 0, dkv_off_qo = 2 * BW_T * HDP * 2, dkv_off_rw = dkv_off_qo + QO;
    static constexpr int dkv_off_rh = dkv_off_rw + 128 * HDP * 2, dkv_off_tw = dkv_off_rh + 16 * HDP * 2;
    static constexpr int dkv_off_row = dkv_off_tw + TW, dkv_total = dkv_off_row + 3 * BW_T * 4;
    static_assert(192 * HDP * 2 + BW_T * HDP * 2 <= KV, "tables + dQ staging must fit the K/V double buffer");
    // dkv with the forward's T tables (attn_bwd_dkv64t): K | V, then two stages of { Q | dO, T_w tile [64][68] fp32, lse | D | T_h }
    static constexpr int TWT = BW_T * 68 * 4;
    static constexpr int dkvt_stage = QO + TWT + 3 * BW_T * 4;
    static constexpr int dkvt_off_st = 2 * BW_T * HDP * 2, dkvt_total = dkvt_off_st + 2 * dkvt_stage;
    static_assert(2 * (dkvt_total + 1024) <= 233472, "two CTAs of attn_bwd_dkv64t per SM");
};

template <int HD>
__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_dq64_kernel(const BwdParams p) {
    using Plan = Bwd64Smem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    __nv_bfloat16* sKVb = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dq_off_kv);
    __nv_bfloat16* sTab = sKVb;                                   // [0,64): rel_pos_h rows qt..qt+63, [64,192): rel_pos_w
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dq_off_qo);
    __nv_bfloat16* sdO = sQ + BW_T * HDP;
    float* sTh = reinterpret_cast<float*>(smem + Plan::dq_off_qo);   // after the fragments are in registers
    float* sTw = reinterpret_cast<float*>(smem + Plan::dq_off_tw);
    float* sRow = reinterpret_cast<float*>(smem + Plan::dq_off_row);
    constexpr int THS = 65, TWS = 129;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const long long tok_base = static_cast<long long>(b) * 4096;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;

    // ---- prologue: Q, dO, tables; D = rowsum(dO * o)
    for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const long long tok = tok_base + qt * 64 + r;
        cp_async_16(smem_u32(sQ + r * HDP + c * 8), qbase + tok * ld + c * 8, true);
        cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, true);
    }
    auto stage_tables64 = [&]() {
        for (int i = tid; i < 192 * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            const bool is_w = r >= 64;
            const int row = is_w ? r - 64 : qt + r;                 // 2S-1 = 127 rows exist
            const __nv_bfloat16* tab = is_w ? p.relw : p.relh;
            cp_async_16(smem_u32(sTab + r * HDP + c * 8), tab + (row < 127 ? row : 0) * HD + c * 8, row < 127);
        }
    };
    stage_tables64();
    cp_async_commit();
    {
        const int r = tid >> 1, half = tid & 1;
        const long long tok = tok_base + qt * 64 + r;
        const __nv_bfloat16* po = p.o + tok * p.D + head * HD + half * (HD / 2);
        const __nv_bfloat16* pd = p.dout + tok * p.D + head * HD + half * (HD / 2);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(po + c * 8), bb = *reinterpret_cast<const uint4*>(pd + c * 8);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                d = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(bw4[e] << 16), d);
                d = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(bw4[e] & 0xffff0000u), d);
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        if (half == 0) {
            sRow[r] = d;
            p.dsum[tok * p.heads + head] = d;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    uint32_t qf[KSTEPS][4], dof[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sQ + row * HDP + ks * 16 + koff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            ldmatrix_x4(smem_u32(sdO + row * HDP + ks * 16 + koff), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
        }
    }
    const float dD[2] = {sRow[warp * 16 + g], sRow[warp * 16 + g + 8]};
    __syncthreads();                                   // the Q / dO tiles are dead: their bytes become T_h
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;
    // T_h on table rows qt..qt+63 (64 columns), T_w on all 127 (128 columns)
    for (int which = 0; which < 2; ++which) {
        const __nv_bfloat16* sR = which ? sTab + 64 * HDP : sTab;
        float* sT = which ? sTw : sTh;
        const int ts = which ? TWS : THS, ncols = which ? 128 : 64;
        for (int n0 = 0; n0 < ncols; n0 += 16) {
            float acc[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sR + (n0 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(acc[0], qf[ks], b0, b1);
                mma_bf16_16816(acc[1], qf[ks], b2, b3);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + h * 8 + 2 * t, r0 = warp * 16 + g;
                sT[r0 * ts + col] = acc[h][0] * kLog2eB;
                sT[r0 * ts + col + 1] = acc[h][1] * kLog2eB;
                sT[(r0 + 8) * ts + col] = acc[h][2] * kLog2eB;
                sT[(r0 + 8) * ts + col + 1] = acc[h][3] * kLog2eB;
            }
        }
    }
    __syncwarp();                                      // a warp only reads back its own 16 rows
    // rel_w[q][kx] = T_w[q][qx - kx + 63], qx = row index inside the tile: the same for every key tile -> registers
    float twreg[8][4];
    {
        const float* tw0 = sTw + (warp * 16 + g) * TWS;
        const float* tw1 = tw0 + 8 * TWS;
        const int o0 = warp * 16 + g + 63, o1 = o0 + 8;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            twreg[nt][0] = tw0[o0 - col];
            twreg[nt][1] = tw0[o0 - col - 1];
            twreg[nt][2] = tw1[o1 - col];
            twreg[nt][3] = tw1[o1 - col - 1];
        }
    }
    float* th0p = sTh + (warp * 16 + g) * THS;
    float* th1p = th0p + 8 * THS;
    __syncthreads();                                   // tables dead: the K/V double buffer takes their place

    auto load_tile = [&](int kt, int stage, bool with_v) {
        __nv_bfloat16* sK = sKVb + stage * 2 * BW_T * HDP;
        __nv_bfloat16* sV = sK + BW_T * HDP;
        for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            const long long row = (tok_base + kt * 64 + r) * ld + c * 8;
            cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, true);
            if (with_v) cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, true);
        }
    };
    auto logits64 = [&](const __nv_bfloat16* sK, int kt, float (&s)[8][4]) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
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
        const float t0 = th0p[63 - kt], t1 = th1p[63 - kt];      // T_h column of table row qt + 63 - kt = qy - ky + 63
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = fmaf(s[nt][0], p.scale_log2, t0 + twreg[nt][0]);
            s[nt][1] = fmaf(s[nt][1], p.scale_log2, t0 + twreg[nt][1]);
            s[nt][2] = fmaf(s[nt][2], p.scale_log2, t1 + twreg[nt][2]);
            s[nt][3] = fmaf(s[nt][3], p.scale_log2, t1 + twreg[nt][3]);
        }
    };

    // ---- pass 1: row log-sum-exp (K tiles only) -- unless the training forward left it (BwdParams::lse_ready)
    float lse[2];
    if (p.lse_ready) {
#pragma unroll
        for (int h = 0; h < 2; ++h) lse[h] = p.lse[(tok_base + qt * 64 + warp * 16 + g + h * 8) * p.heads + head];
    } else {
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    load_tile(0, 0, false);
    cp_async_commit();
    for (int kt = 0; kt < 64; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < 64) load_tile(kt + 1, stage ^ 1, false);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        float s[8][4];
        logits64(sKVb + stage * 2 * BW_T * HDP, kt, s);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            l_run[h] *= ex2f(m_run[h] - m_new);
            m_run[h] = m_new;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            l_run[0] += ex2f(s[nt][0] - m_run[0]) + ex2f(s[nt][1] - m_run[0]);
            l_run[1] += ex2f(s[nt][2] - m_run[1]) + ex2f(s[nt][3] - m_run[1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        lse[h] = m_run[h] + log2f(l_run[h]);
        if (t == 0) p.lse[(tok_base + qt * 64 + warp * 16 + g + h * 8) * p.heads + head] = lse[h];
    }
    }

    // ---- pass 2
    float dq[DT][4], dsw[8][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dsw[i][0] = dsw[i][1] = dsw[i][2] = dsw[i][3] = 0.f;
    load_tile(0, 0, true);
    cp_async_commit();
    for (int kt = 0; kt < 64; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < 64) load_tile(kt + 1, stage ^ 1, true);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const __nv_bfloat16* sK = sKVb + stage * 2 * BW_T * HDP;
        const __nv_bfloat16* sV = sK + BW_T * HDP;
        float s[8][4];
        logits64(sK, kt, s);
        float dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sV + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], dof[ks], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], dof[ks], b2, b3);
            }
        }
        uint32_t dsf[4][4];
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float d0 = ex2f(s[nt][0] - lse[0]) * (dp[nt][0] - dD[0]);
            const float d1 = ex2f(s[nt][1] - lse[0]) * (dp[nt][1] - dD[0]);
            const float d2 = ex2f(s[nt][2] - lse[1]) * (dp[nt][2] - dD[1]);
            const float d3 = ex2f(s[nt][3] - lse[1]) * (dp[nt][3] - dD[1]);
            rs0 += d0 + d1;
            rs1 += d2 + d3;
            dsw[nt][0] += d0; dsw[nt][1] += d1; dsw[nt][2] += d2; dsw[nt][3] += d3;
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(d0, d1);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(d2, d3);
        }
        // dSh[q][ky = kt] = row sum of dS over this key tile; it takes the place of the T_h value just consumed
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        if (t == 0) {
            th0p[63 - kt] = rs0;
            th1p[63 - kt] = rs1;
        }
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sK + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dq[2 * dpi], dsf[kk], b0, b1);
                    mma_bf16_16816(dq[2 * dpi + 1], dsf[kk], b2, b3);
                }
            }
        }
        __syncthreads();
    }

    // ---- rel-pos part of dQ.  Tables come back into the (dead) K/V buffer; dSw goes where T_w was.
    stage_tables64();
    cp_async_commit();
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        dq[i][0] *= p.scale; dq[i][1] *= p.scale; dq[i][2] *= p.scale; dq[i][3] *= p.scale;
    }
    {
        float* w0 = sTw + (warp * 16 + g) * THS;            // dSw [64][65]
        float* w1 = w0 + 8 * THS;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            w0[col] = dsw[nt][0]; w0[col + 1] = dsw[nt][1];
            w1[col] = dsw[nt][2]; w1[col + 1] = dsw[nt][3];
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    {
        const int vrow = lane & 15;
        const int vcol = (lane >> 4) * 8;
        // h: dT_h[q][table row qt + c] = dSh[q][c]  ->  A = sTh as it is, B = rel_pos_h rows qt..qt+63
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int c0 = kk * 16 + 2 * t;
            uint32_t af[4];
            af[0] = pack_bf16(th0p[c0], th0p[c0 + 1]);
            af[1] = pack_bf16(th1p[c0], th1p[c0 + 1]);
            af[2] = pack_bf16(th0p[c0 + 8], th0p[c0 + 9]);
            af[3] = pack_bf16(th1p[c0 + 8], th1p[c0 + 9]);
#pragma unroll
            for (int dpi = 0; dpi < DT / 2; ++dpi) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(sTab + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                mma_bf16_16816(dq[2 * dpi], af, b0, b1);
                mma_bf16_16816(dq[2 * dpi + 1], af, b2, b3);
            }
        }
        // w: dT_w[q][j] = dSw[q][qx + 63 - j] (0 outside [0, 64))
        const float* w0 = sTw + (warp * 16 + g) * THS;
        const float* w1 = w0 + 8 * THS;
        const int o0 = warp * 16 + g + 63, o1 = o0 + 8;
        auto pick = [&](const float* row, int off, int j) -> float {
            const int x = off - j;
            return (x >= 0 && x < 64) ? row[x] : 0.f;
        };
#pragma unroll 2
        for (int kk = 0; kk < 8; ++kk) {
            const int j0 = kk * 16 + 2 * t;
            uint32_t af[4];
            af[0] = pack_bf16(pick(w0, o0, j0), pick(w0, o0, j0 + 1));
            af[1] = pack_bf16(pick(w1, o1, j0), pick(w1, o1, j0 + 1));
            af[2] = pack_bf16(pick(w0, o0, j0 + 8), pick(w0, o0, j0 + 9));
            af[3] = pack_bf16(pick(w1, o1, j0 + 8), pick(w1, o1, j0 + 9));
#pragma unroll
            for (int dpi = 0; dpi < DT / 2; ++dpi) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(sTab + (64 + kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                mma_bf16_16816(dq[2 * dpi], af, b0, b1);
                mma_bf16_16816(dq[2 * dpi + 1], af, b2, b3);
            }
        }
    }
    // ---- store dQ, staged behind the tables in the K/V buffer
    __nv_bfloat16* sO = sTab + 192 * HDP + warp * 16 * HDP;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(sO + g * HDP + i * 8 + 2 * t) = pack_bf16(dq[i][0], dq[i][1]);
        *reinterpret_cast<uint32_t*>(sO + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dq[i][2], dq[i][3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        *reinterpret_cast<uint4*>(p.dqkv + (tok_base + qt * 64 + warp * 16 + r) * ld + head * HD + c * 8) =
            *reinterpret_cast<const uint4*>(sO + r * HDP + c * 8);
    }
}

template <int HD>
__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_dkv64_kernel(const BwdParams p) {
    using Plan = Bwd64Smem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8;
    extern __shared__ __align__(16) unsigned char smem[];
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dkv_off_kv);
    __nv_bfloat16* sV = sK + BW_T * HDP;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dkv_off_qo);
    __nv_bfloat16* sdO = sQ + BW_T * HDP;
    __nv_bfloat16* sRw = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dkv_off_rw);
    __nv_bfloat16* sRh = reinterpret_cast<__nv_bfloat16*>(smem + Plan::dkv_off_rh);
    float* sTw = reinterpret_cast<float*>(smem + Plan::dkv_off_tw);
    float* sThc = reinterpret_cast<float*>(smem + Plan::dkv_off_row);
    float* sLse = sThc + BW_T;
    float* sD = sLse + BW_T;
    constexpr int TWS = 129;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int kt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const long long tok_base = static_cast<long long>(b) * 4096;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;

    for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const long long row = (tok_base + kt * 64 + r) * ld + c * 8;
        cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, true);
        cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, true);
    }
    for (int i = tid; i < 128 * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        cp_async_16(smem_u32(sRw + r * HDP + c * 8), p.relw + (r < 127 ? r : 0) * HD + c * 8, r < 127);
    }
    for (int i = tid; i < 8 * CH; i += BW_THREADS)                 // rows 8..15 of the rel_pos_h slot stay zero
        *reinterpret_cast<uint4*>(sRh + (8 + i / CH) * HDP + (i % CH) * 8) = make_uint4(0, 0, 0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[KSTEPS][4], vf[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sK + row * HDP + ks * 16 + koff), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
            ldmatrix_x4(smem_u32(sV + row * HDP + ks * 16 + koff), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
        }
    }
    const int kx0 = warp * 16 + g, kx1 = kx0 + 8;               // this thread's two key rows = kx
    float dk[DT][4], dv[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;

    for (int qt = 0; qt < 64; ++qt) {
        __syncthreads();                                   // previous query tile fully consumed
        for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            const long long tok = tok_base + qt * 64 + r;
            cp_async_16(smem_u32(sQ + r * HDP + c * 8), qbase + tok * ld + c * 8, true);
            cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, true);
        }
        {   // the ONE rel_pos_h row this tile pair needs: qy - ky + 63 (rows 1..7 of the slot: the following table rows, unused)
            const int j0 = qt - kt + 63;
            for (int i = tid; i < 8 * CH; i += BW_THREADS) {
                const int r = i / CH, c = i - r * CH;
                cp_async_16(smem_u32(sRh + r * HDP + c * 8), p.relh + (j0 + r < 127 ? j0 + r : 0) * HD + c * 8, j0 + r < 127);
            }
        }
        if (tid < BW_T) {
            const long long tok = tok_base + qt * 64 + tid;
            sLse[tid] = p.lse[tok * p.heads + head];
            sD[tid] = p.dsum[tok * p.heads + head];
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        {   // T_w of this query tile (each warp projects 16 query rows on all 127 table rows) + the one T_h column
            uint32_t qf[KSTEPS][4];
            const int row = warp * 16 + (lane & 15);
            const int koff = (lane >> 4) * 8;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
                ldmatrix_x4(smem_u32(sQ + row * HDP + ks * 16 + koff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            for (int n0 = 0; n0 < 128; n0 += 16) {
                float acc[2][4] = {};
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(smem_u32(sRw + (n0 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                    mma_bf16_16816(acc[0], qf[ks], b0, b1);
                    mma_bf16_16816(acc[1], qf[ks], b2, b3);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = n0 + h * 8 + 2 * t, r0 = warp * 16 + g;
                    sTw[r0 * TWS + col] = acc[h][0] * kLog2eB;
                    sTw[r0 * TWS + col + 1] = acc[h][1] * kLog2eB;
                    sTw[(r0 + 8) * TWS + col] = acc[h][2] * kLog2eB;
                    sTw[(r0 + 8) * TWS + col + 1] = acc[h][3] * kLog2eB;
                }
            }
            float acc[4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sRh + krow * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(acc, qf[ks], b0, b1);
            }
            if (t == 0) {                                  // column 0 of the n-tile = table row qy - ky + 63
                sThc[warp * 16 + g] = acc[0] * kLog2eB;
                sThc[warp * 16 + g + 8] = acc[2] * kLog2eB;
            }
        }
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sQ + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], kf[ks], b0, b1);
                mma_bf16_16816(s[2 * np + 1], kf[ks], b2, b3);
                ldmatrix_x4(smem_u32(sdO + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], vf[ks], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], vf[ks], b2, b3);
            }
        }
        uint32_t pf[4][4], dsf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = nt * 8 + 2 * t + e;             // query column = qx
                const float lse = sLse[q], dd = sD[q], thc = sThc[q];
                const float* tw = sTw + q * TWS + q + 63;
                const float l0 = fmaf(s[nt][e], p.scale_log2, thc + tw[-kx0]);
                const float l1 = fmaf(s[nt][2 + e], p.scale_log2, thc + tw[-kx1]);
                pv[e] = ex2f(l0 - lse);
                pv[2 + e] = ex2f(l1 - lse);
                dsv[e] = pv[e] * (dp[nt][e] - dd);
                dsv[2 + e] = pv[2 + e] * (dp[nt][2 + e] - dd);
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sdO + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dv[2 * dpi], pf[kk], b0, b1);
                    mma_bf16_16816(dv[2 * dpi + 1], pf[kk], b2, b3);
                    ldmatrix_x4_trans(smem_u32(sQ + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dk[2 * dpi], dsf[kk], b0, b1);
                    mma_bf16_16816(dk[2 * dpi + 1], dsf[kk], b2, b3);
                }
            }
        }
    }
    __syncthreads();
    __nv_bfloat16* oK = sK + warp * 16 * HDP;
    __nv_bfloat16* oV = sV + warp * 16 * HDP;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(oK + g * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][0] * p.scale, dk[i][1] * p.scale);
        *reinterpret_cast<uint32_t*>(oK + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][2] * p.scale, dk[i][3] * p.scale);
        *reinterpret_cast<uint32_t*>(oV + g * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][0], dv[i][1]);
        *reinterpret_cast<uint32_t*>(oV + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][2], dv[i][3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        __nv_bfloat16* dst = p.dqkv + (tok_base + kt * 64 + warp * 16 + r) * ld + head * HD + c * 8;
        *reinterpret_cast<uint4*>(dst + p.D) = *reinterpret_cast<const uint4*>(oK + r * HDP + c * 8);
        *reinterpret_cast<uint4*>(dst + 2 * p.D) = *reinterpret_cast<const uint4*>(oV + r * HDP + c * 8);
    }
}

// dK / dV of the 64 x 64 global blocks with the forward's T tables (BwdParams::tw_tab / th_tab): nothing is projected on the
// rel-pos tables here (that was a third of the kernel's MMAs), and the query tiles are double-buffered: the loads of
// tile qt + 1 (Q, dO, the [64][64] T_w tile, lse, D, the T_h column) run under the MMAs of tile qt.
template <int HD>
__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_dkv64t_kernel(const BwdParams p) {
    using Plan = Bwd64Smem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8, TWS = 68;
    extern __shared__ __align__(16) unsigned char smem[];
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sV = sK + BW_T * HDP;
    auto stQ = [&](int st) { return reinterpret_cast<__nv_bfloat16*>(smem + Plan::dkvt_off_st + st * Plan::dkvt_stage); };
    auto stTw = [&](int st) { return reinterpret_cast<float*>(smem + Plan::dkvt_off_st + st * Plan::dkvt_stage + Plan::QO); };
    auto stRow = [&](int st) { return stTw(st) + BW_T * TWS; };      // lse[64] | D[64] | T_h column[64]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int kt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const long long tok_base = static_cast<long long>(b) * 4096;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;
    const __nv_bfloat16* kbase = qbase + p.D;
    const __nv_bfloat16* vbase = qbase + 2 * p.D;

    auto load_tile = [&](int qt, int st) {
        __nv_bfloat16* sQ = stQ(st);
        __nv_bfloat16* sdO = sQ + BW_T * HDP;
        for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
            const int r = i / CH, c = i - r * CH;
            const long long tok = tok_base + qt * 64 + r;
            cp_async_16(smem_u32(sQ + r * HDP + c * 8), qbase + tok * ld + c * 8, true);
            cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, true);
        }
        float* sTw = stTw(st);
        for (int i = tid; i < BW_T * 16; i += BW_THREADS) {
            const int r = i >> 4, c = i & 15;
            const long long row = (tok_base + qt * 64 + r) * p.heads + head;
            cp_async_16(smem_u32(sTw + r * TWS + c * 4), p.tw_tab + row * 64 + c * 4, true);
        }
        if (tid < BW_T) {
            float* sRow = stRow(st);
            const long long row = (tok_base + qt * 64 + tid) * p.heads + head;
            cp_async_4(smem_u32(sRow + BW_T + tid), p.dsum + row);
            cp_async_4(smem_u32(sRow + 2 * BW_T + tid), p.th_tab + row * 64 + kt);
        }
    };

    for (int i = tid; i < BW_T * CH; i += BW_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const long long row = (tok_base + kt * 64 + r) * ld + c * 8;
        cp_async_16(smem_u32(sK + r * HDP + c * 8), kbase + row, true);
        cp_async_16(smem_u32(sV + r * HDP + c * 8), vbase + row, true);
    }
    load_tile(0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[KSTEPS][4], vf[KSTEPS][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int koff = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sK + row * HDP + ks * 16 + koff), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
            ldmatrix_x4(smem_u32(sV + row * HDP + ks * 16 + koff), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
        }
    }
    const int kx0 = warp * 16 + g, kx1 = kx0 + 8;               // this thread's two key rows = kx
    float dk[DT][4], dv[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;

    for (int qt = 0; qt < 64; ++qt) {
        const int st = qt & 1;
        // (tile qt is resident: the prologue / the previous iteration waited for it; every warp has left tile qt - 1)
        if (qt + 1 < 64) load_tile(qt + 1, st ^ 1);
        cp_async_commit();
        const __nv_bfloat16* sQ = stQ(st);
        const __nv_bfloat16* sdO = sQ + BW_T * HDP;
        const float* sTw = stTw(st);
        const float* sD = stRow(st) + BW_T;
        const float* sThc = sD + BW_T;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sQ + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], kf[ks], b0, b1);
                mma_bf16_16816(s[2 * np + 1], kf[ks], b2, b3);
                ldmatrix_x4(smem_u32(sdO + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], vf[ks], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], vf[ks], b2, b3);
            }
        }
        uint32_t pf[4][4], dsf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = nt * 8 + 2 * t + e;             // query column = qx
                const float c = sThc[q], dd = sD[q];          // (the row's log-sum-exp is already inside the T_w table)
                const float* tw = sTw + q * TWS;              // bank = (4 q + kx) mod 32: distinct over the warp
                const float l0 = fmaf(s[nt][e], p.scale_log2, c + tw[kx0]);
                const float l1 = fmaf(s[nt][2 + e], p.scale_log2, c + tw[kx1]);
                pv[e] = ex2f(l0);
                pv[2 + e] = ex2f(l1);
                dsv[e] = pv[e] * (dp[nt][e] - dd);
                dsv[2 + e] = pv[2 + e] * (dp[nt][2 + e] - dd);
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        {
            const int vrow = lane & 15;
            const int vcol = (lane >> 4) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sdO + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dv[2 * dpi], pf[kk], b0, b1);
                    mma_bf16_16816(dv[2 * dpi + 1], pf[kk], b2, b3);
                    ldmatrix_x4_trans(smem_u32(sQ + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dk[2 * dpi], dsf[kk], b0, b1);
                    mma_bf16_16816(dk[2 * dpi + 1], dsf[kk], b2, b3);
                }
            }
        }
        cp_async_wait<0>();                                  // tile qt + 1 has landed (this thread's part)
        __syncthreads();                                     // ... everybody's; and everybody has left tile qt
    }
    __nv_bfloat16* oK = sK + warp * 16 * HDP;
    __nv_bfloat16* oV = sV + warp * 16 * HDP;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        *reinterpret_cast<uint32_t*>(oK + g * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][0] * p.scale, dk[i][1] * p.scale);
        *reinterpret_cast<uint32_t*>(oK + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dk[i][2] * p.scale, dk[i][3] * p.scale);
        *reinterpret_cast<uint32_t*>(oV + g * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][0], dv[i][1]);
        *reinterpret_cast<uint32_t*>(oV + (g + 8) * HDP + i * 8 + 2 * t) = pack_bf16(dv[i][2], dv[i][3]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
        const int r = i / CH, c = i - r * CH;
        __nv_bfloat16* dst = p.dqkv + (tok_base + kt * 64 + warp * 16 + r) * ld + head * HD + c * 8;
        *reinterpret_cast<uint4*>(dst + p.D) = *reinterpret_cast<const uint4*>(oK + r * HDP + c * 8);
        *reinterpret_cast<uint4*>(dst + 2 * p.D) = *reinterpret_cast<const uint4*>(oV + r * HDP + c * 8);
    }
}

// ================================================================================================
// 14 x 14 windows: one CTA per (window, head) with everything resident.  The generic pair above cuts the 196 tokens of
// a window into 4 x 4 tiles of 64 (the fourth holds 4 rows) and launches 25 600 small CTAs per block, each a chain of
// load -> sync -> mma steps: measured 2.25 ms per block for 8 images against 0.14 ms for the forward.  Here 13 warps own
// 16 rows each of the 208-row (196 + padding) Q, K, V, dO tiles in shared memory:
//   phase A1  (rows = queries)  T = q . rel_pos for the warp's rows, row log-sum-exp (online over 13 key steps)
//   phase B   (rows = keys)     per 16-query step: S^T and dP^T blocks -> P^T, dS^T fragments -> dV += P^T dO, dK += dS^T Q
//   phase A2  (rows = queries)  per 16-key step: S and dP blocks -> dS fragment -> dQ += dS K, dSh / dSw (indicator products);
//                               then dQ += dT_h R_h + dT_w R_w
// No fragment array is ever held for a whole row: every 16 x 16 block of P / dS is consumed as the k-step it is.
constexpr int WB_S = 14, WB_L = 196, WB_LP = 208, WB_WARPS = 13, WB_THREADS = WB_WARPS * 32, WB_TS = 33;
template <int HD>
struct WinBwdSmem {
    static constexpr int HDP = HD + 8;
    static constexpr int TILE = WB_LP * HDP * 2;
    static constexpr int off_q = 0, off_k = TILE, off_v = 2 * TILE, off_do = 3 * TILE, off_rel = 4 * TILE;
    static constexpr int off_th = off_rel + 2 * 32 * HDP * 2, off_tw = off_th + WB_LP * WB_TS * 4;
    static constexpr int off_lse = off_tw + WB_LP * WB_TS * 4, off_d = off_lse + WB_LP * 4, off_yx = off_d + WB_LP * 4;
    static constexpr int off_gt = off_yx + WB_LP * 4, total = off_gt + WB_LP * 4;
    static_assert(total <= 227 * 1024, "window backward shared memory");
};

template <int HD>
__global__ void __launch_bounds__(WB_THREADS, 1) attn_bwd_win14_kernel(const BwdParams p) {
    using Plan = WinBwdSmem<HD>;
    constexpr int HDP = Plan::HDP, KSTEPS = HD / 16, DT = HD / 8, CH = HD / 8, NP = WB_LP / 16;
    extern __shared__ __align__(16) unsigned char smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + Plan::off_q);
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + Plan::off_k);
    __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(smem + Plan::off_v);
    __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(smem + Plan::off_do);
    __nv_bfloat16* sRel = reinterpret_cast<__nv_bfloat16*>(smem + Plan::off_rel);     // [0,32): rel_pos_h, [32,64): rel_pos_w
    float* sTh = reinterpret_cast<float*>(smem + Plan::off_th);
    float* sTw = reinterpret_cast<float*>(smem + Plan::off_tw);
    float* sLse = reinterpret_cast<float*>(smem + Plan::off_lse);
    float* sD = reinterpret_cast<float*>(smem + Plan::off_d);
    short2* sYx = reinterpret_cast<short2*>(smem + Plan::off_yx);
    int* sGt = reinterpret_cast<int*>(smem + Plan::off_gt);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int head = blockIdx.x, bw = blockIdx.y;
    const int G = p.G, win = bw % (p.nW * p.nW);
    const int wy = win / p.nW, wx = win % p.nW;
    const long long tok_base = static_cast<long long>(bw / (p.nW * p.nW)) * G * G;
    const int ld = 3 * p.D;
    const __nv_bfloat16* qbase = p.qkv + head * HD;

    // ---- phase 0: the window's Q, K, V, dO rows (zero for padding), tables, D = rowsum(dO * o)
    for (int r = tid; r < WB_LP; r += WB_THREADS) {
        int ly = 0, lx = 0, gt = -1;
        if (r < WB_L) {
            ly = r / WB_S;
            lx = r - ly * WB_S;
            const int gy = wy * WB_S + ly, gx = wx * WB_S + lx;
            if (gy < G && gx < G) gt = gy * G + gx;
        }
        sYx[r] = make_short2(static_cast<short>(ly), static_cast<short>(lx));
        sGt[r] = gt;
    }
    for (int i = tid; i < WB_LP * CH; i += WB_THREADS) {
        const int r = i / CH, c = i - r * CH;
        int gt = -1;
        if (r < WB_L) {
            const int ly = r / WB_S, lx = r - ly * WB_S;
            const int gy = wy * WB_S + ly, gx = wx * WB_S + lx;
            if (gy < G && gx < G) gt = gy * G + gx;
        }
        const long long tok = tok_base + (gt < 0 ? 0 : gt);
        const __nv_bfloat16* src = qbase + tok * ld + c * 8;
        cp_async_16(smem_u32(sQ + r * HDP + c * 8), src, gt >= 0);
        cp_async_16(smem_u32(sK + r * HDP + c * 8), src + p.D, gt >= 0);
        cp_async_16(smem_u32(sV + r * HDP + c * 8), src + 2 * p.D, gt >= 0);
        cp_async_16(smem_u32(sdO + r * HDP + c * 8), p.dout + tok * p.D + head * HD + c * 8, gt >= 0);
    }
    for (int i = tid; i < 64 * CH; i += WB_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const bool is_w = r >= 32;
        const int row = is_w ? r - 32 : r;
        const __nv_bfloat16* tab = is_w ? p.relw : p.relh;
        cp_async_16(smem_u32(sRel + r * HDP + c * 8), tab + (row < 27 ? row : 0) * HD + c * 8, row < 27);
    }
    cp_async_commit();
    {
        const int r = tid >> 1, half = tid & 1;                   // 416 threads = 2 per row
        int gt = -1;
        if (r < WB_L) {
            const int ly = r / WB_S, lx = r - ly * WB_S;
            const int gy = wy * WB_S + ly, gx = wx * WB_S + lx;
            if (gy < G && gx < G) gt = gy * G + gx;
        }
        float d = 0.f;
        if (gt >= 0) {
            const __nv_bfloat16* po = p.o + (tok_base + gt) * p.D + head * HD + half * (HD / 2);
            const __nv_bfloat16* pd = p.dout + (tok_base + gt) * p.D + head * HD + half * (HD / 2);
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                const uint4 a = *reinterpret_cast<const uint4*>(po + c * 8), b = *reinterpret_cast<const uint4*>(pd + c * 8);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    d = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(bw4[e] << 16), d);
                    d = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(bw4[e] & 0xffff0000u), d);
                }
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        if (half == 0) sD[r] = d;
    }
    cp_async_wait<0>();
    __syncthreads();

    const int krow = (lane & 7) + ((lane >> 4) << 3);
    const int kcol = ((lane >> 3) & 1) * 8;
    const int vrow = lane & 15;
    const int vcol = (lane >> 4) * 8;
    const int arow = warp * 16 + (lane & 15);
    const int akoff = (lane >> 4) * 8;
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const short2 yx0 = sYx[r0], yx1 = sYx[r1];

    // ---- phase A1: T rows and the row log-sum-exp of this warp's 16 queries
    uint32_t qf[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
        ldmatrix_x4(smem_u32(sQ + arow * HDP + ks * 16 + akoff), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    for (int which = 0; which < 2; ++which) {
        const __nv_bfloat16* sR = sRel + which * 32 * HDP;
        float* sT = which ? sTw : sTh;
#pragma unroll
        for (int n0 = 0; n0 < 32; n0 += 16) {
            float acc[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sR + (n0 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(acc[0], qf[ks], b0, b1);
                mma_bf16_16816(acc[1], qf[ks], b2, b3);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + h * 8 + 2 * t;
                sT[r0 * WB_TS + col] = acc[h][0] * kLog2eB;
                sT[r0 * WB_TS + col + 1] = acc[h][1] * kLog2eB;
                sT[r1 * WB_TS + col] = acc[h][2] * kLog2eB;
                sT[r1 * WB_TS + col + 1] = acc[h][3] * kLog2eB;
            }
        }
    }
    __syncwarp();
    const float* th0 = sTh + r0 * WB_TS + yx0.x + 13;
    const float* th1 = sTh + r1 * WB_TS + yx1.x + 13;
    const float* tw0 = sTw + r0 * WB_TS + yx0.y + 13;
    const float* tw1 = sTw + r1 * WB_TS + yx1.y + 13;
    // logits (log2 domain) of the 16-key step np for this thread's 2 rows x 4 columns; keys >= 196 -> -inf
    auto logit_block = [&](const float (&acc)[2][4], int np, float (&l)[2][4]) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = np * 16 + h2 * 8 + 2 * t + e;
                const short2 k = sYx[col];
                const bool live = col < WB_L;
                const float v0 = fmaf(acc[h2][e], p.scale_log2, th0[-k.x] + tw0[-k.y]);
                const float v1 = fmaf(acc[h2][2 + e], p.scale_log2, th1[-k.x] + tw1[-k.y]);
                l[h2][e] = live ? v0 : -INFINITY;
                l[h2][2 + e] = live ? v1 : -INFINITY;
            }
    };
    float lse[2];
    if (p.lse_ready) {          // left by the training forward; rows that are not tokens of the grid take no part (P = 0)
        const int gt0 = sGt[r0], gt1 = sGt[r1];
        lse[0] = gt0 >= 0 ? p.lse[(tok_base + gt0) * p.heads + head] : INFINITY;
        lse[1] = gt1 >= 0 ? p.lse[(tok_base + gt1) * p.heads + head] : INFINITY;
        if (t == 0) {
            sLse[r0] = lse[0];
            sLse[r1] = lse[1];
        }
    } else {
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll 1
        for (int np = 0; np < NP; ++np) {
            float acc[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sK + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(acc[0], qf[ks], b0, b1);
                mma_bf16_16816(acc[1], qf[ks], b2, b3);
            }
            float l[2][4];
            logit_block(acc, np, l);
            float mx0 = fmaxf(fmaxf(l[0][0], l[0][1]), fmaxf(l[1][0], l[1][1]));
            float mx1 = fmaxf(fmaxf(l[0][2], l[0][3]), fmaxf(l[1][2], l[1][3]));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float n0 = fmaxf(m_run[0], mx0), n1 = fmaxf(m_run[1], mx1);     // finite from the first step on
            l_run[0] = l_run[0] * ex2f(m_run[0] - n0) + ex2f(l[0][0] - n0) + ex2f(l[0][1] - n0) + ex2f(l[1][0] - n0) + ex2f(l[1][1] - n0);
            l_run[1] = l_run[1] * ex2f(m_run[1] - n1) + ex2f(l[0][2] - n1) + ex2f(l[0][3] - n1) + ex2f(l[1][2] - n1) + ex2f(l[1][3] - n1);
            m_run[0] = n0;
            m_run[1] = n1;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
            lse[h] = m_run[h] + log2f(l_run[h]);
        }
        if (t == 0) {
            sLse[r0] = r0 < WB_L ? lse[0] : INFINITY;             // rows beyond the window: P = exp2(l - inf) = 0 in phase B
            sLse[r1] = r1 < WB_L ? lse[1] : INFINITY;
        }
    }
    __syncthreads();

    // ---- phase B: this warp's 16 KEY rows
    {
        uint32_t kf[KSTEPS][4], vf[KSTEPS][4];
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            ldmatrix_x4(smem_u32(sK + arow * HDP + ks * 16 + akoff), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
            ldmatrix_x4(smem_u32(sV + arow * HDP + ks * 16 + akoff), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
        }
        float dk[DT][4], dv[DT][4];
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
            dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
        }
        const bool kl0 = r0 < WB_L, kl1 = r1 < WB_L;
#pragma unroll 1
        for (int np = 0; np < NP; ++np) {
            float as[2][4] = {}, ap[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sQ + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(as[0], kf[ks], b0, b1);
                mma_bf16_16816(as[1], kf[ks], b2, b3);
                ldmatrix_x4(smem_u32(sdO + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(ap[0], vf[ks], b0, b1);
                mma_bf16_16816(ap[1], vf[ks], b2, b3);
            }
            uint32_t pf[4], dsf[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                float pv[4], dsv[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = np * 16 + h2 * 8 + 2 * t + e;        // query column
                    const short2 qy = sYx[q];
                    const float lq = sLse[q], dd = sD[q];
                    const float* thq = sTh + q * WB_TS + qy.x + 13;
                    const float* twq = sTw + q * WB_TS + qy.y + 13;
                    const float l0 = fmaf(as[h2][e], p.scale_log2, thq[-yx0.x] + twq[-yx0.y]);
                    const float l1 = fmaf(as[h2][2 + e], p.scale_log2, thq[-yx1.x] + twq[-yx1.y]);
                    pv[e] = kl0 ? ex2f(l0 - lq) : 0.f;
                    pv[2 + e] = kl1 ? ex2f(l1 - lq) : 0.f;
                    dsv[e] = pv[e] * (ap[h2][e] - dd);
                    dsv[2 + e] = pv[2 + e] * (ap[h2][2 + e] - dd);
                }
                pf[h2 * 2 + 0] = pack_bf16(pv[0], pv[1]);
                pf[h2 * 2 + 1] = pack_bf16(pv[2], pv[3]);
                dsf[h2 * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
                dsf[h2 * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
            }
#pragma unroll
            for (int dpi = 0; dpi < DT / 2; ++dpi) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(sdO + (np * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                mma_bf16_16816(dv[2 * dpi], pf, b0, b1);
                mma_bf16_16816(dv[2 * dpi + 1], pf, b2, b3);
                ldmatrix_x4_trans(smem_u32(sQ + (np * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                mma_bf16_16816(dk[2 * dpi], dsf, b0, b1);
                mma_bf16_16816(dk[2 * dpi + 1], dsf, b2, b3);
            }
        }
        const int gt0 = sGt[r0], gt1 = sGt[r1];
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            if (gt0 >= 0) {
                __nv_bfloat16* dst = p.dqkv + (tok_base + gt0) * ld + head * HD + i * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(dst + p.D) = pack_bf16(dk[i][0] * p.scale, dk[i][1] * p.scale);
                *reinterpret_cast<uint32_t*>(dst + 2 * p.D) = pack_bf16(dv[i][0], dv[i][1]);
            }
            if (gt1 >= 0) {
                __nv_bfloat16* dst = p.dqkv + (tok_base + gt1) * ld + head * HD + i * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(dst + p.D) = pack_bf16(dk[i][2] * p.scale, dk[i][3] * p.scale);
                *reinterpret_cast<uint32_t*>(dst + 2 * p.D) = pack_bf16(dv[i][2], dv[i][3]);
            }
        }
    }
    __syncthreads();             // every warp is done with the T rows of the others: they may be overwritten below

    // ---- phase A2: this warp's 16 QUERY rows again
    {
        uint32_t dof[KSTEPS][4];
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
            ldmatrix_x4(smem_u32(sdO + arow * HDP + ks * 16 + akoff), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
        const float d0 = sD[r0], d1 = sD[r1];
        float dq[DT][4], ah[2][4], aw[2][4];
#pragma unroll
        for (int i = 0; i < DT; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ah[i][0] = ah[i][1] = ah[i][2] = ah[i][3] = 0.f;
            aw[i][0] = aw[i][1] = aw[i][2] = aw[i][3] = 0.f;
        }
#pragma unroll 1
        for (int np = 0; np < NP; ++np) {
            float as[2][4] = {}, ap[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(sK + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(as[0], qf[ks], b0, b1);
                mma_bf16_16816(as[1], qf[ks], b2, b3);
                ldmatrix_x4(smem_u32(sV + (np * 16 + krow) * HDP + ks * 16 + kcol), b0, b1, b2, b3);
                mma_bf16_16816(ap[0], dof[ks], b0, b1);
                mma_bf16_16816(ap[1], dof[ks], b2, b3);
            }
            float l[2][4];
            logit_block(as, np, l);
            uint32_t dsf[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const float e0 = ex2f(l[h2][0] - lse[0]) * (ap[h2][0] - d0), e1 = ex2f(l[h2][1] - lse[0]) * (ap[h2][1] - d0);
                const float e2 = ex2f(l[h2][2] - lse[1]) * (ap[h2][2] - d1), e3 = ex2f(l[h2][3] - lse[1]) * (ap[h2][3] - d1);
                dsf[h2 * 2 + 0] = pack_bf16(e0, e1);
                dsf[h2 * 2 + 1] = pack_bf16(e2, e3);
            }
#pragma unroll
            for (int dpi = 0; dpi < DT / 2; ++dpi) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(sK + (np * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                mma_bf16_16816(dq[2 * dpi], dsf, b0, b1);
                mma_bf16_16816(dq[2 * dpi + 1], dsf, b2, b3);
            }
            const short2 y0 = sYx[np * 16 + 2 * t], y1 = sYx[np * 16 + 2 * t + 1];
            const short2 y2 = sYx[np * 16 + 2 * t + 8], y3 = sYx[np * 16 + 2 * t + 9];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = j * 8 + g;
                const uint32_t h0 = (y0.x == n ? 0x3F80u : 0u) | (y1.x == n ? 0x3F800000u : 0u);
                const uint32_t h1 = (y2.x == n ? 0x3F80u : 0u) | (y3.x == n ? 0x3F800000u : 0u);
                const uint32_t w0 = (y0.y == n ? 0x3F80u : 0u) | (y1.y == n ? 0x3F800000u : 0u);
                const uint32_t w1 = (y2.y == n ? 0x3F80u : 0u) | (y3.y == n ? 0x3F800000u : 0u);
                mma_bf16_16816(ah[j], dsf, h0, h1);
                mma_bf16_16816(aw[j], dsf, w0, w1);
            }
        }
        // (keys 196..207 have dS = 0, so their (0, 0) coordinates add nothing to dSh / dSw)
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            dq[i][0] *= p.scale; dq[i][1] *= p.scale; dq[i][2] *= p.scale; dq[i][3] *= p.scale;
        }
        // dSh / dSw into this warp's own (now private) T rows, then dQ += dT_h R_h + dT_w R_w with dT[q][j] = dS*[q][q* + 13 - j]
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = j * 8 + 2 * t;
            sTh[r0 * WB_TS + col] = ah[j][0]; sTh[r0 * WB_TS + col + 1] = ah[j][1];
            sTh[r1 * WB_TS + col] = ah[j][2]; sTh[r1 * WB_TS + col + 1] = ah[j][3];
            sTw[r0 * WB_TS + col] = aw[j][0]; sTw[r0 * WB_TS + col + 1] = aw[j][1];
            sTw[r1 * WB_TS + col] = aw[j][2]; sTw[r1 * WB_TS + col + 1] = aw[j][3];
        }
        __syncwarp();
        for (int which = 0; which < 2; ++which) {
            const float* row0 = (which ? sTw : sTh) + r0 * WB_TS;
            const float* row1 = (which ? sTw : sTh) + r1 * WB_TS;
            const __nv_bfloat16* sR = sRel + which * 32 * HDP;
            const int o0 = (which ? yx0.y : yx0.x) + 13, o1 = (which ? yx1.y : yx1.x) + 13;
            auto pick = [&](const float* row, int off, int j) -> float {
                const int y = off - j;
                return (y >= 0 && y < WB_S) ? row[y] : 0.f;
            };
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int j0 = kk * 16 + 2 * t;
                uint32_t af[4];
                af[0] = pack_bf16(pick(row0, o0, j0), pick(row0, o0, j0 + 1));
                af[1] = pack_bf16(pick(row1, o1, j0), pick(row1, o1, j0 + 1));
                af[2] = pack_bf16(pick(row0, o0, j0 + 8), pick(row0, o0, j0 + 9));
                af[3] = pack_bf16(pick(row1, o1, j0 + 8), pick(row1, o1, j0 + 9));
#pragma unroll
                for (int dpi = 0; dpi < DT / 2; ++dpi) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(smem_u32(sR + (kk * 16 + vrow) * HDP + dpi * 16 + vcol), b0, b1, b2, b3);
                    mma_bf16_16816(dq[2 * dpi], af, b0, b1);
                    mma_bf16_16816(dq[2 * dpi + 1], af, b2, b3);
                }
            }
        }
        const int gt0 = sGt[r0], gt1 = sGt[r1];
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            if (gt0 >= 0)
                *reinterpret_cast<uint32_t*>(p.dqkv + (tok_base + gt0) * ld + head * HD + i * 8 + 2 * t) = pack_bf16(dq[i][0], dq[i][1]);
            if (gt1 >= 0)
                *reinterpret_cast<uint32_t*>(p.dqkv + (tok_base + gt1) * ld + head * HD + i * 8 + 2 * t) = pack_bf16(dq[i][2], dq[i][3]);
        }
    }
}

template <int HD>
int launch_bwd(const BwdParams& p, int batch, cudaStream_t stream) {
    if (p.S == 64 && p.G == 64 && HD >= 64 && option(OPT_ATTN_LEGACY) == 0) {   // the 64 x 64 global blocks: specialised kernels
        static unsigned char g0[kMaxDevices] = {}, g1[kMaxDevices] = {};
        dim3 grid64(64, p.heads, batch);
        const bool tabs = p.tw_tab != nullptr && p.th_tab != nullptr && p.lse_ready;
        // tcgen05 kernels (attention_bwd_tc.cu) when the training forward left its tables; cvlm_set_option("bwd_tc", v):
        // 0 = mma.sync kernels only, 1 = dK / dV on tcgen05, 2 (default) = dQ as well
        const int tc = tabs ? option(OPT_BWD_TC) : 0;
        if (tc >= 2) {
            CVLM_TRY(attention_bwd_dq_tc_launch(p.qkv, p.o, p.dout, p.dqkv, p.dsum, p.tw_tab, p.th_tab, p.relh, p.relw, batch,
                                                p.heads, HD, stream));
        } else {
            CVLM_TRY(ensure_dynamic_smem(attn_bwd_dq64_kernel<HD>, Bwd64Smem<HD>::dq_total, g0));
            attn_bwd_dq64_kernel<HD><<<grid64, BW_THREADS, Bwd64Smem<HD>::dq_total, stream>>>(p);
            CVLM_CUDA(cudaGetLastError());
        }
        if (tc >= 1)
            return attention_bwd_dkv_tc_launch(p.qkv, p.dout, p.dqkv, p.dsum, p.tw_tab, p.th_tab, batch, p.heads, HD, stream);
        if (p.tw_tab != nullptr && p.th_tab != nullptr && p.lse_ready) {
            static unsigned char g2[kMaxDevices] = {};
            CVLM_TRY(ensure_dynamic_smem(attn_bwd_dkv64t_kernel<HD>, Bwd64Smem<HD>::dkvt_total, g2));
            attn_bwd_dkv64t_kernel<HD><<<grid64, BW_THREADS, Bwd64Smem<HD>::dkvt_total, stream>>>(p);
        } else {
            CVLM_TRY(ensure_dynamic_smem(attn_bwd_dkv64_kernel<HD>, Bwd64Smem<HD>::dkv_total, g1));
            attn_bwd_dkv64_kernel<HD><<<grid64, BW_THREADS, Bwd64Smem<HD>::dkv_total, stream>>>(p);
        }
        CVLM_CUDA(cudaGetLastError());
        return CVLM_OK;
    }
    if (p.S == WB_S && HD >= 64 && option(OPT_ATTN_LEGACY) == 0) {   // 14 x 14 windows: one CTA per (window, head)
        static unsigned char gw[kMaxDevices] = {};
        CVLM_TRY(ensure_dynamic_smem(attn_bwd_win14_kernel<HD>, WinBwdSmem<HD>::total, gw));
        dim3 gridw(p.heads, batch * p.nW * p.nW);
        attn_bwd_win14_kernel<HD><<<gridw, WB_THREADS, WinBwdSmem<HD>::total, stream>>>(p);
        CVLM_CUDA(cudaGetLastError());
        return CVLM_OK;
    }
    const BwdSmem<HD> plan(p.TWP, p.TWK);
    if (plan.total > 227 * 1024) return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: shared memory %d", plan.total);
    const int L = p.S * p.S;
    dim3 grid((L + BW_T - 1) / BW_T, p.heads, batch * p.nW * p.nW);
    static unsigned char f0[kMaxDevices] = {}, f1[kMaxDevices] = {}, f2[kMaxDevices] = {};
    // the opt-in size depends on S: always raise it to the maximum this head_dim can need (S = 64)
    const int max_smem = BwdSmem<HD>(128, 128).total;
    if (p.S <= 16) {
        CVLM_TRY(ensure_dynamic_smem(attn_bwd_dq_kernel<HD, 2>, max_smem, f0));
        attn_bwd_dq_kernel<HD, 2><<<grid, BW_THREADS, plan.total, stream>>>(p);
    } else {
        CVLM_TRY(ensure_dynamic_smem(attn_bwd_dq_kernel<HD, 8>, max_smem, f1));
        attn_bwd_dq_kernel<HD, 8><<<grid, BW_THREADS, plan.total, stream>>>(p);
    }
    CVLM_CUDA(cudaGetLastError());
    CVLM_TRY(ensure_dynamic_smem(attn_bwd_dkv_kernel<HD>, max_smem, f2));
    attn_bwd_dkv_kernel<HD><<<grid, BW_THREADS, plan.total, stream>>>(p);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace

// dqkv = d(loss)/d(qkv) given dout = d(loss)/d(attention output); lse / dsum are fp32 [batch*grid^2, heads] scratch.
// Every element of dqkv that belongs to a token of the grid is written (2 launches).
int attention_backward_launch(const void* qkv, const void* o, const void* dout, void* dqkv, float* lse, float* dsum,
                              const void* relh, const void* relw, int batch, int grid_side, int window, int heads,
                              int head_dim, cudaStream_t stream, bool lse_ready, const float* tw_tab, const float* th_tab) {
    if (!qkv || !o || !dout || !dqkv || !lse || !dsum || !relh || !relw)
        return set_error(CVLM_ERR_ARG, "attention_backward: null pointer");
    if (batch <= 0 || grid_side <= 0 || heads <= 0) return set_error(CVLM_ERR_ARG, "attention_backward: bad sizes");
    BwdParams p;
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.o = static_cast<const __nv_bfloat16*>(o);
    p.dout = static_cast<const __nv_bfloat16*>(dout);
    p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    p.lse = lse;
    p.lse_ready = lse_ready ? 1 : 0;
    p.tw_tab = tw_tab;
    p.th_tab = th_tab;
    p.dsum = dsum;
    p.relh = static_cast<const __nv_bfloat16*>(relh);
    p.relw = static_cast<const __nv_bfloat16*>(relw);
    p.G = grid_side;
    p.S = window > 0 ? window : grid_side;
    p.nW = (p.G + p.S - 1) / p.S;
    p.heads = heads;
    p.D = heads * head_dim;
    p.TWP = ((2 * p.S - 1) + 7) / 8 * 8;
    p.TWK = (p.TWP + 15) / 16 * 16;
    p.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    p.scale_log2 = p.scale * kLog2eB;
    if (p.S > 64) return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: window / grid side > 64 tokens");
    if (static_cast<long long>(batch) * p.nW * p.nW > 65535)
        return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: batch * windows exceeds grid.z; split the batch");
    if (head_dim == 80) return launch_bwd<80>(p, batch, stream);
    if (head_dim == 64) return launch_bwd<64>(p, batch, stream);
    if (head_dim == 32) return launch_bwd<32>(p, batch, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: head_dim must be 32, 64 or 80");
}

}  // namespace cvlm
