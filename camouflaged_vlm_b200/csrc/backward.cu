// Side kernels of the adapter-only backward pass (train_ovcos_maskdecoder_edge.py:301-303 trains only
// prompt_generator.*, but the gradient crosses every frozen block: models/sam_maskdecoder_edge.py:429-460):
//   layernorm_train   nn.LayerNorm forward that also keeps the normalised input (bf16) and 1/std per row
//   ln_bwd            LayerNorm backward for one row per warp:  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),
//                     g = dy * gamma (or gamma already folded into the GEMM that produced g), plus the shortcut gradient
//   wgrad_tn          weight gradients: out[i, j] (+)= sum_m A[m, i] * B[m, j] with both operands row-major in m
//                     (activations), mma.sync + ldmatrix.trans, split over m with one partial per split (deterministic),
//                     an optional column of ones appended to B gives the bias gradient in the same pass
//   eg_wgrad          weight gradient of embedding_generator through the reference's scrambled view of the tokens
//   small helpers     fp32 add, partial reduction
// All HBM-bound streaming kernels: 128-bit accesses, fp32 statistics.
#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

namespace {

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int LNB_VMAX = 16;      // float4 per lane -> cols <= 2048

__global__ void __launch_bounds__(256)
layernorm_train_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ xhat, float* __restrict__ rstd_out,
                       long long rows, int cols, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = cols >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
    float4 v[LNB_VMAX];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LNB_VMAX; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = __ldcs(xr + idx);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum_b(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LNB_VMAX; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum_b(q) / cols + eps);
    if (lane == 0) rstd_out[row] = rstd;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < LNB_VMAX; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float4 gg = __ldg(g4 + idx), bb = __ldg(b4 + idx);
            float4 h;
            h.x = (v[i].x - mean) * rstd; h.y = (v[i].y - mean) * rstd;
            h.z = (v[i].z - mean) * rstd; h.w = (v[i].w - mean) * rstd;
            reinterpret_cast<uint2*>(xhat + row * cols)[idx] = make_uint2(pack_bf16(h.x, h.y), pack_bf16(h.z, h.w));
            reinterpret_cast<uint2*>(out + row * cols)[idx] =
                make_uint2(pack_bf16(h.x * gg.x + bb.x, h.y * gg.y + bb.y), pack_bf16(h.z * gg.z + bb.z, h.w * gg.w + bb.w));
        }
    }
}

// XH: where the normalised input comes from -- 0: raw bf16 x + (rstd, -mean*rstd) rows (norm1 fold), 1: saved bf16
// xhat + rstd per row (norm2), 2: fp32 x, statistics recomputed here (neck LayerNorm2d)
// NV > 0: cols == NV * 128 exactly (fully unrolled, no guards: the generic version keeps 2 x 16 float4 slots live and ran at
// a third of the copy bandwidth); NV == 0: any width that is a multiple of 4
template <int XH, int NV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ g_in, const float* __restrict__ gamma, const void* __restrict__ xsrc,
              const float* __restrict__ stat, const float* residual, float* dx, __nv_bfloat16* __restrict__ dxb,
              long long rows, int cols, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = cols >> 2;
    constexpr int SLOTS = NV > 0 ? NV : LNB_VMAX;
    float4 gv[SLOTS], xh[SLOTS];
    float rstd = 0.f, nm = 0.f;
    if (XH == 0) {
        const float2 rn = __ldg(reinterpret_cast<const float2*>(stat) + row);
        rstd = rn.x;
        nm = rn.y;
    } else if (XH == 1) {
        rstd = __ldg(stat + row);
    }
    // ---- normalised input
    if (XH == 2) {
        const float4* xr = reinterpret_cast<const float4*>(static_cast<const float*>(xsrc) + row * cols);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
            const int idx = lane + i * 32;
            if (NV > 0 || idx < nvec) {
                xh[i] = __ldcs(xr + idx);
                s += (xh[i].x + xh[i].y) + (xh[i].z + xh[i].w);
            }
        }
        const float mean = warp_sum_b(s) / cols;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
            const int idx = lane + i * 32;
            if (NV > 0 || idx < nvec) {
                xh[i].x -= mean; xh[i].y -= mean; xh[i].z -= mean; xh[i].w -= mean;
                q += (xh[i].x * xh[i].x + xh[i].y * xh[i].y) + (xh[i].z * xh[i].z + xh[i].w * xh[i].w);
            }
        }
        rstd = 1.0f / sqrtf(warp_sum_b(q) / cols + eps);
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
            xh[i].x *= rstd; xh[i].y *= rstd; xh[i].z *= rstd; xh[i].w *= rstd;
        }
    } else {
        const uint2* xr = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(xsrc) + row * cols);
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
            const int idx = lane + i * 32;
            if (NV > 0 || idx < nvec) {
                const uint2 pk = __ldcs(xr + idx);
                float4 h = make_float4(__uint_as_float(pk.x << 16), __uint_as_float(pk.x & 0xffff0000u),
                                       __uint_as_float(pk.y << 16), __uint_as_float(pk.y & 0xffff0000u));
                if (XH == 0) {
                    h.x = fmaf(h.x, rstd, nm); h.y = fmaf(h.y, rstd, nm); h.z = fmaf(h.z, rstd, nm); h.w = fmaf(h.w, rstd, nm);
                }
                xh[i] = h;
            }
        }
    }
    // ---- g = dy (* gamma), its mean and its mean against xhat
    const float4* gr = reinterpret_cast<const float4*>(g_in + row * cols);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
        const int idx = lane + i * 32;
        if (NV > 0 || idx < nvec) {
            float4 g = __ldcs(gr + idx);
            if (gamma != nullptr) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
                g.x *= w.x; g.y *= w.y; g.z *= w.z; g.w *= w.w;
            }
            gv[i] = g;
            s1 += (g.x + g.y) + (g.z + g.w);
            s2 += (g.x * xh[i].x + g.y * xh[i].y) + (g.z * xh[i].z + g.w * xh[i].w);
        }
    }
    const float m1 = warp_sum_b(s1) / cols, m2 = warp_sum_b(s2) / cols;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
        const int idx = lane + i * 32;
        if (NV > 0 || idx < nvec) {
            float4 d;
            d.x = rstd * (gv[i].x - m1 - xh[i].x * m2);
            d.y = rstd * (gv[i].y - m1 - xh[i].y * m2);
            d.z = rstd * (gv[i].z - m1 - xh[i].z * m2);
            d.w = rstd * (gv[i].w - m1 - xh[i].w * m2);
            if (residual != nullptr) {
                const float4 r = reinterpret_cast<const float4*>(residual + row * cols)[idx];
                d.x += r.x; d.y += r.y; d.z += r.z; d.w += r.w;
            }
            reinterpret_cast<float4*>(dx + row * cols)[idx] = d;
            if (dxb != nullptr)
                reinterpret_cast<uint2*>(dxb + row * cols)[idx] = make_uint2(pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
        }
    }
}

// ------------------------------------------------------------------------------------------------ wgrad (TN)
constexpr int WG_I = 128, WG_M = 32, WG_THREADS = 128;   // static shared memory: 2 stages x (A 8.5 KB + B <= 4.5 KB)

// partial[split][i][j] (+)= sum over this split's rows m of A[m, i0 + i] * B[m, j];  column `ones_col` of B reads as 1
template <int JP>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_tn_kernel(const __nv_bfloat16* __restrict__ A, long long lda, int I, const __nv_bfloat16* __restrict__ B, long long ldb,
                int J, int ones_col, long long M, long long rows_per_split, float* __restrict__ partial, int accumulate) {
    constexpr int LDA = WG_I + 8, LDB = JP + 8, NT = JP / 8;
    __shared__ __align__(16) __nv_bfloat16 sA[2][WG_M * LDA];
    __shared__ __align__(16) __nv_bfloat16 sB[2][WG_M * LDB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int i0 = blockIdx.x * WG_I;
    const long long m_begin = static_cast<long long>(blockIdx.y) * rows_per_split;
    const long long m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
    float acc[2][NT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[a][n][0] = acc[a][n][1] = acc[a][n][2] = acc[a][n][3] = 0.f;

    auto load = [&](long long m0, int stage) {
        for (int i = tid; i < WG_M * (WG_I / 8); i += WG_THREADS) {
            const int r = i / (WG_I / 8), c = i - r * (WG_I / 8);
            const bool ok = m0 + r < m_end && i0 + c * 8 < I;        // I % 8 == 0
            cp_async_16(smem_u32(&sA[stage][r * LDA + c * 8]), A + (ok ? (m0 + r) * lda + i0 + c * 8 : 0), ok);
        }
        for (int i = tid; i < WG_M * (JP / 8); i += WG_THREADS) {
            const int r = i / (JP / 8), c = i - r * (JP / 8);
            const bool ok = m0 + r < m_end && c * 8 < J;             // J % 8 == 0
            cp_async_16(smem_u32(&sB[stage][r * LDB + c * 8]), B + (ok ? (m0 + r) * ldb + c * 8 : 0), ok);
        }
    };
    const long long nchunk = (m_end - m_begin + WG_M - 1) / WG_M;
    if (nchunk > 0) load(m_begin, 0);
    cp_async_commit();
    for (long long ch = 0; ch < nchunk; ++ch) {
        const int stage = static_cast<int>(ch & 1);
        if (ch + 1 < nchunk) load(m_begin + (ch + 1) * WG_M, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (ones_col >= 0) {       // bias gradient: a column of ones for the rows that exist
            const long long m0 = m_begin + ch * WG_M;
            if (tid < WG_M) sB[stage][tid * LDB + ones_col] = __float2bfloat16(m0 + tid < m_end ? 1.0f : 0.0f);
            __syncthreads();
        }
        const __nv_bfloat16* a_s = sA[stage];
        const __nv_bfloat16* b_s = sB[stage];
#pragma unroll
        for (int kk = 0; kk < WG_M / 16; ++kk) {
            uint32_t af[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                // A operand (i x m) from the [m][i] tile: transposed 8x8 blocks
                const int row = kk * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int col = warp * 32 + a * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(smem_u32(a_s + row * LDA + col), af[a][0], af[a][1], af[a][2], af[a][3]);
            }
#pragma unroll
            for (int jp = 0; jp < NT / 2; ++jp) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(b_s + (kk * 16 + (lane & 15)) * LDB + jp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    mma_bf16_16816(acc[a][2 * jp], af[a], b0, b1);
                    mma_bf16_16816(acc[a][2 * jp + 1], af[a], b2, b3);
                }
            }
        }
        __syncthreads();
    }
    float* out = partial + static_cast<long long>(blockIdx.y) * I * JP;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int col = n * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = i0 + warp * 32 + a * 16 + g + h * 8;
                if (row < I) {
                    float2* dst = reinterpret_cast<float2*>(out + static_cast<long long>(row) * JP + col);
                    float2 v = make_float2(acc[a][n][2 * h], acc[a][n][2 * h + 1]);
                    if (accumulate) {
                        const float2 old = *dst;
                        v.x += old.x;
                        v.y += old.y;
                    }
                    *dst = v;
                }
            }
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long long n, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<long long>(k) * n + i];
    out[i] = s;
}

__global__ void add_f32_kernel(float4* __restrict__ y, const float4* __restrict__ x, long long nvec) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 a = y[i];
        const float4 b = __ldcs(x + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        y[i] = a;
    }
}

// Weight gradient of embedding_generator (image_encoder.py:278-281).  The reference feeds it a SCRAMBLED view of the
// patch embedding: the flat [T * D] buffer of an image re-read as X[c'][t'] = flat[c' * T + t'], emb[t'] = X[:, t'] W^T.
//   partial[b][c'][r] = sum_t' X_b[c'][t'] * dt_b[t'][r]       (dW_eg^T, summed over images by wgrad_reduce)
// X is contiguous along t', dt along r: plain fp32 FMAs out of shared memory (13 GFLOP per 32 ViT-H images).
constexpr int EG_C = 32, EG_T = 64, EG_RMAX = 40;
__global__ void __launch_bounds__(256)
eg_wgrad_kernel(const __nv_bfloat16* __restrict__ x0b, const __nv_bfloat16* __restrict__ dt, int T, int D, int RP,
                float* __restrict__ partial) {
    __shared__ float sX[EG_C][EG_T + 1];
    __shared__ float sD[EG_T][EG_RMAX];
    const int b = blockIdx.y, c0 = blockIdx.x * EG_C, tid = threadIdx.x;
    const __nv_bfloat16* xb = x0b + static_cast<long long>(b) * T * D;       // flat image buffer
    const __nv_bfloat16* db = dt + static_cast<long long>(b) * T * RP;
    const int nout = EG_C * RP;                                            // <= 1280 outputs, 5 per thread
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t0 = 0; t0 < T; t0 += EG_T) {
        for (int i = tid; i < EG_C * EG_T; i += 256) {
            const int c = i / EG_T, tt = i - c * EG_T;
            sX[c][tt] = (c0 + c < D && t0 + tt < T) ? __bfloat162float(xb[static_cast<long long>(c0 + c) * T + t0 + tt]) : 0.f;
        }
        for (int i = tid; i < EG_T * RP; i += 256) {
            const int tt = i / RP, r = i - tt * RP;
            sD[tt][r] = t0 + tt < T ? __bfloat162float(db[static_cast<long long>(t0 + tt) * RP + r]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int o = tid + k * 256;
            if (o < nout) {
                const int c = o / RP, r = o - c * RP;
                float a = acc[k];
#pragma unroll 8
                for (int tt = 0; tt < EG_T; ++tt) a = fmaf(sX[c][tt], sD[tt][r], a);
                acc[k] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int o = tid + k * 256;
        if (o < nout) {
            const int c = o / RP, r = o - c * RP;
            if (c0 + c < D) partial[(static_cast<long long>(b) * D + c0 + c) * RP + r] = acc[k];
        }
    }
}

}  // namespace

int layernorm_train_launch(const float* x, const float* gamma, const float* beta, void* out_bf16, void* xhat_bf16,
                           float* rstd, long long rows, int cols, float eps, cudaStream_t stream) {
    if (!x || !gamma || !beta || !out_bf16 || !xhat_bf16 || !rstd) return set_error(CVLM_ERR_ARG, "layernorm_train: null pointer");
    if (rows <= 0 || cols <= 0 || (cols & 3) || cols > LNB_VMAX * 128)
        return set_error(CVLM_ERR_ARG, "layernorm_train: cols=%d must be a multiple of 4 and <= %d", cols, LNB_VMAX * 128);
    const int wpb = 8;
    layernorm_train_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        x, gamma, beta, static_cast<__nv_bfloat16*>(out_bf16), static_cast<__nv_bfloat16*>(xhat_bf16), rstd, rows, cols, eps);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// xhat_kind 0: xsrc = raw bf16 x, stat = (rstd, -mean*rstd) rows; 1: xsrc = bf16 xhat, stat = rstd rows; 2: xsrc = fp32 x
// (statistics recomputed with eps).  g = d(loss)/d(LayerNorm output) [* gamma when gamma != NULL];
// dx = (residual ? residual : 0) + LayerNorm backward; dx may alias residual and g; dxb (optional) = bf16(dx).
int ln_bwd_launch(int xhat_kind, const float* g, const float* gamma, const void* xsrc, const float* stat,
                  const float* residual, float* dx, void* dxb, long long rows, int cols, float eps, cudaStream_t stream) {
    if (!g || !xsrc || !dx || (xhat_kind != 2 && !stat)) return set_error(CVLM_ERR_ARG, "ln_bwd: null pointer");
    if (rows <= 0 || cols <= 0 || (cols & 3) || cols > LNB_VMAX * 128)
        return set_error(CVLM_ERR_ARG, "ln_bwd: cols=%d must be a multiple of 4 and <= %d", cols, LNB_VMAX * 128);
    const int wpb = 8;
    const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
    __nv_bfloat16* db = static_cast<__nv_bfloat16*>(dxb);
#define CVLM_LNB(XH, NV) ln_bwd_kernel<XH, NV><<<grid, wpb * 32, 0, stream>>>(g, gamma, xsrc, stat, residual, dx, db, rows, cols, eps)
#define CVLM_LNB_W(XH)                                  \
    switch (cols) {                                     \
        case 256: CVLM_LNB(XH, 2); break;               \
        case 768: CVLM_LNB(XH, 6); break;               \
        case 1024: CVLM_LNB(XH, 8); break;              \
        case 1280: CVLM_LNB(XH, 10); break;             \
        default: CVLM_LNB(XH, 0); break;                \
    }
    if (xhat_kind == 0) { CVLM_LNB_W(0) }
    else if (xhat_kind == 1) { CVLM_LNB_W(1) }
    else if (xhat_kind == 2) { CVLM_LNB_W(2) }
    else return set_error(CVLM_ERR_ARG, "ln_bwd: xhat_kind %d", xhat_kind);
#undef CVLM_LNB_W
#undef CVLM_LNB
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

int wgrad_splits(long long m) {
    long long s = (m + 2047) / 2048;          // >= 2048 rows per split
    const int cap = 2 * device_sm_count();
    return static_cast<int>(s < 1 ? 1 : s > cap ? cap : s);
}
int wgrad_jp(int j, int ones) { return ((j + (ones ? 1 : 0)) + 15) / 16 * 16; }

// partial: fp32 [splits, I, JP] with JP = wgrad_jp(J, ones), splits = wgrad_splits(M); accumulate != 0 adds to what is
// already there (every element has exactly one writer per launch -> deterministic).  A: bf16 [M, I] (row stride lda),
// B: bf16 [M, J] (row stride ldb); I, J multiples of 8, J + ones <= 64.
int wgrad_tn_launch(const void* A, long long lda, int I, const void* B, long long ldb, int J, int ones, long long M,
                    float* partial, int accumulate, cudaStream_t stream) {
    if (!A || !B || !partial) return set_error(CVLM_ERR_ARG, "wgrad: null pointer");
    if (I <= 0 || J <= 0 || (I & 7) || (J & 7) || (lda & 7) || (ldb & 7) || J + (ones ? 1 : 0) > 64 || M <= 0)
        return set_error(CVLM_ERR_ARG, "wgrad: I=%d J=%d must be multiples of 8, J (+1) <= 64", I, J);
    const int jp = wgrad_jp(J, ones), splits = wgrad_splits(M);
    const long long rps = ((M + splits - 1) / splits + WG_M - 1) / WG_M * WG_M;
    dim3 grid((I + WG_I - 1) / WG_I, splits);
    const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(A);
    const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(B);
    const int oc = ones ? J : -1;
    switch (jp) {
        case 16: wgrad_tn_kernel<16><<<grid, WG_THREADS, 0, stream>>>(a, lda, I, b, ldb, J, oc, M, rps, partial, accumulate); break;
        case 32: wgrad_tn_kernel<32><<<grid, WG_THREADS, 0, stream>>>(a, lda, I, b, ldb, J, oc, M, rps, partial, accumulate); break;
        case 48: wgrad_tn_kernel<48><<<grid, WG_THREADS, 0, stream>>>(a, lda, I, b, ldb, J, oc, M, rps, partial, accumulate); break;
        case 64: wgrad_tn_kernel<64><<<grid, WG_THREADS, 0, stream>>>(a, lda, I, b, ldb, J, oc, M, rps, partial, accumulate); break;
        default: return set_error(CVLM_ERR_ARG, "wgrad: J=%d", J);
    }
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// out[i * JP + j] = sum over splits
int wgrad_reduce_launch(const float* partial, int splits, long long n, float* out, cudaStream_t stream) {
    if (!partial || !out || n <= 0 || splits <= 0) return set_error(CVLM_ERR_ARG, "wgrad_reduce: bad argument");
    wgrad_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(partial, splits, n, out);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

int add_f32_launch(float* y, const float* x, long long n, cudaStream_t stream) {
    if (!x || !y || n <= 0 || (n & 3)) return set_error(CVLM_ERR_ARG, "add_f32: bad argument");
    const long long nvec = n >> 2;
    const long long want = (nvec + 255) / 256;
    add_f32_kernel<<<static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(
        reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), nvec);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// partial: fp32 [batch, D, RP]; x0b: bf16 [batch * T, D] (the patch embedding without pos_embed), dt: bf16 [batch * T, RP]
int eg_wgrad_launch(const void* x0b, const void* dt, int batch, int tokens, int dim, int rank_pad, float* partial,
                    cudaStream_t stream) {
    if (!x0b || !dt || !partial || batch <= 0 || tokens <= 0 || dim <= 0 || rank_pad <= 0 || rank_pad > EG_RMAX)
        return set_error(CVLM_ERR_ARG, "eg_wgrad: bad argument");
    dim3 grid((dim + EG_C - 1) / EG_C, batch);
    eg_wgrad_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x0b), static_cast<const __nv_bfloat16*>(dt),
                                              tokens, dim, rank_pad, partial);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace cvlm
