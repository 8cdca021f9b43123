// HBM-bound side kernels of the encoder: LayerNorm (tokens), LayerNorm2d + NHWC->NCHW (neck output),
// im2col for the two 16x16 stride-16 convolutions and for the neck's 3x3 conv, fp32->bf16 cast, and
// the adapter's embedding_generator (which in the reference reads a SCRAMBLED view of the tokens).
// All of them are pure streaming: 128-bit accesses, one pass over the data, fp32 statistics.
#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// nn.LayerNorm(dim, eps) over the last dim of [rows, cols] fp32 (Block.norm1 / norm2,
// image_encoder.py:431, 444; eps bound by the caller, models/sam_maskdecoder_edge.py:129).
// One warp per row, the row lives in registers, two-pass mean / biased variance like ATen.
constexpr int LN_VMAX = 16;   // float4 per lane -> cols <= 2048
// NV > 0: cols == NV * 128 exactly (ViT-B/L/H widths and the 256-wide neck): fully unrolled, no guards,
// few enough registers for 3 CTAs (24 warps) per SM.  NV == 0: generic width (cols % 4 == 0, <= 2048).
template <bool kOutBf16, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ out, long long rows, int cols, float eps) {
    constexpr int SLOTS = NV > 0 ? NV : LN_VMAX;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = cols >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
    float4 v[SLOTS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
        const int idx = lane + i * 32;
        if (NV > 0 || idx < nvec) {
            v[i] = __ldcs(xr + idx);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
        const int idx = lane + i * 32;
        if (NV > 0 || idx < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / cols + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
        const int idx = lane + i * 32;
        if (NV > 0 || idx < nvec) {
            const float4 gg = __ldg(g4 + idx), bb = __ldg(b4 + idx);
            float4 y;
            y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
            y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
            y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
            y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
            if (kOutBf16) {
                uint2 pk = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
                reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + row * cols)[idx] = pk;
            } else {
                reinterpret_cast<float4*>(static_cast<float*>(out) + row * cols)[idx] = y;
            }
        }
    }
}

template <bool kOutBf16>
static void layernorm_dispatch(const float* x, const float* gamma, const float* beta, void* out, long long rows, int cols,
                               float eps, cudaStream_t stream) {
    const int wpb = 8;
    const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
#define CVLM_LN(NV) layernorm_kernel<kOutBf16, NV><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, cols, eps)
    switch (cols) {
        case 256: CVLM_LN(2); break;
        case 768: CVLM_LN(6); break;
        case 1024: CVLM_LN(8); break;
        case 1280: CVLM_LN(10); break;
        default: CVLM_LN(0); break;
    }
#undef CVLM_LN
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, int out_bf16, long long rows,
                     int cols, float eps, cudaStream_t stream) {
    if (!x || !gamma || !beta || !out) return set_error(CVLM_ERR_ARG, "layernorm: null pointer");
    if (rows <= 0 || cols <= 0 || (cols & 3) || cols > LN_VMAX * 128)
        return set_error(CVLM_ERR_ARG, "layernorm: cols=%d must be a multiple of 4 and <= %d", cols, LN_VMAX * 128);
    if (out_bf16)
        layernorm_dispatch<true>(x, gamma, beta, out, rows, cols, eps, stream);
    else
        layernorm_dispatch<false>(x, gamma, beta, out, rows, cols, eps, stream);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ out,
                                                        long long nvec) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = __ldcs(x + i);
        out[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}
// (sum, sum of squares) partials per 128-column group -> (rstd, -mean * rstd) per row: the statistics half of
// nn.LayerNorm (biased variance, eps inside the square root) for the GEMM epilogue that applies them.
__global__ void ln_finalize_kernel(const float2* __restrict__ partial, int groups, float inv_k, float eps,
                                   float2* __restrict__ out, long long rows) {
    const long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float s1 = 0.f, s2 = 0.f;
    for (int g = 0; g < groups; ++g) {
        const float2 p = __ldg(partial + row * groups + g);
        s1 += p.x;
        s2 += p.y;
    }
    const float mu = s1 * inv_k;
    const float var = fmaxf(fmaf(s2, inv_k, -mu * mu), 0.f);
    const float rstd = rsqrtf(var + eps);
    out[row] = make_float2(rstd, -mu * rstd);
}

int ln_finalize_launch(const float* partial, int groups, int width, float eps, float* rows_out, long long rows,
                       cudaStream_t stream) {
    if (!partial || !rows_out || groups <= 0 || width <= 0 || rows <= 0) return set_error(CVLM_ERR_ARG, "ln_finalize: bad argument");
    const int threads = 256;
    const long long blocks = (rows + threads - 1) / threads;
    ln_finalize_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        reinterpret_cast<const float2*>(partial), groups, 1.0f / static_cast<float>(width), eps,
        reinterpret_cast<float2*>(rows_out), rows);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

int cast_bf16_launch(const float* x, void* out, long long n, cudaStream_t stream) {
    if (!x || !out || n <= 0 || (n & 3)) return set_error(CVLM_ERR_ARG, "cast: bad arguments");
    const long long nvec = n >> 2;
    const unsigned grid = static_cast<unsigned>(nvec / 256 + 1 < 148 * 16 ? nvec / 256 + 1 : 148 * 16);
    cast_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), static_cast<uint2*>(out), nvec);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// im2col of a 16x16 / stride-16 convolution (PatchEmbed :651-659, PatchEmbed2 :369, 379):
// img fp32 NCHW [B, C, H, H]  ->  bf16 [B*g*g, C*256], column = c*256 + ky*16 + kx, which is the
// flattening order of the conv weight [out, C, 16, 16].  One thread moves one 16-pixel patch row.
__global__ void __launch_bounds__(256) im2col_patch16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                             int chans, int H, int g, long long total) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int tx = static_cast<int>(r % g); r /= g;
        const int ky = static_cast<int>(r % 16); r /= 16;
        const int ty = static_cast<int>(r % g); r /= g;
        const int c = static_cast<int>(r % chans);
        const long long b = r / chans;
        const float4* src = reinterpret_cast<const float4*>(img + ((b * chans + c) * H + (ty * 16 + ky)) * H + tx * 16);
        const float4 a0 = __ldcs(src), a1 = __ldcs(src + 1), a2 = __ldcs(src + 2), a3 = __ldcs(src + 3);
        uint4* dst = reinterpret_cast<uint4*>(out + ((b * g + ty) * g + tx) * static_cast<long long>(chans * 256) + c * 256 + ky * 16);
        dst[0] = make_uint4(pack_bf16(a0.x, a0.y), pack_bf16(a0.z, a0.w), pack_bf16(a1.x, a1.y), pack_bf16(a1.z, a1.w));
        dst[1] = make_uint4(pack_bf16(a2.x, a2.y), pack_bf16(a2.z, a2.w), pack_bf16(a3.x, a3.y), pack_bf16(a3.z, a3.w));
    }
}
int im2col_patch_launch(const float* img, void* out, int batch, int chans, int img_size, int patch, cudaStream_t stream) {
    if (!img || !out) return set_error(CVLM_ERR_ARG, "im2col: null pointer");
    if (patch != 16 || img_size % 16) return set_error(CVLM_ERR_UNSUPPORTED, "im2col: only 16x16 stride-16 patches");
    const int g = img_size / 16;
    const long long total = static_cast<long long>(batch) * chans * g * 16 * g;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    im2col_patch16_kernel<<<grid, 256, 0, stream>>>(img, static_cast<__nv_bfloat16*>(out), chans, img_size, g, total);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// im2col of the neck's 3x3 / pad-1 conv (image_encoder.py:106-112) on NHWC bf16 [B, g, g, C]:
// out bf16 [B*g*g, 9*C], column = (ky*3+kx)*C + c; taps outside the grid are zero.
__global__ void __launch_bounds__(256) im2col_3x3_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int g,
                                                         int cvec, long long total) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int cv = static_cast<int>(r % cvec); r /= cvec;
        const int tap = static_cast<int>(r % 9); r /= 9;
        const int tx = static_cast<int>(r % g); r /= g;
        const int ty = static_cast<int>(r % g);
        const long long b = r / g;
        const int sy = ty + tap / 3 - 1, sx = tx + tap % 3 - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (sy >= 0 && sy < g && sx >= 0 && sx < g) v = __ldg(x + ((b * g + sy) * g + sx) * cvec + cv);
        out[i] = v;
    }
}
int im2col_3x3_launch(const void* x, void* out, int batch, int g, int chans, cudaStream_t stream) {
    if (!x || !out || (chans & 7)) return set_error(CVLM_ERR_ARG, "im2col3x3: channels must be a multiple of 8");
    const int cvec = chans / 8;
    const long long total = static_cast<long long>(batch) * g * g * 9 * cvec;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    im2col_3x3_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), g, cvec, total);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm2d (common.py:31-43) of the neck's last conv, fused with the NHWC -> NCHW permute the
// reference output layout needs: x fp32 [B*T, C] token-major -> out fp32 [B, C, T].
// A block normalises 32 tokens (one warp per token, 4 tokens per warp) into a padded smem tile,
// then writes it transposed so that stores run along the token axis.
__global__ void __launch_bounds__(256) layernorm_nchw_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ out,
                                                             int T, int C, float eps) {
    extern __shared__ float tile[];   // [C][33]
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int tt = warp; tt < 32; tt += 8) {
        const int tkn = t0 + tt;
        if (tkn >= T) continue;
        const float* xr = x + (static_cast<long long>(b) * T + tkn) * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += xr[c];
        const float mean = warp_sum(s) / C;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = xr[c] - mean;
            q += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
        for (int c = lane; c < C; c += 32) tile[c * 33 + tt] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    }
    __syncthreads();
    const int tkn = t0 + lane;
    if (tkn < T)
        for (int c = warp; c < C; c += 8) out[(static_cast<long long>(b) * C + c) * T + tkn] = tile[c * 33 + lane];
}
int layernorm_nchw_out_launch(const float* x, const float* gamma, const float* beta, float* out, int batch, int tokens,
                              int chans, float eps, cudaStream_t stream) {
    if (!x || !gamma || !beta || !out) return set_error(CVLM_ERR_ARG, "layernorm2d: null pointer");
    const int smem = chans * 33 * 4;
    if (smem > 200 * 1024) return set_error(CVLM_ERR_UNSUPPORTED, "layernorm2d: too many channels");
    CVLM_CUDA(cudaFuncSetAttribute(layernorm_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((tokens + 31) / 32, batch);
    layernorm_nchw_kernel<<<grid, 256, smem, stream>>>(x, gamma, beta, out, tokens, chans, eps);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// Adapter input  t = handcrafted + embedding  (get_prompt, image_encoder.py:293).
// PromptGenerator.init_embeddings (:278-281) reshapes the NHWC patch embedding of each image to
// (C, H*W) WITHOUT permuting first, so embedding "token t', feature c'" is flat[c'*T + t'] of the
// image's [T, C] buffer.  We reproduce that read pattern directly on the fp32 residual stream x0
// (= patch embedding + pos_embed); the pos_embed part and the bias are input-independent and arrive
// pre-folded in `cst` ([T, RP] = b_eg - W_eg * scramble(pos_embed)).
//   t[b, t', j] = bf16( hand[b*T+t', j] + cst[t', j] + sum_c' w_t[c', j] * x0[b][c'*T + t'] )
// Reads along t' are contiguous, w_t rows are smem broadcasts, fp32 FMA like the reference's addmm.
template <int RP>
__global__ void __launch_bounds__(128) adapter_embed_kernel(const float* __restrict__ x0, const float* __restrict__ hand,
                                                            const float* __restrict__ w_t, const float* __restrict__ cst,
                                                            __nv_bfloat16* __restrict__ t_out, int T, int D, int rank) {
    constexpr int CHUNK = 64;
    __shared__ __align__(16) float sw[CHUNK * RP];
    static_assert(RP % 4 == 0, "adapter_embed: float4 weight reads");
    const int b = blockIdx.y;
    const int tp = blockIdx.x * 128 + threadIdx.x;
    const bool ok = tp < T;
    const float* xb = x0 + static_cast<long long>(b) * T * D;
    float acc[RP];
#pragma unroll
    for (int j = 0; j < RP; ++j) acc[j] = 0.f;
    for (int c0 = 0; c0 < D; c0 += CHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < CHUNK * RP; i += 128) {
            const int cc = c0 + i / RP;
            sw[i] = cc < D ? w_t[static_cast<long long>(cc) * RP + i % RP] : 0.f;
        }
        __syncthreads();
        if (ok) {
            const int cend = D - c0 < CHUNK ? D - c0 : CHUNK;
            int c = 0;
            for (; c + 8 <= cend; c += 8) {          // eight independent loads in flight, then 8 x RP FMAs
                float xv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) xv[k] = __ldg(xb + static_cast<long long>(c0 + c + k) * T + tp);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4* w4 = reinterpret_cast<const float4*>(sw + (c + k) * RP);
#pragma unroll
                    for (int j = 0; j < RP; j += 4) {
                        const float4 w = w4[j / 4];
                        acc[j] = fmaf(w.x, xv[k], acc[j]);
                        acc[j + 1] = fmaf(w.y, xv[k], acc[j + 1]);
                        acc[j + 2] = fmaf(w.z, xv[k], acc[j + 2]);
                        acc[j + 3] = fmaf(w.w, xv[k], acc[j + 3]);
                    }
                }
            }
            for (; c < cend; ++c) {
                const float xv = __ldg(xb + static_cast<long long>(c0 + c) * T + tp);
#pragma unroll
                for (int j = 0; j < RP; ++j) acc[j] = fmaf(sw[c * RP + j], xv, acc[j]);
            }
        }
    }
    if (!ok) return;
    const long long row = static_cast<long long>(b) * T + tp;
#pragma unroll
    for (int j = 0; j < RP; j += 2) {
        float v0 = 0.f, v1 = 0.f;
        if (j < rank) v0 = acc[j] + hand[row * RP + j] + cst[static_cast<long long>(tp) * RP + j];
        if (j + 1 < rank) v1 = acc[j + 1] + hand[row * RP + j + 1] + cst[static_cast<long long>(tp) * RP + j + 1];
        *reinterpret_cast<uint32_t*>(t_out + row * RP + j) = pack_bf16(v0, v1);
    }
}
int adapter_embed_launch(const float* x0, const float* hand, const float* w_t, const float* cst, void* t_out, int batch,
                         int tokens, int dim, int rank, int rank_pad, cudaStream_t stream) {
    if (!x0 || !hand || !w_t || !cst || !t_out) return set_error(CVLM_ERR_ARG, "adapter_embed: null pointer");
    if (rank <= 0 || rank > rank_pad) return set_error(CVLM_ERR_ARG, "adapter_embed: bad rank");
    dim3 grid((tokens + 127) / 128, batch);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(t_out);
    switch (rank_pad) {
        case 8: adapter_embed_kernel<8><<<grid, 128, 0, stream>>>(x0, hand, w_t, cst, o, tokens, dim, rank); break;
        case 16: adapter_embed_kernel<16><<<grid, 128, 0, stream>>>(x0, hand, w_t, cst, o, tokens, dim, rank); break;
        case 24: adapter_embed_kernel<24><<<grid, 128, 0, stream>>>(x0, hand, w_t, cst, o, tokens, dim, rank); break;
        case 32: adapter_embed_kernel<32><<<grid, 128, 0, stream>>>(x0, hand, w_t, cst, o, tokens, dim, rank); break;
        case 40: adapter_embed_kernel<40><<<grid, 128, 0, stream>>>(x0, hand, w_t, cst, o, tokens, dim, rank); break;
        default: return set_error(CVLM_ERR_UNSUPPORTED, "adapter_embed: rank_pad %d not in {8,16,24,32,40}", rank_pad);
    }
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace cvlm
