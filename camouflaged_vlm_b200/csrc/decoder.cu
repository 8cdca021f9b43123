// Mask decoder runtime: MaskDecoder_Edge.forward / predict_masks
// (models/mmseg/models/sam/mask_decoder_edge.py:96-190) with TwoWayTransformer_MaskDecoder_Edge
// (transformer_maskdecoder_edge.py:62-214), batched over images -- the reference only works at batch 1
// (repeat_interleave of the embeddings by the prompt batch), here image b is decoded with prompt b.
//
// Split of the work
//   image side  (4096 keys x 256 per image, and the 128^2 / 256^2 up-scaling maps): every projection, both transposed
//               convolutions (k2 s2: four GEMM column groups per input pixel) and both 3x3 convolutions (implicit GEMM,
//               no im2col) run on the tcgen05 GEMM (gemm_tcgen05.cu) with bf16 operands and fp32 accumulation; `keys`
//               stays fp32.
//   token side  (6 output tokens + the prompt tokens per image): one CTA cluster per image runs a whole transformer layer
//               in fp32 (self attention, token->image attention over the 4096 keys, token->prompt attention, MLP, and the
//               key / value projections the image side needs next).
//
// Data layout in HBM for a chunk of B images (T = g*g, M = B*T; P1 = 4M, P2 = 16M pixels of the up-scaled maps):
//   keys  fp32 [M,256]   image tokens (NHWC), updated in place by the out_proj epilogues + LayerNorm
//   kb    bf16 [M,256]   keys;   kpb bf16 [M,256]  keys + image_pe;   eb bf16 [M,256]  the raw embedding
//   kimg / vimg / qi  fp32 [M,128]   projections for the cross attentions;  ao bf16 [M,128] attention output
//   Up-scaled maps keep NESTED pixel order: pixel (y,x) of the g x g map owns the contiguous rows
//   ((y*g+x)*4 + s1)*4 + s2, s = 2*dy+dx of each transposed convolution, so a k2 s2 transposed convolution is a plain
//   reshape of the GEMM output.  The two 3x3 convolutions run as implicit GEMMs over zero-bordered RASTER maps
//   (GemmArgs::conv_pitch); dec_pad_raster_kernel and the final mask kernel translate between the two orders.
#include <string.h>

#include <string>
#include <unordered_map>

#include <cooperative_groups.h>

#include "cvlm_internal.cuh"

namespace cg = cooperative_groups;

struct cvlm_decoder {
    cvlm_decoder_config cfg;
    std::unordered_map<std::string, std::pair<const void*, long long>> tensors;
    long long last_launches = 0;
};

namespace cvlm {

namespace {

constexpr int DD = 256;        // transformer_dim
constexpr int DI = 128;        // inner width of the cross attentions (attention_downsample_rate 2)
constexpr int NH = 8;          // heads
constexpr int NTOK = 6;        // iou token, 4 mask tokens, edge token (mask_decoder_edge.py:157)
constexpr int NMASK = 4;
constexpr int MAXC = 8;        // prompt tokens per image supported by the kernels
constexpr int DQ = 64, DE = 32;
constexpr int TOK_THREADS = 512;
constexpr int RED_FLOATS = 16 * NTOK * 192;     // scratch of the token kernels (floats)

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// prep: NCHW -> token-major.  keys = emb + dense, kb = bf16(keys), kpb = bf16(keys + pe), eb = bf16(emb), pos_t = pe.
// One CTA per 32 tokens of one image; the 256 x 32 tile goes through shared memory three times (emb, dense, pe).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_prep_kernel(const float* __restrict__ emb, const float* __restrict__ dense,
                                                       long long dense_bstride, const float* __restrict__ pe,
                                                       float* __restrict__ keys, __nv_bfloat16* __restrict__ kb,
                                                       __nv_bfloat16* __restrict__ kpb, __nv_bfloat16* __restrict__ eb,
                                                       float* __restrict__ pos_t, int T) {
    __shared__ float tile[DD][33];
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float val[4][8];
    auto load = [&](const float* src) {
        __syncthreads();
        for (int c = warp; c < DD; c += 8) tile[c][lane] = src[static_cast<long long>(c) * T + t0 + lane];
        __syncthreads();
    };
    load(emb + static_cast<long long>(b) * DD * T);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = static_cast<long long>(b) * T + t0 + warp * 4 + i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            val[i][j] = tile[lane + 32 * j][warp * 4 + i];
            eb[row * DD + lane + 32 * j] = __float2bfloat16_rn(val[i][j]);
        }
    }
    load(dense + static_cast<long long>(b) * dense_bstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = static_cast<long long>(b) * T + t0 + warp * 4 + i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            val[i][j] = __fadd_rn(val[i][j], tile[lane + 32 * j][warp * 4 + i]);
            keys[row * DD + lane + 32 * j] = val[i][j];
            kb[row * DD + lane + 32 * j] = __float2bfloat16_rn(val[i][j]);
        }
    }
    load(pe);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = static_cast<long long>(b) * T + t0 + warp * 4 + i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p = tile[lane + 32 * j][warp * 4 + i];
            kpb[row * DD + lane + 32 * j] = __float2bfloat16_rn(__fadd_rn(val[i][j], p));
            if (b == 0) pos_t[static_cast<long long>(t0 + warp * 4 + i) * DD + lane + 32 * j] = p;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row LayerNorm over C channels (nn.LayerNorm of the transformer, LayerNorm2d of the up-scaling stacks: both are
// (x - mean) / sqrt(var + eps) * w + b over the channel vector of one pixel), optional GELU, one warp per row.
// Outputs (each optional): fp32, bf16, bf16 of (result + pos[row % T]).  In place on x is allowed.
// ------------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) dec_ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, int gelu,
                                                          const float* __restrict__ pos, int T, float* out_f32,
                                                          __nv_bfloat16* __restrict__ out_b16,
                                                          __nv_bfloat16* __restrict__ out_pos_b16, long long rows,
                                                          int pad_pitch) {
    // a row is covered by C/4 lanes holding 4 consecutive channels each: one warp normalises 128/C rows at a time, and
    // RPT such groups of rows per thread so that enough loads are in flight for the narrow maps
    constexpr int LPR = C / 4;              // lanes per row: 64 (C = 256, two passes of the warp) or 16 (C = 64)
    constexpr int RPW = LPR >= 32 ? 1 : 32 / LPR;
    constexpr int PASS = LPR >= 32 ? LPR / 32 : 1;
    constexpr int RPT = C >= 128 ? 1 : 2;
    const int lane = threadIdx.x & 31;
    const int c0 = (RPW > 1 ? lane % LPR : lane) * 4;
    long long row[RPT];
    float4 v[RPT][PASS];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = ((static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * RPT + r) * RPW + (RPW > 1 ? lane / LPR : 0);
#pragma unroll
        for (int j = 0; j < PASS; ++j)
            v[r][j] = row[r] < rows ? *reinterpret_cast<const float4*>(x + row[r] * C + c0 + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PASS; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
#pragma unroll
        for (int o = (LPR >= 32 ? 16 : LPR / 2); o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PASS; ++j) {
            v[r][j].x -= mean; v[r][j].y -= mean; v[r][j].z -= mean; v[r][j].w -= mean;
            q += fmaf(v[r][j].x, v[r][j].x, v[r][j].y * v[r][j].y) + fmaf(v[r][j].z, v[r][j].z, v[r][j].w * v[r][j].w);
        }
#pragma unroll
        for (int o = (LPR >= 32 ? 16 : LPR / 2); o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.f / sqrtf(q * (1.f / C) + eps);
        if (row[r] >= rows) continue;
        // padded raster maps (pad_pitch > 0): the border rows feed the next implicit convolution and must stay zero
        bool border = false;
        if (pad_pitch > 0) {
            const unsigned pp = static_cast<unsigned>(pad_pitch) * pad_pitch;
            const unsigned pr = static_cast<unsigned>(row[r]) % pp;      // rows < 2^31 (checked by the host)
            const unsigned py = pr / pad_pitch, px = pr % pad_pitch;
            border = py == 0 || py == static_cast<unsigned>(pad_pitch) - 1 || px == 0 || px == static_cast<unsigned>(pad_pitch) - 1;
        }
#pragma unroll
        for (int j = 0; j < PASS; ++j) {
            const int c = c0 + 128 * j;
            const long long off = row[r] * C + c;
            if (border) {
                if (out_b16) *reinterpret_cast<uint2*>(out_b16 + off) = make_uint2(0u, 0u);
                continue;
            }
            const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
            float4 y = make_float4(fmaf(v[r][j].x * rstd, gm.x, bt.x), fmaf(v[r][j].y * rstd, gm.y, bt.y),
                                   fmaf(v[r][j].z * rstd, gm.z, bt.z), fmaf(v[r][j].w * rstd, gm.w, bt.w));
            if (gelu) y = make_float4(gelu_exact(y.x), gelu_exact(y.y), gelu_exact(y.z), gelu_exact(y.w));
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = y;
            if (out_b16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
                *reinterpret_cast<uint2*>(out_b16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
            if (out_pos_b16) {
                const float4 p = *reinterpret_cast<const float4*>(pos + (row[r] % T) * C + c);
                __nv_bfloat162 lo = __floats2bfloat162_rn(y.x + p.x, y.y + p.y), hi = __floats2bfloat162_rn(y.z + p.z, y.w + p.w);
                *reinterpret_cast<uint2*>(out_pos_b16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Image -> token / image -> prompt attention (Attention.forward with 4096 queries and nk <= 8 keys): one thread per
// (image token, head), 16 channels per head.  q fp32 [M,128]; kt / vt fp32 [B, nk, 128]; out bf16 [M,128].
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_few_key_attn_kernel(const float* __restrict__ q, const float* __restrict__ kt,
                                                               const float* __restrict__ vt, int nk,
                                                               __nv_bfloat16* __restrict__ out, int T) {
    // keys / values of this image in shared memory as [key][quarter of the head][head] float4: the 8 heads of a row sit
    // in consecutive lanes, so every 128-bit shared load of a quarter-warp is conflict free
    __shared__ float4 ks[MAXC][4][NH], vs[MAXC][4][NH];
    const long long row0 = static_cast<long long>(blockIdx.x) * 32;
    const int b = static_cast<int>(row0 / T);
    for (int i = threadIdx.x; i < nk * (DI / 4); i += 256) {
        const int j = i / (DI / 4), c4 = i % (DI / 4);          // float4 c4 of key j: head c4 / 4, quarter c4 % 4
        ks[j][c4 & 3][c4 >> 2] = reinterpret_cast<const float4*>(kt + static_cast<long long>(b) * nk * DI)[i];
        vs[j][c4 & 3][c4 >> 2] = reinterpret_cast<const float4*>(vt + static_cast<long long>(b) * nk * DI)[i];
    }
    __syncthreads();
    const int h = threadIdx.x & 7;
    const long long row = row0 + (threadIdx.x >> 3);
    float4 qv[4];
    const float4* qp = reinterpret_cast<const float4*>(q + row * DI + h * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = qp[i];
    float lg[MAXC], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
        lg[j] = -INFINITY;
        if (j < nk) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 kk = ks[j][i][h];
                s = fmaf(qv[i].x, kk.x, fmaf(qv[i].y, kk.y, fmaf(qv[i].z, kk.z, fmaf(qv[i].w, kk.w, s))));
            }
            lg[j] = s * 0.25f;     // 1 / sqrt(16)
            mx = fmaxf(mx, lg[j]);
        }
    }
    float l = 0.f;
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
        if (j < nk) {
            const float p = expf(lg[j] - mx);
            l += p;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 vv = vs[j][i][h];
                acc[i].x = fmaf(p, vv.x, acc[i].x); acc[i].y = fmaf(p, vv.y, acc[i].y);
                acc[i].z = fmaf(p, vv.z, acc[i].z); acc[i].w = fmaf(p, vv.w, acc[i].w);
            }
        }
    }
    const float inv = 1.f / l;
    __nv_bfloat162 o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o[2 * i] = __floats2bfloat162_rn(acc[i].x * inv, acc[i].y * inv);
        o[2 * i + 1] = __floats2bfloat162_rn(acc[i].z * inv, acc[i].w * inv);
    }
    uint4* op = reinterpret_cast<uint4*>(out + row * DI + h * 16);
    op[0] = *reinterpret_cast<uint4*>(&o[0]);
    op[1] = *reinterpret_cast<uint4*>(&o[4]);
}

// ------------------------------------------------------------------------------------------------------------------
// Token side.  Helpers run with the whole CTA (TOK_THREADS threads); activations live in shared memory.
// ------------------------------------------------------------------------------------------------------------------
struct AttnW {           // one Attention module (transformer_maskdecoder_edge.py:217-272), fp32 parameters
    const float *qw, *qb, *kw, *kb, *vw, *vb, *ow, *ob;
};
struct NormW {
    const float *w, *b;
};
struct LayerW {
    AttnW self_attn, t2i, t2c, i2t, i2c;
    NormW n1, n2, n2c, n3;
    const float *l1w, *l1b, *l2w, *l2b;
    int mlp_dim;
};
struct Mlp3W {
    const float *w0, *b0, *w1, *b1, *w2, *b2;
};
struct FinalW {
    AttnW t2i;
    NormW nf;
    Mlp3W hyper[NMASK + 1];   // output_hypernetworks_mlps[0..3], edge_mlp
    Mlp3W iou;
};

// The token kernels run one thread-block CLUSTER per image (1, 2, 4 or 8 CTAs, chosen by the host so that the batch fills
// the machine once).  Every CTA keeps a full copy of the (tiny) activations in its own shared memory; a linear layer or the
// token->image attention is split across the ranks, and each rank writes its slice into every rank's copy through
// distributed shared memory.  Cheap steps (LayerNorm, 6x6 attention, adds) are simply repeated by every rank on identical
// data.  Each slice has exactly one producer, so all copies stay bit-identical and nothing depends on scheduling.
struct Cl {
    int rank, size;
};
__device__ __forceinline__ void cl_sync(const Cl& cl) {
    if (cl.size > 1) cg::this_cluster().sync();
    else __syncthreads();
}
__device__ __forceinline__ float* cl_remote(float* p, const Cl& cl, int r) {
    return cl.size > 1 ? cg::this_cluster().map_shared_rank(p, r) : p;
}
constexpr int SYNC_BEFORE = 1;   // cluster barrier before this step publishes (everyone is done reading the destination)
constexpr int SYNC_AFTER = 2;    // cluster barrier after it (the destination is complete everywhere)
constexpr int SYNC_BOTH = 3;

// ys[t][o] = act(sum_k xs[t][k] * W[o][k] + b[o]) for t < NT, with the weight stored TRANSPOSED (Wt[k][o], packed by the
// host as "<name>.weight.t").  Rank r of the cluster owns a contiguous block of output columns; inside the CTA a thread
// owns four adjacent columns (a warp reads contiguous runs of a weight row) and a slice of the K range, the partial sums
// meet in `red` and are added in a fixed order.  Several independent linears may share one pair of barriers (`sync`).
// out % 4 == 0, in % 32 == 0.
template <int NT>
__device__ void tok_linear(const float* __restrict__ Wt, const float* __restrict__ bias, const float* xs, int ldx, int in,
                           float* ys, int ldy, int out, bool relu, float* red, const Cl& cl, int sync) {
    const int nthr = blockDim.x, tid = threadIdx.x;
    const int per = ((out + cl.size * 4 - 1) / (cl.size * 4)) * 4;           // columns per rank
    const int c_lo = min(out, cl.rank * per), ncol = min(out, c_lo + per) - c_lo;
    const int cthreads = ncol / 4;
    int split = 1;
    if (cthreads > 0 && cthreads < nthr) {
        split = nthr / cthreads;
        if (split > in / 4) split = in / 4;
        while (split > 1 && (in % (split * 4) || split * NT * ncol > RED_FLOATS)) split >>= 1;
    }
    const int kchunk = in / split;
    // ---- partial sums of this rank's columns
    for (int c = tid % max(cthreads, 1), ks = tid / max(cthreads, 1); cthreads > 0 && c < cthreads && ks < split; c += nthr) {
        float acc[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
        const float4* wp = reinterpret_cast<const float4*>(Wt + c_lo) + c;
        const int ldw = out / 4;
#pragma unroll 2
        for (int k = ks * kchunk; k < (ks + 1) * kchunk; k += 4) {
            float4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = __ldg(wp + static_cast<long long>(k + u) * ldw);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + t * ldx + k);
                acc[t][0] = fmaf(xv.x, w[0].x, acc[t][0]); acc[t][1] = fmaf(xv.x, w[0].y, acc[t][1]);
                acc[t][2] = fmaf(xv.x, w[0].z, acc[t][2]); acc[t][3] = fmaf(xv.x, w[0].w, acc[t][3]);
                acc[t][0] = fmaf(xv.y, w[1].x, acc[t][0]); acc[t][1] = fmaf(xv.y, w[1].y, acc[t][1]);
                acc[t][2] = fmaf(xv.y, w[1].z, acc[t][2]); acc[t][3] = fmaf(xv.y, w[1].w, acc[t][3]);
                acc[t][0] = fmaf(xv.z, w[2].x, acc[t][0]); acc[t][1] = fmaf(xv.z, w[2].y, acc[t][1]);
                acc[t][2] = fmaf(xv.z, w[2].z, acc[t][2]); acc[t][3] = fmaf(xv.z, w[2].w, acc[t][3]);
                acc[t][0] = fmaf(xv.w, w[3].x, acc[t][0]); acc[t][1] = fmaf(xv.w, w[3].y, acc[t][1]);
                acc[t][2] = fmaf(xv.w, w[3].z, acc[t][2]); acc[t][3] = fmaf(xv.w, w[3].w, acc[t][3]);
            }
        }
        if (split > 1) {
#pragma unroll
            for (int t = 0; t < NT; ++t)
                *reinterpret_cast<float4*>(red + (ks * NT + t) * ncol + 4 * c) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        } else {
            // wide layer: this thread owns whole columns; finish them here and keep them in `red` as [NT][ncol]
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) bv = *reinterpret_cast<const float4*>(bias + c_lo + 4 * c);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float4 y = make_float4(acc[t][0] + bv.x, acc[t][1] + bv.y, acc[t][2] + bv.z, acc[t][3] + bv.w);
                if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
                *reinterpret_cast<float4*>(red + t * ncol + 4 * c) = y;
            }
        }
    }
    __syncthreads();
    if (split > 1) {
        // fold the K slices: red[0][t][o] <- act(bias + sum_s red[s][t][o])
        for (int i = tid; i < NT * ncol; i += nthr) {
            const int o = i % ncol;
            float y = bias ? bias[c_lo + o] : 0.f;
            for (int s2 = 0; s2 < split; ++s2) y += red[s2 * NT * ncol + i];
            red[i] = relu ? fmaxf(y, 0.f) : y;      // slot (0, t, o) == i: only read by this thread in the loop above
        }
        __syncthreads();
    }
    // ---- publish this rank's columns into every rank's copy of ys
    if (sync & SYNC_BEFORE) cl_sync(cl);
    for (int r = 0; r < cl.size; ++r) {
        float* dst = cl_remote(ys, cl, r);
        for (int i = tid; i < NT * ncol; i += nthr) dst[(i / ncol) * ldy + c_lo + i % ncol] = red[i];
    }
    if (sync & SYNC_AFTER) cl_sync(cl);
    else __syncthreads();       // `red` is reused by the next step
}

// dst[t][:] = LayerNorm(a[t][:] + (r ? r[t][:] : 0)) over 256 channels, eps 1e-5 (nn.LayerNorm default); one warp per token
__device__ void tok_add_ln(float* dst, const float* a, const float* r, NormW n, float eps, int nt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nt) {
        float v[8], s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            v[j] = a[warp * DD + c] + (r ? r[warp * DD + c] : 0.f);
            s += v[j];
        }
        const float mean = warp_sum(s) * (1.f / DD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] -= mean;
            q = fmaf(v[j], v[j], q);
        }
        const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / DD) + eps);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            dst[warp * DD + c] = fmaf(v[j] * rstd, n.w[c], n.b[c]);
        }
    }
    __syncthreads();
}

// dst = a + s * b, elementwise over n floats
__device__ void tok_axpy(float* dst, const float* a, const float* b, float s, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = a[i] + s * b[i];
    __syncthreads();
}

// softmax(q k^T / sqrt(hd)) v for a handful of tokens, all in shared memory; one thread per (head, query)
__device__ void tok_small_attention(const float* q, const float* k, const float* v, int ld, int nq, int nk, int hd,
                                    float* out) {
    const float scale = rsqrtf(static_cast<float>(hd));
    for (int idx = threadIdx.x; idx < NH * nq; idx += blockDim.x) {
        const int h = idx / nq, i = idx % nq;
        float lg[MAXC], mx = -INFINITY;
        for (int j = 0; j < nk; ++j) {
            float s = 0.f;
            for (int d = 0; d < hd; ++d) s = fmaf(q[i * ld + h * hd + d], k[j * ld + h * hd + d], s);
            lg[j] = s * scale;
            mx = fmaxf(mx, lg[j]);
        }
        float l = 0.f;
        for (int j = 0; j < nk; ++j) {
            lg[j] = expf(lg[j] - mx);
            l += lg[j];
        }
        const float inv = 1.f / l;
        for (int d = 0; d < hd; ++d) {
            float a = 0.f;
            for (int j = 0; j < nk; ++j) a = fmaf(lg[j], v[j * ld + h * hd + d], a);
            out[i * ld + h * hd + d] = a * inv;
        }
    }
    __syncthreads();
}

// Token -> image attention: 6 queries (smem, [6][128]) against the T keys of this image (fp32 [T,128] in HBM).
// Rank r of the cluster takes the keys [r T / size, (r+1) T / size).  A warp reads one whole key row per load (lane = 4
// channels, 4 lanes per head: 512 contiguous bytes), the 16-channel dot products are finished with two shuffles, and every
// lane carries an online-softmax state (m, l, 4 channels of the output) for the 6 tokens of its head.  The 16 warps'
// states are merged through `red` ([16][6][192] floats) in warp order, the ranks' merged states through `part`
// ([size][6][192] in every CTA) in rank order, so the result does not depend on scheduling.
__device__ void tok_image_attention(const float* qs, const float* __restrict__ kimg, const float* __restrict__ vimg,
                                    int T, float* out, float* red, float* part, const Cl& cl) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr int STRIDE = DI + 32 + 32;
    float4 q[NTOK], acc[NTOK];
    float m[NTOK], l[NTOK];
#pragma unroll
    for (int t = 0; t < NTOK; ++t) {
        q[t] = *reinterpret_cast<const float4*>(qs + t * DI + 4 * lane);
        q[t].x *= 0.25f; q[t].y *= 0.25f; q[t].z *= 0.25f; q[t].w *= 0.25f;      // 1 / sqrt(16)
        acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        m[t] = -INFINITY;
        l[t] = 0.f;
    }
    const float4* kp = reinterpret_cast<const float4*>(kimg) + lane;
    const float4* vp = reinterpret_cast<const float4*>(vimg) + lane;
    const int k_lo = (T / cl.size) * cl.rank, k_hi = cl.rank == cl.size - 1 ? T : k_lo + T / cl.size;
    for (int key = k_lo + warp; key < k_hi; key += 2 * nw) {
        const bool two = key + nw < k_hi;
        const float4 k0 = __ldg(kp + static_cast<long long>(key) * (DI / 4));
        const float4 v0 = __ldg(vp + static_cast<long long>(key) * (DI / 4));
        const float4 k1 = two ? __ldg(kp + static_cast<long long>(key + nw) * (DI / 4)) : k0;
        const float4 v1 = two ? __ldg(vp + static_cast<long long>(key + nw) * (DI / 4)) : v0;
#pragma unroll
        for (int t = 0; t < NTOK; ++t) {
            float s0 = fmaf(q[t].x, k0.x, fmaf(q[t].y, k0.y, fmaf(q[t].z, k0.z, q[t].w * k0.w)));
            float s1 = fmaf(q[t].x, k1.x, fmaf(q[t].y, k1.y, fmaf(q[t].z, k1.z, q[t].w * k1.w)));
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            if (!two) s1 = -INFINITY;
            const float mn = fmaxf(m[t], fmaxf(s0, s1));
            const float c = __expf(m[t] - mn), p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
            l[t] = fmaf(l[t], c, p0 + p1);
            acc[t].x = fmaf(acc[t].x, c, fmaf(p0, v0.x, p1 * v1.x));
            acc[t].y = fmaf(acc[t].y, c, fmaf(p0, v0.y, p1 * v1.y));
            acc[t].z = fmaf(acc[t].z, c, fmaf(p0, v0.z, p1 * v1.z));
            acc[t].w = fmaf(acc[t].w, c, fmaf(p0, v0.w, p1 * v1.w));
            m[t] = mn;
        }
    }
#pragma unroll
    for (int t = 0; t < NTOK; ++t) {
        float* r = red + (warp * NTOK + t) * STRIDE;
        *reinterpret_cast<float4*>(r + 4 * lane) = acc[t];
        r[DI + lane] = m[t];
        r[DI + 32 + lane] = l[t];
    }
    __syncthreads();
    // merge `n` states laid out [n][NTOK][STRIDE] at `src` for (token t, lane ln); returns the un-normalised state
    auto merge = [&](const float* src, int n, int t, int ln, float4& a, float& mm, float& lt) {
        mm = -INFINITY;
        for (int w = 0; w < n; ++w) mm = fmaxf(mm, src[(w * NTOK + t) * STRIDE + DI + ln]);
        lt = 0.f;
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < n; ++w) {
            const float* r = src + (w * NTOK + t) * STRIDE;
            const float mw = r[DI + ln];
            const float c = (mw == -INFINITY) ? 0.f : __expf(mw - mm);
            lt = fmaf(r[DI + 32 + ln], c, lt);
            const float4 av = *reinterpret_cast<const float4*>(r + 4 * ln);
            a.x = fmaf(av.x, c, a.x); a.y = fmaf(av.y, c, a.y); a.z = fmaf(av.z, c, a.z); a.w = fmaf(av.w, c, a.w);
        }
    };
    float4 a[1];
    float mm = 0.f, lt = 0.f;
    const bool owner = threadIdx.x < NTOK * 32;      // one thread per (token, lane)
    const int t = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (owner) merge(red, nw, t, ln, a[0], mm, lt);
    if (cl.size > 1) {
        cl_sync(cl);                                  // every rank is done reading `part` (previous use) and `out`
        if (owner) {
            for (int r = 0; r < cl.size; ++r) {
                float* dst = cl_remote(part, cl, r) + (cl.rank * NTOK + t) * STRIDE;
                *reinterpret_cast<float4*>(dst + 4 * ln) = a[0];
                dst[DI + ln] = mm;
                dst[DI + 32 + ln] = lt;
            }
        }
        cl_sync(cl);
        if (owner) merge(part, cl.size, t, ln, a[0], mm, lt);
    }
    if (owner) {
        const float inv = 1.f / lt;
        *reinterpret_cast<float4*>(out + t * DI + 4 * ln) = make_float4(a[0].x * inv, a[0].y * inv, a[0].z * inv, a[0].w * inv);
    }
    __syncthreads();
}

// Shared-memory plan of the token kernels (floats)
constexpr int SM_Q = 0;                       // queries          [6][256]
constexpr int SM_PE = SM_Q + NTOK * DD;       // query_pe         [6][256]
constexpr int SM_X = SM_PE + NTOK * DD;       // queries (+ pe)   [6][256]
constexpr int SM_T1 = SM_X + NTOK * DD;       // scratch          [8][256] each
constexpr int SM_T2 = SM_T1 + MAXC * DD;
constexpr int SM_T3 = SM_T2 + MAXC * DD;
constexpr int SM_T4 = SM_T3 + MAXC * DD;
constexpr int SM_CND = SM_T4 + MAXC * DD;     // prompt tokens    [8][256]
constexpr int SM_CN2 = SM_CND + MAXC * DD;    // 2 * prompt       [8][256]  (k = cond + cond, transformer_maskdecoder_edge.py:188)
constexpr int SM_RED = SM_CN2 + MAXC * DD;    // split-K partials / attention merge of the 16 warps [16][6][192]
constexpr int SM_HID = SM_RED + RED_FLOATS;   // MLP hidden [6][mlp_dim]; the ranks' attention states [8][6][192] before that
constexpr int HID_MIN_FLOATS = 8 * NTOK * 192;

// One TwoWayAttentionBlock, token side (transformer_maskdecoder_edge.py:161-199), for image blockIdx.x / cluster size; then
// the key / value projections that the image side of the same block consumes (:201-212).
__global__ void __launch_bounds__(TOK_THREADS) dec_token_layer_kernel(
    LayerW w, int first, int nc, int T, float eps, const float* __restrict__ tokens0, const float* __restrict__ cond,
    float* __restrict__ qbuf, const float* __restrict__ kimg, const float* __restrict__ vimg, float* __restrict__ kc,
    float* __restrict__ vc, float* __restrict__ ktok, float* __restrict__ vtok) {
    extern __shared__ __align__(16) float sm[];
    const Cl cl{static_cast<int>(cg::this_cluster().block_rank()), static_cast<int>(cg::this_cluster().num_blocks())};
    const int b = blockIdx.x / cl.size;
    float *Q = sm + SM_Q, *PE = sm + SM_PE, *X = sm + SM_X, *T1 = sm + SM_T1, *T2 = sm + SM_T2, *T3 = sm + SM_T3,
          *T4 = sm + SM_T4, *CND = sm + SM_CND, *CN2 = sm + SM_CN2, *HID = sm + SM_HID, *RED = sm + SM_RED;
    for (int i = threadIdx.x; i < NTOK * DD; i += blockDim.x) {
        PE[i] = tokens0[i];
        Q[i] = first ? tokens0[i] : qbuf[static_cast<long long>(b) * NTOK * DD + i];
    }
    for (int i = threadIdx.x; i < MAXC * DD; i += blockDim.x) {
        const float c = i < nc * DD ? cond[static_cast<long long>(b) * nc * DD + i] : 0.f;
        CND[i] = c;
        CN2[i] = c + c;
    }
    __syncthreads();
    const float* kb_ = kimg + static_cast<long long>(b) * T * DI;
    const float* vb_ = vimg + static_cast<long long>(b) * T * DI;
    // ---- self attention (:163-170); the first layer skips the positional term and the shortcut
    tok_axpy(X, Q, PE, first ? 0.f : 1.f, NTOK * DD);
    tok_linear<NTOK>(w.self_attn.qw, w.self_attn.qb, X, DD, DD, T1, DD, DD, false, RED, cl, SYNC_BEFORE);
    tok_linear<NTOK>(w.self_attn.kw, w.self_attn.kb, X, DD, DD, T2, DD, DD, false, RED, cl, 0);
    tok_linear<NTOK>(w.self_attn.vw, w.self_attn.vb, Q, DD, DD, T3, DD, DD, false, RED, cl, SYNC_AFTER);
    tok_small_attention(T1, T2, T3, DD, NTOK, NTOK, DD / NH, T4);
    tok_linear<NTOK>(w.self_attn.ow, w.self_attn.ob, T4, DD, DD, T1, DD, DD, false, RED, cl, SYNC_BOTH);
    tok_add_ln(Q, T1, first ? nullptr : Q, w.n1, eps, NTOK);
    // ---- token -> image cross attention (:172-177); kimg = (keys + pe) Wk, vimg = keys Wv come from the GEMMs
    tok_axpy(X, Q, PE, 1.f, NTOK * DD);
    tok_linear<NTOK>(w.t2i.qw, w.t2i.qb, X, DD, DD, T1, DI, DI, false, RED, cl, SYNC_BOTH);
    tok_image_attention(T1, kb_, vb_, T, T2, RED, HID, cl);
    tok_linear<NTOK>(w.t2i.ow, w.t2i.ob, T2, DI, DI, T1, DD, DD, false, RED, cl, SYNC_BOTH);
    tok_add_ln(Q, Q, T1, w.n2, eps, NTOK);
    // ---- token -> prompt cross attention (:179-192): q = queries + pe, k = cond + cond, v = cond
    tok_axpy(X, Q, PE, 1.f, NTOK * DD);
    tok_linear<NTOK>(w.t2c.qw, w.t2c.qb, X, DD, DD, T1, DI, DI, false, RED, cl, SYNC_BEFORE);
    tok_linear<MAXC>(w.t2c.kw, w.t2c.kb, CN2, DD, DD, T2, DI, DI, false, RED, cl, 0);
    tok_linear<MAXC>(w.t2c.vw, w.t2c.vb, CND, DD, DD, T3, DI, DI, false, RED, cl, SYNC_AFTER);
    tok_small_attention(T1, T2, T3, DI, NTOK, nc, DI / NH, T4);
    tok_linear<NTOK>(w.t2c.ow, w.t2c.ob, T4, DI, DI, T1, DD, DD, false, RED, cl, SYNC_BOTH);
    tok_add_ln(Q, Q, T1, w.n2c, eps, NTOK);
    // ---- MLP (:194-197), ReLU
    tok_linear<NTOK>(w.l1w, w.l1b, Q, DD, DD, HID, w.mlp_dim, w.mlp_dim, true, RED, cl, SYNC_BOTH);
    tok_linear<NTOK>(w.l2w, w.l2b, HID, w.mlp_dim, w.mlp_dim, T1, DD, DD, false, RED, cl, SYNC_BOTH);
    tok_add_ln(Q, Q, T1, w.n3, eps, NTOK);
    // ---- keys / values of the two image-side attentions of this block (:201-212)
    tok_axpy(X, Q, PE, 1.f, NTOK * DD);
    tok_linear<MAXC>(w.i2c.kw, w.i2c.kb, CN2, DD, DD, T1, DI, DI, false, RED, cl, SYNC_BEFORE);
    tok_linear<MAXC>(w.i2c.vw, w.i2c.vb, CND, DD, DD, T2, DI, DI, false, RED, cl, 0);
    tok_linear<NTOK>(w.i2t.kw, w.i2t.kb, X, DD, DD, T3, DI, DI, false, RED, cl, 0);
    tok_linear<NTOK>(w.i2t.vw, w.i2t.vb, Q, DD, DD, T4, DI, DI, false, RED, cl, SYNC_AFTER);
    if (cl.rank == 0) {
        for (int i = threadIdx.x; i < NTOK * DD; i += blockDim.x) qbuf[static_cast<long long>(b) * NTOK * DD + i] = Q[i];
        for (int i = threadIdx.x; i < nc * DI; i += blockDim.x) {
            kc[static_cast<long long>(b) * nc * DI + i] = T1[i];
            vc[static_cast<long long>(b) * nc * DI + i] = T2[i];
        }
        for (int i = threadIdx.x; i < NTOK * DI; i += blockDim.x) {
            ktok[static_cast<long long>(b) * NTOK * DI + i] = T3[i];
            vtok[static_cast<long long>(b) * NTOK * DI + i] = T4[i];
        }
    }
    cl_sync(cl);      // no CTA may exit while a peer can still write into its shared memory
}

// final_attn_token_to_image + norm_final_attn (transformer_maskdecoder_edge.py:98-109), then the hyper-network MLPs,
// the edge MLP and the IoU head (mask_decoder_edge.py:160-186).  hyper fp32 [B,5,32], iou fp32 [B,4], hs fp32 [B,6,256].
__global__ void __launch_bounds__(TOK_THREADS) dec_token_final_kernel(FinalW w, int T, float eps,
                                                                      const float* __restrict__ tokens0,
                                                                      const float* __restrict__ qbuf,
                                                                      const float* __restrict__ kimg,
                                                                      const float* __restrict__ vimg,
                                                                      float* __restrict__ hs, float* __restrict__ hyper,
                                                                      float* __restrict__ iou) {
    extern __shared__ __align__(16) float sm[];
    const Cl cl{static_cast<int>(cg::this_cluster().block_rank()), static_cast<int>(cg::this_cluster().num_blocks())};
    const int b = blockIdx.x / cl.size;
    float *Q = sm + SM_Q, *PE = sm + SM_PE, *X = sm + SM_X, *T1 = sm + SM_T1, *T2 = sm + SM_T2, *T3 = sm + SM_T3,
          *RED = sm + SM_RED, *HID = sm + SM_HID;
    for (int i = threadIdx.x; i < NTOK * DD; i += blockDim.x) {
        PE[i] = tokens0[i];
        Q[i] = qbuf[static_cast<long long>(b) * NTOK * DD + i];
    }
    __syncthreads();
    tok_axpy(X, Q, PE, 1.f, NTOK * DD);
    tok_linear<NTOK>(w.t2i.qw, w.t2i.qb, X, DD, DD, T1, DI, DI, false, RED, cl, SYNC_BOTH);
    tok_image_attention(T1, kimg + static_cast<long long>(b) * T * DI, vimg + static_cast<long long>(b) * T * DI, T, T2, RED, HID, cl);
    tok_linear<NTOK>(w.t2i.ow, w.t2i.ob, T2, DI, DI, T1, DD, DD, false, RED, cl, SYNC_BOTH);
    tok_add_ln(Q, Q, T1, w.nf, eps, NTOK);
    if (cl.rank == 0)
        for (int i = threadIdx.x; i < NTOK * DD; i += blockDim.x) hs[static_cast<long long>(b) * NTOK * DD + i] = Q[i];
    // six independent 3-layer MLPs, layer by layer: hyper-networks 0..3 on the mask tokens, edge_mlp on the edge token,
    // the IoU head on the IoU token
    constexpr int NMLP = NMASK + 2;
    for (int i = 0; i < NMLP; ++i) {
        const Mlp3W& m = i <= NMASK ? w.hyper[i] : w.iou;
        const float* x = i <= NMASK ? Q + (1 + i) * DD : Q;
        tok_linear<1>(m.w0, m.b0, x, DD, DD, T1 + i * DD, DD, DD, true, RED, cl, (i == 0 ? SYNC_BEFORE : 0) | (i == NMLP - 1 ? SYNC_AFTER : 0));
    }
    for (int i = 0; i < NMLP; ++i) {
        const Mlp3W& m = i <= NMASK ? w.hyper[i] : w.iou;
        tok_linear<1>(m.w1, m.b1, T1 + i * DD, DD, DD, T2 + i * DD, DD, DD, true, RED, cl, (i == 0 ? SYNC_BEFORE : 0) | (i == NMLP - 1 ? SYNC_AFTER : 0));
    }
    for (int i = 0; i < NMLP; ++i) {
        const Mlp3W& m = i <= NMASK ? w.hyper[i] : w.iou;
        const int od = i <= NMASK ? DE : NMASK;
        tok_linear<1>(m.w2, m.b2, T2 + i * DD, DD, DD, T3 + i * DE, od, od, false, RED, cl, (i == 0 ? SYNC_BEFORE : 0) | (i == NMLP - 1 ? SYNC_AFTER : 0));
    }
    if (cl.rank == 0) {
        for (int i = threadIdx.x; i < (NMASK + 1) * DE; i += blockDim.x) hyper[static_cast<long long>(b) * (NMASK + 1) * DE + i] = T3[i];
        for (int i = threadIdx.x; i < NMASK; i += blockDim.x) iou[b * NMASK + i] = T3[(NMASK + 1) * DE + i];
    }
    cl_sync(cl);
}

// ------------------------------------------------------------------------------------------------------------------
// The 3x3 (transposed) convolutions run as implicit GEMMs (GemmArgs::conv_pitch) over maps in PADDED RASTER order:
// image b, pixel (Y, X) of the S x S map lives at row b*(S+2)^2 + (Y+1)*(S+2) + (X+1) of a [*, 64] bf16 matrix whose
// border rows are zero.  raster -> nested index helper, and the kernel that re-lays the nested 32-channel up-scaled map
// out that way (channels 32..63 and the border are zero).  One thread per 16-byte chunk; grid.y = image.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int raster_to_nested(int Y, int X, int g) {
    return ((((Y >> 2) * g + (X >> 2)) * 4 + ((Y >> 1) & 1) * 2 + ((X >> 1) & 1)) << 2) + (Y & 1) * 2 + (X & 1);
}

__global__ void __launch_bounds__(256) dec_pad_raster_kernel(const __nv_bfloat16* __restrict__ x_nested,
                                                             __nv_bfloat16* __restrict__ out, int g) {
    const int S = 4 * g, pitch = S + 2, PP = pitch * pitch;
    const unsigned idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= static_cast<unsigned>(PP) * 8u) return;
    const int pr = static_cast<int>(idx >> 3), chunk = static_cast<int>(idx & 7);
    const int Y = pr / pitch - 1, X = pr % pitch - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (chunk < DE / 8 && Y >= 0 && Y < S && X >= 0 && X < S)
        v = __ldg(reinterpret_cast<const uint4*>(x_nested + (static_cast<long long>(blockIdx.y) * S * S + raster_to_nested(Y, X, g)) * DE) + chunk);
    reinterpret_cast<uint4*>(out + (static_cast<long long>(blockIdx.y) * PP + pr) * DQ)[chunk] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// masks = hyper[:4] . upscaled, edge = sigmoid(hyper[4] . edge_embedding), masks = masks * edge + masks
// (mask_decoder_edge.py:176-183).  `up` and the edge features are in nested order, the output of
// embedding_maskfeature in padded raster order, the outputs in raster order.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_masks_kernel(const float* __restrict__ up, const float* __restrict__ ee,
                                                        const float* __restrict__ ef, const float* __restrict__ hyper,
                                                        float* __restrict__ masks, float* __restrict__ edge, int g) {
    // 8 lanes per pixel (one float4 of the 32 channels each: a warp reads four whole 128-byte rows per load), pixels
    // walked in nested order, the five 32-term dot products finished with three shuffles
    __shared__ float4 hy[NMASK + 1][DE / 4];
    const int b = blockIdx.y, S = 4 * g, P2 = S * S, pitch = S + 2;
    if (threadIdx.x < (NMASK + 1) * DE / 4)
        hy[threadIdx.x / (DE / 4)][threadIdx.x % (DE / 4)] = reinterpret_cast<const float4*>(hyper + b * (NMASK + 1) * DE)[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;      // P2 is a multiple of 32
    const int p = n >> 4, s1 = (n >> 2) & 3, s2 = n & 3;
    const int Y = (p / g) * 4 + (s1 >> 1) * 2 + (s2 >> 1), X = (p % g) * 4 + (s1 & 1) * 2 + (s2 & 1);
    const long long row = static_cast<long long>(b) * P2 + n;                                          // nested maps
    const long long prow = static_cast<long long>(b) * pitch * pitch + (Y + 1) * pitch + (X + 1);     // padded raster map
    const float4 u = reinterpret_cast<const float4*>(up + row * DE)[sub];
    const float4 c = reinterpret_cast<const float4*>(ee + prow * DE)[sub];
    const float4 f = reinterpret_cast<const float4*>(ef + row * DE)[sub];
    // edge_embedding = embedding_maskfeature(upscaled) + edge features (mask_decoder_edge.py:168)
    const float4 e = make_float4(c.x + f.x, c.y + f.y, c.z + f.z, c.w + f.w);
    float d[NMASK + 1];
#pragma unroll
    for (int i = 0; i < NMASK; ++i) {
        const float4 hv = hy[i][sub];
        d[i] = fmaf(hv.x, u.x, fmaf(hv.y, u.y, fmaf(hv.z, u.z, hv.w * u.w)));
    }
    {
        const float4 hv = hy[NMASK][sub];
        d[NMASK] = fmaf(hv.x, e.x, fmaf(hv.y, e.y, fmaf(hv.z, e.z, hv.w * e.w)));
    }
#pragma unroll
    for (int i = 0; i <= NMASK; ++i) {
        d[i] += __shfl_xor_sync(0xffffffffu, d[i], 1);
        d[i] += __shfl_xor_sync(0xffffffffu, d[i], 2);
        d[i] += __shfl_xor_sync(0xffffffffu, d[i], 4);
    }
    if (sub == 0) {
        const float ed = 1.f / (1.f + expf(-d[NMASK]));
        const int pix = Y * S + X;
        edge[static_cast<long long>(b) * P2 + pix] = ed;
#pragma unroll
        for (int i = 0; i < NMASK; ++i)
            masks[(static_cast<long long>(b) * NMASK + i) * P2 + pix] = __fadd_rn(__fmul_rn(d[i], ed), d[i]);
    }
}

// F.interpolate(mode="bilinear", align_corners=False) on fp32 NCHW planes (postprocess_masks,
// models/sam_maskdecoder_edge.py:380-387)
template <int VEC>
__global__ void __launch_bounds__(256) bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              int ih, int iw, int oh, int ow, float sh, float sw) {
    // block = (64 * VEC) x 4 output pixels of plane blockIdx.z; a thread writes VEC consecutive pixels of one row
    const int ox0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * VEC, oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ox0 >= ow || oy >= oh) return;
    const float fy = fmaxf(sh * (oy + 0.5f) - 0.5f, 0.f);
    const int y0 = min(static_cast<int>(fy), ih - 1), y1 = min(y0 + 1, ih - 1);
    const float ly = fy - y0, hy = 1.f - ly;
    const float* p = in + static_cast<long long>(blockIdx.z) * ih * iw;
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int ox = ox0 + j;
        const float fx = fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
        const int x0 = min(static_cast<int>(fx), iw - 1), x1 = min(x0 + 1, iw - 1);
        const float lx = fx - x0, hx = 1.f - lx;
        v[j] = hy * (hx * __ldg(p + y0 * iw + x0) + lx * __ldg(p + y0 * iw + x1)) +
               ly * (hx * __ldg(p + y1 * iw + x0) + lx * __ldg(p + y1 * iw + x1));
    }
    float* o = out + (static_cast<long long>(blockIdx.z) * oh + oy) * ow + ox0;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else o[0] = v[0];
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct DecWorkspace {
    long long keys, kb, kpb, eb, pos, kimg, vimg, qi, ao, u1, h1, up, upb, ef, upp, m1, mh, ee, qbuf, kc, vc, ktok, vtok,
        hyper, hs, total;
};

DecWorkspace plan(const cvlm_decoder_config& c, int batch) {
    const long long T = static_cast<long long>(c.grid) * c.grid, M = T * batch, P1 = 4 * M, P2 = 16 * M;
    DecWorkspace w;
    long long off = 0;
    auto take = [&](long long bytes) {
        const long long o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    w.keys = take(M * DD * 4);
    w.kb = take(M * DD * 2);
    w.kpb = take(M * DD * 2);
    w.eb = take(M * DD * 2);
    w.pos = take(T * DD * 4);
    w.kimg = take(M * DI * 4);
    w.vimg = take(M * DI * 4);
    w.qi = take(M * DI * 4);
    w.ao = take(M * DI * 2);
    w.u1 = take(M * DD * 4);
    w.h1 = take(P1 * DQ * 2);
    w.up = take(P2 * DE * 4);
    w.upb = take(P2 * DE * 2);
    w.ef = take(P2 * DE * 4);
    const long long PP = static_cast<long long>(batch) * (4 * c.grid + 2) * (4 * c.grid + 2);   // padded raster pixels
    w.upp = take(PP * DQ * 2);
    w.m1 = take(PP * DQ * 4);
    w.mh = take(PP * DQ * 2);
    w.ee = take(PP * DE * 4);
    w.qbuf = take(static_cast<long long>(batch) * NTOK * DD * 4);
    w.kc = take(static_cast<long long>(batch) * MAXC * DI * 4);
    w.vc = take(static_cast<long long>(batch) * MAXC * DI * 4);
    w.ktok = take(static_cast<long long>(batch) * NTOK * DI * 4);
    w.vtok = take(static_cast<long long>(batch) * NTOK * DI * 4);
    w.hyper = take(static_cast<long long>(batch) * (NMASK + 1) * DE * 4);
    w.hs = take(static_cast<long long>(batch) * NTOK * DD * 4);
    w.total = off;
    return w;
}

int check_config(const cvlm_decoder_config& c) {
    if (c.dim != DD || c.heads != NH || c.num_mask_tokens != NMASK)
        return set_error(CVLM_ERR_UNSUPPORTED, "decoder: transformer_dim %d / heads %d / mask tokens %d (kernels are built for 256 / 8 / 4)",
                         c.dim, c.heads, c.num_mask_tokens);
    if (c.depth < 1 || c.depth > 8) return set_error(CVLM_ERR_UNSUPPORTED, "decoder: depth %d outside [1, 8]", c.depth);
    if (c.mlp_dim < 128 || c.mlp_dim % 128 || c.mlp_dim > 4096)
        return set_error(CVLM_ERR_UNSUPPORTED, "decoder: mlp_dim %d must be a multiple of 128 up to 4096", c.mlp_dim);
    if (c.grid < 8 || c.grid % 8 || c.grid > 128)
        return set_error(CVLM_ERR_UNSUPPORTED, "decoder: embedding side %d must be a multiple of 8 in [8, 128]", c.grid);
    if (c.cond_tokens < 1 || c.cond_tokens > MAXC)
        return set_error(CVLM_ERR_UNSUPPORTED, "decoder: %d prompt tokens per image outside [1, %d]", c.cond_tokens, MAXC);
    return CVLM_OK;
}

struct DecFwd {
    cvlm_decoder* dec;
    cudaStream_t stream;
    long long launches = 0;
    char name[160];

    int get(const char* nm, long long bytes, const void** out) {
        auto it = dec->tensors.find(nm);
        if (it == dec->tensors.end()) return set_error(CVLM_ERR_MISSING, "decoder tensor '%s' was not registered", nm);
        if (bytes > 0 && it->second.second != bytes)
            return set_error(CVLM_ERR_MISSING, "decoder tensor '%s' has %lld bytes, expected %lld", nm, it->second.second, bytes);
        *out = it->second.first;
        return CVLM_OK;
    }
    int getf(const char* prefix, const char* suffix, long long elems, const float** out) {
        snprintf(name, sizeof(name), "%s%s", prefix, suffix);
        const void* p = nullptr;
        CVLM_TRY(get(name, elems * 4, &p));
        *out = static_cast<const float*>(p);
        return CVLM_OK;
    }
    int geth(const char* prefix, const char* suffix, long long elems, const void** out) {
        snprintf(name, sizeof(name), "%s%s", prefix, suffix);
        return get(name, elems * 2, out);
    }
    int attn(const char* prefix, int inner, AttnW* a) {
        char p[128];
        snprintf(p, sizeof(p), "%s", prefix);
        CVLM_TRY(getf(p, ".q_proj.weight.t", static_cast<long long>(inner) * DD, &a->qw));
        CVLM_TRY(getf(p, ".q_proj.bias", inner, &a->qb));
        CVLM_TRY(getf(p, ".k_proj.weight.t", static_cast<long long>(inner) * DD, &a->kw));
        CVLM_TRY(getf(p, ".k_proj.bias", inner, &a->kb));
        CVLM_TRY(getf(p, ".v_proj.weight.t", static_cast<long long>(inner) * DD, &a->vw));
        CVLM_TRY(getf(p, ".v_proj.bias", inner, &a->vb));
        CVLM_TRY(getf(p, ".out_proj.weight.t", static_cast<long long>(inner) * DD, &a->ow));
        CVLM_TRY(getf(p, ".out_proj.bias", DD, &a->ob));
        return CVLM_OK;
    }
    int norm(const char* prefix, int c, NormW* n) {
        char p[128];
        snprintf(p, sizeof(p), "%s", prefix);
        CVLM_TRY(getf(p, ".weight", c, &n->w));
        CVLM_TRY(getf(p, ".bias", c, &n->b));
        return CVLM_OK;
    }
    int mlp3(const char* prefix, int hidden, int out, Mlp3W* m) {
        char p[128];
        snprintf(p, sizeof(p), "%s", prefix);
        CVLM_TRY(getf(p, ".layers.0.weight.t", static_cast<long long>(hidden) * DD, &m->w0));
        CVLM_TRY(getf(p, ".layers.0.bias", hidden, &m->b0));
        CVLM_TRY(getf(p, ".layers.1.weight.t", static_cast<long long>(hidden) * hidden, &m->w1));
        CVLM_TRY(getf(p, ".layers.1.bias", hidden, &m->b1));
        CVLM_TRY(getf(p, ".layers.2.weight.t", static_cast<long long>(out) * hidden, &m->w2));
        CVLM_TRY(getf(p, ".layers.2.bias", out, &m->b2));
        return CVLM_OK;
    }
    // out[M, n] = a[M, k] * w[n, k]^T + bias (+ residual), fp32 output unless out_bf16; optional GELU / bf16 copy
    int gemm(const void* a, long long m, int k, const void* w, int n, const float* bias, const float* residual, void* out,
             int gelu = 0, void* copy = nullptr, int conv_pitch = 0) {
        GemmArgs g;
        g.a1 = a; g.lda1 = conv_pitch > 0 ? DQ : k; g.b1 = w; g.ldb1 = k; g.k1 = k;
        g.conv_pitch = conv_pitch;
        g.m = static_cast<int>(m); g.n = n;
        g.bias = bias; g.residual = residual; g.ldr = n;
        g.out = out; g.ldo = n; g.out_bf16 = 0; g.act_gelu = gelu;
        g.copy_bf16 = copy; g.ldc = n;
        ++launches;
        return gemm_launch(g, stream);
    }
};

template <int C>
int ln_rows(DecFwd& f, const float* x, const float* gamma, const float* beta, float eps, int gelu, const float* pos, int T,
            float* out_f32, void* out_b16, void* out_pos_b16, long long rows, int pad_pitch = 0) {
    ++f.launches;
    constexpr int rows_per_block = 8 * (C >= 128 ? 1 : 2 * (128 / C));
    dec_ln_rows_kernel<C><<<static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block), 256, 0, f.stream>>>(
        x, gamma, beta, eps, gelu, pos, T, out_f32, static_cast<__nv_bfloat16*>(out_b16),
        static_cast<__nv_bfloat16*>(out_pos_b16), rows, pad_pitch);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

int decoder_forward(cvlm_decoder* dec, const float* emb, const float* image_pe, const float* sparse, const float* dense,
                    int dense_batch, int batch, float* masks, float* edge, float* iou, float* hs_out, void* workspace,
                    long long workspace_bytes, cudaStream_t stream) {
    const cvlm_decoder_config& c = dec->cfg;
    if (!emb || !image_pe || !sparse || !dense || !masks || !edge || !iou || !workspace)
        return set_error(CVLM_ERR_ARG, "decoder_forward: null pointer");
    if (batch < 1) return set_error(CVLM_ERR_ARG, "decoder_forward: batch %d", batch);
    if (dense_batch != 1 && dense_batch != batch)
        return set_error(CVLM_ERR_ARG, "decoder_forward: dense prompt batch %d must be 1 or %d", dense_batch, batch);
    const DecWorkspace ws = plan(c, batch);
    if (workspace_bytes < ws.total)
        return set_error(CVLM_ERR_WORKSPACE, "decoder_forward: workspace %lld < %lld bytes", workspace_bytes, ws.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error(CVLM_ERR_ARG, "decoder_forward: workspace must be 256-byte aligned");
    const int g = c.grid, T = g * g, nc = c.cond_tokens;
    const long long M = static_cast<long long>(T) * batch, P1 = 4 * M;
    if (static_cast<long long>(batch) * (4 * g + 2) * (4 * g + 2) >= (1ll << 31)) return set_error(CVLM_ERR_UNSUPPORTED, "decoder_forward: batch %d too large for one call (chunk it)", batch);
    char* base = static_cast<char*>(workspace);
    auto F32 = [&](long long off) { return reinterpret_cast<float*>(base + off); };
    auto PTR = [&](long long off) { return static_cast<void*>(base + off); };
    DecFwd f{dec, stream};
    const int hid_floats = NTOK * c.mlp_dim > HID_MIN_FLOATS ? NTOK * c.mlp_dim : HID_MIN_FLOATS;
    const size_t tok_smem = static_cast<size_t>(SM_HID + hid_floats) * sizeof(float);
    if (tok_smem > 220 * 1024) return set_error(CVLM_ERR_UNSUPPORTED, "decoder_forward: mlp_dim %d needs too much shared memory", c.mlp_dim);
    // per-device one-time setup (function attributes belong to the device's context); one process may drive several GPUs
    int dev = 0, sms = 148;
    CVLM_CUDA(cudaGetDevice(&dev));
    CVLM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // first calls may come from several host threads at once: the flags are acquire/release bytes and the probe results
    // relaxed ints that every thread would compute to the same value (a lost race repeats idempotent work)
    static unsigned char attr_layer[kMaxDevices] = {}, attr_final[kMaxDevices] = {};
    static int cluster_ok[kMaxDevices][9] = {};      // per cluster size: 0 unknown, 1 schedulable, -1 not
    if (dev < 0 || dev >= kMaxDevices) return set_error(CVLM_ERR_UNSUPPORTED, "decoder_forward: device index %d", dev);
    const int di = dev;
    CVLM_TRY(ensure_dynamic_smem(dec_token_layer_kernel, 220 * 1024, attr_layer));
    CVLM_TRY(ensure_dynamic_smem(dec_token_final_kernel, 220 * 1024, attr_final));
    // CTAs per image on the token side: as many as keep the whole batch in one wave (option "dec_cluster" overrides: tests)
    int csize = 8;
    while (csize > 1 && batch * csize > sms) csize >>= 1;
    {
        const int v = option(OPT_DEC_CLUSTER);       // cvlm_set_option("dec_cluster", n): tests
        if (v == 1 || v == 2 || v == 4 || v == 8) csize = v;
    }
    // a cluster of `csize` CTAs with this much shared memory each must be schedulable: ask once per size, fall back
    for (; csize > 1; csize >>= 1) {
        if (__atomic_load_n(&cluster_ok[di][csize], __ATOMIC_RELAXED) == 0) {
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = csize;
            qa[0].val.clusterDim.y = 1;
            qa[0].val.clusterDim.z = 1;
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(csize);
            qc.blockDim = dim3(TOK_THREADS);
            qc.dynamicSmemBytes = tok_smem;
            qc.attrs = qa;
            qc.numAttrs = 1;
            int n = 0;
            const cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, dec_token_layer_kernel, &qc);
            if (qe != cudaSuccess) cudaGetLastError();
            __atomic_store_n(&cluster_ok[di][csize], (qe == cudaSuccess && n > 0) ? 1 : -1, __ATOMIC_RELAXED);
        }
        if (cluster_ok[di][csize] == 1) break;
    }
    cudaLaunchAttribute tok_attr[1];
    tok_attr[0].id = cudaLaunchAttributeClusterDimension;
    tok_attr[0].val.clusterDim.x = csize;
    tok_attr[0].val.clusterDim.y = 1;
    tok_attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t tok_cfg = {};
    tok_cfg.gridDim = dim3(batch * csize);
    tok_cfg.blockDim = dim3(TOK_THREADS);
    tok_cfg.dynamicSmemBytes = tok_smem;
    tok_cfg.stream = stream;
    tok_cfg.attrs = tok_attr;
    tok_cfg.numAttrs = 1;

    // tokens: [iou_token; mask_tokens; edge_token] (mask_decoder_edge.py:157), packed by the host side
    const float* tokens0 = nullptr;
    CVLM_TRY(f.getf("tokens", "", NTOK * DD, &tokens0));

    // ---- layout + prompt add
    ++f.launches;
    dec_prep_kernel<<<dim3(T / 32, batch), 256, 0, stream>>>(
        emb, dense, dense_batch == 1 ? 0 : static_cast<long long>(DD) * T, image_pe, F32(ws.keys),
        static_cast<__nv_bfloat16*>(PTR(ws.kb)), static_cast<__nv_bfloat16*>(PTR(ws.kpb)),
        static_cast<__nv_bfloat16*>(PTR(ws.eb)), F32(ws.pos), T);
    CVLM_CUDA(cudaGetLastError());

    // ---- embedding_encoder(image_embeddings) (mask_decoder_edge.py:124, :82-87): edge features, nested order
    {
        const void *w0, *w3;
        const float *b0, *b3, *lw, *lb;
        CVLM_TRY(f.geth("embedding_encoder.0", ".wp", 4ll * DQ * DD, &w0));
        CVLM_TRY(f.getf("embedding_encoder.0", ".bp", 4 * DQ, &b0));
        CVLM_TRY(f.getf("embedding_encoder.1", ".weight", DQ, &lw));
        CVLM_TRY(f.getf("embedding_encoder.1", ".bias", DQ, &lb));
        CVLM_TRY(f.geth("embedding_encoder.3", ".wp", 4ll * DE * DQ, &w3));
        CVLM_TRY(f.getf("embedding_encoder.3", ".bp", 4 * DE, &b3));
        CVLM_TRY(f.gemm(PTR(ws.eb), M, DD, w0, 4 * DQ, b0, nullptr, PTR(ws.u1)));
        CVLM_TRY(ln_rows<DQ>(f, F32(ws.u1), lw, lb, c.ln2d_eps, 1, nullptr, 1, nullptr, PTR(ws.h1), nullptr, P1));
        CVLM_TRY(f.gemm(PTR(ws.h1), P1, DQ, w3, 4 * DE, b3, nullptr, PTR(ws.ef)));
    }

    // ---- TwoWayTransformer
    char pre[128], sub[160];
    for (int i = 0; i < c.depth; ++i) {
        snprintf(pre, sizeof(pre), "transformer.layers.%d", i);
        LayerW lw;
        lw.mlp_dim = c.mlp_dim;
        snprintf(sub, sizeof(sub), "%s.self_attn", pre);                  CVLM_TRY(f.attn(sub, DD, &lw.self_attn));
        snprintf(sub, sizeof(sub), "%s.cross_attn_token_to_image", pre);  CVLM_TRY(f.attn(sub, DI, &lw.t2i));
        snprintf(sub, sizeof(sub), "%s.cross_attn_token_to_cond", pre);   CVLM_TRY(f.attn(sub, DI, &lw.t2c));
        snprintf(sub, sizeof(sub), "%s.cross_attn_image_to_token", pre);  CVLM_TRY(f.attn(sub, DI, &lw.i2t));
        snprintf(sub, sizeof(sub), "%s.cross_attn_image_to_cond", pre);   CVLM_TRY(f.attn(sub, DI, &lw.i2c));
        snprintf(sub, sizeof(sub), "%s.norm1", pre);                      CVLM_TRY(f.norm(sub, DD, &lw.n1));
        snprintf(sub, sizeof(sub), "%s.norm2", pre);                      CVLM_TRY(f.norm(sub, DD, &lw.n2));
        snprintf(sub, sizeof(sub), "%s.norm2_cond", pre);                 CVLM_TRY(f.norm(sub, DD, &lw.n2c));
        snprintf(sub, sizeof(sub), "%s.norm3", pre);                      CVLM_TRY(f.norm(sub, DD, &lw.n3));
        NormW n4, n4c;
        snprintf(sub, sizeof(sub), "%s.norm4", pre);                      CVLM_TRY(f.norm(sub, DD, &n4));
        snprintf(sub, sizeof(sub), "%s.norm4_cond", pre);                 CVLM_TRY(f.norm(sub, DD, &n4c));
        snprintf(sub, sizeof(sub), "%s.mlp.lin1", pre);
        CVLM_TRY(f.getf(sub, ".weight.t", static_cast<long long>(c.mlp_dim) * DD, &lw.l1w));
        CVLM_TRY(f.getf(sub, ".bias", c.mlp_dim, &lw.l1b));
        snprintf(sub, sizeof(sub), "%s.mlp.lin2", pre);
        CVLM_TRY(f.getf(sub, ".weight.t", static_cast<long long>(c.mlp_dim) * DD, &lw.l2w));
        CVLM_TRY(f.getf(sub, ".bias", DD, &lw.l2b));
        const void *wk, *wv, *wqc, *woc, *wqt, *wot;
        snprintf(sub, sizeof(sub), "%s.cross_attn_token_to_image", pre);
        CVLM_TRY(f.geth(sub, ".k_proj.weight.bf16", static_cast<long long>(DI) * DD, &wk));
        CVLM_TRY(f.geth(sub, ".v_proj.weight.bf16", static_cast<long long>(DI) * DD, &wv));
        snprintf(sub, sizeof(sub), "%s.cross_attn_image_to_cond", pre);
        CVLM_TRY(f.geth(sub, ".q_proj.weight.bf16", static_cast<long long>(DI) * DD, &wqc));
        CVLM_TRY(f.geth(sub, ".out_proj.weight.bf16", static_cast<long long>(DI) * DD, &woc));
        snprintf(sub, sizeof(sub), "%s.cross_attn_image_to_token", pre);
        CVLM_TRY(f.geth(sub, ".q_proj.weight.bf16", static_cast<long long>(DI) * DD, &wqt));
        CVLM_TRY(f.geth(sub, ".out_proj.weight.bf16", static_cast<long long>(DI) * DD, &wot));

        // token side needs k = (keys + pe) Wk, v = keys Wv of the CURRENT keys
        CVLM_TRY(f.gemm(PTR(ws.kpb), M, DD, wk, DI, lw.t2i.kb, nullptr, PTR(ws.kimg)));
        CVLM_TRY(f.gemm(PTR(ws.kb), M, DD, wv, DI, lw.t2i.vb, nullptr, PTR(ws.vimg)));
        ++f.launches;
        CVLM_CUDA(cudaLaunchKernelEx(&tok_cfg, dec_token_layer_kernel, lw, i == 0 ? 1 : 0, nc, T, c.ln_eps, tokens0, sparse,
                                     F32(ws.qbuf), static_cast<const float*>(F32(ws.kimg)), static_cast<const float*>(F32(ws.vimg)),
                                     F32(ws.kc), F32(ws.vc), F32(ws.ktok), F32(ws.vtok)));
        // image -> prompt attention, norm4_cond (:201-206)
        CVLM_TRY(f.gemm(PTR(ws.kpb), M, DD, wqc, DI, lw.i2c.qb, nullptr, PTR(ws.qi)));
        ++f.launches;
        dec_few_key_attn_kernel<<<static_cast<unsigned>(M / 32), 256, 0, stream>>>(F32(ws.qi), F32(ws.kc), F32(ws.vc), nc,
                                                                                   static_cast<__nv_bfloat16*>(PTR(ws.ao)), T);
        CVLM_CUDA(cudaGetLastError());
        CVLM_TRY(f.gemm(PTR(ws.ao), M, DI, woc, DD, lw.i2c.ob, F32(ws.keys), PTR(ws.keys)));
        CVLM_TRY(ln_rows<DD>(f, F32(ws.keys), n4c.w, n4c.b, c.ln_eps, 0, F32(ws.pos), T, F32(ws.keys), PTR(ws.kb), PTR(ws.kpb), M));
        // image -> token attention, norm4 (:208-212)
        CVLM_TRY(f.gemm(PTR(ws.kpb), M, DD, wqt, DI, lw.i2t.qb, nullptr, PTR(ws.qi)));
        ++f.launches;
        dec_few_key_attn_kernel<<<static_cast<unsigned>(M / 32), 256, 0, stream>>>(F32(ws.qi), F32(ws.ktok), F32(ws.vtok), NTOK,
                                                                                   static_cast<__nv_bfloat16*>(PTR(ws.ao)), T);
        CVLM_CUDA(cudaGetLastError());
        CVLM_TRY(f.gemm(PTR(ws.ao), M, DI, wot, DD, lw.i2t.ob, F32(ws.keys), PTR(ws.keys)));
        CVLM_TRY(ln_rows<DD>(f, F32(ws.keys), n4.w, n4.b, c.ln_eps, 0, F32(ws.pos), T, F32(ws.keys), PTR(ws.kb), PTR(ws.kpb), M));
    }
    // ---- final token -> image attention, hyper-networks, IoU head
    {
        FinalW fw;
        CVLM_TRY(f.attn("transformer.final_attn_token_to_image", DI, &fw.t2i));
        CVLM_TRY(f.norm("transformer.norm_final_attn", DD, &fw.nf));
        for (int i = 0; i < NMASK; ++i) {
            snprintf(sub, sizeof(sub), "output_hypernetworks_mlps.%d", i);
            CVLM_TRY(f.mlp3(sub, DD, DE, &fw.hyper[i]));
        }
        CVLM_TRY(f.mlp3("edge_mlp", DD, DE, &fw.hyper[NMASK]));
        CVLM_TRY(f.mlp3("iou_prediction_head", DD, NMASK, &fw.iou));
        const void *wk, *wv;
        CVLM_TRY(f.geth("transformer.final_attn_token_to_image", ".k_proj.weight.bf16", static_cast<long long>(DI) * DD, &wk));
        CVLM_TRY(f.geth("transformer.final_attn_token_to_image", ".v_proj.weight.bf16", static_cast<long long>(DI) * DD, &wv));
        CVLM_TRY(f.gemm(PTR(ws.kpb), M, DD, wk, DI, fw.t2i.kb, nullptr, PTR(ws.kimg)));
        CVLM_TRY(f.gemm(PTR(ws.kb), M, DD, wv, DI, fw.t2i.vb, nullptr, PTR(ws.vimg)));
        ++f.launches;
        CVLM_CUDA(cudaLaunchKernelEx(&tok_cfg, dec_token_final_kernel, fw, T, c.ln_eps, tokens0, static_cast<const float*>(F32(ws.qbuf)),
                                     static_cast<const float*>(F32(ws.kimg)), static_cast<const float*>(F32(ws.vimg)),
                                     hs_out ? hs_out : F32(ws.hs), F32(ws.hyper), iou));
    }
    // ---- output_upscaling(keys) (mask_decoder_edge.py:165-166, :53-59)
    {
        const void *w0, *w3;
        const float *b0, *b3, *lw, *lb;
        CVLM_TRY(f.geth("output_upscaling.0", ".wp", 4ll * DQ * DD, &w0));
        CVLM_TRY(f.getf("output_upscaling.0", ".bp", 4 * DQ, &b0));
        CVLM_TRY(f.getf("output_upscaling.1", ".weight", DQ, &lw));
        CVLM_TRY(f.getf("output_upscaling.1", ".bias", DQ, &lb));
        CVLM_TRY(f.geth("output_upscaling.3", ".wp", 4ll * DE * DQ, &w3));
        CVLM_TRY(f.getf("output_upscaling.3", ".bp", 4 * DE, &b3));
        CVLM_TRY(f.gemm(PTR(ws.kb), M, DD, w0, 4 * DQ, b0, nullptr, PTR(ws.u1)));
        CVLM_TRY(ln_rows<DQ>(f, F32(ws.u1), lw, lb, c.ln2d_eps, 1, nullptr, 1, nullptr, PTR(ws.h1), nullptr, P1));
        CVLM_TRY(f.gemm(PTR(ws.h1), P1, DQ, w3, 4 * DE, b3, nullptr, PTR(ws.up), 1, PTR(ws.upb)));
    }
    // ---- embedding_maskfeature(upscaled) (:168, :88-93): two 3x3 s1 p1 transposed convolutions = 3x3 convolutions with
    // the flipped kernel, as implicit GEMMs over zero-bordered raster maps (no im2col); the edge features are added by the
    // mask kernel
    {
        const void *w0, *w3;
        const float *b0, *b3, *lw, *lb;
        const int pitch = 4 * g + 2;
        const long long PP = static_cast<long long>(batch) * pitch * pitch;
        CVLM_TRY(f.geth("embedding_maskfeature.0", ".wp", static_cast<long long>(DQ) * 9 * DQ, &w0));
        CVLM_TRY(f.getf("embedding_maskfeature.0", ".bias", DQ, &b0));
        CVLM_TRY(f.getf("embedding_maskfeature.1", ".weight", DQ, &lw));
        CVLM_TRY(f.getf("embedding_maskfeature.1", ".bias", DQ, &lb));
        CVLM_TRY(f.geth("embedding_maskfeature.3", ".wp", static_cast<long long>(DE) * 9 * DQ, &w3));
        CVLM_TRY(f.getf("embedding_maskfeature.3", ".bias", DE, &b3));
        ++f.launches;
        dec_pad_raster_kernel<<<dim3((pitch * pitch * 8 + 255) / 256, batch), 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(PTR(ws.upb)), static_cast<__nv_bfloat16*>(PTR(ws.upp)), g);
        CVLM_CUDA(cudaGetLastError());
        CVLM_TRY(f.gemm(PTR(ws.upp), PP, 9 * DQ, w0, DQ, b0, nullptr, PTR(ws.m1), 0, nullptr, pitch));
        CVLM_TRY(ln_rows<DQ>(f, F32(ws.m1), lw, lb, c.ln2d_eps, 1, nullptr, 1, nullptr, PTR(ws.mh), nullptr, PP, pitch));
        CVLM_TRY(f.gemm(PTR(ws.mh), PP, 9 * DQ, w3, DE, b3, nullptr, PTR(ws.ee), 0, nullptr, pitch));
    }
    // ---- masks, edge
    ++f.launches;
    dec_masks_kernel<<<dim3(16 * T / 32, batch), 256, 0, stream>>>(F32(ws.up), F32(ws.ee), F32(ws.ef), F32(ws.hyper), masks, edge, g);
    CVLM_CUDA(cudaGetLastError());
    dec->last_launches = f.launches;
    return CVLM_OK;
}

}  // namespace

int bilinear_resize_launch(const float* in, float* out, long long planes, int ih, int iw, int oh, int ow, cudaStream_t stream) {
    if (!in || !out || planes < 1 || ih < 1 || iw < 1 || oh < 1 || ow < 1) return set_error(CVLM_ERR_ARG, "bilinear_resize: bad arguments");
    if (planes > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "bilinear_resize: more than 65535 planes in one call");
    const float sh = static_cast<float>(ih) / oh, sw = static_cast<float>(iw) / ow;
    if (ow % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        bilinear_resize_kernel<4><<<dim3((ow + 255) / 256, (oh + 3) / 4, static_cast<unsigned>(planes)), 256, 0, stream>>>(
            in, out, ih, iw, oh, ow, sh, sw);
    else
        bilinear_resize_kernel<1><<<dim3((ow + 63) / 64, (oh + 3) / 4, static_cast<unsigned>(planes)), 256, 0, stream>>>(
            in, out, ih, iw, oh, ow, sh, sw);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace cvlm

extern "C" {

int cvlm_decoder_create(const cvlm_decoder_config* cfg, cvlm_decoder** out) {
    if (!cfg || !out) return cvlm::set_error(CVLM_ERR_ARG, "decoder_create: null argument");
    CVLM_TRY(cvlm::check_config(*cfg));
    cvlm_decoder* d = new cvlm_decoder();
    d->cfg = *cfg;
    *out = d;
    return CVLM_OK;
}

int cvlm_decoder_destroy(cvlm_decoder* dec) {
    delete dec;
    return CVLM_OK;
}

int cvlm_decoder_set_tensor(cvlm_decoder* dec, const char* name, const void* dev_ptr, long long nbytes) {
    if (!dec || !name || !dev_ptr || nbytes <= 0) return cvlm::set_error(CVLM_ERR_ARG, "decoder_set_tensor: bad argument");
    if (reinterpret_cast<uintptr_t>(dev_ptr) & 15) return cvlm::set_error(CVLM_ERR_ARG, "decoder_set_tensor: '%s' is not 16-byte aligned", name);
    dec->tensors[name] = {dev_ptr, nbytes};
    return CVLM_OK;
}

long long cvlm_decoder_workspace_bytes(const cvlm_decoder* dec, int batch) {
    if (!dec || batch < 1) return cvlm::set_error(CVLM_ERR_ARG, "decoder_workspace_bytes: bad argument");
    return cvlm::plan(dec->cfg, batch).total;
}

long long cvlm_decoder_last_launches(const cvlm_decoder* dec) { return dec ? dec->last_launches : 0; }

int cvlm_decoder_forward(cvlm_decoder* dec, const float* image_embeddings, const float* image_pe, const float* sparse_prompt,
                         const float* dense_prompt, int dense_batch, int batch, float* masks, float* edge, float* iou_pred,
                         float* hs, void* workspace, long long workspace_bytes, cvlm_stream_t stream) {
    if (!dec) return cvlm::set_error(CVLM_ERR_ARG, "decoder_forward: null handle");
    return cvlm::decoder_forward(dec, image_embeddings, image_pe, sparse_prompt, dense_prompt, dense_batch, batch, masks, edge,
                                 iou_pred, hs, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int cvlm_op_bilinear_resize(const float* in, float* out, long long planes, int in_h, int in_w, int out_h, int out_w,
                            cvlm_stream_t stream) {
    return cvlm::bilinear_resize_launch(in, out, planes, in_h, in_w, out_h, out_w, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
