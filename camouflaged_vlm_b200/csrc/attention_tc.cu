// Global attention (token grid 64x64, 4096 keys) on the 5th-generation tensor cores.
//
// Same contract as attention.cu (which keeps serving windows and other grid sizes): bias-free K/V in the
// token-major qkv buffer, decomposed rel-pos from the unscaled q added inside the softmax tile
// (image_encoder.py:493-501, 589-625).  One CTA owns 128 queries (two grid rows) of one (image, head) and
// streams the 32 key tiles of 128 keys (two grid rows each):
//
//   warp 0      TMA producer.  Q tile once; rel_pos_w, a 128-row slice of rel_pos_h and the K tiles go
//               through one 3-stage ring, V tiles through another (cp.async.bulk.tensor, SWIZZLE_128B).
//               head_dim 80 = a 64-wide tile + a 16-wide tile (columns 80..127 of the second box are
//               out of bounds in the [d, head, token] tensor map and arrive as zeros).
//   warp 1      MMA issuer (one lane):  S = Q K^T  (tcgen05.mma SS, M128 N128, K = 64 + 16) into one of two
//               TMEM S buffers;  O += P V  (tcgen05.mma TS: P read from TMEM, V MN-major from smem, N = hd).
//               The first two "tiles" of the K ring are the rel-pos tables: T_w = Q rel_pos_w^T and
//               T_h = Q rel_pos_h[y0..]^T come out of the same MMA path.
//   warp 2      TMEM allocator (512 columns: S0, S1, O, P0, P1).
//   warps 4-7   softmax, one thread per query row (TMEM lane = row, no shuffles).  Pass A finds the raw row
//               maximum, pass B computes exp2(scale*s + rel_h + rel_w - m) and writes bf16 P back to TMEM.
//               Rescaling of O is lazy: it only happens when the (upper bound of the) row maximum grows by
//               more than 2^8, so O stays in TMEM for almost every tile.
//
// Since a key tile is exactly two grid rows, the h-term is one scalar per (query, grid row) and the w-term
// is the same 64 values for every tile (kept in registers).
#include "cvlm_internal.cuh"
#include "ptx.cuh"
#include "tensor_map.cuh"

namespace cvlm {

constexpr int TC_TILE = 128 * 128;                 // one [128 rows x 64 bf16] swizzled tile, bytes
constexpr int TC_STAGE = 2 * TC_TILE;              // lo (d 0..63) + hi (d 64..127, only 64..79 real)
constexpr int TC_KST = 3, TC_VST = 3;
constexpr int TC_OFF_Q = 0;
constexpr int TC_OFF_K = TC_STAGE;
constexpr int TC_OFF_V = TC_OFF_K + TC_KST * TC_STAGE;
constexpr int TC_OFF_BAR = TC_OFF_V + TC_VST * TC_STAGE;
constexpr int TC_SMEM = TC_OFF_BAR + 256 + 1024;
static_assert(TC_SMEM <= 232448, "attention_tc shared memory");
constexpr int TC_THREADS = 256;
constexpr uint32_t COL_S = 0, COL_O = 256, COL_P = 336;   // TMEM columns: S0 0..127, S1 128..255, O, P0 336.., P1 400..
constexpr float kLog2eTc = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_global_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmRelH,
                      const __grid_constant__ CUtensorMap tmRelW, __nv_bfloat16* __restrict__ out, int heads, int D,
                      float scale_log2) {
    constexpr bool HI = HD > 64;
    constexpr uint32_t LOAD_BYTES = (HI ? 2 : 1) * TC_TILE;
    constexpr int NT = 32;                          // 4096 keys / 128
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base + TC_OFF_Q;
    auto sK = [&](int s) { return base + TC_OFF_K + s * TC_STAGE; };
    auto sV = [&](int s) { return base + TC_OFF_V + s * TC_STAGE; };
    const uint32_t bars = base + TC_OFF_BAR;
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bars + 8u * (4 + s); };
    auto v_full = [&](int s) { return bars + 8u * (7 + s); };
    auto v_empty = [&](int s) { return bars + 8u * (10 + s); };
    auto s_full = [&](int i) { return bars + 8u * (13 + i); };
    auto s_empty = [&](int i) { return bars + 8u * (15 + i); };
    auto p_full = [&](int i) { return bars + 8u * (17 + i); };
    auto p_empty = [&](int i) { return bars + 8u * (19 + i); };
    const uint32_t o_full = bars + 8u * 21;
    const uint32_t tmem_slot = bars + 8u * 22;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int y0 = 2 * qt;                          // first grid row of this query tile
    const int tok0 = b * 4096;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmRelH);
        tma_prefetch_desc(&tmRelW);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 3; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(s_empty(i), 128);
            mbar_init(p_full(i), 128);
            mbar_init(p_empty(i), 1);
        }
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(q_full, LOAD_BYTES);
            tma_load_3d(sQ, &tmQKV, q_full, 0, head, tok0 + qt * 128);
            if (HI) tma_load_3d(sQ + TC_TILE, &tmQKV, q_full, 64, head, tok0 + qt * 128);
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            auto produce_k = [&](int u) {   // u = 0: rel_pos_w, 1: rel_pos_h slice, 2 + j: key tile j
                mbar_wait(k_empty(kst), kph ^ 1u);
                mbar_arrive_expect_tx(k_full(kst), LOAD_BYTES);
                if (u == 0) {
                    tma_load_2d(sK(kst), &tmRelW, k_full(kst), 0, 0);
                    if (HI) tma_load_2d(sK(kst) + TC_TILE, &tmRelW, k_full(kst), 64, 0);
                } else if (u == 1) {
                    tma_load_2d(sK(kst), &tmRelH, k_full(kst), 0, y0);
                    if (HI) tma_load_2d(sK(kst) + TC_TILE, &tmRelH, k_full(kst), 64, y0);
                } else {
                    const int tok = tok0 + (u - 2) * 128;
                    tma_load_3d(sK(kst), &tmQKV, k_full(kst), 0, heads + head, tok);
                    if (HI) tma_load_3d(sK(kst) + TC_TILE, &tmQKV, k_full(kst), 64, heads + head, tok);
                }
                if (++kst == TC_KST) { kst = 0; kph ^= 1u; }
            };
            auto produce_v = [&](int j) {
                mbar_wait(v_empty(vst), vph ^ 1u);
                mbar_arrive_expect_tx(v_full(vst), LOAD_BYTES);
                const int tok = tok0 + j * 128;
                tma_load_3d(sV(vst), &tmQKV, v_full(vst), 0, 2 * heads + head, tok);
                if (HI) tma_load_3d(sV(vst) + TC_TILE, &tmQKV, v_full(vst), 64, 2 * heads + head, tok);
                if (++vst == TC_VST) { vst = 0; vph ^= 1u; }
            };
            produce_k(0);
            produce_k(1);
            produce_k(2);
            for (int j = 0; j < NT; ++j) {
                if (j + 1 < NT) produce_k(j + 3);
                produce_v(j);
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idescS = umma_idesc_bf16(128, 128);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            mbar_wait(q_full, 0);
            tc_fence_after();
            const uint64_t q_lo = umma_desc_k_sw128(sQ), q_hi = umma_desc_k_sw128(sQ + TC_TILE);
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            auto mma_s = [&](int u) {       // S[u & 1] = Q * (K-ring tile u)^T
                const int sb = u & 1;
                if (u >= 2) {
                    mbar_wait(s_empty(sb), static_cast<uint32_t>(((u >> 1) - 1) & 1));
                    tc_fence_after();
                }
                mbar_wait(k_full(kst), kph);
                tc_fence_after();
                const uint32_t d = tmem_base + COL_S + sb * 128;
                const uint64_t k_lo = umma_desc_k_sw128(sK(kst));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(d, q_lo + 2u * k, k_lo + 2u * k, idescS, k > 0);
                if (HI) umma_bf16_ss(d, q_hi, umma_desc_k_sw128(sK(kst) + TC_TILE), idescS, 1);
                umma_commit(k_empty(kst));
                umma_commit(s_full(sb));
                if (++kst == TC_KST) { kst = 0; kph ^= 1u; }
            };
            mma_s(0);
            mma_s(1);
            mma_s(2);
            for (int j = 0; j < NT; ++j) {
                if (j + 1 < NT) mma_s(j + 3);       // S of the next tile runs under this tile's softmax
                const int pb = j & 1;
                mbar_wait(p_full(pb), static_cast<uint32_t>((j >> 1) & 1));
                tc_fence_after();
                mbar_wait(v_full(vst), vph);
                tc_fence_after();
                const uint64_t v_desc = umma_desc_mn_sw128(sV(vst), TC_TILE);
                const uint32_t a = tmem_base + COL_P + pb * 64;
#pragma unroll
                for (int k = 0; k < 8; ++k)         // 16 keys per MMA: 8 TMEM columns of P, 2048 B of V
                    umma_bf16_ts(tmem_base + COL_O, a + 8u * k, v_desc + 128u * k, idescO, (j | k) != 0);
                umma_commit(v_empty(vst));
                umma_commit(p_empty(pb));
                if (j == NT - 1) umma_commit(o_full);
                if (++vst == TC_VST) { vst = 0; vph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ============================================================ softmax: one thread per query row
        const int r = threadIdx.x - 128;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int qrow = r >> 6, qx = r & 63;
        const uint32_t t_s = tmem_base + lane_addr + COL_S;
        const uint32_t t_o = tmem_base + lane_addr + COL_O;
        const uint32_t t_p = tmem_base + lane_addr + COL_P;
        float trw[128], trh[96];                    // this row of T_w / T_h (dynamic indexing -> local memory)
        uint32_t v[32];
        mbar_wait(s_full(0), 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            tmem_ld_32x32(t_s + c * 32, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int i = 0; i < 32; ++i) trw[c * 32 + i] = __uint_as_float(v[i]) * kLog2eTc;
        }
        tc_fence_before();
        mbar_arrive(s_empty(0));
        mbar_wait(s_full(1), 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            tmem_ld_32x32(t_s + 128 + c * 32, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int i = 0; i < 32; ++i) trh[c * 32 + i] = __uint_as_float(v[i]) * kLog2eTc;
        }
        tc_fence_before();
        mbar_arrive(s_empty(1));
        // rel_w[q, kx] = T_w[q][qx - kx + 63] (image_encoder.py:576-586): the same for every key tile
        float tw[64];
        float twmax = -INFINITY;
#pragma unroll
        for (int kx = 0; kx < 64; ++kx) {
            tw[kx] = trw[qx + 63 - kx];
            twmax = fmaxf(twmax, tw[kx]);
        }
        float m_used = -INFINITY, l = 0.f;

        auto scale_o = [&](float f) {   // O row *= f, in tensor memory
            uint32_t w[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld_32x32(t_o + c * 32, w);
                tmem_ld_wait(w);
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * f);
                tmem_st_32x32(t_o + c * 32, w);
            }
            if (HI) {
                uint32_t w16[16];
                tmem_ld_32x16(t_o + 64, w16);
                tmem_ld_wait16(w16);
#pragma unroll
                for (int i = 0; i < 16; ++i) w16[i] = __float_as_uint(__uint_as_float(w16[i]) * f);
                tmem_st_32x16(t_o + 64, w16);
            }
            tmem_st_wait();
        };

        for (int j = 0; j < NT; ++j) {
            const int sb = j & 1, pb = j & 1;
            mbar_wait(s_full(sb), static_cast<uint32_t>(((j + 2) >> 1) & 1));
            tc_fence_after();
            // key rows of this tile: ky = 2j, 2j+1;  rel_h[q, ky] = T_h[q][qy - ky + 63 - y0]
            const float th0 = trh[qrow + 63 - 2 * j], th1 = trh[qrow + 62 - 2 * j];
            // ---- pass A: raw row maximum
            float smax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_32x32(t_s + sb * 128 + c * 32, v);
                tmem_ld_wait(v);
#pragma unroll
                for (int i = 0; i < 32; ++i) smax = fmaxf(smax, __uint_as_float(v[i]));
            }
            // upper bound of the row's logits in this tile; exponentials use the stale reference m_used as
            // long as it is within 2^8 of that bound (lazy rescale)
            const float bound = fmaf(smax, scale_log2, fmaxf(th0, th1) + twmax);
            const bool need = bound > m_used + 8.0f;
            if (__any_sync(0xffffffffu, need)) {
                const float f = need ? ex2f(m_used - bound) : 1.0f;
                if (j > 0) {
                    mbar_wait(p_empty((j - 1) & 1), static_cast<uint32_t>(((j - 1) >> 1) & 1));   // P V of tile j-1 retired
                    tc_fence_after();
                    scale_o(f);
                }
                l *= f;
                if (need) m_used = bound;
            }
            if (j >= 2) {
                mbar_wait(p_empty(pb), static_cast<uint32_t>(((j >> 1) - 1) & 1));                 // P buffer free
                tc_fence_after();
            }
            // ---- pass B: P = exp2(scale*s + rel_h + rel_w - m_used), bf16, back into tensor memory
            const float c0 = th0 - m_used, c1 = th1 - m_used;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_32x32(t_s + sb * 128 + c * 32, v);
                tmem_ld_wait(v);
                if (c == 3) {
                    tc_fence_before();
                    mbar_arrive(s_empty(sb));       // S buffer may be overwritten by the MMA of tile j+2
                }
                const float ch = c < 2 ? c0 : c1;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = ex2f(fmaf(__uint_as_float(v[2 * i]), scale_log2, tw[(c & 1) * 32 + 2 * i] + ch));
                    const float p1 = ex2f(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, tw[(c & 1) * 32 + 2 * i + 1] + ch));
                    l += p0 + p1;
                    pk[i] = pack_bf16(p0, p1);
                }
                tmem_st_32x16(t_p + pb * 64 + c * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full(pb));
        }
        // ---- epilogue: O / l -> bf16 -> global (column = head*HD + d, image_encoder.py:501)
        mbar_wait(o_full, 0);
        tc_fence_after();
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = out + (static_cast<long long>(tok0) + qt * 128 + r) * D + head * HD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(t_o + c * 32, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 pk;
                pk.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                pk.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                pk.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                pk.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + c * 32 + i) = pk;
            }
        }
        if (HI) {
            uint32_t w16[16];
            tmem_ld_32x16(t_o + 64, w16);
            tmem_ld_wait16(w16);
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 pk;
                pk.x = pack_bf16(__uint_as_float(w16[i]) * inv, __uint_as_float(w16[i + 1]) * inv);
                pk.y = pack_bf16(__uint_as_float(w16[i + 2]) * inv, __uint_as_float(w16[i + 3]) * inv);
                pk.z = pack_bf16(__uint_as_float(w16[i + 4]) * inv, __uint_as_float(w16[i + 5]) * inv);
                pk.w = pack_bf16(__uint_as_float(w16[i + 6]) * inv, __uint_as_float(w16[i + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + 64 + i) = pk;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD>
static int launch_global_tc(const void* qkv, void* out, const void* relh, const void* relw, int batch, int heads,
                            cudaStream_t stream) {
    const int D = heads * HD;
    CUtensorMap tq, th, tw;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 3;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = 4096ull * batch;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = 128;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    MapSpec r = map_spec_zero();
    r.rank = 2;
    r.dims[0] = HD; r.dims[1] = 127;
    r.strides[0] = HD * 2ull;
    r.box[0] = 64; r.box[1] = 128;
    r.ptr = relh;
    CVLM_TRY(make_tensor_map_bf16(&th, r));
    r.ptr = relw;
    CVLM_TRY(make_tensor_map_bf16(&tw, r));
    auto kern = attn_global_tc_kernel<HD>;
    CVLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    dim3 grid(32, heads, batch);
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * kLog2eTc;
    kern<<<grid, TC_THREADS, TC_SMEM, stream>>>(tq, th, tw, static_cast<__nv_bfloat16*>(out), heads, D, scale_log2);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// Global attention over a 64x64 token grid, head_dim 64 or 80.
int attention_global_tc_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int heads,
                               int head_dim, cudaStream_t stream) {
    if (batch > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "attention: batch exceeds grid.z; split the batch");
    if (head_dim == 80) return launch_global_tc<80>(qkv, out, relh, relw, batch, heads, stream);
    if (head_dim == 64) return launch_global_tc<64>(qkv, out, relh, relw, batch, heads, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_tc: head_dim %d", head_dim);
}

}  // namespace cvlm
