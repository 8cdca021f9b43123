// 14x14 window attention on the 5th-generation tensor cores.
//
// The reference pads the 64x64 token grid to 70x70 AFTER norm1, cuts it into 25 windows of 196 tokens
// (window_partition, image_encoder.py:507-528), runs attention per window and crops (window_unpartition
// :531-553).  Here a window is one 5-D TMA box {64 d, 1 head-slot, 14 x, 14 y, 1 image} over the
// un-partitioned token-major qkv buffer: the 196 rows land in shared memory in window order and every
// token outside the grid arrives as ZEROS (TMA out-of-bounds fill) -- exactly the reference's padding,
// because K and V carry no bias (packing.py).  Nothing is materialised, nothing is copied back.
//
// Persistent CTAs loop over (image, window, head) items.  Per item: two query tiles of 128 rows
// (rows >= 196 are don't-care), one key tile of 208 columns (196 keys + 12 masked), so there is no
// online softmax.  Per query tile:
//   MMA  T = Q [rel_pos_h ; rel_pos_w]^T   (N = 64)       } tcgen05.mma SS, K = 64 + 16
//   MMA  S = Q K^T                          (N = 208)      }
//   softmax warps (thread = query row): th[ky] / tw[kx] from T, one max pass, one exp2 pass that writes
//        bf16 P over the S columns in place (P chunk c never overtakes the S chunk being read)
//   MMA  O = P V                            (tcgen05.mma TS, V MN-major, N = head_dim, K = 208)
//   epilogue: O / l -> bf16 -> global, only for queries inside the grid.
// Q is double-buffered across items; K / V are reloaded as soon as the item's MMAs that read them retire,
// which overlaps the loads of item i+1 with the softmax / P V tail of item i.
#include "cvlm_internal.cuh"
#include "ptx.cuh"
#include "tensor_map.cuh"

namespace cvlm {

constexpr int WN_WS = 14, WN_L = 196, WN_ROWS = 208;
constexpr int WN_TILE = WN_ROWS * 128;               // [208 rows x 64 bf16], 26 KB (26 swizzle atoms of 1 KB)
constexpr int WN_SLOT = 2 * WN_TILE;                 // lo + hi
constexpr int WN_OFF_Q = 0;                          // two Q slots
constexpr int WN_OFF_K = 2 * WN_SLOT;
constexpr int WN_OFF_V = WN_OFF_K + WN_SLOT;
constexpr int WN_OFF_REL = WN_OFF_V + WN_SLOT;       // [64 rows x 64] lo + hi: rows 0..26 rel_pos_h, 32..58 rel_pos_w
constexpr int WN_REL_TILE = 64 * 128;
constexpr int WN_OFF_BAR = WN_OFF_REL + 2 * WN_REL_TILE;
constexpr int WN_SMEM = WN_OFF_BAR + 256 + 1024;
static_assert(WN_SMEM <= 232448, "attention_tc_win shared memory");
constexpr int WN_BOX_BYTES = 64 * WN_L * 2;          // bytes one TMA box delivers
constexpr uint32_t WCOL_T = 0, WCOL_S = 64, WCOL_O = 272;   // TMEM: T 0..63, S/P 64..271, O0 272..351, O1 352..431
constexpr float kLog2eW = 1.4426950408889634f;

__device__ __forceinline__ float ex2w(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct WinItem {
    int b, wy, wx, head, ntiles;
};

template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_window_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmRelH,
                      const __grid_constant__ CUtensorMap tmRelW, __nv_bfloat16* __restrict__ out, int G, int nW,
                      int heads, int D, int num_items, float scale_log2) {
    constexpr bool HI = HD > 64;
    constexpr uint32_t LOAD_BYTES = (HI ? 2 : 1) * WN_BOX_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto sQ = [&](int s) { return base + WN_OFF_Q + s * WN_SLOT; };
    const uint32_t sK = base + WN_OFF_K, sV = base + WN_OFF_V, sRel = base + WN_OFF_REL;
    const uint32_t bars = base + WN_OFF_BAR;
    const uint32_t rel_full = bars;
    auto q_full = [&](int s) { return bars + 8u * (1 + s); };
    auto q_empty = [&](int s) { return bars + 8u * (3 + s); };
    const uint32_t k_full = bars + 8u * 5, k_empty = bars + 8u * 6, v_full = bars + 8u * 7, v_empty = bars + 8u * 8;
    const uint32_t s_full = bars + 8u * 9, p_full = bars + 8u * 10;
    auto o_full = [&](int t) { return bars + 8u * (11 + t); };
    auto o_empty = [&](int t) { return bars + 8u * (13 + t); };
    const uint32_t tmem_slot = bars + 8u * 15;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + WN_OFF_BAR + 8 * 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto decode = [&](int it) {
        WinItem w;
        w.head = it % heads;
        const int rest = it / heads;
        const int win = rest % (nW * nW);
        w.b = rest / (nW * nW);
        w.wy = win / nW;
        w.wx = win % nW;
        // the second query tile starts at window row 9 (token 128 = 9*14 + 2): all padding for bottom-edge windows
        w.ntiles = (w.wy * WN_WS + 9 < G) ? 2 : 1;
        return w;
    };

    // V rows 196..207 are read by the last k-step of P V (against P == 0) but never written by TMA: zero them once
    for (int i = threadIdx.x; i < 2 * 12 * 8; i += blockDim.x) {
        const int tile = i / 96, rest = i % 96;
        *reinterpret_cast<uint4*>(base_ptr + WN_OFF_V + tile * WN_TILE + WN_L * 128 + rest * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmRelH);
        tma_prefetch_desc(&tmRelW);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(rel_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(q_full(s), 1);
            mbar_init(q_empty(s), 1);
            mbar_init(o_full(s), 1);
            mbar_init(o_empty(s), 128);
        }
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
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
            mbar_arrive_expect_tx(rel_full, (HI ? 4 : 2) * 32 * 128);
            tma_load_2d(sRel, &tmRelH, rel_full, 0, 0);
            tma_load_2d(sRel + 32 * 128, &tmRelW, rel_full, 0, 0);
            if (HI) {
                tma_load_2d(sRel + WN_REL_TILE, &tmRelH, rel_full, 64, 0);
                tma_load_2d(sRel + WN_REL_TILE + 32 * 128, &tmRelW, rel_full, 64, 0);
            }
            int n = 0;
            for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
                const WinItem w = decode(it);
                const int qs = n & 1;
                const uint32_t ph = static_cast<uint32_t>(n & 1);          // K / V: one slot, phase flips per item
                const uint32_t qph = static_cast<uint32_t>((n >> 1) & 1);  // Q: two slots
                const int x0 = w.wx * WN_WS, y0 = w.wy * WN_WS;
                mbar_wait(q_empty(qs), qph ^ 1u);
                mbar_arrive_expect_tx(q_full(qs), LOAD_BYTES);
                tma_load_5d(sQ(qs), &tmQKV, q_full(qs), 0, w.head, x0, y0, w.b);
                if (HI) tma_load_5d(sQ(qs) + WN_TILE, &tmQKV, q_full(qs), 64, w.head, x0, y0, w.b);
                mbar_wait(k_empty, ph ^ 1u);
                mbar_arrive_expect_tx(k_full, LOAD_BYTES);
                tma_load_5d(sK, &tmQKV, k_full, 0, heads + w.head, x0, y0, w.b);
                if (HI) tma_load_5d(sK + WN_TILE, &tmQKV, k_full, 64, heads + w.head, x0, y0, w.b);
                mbar_wait(v_empty, ph ^ 1u);
                mbar_arrive_expect_tx(v_full, LOAD_BYTES);
                tma_load_5d(sV, &tmQKV, v_full, 0, 2 * heads + w.head, x0, y0, w.b);
                if (HI) tma_load_5d(sV + WN_TILE, &tmQKV, v_full, 64, 2 * heads + w.head, x0, y0, w.b);
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idescT = umma_idesc_bf16(128, 64);
            constexpr uint32_t idescS = umma_idesc_bf16(128, WN_ROWS);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            mbar_wait(rel_full, 0);
            tc_fence_after();
            const uint64_t rel_lo = umma_desc_k_sw128(sRel), rel_hi = umma_desc_k_sw128(sRel + WN_REL_TILE);
            const uint64_t k_lo = umma_desc_k_sw128(sK), k_hi = umma_desc_k_sw128(sK + WN_TILE);
            const uint64_t v_desc = umma_desc_mn_sw128(sV, WN_TILE);
            int n = 0;
            uint32_t tile_count = 0;                 // q-tiles processed so far (phase of s_full / p_full)
            uint32_t o_uses[2] = {0, 0};             // how many times O[t] has been produced
            for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
                const WinItem w = decode(it);
                const int qs = n & 1;
                mbar_wait(q_full(qs), static_cast<uint32_t>((n >> 1) & 1));
                mbar_wait(k_full, static_cast<uint32_t>(n & 1));
                tc_fence_after();
                for (int t = 0; t < w.ntiles; ++t, ++tile_count) {
                    const uint64_t q_lo = umma_desc_k_sw128(sQ(qs) + t * 128 * 128);
                    const uint64_t q_hi = umma_desc_k_sw128(sQ(qs) + WN_TILE + t * 128 * 128);
                    // T and S of this query tile.  They may overwrite the previous tile's T / P: tcgen05.mma
                    // executes in issue order behind the previous P V, and the previous softmax finished
                    // reading T before it signalled p_full.
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + WCOL_T, q_lo + 2u * k, rel_lo + 2u * k, idescT, k > 0);
                    if (HI) umma_bf16_ss(tmem_base + WCOL_T, q_hi, rel_hi, idescT, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + WCOL_S, q_lo + 2u * k, k_lo + 2u * k, idescS, k > 0);
                    if (HI) umma_bf16_ss(tmem_base + WCOL_S, q_hi, k_hi, idescS, 1);
                    umma_commit(s_full);
                    if (t == w.ntiles - 1) {
                        umma_commit(q_empty(qs));
                        umma_commit(k_empty);
                    }
                    // O[t] = P V
                    mbar_wait(p_full, tile_count & 1u);
                    tc_fence_after();
                    if (t == 0) {
                        mbar_wait(v_full, static_cast<uint32_t>(n & 1));
                        tc_fence_after();
                    }
                    if (o_uses[t] > 0) {
                        mbar_wait(o_empty(t), (o_uses[t] - 1) & 1u);      // epilogue of the previous use has read O[t]
                        tc_fence_after();
                    }
#pragma unroll
                    for (int k = 0; k < WN_ROWS / 16; ++k)
                        umma_bf16_ts(tmem_base + WCOL_O + t * 80, tmem_base + WCOL_S + 8u * k, v_desc + 128u * k, idescO, k > 0);
                    umma_commit(o_full(t));
                    ++o_uses[t];
                    if (t == w.ntiles - 1) umma_commit(v_empty);
                }
            }
        }
    } else if (warp >= 4) {
        // ============================================================ softmax + epilogue: thread = query row
        const int r = threadIdx.x - 128;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_t = tmem_base + lane_addr + WCOL_T;
        const uint32_t t_s = tmem_base + lane_addr + WCOL_S;
        uint32_t tile_count = 0;
        uint32_t o_uses[2] = {0, 0};
        for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
            const WinItem w = decode(it);
            for (int t = 0; t < w.ntiles; ++t, ++tile_count) {
                const int l_q = t * 128 + r;
                const bool q_in_window = l_q < WN_L;
                const int qy = q_in_window ? l_q / WN_WS : 0, qx = q_in_window ? l_q - qy * WN_WS : 0;
                mbar_wait(s_full, tile_count & 1u);
                tc_fence_after();
                // ---- this row of T = q . [rel_pos_h ; rel_pos_w]  (unscaled q, image_encoder.py:618-619)
                float trow[64];
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld_32x32(t_t + c * 32, v);
                    tmem_ld_wait(v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) trow[c * 32 + i] = __uint_as_float(v[i]) * kLog2eW;
                }
                float th[WN_WS], tw[WN_WS];
                float thmax = -INFINITY, twmax = -INFINITY;
#pragma unroll
                for (int k = 0; k < WN_WS; ++k) {
                    th[k] = trow[qy + 13 - k];            // rel_h[q, ky] = T_h[q][qy - ky + 13]
                    tw[k] = trow[32 + qx + 13 - k];       // rel_w[q, kx] = T_w[q][qx - kx + 13]
                    thmax = fmaxf(thmax, th[k]);
                    twmax = fmaxf(twmax, tw[k]);
                }
                // ---- pass A: raw row maximum over the 196 real columns
                float smax = -INFINITY;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    tmem_ld_32x32(t_s + c * 32, v);
                    tmem_ld_wait(v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) smax = fmaxf(smax, __uint_as_float(v[i]));
                }
                {
                    uint32_t v16[16];
                    tmem_ld_32x16(t_s + 192, v16);
                    tmem_ld_wait16(v16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) smax = fmaxf(smax, __uint_as_float(v16[i]));
                }
                // reference point >= every logit of the row (bounds the exponentials by 1)
                const float m = fmaf(smax, scale_log2, thmax + twmax);
#pragma unroll
                for (int k = 0; k < WN_WS; ++k) th[k] -= m;
                // ---- pass B: P = exp2(scale*s + rel_h + rel_w - m) as bf16, written over S in place
                float l = 0.f;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    tmem_ld_32x32(t_s + c * 32, v);
                    tmem_ld_wait(v);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int k0 = c * 32 + 2 * i, k1 = k0 + 1;            // compile-time key indices
                        const float p0 = ex2w(fmaf(__uint_as_float(v[2 * i]), scale_log2, th[k0 / WN_WS] + tw[k0 % WN_WS]));
                        const float p1 = ex2w(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, th[k1 / WN_WS] + tw[k1 % WN_WS]));
                        l += p0 + p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x16(t_s + c * 16, pk);
                }
                {
                    uint32_t v16[16];
                    tmem_ld_32x16(t_s + 192, v16);
                    tmem_ld_wait16(v16);
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k0 = 192 + 2 * i, k1 = k0 + 1;
                        float p0 = 0.f, p1 = 0.f;                               // keys 196..207 do not exist
                        if (k0 < WN_L) p0 = ex2w(fmaf(__uint_as_float(v16[2 * i]), scale_log2, th[(k0 < WN_L ? k0 : 0) / WN_WS] + tw[k0 % WN_WS]));
                        if (k1 < WN_L) p1 = ex2w(fmaf(__uint_as_float(v16[2 * i + 1]), scale_log2, th[(k1 < WN_L ? k1 : 0) / WN_WS] + tw[k1 % WN_WS]));
                        l += p0 + p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x8(t_s + 96, pk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_full);
                // ---- epilogue of this tile: O[t] / l -> bf16 -> global, queries inside the grid only
                mbar_wait(o_full(t), o_uses[t] & 1u);
                ++o_uses[t];
                tc_fence_after();
                const int gy = w.wy * WN_WS + qy, gx = w.wx * WN_WS + qx;
                const bool store = q_in_window && gy < G && gx < G;
                const float inv = 1.0f / l;
                __nv_bfloat16* orow = out + ((static_cast<long long>(w.b) * G + gy) * G + gx) * D + w.head * HD;
                const uint32_t t_o = tmem_base + lane_addr + WCOL_O + t * 80;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld_32x32(t_o + c * 32, v);
                    tmem_ld_wait(v);
                    if (store) {
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
                }
                if (HI) {
                    uint32_t w16[16];
                    tmem_ld_32x16(t_o + 64, w16);
                    tmem_ld_wait16(w16);
                    if (store) {
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
                mbar_arrive(o_empty(t));
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
static int launch_window_tc(const void* qkv, void* out, const void* relh, const void* relw, int batch, int G, int heads,
                            cudaStream_t stream) {
    const int D = heads * HD;
    const int nW = (G + WN_WS - 1) / WN_WS;
    CUtensorMap tq, th, tw;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 5;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = G; s.dims[3] = G; s.dims[4] = batch;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2; s.strides[2] = 3ull * D * 2 * G;
    s.strides[3] = 3ull * D * 2 * G * G;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = WN_WS; s.box[3] = WN_WS; s.box[4] = 1;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    MapSpec r = map_spec_zero();
    r.rank = 2;
    r.dims[0] = HD; r.dims[1] = 2 * WN_WS - 1;
    r.strides[0] = HD * 2ull;
    r.box[0] = 64; r.box[1] = 32;
    r.ptr = relh;
    CVLM_TRY(make_tensor_map_bf16(&th, r));
    r.ptr = relw;
    CVLM_TRY(make_tensor_map_bf16(&tw, r));
    auto kern = attn_window_tc_kernel<HD>;
    CVLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WN_SMEM));
    const int items = batch * nW * nW * heads;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = items < sms ? items : sms;
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * kLog2eW;
    kern<<<grid, 256, WN_SMEM, stream>>>(tq, th, tw, static_cast<__nv_bfloat16*>(out), G, nW, heads, D, items, scale_log2);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// 14x14 window attention, head_dim 64 or 80, any token grid.
int attention_window_tc_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int grid_side,
                               int heads, int head_dim, cudaStream_t stream) {
    if (head_dim == 80) return launch_window_tc<80>(qkv, out, relh, relw, batch, grid_side, heads, stream);
    if (head_dim == 64) return launch_window_tc<64>(qkv, out, relh, relw, batch, grid_side, heads, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_window_tc: head_dim %d", head_dim);
}

}  // namespace cvlm
