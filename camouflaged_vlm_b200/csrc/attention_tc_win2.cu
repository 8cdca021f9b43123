// 14x14 window attention, second generation: one softmax warpgroup per query tile, the rel-pos tables folded
// into the key tile, single-pass softmax.
//
// The reference pads the 64x64 token grid to 70x70 AFTER norm1, cuts it into 25 windows of 196 tokens
// (window_partition, image_encoder.py:507-528), runs attention per window and crops (window_unpartition
// :531-553).  Here a window is one 5-D TMA box {64 d, 1 head-slot, 14 x, 14 y, 1 image} over the
// un-partitioned token-major qkv buffer: the 196 rows land in shared memory in window order and every
// token outside the grid arrives as ZEROS (TMA out-of-bounds fill) -- exactly the reference's padding,
// because K and V carry no bias (packing.py).  Nothing is materialised, nothing is copied back.
//
// Persistent CTAs loop over (image, window, head) items; an item is two query tiles of 128 rows (tile 1
// holds queries 128..195; it is skipped for bottom-edge windows where those are all padding).
//   warp 0       TMA producer (Q, K and V single slots, reloaded as soon as the MMAs that read them retire; the
//                next item's S' cannot start before this item's epilogue anyway, so the loads are hidden).
//   warp 1       MMA issuer.  The key tile K' has 256 rows: 196 keys, then rel_pos_h (27 rows) and rel_pos_w
//                (27 rows) written once per CTA, so ONE tcgen05.mma chain  S' = Q K'^T  (M128 N256, K = 64 + 16)
//                yields the logits (columns 0..195) and T_h / T_w = q . rel_pos (columns 196..249, unscaled q,
//                image_encoder.py:618-619).  O = P V is a TS chain (P from TMEM, V MN-major, K = 208).
//   warp 2       TMEM allocator: 512 columns = one 256-column S' region per query tile; bf16 P is written over
//                the consumed S columns [0,104) and O lands in [112,192) of the same region.
//   warps 4-7    softmax + epilogue of query tile 0, warps 8-11 of tile 1 (thread = query row).  A row needs
//                th[ky] = T_h[qy - ky + 13] and tw[kx] = T_w[qx - kx + 13]: a per-thread dynamic index.  Each
//                thread parks its 54 T values in a shared-memory scratch laid out [column][lane] (bank = lane,
//                conflict free for any index) and reads the 28 it needs back -- registers cannot be indexed
//                and local memory serialises 32 ways on divergent indices.
// Softmax is single pass with a per-row reference that starts at the row's first logit and is only raised
// when a 32-column chunk sums above 2^30 (see attention_tc2.cu); arithmetic is packed (FFMA2 / FADD2).
#include "cvlm_internal.cuh"
#include "ptx.cuh"
#include "tensor_map.cuh"

namespace cvlm {

constexpr int W2_WS = 14, W2_L = 196, W2_ROWS = 208, W2_KROWS = 256;
constexpr int W2_TILE = W2_ROWS * 128;               // [208 rows x 64 bf16], 26 KB
constexpr int W2_SLOT = 2 * W2_TILE;                 // lo + hi
constexpr int W2_KTILE = W2_KROWS * 128;             // [256 rows x 64 bf16], 32 KB
constexpr int W2_OFF_V = 0;
constexpr int W2_OFF_Q = W2_SLOT;                    // query tile 1 reads 48 rows past the slot's 208 (don't-care rows),
constexpr int W2_OFF_K = W2_OFF_Q + W2_SLOT;         // so K' (lo + hi, 64 KB) follows the Q slot
constexpr int W2_OFF_STG = W2_OFF_K + 2 * W2_KTILE;  // T scratch: 7 active softmax warps x [54 columns][32 lanes] fp32
constexpr int W2_STG_WARP = 54 * 128;
constexpr int W2_OFF_BAR = W2_OFF_STG + 7 * W2_STG_WARP;
constexpr int W2_SMEM = W2_OFF_BAR + 256 + 1024;
static_assert(W2_SMEM <= 232448, "attention_tc_win2 shared memory");
constexpr int W2_BOX_BYTES = 64 * W2_L * 2;          // bytes one TMA box delivers
constexpr int W2_THREADS = 384;
constexpr int W2_REL_H = 196, W2_REL_W = 223;        // K' rows / S' columns of the rel-pos tables
constexpr uint32_t W2_COL_O = 112;
constexpr float kLog2eW2 = 1.4426950408889634f;
constexpr float W2_CHUNK_LIMIT = 1073741824.0f;      // 2^30
// Bottleneck probes (tools/attn_probe.sh builds side libraries with -DW2_PROBE=n; the shipped library has 0):
// bit 0 = no exponentials (P = masked S bits), bit 1 = Q / K / V are loaded for the first item only, bit 2 = the
// d 64..79 boxes are not loaded (half the TMA box rows), bit 3 = no global stores.  Results are garbage by design.
#ifndef W2_PROBE
#define W2_PROBE 0
#endif

struct Win2Item {
    int b, wy, wx, head, ntiles;
};

// Columns [32*C, 32*C + NP*2) of a row: p = exp2(s * scale + tw[kx] + (th[ky] - m)); sum returned, bf16 pairs packed.
// Key k -> (ky, kx) = (k / 14, k % 14) are compile-time; pairs never straddle a window row (14 is even).
template <int C, int NP>
__device__ __forceinline__ float w2_chunk_exp(const uint32_t* v, const uint64_t (&tw2)[7], const uint64_t (&thm2)[14],
                                              uint64_t scale2, uint32_t* pk) {
    uint64_t acc0 = 0ull, acc1 = 0ull;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int k0 = C * 32 + 2 * i;
        if (W2_PROBE & 1) {
            pk[i] = v[2 * i] & 0x3f803f80u;
        } else if (k0 < W2_L) {
            const uint64_t a = f32x2_add(tw2[(k0 % W2_WS) / 2], thm2[k0 / W2_WS]);
            const uint64_t x = f32x2_fma(f32x2_pack_bits(v[2 * i], v[2 * i + 1]), scale2, a);
            float x0, x1;
            f32x2_unpack(x, x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            if (i & 1) acc1 = f32x2_add(acc1, f32x2_pack(p0, p1));
            else acc0 = f32x2_add(acc0, f32x2_pack(p0, p1));
            pk[i] = pack_bf16(p0, p1);
        } else {
            pk[i] = 0u;                               // keys 196..207 do not exist
        }
    }
    float s0, s1;
    f32x2_unpack(f32x2_add(acc0, acc1), s0, s1);
    return s0 + s1;
}

template <int C, int NP>
__device__ __forceinline__ float w2_chunk_max(const uint32_t* v, const float (&tw)[14], const float (&th)[14], float scale_log2) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
        const int k = C * 32 + i;
        if (k < W2_L) mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), scale_log2, tw[k % W2_WS] + th[k / W2_WS]));
    }
    return mx;
}

__device__ __forceinline__ void w2_scale_p(uint32_t addr, float f) {
    uint32_t w[16];
    tmem_ld_32x16(addr, w);
    tmem_ld_wait16(w);
#pragma unroll
    for (int i = 0; i < 16; ++i)
        w[i] = pack_bf16(__uint_as_float(w[i] << 16) * f, __uint_as_float(w[i] & 0xffff0000u) * f);
    tmem_st_32x16(addr, w);
}

// One chunk with the range guard.  `valid` masks rows that are not queries of the window (garbage in, ignored).
template <int C, int NP>
__device__ __forceinline__ void w2_chunk(const uint32_t* v, const float (&tw)[14], const float (&th)[14], const uint64_t (&tw2)[7],
                                         uint64_t (&thm2)[14], float& m_used, float& l, uint64_t scale2, float scale_log2,
                                         uint32_t t_s, bool valid) {
    uint32_t pk[16];
    float cs = w2_chunk_exp<C, NP>(v, tw2, thm2, scale2, pk);
    if (__any_sync(0xffffffffu, valid && !(cs < W2_CHUNK_LIMIT))) {
        const float m_new = fmaxf(m_used, w2_chunk_max<C, NP>(v, tw, th, scale_log2));
        const float f = ex2_approx(m_used - m_new);
        l *= f;
#pragma unroll
        for (int cp = 0; cp < C; ++cp) w2_scale_p(t_s + cp * 16, f);
        m_used = m_new;
#pragma unroll
        for (int k = 0; k < W2_WS; ++k) thm2[k] = f32x2_pack(th[k] - m_used, th[k] - m_used);
        cs = w2_chunk_exp<C, NP>(v, tw2, thm2, scale2, pk);
    }
    l += cs;
    if (NP == 16) {
        tmem_st_32x16(t_s + C * 16, *reinterpret_cast<uint32_t(*)[16]>(pk));
    } else {
        tmem_st_32x8(t_s + C * 16, *reinterpret_cast<uint32_t(*)[8]>(pk));
    }
}

template <int HD>
__global__ void __launch_bounds__(W2_THREADS, 1)
attn_window_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __nv_bfloat16* __restrict__ relh,
                       const __nv_bfloat16* __restrict__ relw, __nv_bfloat16* __restrict__ out, int G, int nW, int heads,
                       int D, int num_items, float scale_log2, float* __restrict__ lse_out) {
    constexpr bool HI = HD > 64;
    constexpr bool LOAD_HI = HI && !(W2_PROBE & 4);
    constexpr uint32_t LOAD_BYTES = (LOAD_HI ? 2 : 1) * W2_BOX_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sK = base + W2_OFF_K, sV = base + W2_OFF_V;
    const uint32_t sQ = base + W2_OFF_Q;
    const uint32_t bars = base + W2_OFF_BAR;
    const uint32_t q_full = bars, q_empty = bars + 8u * 2;
    const uint32_t k_full = bars + 8u * 4, k_empty = bars + 8u * 5, v_full = bars + 8u * 6, v_empty = bars + 8u * 7;
    auto s_full = [&](int t) { return bars + 8u * (8 + t); };
    auto p_full = [&](int t) { return bars + 8u * (10 + t); };
    auto o_full = [&](int t) { return bars + 8u * (12 + t); };
    auto o_empty = [&](int t) { return bars + 8u * (14 + t); };
    const uint32_t tmem_slot = bars + 8u * 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + W2_OFF_BAR + 8 * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto decode = [&](int it) {
        Win2Item w;
        w.head = it % heads;
        const int rest = it / heads;
        const int win = rest % (nW * nW);
        w.b = rest / (nW * nW);
        w.wy = win / nW;
        w.wx = win % nW;
        // the second query tile starts at window row 9 (token 128 = 9*14 + 2): all padding for bottom-edge windows
        w.ntiles = (w.wy * W2_WS + 9 < G) ? 2 : 1;
        return w;
    };

    // ---- one-time shared-memory setup (generic proxy writes, fenced for the async proxy below)
    // K' rows 196..255: rel_pos_h (27 rows), rel_pos_w (27 rows), zeros; 128-byte swizzle by hand:
    // 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4).
    for (int i = threadIdx.x; i < 2 * 60 * 8; i += blockDim.x) {
        const int tile = i / 480, rest = i % 480;
        const int row = W2_L + rest / 8, chunk = rest % 8;
        const int d0 = tile * 64 + chunk * 8;                 // first head-dim element of this 16-byte chunk
        uint4 val = make_uint4(0, 0, 0, 0);
        if (d0 < HD) {
            if (row < W2_REL_W) val = *reinterpret_cast<const uint4*>(relh + (row - W2_REL_H) * HD + d0);
            else if (row < W2_REL_W + 27) val = *reinterpret_cast<const uint4*>(relw + (row - W2_REL_W) * HD + d0);
        }
        *reinterpret_cast<uint4*>(base_ptr + W2_OFF_K + tile * W2_KTILE + row * 128 + ((chunk ^ (row & 7)) << 4)) = val;
    }
    // V rows 196..207 are read by the last k-step of P V (against P == 0) but never written by TMA: zero them once
    for (int i = threadIdx.x; i < 2 * 12 * 8; i += blockDim.x) {
        const int tile = i / 96, rest = i % 96;
        *reinterpret_cast<uint4*>(base_ptr + W2_OFF_V + tile * W2_TILE + W2_L * 128 + rest * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(s_full(s), 1);
            mbar_init(p_full(s), 128);
            mbar_init(o_full(s), 1);
            mbar_init(o_empty(s), 128);
        }
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
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

    if (warp < 4) {
        setmaxnreg_dec<64>();
        if (warp == 0 && elect_one_sync()) {
            // ============================================================ TMA producer
            int n = 0;
            for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
                const Win2Item w = decode(it);
                const uint32_t ph = static_cast<uint32_t>(n & 1);          // one slot each: the phase flips per item
                const int x0 = w.wx * W2_WS, y0 = w.wy * W2_WS;
                if ((W2_PROBE & 2) && n > 0) break;
                mbar_wait_relaxed(q_empty, ph ^ 1u);
                mbar_arrive_expect_tx(q_full, LOAD_BYTES);
                tma_load_5d(sQ, &tmQKV, q_full, 0, w.head, x0, y0, w.b);
                if (LOAD_HI) tma_load_5d(sQ + W2_TILE, &tmQKV, q_full, 64, w.head, x0, y0, w.b);
                mbar_wait_relaxed(k_empty, ph ^ 1u);
                mbar_arrive_expect_tx(k_full, LOAD_BYTES);
                tma_load_5d(sK, &tmQKV, k_full, 0, heads + w.head, x0, y0, w.b);
                if (LOAD_HI) tma_load_5d(sK + W2_KTILE, &tmQKV, k_full, 64, heads + w.head, x0, y0, w.b);
                mbar_wait_relaxed(v_empty, ph ^ 1u);
                mbar_arrive_expect_tx(v_full, LOAD_BYTES);
                tma_load_5d(sV, &tmQKV, v_full, 0, 2 * heads + w.head, x0, y0, w.b);
                if (LOAD_HI) tma_load_5d(sV + W2_TILE, &tmQKV, v_full, 64, 2 * heads + w.head, x0, y0, w.b);
            }
        } else if (warp == 1 && elect_one_sync()) {
            // ============================================================ MMA issuer
            constexpr uint32_t idescS = umma_idesc_bf16(128, W2_KROWS);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            const uint64_t k_lo = umma_desc_k_sw128(sK), k_hi = umma_desc_k_sw128(sK + W2_KTILE);
            const uint64_t v_desc = umma_desc_mn_sw128(sV, W2_TILE);
            int n = 0;
            uint32_t uses[2] = {0, 0};               // how many times tile t has been processed
            for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
                const Win2Item w = decode(it);
                if (!(W2_PROBE & 2) || n == 0) {
                    mbar_wait(q_full, static_cast<uint32_t>(n & 1));
                    mbar_wait(k_full, static_cast<uint32_t>(n & 1));
                }
                tc_fence_after();
                for (int t = 0; t < w.ntiles; ++t) {
                    if (uses[t] > 0) {
                        mbar_wait(o_empty(t), (uses[t] - 1) & 1u);     // epilogue of the previous use has read O_t
                        tc_fence_after();
                    }
                    const uint32_t d = tmem_base + t * 256;
                    const uint64_t q_lo = umma_desc_k_sw128(sQ + t * 128 * 128);
                    const uint64_t q_hi = umma_desc_k_sw128(sQ + W2_TILE + t * 128 * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(d, q_lo + 2u * k, k_lo + 2u * k, idescS, k > 0);
                    if (HI) umma_bf16_ss(d, q_hi, k_hi, idescS, 1);
                    umma_commit(s_full(t));
                }
                umma_commit(q_empty);
                umma_commit(k_empty);
                for (int t = 0; t < w.ntiles; ++t) {
                    mbar_wait(p_full(t), uses[t] & 1u);
                    tc_fence_after();
                    if (t == 0 && (!(W2_PROBE & 2) || n == 0)) {
                        mbar_wait(v_full, static_cast<uint32_t>(n & 1));
                        tc_fence_after();
                    }
                    const uint32_t a = tmem_base + t * 256;
#pragma unroll
                    for (int k = 0; k < W2_ROWS / 16; ++k)
                        umma_bf16_ts(a + W2_COL_O, a + 8u * k, v_desc + 128u * k, idescO, k > 0);
                    umma_commit(o_full(t));
                    ++uses[t];
                }
                umma_commit(v_empty);
            }
        }
    } else {
        // ============================================================ softmax + epilogue: thread = query row
        setmaxnreg_inc<216>();
        const int t = warp >= 8 ? 1 : 0;
        const int r = threadIdx.x - 128 - t * 128;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_base + lane_addr + t * 256;
        const int l_q = t * 128 + r;
        const bool q_in_window = l_q < W2_L;
        const bool warp_active = t * 128 + (warp & 3) * 32 < W2_L;        // any query of the window in this warp
        const int qy = q_in_window ? l_q / W2_WS : 0, qx = q_in_window ? l_q - qy * W2_WS : 0;
        const uint64_t scale2 = f32x2_pack(scale_log2, scale_log2);
        // this thread's column of the T scratch (warps of tile 1 use slots 4..6; the all-padding warp has none)
        const uint32_t stg = base + W2_OFF_STG + (t * 4 + (warp & 3)) * W2_STG_WARP + lane * 4;
        uint32_t use = 0;
        for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
            const Win2Item w = decode(it);
            if (t >= w.ntiles) continue;
            mbar_wait(s_full(t), use & 1u);
            tc_fence_after();
            float l = 1.f, m_row = 0.f;
            if (warp_active) {
                // ---- this row of T = q . [rel_pos_h ; rel_pos_w]: S' columns 192..255 hold keys 192..195, T_h, T_w
                uint32_t va[32], vb[32];
                tmem_ld_32x32(t_s + 192, va);
                tmem_ld_32x32(t_s + 224, vb);
                tmem_ld_wait(va);
                tmem_ld_wait(vb);
                const uint32_t tail[4] = {va[0], va[1], va[2], va[3]};     // logits of keys 192..195
#pragma unroll
                for (int c = 0; c < 54; ++c)                               // T column c = S' column 196 + c
                    st_shared_b32(stg + c * 128, c < 28 ? va[4 + c] : vb[c - 28]);
                tmem_ld_32x32(t_s, va);                                    // chunk 0 in flight
                float th[W2_WS], tw[W2_WS];
#pragma unroll
                for (int k = 0; k < W2_WS; ++k) {
                    th[k] = ld_shared_f32(stg + (qy + 13 - k) * 128) * kLog2eW2;        // rel_h[q, ky] = T_h[q][qy - ky + 13]
                    tw[k] = ld_shared_f32(stg + (27 + qx + 13 - k) * 128) * kLog2eW2;   // rel_w[q, kx] = T_w[q][qx - kx + 13]
                }
                uint64_t tw2[7], thm2[W2_WS];
#pragma unroll
                for (int k = 0; k < 7; ++k) tw2[k] = f32x2_pack(tw[2 * k], tw[2 * k + 1]);
                tmem_ld_wait(va);
                float m_used = fmaf(__uint_as_float(va[0]), scale_log2, th[0] + tw[0]);   // the row's first logit
#pragma unroll
                for (int k = 0; k < W2_WS; ++k) thm2[k] = f32x2_pack(th[k] - m_used, th[k] - m_used);
                l = 0.f;
                tmem_ld_32x32(t_s + 32, vb);
                w2_chunk<0, 16>(va, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                tmem_ld_wait(vb);
                tmem_ld_32x32(t_s + 64, va);
                w2_chunk<1, 16>(vb, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                tmem_ld_wait(va);
                tmem_ld_32x32(t_s + 96, vb);
                w2_chunk<2, 16>(va, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                tmem_ld_wait(vb);
                tmem_ld_32x32(t_s + 128, va);
                w2_chunk<3, 16>(vb, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                tmem_ld_wait(va);
                tmem_ld_32x32(t_s + 160, vb);
                w2_chunk<4, 16>(va, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                tmem_ld_wait(vb);
                w2_chunk<5, 16>(vb, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                // keys 192..195 (+ 12 zero columns up to 208): their logits were saved before P overwrote anything
                uint32_t tail16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) tail16[i] = i < 4 ? tail[i] : 0u;
                w2_chunk<6, 8>(tail16, tw, th, tw2, thm2, m_used, l, scale2, scale_log2, t_s, q_in_window);
                m_row = m_used;
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(p_full(t));
            // ---- epilogue of this tile: O_t / l -> bf16 -> global, queries inside the grid only
            mbar_wait(o_full(t), use & 1u);
            ++use;
            tc_fence_after();
            if (warp_active) {
                const int gy = w.wy * W2_WS + qy, gx = w.wx * W2_WS + qx;
                const bool store = q_in_window && gy < G && gx < G;
                const float inv = 1.0f / l;
                __nv_bfloat16* orow = out + ((static_cast<long long>(w.b) * G + gy) * G + gx) * D + w.head * HD;
                const uint32_t t_o = t_s + W2_COL_O;
                uint32_t o0[32], o1[32], o2[16];
                tmem_ld_32x32(t_o, o0);
                tmem_ld_32x32(t_o + 32, o1);
                if (HI) tmem_ld_32x16(t_o + 64, o2);
                tmem_ld_wait(o0);
                tmem_ld_wait(o1);
                if (HI) tmem_ld_wait16(o2);
                // O is in registers: give the TMEM region back before the (slow, scattered) global stores -- the next
                // item's S' cannot start until this arrives (timeline: the stores took ~1800 of ~9100 cycles per item)
                tc_fence_before();
                mbar_arrive(o_empty(t));
                // training forward: the row's log-sum-exp (log2 domain, rel-pos terms included) for the backward kernels
                if (store && lse_out != nullptr)
                    lse_out[((static_cast<long long>(w.b) * G + gy) * G + gx) * heads + w.head] = m_row + log2f(l);
                if (store && !(W2_PROBE & 8)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 16) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            pk[e] = pack_bf16(__uint_as_float(o0[i + 2 * e]) * inv, __uint_as_float(o0[i + 2 * e + 1]) * inv);
                        st_global_v8(orow + i, pk);
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 16) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            pk[e] = pack_bf16(__uint_as_float(o1[i + 2 * e]) * inv, __uint_as_float(o1[i + 2 * e + 1]) * inv);
                        st_global_v8(orow + 32 + i, pk);
                    }
                    if (HI) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            pk[e] = pack_bf16(__uint_as_float(o2[2 * e]) * inv, __uint_as_float(o2[2 * e + 1]) * inv);
                        st_global_v8(orow + 64, pk);
                    }
                }
            } else {
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
static int launch_window_tc2(const void* qkv, void* out, const void* relh, const void* relw, int batch, int G, int heads,
                             cudaStream_t stream, float* lse_out) {
    const int D = heads * HD;
    const int nW = (G + W2_WS - 1) / W2_WS;
    CUtensorMap tq;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 5;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = G; s.dims[3] = G; s.dims[4] = batch;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2; s.strides[2] = 3ull * D * 2 * G;
    s.strides[3] = 3ull * D * 2 * G * G;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = W2_WS; s.box[3] = W2_WS; s.box[4] = 1;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    if ((reinterpret_cast<uintptr_t>(relh) & 15) || (reinterpret_cast<uintptr_t>(relw) & 15))
        return set_error(CVLM_ERR_ARG, "attention: rel-pos tables must be 16-byte aligned");
    auto kern = attn_window_tc2_kernel<HD>;
    static unsigned char ready[kMaxDevices] = {};         // per (instantiation, device)
    {
        int dev = 0;
        CVLM_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices || !__atomic_load_n(&ready[dev], __ATOMIC_ACQUIRE)) {
            // setmaxnreg: 4 control warps shrink to 64, 8 softmax warps grow to 216 out of what those released
            cudaFuncAttributes fa;
            CVLM_CUDA(cudaFuncGetAttributes(&fa, kern));
            if ((fa.numRegs - 64) * 128 < (216 - fa.numRegs) * 256)
                return set_error(CVLM_ERR_UNSUPPORTED, "attention_tc_win2: built with %d registers per thread; the setmaxnreg pool would be short", fa.numRegs);
        }
    }
    CVLM_TRY(ensure_dynamic_smem(kern, W2_SMEM, ready));
    const int items = batch * nW * nW * heads;
    const int sms = device_sm_count();
    const int grid = items < sms ? items : sms;
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * kLog2eW2;
    kern<<<grid, W2_THREADS, W2_SMEM, stream>>>(tq, static_cast<const __nv_bfloat16*>(relh), static_cast<const __nv_bfloat16*>(relw),
                                                static_cast<__nv_bfloat16*>(out), G, nW, heads, D, items, scale_log2, lse_out);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// 14x14 window attention, head_dim 64 or 80, any token grid.
int attention_window_tc2_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int grid_side,
                                int heads, int head_dim, cudaStream_t stream, float* lse_out) {
    if (head_dim == 80) return launch_window_tc2<80>(qkv, out, relh, relw, batch, grid_side, heads, stream, lse_out);
    if (head_dim == 64) return launch_window_tc2<64>(qkv, out, relh, relw, batch, grid_side, heads, stream, lse_out);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_window_tc2: head_dim %d", head_dim);
}

}  // namespace cvlm
