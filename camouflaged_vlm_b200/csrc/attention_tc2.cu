// Global attention (token grid 64x64, 4096 keys), second generation.
//
// Contract as in attention.cu: bias-free K/V in the token-major qkv buffer, decomposed rel-pos from the
// unscaled q added inside the softmax tile (image_encoder.py:493-501, 589-625).  One CTA owns 256 queries
// (two query tiles of 128 = four grid rows) of one (image, head) and streams 64 key units of 64 keys (one
// grid row each), so the rel-pos h-term is one scalar per (query, unit) and the w-term is the same 64 values
// for every unit (kept in registers).
//
//   warp 0       TMA producer: both Q tiles once; the rel-pos tables (4 units: rel_pos_w rows 0..127, a
//                128-row slice of rel_pos_h) and the K units go through a 4-stage ring, V units through another
//                (cp.async.bulk.tensor, SWIZZLE_128B; head_dim 80 = a 64-wide tile + a 16-wide tile whose
//                columns 80..127 are out of bounds and arrive as zeros).
//   warp 1       MMA issuer (one lane).  S_t(u) = Q_t * unit_u^T  (tcgen05.mma SS, M128 N64, K = 64 + 16) goes
//                into one of TWO S buffers per query tile, so it is issued a full unit ahead of the softmax
//                that consumes it;  O_t += P_t(u) V_u  (tcgen05.mma TS: P read from TMEM, V MN-major from smem,
//                N = hd) follows as soon as the softmax hands P over.  The tensor pipe executes in issue order,
//                so S_t(u+2) may overwrite the columns P_t(u) lived in.
//   warp 2       TMEM allocator (512 columns: S[t][b] at (2t+b)*64, O_t at 256 + 80t).
//   warps 4-7    softmax of query tile 0, warps 8-11 of tile 1; one thread per query row (TMEM lane = row).
//
// Softmax is single pass and "optimistic": P = exp2(scale*s + rel_h + rel_w - m) with a per-row reference m
// that only has to keep the exponentials inside the fp32/bf16 range (softmax is shift invariant; P, l and O
// carry no absolute scale).  m starts at the row's first logit and is raised -- rescaling l and O -- only
// when a unit's 64 exponentials sum above 2^30, which also catches inf.  The unit's S row is read from
// tensor memory once, into registers, and bf16 P is written over it.  Arithmetic is packed (FFMA2 / FADD2);
// POLY of every 16 exponential pairs are evaluated on the FMA pipe (Cody-Waite split + cubic, 7.5e-5 relative
// error, far below the bf16 rounding of P) instead of the MUFU, which is the busiest unit of this kernel.
#include <stdlib.h>

#include "cvlm_internal.cuh"
#include "ptx.cuh"
#include "tensor_map.cuh"

namespace cvlm {

constexpr int G2_UNIT = 64;                        // keys per unit
constexpr int G2_QTILE = 128 * 128;                // one [128 rows x 64 bf16] swizzled tile, bytes
constexpr int G2_QSTAGE = 2 * G2_QTILE;            // lo (d 0..63) + hi (d 64..127, only 64..79 real)
constexpr int G2_KTILE = G2_UNIT * 128;            // one [64 rows x 64 bf16] tile, 8 KB
constexpr int G2_KSTAGE = 2 * G2_KTILE;
constexpr int G2_KST = 4, G2_VST = 4;
constexpr int G2_OFF_Q = 0;                        // two query tiles
constexpr int G2_OFF_K = 2 * G2_QSTAGE;
constexpr int G2_OFF_V = G2_OFF_K + G2_KST * G2_KSTAGE;
constexpr int G2_OFF_BAR = G2_OFF_V + G2_VST * G2_KSTAGE;
constexpr int G2_SMEM = G2_OFF_BAR + 512 + 1024;
static_assert(G2_SMEM <= 232448, "attention_tc2 shared memory");
constexpr int G2_THREADS = 384;
constexpr uint32_t G2_COL_O = 256, G2_O_STRIDE = 80;
constexpr float kLog2eG2 = 1.4426950408889634f;
constexpr float G2_UNIT_LIMIT = 1073741824.0f;     // 2^30
// Bottleneck probes (tools/attn_probe.sh builds side libraries with -DG2_PROBE=n; the shipped library has 0):
// bit 0 = the softmax warps skip the exponentials (P = raw S bits), bit 1 = K / V units are loaded for the first
// ring fill only and the MMAs re-read those stages (no steady-state TMA traffic), bit 2 = the 16-wide d 64..79 boxes of
// K and V are not loaded (half the TMA box rows).  Results are garbage by design.
#ifndef G2_PROBE
#define G2_PROBE 0
#endif

// 2^x for a pair on the FMA pipe: n = round(x), r = x - n in [-0.5, 0.5], cubic in r, exponent added as an integer.
__device__ __forceinline__ void g2_exp2_poly(uint64_t x2, float& p0, float& p1) {
    float x0, x1;
    f32x2_unpack(x2, x0, x1);
    x0 = fminf(fmaxf(x0, -125.0f), 126.0f);     // above: 2^126 trips the caller's range guard like the MUFU's inf does
    x1 = fminf(fmaxf(x1, -125.0f), 126.0f);
    const uint64_t magic = f32x2_pack(12582912.0f, 12582912.0f);          // 1.5 * 2^23
    const uint64_t nmagic = f32x2_pack(-12582912.0f, -12582912.0f);
    const uint64_t xc = f32x2_pack(x0, x1);
    const uint64_t t = f32x2_add(xc, magic);                              // low mantissa bits = round(x)
    const uint64_t nf = f32x2_add(t, nmagic);
    const uint64_t r = f32x2_fma(nf, f32x2_pack(-1.0f, -1.0f), xc);
    uint64_t p = f32x2_fma(f32x2_pack(0.0551716685f, 0.0551716685f), r, f32x2_pack(0.2426111251f, 0.2426111251f));
    p = f32x2_fma(p, r, f32x2_pack(0.6932609677f, 0.6932609677f));
    p = f32x2_fma(p, r, f32x2_pack(0.9999280572f, 0.9999280572f));
    float q0, q1, t0, t1;
    f32x2_unpack(p, q0, q1);
    f32x2_unpack(t, t0, t1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// One 32-column chunk (CC = 0 / 1: columns 0..31 / 32..63 of the unit): p = exp2(s * scale + (tw + ch)).
// Accumulates the row sum into acc (two packed partial sums) and packs bf16 pairs.
template <int CC, int POLY>
__device__ __forceinline__ void g2_chunk_exp(const uint32_t (&v)[32], const uint64_t (&tw2)[32], uint64_t ch2, uint64_t scale2,
                                             uint64_t& acc0, uint64_t& acc1, uint32_t* pk) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t a = f32x2_add(tw2[CC * 16 + i], ch2);
        const uint64_t x = f32x2_fma(f32x2_pack_bits(v[2 * i], v[2 * i + 1]), scale2, a);
        float p0, p1;
        // spread the polynomial pairs evenly through the chunk so the MUFU never idles behind them
        const bool poly = POLY > 0 && ((i + 1) * POLY) / 16 != (i * POLY) / 16;
        if (poly) {
            g2_exp2_poly(x, p0, p1);
        } else {
            float x0, x1;
            f32x2_unpack(x, x0, x1);
            p0 = ex2_approx(x0);
            p1 = ex2_approx(x1);
        }
        if (i & 1) acc1 = f32x2_add(acc1, f32x2_pack(p0, p1));
        else acc0 = f32x2_add(acc0, f32x2_pack(p0, p1));
        pk[i] = pack_bf16(p0, p1);
    }
}

// row *= f for NCOLS fp32 TMEM columns (NCOLS = 64 or 80)
template <int NCOLS>
__device__ __forceinline__ void g2_scale_cols(uint32_t addr, float f) {
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(addr + c * 32, w);
        tmem_ld_wait(w);
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * f);
        tmem_st_32x32(addr + c * 32, w);
    }
    if (NCOLS > 64) {
        uint32_t w16[16];
        tmem_ld_32x16(addr + 64, w16);
        tmem_ld_wait16(w16);
#pragma unroll
        for (int i = 0; i < 16; ++i) w16[i] = __float_as_uint(__uint_as_float(w16[i]) * f);
        tmem_st_32x16(addr + 64, w16);
    }
    tmem_st_wait();
}

// PLAIN = true: the same pipeline as a plain softmax(q k^T / sqrt(hd)) v over `L` tokens per image (any L >= 1, no
// relative-position terms): key units beyond L are masked in the last unit, query rows beyond L are not stored.
template <int HD, int POLY, bool PLAIN>
__global__ void __launch_bounds__(G2_THREADS, 1)
attn_global_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const __grid_constant__ CUtensorMap tmRelH, const __grid_constant__ CUtensorMap tmRelW,
                       __nv_bfloat16* __restrict__ out, int heads, int D, float scale_log2, int L,
                       float* __restrict__ lse_out, float* __restrict__ tw_out, float* __restrict__ th_out) {
    constexpr bool HI = HD > 64;
    constexpr uint32_t Q_BYTES = (HI ? 2 : 1) * G2_QTILE;
    constexpr uint32_t KV_BYTES = (HI ? 2 : 1) * G2_KTILE;
    const int NK = PLAIN ? (L + G2_UNIT - 1) / G2_UNIT : 4096 / G2_UNIT;   // key units (64 for the 64x64 grid)
    constexpr int NREL = PLAIN ? 0 : 4;             // rel_pos_w rows 0..63 / 64..127, rel_pos_h slice rows 0..63 / 64..127
    const int NU = NK + NREL;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    auto sQ = [&](int t) { return base + G2_OFF_Q + t * G2_QSTAGE; };
    auto sK = [&](int s) { return base + G2_OFF_K + s * G2_KSTAGE; };
    auto sV = [&](int s) { return base + G2_OFF_V + s * G2_KSTAGE; };
    const uint32_t bars = base + G2_OFF_BAR;
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bars + 8u * (5 + s); };
    auto v_full = [&](int s) { return bars + 8u * (9 + s); };
    auto v_empty = [&](int s) { return bars + 8u * (13 + s); };
    auto s_full = [&](int t, int b) { return bars + 8u * (17 + 2 * t + b); };
    auto p_full = [&](int t, int b) { return bars + 8u * (21 + 2 * t + b); };
    auto pv_done = [&](int t) { return bars + 8u * (25 + t); };
    auto o_full = [&](int t) { return bars + 8u * (27 + t); };
    const uint32_t tmem_slot = bars + 8u * 29;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int y0 = 4 * qb;                          // first grid row of this CTA's queries
    const int tok0 = b * (PLAIN ? L : 4096);
    // PLAIN: the last query block of a sequence may hold fewer than 129 rows; its second tile is then all padding and is
    // skipped altogether (no S / PV MMAs, its softmax warpgroup idles): 581 tokens are 4.5 tiles, not 6
    const int ntile = (PLAIN && qb * 256 + 128 >= L) ? 1 : 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmRelH);
        tma_prefetch_desc(&tmRelW);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 4; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
            mbar_init(s_full(s >> 1, s & 1), 1);
            mbar_init(p_full(s >> 1, s & 1), 128);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(pv_done(t), 1);
            mbar_init(o_full(t), 1);
        }
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
            mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int tok = tok0 + qb * 256 + t * 128;
                tma_load_3d(sQ(t), &tmQ, q_full, 0, head, tok);
                if (HI) tma_load_3d(sQ(t) + G2_QTILE, &tmQ, q_full, 64, head, tok);
            }
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            auto produce_k = [&](int u) {
                if ((G2_PROBE & 2) && u >= NREL + G2_KST) return;
                mbar_wait_relaxed(k_empty(kst), kph ^ 1u);
                const bool skip_hi = (G2_PROBE & 4) && u >= NREL;     // probe: 64 box rows per unit instead of 128
                mbar_arrive_expect_tx(k_full(kst), skip_hi ? G2_KTILE : KV_BYTES);
                if (!PLAIN && u < 2) {
                    tma_load_2d(sK(kst), &tmRelW, k_full(kst), 0, u * 64);
                    if (HI) tma_load_2d(sK(kst) + G2_KTILE, &tmRelW, k_full(kst), 64, u * 64);
                } else if (!PLAIN && u < 4) {
                    tma_load_2d(sK(kst), &tmRelH, k_full(kst), 0, y0 + (u - 2) * 64);
                    if (HI) tma_load_2d(sK(kst) + G2_KTILE, &tmRelH, k_full(kst), 64, y0 + (u - 2) * 64);
                } else {
                    const int tok = tok0 + (u - NREL) * G2_UNIT;
                    tma_load_3d(sK(kst), &tmKV, k_full(kst), 0, heads + head, tok);
                    if (HI && !skip_hi) tma_load_3d(sK(kst) + G2_KTILE, &tmKV, k_full(kst), 64, heads + head, tok);
                }
                if (++kst == G2_KST) { kst = 0; kph ^= 1u; }
            };
            auto produce_v = [&](int j) {
                if ((G2_PROBE & 2) && j >= G2_VST) return;
                mbar_wait_relaxed(v_empty(vst), vph ^ 1u);
                mbar_arrive_expect_tx(v_full(vst), (G2_PROBE & 4) ? G2_KTILE : KV_BYTES);
                const int tok = tok0 + j * G2_UNIT;
                tma_load_3d(sV(vst), &tmKV, v_full(vst), 0, 2 * heads + head, tok);
                if (HI && !(G2_PROBE & 4)) tma_load_3d(sV(vst) + G2_KTILE, &tmKV, v_full(vst), 64, 2 * heads + head, tok);
                if (++vst == G2_VST) { vst = 0; vph ^= 1u; }
            };
            for (int u = 0; u < NREL; ++u) produce_k(u);
            for (int j = 0; j < NK; ++j) {
                produce_k(j + NREL);
                produce_v(j);
            }
        } else if (warp == 1 && elect_one_sync()) {
            // ============================================================ MMA issuer
            constexpr uint32_t idescS = umma_idesc_bf16(128, G2_UNIT);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            mbar_wait(q_full, 0);
            tc_fence_after();
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            auto issue_s = [&](int u) {             // S_t(u) for both query tiles into buffer u & 1
                if (!(G2_PROBE & 2) || u < NREL + G2_KST) mbar_wait(k_full(kst), kph);
                tc_fence_after();
                const uint64_t k_lo = umma_desc_k_sw128(sK(kst)), k_hi = umma_desc_k_sw128(sK(kst) + G2_KTILE);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t >= ntile) break;
                    const uint32_t d = tmem_base + (2 * t + (u & 1)) * 64;
                    const uint64_t q_lo = umma_desc_k_sw128(sQ(t));
                    if (G2_PROBE & 8) {         // probe: A operand from (uninitialised) tensor memory instead of shared memory
                        const uint32_t qa = tmem_base + 416 + 40 * t;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ts(d, qa + 8u * k, k_lo + 2u * k, idescS, k > 0);
                        if (HI) umma_bf16_ts(d, qa + 32u, k_hi, idescS, 1);
                        umma_commit(s_full(t, u & 1));
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(d, q_lo + 2u * k, k_lo + 2u * k, idescS, k > 0);
                    if (HI) umma_bf16_ss(d, umma_desc_k_sw128(sQ(t) + G2_QTILE), k_hi, idescS, 1);
                    umma_commit(s_full(t, u & 1));
                }
                umma_commit(k_empty(kst));
                if (++kst == G2_KST) { kst = 0; kph ^= 1u; }
            };
            issue_s(0);
            for (int u = 0; u < NU; ++u) {
                if (u + 1 < NU) issue_s(u + 1);     // runs under the softmax of unit u
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t >= ntile) break;
                    mbar_wait(p_full(t, u & 1), static_cast<uint32_t>((u >> 1) & 1));   // S_t(u) consumed / P_t(u) written
                    tc_fence_after();
                    if (u >= NREL) {
                        if (t == 0 && (!(G2_PROBE & 2) || u - NREL < G2_VST)) {
                            mbar_wait(v_full(vst), vph);
                            tc_fence_after();
                        }
                        const uint64_t v_desc = umma_desc_mn_sw128(sV(vst), G2_KTILE);
                        const uint32_t a = tmem_base + (2 * t + (u & 1)) * 64;
                        const uint32_t d = tmem_base + G2_COL_O + t * G2_O_STRIDE;
#pragma unroll
                        for (int k = 0; k < G2_UNIT / 16; ++k)  // 16 keys per MMA: 8 TMEM columns of P, 2048 B of V
                            umma_bf16_ts(d, a + 8u * k, v_desc + 128u * k, idescO, (u > NREL || k > 0) ? 1u : 0u);
                        umma_commit(pv_done(t));
                        if (u == NU - 1) umma_commit(o_full(t));
                        if (t == ntile - 1) {
                            umma_commit(v_empty(vst));
                            if (++vst == G2_VST) { vst = 0; vph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else {
        // ============================================================ softmax: one thread per query row
        setmaxnreg_inc<216>();
        const int t = warp >= 8 ? 1 : 0;
        const int r = threadIdx.x - 128 - t * 128;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int qrow = r >> 6, qx = r & 63;
        if (t < ntile) {
        const uint32_t t_s = tmem_base + lane_addr + t * 128;
        const uint32_t t_o = tmem_base + lane_addr + G2_COL_O + t * G2_O_STRIDE;
        float trow[PLAIN ? 1 : 256];                // this row of T_w (0..127) and T_h (128..255): dynamic indexing -> local memory
        uint32_t va[32], vb[32];
#pragma unroll 1
        for (int u = 0; u < NREL; ++u) {
            mbar_wait(s_full(t, u & 1), static_cast<uint32_t>((u >> 1) & 1));
            tc_fence_after();
            tmem_ld_32x32(t_s + (u & 1) * 64, va);
            tmem_ld_32x32(t_s + (u & 1) * 64 + 32, vb);
            tmem_ld_wait(va);
            tmem_ld_wait(vb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                trow[u * 64 + i] = __uint_as_float(va[i]) * kLog2eG2;
                trow[u * 64 + 32 + i] = __uint_as_float(vb[i]) * kLog2eG2;
            }
            tc_fence_before();
            mbar_arrive(p_full(t, u & 1));
        }
        // rel_w[q, kx] = T_w[q][qx - kx + 63] (image_encoder.py:576-586): the same for every key unit
        uint64_t tw2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tw2[i] = PLAIN ? 0ull : f32x2_pack(trow[PLAIN ? 0 : qx + 63 - 2 * i], trow[PLAIN ? 0 : qx + 62 - 2 * i]);
        const uint64_t scale2 = f32x2_pack(scale_log2, scale_log2);
        float m_used = 0.f, l = 0.f;

#pragma unroll 1
        for (int j = 0; j < NK; ++j) {
            const int u = j + NREL, bb = u & 1;
            // key row ky = j;  rel_h[q, ky] = T_h[q][qy - ky + 63 - y0],  qy = y0 + 2t + qrow
            const float th = PLAIN ? 0.f : trow[PLAIN ? 0 : 128 + 2 * t + qrow + 63 - j];
            mbar_wait(s_full(t, bb), static_cast<uint32_t>((u >> 1) & 1));
            tc_fence_after();
            const uint32_t t_su = t_s + bb * 64;
            tmem_ld_32x32(t_su, va);
            tmem_ld_32x32(t_su + 32, vb);
            tmem_ld_wait(va);
            tmem_ld_wait(vb);
            if (PLAIN && j == NK - 1 && L - j * G2_UNIT < G2_UNIT) {
                // keys beyond the sequence: logit -inf, so their exponential is (as good as) zero
                const int nvalid = L - j * G2_UNIT;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= nvalid) va[i] = 0xff800000u;
                    if (32 + i >= nvalid) vb[i] = 0xff800000u;
                }
            }
            if (j == 0) {
                float t0, t1;
                f32x2_unpack(tw2[0], t0, t1);
                m_used = fmaf(__uint_as_float(va[0]), scale_log2, t0 + th);   // the row's first logit
            }
            uint32_t pk[32];
            uint64_t acc0 = 0ull, acc1 = 0ull;
            if (G2_PROBE & 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    pk[i] = va[2 * i] & 0x3f803f80u;
                    pk[16 + i] = vb[2 * i] & 0x3f803f80u;
                }
            } else {
                const float ch = th - m_used;
                const uint64_t ch2 = f32x2_pack(ch, ch);
                g2_chunk_exp<0, POLY>(va, tw2, ch2, scale2, acc0, acc1, pk);
                g2_chunk_exp<1, POLY>(vb, tw2, ch2, scale2, acc0, acc1, pk + 16);
            }
            float s0, s1;
            f32x2_unpack(f32x2_add(acc0, acc1), s0, s1);
            float us = s0 + s1;
            if (__any_sync(0xffffffffu, !(us < G2_UNIT_LIMIT))) {
                // rare: the reference is too low for this unit.  Raise it to the unit's true maximum, rescale what
                // has been accumulated with the old one (l and O, once the P V of the previous unit retired), redo.
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float t0, t1, t2, t3;
                    f32x2_unpack(tw2[i], t0, t1);
                    f32x2_unpack(tw2[16 + i], t2, t3);
                    mx = fmaxf(mx, fmaf(__uint_as_float(va[2 * i]), scale_log2, t0 + th));
                    mx = fmaxf(mx, fmaf(__uint_as_float(va[2 * i + 1]), scale_log2, t1 + th));
                    mx = fmaxf(mx, fmaf(__uint_as_float(vb[2 * i]), scale_log2, t2 + th));
                    mx = fmaxf(mx, fmaf(__uint_as_float(vb[2 * i + 1]), scale_log2, t3 + th));
                }
                const float m_new = fmaxf(m_used, mx);
                const float f = ex2_approx(m_used - m_new);
                l *= f;
                if (j > 0) {
                    mbar_wait(pv_done(t), static_cast<uint32_t>((j - 1) & 1));
                    tc_fence_after();
                    g2_scale_cols<HD>(t_o, f);
                }
                m_used = m_new;
                const float ch = th - m_used;
                const uint64_t ch2 = f32x2_pack(ch, ch);
                acc0 = 0ull;
                acc1 = 0ull;
                g2_chunk_exp<0, 0>(va, tw2, ch2, scale2, acc0, acc1, pk);
                g2_chunk_exp<1, 0>(vb, tw2, ch2, scale2, acc0, acc1, pk + 16);
                f32x2_unpack(f32x2_add(acc0, acc1), s0, s1);
                us = s0 + s1;
            }
            l += us;
            tmem_st_32x32(t_su, *reinterpret_cast<uint32_t(*)[32]>(pk));
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full(t, bb));
        }
        // ---- epilogue: O / l -> bf16 -> global (column = head*HD + d, image_encoder.py:501)
        mbar_wait(o_full(t), 0);
        tc_fence_after();
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = out + (static_cast<long long>(tok0) + qb * 256 + t * 128 + r) * D + head * HD;
        const bool live = !PLAIN || qb * 256 + t * 128 + r < L;        // query rows beyond the sequence are not stored
        // training forward: the row's log-sum-exp (log2 domain, rel-pos terms included) for the backward kernels
        const float lse_row = m_used + log2f(l);
        if (lse_out != nullptr && live)
            lse_out[(static_cast<long long>(tok0) + qb * 256 + t * 128 + r) * heads + head] = lse_row;
        if (!PLAIN && tw_out != nullptr && th_out != nullptr) {
            // ... and this row's rel-pos terms as the dK / dV kernels want them (they walk the queries as columns and cannot
            // project them themselves without a third of their MMAs): tw[kx] = T_w[q][qx - kx + 63] - lse[q] (so that
            // P = exp2(s * scale + tw + th) needs no further per-query term), th[ky] = T_h[q][qy - ky + 63]
            const long long row = (static_cast<long long>(tok0) + qb * 256 + t * 128 + r) * heads + head;
            float4* twp = reinterpret_cast<float4*>(tw_out + row * 64);
            float4* thp = reinterpret_cast<float4*>(th_out + row * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a0, a1, a2, a3;
                f32x2_unpack(tw2[2 * i], a0, a1);
                f32x2_unpack(tw2[2 * i + 1], a2, a3);
                twp[i] = make_float4(a0 - lse_row, a1 - lse_row, a2 - lse_row, a3 - lse_row);
            }
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int o = 128 + 2 * t + qrow + 63 - 4 * i;       // ky = 4 i .. 4 i + 3
                thp[i] = make_float4(trow[PLAIN ? 0 : o], trow[PLAIN ? 0 : o - 1], trow[PLAIN ? 0 : o - 2], trow[PLAIN ? 0 : o - 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(t_o + c * 32, va);
            tmem_ld_wait(va);
            if (!live) continue;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 o4;
                o4.x = pack_bf16(__uint_as_float(va[i]) * inv, __uint_as_float(va[i + 1]) * inv);
                o4.y = pack_bf16(__uint_as_float(va[i + 2]) * inv, __uint_as_float(va[i + 3]) * inv);
                o4.z = pack_bf16(__uint_as_float(va[i + 4]) * inv, __uint_as_float(va[i + 5]) * inv);
                o4.w = pack_bf16(__uint_as_float(va[i + 6]) * inv, __uint_as_float(va[i + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + c * 32 + i) = o4;
            }
        }
        if (HI) {
            uint32_t w16[16];
            tmem_ld_32x16(t_o + 64, w16);
            tmem_ld_wait16(w16);
#pragma unroll
            for (int i = 0; i < (live ? 16 : 0); i += 8) {
                uint4 o4;
                o4.x = pack_bf16(__uint_as_float(w16[i]) * inv, __uint_as_float(w16[i + 1]) * inv);
                o4.y = pack_bf16(__uint_as_float(w16[i + 2]) * inv, __uint_as_float(w16[i + 3]) * inv);
                o4.z = pack_bf16(__uint_as_float(w16[i + 4]) * inv, __uint_as_float(w16[i + 5]) * inv);
                o4.w = pack_bf16(__uint_as_float(w16[i + 6]) * inv, __uint_as_float(w16[i + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + 64 + i) = o4;
            }
        }
        }   // t < ntile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD, int POLY, bool PLAIN>
static int launch_global_tc2(const void* qkv, void* out, const void* relh, const void* relw, int batch, int heads,
                             int tokens, cudaStream_t stream, float* lse_out = nullptr, float* tw_out = nullptr,
                             float* th_out = nullptr) {
    const int D = heads * HD;
    CUtensorMap tq, tkv, th, tw;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 3;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = static_cast<unsigned long long>(tokens) * batch;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = 128;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    s.box[2] = G2_UNIT;
    CVLM_TRY(make_tensor_map_bf16(&tkv, s));
    MapSpec r = map_spec_zero();
    r.rank = 2;
    r.dims[0] = HD; r.dims[1] = 127;
    r.strides[0] = HD * 2ull;
    r.box[0] = 64; r.box[1] = G2_UNIT;
    if (PLAIN) {
        th = tkv;       // never read
        tw = tkv;
    } else {
        r.ptr = relh;
        CVLM_TRY(make_tensor_map_bf16(&th, r));
        r.ptr = relw;
        CVLM_TRY(make_tensor_map_bf16(&tw, r));
    }
    auto kern = attn_global_tc2_kernel<HD, POLY, PLAIN>;
    static unsigned char ready[kMaxDevices] = {};         // per (instantiation, device)
    {
        int dev = 0;
        CVLM_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices || !__atomic_load_n(&ready[dev], __ATOMIC_ACQUIRE)) {
            // setmaxnreg: 4 control warps shrink to 64, 8 softmax warps grow to 216 out of what those released
            cudaFuncAttributes fa;
            CVLM_CUDA(cudaFuncGetAttributes(&fa, kern));
            if ((fa.numRegs - 64) * 128 < (216 - fa.numRegs) * 256)
                return set_error(CVLM_ERR_UNSUPPORTED, "attention_tc2: built with %d registers per thread; the setmaxnreg pool would be short", fa.numRegs);
        }
    }
    CVLM_TRY(ensure_dynamic_smem(kern, G2_SMEM, ready));
    dim3 grid((tokens + 255) / 256, heads, batch);
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * kLog2eG2;
    kern<<<grid, G2_THREADS, G2_SMEM, stream>>>(tq, tkv, th, tw, static_cast<__nv_bfloat16*>(out), heads, D, scale_log2, tokens, lse_out, tw_out, th_out);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

template <int HD>
static int launch_global_tc2_poly(int poly, const void* qkv, void* out, const void* relh, const void* relw, int batch,
                                  int heads, cudaStream_t stream, float* lse_out, float* tw_out, float* th_out) {
    switch (poly) {
        case 0: return launch_global_tc2<HD, 0, false>(qkv, out, relh, relw, batch, heads, 4096, stream, lse_out, tw_out, th_out);
        case 2: return launch_global_tc2<HD, 2, false>(qkv, out, relh, relw, batch, heads, 4096, stream, lse_out, tw_out, th_out);
        case 4: return launch_global_tc2<HD, 4, false>(qkv, out, relh, relw, batch, heads, 4096, stream, lse_out, tw_out, th_out);
        case 6: return launch_global_tc2<HD, 6, false>(qkv, out, relh, relw, batch, heads, 4096, stream, lse_out, tw_out, th_out);
        case 8: return launch_global_tc2<HD, 8, false>(qkv, out, relh, relw, batch, heads, 4096, stream, lse_out, tw_out, th_out);
        default: return set_error(CVLM_ERR_ARG, "CVLM_ATTN_POLY must be 0, 2, 4, 6 or 8");
    }
}

// Global attention over a 64x64 token grid, head_dim 64 or 80.
// CVLM_ATTN_POLY = how many of every 16 exponential pairs run on the FMA pipe (tuning knob; default below).
int attention_global_tc2_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int heads,
                                int head_dim, cudaStream_t stream, float* lse_out, float* tw_out, float* th_out) {
    if (batch > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "attention: batch exceeds grid.z; split the batch");
    const int poly = option(OPT_ATTN_POLY);               // cvlm_set_option("attn_poly", n); default 4
    if (head_dim == 80) return launch_global_tc2_poly<80>(poly, qkv, out, relh, relw, batch, heads, stream, lse_out, tw_out, th_out);
    if (head_dim == 64) return launch_global_tc2_poly<64>(poly, qkv, out, relh, relw, batch, heads, stream, lse_out, tw_out, th_out);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_tc2: head_dim %d", head_dim);
}

// Plain attention softmax(q k^T / sqrt(hd)) v over `tokens` tokens per image (no relative-position terms): the visual
// tower of Alpha-CLIP (alpha_clip_rw/model.py:224-256 with the default arguments) on the same tcgen05
// pipeline.  qkv bf16 [batch * tokens, 3, heads, hd] (the output of in_proj), out bf16 [batch * tokens, heads * hd].
int attention_plain_tc2_launch(const void* qkv, void* out, int batch, int tokens, int heads, int head_dim,
                               cudaStream_t stream) {
    if (batch > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "attention: batch exceeds grid.z; split the batch");
    if (tokens < 1 || heads < 1) return set_error(CVLM_ERR_ARG, "attention_plain: tokens %d heads %d", tokens, heads);
    if (static_cast<long long>(tokens) * batch >= (1ll << 31)) return set_error(CVLM_ERR_UNSUPPORTED, "attention_plain: too many tokens");
    if (head_dim == 80) return launch_global_tc2<80, 4, true>(qkv, out, nullptr, nullptr, batch, heads, tokens, stream);
    if (head_dim == 64) return launch_global_tc2<64, 4, true>(qkv, out, nullptr, nullptr, batch, heads, tokens, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_plain: head_dim %d (64 or 80)", head_dim);
}

}  // namespace cvlm
