// dK / dV of the 64 x 64 global-attention blocks on tcgen05 (adapter-only training; the reference gets these
// from torch autograd through Attention.forward, image_encoder.py:493-501, 589-625).
//
// The forward kernel (attention_tc2.cu) with the roles swapped: one CTA owns 128 KEYS (two grid rows) of one (image, head)
// and streams the 64 query units of 64 queries (one grid row each).  Thread = key row (TMEM lane), columns = queries:
//   S^T(u)  = K Q_u^T           tcgen05.mma SS, M128 N64, K = 64 + 16      -> TMEM, two buffers
//   dP^T(u) = V dO_u^T          the same with V / dO
//   P^T  = exp2(S^T * scale + tw[q][kx] + th[q][ky])      (tw carries -lse[q]: BwdParams::tw_tab; no running max, no sums)
//   dS^T = P^T * (dP^T - D[q])                              bf16, written over the consumed S^T / dP^T columns
//   dV  += P^T dO_u,  dK += dS^T Q_u      tcgen05.mma TS (A from TMEM, Q_u / dO_u MN-major from the same shared-memory tiles
//                                         the SS MMAs read K-major), N = head_dim, accumulators in TMEM for the whole CTA
// Nothing is reduced along a row, so the two softmax warpgroups simply split the unit's 64 query columns (32 each) over all
// 128 key rows.  The per-query terms come from the training forward (no projection on the rel-pos tables here): the T_w
// tile [64 q][64 kx] fp32 arrives by TMA with the unit, {T_h[ky0], D, T_h[ky0 + 1], D} per query is staged by warp 3.
//
//   warp 0  TMA producer (K / V tiles once; Q_u, dO_u, T_w tile through a 3-stage ring)
//   warp 1  MMA issuer (one lane): S^T(u), dP^T(u) one unit ahead of dV / dK (u - 1)
//   warp 2  TMEM allocator (512 columns: S^T[b] at 64 b, dP^T[b] at 128 + 64 b, dV at 256, dK at 336)
//   warp 3  per-query scalars of the unit -> shared memory
//   warps 4-7 / 8-11  query columns 0..31 / 32..63 of every unit
#include "cvlm_internal.cuh"
#include "ptx.cuh"
#include "tensor_map.cuh"

namespace cvlm {

namespace {

constexpr int BT_UNIT = 64;                         // queries per unit
constexpr int BT_KT = 128 * 128;                    // one [128 rows x 64 bf16] swizzled tile, bytes
constexpr int BT_UT = BT_UNIT * 128;                // one [64 rows x 64 bf16] tile
constexpr int BT_TW = BT_UNIT * 64 * 4;             // T_w tile [64 q][64 kx] fp32
constexpr int BT_SC = BT_UNIT * 16;                 // {th0, D, th1, D} per query
constexpr int BT_NST = 3;
constexpr int BT_STAGE = 4 * BT_UT + BT_TW + BT_SC; // Q lo|hi, dO lo|hi, T_w, scalars
constexpr int BT_OFF_K = 0, BT_OFF_V = 2 * BT_KT, BT_OFF_ST = 4 * BT_KT;
constexpr int BT_OFF_BAR = BT_OFF_ST + BT_NST * BT_STAGE;
constexpr int BT_SMEM = BT_OFF_BAR + 256 + 1024;
static_assert(BT_SMEM <= 232448, "attention_bwd_tc shared memory");
static_assert(BT_STAGE % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzled tiles");
constexpr int BT_THREADS = 384;
constexpr uint32_t BT_COL_S = 0, BT_COL_DP = 128, BT_COL_DV = 256, BT_COL_DK = 336;
constexpr float kLog2eBT = 1.4426950408889634f;

struct BtParams {
    const float* th_tab;          // [B*4096, heads, 64]
    const float* dsum;            // [B*4096, heads]
    __nv_bfloat16* dqkv;          // [B*4096, 3*D]
    int heads, D;
    float scale, scale_log2;
};

__device__ __forceinline__ float bt_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                       const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmTW, const BtParams p) {
    constexpr bool HI = HD > 64;
    constexpr uint32_t KV_BYTES = (HI ? 2 : 1) * BT_KT;
    constexpr uint32_t U_BYTES = (HI ? 2 : 1) * BT_UT;
    constexpr int NU = 4096 / BT_UNIT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sK = base + BT_OFF_K, sV = base + BT_OFF_V;
    auto sQ = [&](int st) { return base + BT_OFF_ST + st * BT_STAGE; };
    auto sDO = [&](int st) { return sQ(st) + 2 * BT_UT; };
    auto sTW = [&](int st) { return sQ(st) + 4 * BT_UT; };
    auto sSC = [&](int st) { return sTW(st) + BT_TW; };
    const uint32_t bars = base + BT_OFF_BAR;
    const uint32_t kv_full = bars;
    auto st_full = [&](int s) { return bars + 8u * (1 + s); };
    auto st_empty = [&](int s) { return bars + 8u * (4 + s); };
    auto sdp_full = [&](int b) { return bars + 8u * (7 + b); };
    auto pds_full = [&](int b) { return bars + 8u * (9 + b); };
    const uint32_t acc_full = bars + 8u * 11;
    const uint32_t tmem_slot = bars + 8u * 12;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + BT_OFF_BAR + 8 * 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int heads = p.heads;
    const int tok0 = b * 4096;
    const int ky0 = 2 * kt;                          // the tile's two key rows

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmKV128);
        tma_prefetch_desc(&tmQ64);
        tma_prefetch_desc(&tmDO64);
        tma_prefetch_desc(&tmTW);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int s = 0; s < BT_NST; ++s) {
            mbar_init(st_full(s), 2);                // the TMA producer's expect_tx arrive + warp 3's arrive
            mbar_init(st_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(sdp_full(i), 1);
            mbar_init(pds_full(i), 256);
        }
        mbar_init(acc_full, 1);
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
            mbar_arrive_expect_tx(kv_full, 2 * KV_BYTES);
            const int ktok = tok0 + kt * 128;
            tma_load_3d(sK, &tmKV128, kv_full, 0, heads + head, ktok);
            if (HI) tma_load_3d(sK + BT_KT, &tmKV128, kv_full, 64, heads + head, ktok);
            tma_load_3d(sV, &tmKV128, kv_full, 0, 2 * heads + head, ktok);
            if (HI) tma_load_3d(sV + BT_KT, &tmKV128, kv_full, 64, 2 * heads + head, ktok);
            int st = 0;
            uint32_t ph = 0;
            for (int u = 0; u < NU; ++u) {
                mbar_wait_relaxed(st_empty(st), ph ^ 1u);
                mbar_arrive_expect_tx(st_full(st), 2 * U_BYTES + BT_TW);
                const int qtok = tok0 + u * BT_UNIT;
                tma_load_3d(sQ(st), &tmQ64, st_full(st), 0, head, qtok);
                if (HI) tma_load_3d(sQ(st) + BT_UT, &tmQ64, st_full(st), 64, head, qtok);
                tma_load_3d(sDO(st), &tmDO64, st_full(st), 0, head, qtok);
                if (HI) tma_load_3d(sDO(st) + BT_UT, &tmDO64, st_full(st), 64, head, qtok);
                tma_load_3d(sTW(st), &tmTW, st_full(st), 0, head, qtok);
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
        } else if (warp == 1 && elect_one_sync()) {
            // ============================================================ MMA issuer
            constexpr uint32_t idescS = umma_idesc_bf16(128, BT_UNIT);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            mbar_wait(kv_full, 0);
            tc_fence_after();
            const uint64_t k_lo = umma_desc_k_sw128(sK), k_hi = umma_desc_k_sw128(sK + BT_KT);
            const uint64_t v_lo = umma_desc_k_sw128(sV), v_hi = umma_desc_k_sw128(sV + BT_KT);
            auto issue_acc = [&](int j) {            // dV += P^T(j) dO_j, dK += dS^T(j) Q_j
                const int stj = j % BT_NST, bj = j & 1;
                mbar_wait(pds_full(bj), static_cast<uint32_t>((j >> 1) & 1));
                tc_fence_after();
                const uint64_t do_mn = umma_desc_mn_sw128(sDO(stj), BT_UT), q_mn = umma_desc_mn_sw128(sQ(stj), BT_UT);
                const uint32_t aP = tmem_base + BT_COL_S + bj * 64, aDS = tmem_base + BT_COL_DP + bj * 64;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {     // 16 queries per MMA; P / dS of queries 32.. live at columns 32..
                    const uint32_t pc = (kk & 1) * 8u + (kk >> 1) * 32u;
                    umma_bf16_ts(tmem_base + BT_COL_DV, aP + pc, do_mn + 128u * kk, idescO, (j > 0 || kk > 0) ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t pc = (kk & 1) * 8u + (kk >> 1) * 32u;
                    umma_bf16_ts(tmem_base + BT_COL_DK, aDS + pc, q_mn + 128u * kk, idescO, (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(st_empty(stj));
            };
            int st = 0;
            uint32_t ph = 0;
            for (int u = 0; u < NU; ++u) {
                mbar_wait(st_full(st), ph);
                tc_fence_after();
                const int bu = u & 1;
                const uint64_t q_lo = umma_desc_k_sw128(sQ(st)), q_hi = umma_desc_k_sw128(sQ(st) + BT_UT);
                const uint64_t d_lo = umma_desc_k_sw128(sDO(st)), d_hi = umma_desc_k_sw128(sDO(st) + BT_UT);
                const uint32_t tS = tmem_base + BT_COL_S + bu * 64, tDP = tmem_base + BT_COL_DP + bu * 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, k_lo + 2u * k, q_lo + 2u * k, idescS, k > 0);
                if (HI) umma_bf16_ss(tS, k_hi, q_hi, idescS, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tDP, v_lo + 2u * k, d_lo + 2u * k, idescS, k > 0);
                if (HI) umma_bf16_ss(tDP, v_hi, d_hi, idescS, 1);
                umma_commit(sdp_full(bu));
                if (u >= 1) issue_acc(u - 1);
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
            issue_acc(NU - 1);
            umma_commit(acc_full);
        } else if (warp == 3) {
            // ============================================================ per-query scalars of every unit
            int st = 0;
            uint32_t ph = 0;
            for (int u = 0; u < NU; ++u) {
                if (lane == 0) mbar_wait_relaxed(st_empty(st), ph ^ 1u);
                __syncwarp();
                float4* dst = reinterpret_cast<float4*>(base_ptr + (sSC(st) - base));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = lane + 32 * i;
                    const long long row = (static_cast<long long>(tok0) + u * BT_UNIT + q) * heads + head;
                    const float2 th = *reinterpret_cast<const float2*>(p.th_tab + row * 64 + ky0);
                    const float d = p.dsum[row];
                    dst[q] = make_float4(th.x, d, th.y, d);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(st_full(st));
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
        }
    } else {
        // ============================================================ P^T / dS^T: thread = key row, 32 query columns
        setmaxnreg_inc<216>();
        const int w = warp >= 8 ? 1 : 0;
        const int r = threadIdx.x - 128 - w * 128;          // key row 0..127 of the tile
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int kx = r & 63, kysel = r >> 6;
        const uint32_t col_in = 32u * w;                     // S^T / dP^T columns this warpgroup reads
        const uint32_t col_out = 32u * w;                    // ... and the 16 packed columns it writes (its own range)
        int st = 0;
        uint32_t ph = 0;
#pragma unroll 1
        for (int u = 0; u < NU; ++u) {
            const int bu = u & 1;
            mbar_wait(st_full(st), ph);
            mbar_wait(sdp_full(bu), static_cast<uint32_t>((u >> 1) & 1));
            tc_fence_after();
            const uint32_t tS = tmem_base + lane_addr + BT_COL_S + bu * 64, tDP = tmem_base + lane_addr + BT_COL_DP + bu * 64;
            uint32_t vs[32], vd[32];
            tmem_ld_32x32(tS + col_in, vs);
            tmem_ld_32x32(tDP + col_in, vd);
            tmem_ld_wait(vs);
            tmem_ld_wait(vd);
            const uint32_t tw_addr = sTW(st) + (32u * w * 64u + kx) * 4u;       // + q * 256
            const uint32_t sc_addr = sSC(st) + 32u * w * 16u + kysel * 8u;      // + q * 16: {th[ky], D}
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float pv[2], dv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float tw = ld_shared_f32(tw_addr + (c + e) * 256);
                    const float2 sc = ld_shared_f2(sc_addr + (c + e) * 16);
                    const float x = fmaf(__uint_as_float(vs[c + e]), p.scale_log2, tw + sc.x);
                    pv[e] = bt_ex2(x);
                    dv[e] = pv[e] * (__uint_as_float(vd[c + e]) - sc.y);
                }
                pk[c >> 1] = pack_bf16(pv[0], pv[1]);
                dk[c >> 1] = pack_bf16(dv[0], dv[1]);
            }
            tmem_st_32x16(tS + col_out, pk);
            tmem_st_32x16(tDP + col_out, dk);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(pds_full(bu));
            if (++st == BT_NST) { st = 0; ph ^= 1u; }
        }
        // ---- epilogue: warpgroup 0 stores dK (scaled), warpgroup 1 dV
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_addr + (w == 0 ? BT_COL_DK : BT_COL_DV);
        const float f = w == 0 ? p.scale : 1.0f;
        __nv_bfloat16* orow = p.dqkv + (static_cast<long long>(tok0) + kt * 128 + r) * (3LL * p.D) + (w == 0 ? 1 : 2) * p.D + head * HD;
        uint32_t va[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(t_acc + c * 32, va);
            tmem_ld_wait(va);
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 o4;
                o4.x = pack_bf16(__uint_as_float(va[i]) * f, __uint_as_float(va[i + 1]) * f);
                o4.y = pack_bf16(__uint_as_float(va[i + 2]) * f, __uint_as_float(va[i + 3]) * f);
                o4.z = pack_bf16(__uint_as_float(va[i + 4]) * f, __uint_as_float(va[i + 5]) * f);
                o4.w = pack_bf16(__uint_as_float(va[i + 6]) * f, __uint_as_float(va[i + 7]) * f);
                *reinterpret_cast<uint4*>(orow + c * 32 + i) = o4;
            }
        }
        if (HI) {
            uint32_t w16[16];
            tmem_ld_32x16(t_acc + 64, w16);
            tmem_ld_wait16(w16);
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 o4;
                o4.x = pack_bf16(__uint_as_float(w16[i]) * f, __uint_as_float(w16[i + 1]) * f);
                o4.y = pack_bf16(__uint_as_float(w16[i + 2]) * f, __uint_as_float(w16[i + 3]) * f);
                o4.z = pack_bf16(__uint_as_float(w16[i + 4]) * f, __uint_as_float(w16[i + 5]) * f);
                o4.w = pack_bf16(__uint_as_float(w16[i + 6]) * f, __uint_as_float(w16[i + 7]) * f);
                *reinterpret_cast<uint4*>(orow + 64 + i) = o4;
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

// ------------------------------------------------------------------------------------------------------------------
// dQ of the same blocks (and D = rowsum(dO o), which the dK / dV kernel reads).  One CTA owns 128 QUERIES (two grid rows) and
// streams the 64 key units (one grid row each); thread = query row, and -- because every per-row term (lse, D, the rel-pos
// terms) is known before the first unit -- the two softmax warpgroups again split the unit's 64 key columns without talking
// to each other:
//   S(u) = Q K_u^T,  dP(u) = dO V_u^T          SS, M128 N64 -> TMEM (two buffers each)
//   dS = exp2(S * scale + tw[kx] + th[ky = u]) * (dP - D)      dSh[q][u] += sum_kx dS (a scalar per unit and thread),
//                                                              dSw[q][kx] += dS (32 registers per thread)
//   dQ += (scale * dS) K_u                      TS, K_u MN-major from the tile the SS MMA read K-major
// and at the end  dQ += dT_h R_h + dT_w R_w  with  dT_h[q][j] = dSh[q][qy + 63 - j],  dT_w[q][j] = dSw[q][qx + 63 - j]:
// each warpgroup writes its PARTIAL dT rows (bf16) over the retired S / dP buffers and the MMA adds all four products; the
// two tables travel through the unit ring as "units" 64 and 65.
constexpr int BQ_STAGE = 4 * BT_UT;                 // K lo|hi, V lo|hi (or one rel-pos table: rows 0..63 | 64..127)
constexpr int BQ_OFF_Q = 0, BQ_OFF_DO = 2 * BT_KT, BQ_OFF_ST = 4 * BT_KT;
constexpr int BQ_OFF_BAR = BQ_OFF_ST + BT_NST * BQ_STAGE;
constexpr int BQ_SMEM = BQ_OFF_BAR + 256 + 1024;
constexpr uint32_t BQ_COL_DQ = 256;

struct BqParams {
    const float* tw_tab;          // [B*4096, heads, 64]  T_w[q][qx - kx + 63] - lse[q]
    const float* th_tab;          // [B*4096, heads, 64]  T_h[q][qy - ky + 63]
    const __nv_bfloat16* o;       // [B*4096, D] forward output
    const __nv_bfloat16* dout;    // [B*4096, D]
    float* dsum;                  // [B*4096, heads] out: D
    __nv_bfloat16* dqkv;          // [B*4096, 3*D]
    int heads, D;
    float scale, scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmDO128,
                      const __grid_constant__ CUtensorMap tmKV64, const __grid_constant__ CUtensorMap tmRelH,
                      const __grid_constant__ CUtensorMap tmRelW, const BqParams p) {
    constexpr bool HI = HD > 64;
    constexpr uint32_t T_BYTES = (HI ? 2 : 1) * BT_KT;
    constexpr uint32_t U_BYTES = (HI ? 2 : 1) * BT_UT;
    constexpr int NK = 4096 / BT_UNIT;               // key units; units NK, NK + 1 of the ring are the two tables
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base + BQ_OFF_Q, sDO = base + BQ_OFF_DO;
    auto sK = [&](int st) { return base + BQ_OFF_ST + st * BQ_STAGE; };
    auto sV = [&](int st) { return sK(st) + 2 * BT_UT; };
    const uint32_t bars = base + BQ_OFF_BAR;
    const uint32_t qd_full = bars;
    auto st_full = [&](int s) { return bars + 8u * (1 + s); };
    auto st_empty = [&](int s) { return bars + 8u * (4 + s); };
    auto sdp_full = [&](int b) { return bars + 8u * (7 + b); };
    auto ds_full = [&](int b) { return bars + 8u * (9 + b); };
    const uint32_t dq_done = bars + 8u * 11, dt_full = bars + 8u * 12, acc_full = bars + 8u * 13;
    const uint32_t tmem_slot = bars + 8u * 14;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + BQ_OFF_BAR + 8 * 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int heads = p.heads;
    const int tok0 = b * 4096;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ128);
        tma_prefetch_desc(&tmDO128);
        tma_prefetch_desc(&tmKV64);
        tma_prefetch_desc(&tmRelH);
        tma_prefetch_desc(&tmRelW);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(qd_full, 1);
        for (int s = 0; s < BT_NST; ++s) {
            mbar_init(st_full(s), 1);
            mbar_init(st_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(sdp_full(i), 1);
            mbar_init(ds_full(i), 256);
        }
        mbar_init(dq_done, 1);
        mbar_init(dt_full, 256);
        mbar_init(acc_full, 1);
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
            mbar_arrive_expect_tx(qd_full, 2 * T_BYTES);
            const int qtok = tok0 + qt * 128;
            tma_load_3d(sQ, &tmQ128, qd_full, 0, head, qtok);
            if (HI) tma_load_3d(sQ + BT_KT, &tmQ128, qd_full, 64, head, qtok);
            tma_load_3d(sDO, &tmDO128, qd_full, 0, head, qtok);
            if (HI) tma_load_3d(sDO + BT_KT, &tmDO128, qd_full, 64, head, qtok);
            int st = 0;
            uint32_t ph = 0;
            for (int u = 0; u < NK + 2; ++u) {
                mbar_wait_relaxed(st_empty(st), ph ^ 1u);
                mbar_arrive_expect_tx(st_full(st), 2 * U_BYTES);
                if (u < NK) {
                    const int ktok = tok0 + u * BT_UNIT;
                    tma_load_3d(sK(st), &tmKV64, st_full(st), 0, heads + head, ktok);
                    if (HI) tma_load_3d(sK(st) + BT_UT, &tmKV64, st_full(st), 64, heads + head, ktok);
                    tma_load_3d(sV(st), &tmKV64, st_full(st), 0, 2 * heads + head, ktok);
                    if (HI) tma_load_3d(sV(st) + BT_UT, &tmKV64, st_full(st), 64, 2 * heads + head, ktok);
                } else {
                    const CUtensorMap* tm = u == NK ? &tmRelH : &tmRelW;     // rows 0..63 | 64..127 (row 127: zeros)
                    tma_load_2d(sK(st), tm, st_full(st), 0, 0);
                    if (HI) tma_load_2d(sK(st) + BT_UT, tm, st_full(st), 64, 0);
                    tma_load_2d(sV(st), tm, st_full(st), 0, 64);
                    if (HI) tma_load_2d(sV(st) + BT_UT, tm, st_full(st), 64, 64);
                }
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
        } else if (warp == 1 && elect_one_sync()) {
            // ============================================================ MMA issuer
            constexpr uint32_t idescS = umma_idesc_bf16(128, BT_UNIT);
            constexpr uint32_t idescO = umma_idesc_bf16_bmn(128, HD);
            mbar_wait(qd_full, 0);
            tc_fence_after();
            const uint64_t q_lo = umma_desc_k_sw128(sQ), q_hi = umma_desc_k_sw128(sQ + BT_KT);
            const uint64_t d_lo = umma_desc_k_sw128(sDO), d_hi = umma_desc_k_sw128(sDO + BT_KT);
            const uint32_t tDQ = tmem_base + BQ_COL_DQ;
            auto issue_dq = [&](int j) {             // dQ += (scale dS(j)) K_j
                const int stj = j % BT_NST, bj = j & 1;
                mbar_wait(ds_full(bj), static_cast<uint32_t>((j >> 1) & 1));
                tc_fence_after();
                const uint64_t k_mn = umma_desc_mn_sw128(sK(stj), BT_UT);
                const uint32_t aDS = tmem_base + BT_COL_S + bj * 64;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t pc = (kk & 1) * 8u + (kk >> 1) * 32u;
                    umma_bf16_ts(tDQ, aDS + pc, k_mn + 128u * kk, idescO, (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(st_empty(stj));
            };
            int st = 0;
            uint32_t ph = 0;
            for (int u = 0; u < NK; ++u) {
                mbar_wait(st_full(st), ph);
                tc_fence_after();
                const int bu = u & 1;
                const uint64_t k_lo = umma_desc_k_sw128(sK(st)), k_hi = umma_desc_k_sw128(sK(st) + BT_UT);
                const uint64_t v_lo = umma_desc_k_sw128(sV(st)), v_hi = umma_desc_k_sw128(sV(st) + BT_UT);
                const uint32_t tS = tmem_base + BT_COL_S + bu * 64, tDP = tmem_base + BT_COL_DP + bu * 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, q_lo + 2u * k, k_lo + 2u * k, idescS, k > 0);
                if (HI) umma_bf16_ss(tS, q_hi, k_hi, idescS, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tDP, d_lo + 2u * k, v_lo + 2u * k, idescS, k > 0);
                if (HI) umma_bf16_ss(tDP, d_hi, v_hi, idescS, 1);
                umma_commit(sdp_full(bu));
                if (u >= 1) issue_dq(u - 1);
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
            issue_dq(NK - 1);
            umma_commit(dq_done);                    // S / dP buffers are free once this fires: the dT rows go there
            // ---- dQ += dT_h R_h + dT_w R_w: four partial dT operands (two warpgroups x {h, w}), 128 table rows each
            int st_tab[2];
            for (int i = 0; i < 2; ++i) {
                mbar_wait(st_full(st), ph);
                st_tab[i] = st;
                if (++st == BT_NST) { st = 0; ph ^= 1u; }
            }
            mbar_wait(dt_full, 0);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 2; ++i) {            // i = 0: rel_pos_h, 1: rel_pos_w
                const uint64_t r0 = umma_desc_mn_sw128(sK(st_tab[i]), BT_UT), r1 = umma_desc_mn_sw128(sV(st_tab[i]), BT_UT);
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const uint32_t aT = tmem_base + (i == 0 ? BT_COL_S : BT_COL_DP) + w * 64;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16_ts(tDQ, aT + 8u * kk, (kk < 4 ? r0 + 128u * kk : r1 + 128u * (kk - 4)), idescO, 1u);
                }
            }
            umma_commit(acc_full);
        }
    } else {
        // ============================================================ dS: thread = query row, 32 key columns
        setmaxnreg_inc<216>();
        const int w = warp >= 8 ? 1 : 0;
        const int r = threadIdx.x - 128 - w * 128;          // query row 0..127 of the tile
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int qy_rel = r >> 6, qx = r & 63;              // qy = 2 qt + qy_rel
        const long long tok = static_cast<long long>(tok0) + qt * 128 + r;
        const long long row = tok * heads + head;
        // ---- per-row terms: D = rowsum(dO o); tw for this warpgroup's 32 kx; the whole th row (indexed by the unit)
        float dsum = 0.f;
        {
            const __nv_bfloat16* po = p.o + tok * p.D + head * HD;
            const __nv_bfloat16* pd = p.dout + tok * p.D + head * HD;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                const uint4 a = *reinterpret_cast<const uint4*>(po + c * 8), g = *reinterpret_cast<const uint4*>(pd + c * 8);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dsum = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), dsum);
                    dsum = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(gw[e] & 0xffff0000u), dsum);
                }
            }
            if (w == 0) p.dsum[row] = dsum;
        }
        float tw[32], th_l[64], dsw[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(p.tw_tab + row * 64 + 32 * w + 4 * i);
            tw[4 * i] = v.x; tw[4 * i + 1] = v.y; tw[4 * i + 2] = v.z; tw[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(p.th_tab + row * 64 + 4 * i);
            th_l[4 * i] = v.x; th_l[4 * i + 1] = v.y; th_l[4 * i + 2] = v.z; th_l[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) dsw[i] = 0.f;
        const uint32_t col = 32u * w;
#pragma unroll 1
        for (int u = 0; u < NK; ++u) {
            const int bu = u & 1;
            const float th = th_l[u];
            mbar_wait(sdp_full(bu), static_cast<uint32_t>((u >> 1) & 1));
            tc_fence_after();
            const uint32_t tS = tmem_base + lane_addr + BT_COL_S + bu * 64, tDP = tmem_base + lane_addr + BT_COL_DP + bu * 64;
            uint32_t vs[32], vd[32];
            tmem_ld_32x32(tS + col, vs);
            tmem_ld_32x32(tDP + col, vd);
            tmem_ld_wait(vs);
            tmem_ld_wait(vd);
            uint32_t pk[16];
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float ds[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float x = fmaf(__uint_as_float(vs[c + e]), p.scale_log2, tw[c + e] + th);
                    ds[e] = bt_ex2(x) * (__uint_as_float(vd[c + e]) - dsum);
                    dsw[c + e] += ds[e];
                    rs += ds[e];
                }
                pk[c >> 1] = pack_bf16(ds[0] * p.scale, ds[1] * p.scale);
            }
            th_l[u] = rs;                                // dSh[q][ky = u] (this warpgroup's share) takes the consumed slot
            tmem_st_32x16(tS + col, pk);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(ds_full(bu));
        }
        // ---- partial dT rows over the retired buffers: dT_h at S[w], dT_w at dP[w]; column c holds table rows 2c, 2c + 1
        float dsw_l[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) dsw_l[i] = dsw[i];
        mbar_wait(dq_done, 0);
        tc_fence_after();
        const int qy = 2 * qt + qy_rel;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t hk[32], wk[32];
#pragma unroll 4
            for (int c = 0; c < 32; ++c) {
                float hv[2], wv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = half * 64 + 2 * c + e;                 // table row
                    const int ky = qy + 63 - j, kxw = qx + 63 - j - 32 * w;
                    hv[e] = (ky >= 0 && ky < 64) ? th_l[ky] : 0.f;
                    wv[e] = (kxw >= 0 && kxw < 32) ? dsw_l[kxw] : 0.f;
                }
                hk[c] = pack_bf16(hv[0], hv[1]);
                wk[c] = pack_bf16(wv[0], wv[1]);
            }
            tmem_st_32x32(tmem_base + lane_addr + BT_COL_S + w * 64 + half * 32, hk);
            tmem_st_32x32(tmem_base + lane_addr + BT_COL_DP + w * 64 + half * 32, wk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(dt_full);
        // ---- epilogue: dQ -> bf16 -> global; warpgroup 0 stores d 0..31 (and 64..79), warpgroup 1 d 32..63
        mbar_wait(acc_full, 0);
        tc_fence_after();
        __nv_bfloat16* orow = p.dqkv + tok * (3LL * p.D) + head * HD;
        uint32_t va[32];
        tmem_ld_32x32(tmem_base + lane_addr + BQ_COL_DQ + 32 * w, va);
        tmem_ld_wait(va);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            uint4 o4;
            o4.x = pack_bf16(__uint_as_float(va[i]), __uint_as_float(va[i + 1]));
            o4.y = pack_bf16(__uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
            o4.z = pack_bf16(__uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
            o4.w = pack_bf16(__uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
            *reinterpret_cast<uint4*>(orow + 32 * w + i) = o4;
        }
        if (HI && w == 0) {
            uint32_t w16[16];
            tmem_ld_32x16(tmem_base + lane_addr + BQ_COL_DQ + 64, w16);
            tmem_ld_wait16(w16);
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 o4;
                o4.x = pack_bf16(__uint_as_float(w16[i]), __uint_as_float(w16[i + 1]));
                o4.y = pack_bf16(__uint_as_float(w16[i + 2]), __uint_as_float(w16[i + 3]));
                o4.z = pack_bf16(__uint_as_float(w16[i + 4]), __uint_as_float(w16[i + 5]));
                o4.w = pack_bf16(__uint_as_float(w16[i + 6]), __uint_as_float(w16[i + 7]));
                *reinterpret_cast<uint4*>(orow + 64 + i) = o4;
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
int launch_dq_tc(const void* qkv, const void* o, const void* dout, void* dqkv, float* dsum, const float* tw_tab,
                 const float* th_tab, const void* relh, const void* relw, int batch, int heads, cudaStream_t stream) {
    const int D = heads * HD;
    const unsigned long long M = 4096ull * batch;
    CUtensorMap tq, tdo, tkv, trh, trw;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 3;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = M;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = 128;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    s.box[2] = BT_UNIT;
    CVLM_TRY(make_tensor_map_bf16(&tkv, s));
    MapSpec d = map_spec_zero();
    d.ptr = dout;
    d.rank = 3;
    d.dims[0] = HD; d.dims[1] = heads; d.dims[2] = M;
    d.strides[0] = HD * 2ull; d.strides[1] = D * 2ull;
    d.box[0] = 64; d.box[1] = 1; d.box[2] = 128;
    CVLM_TRY(make_tensor_map_bf16(&tdo, d));
    MapSpec r = map_spec_zero();
    r.rank = 2;
    r.dims[0] = HD; r.dims[1] = 127;
    r.strides[0] = HD * 2ull;
    r.box[0] = 64; r.box[1] = 64;
    r.ptr = relh;
    CVLM_TRY(make_tensor_map_bf16(&trh, r));
    r.ptr = relw;
    CVLM_TRY(make_tensor_map_bf16(&trw, r));
    auto kern = attn_bwd_dq_tc_kernel<HD>;
    static unsigned char ready[kMaxDevices] = {};
    {
        int dev = 0;
        CVLM_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices || !__atomic_load_n(&ready[dev], __ATOMIC_ACQUIRE)) {
            cudaFuncAttributes fa;
            CVLM_CUDA(cudaFuncGetAttributes(&fa, kern));
            if ((fa.numRegs - 64) * 128 < (216 - fa.numRegs) * 256)
                return set_error(CVLM_ERR_UNSUPPORTED, "attention_bwd_tc: built with %d registers per thread; the setmaxnreg pool would be short", fa.numRegs);
        }
    }
    CVLM_TRY(ensure_dynamic_smem(kern, BQ_SMEM, ready));
    BqParams p;
    p.tw_tab = tw_tab;
    p.th_tab = th_tab;
    p.o = static_cast<const __nv_bfloat16*>(o);
    p.dout = static_cast<const __nv_bfloat16*>(dout);
    p.dsum = dsum;
    p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    p.heads = heads;
    p.D = D;
    p.scale = 1.0f / sqrtf(static_cast<float>(HD));
    p.scale_log2 = p.scale * kLog2eBT;
    dim3 grid(4096 / 128, heads, batch);
    kern<<<grid, BT_THREADS, BQ_SMEM, stream>>>(tq, tdo, tkv, trh, trw, p);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

template <int HD>
int launch_dkv_tc(const void* qkv, const void* dout, void* dqkv, const float* dsum, const float* tw_tab, const float* th_tab,
                  int batch, int heads, cudaStream_t stream) {
    const int D = heads * HD;
    const unsigned long long M = 4096ull * batch;
    CUtensorMap tkv, tq, tdo, ttw;
    MapSpec s = map_spec_zero();
    s.ptr = qkv;
    s.rank = 3;
    s.dims[0] = HD; s.dims[1] = 3ull * heads; s.dims[2] = M;
    s.strides[0] = HD * 2ull; s.strides[1] = 3ull * D * 2;
    s.box[0] = 64; s.box[1] = 1; s.box[2] = 128;
    CVLM_TRY(make_tensor_map_bf16(&tkv, s));
    s.box[2] = BT_UNIT;
    CVLM_TRY(make_tensor_map_bf16(&tq, s));
    MapSpec d = map_spec_zero();
    d.ptr = dout;
    d.rank = 3;
    d.dims[0] = HD; d.dims[1] = heads; d.dims[2] = M;
    d.strides[0] = HD * 2ull; d.strides[1] = D * 2ull;
    d.box[0] = 64; d.box[1] = 1; d.box[2] = BT_UNIT;
    CVLM_TRY(make_tensor_map_bf16(&tdo, d));
    MapSpec t = map_spec_zero();
    t.ptr = tw_tab;
    t.rank = 3;
    t.dims[0] = 64; t.dims[1] = heads; t.dims[2] = M;
    t.strides[0] = 64 * 4ull; t.strides[1] = 64ull * 4 * heads;
    t.box[0] = 64; t.box[1] = 1; t.box[2] = BT_UNIT;
    CVLM_TRY(make_tensor_map_f32_plain(&ttw, t));
    auto kern = attn_bwd_dkv_tc_kernel<HD>;
    static unsigned char ready[kMaxDevices] = {};
    {
        int dev = 0;
        CVLM_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices || !__atomic_load_n(&ready[dev], __ATOMIC_ACQUIRE)) {
            cudaFuncAttributes fa;
            CVLM_CUDA(cudaFuncGetAttributes(&fa, kern));
            if ((fa.numRegs - 64) * 128 < (216 - fa.numRegs) * 256)
                return set_error(CVLM_ERR_UNSUPPORTED, "attention_bwd_tc: built with %d registers per thread; the setmaxnreg pool would be short", fa.numRegs);
        }
    }
    CVLM_TRY(ensure_dynamic_smem(kern, BT_SMEM, ready));
    BtParams p;
    p.th_tab = th_tab;
    p.dsum = dsum;
    p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    p.heads = heads;
    p.D = D;
    p.scale = 1.0f / sqrtf(static_cast<float>(HD));
    p.scale_log2 = p.scale * kLog2eBT;
    dim3 grid(4096 / 128, heads, batch);
    kern<<<grid, BT_THREADS, BT_SMEM, stream>>>(tkv, tq, tdo, ttw, p);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace

// dK and dV columns of dqkv for the 64 x 64 global blocks (head_dim 64 or 80), given the forward's tables and the D = rowsum(dO o)
// the dQ kernel has just written.
int attention_bwd_dkv_tc_launch(const void* qkv, const void* dout, void* dqkv, const float* dsum, const float* tw_tab,
                                const float* th_tab, int batch, int heads, int head_dim, cudaStream_t stream) {
    if (batch > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: batch exceeds grid.z; split the batch");
    if (head_dim == 80) return launch_dkv_tc<80>(qkv, dout, dqkv, dsum, tw_tab, th_tab, batch, heads, stream);
    if (head_dim == 64) return launch_dkv_tc<64>(qkv, dout, dqkv, dsum, tw_tab, th_tab, batch, heads, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_bwd_dkv_tc: head_dim %d", head_dim);
}

// dQ columns of dqkv and D = rowsum(dO o) (into dsum) for the 64 x 64 global blocks, given the forward's tables.
int attention_bwd_dq_tc_launch(const void* qkv, const void* o, const void* dout, void* dqkv, float* dsum, const float* tw_tab,
                               const float* th_tab, const void* relh, const void* relw, int batch, int heads, int head_dim,
                               cudaStream_t stream) {
    if (batch > 65535) return set_error(CVLM_ERR_UNSUPPORTED, "attention_backward: batch exceeds grid.z; split the batch");
    if ((reinterpret_cast<uintptr_t>(relh) & 15) || (reinterpret_cast<uintptr_t>(relw) & 15))
        return set_error(CVLM_ERR_ARG, "attention_backward: rel-pos tables must be 16-byte aligned");
    if (head_dim == 80) return launch_dq_tc<80>(qkv, o, dout, dqkv, dsum, tw_tab, th_tab, relh, relw, batch, heads, stream);
    if (head_dim == 64) return launch_dq_tc<64>(qkv, o, dout, dqkv, dsum, tw_tab, th_tab, relh, relw, batch, heads, stream);
    return set_error(CVLM_ERR_UNSUPPORTED, "attention_bwd_dq_tc: head_dim %d", head_dim);
}

}  // namespace cvlm
