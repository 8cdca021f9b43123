// Persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A1 * B1^T [+ A2 * B2^T]).
//
// Carries every nn.Linear / 1x1 / im2col'd conv of the encoder:
//   PatchEmbed.proj (image_encoder.py:651-659, + pos_embed :139-140 as a row-periodic residual),
//   Attention.qkv / .proj (:476-477, :491, :502), MLPBlock.lin1 / lin2 (common.py:21-26, GELU and
//   the shortcut adds of Block.forward :443-444 fused), the adapter linears (:252-259, :287-296 -- the
//   rank-r prompt of the NEXT block rides along as a second K segment of lin2), neck convs (:98-114).
//
// Design: one CTA per SM, 128x256 output tile, 64-wide K blocks.
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (SWIZZLE_128B) into a 4-stage smem ring
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M128 N256 K16, bf16 -> fp32 in TMEM)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered)
//   warps 4-11  epilogue: tcgen05.ld -> registers -> bias / GELU / residual -> global
// The accumulator double buffer lets the MMA of tile i+1 overlap the epilogue of tile i.
// Operands are K-major in global memory (activations [M,K], weights [N,K]); ragged M/N/K edges
// are handled by TMA out-of-bounds zero fill plus predicated stores.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, STAGES_PAIR = 6, UMMA_K = 16;
constexpr int A_STAGE = BM * BK * 2;     // 16 KB
constexpr int B_STAGE = BN * BK * 2;     // 32 KB
constexpr int EPI_WARPS = 8;             // (16 warps / 5 stages measured slower on every GEMM: spills + a shallower ring)
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;   // 4 control warps + the epilogue warps
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);       // accumulator columns per epilogue warp (64)
constexpr int EPI_CHUNKS = EPI_COLS / 32;
static_assert(EPI_COLS == CVLM_STAT_GROUP_COLS, "row statistics are written per epilogue warp");

constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGING = EPI_WARPS * 4096;   // one 32x32 fp32 transpose tile per epilogue warp
constexpr int GEMM_SMEM = STAGES * (A_STAGE + B_STAGE) + EPI_STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int GEMM_SMEM_PAIR = STAGES_PAIR * (A_STAGE + B_STAGE / 2) + EPI_STAGING + 1024 + 256;
static_assert(GEMM_SMEM <= 232448 && GEMM_SMEM_PAIR <= 232448, "exceeds the 227 KB shared memory of one CTA");

struct GemmEpi {
    const float* bias;
    const float* residual;
    long long ldr;
    int res_row_mod;
    void* out;
    long long ldo;
    int out_bf16;
    int act_gelu;
    // LayerNorm folded around the GEMMs (see GemmArgs): producer side ...
    __nv_bfloat16* copy_bf16;     // fp32 path: also store bf16(out) here (the next GEMM's A operand), row stride ldc
    long long ldc;
    float2* stat_out;             // fp32 path: per (row, column group) partial (sum, sum of squares) of out
    int stat_groups;
    // ... and consumer side
    const float2* ln_rows;        // bf16 path: (rstd, -mean * rstd) of this GEMM's A rows (ln_finalize_kernel)
    const float* ln_colsum;       // [n] sum_k B[n, k]
    // batched mode (GemmArgs::batch): tile list = batch x tiles; batch b reads B at K offset b * b_k_off and owns
    // output / residual rows [b * batch_rows, b * batch_rows + M)
    int tiles_per_batch;
    long long batch_rows;
    int b_k_off;
    int act_abs;                  // fp32 path: store |value|
    int conv_pitch;               // > 0: implicit 3x3 convolution over a zero-bordered raster map (GemmArgs::conv_pitch)
    int n_run;                    // run-time-epilogue kernel: accumulator columns the MMA really computes (64, 128, 192 or
                                  // 256): a narrow output does not pay for a 256-wide tile on the tensor pipe
    const __nv_bfloat16* aux;     // bf16 path, backward of a GELU layer: out = acc * gelu'(aux[row, col])
    long long ldaux;
    const float2* shift_rows;     // producer: per-row shift c = -y / x taken from the PREVIOUS LayerNorm rows (GemmArgs::shift_rows)
};

// Exact-erf GELU (nn.GELU default, common.py:26): gelu(x) = x * Phi(x) with Phi from Abramowitz-Stegun 7.1.26
// (|abs err of erf| <= 1.5e-7, far below the bf16 rounding of the output).  Written as
//   gelu(x) = max(x, 0) - |x| * E,   E = 0.5 * poly(t) * t * exp(-z^2),  z = |x| / sqrt(2),  t = 1 / (1 + p z)
// with the log2(e) of the exponential folded into z: 13 instructions, two of them MUFU (rcp, ex2), instead
// of libdevice erff's branchy ~30 -- the lin1 epilogue was the bottleneck of that GEMM (profiles/r1a, r1e).
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    const float zs = ax * 0.8493218002880191f;                 // z * sqrt(log2 e)
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2727374808792225f, zs, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-zs * zs));
    float p = fmaf(t, 0.5307027145f, -0.7265760135f);
    p = fmaf(p, t, 0.7107068705f);
    p = fmaf(p, t, -0.142248368f);
    p = fmaf(p, t, 0.127414796f);
    return fmaf(-ax, p * t * e, fmaxf(x, 0.0f));
}

// The same on a pair (FFMA2 / FMUL2): 9 issue slots per element instead of 14 -- the lin1 epilogue is issue bound
// (profiles/ncu_gemm_r1e: 55 % issue-active with two epilogue warps per scheduler).  The polynomial carries the
// minus sign of "- |x| * E".
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x2) {
    // the scalar MUFU / max ops live inside asm blocks that unpack and re-pack the pair themselves: ptxas then
    // allocates their results straight into the halves of the 64-bit register (no MOVs between the packed ops)
    const uint64_t ax2 = f32x2_abs(x2);
    const uint64_t zs2 = f32x2_mul(ax2, f32x2_pack(0.8493218002880191f, 0.8493218002880191f));
    const uint64_t t2 = f32x2_rcp(f32x2_fma(zs2, f32x2_pack(0.2727374808792225f, 0.2727374808792225f), f32x2_pack(1.0f, 1.0f)));
    const uint64_t e2 = f32x2_ex2_neg(f32x2_mul(zs2, zs2));
    uint64_t p = f32x2_fma(t2, f32x2_pack(-0.5307027145f, -0.5307027145f), f32x2_pack(0.7265760135f, 0.7265760135f));
    p = f32x2_fma(p, t2, f32x2_pack(-0.7107068705f, -0.7107068705f));
    p = f32x2_fma(p, t2, f32x2_pack(0.142248368f, 0.142248368f));
    p = f32x2_fma(p, t2, f32x2_pack(-0.127414796f, -0.127414796f));
    p = f32x2_mul(f32x2_mul(p, t2), e2);
    return f32x2_fma(ax2, p, f32x2_relu(x2));
}

// gelu_erf2 that also returns gelu'(x) = Phi(x) + x phi(x) from the same intermediates (the training forward stores it
// as bf16 for the backward pass: six more instructions here instead of a second erf + exp there).
__device__ __forceinline__ uint64_t gelu_erf2_grad(uint64_t x2, uint64_t& g2) {
    const uint64_t ax2 = f32x2_abs(x2);
    const uint64_t zs2 = f32x2_mul(ax2, f32x2_pack(0.8493218002880191f, 0.8493218002880191f));
    const uint64_t t2 = f32x2_rcp(f32x2_fma(zs2, f32x2_pack(0.2727374808792225f, 0.2727374808792225f), f32x2_pack(1.0f, 1.0f)));
    const uint64_t e2 = f32x2_ex2_neg(f32x2_mul(zs2, zs2));                 // exp(-x^2 / 2)
    uint64_t p = f32x2_fma(t2, f32x2_pack(-0.5307027145f, -0.5307027145f), f32x2_pack(0.7265760135f, 0.7265760135f));
    p = f32x2_fma(p, t2, f32x2_pack(-0.7107068705f, -0.7107068705f));
    p = f32x2_fma(p, t2, f32x2_pack(0.142248368f, 0.142248368f));
    p = f32x2_fma(p, t2, f32x2_pack(-0.127414796f, -0.127414796f));
    p = f32x2_mul(f32x2_mul(p, t2), e2);                                    // -0.5 erfc(|x| / sqrt 2)
    float x0, x1, h0, h1, e0, e1;
    f32x2_unpack(x2, x0, x1);
    f32x2_unpack(p, h0, h1);
    f32x2_unpack(e2, e0, e1);
    const float c0 = x0 >= 0.f ? 1.0f + h0 : -h0, c1 = x1 >= 0.f ? 1.0f + h1 : -h1;      // Phi(x)
    g2 = f32x2_pack(fmaf(x0 * 0.3989422804014327f, e0, c0), fmaf(x1 * 0.3989422804014327f, e1, c1));
    return f32x2_fma(ax2, p, f32x2_relu(x2));
}

// PAIR = false: one CTA per 128 x 256 tile (small problems).
// PAIR = true : a cluster of two CTAs (one TPC) per 256 x 256 tile, tcgen05.mma.cta_group::2.  Each CTA stages its own
//   128 rows of A and HALF of the B tile (128 of the 256 weight rows), so the tensor core's shared-memory operand
//   traffic per CTA drops from 48 KB to 32 KB per k-block -- the 1-CTA kernel keeps the SM's shared-memory pipe
//   94 % busy on the qkv GEMM (profiles/ncu_gemm_r1e) and that, not the tensor pipe, was its ceiling.  The leader
//   CTA (rank 0) issues the MMAs for both; TMA completions of both CTAs land on the leader's "full" barrier,
//   tcgen05.commit multicasts "empty" / "accumulator full" to both CTAs, and the epilogue warps of both CTAs
//   hand the accumulator back by arriving on the leader's "accumulator empty" barrier.
// FEAT selects which optional epilogue / scheduling features are compiled in (the plain proj and lin1 GEMMs have no
// slack in their epilogues: carrying the unused branches cost them 15 % and 3 %, profiles/ncu_gemm_r1l vs r1i):
//   F_PRODUCER  bf16 copy + row statistics (norm1 fold, fp32 path)      F_CONSUMER  LayerNorm applied to the accumulator
//   F_RARE      batched mode, |.| output, row-periodic residual (stem / high-pass GEMMs)
//   F_STATIC    the epilogue type is a compile-time fact: F_BF16 (bf16 output), F_GELU, F_RES (fp32 residual) say which;
//               without it those three are read from GemmEpi at run time (stem / adapter / neck GEMMs)
constexpr int F_PRODUCER = 1, F_CONSUMER = 2, F_RARE = 4, F_STATIC = 8, F_BF16 = 16, F_GELU = 32, F_RES = 64, F_QUICK = 128;
//   F_DGELU     bf16 path: multiply by aux = gelu'(pre-activation) (backward of a GELU layer).  On the bf16 path
//               F_PRODUCER means "also store what that backward needs in copy_bf16": gelu'(pre) for an erf-GELU layer,
//               the pre-activation value otherwise.
constexpr int F_DGELU = 256;
// QuickGELU of CLIP (alpha_clip_rw/model.py:171-173): x * sigmoid(1.702 x); GemmArgs::act_gelu == 2
__device__ __forceinline__ float quick_gelu(float x) {
    return x * __frcp_rn(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * x));
}

// the same on a pair: x / (1 + 2^(-1.702 log2(e) x)), packed multiplies / adds around the two MUFU pairs
__device__ __forceinline__ uint64_t quick_gelu2(uint64_t x) {
    const uint64_t c = f32x2_pack(1.702f * 1.4426950408889634f, 1.702f * 1.4426950408889634f);
    const uint64_t one = f32x2_pack(1.0f, 1.0f);
    return f32x2_mul(x, f32x2_rcp(f32x2_add(one, f32x2_ex2_neg(f32x2_mul(x, c)))));
}

template <bool PAIR, int FEAT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                         const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                         int kblocks1, int kblocks2, int M, int N, int tiles_m, int tiles_n, GemmEpi ep) {
    constexpr int NSTAGE = PAIR ? STAGES_PAIR : STAGES;
    constexpr int B_BYTES = PAIR ? B_STAGE / 2 : B_STAGE;      // bytes of B this CTA stages per k-block
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = base + NSTAGE * A_STAGE;
    const uint32_t staging = sB + NSTAGE * B_BYTES;
    const uint32_t bars = staging + EPI_STAGING;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = kblocks1 + kblocks2;
    const int num_tiles = tiles_m * tiles_n;
    // work unit of this CTA (PAIR: of this cluster) and its stride over the tile list
    const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
    const int first_tile = PAIR ? blockIdx.x >> 1 : blockIdx.x;
    const int tile_stride = PAIR ? gridDim.x >> 1 : gridDim.x;
    constexpr int TILE_M = PAIR ? 2 * BM : BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
        if (kblocks2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), PAIR ? 2 * EPI_WARPS : EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) {
            tmem_alloc_pair(tmem_slot, TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();      // peer barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            // bytes of B one box brings in: the run-time kernel trims the box to the n_run columns it computes
            const uint32_t b_tx = (FEAT & F_RARE) ? static_cast<uint32_t>(ep.n_run) * (PAIR ? BK : 2 * BK) : B_BYTES;
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
                const int bidx = (FEAT & F_RARE) ? tile / ep.tiles_per_batch : 0, tl = tile - bidx * ep.tiles_per_batch;
                const int m0 = (tl / tiles_n) * TILE_M + cta_rank * BM;
                // PAIR: each CTA stages its half of the B rows the instruction reads (n_run / 2 of them)
                const int n0 = (tl % tiles_n) * ((FEAT & F_RARE) ? ep.n_run : BN) + (PAIR ? cta_rank * (((FEAT & F_RARE) ? ep.n_run : BN) / 2) : 0);
                const int bk = (FEAT & F_RARE) ? bidx * ep.b_k_off : 0;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const bool second = kb >= kblocks1;
                    const int kc = (second ? kb - kblocks1 : kb) * BK;
                    // implicit 3x3 convolution: k-block kb is tap (kb / 3, kb % 3); its A tile is the same 64 channels
                    // of the pixel rows shifted by that tap (rows outside the map read as zero)
                    int a_col = kc, a_row = m0;
                    if ((FEAT & F_RARE) && ep.conv_pitch > 0) {
                        a_col = 0;
                        a_row = m0 + (kb / 3 - 1) * ep.conv_pitch + (kb % 3 - 1);
                    }
                    if (PAIR) {
                        // the leader arms its barrier with the bytes of both CTAs; both CTAs' copies complete on it
                        if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * (A_STAGE + b_tx));
                        const uint32_t lead_bar = full_bar(stage) & kPeerBitMask;
                        tma_load_2d_pair(sA + stage * A_STAGE, second ? &tmA2 : &tmA1, lead_bar, a_col, a_row);
                        tma_load_2d_pair(sB + stage * B_BYTES, second ? &tmB2 : &tmB1, lead_bar, kc + bk, n0);
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), A_STAGE + b_tx);
                        tma_load_2d(sA + stage * A_STAGE, second ? &tmA2 : &tmA1, full_bar(stage), a_col, a_row);
                        tma_load_2d(sB + stage * B_BYTES, second ? &tmB2 : &tmB1, full_bar(stage), kc + bk, n0);
                    }
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single lane; PAIR: of the leader CTA only) =====================
        if (cta_rank == 0 && elect_one_sync()) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, (FEAT & F_RARE) ? ep.n_run : BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++iter) {
                const int acc = iter & 1;
                mbar_wait(tempty_bar(acc), (((iter >> 1) & 1) ^ 1u));   // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_k_sw128(sA + stage * A_STAGE);
                    const uint64_t bdesc = umma_desc_k_sw128(sB + stage * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128 B swizzle span: +2 in 16-byte units
                        if (PAIR) umma_bf16_ss_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
                        else umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
                    }
                    if (PAIR) {
                        umma_commit_pair(empty_bar(stage));                // smem slot free in BOTH CTAs
                        if (kb == kblocks - 1) umma_commit_pair(tfull_bar(acc));
                    } else {
                        umma_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
                        if (kb == kblocks - 1) umma_commit(tfull_bar(acc));   // accumulator complete
                    }
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        // TMEM hands each thread one accumulator ROW (lane = row).  Storing that way touches 32 different
        // 128-byte lines per warp instruction and the LSU serialises them (profiles/r1a: tensor pipe 28-58 %
        // active on the K=1280 GEMMs).  Each warp therefore transposes its 32x32 fp32 chunk through a
        // private, XOR-swizzled 4 KB shared-memory tile so that global loads/stores run along rows:
        // 8 lanes x 16 B = one full 128-byte line of the fp32 output / residual per row.
        const bool out_bf16 = (FEAT & F_STATIC) ? (FEAT & F_BF16) != 0 : ep.out_bf16 != 0;
        const bool act_gelu = (FEAT & F_STATIC) ? (FEAT & F_GELU) != 0 : ep.act_gelu != 0;
        const bool act_quick = (FEAT & F_STATIC) ? (FEAT & F_QUICK) != 0 : ep.act_gelu == 2;      // QuickGELU instead of erf
        const bool has_res = (FEAT & F_STATIC) ? (FEAT & F_RES) != 0 : ep.residual != nullptr;
        const int ew = warp - 4;
        const int quad = warp & 3;          // TMEM lane quadrant this warp may touch (warp_id % 4)
        const int half = ew >> 2;           // which half of the accumulator columns
        // the two warps of a lane quadrant split the columns the MMA really computed (n_run <= 256 in the run-time kernel)
        const int epi_cols = (FEAT & F_RARE) ? ep.n_run / 2 : EPI_COLS;
        const int epi_chunks = (FEAT & F_RARE) ? ep.n_run / 64 : EPI_CHUNKS;
        const uint32_t stg = staging + ew * 4096;
        const float* __restrict__ bias = ep.bias;
        int iter = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++iter) {
            const int acc = iter & 1;
            const int bidx = (FEAT & F_RARE) ? tile / ep.tiles_per_batch : 0, tl = tile - bidx * ep.tiles_per_batch;
            const int m0 = (tl / tiles_n) * TILE_M + cta_rank * BM, n0 = (tl % tiles_n) * ((FEAT & F_RARE) ? ep.n_run : BN);
            const int row_base = m0 + quad * 32;
            const long long brow = (FEAT & F_RARE) ? bidx * ep.batch_rows : 0;   // first output row of this batch entry
            // LayerNorm of the A rows, applied to the accumulator (GemmArgs::ln_stat): this lane's 4 rows of the
            // bf16 path.  out = rstd * (acc - mu * colsum) + bias  ==  LN(x) * W^T + b  with gamma folded into W.
            float ln_rs[4] = {1.f, 1.f, 1.f, 1.f}, ln_nm[4] = {0.f, 0.f, 0.f, 0.f};
            if ((FEAT & F_CONSUMER) && ep.ln_rows != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = row_base + 8 * i + (lane >> 2);
                    if (row < M) {
                        const float2 rn = __ldg(ep.ln_rows + row);
                        ln_rs[i] = rn.x;
                        ln_nm[i] = rn.y;
                    }
                }
            }
            float st1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, st2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // producer: the bf16 copy and the statistics are taken of (x - c) with c the row's mean one block earlier, so
            // that neither the bf16 rounding nor E[x^2] - mean^2 see a large common offset of the row
            float shc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if ((FEAT & F_PRODUCER) && !out_bf16 && ep.shift_rows != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = row_base + 4 * i + (lane >> 3);
                    if (row < M) {
                        const float2 rn = __ldg(ep.shift_rows + row);
                        shc[i] = -rn.y / rn.x;
                    }
                }
            }
            mbar_wait(tfull_bar(acc), (iter >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < epi_chunks; ++c) {
                const int col0 = n0 + half * epi_cols + c * 32;
                // fp32 path: issue the chunk's 8 residual loads first, as one batch.  out may alias residual
                // (in-place shortcut add), so loads interleaved with the stores below would be serialised
                // behind them by the compiler and expose a full memory latency per row group.
                // bf16 path: this lane's 8 bias / column-sum values, requested before the TMEM load and the transpose
                // so that their latency is off the critical path
                float4 pb0 = make_float4(0.f, 0.f, 0.f, 0.f), pb1 = pb0, pc0 = pb0, pc1 = pb0;
                if (out_bf16) {
                    const int col = col0 + (lane & 3) * 8;
                    if (col < N) {
                        if (bias) {
                            pb0 = __ldg(reinterpret_cast<const float4*>(bias + col));
                            pb1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                        }
                        if ((FEAT & F_CONSUMER) && ep.ln_rows != nullptr) {
                            pc0 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col));
                            pc1 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col + 4));
                        }
                    }
                }
                float4 rr[8];
                const bool use_res = has_res && !out_bf16;
                if (use_res) {
                    const int col = col0 + (lane & 7) * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + 4 * i + (lane >> 3);
                        rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < M && col < N) {
                            const long long rrow = ((FEAT & F_RARE) && ep.res_row_mod > 0) ? row % ep.res_row_mod : brow + row;
                            rr[i] = *reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + col);
                        }
                    }
                }
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * epi_cols + c * 32, v);
                tmem_ld_wait(v);
                if (c == epi_chunks - 1) {
                    // all of this warp's TMEM reads are done: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(tempty_bar(acc) & kPeerBitMask);   // the leader's barrier
                        else mbar_arrive(tempty_bar(acc));
                    }
                }
                if (col0 >= N || row_base >= M) continue;   // warp-uniform
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    st_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                __syncwarp();
                if (!out_bf16) {
                    const int cq = lane & 7, rsub = lane >> 3;
                    const int col = col0 + cq * 4;
                    const bool col_ok = col < N;           // N % 8 == 0: groups of 4 are all-or-nothing
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias && col_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + col));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + rsub;
                        const int row = row_base + r;
                        float4 x = ld_shared_f4(stg + r * 128 + ((cq ^ (r & 7)) << 4));
                        if (row < M && col_ok) {
                            x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                            if (act_gelu) {
                                if (act_quick) {
                                    x.x = quick_gelu(x.x); x.y = quick_gelu(x.y); x.z = quick_gelu(x.z); x.w = quick_gelu(x.w);
                                } else {
                                    x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
                                }
                            }
                            if (use_res) {
                                x.x += rr[i].x; x.y += rr[i].y; x.z += rr[i].z; x.w += rr[i].w;
                            }
                            if ((FEAT & F_RARE) && ep.act_abs) {
                                x.x = fabsf(x.x); x.y = fabsf(x.y); x.z = fabsf(x.z); x.w = fabsf(x.w);
                            }
                            *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + (brow + row) * ep.ldo + col) = x;
                            if (FEAT & F_PRODUCER) {
                                const float c = shc[i];
                                x.x -= c; x.y -= c; x.z -= c; x.w -= c;      // (out has been stored: x is dead after this)
                            }
                            if ((FEAT & F_PRODUCER) && ep.copy_bf16 != nullptr) {
                                uint2 pk;
                                pk.x = pack_bf16(x.x, x.y);
                                pk.y = pack_bf16(x.z, x.w);
                                *reinterpret_cast<uint2*>(ep.copy_bf16 + static_cast<long long>(row) * ep.ldc + col) = pk;
                            }
                            if ((FEAT & F_PRODUCER) && ep.stat_out != nullptr) {
                                st1[i] += (x.x + x.y) + (x.z + x.w);
                                st2[i] += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
                            }
                        }
                    }
                } else {
                    // bf16 output (qkv, lin1, adapter): 8 columns per lane, packed fp32 pairs all the way
                    const int c8 = lane & 3, rsub = lane >> 2;
                    const int col = col0 + c8 * 8;
                    const bool col_ok = col < N;           // N % 8 == 0
                    const bool use_ln = (FEAT & F_CONSUMER) && ep.ln_rows != nullptr;
                    uint64_t bb[4], cs[4];
                    bb[0] = f32x2_pack(pb0.x, pb0.y); bb[1] = f32x2_pack(pb0.z, pb0.w);
                    bb[2] = f32x2_pack(pb1.x, pb1.y); bb[3] = f32x2_pack(pb1.z, pb1.w);
                    cs[0] = f32x2_pack(pc0.x, pc0.y); cs[1] = f32x2_pack(pc0.z, pc0.w);
                    cs[2] = f32x2_pack(pc1.x, pc1.y); cs[3] = f32x2_pack(pc1.z, pc1.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = 8 * i + rsub;
                        const int row = row_base + r;
                        const float4 x0 = ld_shared_f4(stg + r * 128 + (((2 * c8) ^ (r & 7)) << 4));
                        const float4 x1 = ld_shared_f4(stg + r * 128 + (((2 * c8 + 1) ^ (r & 7)) << 4));
                        if (row < M && col_ok) {
                            uint64_t y[4] = {f32x2_pack(x0.x, x0.y), f32x2_pack(x0.z, x0.w), f32x2_pack(x1.x, x1.y),
                                             f32x2_pack(x1.z, x1.w)};
                            if (use_ln) {
                                const uint64_t rs2 = f32x2_pack(ln_rs[i], ln_rs[i]), nm2 = f32x2_pack(ln_nm[i], ln_nm[i]);
#pragma unroll
                                for (int e = 0; e < 4; ++e) y[e] = f32x2_fma(y[e], rs2, f32x2_fma(cs[e], nm2, bb[e]));
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) y[e] = f32x2_add(y[e], bb[e]);
                            }
                            const bool keep = (FEAT & F_PRODUCER) && ep.copy_bf16 != nullptr;
                            // training forward: what the backward pass needs of this layer goes to copy_bf16 -- gelu'(pre)
                            // for an erf-GELU layer (the frozen trunk needs no weight gradient, so the pre-activation
                            // itself is never read again), the pre-activation value otherwise
                            const bool keep_grad = keep && act_gelu && !act_quick;
                            if (keep && !keep_grad) {
                                uint4 pre;
                                float lo, hi;
                                f32x2_unpack(y[0], lo, hi); pre.x = pack_bf16(lo, hi);
                                f32x2_unpack(y[1], lo, hi); pre.y = pack_bf16(lo, hi);
                                f32x2_unpack(y[2], lo, hi); pre.z = pack_bf16(lo, hi);
                                f32x2_unpack(y[3], lo, hi); pre.w = pack_bf16(lo, hi);
                                *reinterpret_cast<uint4*>(ep.copy_bf16 + (brow + row) * ep.ldc + col) = pre;
                            }
                            if (keep_grad) {
                                uint64_t gq[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) y[e] = gelu_erf2_grad(y[e], gq[e]);
                                uint4 gr;
                                float lo, hi;
                                f32x2_unpack(gq[0], lo, hi); gr.x = pack_bf16(lo, hi);
                                f32x2_unpack(gq[1], lo, hi); gr.y = pack_bf16(lo, hi);
                                f32x2_unpack(gq[2], lo, hi); gr.z = pack_bf16(lo, hi);
                                f32x2_unpack(gq[3], lo, hi); gr.w = pack_bf16(lo, hi);
                                *reinterpret_cast<uint4*>(ep.copy_bf16 + (brow + row) * ep.ldc + col) = gr;
                            } else if (act_gelu) {
                                if (act_quick) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) y[e] = quick_gelu2(y[e]);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) y[e] = gelu_erf2(y[e]);
                                }
                            }
                            if (((FEAT & F_DGELU) || !(FEAT & F_STATIC)) && ep.aux != nullptr) {
                                const uint4 ax = *reinterpret_cast<const uint4*>(ep.aux + (brow + row) * ep.ldaux + col);
                                const uint32_t aw[4] = {ax.x, ax.y, ax.z, ax.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float lo, hi;
                                    f32x2_unpack(y[e], lo, hi);
                                    lo *= __uint_as_float(aw[e] << 16);            // aux holds gelu'(pre), see keep_grad
                                    hi *= __uint_as_float(aw[e] & 0xffff0000u);
                                    y[e] = f32x2_pack(lo, hi);
                                }
                            }
                            if (has_res) {
                                const long long rrow = ((FEAT & F_RARE) && ep.res_row_mod > 0) ? row % ep.res_row_mod : brow + row;
                                const float4 r0 = *reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + col);
                                const float4 r1 = *reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + col + 4);
                                y[0] = f32x2_add(y[0], f32x2_pack(r0.x, r0.y)); y[1] = f32x2_add(y[1], f32x2_pack(r0.z, r0.w));
                                y[2] = f32x2_add(y[2], f32x2_pack(r1.x, r1.y)); y[3] = f32x2_add(y[3], f32x2_pack(r1.z, r1.w));
                            }
                            uint4 pk;
                            float lo, hi;
                            f32x2_unpack(y[0], lo, hi); pk.x = pack_bf16(lo, hi);
                            f32x2_unpack(y[1], lo, hi); pk.y = pack_bf16(lo, hi);
                            f32x2_unpack(y[2], lo, hi); pk.z = pack_bf16(lo, hi);
                            f32x2_unpack(y[3], lo, hi); pk.w = pack_bf16(lo, hi);
                            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + (brow + row) * ep.ldo + col) = pk;
                        }
                    }
                }
                __syncwarp();   // staging tile is rewritten by the next chunk
            }
            if ((FEAT & F_PRODUCER) && ep.stat_out != nullptr && !out_bf16 && row_base < M && n0 + half * EPI_COLS < N) {
                // (sum, sum of squares) of this warp's 32 rows over its EPI_COLS columns: 8 lanes share a row
                const int grp = (n0 + half * EPI_COLS) / EPI_COLS;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a1 = st1[i], a2 = st2[i];
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                    }
                    const int row = row_base + 4 * i + (lane >> 3);
                    if ((lane & 7) == 0 && row < M)
                        ep.stat_out[static_cast<long long>(row) * ep.stat_groups + grp] = make_float2(a1, a2);
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    long long rows, cols, ld;
    int box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        h = h * 1000003u ^ static_cast<size_t>(k.cols);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.box_rows);
        return h;
    }
};

// bf16 row-major [rows, cols] with row stride ld (elements); box = box_rows x 64, 128-byte swizzle.
// Tensor maps only depend on (pointer, shape), so they are memoised: a forward pass re-uses ~10 maps.
static int make_tensor_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld,
                           int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return CVLM_OK;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15))
        return set_error(CVLM_ERR_ARG, "gemm: operand pointer / row stride must be 16-byte aligned");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return CVLM_OK;
}

static int num_sms() { return device_sm_count(); }     // per device (cvlm_internal.cuh)

template <bool PAIR, int FEAT>
static int launch_kernel(int total_tiles, cudaStream_t stream, const CUtensorMap& tA1, const CUtensorMap& tB1,
                         const CUtensorMap& tA2, const CUtensorMap& tB2, int kb1, int kb2, int m, int n, int tiles_m,
                         int tiles_n, const GemmEpi& ep) {
    auto kern = gemm_bf16_tcgen05_kernel<PAIR, FEAT>;
    constexpr int smem = PAIR ? GEMM_SMEM_PAIR : GEMM_SMEM;
    static unsigned char attr_set[kMaxDevices] = {};      // per (instantiation, device): attributes belong to a context
    CVLM_TRY(ensure_dynamic_smem(kern, smem, attr_set));
    if (PAIR) {   // clusters of two CTAs (one TPC): cta_group::2
        const int clusters = total_tiles < num_sms() / 2 ? total_tiles : num_sms() / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CVLM_CUDA(cudaLaunchKernelEx(&cfg, kern, tA1, tB1, tA2, tB2, kb1, kb2, m, n, tiles_m, tiles_n, ep));
    } else {
        const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
        kern<<<grid, GEMM_THREADS, smem, stream>>>(tA1, tB1, tA2, tB2, kb1, kb2, m, n, tiles_m, tiles_n, ep);
    }
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

// The four GEMMs of a transformer block get fully specialised kernels; everything else runs a variant that reads the
// epilogue type at run time.
constexpr int V_QKV = F_STATIC | F_BF16 | F_CONSUMER;           // norm1 + qkv
constexpr int V_PROJ = F_STATIC | F_RES;                        // proj + shortcut
constexpr int V_LIN1 = F_STATIC | F_BF16 | F_GELU;              // lin1 + GELU
constexpr int V_LIN2 = F_STATIC | F_RES | F_PRODUCER;           // lin2 + shortcut, feeds the next norm1
constexpr int V_PLAIN = F_STATIC | F_BF16;                      // bias only, bf16 out (in_proj of the CLIP tower)
constexpr int V_LIN1Q = F_STATIC | F_BF16 | F_GELU | F_QUICK;   // c_fc + QuickGELU of the CLIP tower
constexpr int V_LIN1T = V_LIN1 | F_PRODUCER;                    // lin1 of the training forward: also stores the pre-activation
constexpr int V_DGELU = F_STATIC | F_BF16 | F_DGELU;            // backward of lin2 into the GELU: acc * gelu'(h) -> bf16
constexpr int V_F32 = F_STATIC;                                 // plain fp32 out (dX GEMMs of the backward pass)
constexpr int V_ANY = F_PRODUCER | F_CONSUMER | F_RARE;         // run-time epilogue type, every feature
template <typename... Args>
static int launch_variant(bool pair, int feat, Args&&... args) {
    if (pair) {
        switch (feat) {
            case V_QKV: return launch_kernel<true, V_QKV>(args...);
            case V_PROJ: return launch_kernel<true, V_PROJ>(args...);
            case V_LIN1: return launch_kernel<true, V_LIN1>(args...);
            case V_LIN2: return launch_kernel<true, V_LIN2>(args...);
            case V_PLAIN: return launch_kernel<true, V_PLAIN>(args...);
            case V_LIN1Q: return launch_kernel<true, V_LIN1Q>(args...);
            case V_LIN1T: return launch_kernel<true, V_LIN1T>(args...);
            case V_DGELU: return launch_kernel<true, V_DGELU>(args...);
            case V_F32: return launch_kernel<true, V_F32>(args...);
            default: return launch_kernel<true, V_ANY>(args...);
        }
    }
    switch (feat) {
        case V_QKV: return launch_kernel<false, V_QKV>(args...);
        case V_PROJ: return launch_kernel<false, V_PROJ>(args...);
        case V_LIN1: return launch_kernel<false, V_LIN1>(args...);
        case V_LIN2: return launch_kernel<false, V_LIN2>(args...);
        case V_PLAIN: return launch_kernel<false, V_PLAIN>(args...);
        case V_LIN1Q: return launch_kernel<false, V_LIN1Q>(args...);
        case V_LIN1T: return launch_kernel<false, V_LIN1T>(args...);
        case V_DGELU: return launch_kernel<false, V_DGELU>(args...);
        case V_F32: return launch_kernel<false, V_F32>(args...);
        default: return launch_kernel<false, V_ANY>(args...);
    }
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
    if (!a.a1 || !a.b1 || !a.out) return set_error(CVLM_ERR_ARG, "gemm: null operand");
    if (a.m <= 0 || a.n <= 0 || a.k1 <= 0) return set_error(CVLM_ERR_ARG, "gemm: bad shape m=%d n=%d k=%d", a.m, a.n, a.k1);
    if (a.n % 8) return set_error(CVLM_ERR_ARG, "gemm: n=%d must be a multiple of 8 (pad the output)", a.n);
    if ((a.k2 > 0) != (a.a2 != nullptr && a.b2 != nullptr)) return set_error(CVLM_ERR_ARG, "gemm: inconsistent second K segment");
    if ((reinterpret_cast<uintptr_t>(a.out) & 15) || (a.ldo % 8) || (a.residual && ((reinterpret_cast<uintptr_t>(a.residual) & 15) || (a.ldr % 4))) ||
        (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)))
        return set_error(CVLM_ERR_ARG, "gemm: epilogue pointers / strides must allow 16-byte vectors");
    const int brows = a.b_rows > 0 ? a.b_rows : a.n;
    const int nbatch = a.batch > 0 ? a.batch : 1;
    if (nbatch > 1 && (a.copy_bf16 || a.stat_out || a.ln_rows || a.res_row_mod > 0))
        return set_error(CVLM_ERR_ARG, "gemm: batched mode supports bias / GELU / residual / abs epilogues only");
    if (a.act_abs && a.out_bf16) return set_error(CVLM_ERR_ARG, "gemm: abs needs the fp32 output path");
    // 256 x 256 cluster tiles whenever they fill the machine; cvlm_set_option("gemm_mode", 1 | 2) forces one kernel
    // (the parity tests flip it to cover both kernels on every shape)
    const int mode = option(OPT_GEMM_MODE);
    const int pair_tiles = nbatch * ((a.m + 2 * BM - 1) / (2 * BM)) * ((a.n + BN - 1) / BN);
    const bool pair = mode == 2 ? true : mode == 1 ? false : pair_tiles >= num_sms() / 2;
    // narrow outputs (n <= 192) run the run-time-epilogue kernel with the MMA, the B box and the epilogue trimmed to
    // n_run columns.  The B box matters most: with a 256-row box over a 64-row weight matrix every k-block took 1.1 us
    // (0.5 us with the 128-row boxes of a CTA pair) however little the MMA had to do -- the time followed the number of
    // box rows beyond the tensor, ~11 cycles per zero-filled row (decoder 3x3 convolutions).
    // The 5-D window boxes of the attention kernel show no such cost (tools/win_oob_probe.py).
    int n_run = (a.n > 192 || a.stat_out) ? BN : (a.n + 63) / 64 * 64;
    // few tiles (small m: the batch-1 latency path of the CLIP tower and the decoder): a narrower N tile, run by the same
    // run-time kernel, puts more SMs to work (m = 579, n = 1024: 20 tiles of 256 columns -> 80 of 64)
    if (n_run == BN && !pair && !a.stat_out && !a.ln_rows && nbatch == 1 && mode != 1) {
        const long long tm = (a.m + BM - 1) / BM;
        if (tm * ((a.n + BN - 1) / BN) * 2 <= num_sms()) {
            n_run = 128;
            if (tm * ((a.n + 127) / 128) * 2 <= num_sms()) n_run = 64;
        }
    }
    const int b_box = pair ? n_run / 2 : n_run;
    CUtensorMap tA1, tB1, tA2, tB2;
    if (a.conv_pitch > 0 && (a.k1 != 9 * BK || a.k2 > 0 || nbatch > 1 || a.lda1 != BK || a.ln_rows))
        return set_error(CVLM_ERR_ARG, "gemm: the implicit 3x3 mode needs A = [m, %d] (lda %d), k1 = %d, one K segment, no batching", BK, BK, 9 * BK);
    CVLM_TRY(make_tensor_map(&tA1, a.a1, a.m, a.conv_pitch > 0 ? BK : a.k1, a.lda1, BM));
    // batched mode: batch b reads B columns [b * b_k_off, b * b_k_off + k); with b_k_off == k the map ends exactly
    // where the last batch entry does, so ragged K edges still zero-fill
    if (nbatch > 1 && (a.b_k_off != a.k1 || (a.k2 > 0 && a.k2 != a.k1) || a.k1 % BK))
        return set_error(CVLM_ERR_ARG, "gemm: batched mode needs b_k_off == k1 (== k2), a multiple of %d", BK);
    CVLM_TRY(make_tensor_map(&tB1, a.b1, brows, static_cast<long long>(a.k1) * nbatch, a.ldb1, b_box));
    if (a.k2 > 0) {
        CVLM_TRY(make_tensor_map(&tA2, a.a2, a.m, a.k2, a.lda2, BM));
        CVLM_TRY(make_tensor_map(&tB2, a.b2, brows, static_cast<long long>(a.k2) * nbatch, a.ldb2, b_box));
    } else {
        tA2 = tA1;
        tB2 = tB1;
    }
    if (a.stat_out && a.out_bf16)
        return set_error(CVLM_ERR_ARG, "gemm: row statistics need the fp32 output path");
    if (a.copy_bf16 && ((reinterpret_cast<uintptr_t>(a.copy_bf16) & (a.out_bf16 ? 15 : 7)) || (a.ldc % (a.out_bf16 ? 8 : 4))))
        return set_error(CVLM_ERR_ARG, "gemm: bf16 copy must allow 8-byte (fp32 path) / 16-byte (bf16 path) vectors");
    if (a.aux && (!a.out_bf16 || (reinterpret_cast<uintptr_t>(a.aux) & 15) || (a.ldaux % 8) || a.act_gelu))
        return set_error(CVLM_ERR_ARG, "gemm: the gelu' epilogue needs bf16 output, a 16-byte aligned aux and no activation");
    if (a.copy_bf16 && (a.copy_bf16 == a.a1 || a.copy_bf16 == a.a2))
        return set_error(CVLM_ERR_ARG, "gemm: the bf16 copy must not alias an A operand (other tiles still read it)");
    if (a.stat_out && a.stat_groups != (a.n + EPI_COLS - 1) / EPI_COLS)
        return set_error(CVLM_ERR_ARG, "gemm: stat_groups must be ceil(n / %d)", EPI_COLS);
    if (a.shift_rows && (a.out_bf16 || !a.copy_bf16 || (reinterpret_cast<uintptr_t>(a.shift_rows) & 7)))
        return set_error(CVLM_ERR_ARG, "gemm: shift_rows belongs to the fp32 producer epilogue (bf16 copy + statistics)");
    if (a.ln_rows && (!a.out_bf16 || !a.ln_colsum || (reinterpret_cast<uintptr_t>(a.ln_colsum) & 15) ||
                      (reinterpret_cast<uintptr_t>(a.ln_rows) & 7)))
        return set_error(CVLM_ERR_ARG, "gemm: the LayerNorm epilogue needs bf16 output, column sums and finalized row statistics");
    GemmEpi ep{a.bias, a.residual, a.ldr, a.res_row_mod, a.out, a.ldo, a.out_bf16, a.act_gelu,
               static_cast<__nv_bfloat16*>(a.copy_bf16), a.ldc, reinterpret_cast<float2*>(a.stat_out), a.stat_groups,
               reinterpret_cast<const float2*>(a.ln_rows), a.ln_colsum, 0, a.batch_rows, nbatch > 1 ? a.b_k_off : 0,
               a.act_abs, a.conv_pitch, n_run, static_cast<const __nv_bfloat16*>(a.aux), a.ldaux,
               reinterpret_cast<const float2*>(a.shift_rows)};
    const int kb1 = (a.k1 + BK - 1) / BK, kb2 = a.k2 > 0 ? (a.k2 + BK - 1) / BK : 0;
    const int tiles_n = (a.n + n_run - 1) / n_run;      // n_run == BN except for narrow outputs / few-tile problems
    const int tiles_m = pair ? (a.m + 2 * BM - 1) / (2 * BM) : (a.m + BM - 1) / BM;
    ep.tiles_per_batch = tiles_m * tiles_n;
    const int total_tiles = nbatch * tiles_m * tiles_n;
    // the exact shape of this call's epilogue; launch_variant maps the four block GEMMs to their own kernels
    const bool rare = nbatch > 1 || a.act_abs || a.res_row_mod > 0 || a.conv_pitch > 0;
    const bool producer = a.out_bf16 ? a.copy_bf16 != nullptr : (a.copy_bf16 != nullptr && a.stat_out != nullptr);
    int feat = F_STATIC | (a.out_bf16 ? F_BF16 : 0) | (a.act_gelu ? F_GELU : 0) | (a.act_gelu == 2 ? F_QUICK : 0) | (a.residual ? F_RES : 0) |
               (producer ? F_PRODUCER : 0) | (a.ln_rows ? F_CONSUMER : 0) | (a.aux ? F_DGELU : 0);
    // bias-free GEMMs have static kernels only for the backward shapes (fp32 out, bf16 out, gelu'); the rest read the
    // epilogue type at run time
    const bool no_bias_ok = feat == V_F32 || feat == V_PLAIN || feat == V_DGELU;
    if (rare || (!producer && (a.copy_bf16 || a.stat_out)) || (!a.bias && !no_bias_ok) || n_run < BN) feat = V_ANY;
    return launch_variant(pair, feat, total_tiles, stream, tA1, tB1, tA2, tB2, kb1, kb2, a.m, a.n, nbatch * tiles_m, tiles_n, ep);
}

}  // namespace cvlm
