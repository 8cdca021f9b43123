// Internal declarations shared by the kernels, the op-level C-ABI and the encoder runtime.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cvlm.h"

namespace cvlm {

int set_error(int code, const char* fmt, ...);

#define CVLM_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::cvlm::set_error(CVLM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                     __FILE__, __LINE__);                                           \
    } while (0)

#define CVLM_TRY(expr)              \
    do {                            \
        int _rc = (expr);           \
        if (_rc != CVLM_OK) return _rc; \
    } while (0)

// ---- per-device facts and test/tuning options (encoder.cu).  The library keeps no per-process device state: the SM
// count and the "opt-in shared memory size already set" flags are per DEVICE (function attributes belong to a device's
// context), so one process may drive several GPUs.
constexpr int kMaxDevices = 64;
int device_sm_count();                      // SMs of the CURRENT device (cached per device)
// one-time cudaFuncSetAttribute(MaxDynamicSharedMemorySize) per (call site, device); `flags` is the call site's own
// static array of kMaxDevices zero-initialised bytes
template <typename Kern>
int ensure_dynamic_smem(Kern kern, int bytes, unsigned char* flags) {
    int dev = 0;
    CVLM_CUDA(cudaGetDevice(&dev));
    const int di = dev >= 0 && dev < kMaxDevices ? dev : 0;
    if (__atomic_load_n(&flags[di], __ATOMIC_ACQUIRE) && dev < kMaxDevices) return CVLM_OK;
    CVLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    __atomic_store_n(&flags[di], static_cast<unsigned char>(1), __ATOMIC_RELEASE);
    return CVLM_OK;
}
// cvlm_set_option (include/cvlm.h): read on the launch path as one relaxed atomic load, never getenv
enum Option { OPT_GEMM_MODE = 0, OPT_DEC_CLUSTER, OPT_ATTN_LEGACY, OPT_ATTN_POLY, OPT_WIN_MODE, OPT_BWD_TC, OPT_COUNT };
int option(Option o);

// ---- GEMM (gemm_tcgen05.cu):  out = epilogue(A1 * B1^T [+ A2 * B2^T])
struct GemmArgs {
    const void* a1 = nullptr;  long long lda1 = 0;   // bf16 [M, K1], row stride lda1 elements
    const void* b1 = nullptr;  long long ldb1 = 0;   // bf16 [Nb, K1] (Nb <= N rows actually stored)
    int k1 = 0;
    const void* a2 = nullptr;  long long lda2 = 0;   // optional second K segment
    const void* b2 = nullptr;  long long ldb2 = 0;
    int k2 = 0;
    int m = 0, n = 0;          // logical output shape
    int b_rows = 0;            // rows present in B1/B2 (0 -> n); rows beyond read as zero
    const float* bias = nullptr;       // [n] or null
    const float* residual = nullptr;   // fp32 [*, ldr] or null
    long long ldr = 0;
    int res_row_mod = 0;       // >0: residual row = row % res_row_mod (pos_embed broadcast)
    void* out = nullptr;       // fp32 or bf16 [M, ldo]
    long long ldo = 0;
    int out_bf16 = 0;
    int act_gelu = 0;          // 1: exact-erf GELU (nn.GELU), 2: QuickGELU x * sigmoid(1.702 x) (CLIP)
    // LayerNorm folded around the GEMMs.  nn.LayerNorm (image_encoder.py:430-446 norm1 / norm2) sits between a GEMM
    // that writes the fp32 residual stream and a GEMM that reads its normalised bf16 copy.  Instead of a separate
    // pass over x, the PRODUCER (fp32 path) also stores bf16(x) and, per row and column group, (sum x, sum x^2);
    // ln_finalize turns those into (rstd, -mean * rstd) per row; the CONSUMER runs on the raw bf16(x) with gamma
    // folded into its weights and applies
    //     out = rstd * (acc - mean * colsum) + bias,   colsum[n] = sum_k W'[n, k],  bias = W beta + b
    // in its epilogue.  Each partial has exactly one writer, so results do not depend on scheduling.
    void* copy_bf16 = nullptr;  long long ldc = 0;       // producer: bf16 copy of out (must not alias a1 / a2)
    float* stat_out = nullptr;  int stat_groups = 0;     // producer: fp32 [m, stat_groups, 2], stat_groups = ceil(n / CVLM_STAT_GROUP_COLS)
    // producer, optional: fp32 [m, 2] = the (rstd, -mean * rstd) rows of the PREVIOUS LayerNorm input of the same rows.  The
    // copy and the statistics are then taken of (x - c), c = that earlier mean: LN(x - c) == LN(x), but a row whose
    // |mean| is large against its spread loses nothing to the bf16 rounding of the copy or to E[x^2] - mean^2.
    const float* shift_rows = nullptr;
    const float* ln_rows = nullptr;                      // consumer: fp32 [m, 2] = (rstd, -mean * rstd) of the A rows
    const float* ln_colsum = nullptr;                    // consumer: fp32 [n]
    // Batched mode: `batch` independent products that share A1 / A2.  Entry b multiplies by the B columns
    // [b * b_k_off, b * b_k_off + k) and owns output / residual rows [b * batch_rows, b * batch_rows + m).
    int batch = 1;
    int b_k_off = 0;
    long long batch_rows = 0;
    int act_abs = 0;           // fp32 path: store |value|
    // Implicit 3x3 convolution (stride 1, padding 1) over an NHWC map with 64 channels stored with a one-pixel ZERO
    // border: A = bf16 [m, 64] holds the rows of images of conv_pitch x conv_pitch padded pixels, B = [n, 9 * 64] with
    // column (ty*3+tx)*64 + c.  K-block t of output row r reads A row r + (t/3 - 1) * conv_pitch + (t%3 - 1) (zero outside
    // [0, m)), so no im2col matrix is ever written.  Output rows at border pixels are meaningless; the caller zeroes or
    // ignores them.  k1 must be 576.
    int conv_pitch = 0;
    // Training (adapter-only backward).  bf16 path: `copy_bf16` also receives the PRE-activation value
    // (what the backward of the GELU needs); `aux` (bf16 [m, ldaux]) makes the epilogue multiply by gelu'(aux): the
    // backward of lin2 / shared_mlp into the GELU that precedes them.
    const void* aux = nullptr;  long long ldaux = 0;
};
int gemm_launch(const GemmArgs& a, cudaStream_t stream);

// ---- attention (attention.cu)
// lse_out (optional, training forward): fp32 [batch * grid^2, heads], every token's row log-sum-exp in the log2 domain;
// *lse_written tells whether the kernel that ran produces it (the tcgen05 kernels do, the mma.sync fallback does not).
// tw_out / th_out (optional, 64 x 64 global blocks only; *tabs_written): fp32 [batch * grid^2, heads, 64] each, every
// query's rel-pos terms laid out for the dK / dV kernel (BwdParams::tw_tab / th_tab).
int attention_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int grid_side,
                     int window, int heads, int head_dim, cudaStream_t stream, float* lse_out = nullptr,
                     bool* lse_written = nullptr, float* tw_out = nullptr, float* th_out = nullptr,
                     bool* tabs_written = nullptr);

int attention_global_tc2_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int heads,
                                int head_dim, cudaStream_t stream, float* lse_out = nullptr, float* tw_out = nullptr,
                                float* th_out = nullptr);

int attention_plain_tc2_launch(const void* qkv, void* out, int batch, int tokens, int heads, int head_dim,
                               cudaStream_t stream);

int attention_window_tc2_launch(const void* qkv, void* out, const void* relh, const void* relw, int batch, int grid_side,
                                int heads, int head_dim, cudaStream_t stream, float* lse_out = nullptr);

// ---- backward pass for adapter-only training (attention_bwd.cu, backward.cu)
// lse_ready: `lse` already holds the forward's row log-sum-exp (attention_launch's lse_out); the kernels then skip their
// own sweep over S that would compute it.
int attention_backward_launch(const void* qkv, const void* o, const void* dout, void* dqkv, float* lse, float* dsum,
                              const void* relh, const void* relw, int batch, int grid_side, int window, int heads,
                              int head_dim, cudaStream_t stream, bool lse_ready = false, const float* tw_tab = nullptr,
                              const float* th_tab = nullptr);
// dK / dV of a 64 x 64 global block on tcgen05 (attention_bwd_tc.cu); dsum must hold D = rowsum(dO o) (the dQ kernel writes it)
int attention_bwd_dkv_tc_launch(const void* qkv, const void* dout, void* dqkv, const float* dsum, const float* tw_tab,
                                const float* th_tab, int batch, int heads, int head_dim, cudaStream_t stream);
int attention_bwd_dq_tc_launch(const void* qkv, const void* o, const void* dout, void* dqkv, float* dsum, const float* tw_tab,
                               const float* th_tab, const void* relh, const void* relw, int batch, int heads, int head_dim,
                               cudaStream_t stream);
int layernorm_train_launch(const float* x, const float* gamma, const float* beta, void* out_bf16, void* xhat_bf16,
                           float* rstd, long long rows, int cols, float eps, cudaStream_t stream);
int ln_bwd_launch(int xhat_kind, const float* g, const float* gamma, const void* xsrc, const float* stat,
                  const float* residual, float* dx, void* dxb, long long rows, int cols, float eps, cudaStream_t stream);
int wgrad_splits(long long m);
int wgrad_jp(int j, int ones);
int wgrad_tn_launch(const void* A, long long lda, int I, const void* B, long long ldb, int J, int ones, long long M,
                    float* partial, int accumulate, cudaStream_t stream);
int wgrad_reduce_launch(const float* partial, int splits, long long n, float* out, cudaStream_t stream);
int add_f32_launch(float* y, const float* x, long long n, cudaStream_t stream);
int eg_wgrad_launch(const void* x0b, const void* dt, int batch, int tokens, int dim, int rank_pad, float* partial,
                    cudaStream_t stream);

// ---- row kernels (rowops.cu)
int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, int out_bf16, long long rows,
                     int cols, float eps, cudaStream_t stream);
int ln_finalize_launch(const float* partial, int groups, int width, float eps, float* rows_out, long long rows,
                       cudaStream_t stream);
int cast_bf16_launch(const float* x, void* out, long long n, cudaStream_t stream);
// ---- image input pipeline (preprocess.cu)
long long preprocess_scratch_bytes(int h, int w, int out_size);
int preprocess_u8_launch(const unsigned char* img, int h, int w, float* out, int out_size, const float* mean, const float* stdv,
                         void* scratch, long long scratch_bytes, cudaStream_t stream);
// ---- high-pass filter of the adapter (highpass.cu)
long long high_pass_scratch_bytes(int batch, int chans, int img_size);
int high_pass_launch(const float* images, float* out, int batch, int chans, int img_size, const void* pcat,
                     const void* npr, const void* pi, void* scratch, long long scratch_bytes, cudaStream_t stream);
int im2col_patch_launch(const float* img, void* out, int batch, int chans, int img_size, int patch,
                        cudaStream_t stream);
int im2col_3x3_launch(const void* x_nhwc_bf16, void* out, int batch, int grid_side, int chans, cudaStream_t stream);
int layernorm_nchw_out_launch(const float* x, const float* gamma, const float* beta, float* out_nchw, int batch,
                              int tokens, int chans, float eps, cudaStream_t stream);
int adapter_embed_launch(const float* x0, const float* hand, const float* w_t, const float* cst, void* t_out,
                         int batch, int tokens, int dim, int rank, int rank_pad, cudaStream_t stream);

}  // namespace cvlm
