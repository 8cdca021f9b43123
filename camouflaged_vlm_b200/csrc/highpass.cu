// PromptGenerator.fft (image_encoder.py:332-353): the adapter's high-pass filter, without an FFT.
//
// The reference zeroes the centred low-frequency box of the 2-D spectrum (half-width `line`) and takes |Re ifft2|.
// The box is separable, so the removed part is  LP = P X P^T  with the circulant projector
//     P[n, n'] = (1/N) sum_{u in [-line, line)} exp(2 pi i u (n - n') / N)  =  Pr + i Pi
// (Pi is not zero: the reference's box holds frequency -line but not +line).  Hence
//     high_pass(X) = | X - Pr X Pr^T + Pi X Pi^T |
// which is two passes of the tcgen05 GEMM over bf16 copies -- 0.8 TFLOP per 32 images, against seven memory passes
// over a complex spectrum for the FFT route:
//   cast   Xb = bf16(X)                                   [(image, h), w]
//   G1     Y  = [Pr; Pi] Xb^T   (one GEMM)                [2N, (image, h)]   Y_r = Pr X^T, Y_i = Pi X^T
//   G2     out_image = | X - Pr Y_r^T + Pi Y_i^T |        (batched over images, two K segments, fp32 residual X)
// The matrices are packed on the host (packing.py: hp.pcat = [Pr; Pi] [2n, n]; hp.npr = -Pr, hp.pi = Pi [n, np]).
#include "cvlm_internal.cuh"
#include "ptx.cuh"

namespace cvlm {

// Images whose side is not a multiple of the GEMM's 64-wide k-block: the contraction over h in G2 runs per image
// inside one long K axis, so every image's block of Y columns is padded to np = ceil(n / 64) * 64 with zeros
// (zero rows of Xb -> zero columns of Y; -Pr / Pi are packed with np columns as well).
static long long pad64(long long n) { return (n + 63) / 64 * 64; }

long long high_pass_scratch_bytes(int batch, int chans, int img_size) {
    const long long ni = static_cast<long long>(batch) * chans, n = img_size, np = pad64(n);
    return ni * np * n * 2 /* Xb */ + 2 * n * ni * np * 2 /* Y */;
}

// fp32 [ni, n, n] -> bf16 [ni, np, n], rows >= n zero.  One thread converts 8 pixels.
__global__ void __launch_bounds__(256) cast_pad_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int n,
                                                            int np, long long total8) {
    const int n8 = n / 8;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % n8);
        const long long r = i / n8;
        const int h = static_cast<int>(r % np);
        const long long img = r / np;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (h < n) {
            const float4* src = reinterpret_cast<const float4*>(x + (img * n + h) * n + c8 * 8);
            const float4 a = __ldcs(src), b = __ldcs(src + 1);
            v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
        }
        *reinterpret_cast<uint4*>(out + (img * np + h) * n + c8 * 8) = v;
    }
}

int high_pass_launch(const float* images, float* out, int batch, int chans, int img_size, const void* pcat,
                     const void* npr, const void* pi, void* scratch, long long scratch_bytes, cudaStream_t stream) {
    if (!images || !out || !pcat || !npr || !pi || !scratch) return set_error(CVLM_ERR_ARG, "high_pass: null argument");
    if (batch <= 0 || chans <= 0 || img_size <= 0 || img_size % 8)
        return set_error(CVLM_ERR_UNSUPPORTED, "high_pass: img_size must be a multiple of 8");
    if (scratch_bytes < high_pass_scratch_bytes(batch, chans, img_size))
        return set_error(CVLM_ERR_WORKSPACE, "high_pass: scratch has %lld bytes, need %lld", scratch_bytes,
                         high_pass_scratch_bytes(batch, chans, img_size));
    const long long ni = static_cast<long long>(batch) * chans, n = img_size, np = pad64(n);
    if (ni * np > 0x7fffffffLL) return set_error(CVLM_ERR_UNSUPPORTED, "high_pass: batch too large for one call");
    __nv_bfloat16* xb = static_cast<__nv_bfloat16*>(scratch);
    __nv_bfloat16* y = xb + ni * np * n;
    {
        const long long total8 = ni * np * (n / 8);
        const unsigned grid = static_cast<unsigned>((total8 + 255) / 256 < 148 * 32 ? (total8 + 255) / 256 : 148 * 32);
        cast_pad_rows_kernel<<<grid, 256, 0, stream>>>(images, xb, static_cast<int>(n), static_cast<int>(np), total8);
        CVLM_CUDA(cudaGetLastError());
    }
    {   // Y[m, (image, h)] = sum_w Pcat[m, w] X[image, h, w]
        GemmArgs ga;
        ga.a1 = pcat; ga.lda1 = n; ga.b1 = xb; ga.ldb1 = n; ga.k1 = static_cast<int>(n);
        ga.m = static_cast<int>(2 * n); ga.n = static_cast<int>(ni * np);
        ga.out = y; ga.ldo = ni * np; ga.out_bf16 = 1;
        CVLM_TRY(gemm_launch(ga, stream));
    }
    {   // out[image, h', w'] = | X[image, h', w'] - sum_h Pr[h', h] Y_r[w', (image, h)] + sum_h Pi[h', h] Y_i[w', (image, h)] |
        GemmArgs ga;
        ga.a1 = npr; ga.lda1 = np; ga.b1 = y; ga.ldb1 = ni * np; ga.k1 = static_cast<int>(np);
        ga.a2 = pi; ga.lda2 = np; ga.b2 = y + n * ni * np; ga.ldb2 = ni * np; ga.k2 = static_cast<int>(np);
        ga.m = static_cast<int>(n); ga.n = static_cast<int>(n);
        ga.batch = static_cast<int>(ni); ga.b_k_off = static_cast<int>(np); ga.batch_rows = n;
        ga.residual = images; ga.ldr = n; ga.out = out; ga.ldo = n; ga.act_abs = 1;
        CVLM_TRY(gemm_launch(ga, stream));
    }
    return CVLM_OK;
}

}  // namespace cvlm
