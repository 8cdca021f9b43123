// The reference's image input pipeline on the GPU, bit-exact:
//   transforms.Resize((S, S)) on a PIL image  ->  ToTensor  ->  Normalize     (datasets/wrappers.py:22-27, 61)
// uint8 HWC image of any size in, float32 [3, S, S] out, ready for ImageEncoderViT.forward.
//
// transforms.Resize on a PIL image is Pillow's 8-bit antialiased resample (src/libImaging/Resample.c): two separable
// passes, horizontal first, each rounded to uint8.  Per output coordinate the taps and their weights are computed in
// double precision, rounded to 22-bit fixed point, and the pixel is clip((2^21 + sum(pixel * kk)) >> 22, 0, 255).
// The coefficient tables are computed on the host with the same double-precision expression order as Pillow
// (oracle/preprocess_oracle.py restates it and is pinned against PIL itself); the two passes are integer kernels, so the
// result is identical to the reference's bytes; ToTensor / Normalize are three IEEE fp32 operations per value
// (x / 255, - mean, / std) written with round-to-nearest intrinsics so nothing is contracted or approximated.
#include <math.h>

#include <string.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "cvlm_internal.cuh"

namespace cvlm {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

struct PassTable {
    int out = 0, ksize = 0;
    std::vector<int> bounds;   // [out][2] = (first tap, tap count)
    std::vector<int> kk;       // [out][ksize] fixed-point weights
};

// precompute_coeffs + normalize_coeffs_8bpc (Resample.c), bilinear filter (support 1), box = the whole axis
static PassTable make_table(int in_size, int out_size) {
    PassTable t;
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    t.out = out_size;
    t.ksize = static_cast<int>(ceil(support)) * 2 + 1;
    t.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
    t.kk.assign(static_cast<size_t>(out_size) * t.ksize, 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> w(t.ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double v = (x + xmin - center + 0.5) * ss;
            if (v < 0.0) v = -v;
            w[x] = v < 1.0 ? 1.0 - v : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            const double k = ww != 0.0 ? w[x] / ww : w[x];
            t.kk[static_cast<size_t>(xx) * t.ksize + x] =
                k < 0 ? static_cast<int>(-0.5 + k * (1 << PP_PRECISION_BITS)) : static_cast<int>(0.5 + k * (1 << PP_PRECISION_BITS));
        }
        t.bounds[2 * xx] = xmin;
        t.bounds[2 * xx + 1] = xmax;
    }
    return t;
}

__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> PP_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: u8 [h, w, 3] -> u8 [h, ow, 3]; one thread per output pixel
__global__ void __launch_bounds__(256) pp_horizontal_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                            int h, int w, int ow) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(h) * ow) return;
    const int x = static_cast<int>(i % ow), y = static_cast<int>(i / ow);
    const int xmin = bounds[2 * x], cnt = bounds[2 * x + 1];
    const int* k = kk + static_cast<long long>(x) * ksize;
    const unsigned char* row = in + (static_cast<long long>(y) * w + xmin) * 3;
    int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < cnt; ++t) {
        const int c = __ldg(k + t);
        a0 += row[3 * t] * c;
        a1 += row[3 * t + 1] * c;
        a2 += row[3 * t + 2] * c;
    }
    unsigned char* o = out + i * 3;
    o[0] = static_cast<unsigned char>(clip8(a0));
    o[1] = static_cast<unsigned char>(clip8(a1));
    o[2] = static_cast<unsigned char>(clip8(a2));
}

// vertical pass (or identity when the height already matches) + ToTensor + Normalize: u8 [h, ow, 3] -> fp32 [3, oh, ow]
__global__ void __launch_bounds__(256) pp_vertical_norm_kernel(const unsigned char* __restrict__ in, float* __restrict__ out,
                                                               const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                               int identity, int oh, int ow, float m0, float m1, float m2, float s0,
                                                               float s1, float s2) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(oh) * ow) return;
    const int x = static_cast<int>(i % ow), y = static_cast<int>(i / ow);
    int v0, v1, v2;
    if (identity) {
        const unsigned char* p = in + (static_cast<long long>(y) * ow + x) * 3;
        v0 = p[0]; v1 = p[1]; v2 = p[2];
    } else {
        const int ymin = bounds[2 * y], cnt = bounds[2 * y + 1];
        const int* k = kk + static_cast<long long>(y) * ksize;
        int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < cnt; ++t) {
            const int c = __ldg(k + t);
            const unsigned char* p = in + (static_cast<long long>(ymin + t) * ow + x) * 3;
            a0 += p[0] * c;
            a1 += p[1] * c;
            a2 += p[2] * c;
        }
        v0 = clip8(a0); v1 = clip8(a1); v2 = clip8(a2);
    }
    const long long plane = static_cast<long long>(oh) * ow;
    out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v0), 255.0f), m0), s0);
    out[plane + i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v1), 255.0f), m1), s1);
    out[2 * plane + i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v2), 255.0f), m2), s2);
}

static long long align16(long long v) { return (v + 15) / 16 * 16; }

static int ksize_for(int in_size, int out_size) {
    const double scale = static_cast<double>(in_size) / out_size;
    return static_cast<int>(ceil(scale < 1.0 ? 1.0 : scale)) * 2 + 1;
}

// One pass table in pinned host memory, built once per (input length, output size) and never freed (a few KB to a few MB
// per distinct image size; bounded to 256 entries, after which the oldest is kept and new sizes fall back to rebuilding
// into the same slot under the lock -- sizes seen in practice are a handful).
struct PinnedTable {
    int* data = nullptr;      // bounds [2 * out] followed by kk [out * ksize]
    long long n_bounds = 0, n_kk = 0;
    int ksize = 0;
};
static const PinnedTable* pinned_pass_table(int in_size, int out_size) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, PinnedTable> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(in_size, out_size);
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    if (cache.size() >= 256) return nullptr;
    const PassTable t = make_table(in_size, out_size);
    PinnedTable pt;
    pt.n_bounds = static_cast<long long>(t.bounds.size());
    pt.n_kk = static_cast<long long>(t.kk.size());
    pt.ksize = t.ksize;
    if (cudaHostAlloc(reinterpret_cast<void**>(&pt.data), static_cast<size_t>(pt.n_bounds + pt.n_kk) * sizeof(int), cudaHostAllocPortable) !=
        cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    memcpy(pt.data, t.bounds.data(), t.bounds.size() * sizeof(int));
    memcpy(pt.data + pt.n_bounds, t.kk.data(), t.kk.size() * sizeof(int));
    return &cache.emplace(key, pt).first->second;
}

long long preprocess_scratch_bytes(int h, int w, int out_size) {
    const long long tables = (2LL * out_size + static_cast<long long>(out_size) * ksize_for(w, out_size)) * 4 +
                             (2LL * out_size + static_cast<long long>(out_size) * ksize_for(h, out_size)) * 4;
    return align16(static_cast<long long>(h) * out_size * 3) + align16(tables) + 64;
}

int preprocess_u8_launch(const unsigned char* img, int h, int w, float* out, int out_size, const float* mean, const float* stdv,
                         void* scratch, long long scratch_bytes, cudaStream_t stream) {
    if (!img || !out || !mean || !stdv || !scratch) return set_error(CVLM_ERR_ARG, "preprocess: null argument");
    if (h <= 0 || w <= 0 || out_size <= 0) return set_error(CVLM_ERR_ARG, "preprocess: bad size %dx%d -> %d", h, w, out_size);
    if (scratch_bytes < preprocess_scratch_bytes(h, w, out_size))
        return set_error(CVLM_ERR_WORKSPACE, "preprocess: scratch has %lld bytes, need %lld", scratch_bytes,
                         preprocess_scratch_bytes(h, w, out_size));
    if (reinterpret_cast<uintptr_t>(scratch) & 15) return set_error(CVLM_ERR_ARG, "preprocess: scratch must be 16-byte aligned");
    unsigned char* mid = static_cast<unsigned char*>(scratch);                      // horizontally resampled image
    int* tab = reinterpret_cast<int*>(mid + align16(static_cast<long long>(h) * out_size * 3));
    const bool need_h = w != out_size, need_v = h != out_size;
    // Coefficient tables depend on (input length, output size) only: built once per size and kept in PINNED host memory
    // (pinned_pass_table), so the upload is a true asynchronous copy on `stream` -- no host serialisation, capturable in a
    // CUDA graph -- and nothing is rebuilt per image.
    long long off_hb = 0, off_hk = 0, off_vb = 0, off_vk = 0, used = 0;
    int ks_h = 0, ks_v = 0;
    if (need_h) {
        const PinnedTable* t = pinned_pass_table(w, out_size);
        if (!t) return set_error(CVLM_ERR_CUDA, "preprocess: could not allocate the pinned coefficient table");
        ks_h = t->ksize;
        off_hb = used;
        off_hk = used + t->n_bounds;
        CVLM_CUDA(cudaMemcpyAsync(tab + used, t->data, static_cast<size_t>(t->n_bounds + t->n_kk) * sizeof(int), cudaMemcpyHostToDevice, stream));
        used += t->n_bounds + t->n_kk;
    }
    if (need_v) {
        const PinnedTable* t = pinned_pass_table(h, out_size);
        if (!t) return set_error(CVLM_ERR_CUDA, "preprocess: could not allocate the pinned coefficient table");
        ks_v = t->ksize;
        off_vb = used;
        off_vk = used + t->n_bounds;
        CVLM_CUDA(cudaMemcpyAsync(tab + used, t->data, static_cast<size_t>(t->n_bounds + t->n_kk) * sizeof(int), cudaMemcpyHostToDevice, stream));
        used += t->n_bounds + t->n_kk;
    }
    const unsigned char* vin = img;
    if (need_h) {
        const long long n = static_cast<long long>(h) * out_size;
        pp_horizontal_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(img, mid, tab + off_hb, tab + off_hk, ks_h, h,
                                                                                        w, out_size);
        CVLM_CUDA(cudaGetLastError());
        vin = mid;
    }
    const long long n = static_cast<long long>(out_size) * out_size;
    pp_vertical_norm_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        vin, out, tab + off_vb, tab + off_vk, ks_v, need_v ? 0 : 1, out_size, out_size, mean[0], mean[1], mean[2], stdv[0], stdv[1],
        stdv[2]);
    CVLM_CUDA(cudaGetLastError());
    return CVLM_OK;
}

}  // namespace cvlm
