// Shared by the encoder runtime (encoder.cu: inference schedule, C ABI) and the training schedule (encoder_train.cu).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "cvlm_internal.cuh"

namespace cvlm {

struct TensorRef {
    const void* ptr;
    long long nbytes;
};

}  // namespace cvlm

struct ProfSpan {
    int kind;
    cudaEvent_t start, stop;
};

struct cvlm_encoder {
    cvlm_encoder_config cfg;
    std::unordered_map<std::string, cvlm::TensorRef> tensors;
    long long last_launches = 0;
    bool profiling = false;
    std::vector<ProfSpan> spans;        // spans of the profiled forwards since the last get_profile
    std::vector<cudaEvent_t> event_pool;
    std::vector<cudaEvent_t> host_events;   // cvlm_encoder_forward_host: staging-buffer ready / consumed, stream edges
    // training forward: block i's attention kernel left the row log-sum-exp in the saved region (the tcgen05 kernels do)
    unsigned char lse_saved[CVLM_MAX_DEPTH] = {};
    unsigned char tabs_saved[CVLM_MAX_DEPTH] = {};   // ... and (global blocks) the per-query rel-pos terms for the dK / dV kernel
};

namespace cvlm {

enum ProfKind {
    K_GEMM_QKV = 0, K_GEMM_PROJ, K_GEMM_LIN1, K_GEMM_LIN2, K_GEMM_OTHER, K_ATTN_WINDOW, K_ATTN_GLOBAL, K_LAYERNORM,
    K_OTHER, K_COUNT
};
static_assert(K_COUNT == CVLM_PROF_KINDS, "cvlm.h CVLM_PROF_KINDS out of sync");

static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// layout of the saved-activation region of the training forward (encoder.cu: plan_saved)
struct Saved {
    long long off_x0b, off_php, off_t, off_gall, off_gpre, off_n1, off_n2, off_blocks, block_stride;
    long long b_xb, b_rows1, b_qkv, b_o, b_xh2, b_rstd2, b_hpre, b_lse;     // offsets inside one block's slot
    // per GLOBAL block (index among the global blocks): tw | th tables, fp32 [M, heads, 64] each (attention_tc2.cu)
    long long off_gtab, gtab_stride, gtab_half;
    long long total;
};
Saved plan_saved(const cvlm_encoder_config& c, int batch);
int forward_impl(cvlm_encoder* enc, const float* images, const float* highpass, int batch, float* out,
                 float* const* interm, int stop_after_block, float* x_out, void* workspace, long long workspace_bytes,
                 void* saved, long long saved_bytes, cudaStream_t stream);

struct Fwd {
    cvlm_encoder* enc;
    cudaStream_t stream;
    long long launches = 0;
    char namebuf[96];

    int get(const char* name, long long expect_bytes, const void** out) {
        auto it = enc->tensors.find(name);
        if (it == enc->tensors.end()) return set_error(CVLM_ERR_MISSING, "tensor '%s' was not registered", name);
        if (expect_bytes > 0 && it->second.nbytes != expect_bytes)
            return set_error(CVLM_ERR_MISSING, "tensor '%s' has %lld bytes, expected %lld", name, it->second.nbytes,
                             expect_bytes);
        *out = it->second.ptr;
        return CVLM_OK;
    }
    int getb(int blk, const char* suffix, long long expect_bytes, const void** out) {
        snprintf(namebuf, sizeof(namebuf), "blk.%d.%s", blk, suffix);
        return get(namebuf, expect_bytes, out);
    }
    cudaEvent_t event() {
        if (!enc->event_pool.empty()) {
            cudaEvent_t e = enc->event_pool.back();
            enc->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    // every kernel launch of the forward goes through here: counts it and, when profiling is on,
    // brackets it with CUDA events on the launching stream
    template <typename F>
    int run(int kind, F&& launch) {
        ++launches;
        if (!enc->profiling) return launch();
        ProfSpan sp{kind, event(), event()};
        cudaEventRecord(sp.start, stream);
        const int rc = launch();
        cudaEventRecord(sp.stop, stream);
        enc->spans.push_back(sp);
        return rc;
    }
    int gemm(int kind, const GemmArgs& a) {
        return run(kind, [&] { return gemm_launch(a, stream); });
    }
};

}  // namespace cvlm
