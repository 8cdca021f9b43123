// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry
// point, so the library does not link against libcuda).  Maps depend only on (pointer, shape, box), so
// they are memoised; a forward pass re-uses a few dozen of them.
#pragma once
#include <mutex>
#include <string.h>
#include <unordered_map>

#include "cvlm_internal.cuh"

namespace cvlm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
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

struct MapSpec {
    const void* ptr;
    int rank;
    unsigned long long dims[5];      // elements, innermost first
    unsigned long long strides[4];   // bytes, for dims 1..rank-1
    unsigned box[5];
    bool operator==(const MapSpec& o) const { return memcmp(this, &o, sizeof(MapSpec)) == 0; }
};
struct MapSpecHash {
    size_t operator()(const MapSpec& k) const {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
        size_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(MapSpec); ++i) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
};

// bf16 tensor, 128-byte swizzle, zero fill out of bounds.
inline int make_tensor_map_bf16(CUtensorMap* out, MapSpec spec) {
    static std::mutex mu;
    static std::unordered_map<MapSpec, CUtensorMap, MapSpecHash> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(spec);
        if (it != cache.end()) {
            *out = it->second;
            return CVLM_OK;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (reinterpret_cast<uintptr_t>(spec.ptr) & 15) return set_error(CVLM_ERR_ARG, "tensor map: base not 16-byte aligned");
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    if (spec.rank < 1 || spec.rank > 5) return set_error(CVLM_ERR_ARG, "tensor map: rank %d", spec.rank);
    for (int i = 0; i < spec.rank; ++i) {
        dims[i] = spec.dims[i];
        box[i] = spec.box[i];
    }
    for (int i = 0; i + 1 < spec.rank; ++i) {
        if (spec.strides[i] & 15) return set_error(CVLM_ERR_ARG, "tensor map: stride not a multiple of 16 bytes");
        strides[i] = spec.strides[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, spec.rank, const_cast<void*>(spec.ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 8192) cache.clear();
    cache[spec] = *out;
    return CVLM_OK;
}

// fp32 tensor, no swizzle (rows land in shared memory as they are in global memory), zero fill out of bounds.
inline int make_tensor_map_f32_plain(CUtensorMap* out, MapSpec spec) {
    static std::mutex mu;
    static std::unordered_map<MapSpec, CUtensorMap, MapSpecHash> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(spec);
        if (it != cache.end()) {
            *out = it->second;
            return CVLM_OK;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (reinterpret_cast<uintptr_t>(spec.ptr) & 15) return set_error(CVLM_ERR_ARG, "tensor map: base not 16-byte aligned");
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    if (spec.rank < 1 || spec.rank > 5) return set_error(CVLM_ERR_ARG, "tensor map: rank %d", spec.rank);
    for (int i = 0; i < spec.rank; ++i) {
        dims[i] = spec.dims[i];
        box[i] = spec.box[i];
    }
    for (int i = 0; i + 1 < spec.rank; ++i) {
        if (spec.strides[i] & 15) return set_error(CVLM_ERR_ARG, "tensor map: stride not a multiple of 16 bytes");
        strides[i] = spec.strides[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, spec.rank, const_cast<void*>(spec.ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(CVLM_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 8192) cache.clear();
    cache[spec] = *out;
    return CVLM_OK;
}

inline MapSpec map_spec_zero() {
    MapSpec s;
    memset(&s, 0, sizeof(s));
    return s;
}

}  // namespace cvlm
