// sm100_host.cu -- host side of sm100.cuh: TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#include "sm100.cuh"

namespace eml {
namespace sm100 {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encoder() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}
}  // namespace

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeFn fn = encoder();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return EML_ERR_CUDA;
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; i++) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = 1;
        if (i + 1 < rank) s[i] = strides_bytes[i];
    }
    CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu x %llu, box %u x %u x %u)",
                  (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                  (unsigned long long)(rank > 2 ? d[2] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0);
        return EML_ERR_CUDA;
    }
    return EML_OK;
}

}  // namespace sm100
}  // namespace eml
