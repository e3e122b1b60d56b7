// Host side of the tensor-core path: TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#include <string.h>
#include "kernels.cuh"
#include "tc_common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int dgprf_make_tmap_3d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t chains,
                       uint64_t row_stride_floats, uint64_t chain_stride_floats, uint32_t box_rows, bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        dgprf_set_error("cuTensorMapEncodeTiled entry point not available");
        return DGPRF_ECUDA;
    }
    DGPRF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride_floats % 4) == 0 &&
                  (chains <= 1 || (chain_stride_floats % 4) == 0),
                  "TMA needs 16-byte aligned base and strides (base=%p row_stride=%llu chain_stride=%llu)", (const void*)base,
                  (unsigned long long)row_stride_floats, (unsigned long long)chain_stride_floats);
    const cuuint64_t dims[3] = {cols, rows, chains < 1 ? 1 : chains};
    const cuuint64_t strides[2] = {row_stride_floats * sizeof(float),
                                   (chains <= 1 ? rows * row_stride_floats : chain_stride_floats) * sizeof(float)};
    const cuuint32_t box[3] = {32, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dgprf_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return DGPRF_ECUDA;
    }
    return DGPRF_OK;
}
