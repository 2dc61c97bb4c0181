// Host-side CUtensorMap builders shared by the tcgen05 GEMM (gemm2_tc.cu), attention (attention_tc.cu) and SwinV2 kernels:
// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda), bf16, 128-byte swizzle.
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"

namespace tmap {

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

}  // namespace tmap

// 2-D bf16 row-major tensor [rows][cols] with leading dimension ld (elements); box = {box_cols, box_rows}, 128-byte swizzle.
int rgbnm_make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
    tmap::EncodeTiledFn enc = tmap::get_encode();
    if (!enc) return RGBNM_ERR_CUDA;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
    cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbnm_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
        return RGBNM_ERR_CUDA;
    }
    return RGBNM_OK;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous; pitches ld1, ld2 in elements); box = {box0, box1, 1}, 128-byte swizzle.
int rgbnm_make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long long d0, long long d1, long long d2, long long ld1,
                            long long ld2, int box0, int box1) {
    tmap::EncodeTiledFn enc = tmap::get_encode();
    if (!enc) return RGBNM_ERR_CUDA;
    cuuint64_t dims[3] = {cuuint64_t(d0), cuuint64_t(d1), cuuint64_t(d2)};
    cuuint64_t strides[2] = {cuuint64_t(ld1) * 2, cuuint64_t(ld2) * 2};
    cuuint32_t box[3] = {cuuint32_t(box0), cuuint32_t(box1), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbnm_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(3d)");
        return RGBNM_ERR_CUDA;
    }
    return RGBNM_OK;
}
