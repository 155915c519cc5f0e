// Host side of the TMA tile loads: tensor-map encoding through the driver entry point (libcuda is not linked; the
// function is looked up once with cudaGetDriverEntryPoint).
#include "p3d_pocs_kernels.cuh"

#include <cuda_runtime.h>
#include <cstdlib>
#include <mutex>

namespace p3d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("P3D_NO_TMA")) return;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else cudaGetLastError();
    });
    return fn;
}

bool tma_encode_tile_map(CUtensorMap* map, const void* base, long long slices, int n1, int n2, int elem_bytes, int C, int rows, int row_step) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || !base || slices <= 0) return false;
    const int per = elem_bytes / 8;                           // 8-byte tensor elements per complex value
    const cuuint64_t gdim[3] = {(cuuint64_t)n2 * per, (cuuint64_t)n1, (cuuint64_t)slices};
    const cuuint64_t gstride[2] = {(cuuint64_t)n2 * elem_bytes, (cuuint64_t)n1 * n2 * elem_bytes};
    if (gstride[0] % 16 != 0 || ((uintptr_t)base % 16) != 0) return false;
    // row_step > 1: every row_step-th row of the box (`rows` rows land in shared memory, rows * row_step are traversed)
    const cuuint32_t box[3] = {(cuuint32_t)(C * per), (cuuint32_t)(rows * row_step), 1u};
    const cuuint32_t estr[3] = {1u, (cuuint32_t)row_step, 1u};
    if (rows * row_step > 256) return false;
    static const bool promote = getenv("P3D_TMA_L2_128") != nullptr;
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace p3d
