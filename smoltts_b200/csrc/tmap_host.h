// Host side of the TMA form of umma.cuh: 2-D tensor maps over row-major [rows][K] bf16 buffers with a 64-element x
// box_rows box and the 128-byte swizzle.  cuTensorMapEncodeTiled is fetched through the runtime
// (cudaGetDriverEntryPoint), so the library does not link against libcuda.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smol {

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }
    return fn;
}

// rows x K bf16, row pitch `pitch_elems` (>= K, pitch bytes a multiple of 16), box = 64 elements x box_rows rows.
// Rows / columns outside [rows) x [K) read as zero.  Returns false if the driver rejects the description.
inline bool make_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t K, uint64_t pitch_elems, uint32_t box_rows) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
    const cuuint32_t box[2] = {64u, box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace smol
