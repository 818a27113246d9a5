// Tensor-map (TMA descriptor) encoding without linking libcuda: the encoder comes from cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>  // CUtensorMap and enums only
#include <cuda_runtime.h>

namespace h2svd {

typedef CUresult (*tma_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tma_encode_fn tma_encoder() {
    static tma_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tma_encode_fn>(p);
    }
    return fn;
}

}  // namespace h2svd
