// tma.cuh -- tensor-map encoding (host) and the tensor-copy / mbarrier PTX wrappers (device) shared by the image kernels.
#pragma once
#include <cuda.h>

#include <string>

#include "common.cuh"

namespace slideo {

typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmaEncodeTiledFn tma_encode_fn() {
    static TmaEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SLIDEO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p) throw CudaError(cudaErrorNotSupported, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<TmaEncodeTiledFn>(p);
    }
    return fn;
}

// rank-3 tensor {x, y, image}, no swizzle, out-of-bounds elements (negative coordinates included) read as zero
inline CUtensorMap tma_map_3d(CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t dim_x, uint64_t dim_y, uint64_t dim_z,
                              uint64_t stride_y_bytes, uint64_t stride_z_bytes, uint32_t box_x, uint32_t box_y) {
    CUtensorMap tm;
    const cuuint64_t dims[3] = {dim_x, dim_y, dim_z};
    const cuuint64_t strides[2] = {stride_y_bytes, stride_z_bytes};
    const cuuint32_t box[3] = {box_x, box_y, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    (void)elem_bytes;
    const CUresult r = tma_encode_fn()(&tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return tm;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(tma_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// one box of a rank-3 tensor -> shared memory, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     tma_smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(tma_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
#endif

}  // namespace slideo
