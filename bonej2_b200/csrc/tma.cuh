// tma.cuh -- the Tensor Memory Accelerator pieces the tile kernels use (sm_90+; built here for sm_100a): tiled tensor
// maps encoded on the host through the driver entry point (no link against libcuda), and on the device one mbarrier
// per tile that the bulk-tensor copies complete on.  A box that reaches outside the tensor is filled with zeros,
// which is exactly the reference's "a neighbour outside the image counts as 0" rule for the ridge and the cleanup.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bjt {
namespace tma {

inline PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = []() {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

// [d][h][w] volume of 4-byte elements, box bx x by x bz (x fastest); false: no TMA here, take the plain-load kernel
inline bool make_map_3d_u32(CUtensorMap *m, const void *base, int w, int h, int d, int bx, int by, int bz) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = encode_fn();
    if (!enc || (uintptr_t)base % 16 != 0 || w % 4 != 0 || bx > 256 || by > 256 || bz > 256 || (bx * 4) % 16 != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d};
    const cuuint64_t strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;   // none: 4.5 ms against 2.7 ms for the z pass at 1024^3
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(phase)
        : "memory");
}
// one box of the tensor into shared memory; completes `bytes of the box` on the barrier
__device__ __forceinline__ void load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_addr(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar))
                 : "memory");
}

}  // namespace tma
}  // namespace bjt
