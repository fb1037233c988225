// Minimal TMA (cp.async.bulk.tensor) + mbarrier wrappers for sm_100a, and the host-side tensor-map encoder.
// The driver entry point cuTensorMapEncodeTiled is resolved through the runtime (no -lcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pea {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// try_wait blocks in hardware until the phase completes or the time hint expires: with a generous hint a waiting warp
// sleeps instead of spinning through SYNCS + BRA pairs (measured: 7 % of the fused 3-D kernel's instruction stream)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "PEA_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra PEA_MBAR_DONE_%=;\n\t"
        "bra PEA_MBAR_WAIT_%=;\n"
        "PEA_MBAR_DONE_%=:\n\t}"
        ::"r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
}

// global (5-D tensor map) -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// shared -> global (5-D tensor map); elements of the box that fall outside the tensor are not written.
// Completion is tracked by the issuing thread's bulk async-group (tma_store_commit / tma_store_wait_read).
__device__ __forceinline__ void tma_store_5d(const void* smem_src, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- host: encode a tiled map over a contiguous float32 (B, C, D, H, W) tensor with box (bx, by, bz, bc, 1) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
    // function-local static: resolved once per process, immutable afterwards (not "mutable global state")
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// returns 0 on success, a CUresult (> 0) otherwise
// `plane` = elements between consecutive channel planes (>= D*H*W: the volume may be a z-range view of a deeper buffer)
// `cstride` > 1: the box spans `bc` consecutive channels of which every cstride-th is copied (ceil(bc / cstride) planes
// land in shared memory) -- e.g. the y-shift channels 1, 4, 7, 10 of a 12-channel target in one copy.
inline int encode_volume_map(CUtensorMap* map, const float* base, int B, int C, int D, int H, int W, int bx, int by,
                             int bz, int bc, int64_t plane = 0, int cstride = 1) {
    EncodeTiledFn fn = get_encode_tiled();
    if (!fn) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, (cuuint64_t)B};
    if (plane <= 0) plane = (int64_t)W * H * D;
    const cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)plane * 4,
                                   (cuuint64_t)plane * C * 4};
    const cuuint32_t box[5] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bc, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, (cuuint32_t)cstride, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace pea
