// ht_common.cuh -- shared device/host helpers for the sm_100a kernels:
// error codes of the C-ABI, TMA tensor-map creation through the driver entry
// point (no link-time libcuda dependency, so the .so loads on a CPU-only box),
// and the mbarrier / cp.async.bulk.tensor PTX wrappers the halo-tile kernels use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hydrotools_b200.h"

#define HT_CUDA_TRY(expr)                       \
    do {                                        \
        cudaError_t e__ = (expr);               \
        if (e__ != cudaSuccess)                 \
            return (int)e__;                    \
    } while (0)

#define HT_LAUNCH_CHECK()                       \
    do {                                        \
        cudaError_t e__ = cudaGetLastError();   \
        if (e__ != cudaSuccess)                 \
            return (int)e__;                    \
    } while (0)

namespace ht {

constexpr int kSMs = 148; // B200: 2 dies x 74 SMs

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// TMA tensor maps
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(
    CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
    CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault,
                                    &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D row-major tensor of `elem_bytes`-wide elements: dims (cols, rows), row
// pitch `ld` elements.  Requirements (checked): base 16-B aligned, pitch a
// multiple of 16 B, box_w*elem_bytes a multiple of 16 B, box dims <= 256.
static inline int make_map_2d(CUtensorMap *map, CUtensorMapDataType dt,
                              int elem_bytes, const void *base, int64_t rows,
                              int64_t cols, int64_t ld, int box_w, int box_h,
                              bool nan_fill)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc)
        return HT_ERR_DRIVER;
    if (((uintptr_t)base & 15) || ((ld * elem_bytes) & 15) ||
        ((box_w * elem_bytes) & 15) || box_w > 256 || box_h > 256)
        return HT_ERR_ALIGN;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(ld * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA
                              : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : HT_ERR_DRIVER;
}

// ---------------------------------------------------------------------------
// mbarrier + bulk tensor copy (PTX; SASS: SYNCS.*, UTMALDG)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(count));
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// generic-proxy writes/reads of smem -> visible to / ordered before the async
// proxy (TMA) that is about to overwrite the buffer
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map,
                                            int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// streaming 128-bit store (outputs are written once, never re-read here)
__device__ __forceinline__ void st_stream_f4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
#endif

} // namespace ht
