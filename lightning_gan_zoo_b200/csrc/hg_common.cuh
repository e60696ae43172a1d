// Shared device/host helpers for the sm_100a kernels behind include/hologan_b200.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hologan_b200.h"

namespace hg {

// ---- thread-local error reporting ---------------------------------------------------------------
char *error_buffer();   // defined in c_api.cu
int fail(int code, const char *fmt, ...);

#define HG_REQUIRE(cond, code, ...)                 \
    do {                                            \
        if (!(cond)) return hg::fail(code, __VA_ARGS__); \
    } while (0)

inline int check_launch(const char *what)
{
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HG_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    }
    return HG_OK;
}

// ---- tuning options ------------------------------------------------------------------------------
// Process-wide switches between EQUIVALENT kernels (A/B measurements, tests of both implementations of a pass).  Each is
// read from the environment variable HG_<NAME> once, when the library first needs it, and can be changed afterwards only
// through hg_set_option(); no kernel reads the environment at launch time.  They never change results beyond the
// documented tolerances.  (defined in c_api.cu)
enum Option {
    kOptAdainClNoCluster = 0,   // 1: channels-last AdaIN on the chunked two-kernel path only
    kOptAdainClClusterBwd,      // 1: cluster single-pass kernel for the channels-last AdaIN backward
    kOptTapGemmDual,            // -1 auto, 0 / 1: force one / two tap-GEMM CTAs per SM
    kOptFinalConvMma,           // bit mask: final layer passes on the mma.sync kernels (1 fwd, 2 dx, 4 dw)
    kOptRotateSlab32,           // 1: 32^3 rotate forward on source-slab tiles
    kOptRotateGatherBwd,        // 1: 32^3 rotate backward as a table-free per-voxel gather
    kOptAdainGemmStats,         // 1: generator AdaIN statistics from the tap-GEMM epilogue (read by the Python layer)
    kOptTapGemmPersistent,      // 1: tap GEMMs with more tiles than SMs run one persistent CTA per SM (0: one tile per CTA)
    kOptTapGemmMsub,            // 1: wide tap GEMMs (256-column tiles) process two 128-row sub-tiles per CTA that share every B tile
    kOptTapGemmShareA,          // > 0: forward of the narrow layers loads every shifted A box once for all parity classes of a CTA (value = weight boxes per item)
    kOptAdainClRing,            // 1: chunked channels-last AdaIN backward streams through per-thread cp.async rings
    kOptAdainClSmallRegs,       // 1: one-CTA-per-sample channels-last norm kernels keep their rows in registers (one load pass)
    kOptCount
};
int option(Option o);

// ---- element access -----------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte streaming global accesses (read-once / write-once data: keep L1 for the gather tables)
__device__ __forceinline__ uint4 ld_stream_16(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// 16-byte asynchronous global -> shared copy (LDGSTS, bypasses L1 and the register file).  `valid == false` reads nothing
// and fills the 16 bytes with zeros (src-size 0): halo pixels outside the image.  A staging loop written with these keeps
// every copy of a thread in flight at once; the same loop with `smem = __ldg(global)` pays one memory round trip per
// iteration whenever the compiler cannot hoist the loads over the stores.
__device__ __forceinline__ void cp_async_16_zfill(void *smem_dst, const void *gsrc, bool valid)
{
    const unsigned n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void st_stream_16(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

inline int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

}  // namespace hg
