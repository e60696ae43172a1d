// Device helpers shared by the interleaved-tile rotate kernels (rotate_il.cu: whole-volume tiles at 8^3 / 16^3,
// rotate_slab.cu: source-slab tiles at 32^3): streaming loads / stores, the hashed 16-byte unit placement, the
// lane -> voxel map of an 8x2x2 output block and the bit-exact coordinate chain.
#pragma once

#include "hg_common.cuh"
#include "rotate_common.cuh"

namespace hg {

__device__ __forceinline__ uint32_t ld_stream_4(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ld_stream_2(const void *p)
{
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}

// hashed unit index of voxel v = (z*S + y)*S + x
__device__ __forceinline__ int il_unit(int v, int logS)
{
    const int y = v >> logS, z = v >> (2 * logS);
    return v ^ (((y << 1) ^ (z << 2)) & 7);
}

template <typename T> struct IlUnit;
template <> struct IlUnit<float> {
    static constexpr int CI = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
};
template <> struct IlUnit<__nv_bfloat16> {
    static constexpr int CI = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

// lane -> voxel of the j-th 8x2x2 block: quarter warps (lanes 8q..8q+7) are 2x2x2 sub-blocks
__device__ __forceinline__ void il_block_voxel(int j, int lane, int S, int logS, int &x, int &y, int &z)
{
    // S is a power of two >= 8: S/8 blocks along x, S/2 along y and z
    const int lbx = logS - 3, lby = logS - 1;
    const int bx = j & ((1 << lbx) - 1), t = j >> lbx;
    const int by = t & ((1 << lby) - 1), bz = t >> lby;
    (void)S;
    x = (bx << 3) + (lane & 1) + ((lane >> 3) << 1);
    y = (by << 1) + ((lane >> 1) & 1);
    z = (bz << 1) + ((lane >> 2) & 1);
}

__device__ __forceinline__ void il_coords(const float *__restrict__ m, int ox, int oy, int oz, float &x, float &y, float &z)
{
    const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
    x = row_dot(m, fx, fy, fz);
    y = row_dot(m + 4, fx, fy, fz);
    z = row_dot(m + 8, fx, fy, fz);
}

template <typename T> __device__ __forceinline__ void st_stream_elem(T *p, float v);
template <> __device__ __forceinline__ void st_stream_elem<float>(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
template <> __device__ __forceinline__ void st_stream_elem<__nv_bfloat16>(__nv_bfloat16 *p, float v)
{
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<const uint16_t *>(&h)) : "memory");
}

}  // namespace hg
