// Device functions shared by the rotate-resample kernels (rotate.cu: NCDHW / reference layout,
// rotate_cl.cu: channels-last pipeline layouts).
#pragma once

#include "hg_common.cuh"

namespace hg {

// -------------------------------------------------------------------------------------------------
// Coordinates and corners -- bit-exact with the reference CPU path
// -------------------------------------------------------------------------------------------------
struct Corners {
    int idx[8];     // flat (z*S + y)*S + x of corners a..h (hologan_generator.py:278-287)
    float w[8];     // weights a..h (:309-318)
    bool inside;    // all three coordinates in [0, S-1): the only samples that are not ~0
};

// src = A @ [x y z 1]^T.  torch.matmul on CPU (MKL sgemm, k = 4) is reproduced bit-for-bit by this
// sequential chain: one rounding for m0*x, then three fused multiply-adds.
__device__ __forceinline__ float row_dot(const float *__restrict__ r, float x, float y, float z)
{
    float acc = __fmul_rn(r[0], x);
    acc = __fmaf_rn(r[1], y, acc);
    acc = __fmaf_rn(r[2], z, acc);
    acc = __fmaf_rn(r[3], 1.0f, acc);
    return acc;
}

__device__ __forceinline__ int clampi(int v, int hi) { return min(max(v, 0), hi); }

// Shared-memory placement of voxel (z,y,x): rows keep their order but x is XOR-swizzled with a per-row
// key so that neighbours along y and z (and not only along x) fall into different banks.  The views
// of the hot path are near axis-aligned (azimuth ~270 deg, elevation ~90 deg): without the swizzle a
// warp walking along output x reads source voxels S or S^2 elements apart = one bank, a 16-32-way
// conflict.  ysh = log2(rows per 128-byte bank line).
__device__ __forceinline__ int swz(int z, int y, int x, int S, int logS, int ysh)
{
    return (((z << logS) + y) << logS) + (x ^ ((z ^ (y >> ysh)) & (S - 1)));
}

// Corner indices (floor, +1, clamp: :249-261) and weights from the CLAMPED corner as float against
// the UNCLAMPED coordinate, product order (wx*wy)*wz (:301-318).  Explicit _rn intrinsics keep nvcc
// from contracting / re-associating, so fp32 results carry the reference's bits.
// kSwz: idx[] addresses the swizzled shared-memory tile instead of the linear volume.
template <bool kSwz>
__device__ __forceinline__ void make_corners(float x, float y, float z, int S, int logS, int ysh, Corners &c)
{
    const int fx = __float2int_rd(x), fy = __float2int_rd(y), fz = __float2int_rd(z);
    const int x0 = clampi(fx, S - 1), x1 = clampi(fx + 1, S - 1);
    const int y0 = clampi(fy, S - 1), y1 = clampi(fy + 1, S - 1);
    const int z0 = clampi(fz, S - 1), z1 = clampi(fz + 1, S - 1);
    const float ux = __fsub_rn((float)x1, x), lx = __fsub_rn(x, (float)x0);
    const float uy = __fsub_rn((float)y1, y), ly = __fsub_rn(y, (float)y0);
    const float uz = __fsub_rn((float)z1, z), lz = __fsub_rn(z, (float)z0);
    if (kSwz) {
        const int m = S - 1;
        const int k00 = (z0 ^ (y0 >> ysh)) & m, k01 = (z0 ^ (y1 >> ysh)) & m;
        const int k10 = (z1 ^ (y0 >> ysh)) & m, k11 = (z1 ^ (y1 >> ysh)) & m;
        const int r00 = ((z0 << logS) + y0) << logS, r01 = ((z0 << logS) + y1) << logS;
        const int r10 = ((z1 << logS) + y0) << logS, r11 = ((z1 << logS) + y1) << logS;
        c.idx[0] = r00 + (x0 ^ k00); c.idx[1] = r01 + (x0 ^ k01); c.idx[2] = r00 + (x1 ^ k00); c.idx[3] = r01 + (x1 ^ k01);
        c.idx[4] = r10 + (x0 ^ k10); c.idx[5] = r11 + (x0 ^ k11); c.idx[6] = r10 + (x1 ^ k10); c.idx[7] = r11 + (x1 ^ k11);
    } else {
        const int r00 = (z0 * S + y0) * S, r01 = (z0 * S + y1) * S, r10 = (z1 * S + y0) * S, r11 = (z1 * S + y1) * S;
        c.idx[0] = r00 + x0; c.idx[1] = r01 + x0; c.idx[2] = r00 + x1; c.idx[3] = r01 + x1;
        c.idx[4] = r10 + x0; c.idx[5] = r11 + x0; c.idx[6] = r10 + x1; c.idx[7] = r11 + x1;
    }
    const float uxuy = __fmul_rn(ux, uy), uxly = __fmul_rn(ux, ly), lxuy = __fmul_rn(lx, uy), lxly = __fmul_rn(lx, ly);
    c.w[0] = __fmul_rn(uxuy, uz); c.w[1] = __fmul_rn(uxly, uz); c.w[2] = __fmul_rn(lxuy, uz); c.w[3] = __fmul_rn(lxly, uz);
    c.w[4] = __fmul_rn(uxuy, lz); c.w[5] = __fmul_rn(uxly, lz); c.w[6] = __fmul_rn(lxuy, lz); c.w[7] = __fmul_rn(lxly, lz);
    const float lim = (float)(S - 1);
    c.inside = (x >= 0.f) && (x < lim) && (y >= 0.f) && (y < lim) && (z >= 0.f) && (z < lim);
}

__device__ __forceinline__ void lattice_coords(const float *__restrict__ m, int o, int S, int logS, float &x, float &y,
                                               float &z)
{
    const int ox = o & (S - 1), oy = (o >> logS) & (S - 1), oz = o >> (2 * logS);
    const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
    x = row_dot(m, fx, fy, fz);
    y = row_dot(m + 4, fx, fy, fz);
    z = row_dot(m + 8, fx, fy, fz);
}

// out = ((((((w0*v0 + w1*v1) + w2*v2) + ...) + w7*v7): separate multiply and add, left to right,
// like the reference's `wa*Ia + wb*Ib + ...` tensor expression (:320).
__device__ __forceinline__ float blend8(const Corners &c, const float (&v)[8])
{
    float acc = __fmul_rn(c.w[0], v[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fadd_rn(acc, __fmul_rn(c.w[k], v[k]));
    return acc;
}

}  // namespace hg
