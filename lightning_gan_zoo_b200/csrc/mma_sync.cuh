// Warp-level tensor-core MMA (mma.sync.m16n8k16, bf16 operands, fp32 accumulation) helpers shared by the small
// bandwidth-bound contractions (final layer, the discriminator's first convolution).
// Fragment layouts (PTX ISA, .row.col) with g = lane / 4, q = lane % 4:
//   A (16 x 16): a0 = A[g][2q, 2q+1]   a1 = A[g+8][2q, 2q+1]   a2 = A[g][2q+8, 2q+9]   a3 = A[g+8][2q+8, 2q+9]
//   B (16 x  8): b0 = B[2q, 2q+1][g]   b1 = B[2q+8, 2q+9][g]
//   C (16 x  8): c0, c1 = C[g][2q, 2q+1]   c2, c3 = C[g+8][2q, 2q+1]
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace hg {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds32(const unsigned char *p) { return *reinterpret_cast<const uint32_t *>(p); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

}  // namespace hg
