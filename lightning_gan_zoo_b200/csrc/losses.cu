// The two losses of HOLOGAN.training_step (reference core/lightning_module.py:217-237) in one launch each way:
//   adv = wa * mean_i bce(a[i], ta) + wb * mean_j bce(b[j], tb)        bce = BCEWithLogits against a constant target
//   q   = mean_k (zp[k] - z[k])^2                                       (the latent-reconstruction "q loss")
// D step (:219-229): a = D(real), ta = 1, wa = 1/2, b = D(fake), tb = 0, wb = 1/2.   G step (:231-237): a = D(fake),
// ta = 1, wa = 1, no b.  The stock path is ~15 elementwise/reduce launches forward and as many backward on tensors
// of 64 ... 8192 elements -- pure launch latency.  One CTA, fixed summation order (deterministic).
#include "hg_common.cuh"

namespace hg {

// BCEWithLogits with a constant target t, torch's stable form: (1 - t) * x + max(-x, 0) + log1p(exp(-|x|))
__device__ __forceinline__ float bce_logits(float x, float t)
{
    return (1.f - t) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x)
{
    // 1 / (1 + exp(-x)) without overflow for large |x|
    const float e = expf(-fabsf(x));
    return x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
}

template <int NT> __device__ __forceinline__ float block_sum(float v, float *sh)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = lane < NT / 32 ? sh[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;        // valid in warp 0
}

template <typename T>
__global__ void __launch_bounds__(256) gan_loss_fwd_kernel(const T *__restrict__ a, int na, float ta, float wa,
                                                           const T *__restrict__ b, int nb, float tb, float wb,
                                                           const T *__restrict__ zp, const float *__restrict__ z, int nz,
                                                           float *__restrict__ total, float *__restrict__ parts)
{
    __shared__ float sh[8];
    float sa = 0.f, sb = 0.f, sq = 0.f;
    for (int i = threadIdx.x; i < na; i += 256) sa += bce_logits(to_f32<T>(a[i]), ta);
    for (int i = threadIdx.x; i < nb; i += 256) sb += bce_logits(to_f32<T>(b[i]), tb);
    for (int i = threadIdx.x; i < nz; i += 256) {
        const float d = to_f32<T>(zp[i]) - z[i];
        sq = fmaf(d, d, sq);
    }
    sa = block_sum<256>(sa, sh);
    sb = block_sum<256>(sb, sh);
    sq = block_sum<256>(sq, sh);
    if (threadIdx.x == 0) {
        float adv = wa * (sa / (float)na);
        if (nb > 0) adv += wb * (sb / (float)nb);
        const float q = sq / (float)nz;
        total[0] = adv + q;
        parts[0] = adv;
        parts[1] = q;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gan_loss_bwd_kernel(const float *__restrict__ gout, const T *__restrict__ a, int na,
                                                           float ta, float wa, const T *__restrict__ b, int nb, float tb,
                                                           float wb, const T *__restrict__ zp, const float *__restrict__ z,
                                                           int nz, T *__restrict__ da, T *__restrict__ db, T *__restrict__ dzp)
{
    const float g = gout ? gout[0] : 1.f;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < na) da[i] = from_f32<T>(g * wa / (float)na * (sigmoidf_(to_f32<T>(a[i])) - ta));
    if (i < nb) db[i] = from_f32<T>(g * wb / (float)nb * (sigmoidf_(to_f32<T>(b[i])) - tb));
    if (i < nz) dzp[i] = from_f32<T>(g * 2.f / (float)nz * (to_f32<T>(zp[i]) - z[i]));
}

}  // namespace hg

using namespace hg;

extern "C" int hg_gan_loss_fwd(const void *a, int na, float ta, float wa, const void *b, int nb, float tb, float wb,
                               const void *zp, const float *z, int nz, int dtype, float *total, float *parts, void *stream)
{
    HG_REQUIRE(a && zp && z && total && parts, HG_ERR_INVALID_ARG, "hg_gan_loss_fwd: null pointer");
    HG_REQUIRE(na > 0 && nz > 0 && nb >= 0 && (nb == 0 || b), HG_ERR_INVALID_ARG, "hg_gan_loss_fwd: bad sizes");
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_gan_loss_fwd: dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == HG_F32)
        gan_loss_fwd_kernel<float><<<1, 256, 0, st>>>((const float *)a, na, ta, wa, (const float *)b, nb, tb, wb,
                                                      (const float *)zp, z, nz, total, parts);
    else
        gan_loss_fwd_kernel<__nv_bfloat16><<<1, 256, 0, st>>>((const __nv_bfloat16 *)a, na, ta, wa, (const __nv_bfloat16 *)b,
                                                              nb, tb, wb, (const __nv_bfloat16 *)zp, z, nz, total, parts);
    return check_launch("gan_loss_fwd");
}

extern "C" int hg_gan_loss_bwd(const float *gout, const void *a, int na, float ta, float wa, const void *b, int nb, float tb,
                               float wb, const void *zp, const float *z, int nz, int dtype, void *da, void *db, void *dzp,
                               void *stream)
{
    HG_REQUIRE(a && zp && z && da && dzp, HG_ERR_INVALID_ARG, "hg_gan_loss_bwd: null pointer");
    HG_REQUIRE(na > 0 && nz > 0 && nb >= 0 && (nb == 0 || (b && db)), HG_ERR_INVALID_ARG, "hg_gan_loss_bwd: bad sizes");
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_gan_loss_bwd: dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = max(max(na, nb), nz), grid = (n + 255) / 256;
    if (dtype == HG_F32)
        gan_loss_bwd_kernel<float><<<grid, 256, 0, st>>>(gout, (const float *)a, na, ta, wa, (const float *)b, nb, tb, wb,
                                                         (const float *)zp, z, nz, (float *)da, (float *)db, (float *)dzp);
    else
        gan_loss_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(gout, (const __nv_bfloat16 *)a, na, ta, wa,
                                                                 (const __nv_bfloat16 *)b, nb, tb, wb,
                                                                 (const __nv_bfloat16 *)zp, z, nz, (__nv_bfloat16 *)da,
                                                                 (__nv_bfloat16 *)db, (__nv_bfloat16 *)dzp);
    return check_launch("gan_loss_bwd");
}
