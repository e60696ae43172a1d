// Channels-last AdaIN (+ activation) for the bf16 tensor-core pipeline.
//
// Same arithmetic as adain.cu (reference AdaIn, core/models/hologan_generator.py:333-345, fused with
// the ReLU of :41) but on the layouts the implicit-GEMM kernels produce and consume:
//   input  x : (B, Npos, P, C) bf16 -- the space-to-depth output of hg_convt_fwd (P parity classes,
//              P = 1 for a plain channels-last tensor), rows r = pos * P + cls
//   output y : (B, (2S)^d, C) bf16 plain channels-last: row (pos, cls) lands on its up-sampled pixel,
//              i.e. the depth-to-space shuffle is folded into the store addressing (zero extra passes).
// One CTA = (sample, 16 channels): every row contributes one 32-byte sector, the whole (N x 16) slab
// (N <= 4096 rows) is held in registers, statistics are reduced with shuffles + one smem exchange, and
// the tensor is read exactly once.  The backward makes two passes over its inputs (sums, then dx);
// the second pass hits L2.
#include "hg_common.cuh"

namespace hg {

constexpr int kClThreads = 512;
constexpr int kClGroup = 16;        // channels per CTA

__device__ __forceinline__ void unpack8(const uint4 &u, float *f)
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float *f)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// row r = pos * P + cls of the s2d tensor -> row of the up-sampled channels-last tensor
__device__ __forceinline__ int upsampled_row(int r, int ndim, int logS, int logP)
{
    if (logP == 0) return r;
    const int cls = r & ((1 << logP) - 1), pos = r >> logP;
    const int S = 1 << logS, S2 = 2 * S;
    const int ix = pos & (S - 1), iy = (pos >> logS) & (S - 1);
    const int px = cls & 1, py = (cls >> 1) & 1;
    if (ndim == 2) return (2 * iy + py) * S2 + 2 * ix + px;
    const int iz = pos >> (2 * logS), pz = cls >> 2;
    return ((2 * iz + pz) * S2 + 2 * iy + py) * S2 + 2 * ix + px;
}

// Sum 8 per-thread partials over all threads of the CTA that own the same channel half (t & 1).
// red: [kWarps][16 channels] floats; result is valid in every thread.
template <int kWarps>
__device__ __forceinline__ void cta_sum8(float (&a)[8], float (*red)[kClGroup], int warp, int lane)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o >= 2; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    }
    if (lane < 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = a[j];
    }
    __syncthreads();
    const int half = lane & 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += red[w][half * 8 + j];
    }
}

__device__ __forceinline__ float modulate_cl(float x, float mean, float rstd, float s, float b)
{
    return __fadd_rn(__fmul_rn(s, __fmul_rn(__fsub_rn(x, mean), rstd)), b);
}

template <int VPT, int kThreadsT>
__global__ void __launch_bounds__(kThreadsT) adain_cl_fwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                  const float *__restrict__ scale,
                                                                  const float *__restrict__ bias,
                                                                  __nv_bfloat16 *__restrict__ y,
                                                                  float *__restrict__ save_mean,
                                                                  float *__restrict__ save_rstd, int C, int N, int Nvar,
                                                                  int ndim, int logS, int logP, int sbs, float eps,
                                                                  float slope)
{
    __shared__ float red[2][kThreadsT / 32][kClGroup];
    const int b = blockIdx.y, c0 = blockIdx.x * kClGroup;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, half = t & 1;
    const int nvec = N * 2;
    const size_t base = (size_t)b * N * C + c0 + half * 8;
    const __nv_bfloat16 *xb = x + base;

    uint4 raw[VPT];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int v = t + k * kThreadsT;
        if (v < nvec) {
            raw[k] = ld_stream_16(xb + (size_t)(v >> 1) * C);
            float f[8];
            unpack8(raw[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
    cta_sum8<kThreadsT / 32>(acc, red[0], warp, lane);
    float mean[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mean[j] = acc[j] / (float)N;
        acc[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        if (t + k * kThreadsT < nvec) {
            float f[8];
            unpack8(raw[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = f[j] - mean[j];
                acc[j] += d * d;
            }
        }
    }
    cta_sum8<kThreadsT / 32>(acc, red[1], warp, lane);
    float rstd[8], s[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // Nvar = N - 1: unbiased variance (:338), eps inside the rsqrt (:339); Nvar = N: InstanceNorm2d
        rstd[j] = __frsqrt_rn(acc[j] / (float)Nvar + eps);
        s[j] = scale ? scale[(size_t)b * sbs + c0 + half * 8 + j] : 1.f;
        bb[j] = bias ? bias[(size_t)b * sbs + c0 + half * 8 + j] : 0.f;
    }
    if (t < 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            save_mean[(size_t)b * C + c0 + half * 8 + j] = mean[j];
            save_rstd[(size_t)b * C + c0 + half * 8 + j] = rstd[j];
        }
    }
    __nv_bfloat16 *yb = y + base;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int v = t + k * kThreadsT;
        if (v < nvec) {
            float f[8];
            unpack8(raw[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = modulate_cl(f[j], mean[j], rstd[j], s[j], bb[j]);
                f[j] = p > 0.f ? p : p * slope;
            }
            const int orow = upsampled_row(v >> 1, ndim, logS, logP);
            st_stream_16(yb + (size_t)orow * C, pack8(f));
        }
    }
}

// Backward.  256 threads per (sample, 16 channels); two passes over x / dy (the second one hits L2), each
// pass issues kBwdUnroll independent 16-byte loads of x and of dy per thread before consuming them, so a
// CTA keeps ~64 KB in flight.  Nvar as in the forward.  scale / bias may be null (1 / 0), dscale / dbias may
// be null (not needed: the discriminator's InstanceNorm has no affine parameters).
constexpr int kBwdThreads = 256;
constexpr int kBwdUnroll = 4;

__global__ void __launch_bounds__(kBwdThreads) adain_cl_bwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                   const __nv_bfloat16 *__restrict__ dy,
                                                                   const float *__restrict__ scale,
                                                                   const float *__restrict__ bias,
                                                                   const float *__restrict__ save_mean,
                                                                   const float *__restrict__ save_rstd,
                                                                   __nv_bfloat16 *__restrict__ dx, float *__restrict__ dscale,
                                                                   float *__restrict__ dbias, int C, int N, int Nvar, int ndim,
                                                                   int logS, int logP, int sbs, int dsbs, float slope)
{
    __shared__ float red[2][kBwdThreads / 32][kClGroup];
    const int b = blockIdx.y, c0 = blockIdx.x * kClGroup;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, half = t & 1;
    const int nvec = N * 2;
    const size_t base = (size_t)b * N * C + c0 + half * 8;
    const __nv_bfloat16 *xb = x + base, *gb = dy + base;

    float mean[8], rstd[8], s[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + half * 8 + j;
        mean[j] = save_mean[(size_t)b * C + c];
        rstd[j] = save_rstd[(size_t)b * C + c];
        s[j] = scale ? scale[(size_t)b * sbs + c] : 1.f;
        bb[j] = bias ? bias[(size_t)b * sbs + c] : 0.f;
    }
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    for (int v0 = t; v0 < nvec; v0 += kBwdThreads * kBwdUnroll) {
        uint4 xr[kBwdUnroll], gr[kBwdUnroll];
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            const int v = v0 + u * kBwdThreads;
            if (v < nvec) {
                const int row = v >> 1;
                xr[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)row * C));
                gr[u] = __ldg(reinterpret_cast<const uint4 *>(gb + (size_t)upsampled_row(row, ndim, logS, logP) * C));
            }
        }
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            if (v0 + u * kBwdThreads < nvec) {
                float xf[8], gf[8];
                unpack8(xr[u], xf);
                unpack8(gr[u], gf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float pre = modulate_cl(xf[j], mean[j], rstd[j], s[j], bb[j]);     // the forward's bits
                    const float g = pre > 0.f ? gf[j] : gf[j] * slope;
                    sg[j] += g;
                    sgx[j] += g * ((xf[j] - mean[j]) * rstd[j]);
                }
            }
        }
    }
    cta_sum8<kBwdThreads / 32>(sg, red[0], warp, lane);
    cta_sum8<kBwdThreads / 32>(sgx, red[1], warp, lane);
    if (t < 2 && dscale && dbias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dbias[(size_t)b * dsbs + c0 + half * 8 + j] = sg[j];
            dscale[(size_t)b * dsbs + c0 + half * 8 + j] = sgx[j];
        }
    }
    float k1[8], k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        k1[j] = s[j] * sg[j] / (float)N;
        k2[j] = s[j] * sgx[j] / (float)Nvar;
    }
    __nv_bfloat16 *db = dx + base;
    for (int v0 = t; v0 < nvec; v0 += kBwdThreads * kBwdUnroll) {
        uint4 xr[kBwdUnroll], gr[kBwdUnroll];
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            const int v = v0 + u * kBwdThreads;
            if (v < nvec) {
                const int row = v >> 1;
                xr[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)row * C));
                gr[u] = __ldg(reinterpret_cast<const uint4 *>(gb + (size_t)upsampled_row(row, ndim, logS, logP) * C));
            }
        }
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            const int v = v0 + u * kBwdThreads;
            if (v < nvec) {
                float xf[8], gf[8];
                unpack8(xr[u], xf);
                unpack8(gr[u], gf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (xf[j] - mean[j]) * rstd[j];
                    const float pre = modulate_cl(xf[j], mean[j], rstd[j], s[j], bb[j]);
                    const float g = pre > 0.f ? gf[j] : gf[j] * slope;
                    xf[j] = rstd[j] * (g * s[j] - k1[j] - xh * k2[j]);
                }
                st_stream_16(db + (size_t)(v >> 1) * C, pack8(xf));
            }
        }
    }
}

static int ilog2(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

}  // namespace hg

using namespace hg;

static int cl_check(const char *who, int batch, int channels, int ndim, int size, int classes, int &n, int &logS, int &logP)
{
    HG_REQUIRE(batch > 0 && channels > 0 && size > 0 && classes > 0, HG_ERR_INVALID_ARG, "%s: dims must be positive", who);
    HG_REQUIRE(batch <= 65535, HG_ERR_UNSUPPORTED, "%s: batch > 65535", who);
    HG_REQUIRE(ndim == 2 || ndim == 3, HG_ERR_INVALID_ARG, "%s: ndim must be 2 or 3", who);
    logS = ilog2(size);
    logP = ilog2(classes);
    HG_REQUIRE(logS >= 0 && (classes == 1 || classes == (1 << ndim)), HG_ERR_UNSUPPORTED,
               "%s: size must be a power of two and classes 1 or 2^ndim", who);
    HG_REQUIRE(channels % kClGroup == 0, HG_ERR_UNSUPPORTED, "%s: channels must be a multiple of %d", who, kClGroup);
    long long rows = classes;
    for (int i = 0; i < ndim; ++i) rows *= size;
    HG_REQUIRE(rows >= 2 && rows <= 4096, HG_ERR_UNSUPPORTED, "%s: %lld rows per instance exceed the single-pass limit 4096", who, rows);
    n = (int)rows;
    return HG_OK;
}

extern "C" int hg_adain_cl_fwd(const void *x, const float *scale, const float *bias, void *y, float *save_mean,
                               float *save_rstd, int batch, int channels, int ndim, int size, int classes, int sb_stride,
                               float eps, float neg_slope, int biased_var, void *stream)
{
    HG_REQUIRE(x && y && save_mean && save_rstd, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: null pointer");
    HG_REQUIRE((scale == nullptr) == (bias == nullptr), HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: scale and bias must both be given or both be null");
    int n, logS, logP;
    int rc = cl_check("hg_adain_cl_fwd", batch, channels, ndim, size, classes, n, logS, logP);
    if (rc) return rc;
    HG_REQUIRE(!scale || sb_stride >= channels, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: sb_stride < channels");
    const int nvar = biased_var ? n : n - 1;
    const __nv_bfloat16 *xp = static_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *yp = static_cast<__nv_bfloat16 *>(y);
    dim3 grid(channels / kClGroup, batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int vpt = (n * 2 + kClThreads - 1) / kClThreads;
#define HG_LAUNCH_CL_T(V, T) adain_cl_fwd_kernel<V, T><<<grid, T, 0, st>>>(xp, scale, bias, yp, save_mean, save_rstd, channels, n, nvar, ndim, logS, logP, sb_stride, eps, neg_slope)
#define HG_LAUNCH_CL(V) HG_LAUNCH_CL_T(V, kClThreads)
    // small instances (the discriminator's 16x16 ... 4x4 maps): right-size the CTA, one vector per thread
    if (n * 2 <= 64) HG_LAUNCH_CL_T(1, 64);
    else if (n * 2 <= 128) HG_LAUNCH_CL_T(1, 128);
    else if (n * 2 <= 256) HG_LAUNCH_CL_T(1, 256);
    else if (vpt <= 1) HG_LAUNCH_CL(1);
    else if (vpt <= 2) HG_LAUNCH_CL(2);
    else if (vpt <= 4) HG_LAUNCH_CL(4);
    else if (vpt <= 8) HG_LAUNCH_CL(8);
    else HG_LAUNCH_CL(16);
#undef HG_LAUNCH_CL
#undef HG_LAUNCH_CL_T
    return check_launch("hg_adain_cl_fwd");
}

extern "C" int hg_adain_cl_bwd(const void *x, const void *dy, const float *scale, const float *bias, const float *save_mean,
                               const float *save_rstd, void *dx, float *dscale, float *dbias, int batch, int channels,
                               int ndim, int size, int classes, int sb_stride, int dsb_stride, float neg_slope,
                               int biased_var, void *stream)
{
    HG_REQUIRE(x && dy && save_mean && save_rstd && dx, HG_ERR_INVALID_ARG, "hg_adain_cl_bwd: null pointer");
    HG_REQUIRE((scale == nullptr) == (bias == nullptr) && (dscale == nullptr) == (dbias == nullptr), HG_ERR_INVALID_ARG,
               "hg_adain_cl_bwd: scale/bias and dscale/dbias come in pairs");
    int n, logS, logP;
    int rc = cl_check("hg_adain_cl_bwd", batch, channels, ndim, size, classes, n, logS, logP);
    if (rc) return rc;
    HG_REQUIRE((!scale || sb_stride >= channels) && (!dscale || dsb_stride >= channels), HG_ERR_INVALID_ARG,
               "hg_adain_cl_bwd: stride < channels");
    dim3 grid(channels / kClGroup, batch);
    adain_cl_bwd_kernel<<<grid, kBwdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(dy), scale, bias, save_mean, save_rstd,
        static_cast<__nv_bfloat16 *>(dx), dscale, dbias, channels, n, biased_var ? n : n - 1, ndim, logS, logP, sb_stride,
        dsb_stride, neg_slope);
    return check_launch("hg_adain_cl_bwd");
}
