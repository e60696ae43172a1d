// Adaptive instance norm fused with the activation that follows it at every reference call site.
// Replaces AdaIn (core/models/hologan_generator.py:333-345) + nn.ReLU (:41, :124) and, via
// x_batch_stride == 0, the `self.x.repeat(B, ...)` of :121 (the constant is never materialised).
//
// HBM-bound op: algorithmic traffic fwd = 1 read + 1 write of the tensor, bwd = 2 reads + 1 write.
// Single pass: a (sample, channel) instance (N <= 4096 fp32 / 8192 bf16 elements on the hot path)
// is held in registers by one warp (small N) or one 256-thread CTA, reduced with warp shuffles,
// normalised, modulated, activated and stored -- the tensor is read exactly once.
#include "hg_common.cuh"

namespace hg {

template <int GROUP> __device__ __forceinline__ float group_sum(float v, float *scratch);

template <> __device__ __forceinline__ float group_sum<32>(float v, float *) { return warp_sum(v); }

// scratch: 8 floats per reduction slot, caller provides distinct slots for back-to-back reductions
template <> __device__ __forceinline__ float group_sum<256>(float v, float *scratch)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    float t = scratch[threadIdx.x & 7];
    // 8 partials: reduce over the low 3 lane bits (every lane ends with the full sum)
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    return t;
}

template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *f)
    {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ uint4 pack(const float *f)
    {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&p);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// reference op order (:340-344): normalized = (x - mean) * sigma; scale * normalized; += bias
__device__ __forceinline__ float modulate(float x, float mean, float rstd, float s, float b)
{
    return __fadd_rn(__fmul_rn(s, __fmul_rn(__fsub_rn(x, mean), rstd)), b);
}

template <typename T, int GROUP, int ITEMS>
__global__ void __launch_bounds__(256) adain_fwd_kernel(const T *__restrict__ x, const float *__restrict__ scale,
                                                        const float *__restrict__ bias, T *__restrict__ y,
                                                        float *__restrict__ save_mean, float *__restrict__ save_rstd,
                                                        int BC, int C, int N, long long xbs, int sbs, float eps, float slope, int biased)
{
    constexpr int V = Vec16<T>::N;
    __shared__ float scratch[16];
    const int inst = blockIdx.x * (256 / GROUP) + threadIdx.x / GROUP;
    if (GROUP == 32 && inst >= BC) return;   // warp-uniform; GROUP==256 grids are exact
    const int t = threadIdx.x % GROUP;
    const int b = inst / C, c = inst - b * C;
    const uint4 *xp = reinterpret_cast<const uint4 *>(x + (size_t)b * xbs + (size_t)c * N);
    uint4 *yp = reinterpret_cast<uint4 *>(y + ((size_t)b * C + c) * N);
    const int nvec = N / V;

    float v[ITEMS][V];
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int vi = t + it * GROUP;
        if (vi < nvec) {
            Vec16<T>::unpack(xbs == 0 ? __ldg(xp + vi) : ld_stream_16(xp + vi), v[it]);
#pragma unroll
            for (int j = 0; j < V; ++j) sum += v[it][j];
        }
    }
    sum = group_sum<GROUP>(sum, scratch);
    const float mean = sum / (float)N;
    float m2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if (t + it * GROUP < nvec) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float d = v[it][j] - mean;
                m2 += d * d;
            }
        }
    }
    m2 = group_sum<GROUP>(m2, scratch + 8);
    const float var = m2 / (float)(biased ? N : N - 1);  // unbiased = torch.var default (:338); biased = InstanceNorm2d
    const float rstd = __frsqrt_rn(var + eps);           // :339
    const float s = scale[(size_t)b * sbs + c], bb = bias[(size_t)b * sbs + c];
    if (t == 0) {
        save_mean[inst] = mean;
        save_rstd[inst] = rstd;
    }
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int vi = t + it * GROUP;
        if (vi < nvec) {
            float o[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float p = modulate(v[it][j], mean, rstd, s, bb);
                o[j] = p > 0.f ? p : p * slope;
            }
            st_stream_16(yp + vi, Vec16<T>::pack(o));
        }
    }
}

// kConstX: x is the batch-shared constant (x_batch_stride == 0); one group owns channel c, loops
// over the batch and accumulates dx in registers (deterministic sum over samples).
template <typename T, int GROUP, int ITEMS, bool kConstX>
__global__ void __launch_bounds__(256) adain_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                        const float *__restrict__ scale, const float *__restrict__ bias,
                                                        const float *__restrict__ save_mean,
                                                        const float *__restrict__ save_rstd, T *__restrict__ dx,
                                                        float *__restrict__ dscale, float *__restrict__ dbias, int B, int C,
                                                        int N, long long xbs, int sbs, int dsbs, float slope, int biased)
{
    constexpr int V = Vec16<T>::N;
    // kSplit (the learned constant at warp-sized instances, the hot-path case C = 512, N = 64): one CTA per channel,
    // its 8 warps take the samples b = warp, warp + 8, ... and their partial dx sums meet in shared memory in a fixed
    // order.  A single warp walking all B samples (64 dependent load -> reduce rounds) took 64 us at B = 64.
    constexpr bool kSplit = kConstX && GROUP == 32;
    __shared__ float scratch[16];
    __shared__ float part[kSplit ? 8 * ITEMS * V * 32 : 1];
    const int n_inst = kConstX ? C : B * C;
    const int inst = kSplit ? (int)blockIdx.x : (int)(blockIdx.x * (256 / GROUP) + threadIdx.x / GROUP);
    if (!kSplit && GROUP == 32 && inst >= n_inst) return;
    const int t = threadIdx.x % GROUP;
    const int nvec = N / V;
    const int c = kConstX ? inst : inst % C;
    const int b_begin = kSplit ? (int)(threadIdx.x / 32) : (kConstX ? 0 : inst / C);
    const int b_end = kConstX ? B : b_begin + 1;
    const int b_step = kSplit ? 8 : 1;
    const float inv_nm1 = 1.f / (float)(biased ? N : N - 1), inv_n = 1.f / (float)N;

    float xv[ITEMS][V];
    float acc[ITEMS][V];
    if (kConstX) {
        const uint4 *xp = reinterpret_cast<const uint4 *>(x + (size_t)c * N);
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            if (t + it * GROUP < nvec) Vec16<T>::unpack(__ldg(xp + t + it * GROUP), xv[it]);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[it][j] = 0.f;
        }
    }
    for (int b = b_begin; b < b_end; b += b_step) {
        const size_t bc = (size_t)b * C + c;
        const uint4 *gp = reinterpret_cast<const uint4 *>(dy + bc * N);
        const float mean = save_mean[bc], rstd = save_rstd[bc];
        const float s = scale[(size_t)b * sbs + c], bb = bias[(size_t)b * sbs + c];
        if (!kConstX) {
            const uint4 *xp = reinterpret_cast<const uint4 *>(x + (size_t)b * xbs + (size_t)c * N);
#pragma unroll
            for (int it = 0; it < ITEMS; ++it)
                if (t + it * GROUP < nvec) Vec16<T>::unpack(ld_stream_16(xp + t + it * GROUP), xv[it]);
        }
        float g[ITEMS][V];
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            if (t + it * GROUP < nvec) {
                Vec16<T>::unpack(ld_stream_16(gp + t + it * GROUP), g[it]);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float pre = modulate(xv[it][j], mean, rstd, s, bb);   // same bits as the forward
                    const float gg = pre > 0.f ? g[it][j] : g[it][j] * slope;
                    const float xh = (xv[it][j] - mean) * rstd;
                    g[it][j] = gg;
                    sg += gg;
                    sgx += gg * xh;
                }
            }
        }
        sg = group_sum<GROUP>(sg, scratch);
        sgx = group_sum<GROUP>(sgx, scratch + 8);
        if (t == 0) {
            dbias[(size_t)b * dsbs + c] = sg;
            dscale[(size_t)b * dsbs + c] = sgx;
        }
        // dx = rstd * (dxh - mean(dxh) - xh * sum(dxh*xh)/(N-1)),  dxh = g*s   (unbiased variance)
        const float k1 = s * sg * inv_n, k2 = s * sgx * inv_nm1;
        uint4 *dp = kConstX ? nullptr : reinterpret_cast<uint4 *>(dx + bc * N);
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int vi = t + it * GROUP;
            if (vi < nvec) {
                float o[V];
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float xh = (xv[it][j] - mean) * rstd;
                    o[j] = rstd * (g[it][j] * s - k1 - xh * k2);
                    if (kConstX) acc[it][j] += o[j];
                }
                if (!kConstX) st_stream_16(dp + vi, Vec16<T>::pack(o));
            }
        }
    }
    if (kSplit) {
        const int warp = threadIdx.x / 32;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it)
#pragma unroll
            for (int j = 0; j < V; ++j) part[((warp * ITEMS + it) * V + j) * 32 + t] = acc[it][j];
        __syncthreads();
        if (warp != 0) return;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it)
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum += part[((w * ITEMS + it) * V + j) * 32 + t];     // fixed order
                acc[it][j] = sum;
            }
    }
    if (kConstX) {
        uint4 *dp = reinterpret_cast<uint4 *>(dx + (size_t)c * N);
#pragma unroll
        for (int it = 0; it < ITEMS; ++it)
            if (t + it * GROUP < nvec) dp[t + it * GROUP] = Vec16<T>::pack(acc[it]);
    }
}

template <typename T>
static int adain_fwd_dispatch(const void *x, const float *scale, const float *bias, void *y, float *mean, float *rstd, int B,
                              int C, int N, long long xbs, int sbs, float eps, float slope, int biased, cudaStream_t st)
{
    constexpr int V = Vec16<T>::N;
    const int BC = B * C;
    const T *xp = static_cast<const T *>(x);
    T *yp = static_cast<T *>(y);
    if (N <= 32 * 4 * V) {
        adain_fwd_kernel<T, 32, 4><<<(BC + 7) / 8, 256, 0, st>>>(xp, scale, bias, yp, mean, rstd, BC, C, N, xbs, sbs, eps, slope, biased);
    } else if (N <= 256 * 4 * V) {
        adain_fwd_kernel<T, 256, 4><<<BC, 256, 0, st>>>(xp, scale, bias, yp, mean, rstd, BC, C, N, xbs, sbs, eps, slope, biased);
    } else {
        return fail(HG_ERR_UNSUPPORTED, "hg_adain_act_fwd: N=%d exceeds the single-pass limit %d", N, 256 * 4 * V);
    }
    return check_launch("adain_fwd");
}

template <typename T>
static int adain_bwd_dispatch(const void *x, const void *dy, const float *scale, const float *bias, const float *mean,
                              const float *rstd, void *dx, float *dscale, float *dbias, int B, int C, int N, long long xbs,
                              int sbs, int dsbs, float slope, int biased, cudaStream_t st)
{
    constexpr int V = Vec16<T>::N;
    const T *xp = static_cast<const T *>(x), *gp = static_cast<const T *>(dy);
    T *dp = static_cast<T *>(dx);
    const bool cx = xbs == 0;
    const int n_inst = cx ? C : B * C;
    if (N <= 32 * 4 * V) {
        const int grid = (n_inst + 7) / 8;
        if (cx)                                   // one CTA per channel, the 8 warps split the batch
            adain_bwd_kernel<T, 32, 4, true><<<n_inst, 256, 0, st>>>(xp, gp, scale, bias, mean, rstd, dp, dscale, dbias, B, C, N, xbs, sbs, dsbs, slope, biased);
        else
            adain_bwd_kernel<T, 32, 4, false><<<grid, 256, 0, st>>>(xp, gp, scale, bias, mean, rstd, dp, dscale, dbias, B, C, N, xbs, sbs, dsbs, slope, biased);
    } else if (N <= 256 * 4 * V) {
        if (cx)
            adain_bwd_kernel<T, 256, 4, true><<<n_inst, 256, 0, st>>>(xp, gp, scale, bias, mean, rstd, dp, dscale, dbias, B, C, N, xbs, sbs, dsbs, slope, biased);
        else
            adain_bwd_kernel<T, 256, 4, false><<<n_inst, 256, 0, st>>>(xp, gp, scale, bias, mean, rstd, dp, dscale, dbias, B, C, N, xbs, sbs, dsbs, slope, biased);
    } else {
        return fail(HG_ERR_UNSUPPORTED, "hg_adain_act_bwd: N=%d exceeds the single-pass limit %d", N, 256 * 4 * V);
    }
    return check_launch("adain_bwd");
}

}  // namespace hg

using namespace hg;

extern "C" int hg_adain_act_fwd(const void *x, const float *scale, const float *bias, void *y, float *save_mean,
                                float *save_rstd, int batch, int channels, int n, long long x_batch_stride, int sb_stride,
                                float eps, float neg_slope, int biased_var, int dtype, void *stream)
{
    HG_REQUIRE(x && scale && bias && y && save_mean && save_rstd, HG_ERR_INVALID_ARG, "hg_adain_act_fwd: null pointer");
    HG_REQUIRE(batch > 0 && channels > 0 && n > 0, HG_ERR_INVALID_ARG, "hg_adain_act_fwd: dims must be positive");
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_adain_act_fwd: unknown dtype %d", dtype);
    const int vec = dtype == HG_F32 ? 4 : 8;
    HG_REQUIRE(n % vec == 0 && x_batch_stride % vec == 0, HG_ERR_UNSUPPORTED,
               "hg_adain_act_fwd: N and the batch stride must be multiples of %d elements (16 bytes)", vec);
    HG_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (long long)channels * n, HG_ERR_INVALID_ARG,
               "hg_adain_act_fwd: batch stride smaller than one sample");
    HG_REQUIRE(sb_stride >= channels, HG_ERR_INVALID_ARG, "hg_adain_act_fwd: sb_stride < channels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == HG_F32)
        return adain_fwd_dispatch<float>(x, scale, bias, y, save_mean, save_rstd, batch, channels, n, x_batch_stride, sb_stride, eps, neg_slope, biased_var, st);
    return adain_fwd_dispatch<__nv_bfloat16>(x, scale, bias, y, save_mean, save_rstd, batch, channels, n, x_batch_stride, sb_stride, eps, neg_slope, biased_var, st);
}

extern "C" int hg_adain_act_bwd(const void *x, const void *dy, const float *scale, const float *bias, const float *save_mean,
                                const float *save_rstd, void *dx, float *dscale, float *dbias, int batch, int channels, int n,
                                long long x_batch_stride, int sb_stride, int dsb_stride, float neg_slope, int biased_var,
                                int dtype, void *stream)
{
    HG_REQUIRE(x && dy && scale && bias && save_mean && save_rstd && dx && dscale && dbias, HG_ERR_INVALID_ARG,
               "hg_adain_act_bwd: null pointer");
    HG_REQUIRE(batch > 0 && channels > 0 && n > 0, HG_ERR_INVALID_ARG, "hg_adain_act_bwd: dims must be positive");
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_adain_act_bwd: unknown dtype %d", dtype);
    const int vec = dtype == HG_F32 ? 4 : 8;
    HG_REQUIRE(n % vec == 0 && x_batch_stride % vec == 0, HG_ERR_UNSUPPORTED,
               "hg_adain_act_bwd: N and the batch stride must be multiples of %d elements (16 bytes)", vec);
    HG_REQUIRE(sb_stride >= channels && dsb_stride >= channels, HG_ERR_INVALID_ARG, "hg_adain_act_bwd: stride < channels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == HG_F32)
        return adain_bwd_dispatch<float>(x, dy, scale, bias, save_mean, save_rstd, dx, dscale, dbias, batch, channels, n, x_batch_stride, sb_stride, dsb_stride, neg_slope, biased_var, st);
    return adain_bwd_dispatch<__nv_bfloat16>(x, dy, scale, bias, save_mean, save_rstd, dx, dscale, dbias, batch, channels, n, x_batch_stride, sb_stride, dsb_stride, neg_slope, biased_var, st);
}
