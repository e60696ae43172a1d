// Channels-last AdaIN / InstanceNorm (+ activation) for the bf16 pipeline.
//
// Same arithmetic as adain.cu (reference AdaIn, core/models/hologan_generator.py:333-345, fused with the ReLU
// of :41; with biased_var also the discriminator's InstanceNorm2d + LeakyReLU,
// core/models/hologan_discriminator.py:16-17,21-22) on the layouts the implicit-GEMM kernels produce / consume:
//   input  x : (B, Npos, P, C) bf16 -- the space-to-depth output of hg_convt_fwd (P parity classes, P = 1 for
//              a plain channels-last tensor), rows r = pos * P + cls
//   output y : (B, (2S)^d, C) bf16 plain channels-last: row (pos, cls) lands on its up-sampled pixel, i.e. the
//              depth-to-space shuffle is folded into the store addressing (zero extra passes).
//
// HBM-bound.  Every global access is a FULL ROW (C * 2 bytes contiguous, C/8 lanes x 16 bytes): a sample is
// cut into row chunks, one CTA per (chunk, sample) streams its rows with four independent 16-byte loads in
// flight per thread.
//   forward : stats kernel  -- per-channel sum(x - K), sum((x - K)^2) per chunk (K = the sample's first row, a
//                              pivot that removes the cancellation of E[x^2] - E[x]^2; partials ADD, so chunks
//                              merge in a fixed order without Welford bookkeeping)
//             apply kernel  -- merges the chunk partials, normalises / modulates / activates its rows (second
//                              read of x comes from L2) and stores to the up-sampled row
//   backward: sums kernel   -- sum(g), sum(g * xhat) per chunk, g = dy through the activation
//             apply kernel  -- dx = rstd * (g * s - s * sum(g) / N - xhat * s * sum(g * xhat) / Nvar)
// Small instances (a sample <= 64 KB: the discriminator's maps) run both phases in ONE kernel, one CTA per
// sample, the second pass served by L1.  Algorithmic HBM bytes: fwd 2 * B*N*C*2, bwd 3 * B*N*C*2.
//
// Cluster path (the normal case for the generator's AdaIN sites and the larger discriminator maps): a sample is
// spread over a thread-block CLUSTER of up to 8 CTAs; every thread parks its rows (<= 16 x 16 bytes) in shared
// memory with cp.async (private slots: no CTA barrier needed), the per-CTA partial sums are exchanged through
// distributed shared memory (cluster.map_shared_rank, fixed rank order -> deterministic), and the normalised rows
// are produced from the staged copy: ONE kernel, x is read from HBM exactly once (forward: 1 read + 1 write, backward: 2 reads + 1 write = the algorithmic traffic).
#include <cooperative_groups.h>

#include "hg_common.cuh"

namespace cg = cooperative_groups;

namespace hg {

constexpr int kClThreads = 256;
constexpr int kClBwdClusterThreads = 512;   // cluster backward: two tensors in registers -> twice the threads per CTA
constexpr int kClClusterMaxC = 512;         // channels the cluster path covers (per-CTA partials in static smem)
constexpr int kClClusterFwdRows = 16;       // rows (16-byte units) a thread stages: 64 KB per forward CTA
constexpr int kClClusterBwdRows = 8;        // per tensor: 2 x 64 KB per backward CTA
constexpr int kClUnroll = 8;          // forward: independent 16-byte loads in flight per thread
constexpr int kClUnrollBwd = 4;       // backward streams two tensors: 2 x 4 loads in flight
constexpr int kClMaxChunks = 32;

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

struct ClGeom;
__device__ __forceinline__ int mapped_row(int r, const ClGeom &g);

struct ClGeom {
    int C, N, Nvar, ndim, logS, logP;
    int s2d_out;        // plain (S x S) rows in, 2x2 space-to-depth row order out (classes == -4, see hg_adain_cl_fwd)
    int lanes;          // C / 8 threads per row
    int rows_per_pass;  // kClThreads / lanes
    int chunk_rows;     // rows per CTA (multiple of rows_per_pass)
    int chunks;
};

// Row of the output tensor (forward) / of dy (backward) that belongs to input row r.  Normally the depth-to-space shuffle
// of a transposed convolution's s2d output; with s2d_out the opposite direction: pixel (iy, ix) of a plain S x S map goes
// to row ((iy/2) * S/2 + ix/2) * 4 + (iy%2) * 2 + ix%2, the layout hg_conv5s2_fwd reads (discriminator blocks).
__device__ __forceinline__ int mapped_row(int r, const ClGeom &g)
{
    if (g.s2d_out) {
        const int ix = r & ((1 << g.logS) - 1), iy = r >> g.logS;
        return ((((iy >> 1) << (g.logS - 1)) + (ix >> 1)) << 2) + ((iy & 1) << 1) + (ix & 1);
    }
    return upsampled_row(r, g.ndim, g.logS, g.logP);
}

// Sum the two 8-float partials of all row slots of the CTA in a fixed order; result valid for every thread.
// Row slots that share a warp (lanes < 32) are folded with xor-shuffles first, the remaining <= 8 groups go
// through shared memory laid out [group][value][channel octet] (conflict-free).
template <int NT = kClThreads>
__device__ __forceinline__ void cta_rowslot_sum(float (&a)[8], float (&b)[8], float *red, int lanes, int rows_per_pass,
                                                int cs, int rs)
{
    for (int o = lanes; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
            b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = lanes < 32 ? NT / 32 : rows_per_pass;
    const int grp = lanes < 32 ? warp : rs;
    __syncthreads();
    if (lanes >= 32 || lane < lanes) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[(grp * 16 + j) * lanes + cs] = a[j];
            red[(grp * 16 + 8 + j) * lanes + cs] = b[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
    for (int k = 0; k < groups; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] += red[(k * 16 + j) * lanes + cs];
            b[j] += red[(k * 16 + 8 + j) * lanes + cs];
        }
    }
}

// ---- forward pieces -----------------------------------------------------------------------------------
// s1 += sum(x - K), s2 += sum((x - K)^2) over rows [r0, r1) of this thread's row slot
__device__ __forceinline__ void fwd_accumulate(const __nv_bfloat16 *__restrict__ xb, const ClGeom &g, int r0, int r1, int rs,
                                               const float (&piv)[8], float (&s1)[8], float (&s2)[8])
{
    for (int r = r0 + rs; r < r1; r += kClUnroll * g.rows_per_pass) {
        uint4 raw[kClUnroll];
#pragma unroll
        for (int u = 0; u < kClUnroll; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) raw[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)rr * g.C));
        }
#pragma unroll
        for (int u = 0; u < kClUnroll; ++u) {
            if (r + u * g.rows_per_pass < r1) {
                float f[8];
                unpack8(raw[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = f[j] - piv[j];
                    s1[j] += d;
                    s2[j] = fmaf(d, d, s2[j]);
                }
            }
        }
    }
}

// y = act(x * a + c) with a = scale * rstd, c = bias - mean * a: ONE fma per element (the statistics pass and the
// backward recompute the same a, c from the saved mean / rstd -> same bits, same activation mask).  The kernels are
// issue-bound (ncu: 0.5 instructions per scheduler cycle at 16 warps per SM), so every instruction per element counts.
__device__ __forceinline__ void fwd_apply(const __nv_bfloat16 *__restrict__ xb, __nv_bfloat16 *__restrict__ yb, const ClGeom &g,
                                          int r0, int r1, int rs, const float (&a)[8], const float (&c)[8], float slope)
{
    for (int r = r0 + rs; r < r1; r += kClUnroll * g.rows_per_pass) {
        uint4 raw[kClUnroll];
#pragma unroll
        for (int u = 0; u < kClUnroll; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) raw[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)rr * g.C));
        }
#pragma unroll
        for (int u = 0; u < kClUnroll; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) {
                float f[8];
                unpack8(raw[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float p = fmaf(f[j], a[j], c[j]);
                    f[j] = p > 0.f ? p : p * slope;
                }
                st_stream_16(yb + (size_t)mapped_row(rr, g) * g.C, pack8(f));
            }
        }
    }
}

// a = scale * rstd, c = bias - mean * a (exactly as the cluster kernels compute them)
__device__ __forceinline__ void affine_coeffs(const float *scale, const float *bias, int b, int sbs, int c0, const float (&mean)[8],
                                              const float (&rstd)[8], float (&a)[8], float (&c)[8])
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = __fmul_rn(scale ? scale[(size_t)b * sbs + c0 + j] : 1.f, rstd[j]);
        c[j] = __fsub_rn(bias ? bias[(size_t)b * sbs + c0 + j] : 0.f, __fmul_rn(mean[j], a[j]));
    }
}

// grid (chunks, B).  part[(b * chunks + chunk) * 2C + {0, C} + c]
__global__ void __launch_bounds__(kClThreads, 3) adain_cl_stats_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ part,
                                                                    ClGeom g)
{
    __shared__ float red[kClThreads * 16];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const __nv_bfloat16 *xb = x + (size_t)b * g.N * g.C + cs * 8;
    float piv[8], s1[8], s2[8];
    unpack8(__ldg(reinterpret_cast<const uint4 *>(xb)), piv);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int r0 = chunk * g.chunk_rows, r1 = min(g.N, r0 + g.chunk_rows);
    fwd_accumulate(xb, g, r0, r1, rs, piv, s1, s2);
    cta_rowslot_sum(s1, s2, red, g.lanes, g.rows_per_pass, cs, rs);
    if (rs == 0) {
        float *dst = part + ((size_t)b * g.chunks + chunk) * 2 * g.C + cs * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dst[j] = s1[j];
            dst[g.C + j] = s2[j];
        }
    }
}

// Statistics from the producing GEMM's epilogue (hg_convt_fwd_stats): stats[(g * ld + col) * 2 + {sum, sum of squares}] per
// 32-row group g and s2d column col = cls * C + c of the fp32 conv results.  One thread per (b, c) merges the
// groups x classes partials of its instance in a fixed order (Chan's parallel update on (mean, M2) -- each 32-element
// partial is centred on its own mean first, so nothing cancels across the instance) and writes mean / rstd.
// grid (C / 32, B), 256 threads = 32 channels x 8 slices of the partial list (8 loads in flight per thread, coalesced
// over the channels); the slices meet in shared memory and are merged in a fixed order.
__device__ __forceinline__ void chan_merge(float &mean, float &m2, float &n, float mw, float m2w, float nw)
{
    if (nw == 0.f) return;
    const float nn = n + nw, delta = mw - mean;
    mean += delta * (nw / nn);
    m2 += m2w + delta * delta * (n * nw / nn);
    n = nn;
}
__global__ void __launch_bounds__(256) adain_cl_stats_finalize_kernel(const float *__restrict__ stats, float *__restrict__ save_mean,
                                                                      float *__restrict__ save_rstd, int C, int P, int groups,
                                                                      int nvar, float eps)
{
    __shared__ float sm[8][3][33];
    const int cx = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx, b = blockIdx.y;
    const int ld = P * C, total = groups * P;           // partial q = g * P + p sits at float2 index (b * groups + g) * ld + p * C + c
    const float2 *base = reinterpret_cast<const float2 *>(stats) + (size_t)b * groups * ld + c;
    float mean = 0.f, m2 = 0.f, n = 0.f;
    if (c < C) {
        for (int q0 = slice; q0 < total; q0 += 64) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + 8 * u;
                if (q < total) v[u] = __ldg(base + (size_t)(q / P) * ld + (q % P) * C);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (q0 + 8 * u < total) {
                    const float mw = v[u].x * (1.f / 32.f);
                    chan_merge(mean, m2, n, mw, fmaxf(v[u].y - v[u].x * mw, 0.f), 32.f);
                }
            }
        }
    }
    sm[slice][0][cx] = mean; sm[slice][1][cx] = m2; sm[slice][2][cx] = n;
    __syncthreads();
    if (slice == 0 && c < C) {
        for (int k = 1; k < 8; ++k) chan_merge(mean, m2, n, sm[k][0][cx], sm[k][1][cx], sm[k][2][cx]);
        save_mean[(size_t)b * C + c] = mean;
        save_rstd[(size_t)b * C + c] = __frsqrt_rn(m2 / (float)nvar + eps);
    }
}

// Streaming normalise / modulate / activate pass with the statistics already known (save_mean / save_rstd)
__global__ void __launch_bounds__(kClThreads, 3) adain_cl_apply_stats_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                          const float *__restrict__ scale,
                                                                          const float *__restrict__ bias,
                                                                          const float *__restrict__ save_mean,
                                                                          const float *__restrict__ save_rstd,
                                                                          __nv_bfloat16 *__restrict__ y, ClGeom g, int sbs, float slope)
{
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    float mean[8], rstd[8], a[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mean[j] = save_mean[(size_t)b * g.C + cs * 8 + j];
        rstd[j] = save_rstd[(size_t)b * g.C + cs * 8 + j];
    }
    affine_coeffs(scale, bias, b, sbs, cs * 8, mean, rstd, a, c);
    const int r0 = chunk * g.chunk_rows, r1 = min(g.N, r0 + g.chunk_rows);
    fwd_apply(x + (size_t)b * g.N * g.C + cs * 8, y + (size_t)b * g.N * g.C + cs * 8, g, r0, r1, rs, a, c, slope);
}

// kFused: one CTA per sample computes the statistics itself (chunks == 1); else every CTA merges the chunk
// partials in `part` (fixed order).
template <bool kFused>
__global__ void __launch_bounds__(kClThreads, 3) adain_cl_apply_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                    const float *__restrict__ part,
                                                                    const float *__restrict__ scale,
                                                                    const float *__restrict__ bias,
                                                                    __nv_bfloat16 *__restrict__ y, float *__restrict__ save_mean,
                                                                    float *__restrict__ save_rstd, ClGeom g, int sbs, float eps,
                                                                    float slope)
{
    __shared__ float red[kFused ? kClThreads * 16 : 1];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const __nv_bfloat16 *xb = x + (size_t)b * g.N * g.C + cs * 8;
    float mean[8], rstd[8];
    if (kFused) {
        float piv[8], s1[8], s2[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(xb)), piv);
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        fwd_accumulate(xb, g, 0, g.N, rs, piv, s1, s2);
        cta_rowslot_sum(s1, s2, red, g.lanes, g.rows_per_pass, cs, rs);
        const float inv_n = 1.f / (float)g.N;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = s1[j] * inv_n;                      // mean - K
            mean[j] = piv[j] + d;
            // Nvar = N - 1: unbiased variance (:338), eps inside the rsqrt (:339); Nvar = N: InstanceNorm2d
            const float var = fmaxf(s2[j] - s1[j] * d, 0.f) / (float)g.Nvar;
            rstd[j] = __frsqrt_rn(var + eps);
        }
        if (rs == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                save_mean[(size_t)b * g.C + cs * 8 + j] = mean[j];
                save_rstd[(size_t)b * g.C + cs * 8 + j] = rstd[j];
            }
        }
    } else {
        // merge the chunk partials of adain_cl_stats_kernel (fixed chunk order: every CTA of the sample gets the
        // same bits); the sample's first chunk also publishes mean / rstd for the backward
        float piv[8], s1[8], s2[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(xb)), piv);
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        const float *src = part + (size_t)b * g.chunks * 2 * g.C + cs * 8;
        for (int k = 0; k < g.chunks; ++k) {
            const float4 *a4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C);
            const float4 *q4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C + g.C);
            const float4 a0 = a4[0], a1 = a4[1], q0 = q4[0], q1 = q4[1];
            s1[0] += a0.x; s1[1] += a0.y; s1[2] += a0.z; s1[3] += a0.w; s1[4] += a1.x; s1[5] += a1.y; s1[6] += a1.z; s1[7] += a1.w;
            s2[0] += q0.x; s2[1] += q0.y; s2[2] += q0.z; s2[3] += q0.w; s2[4] += q1.x; s2[5] += q1.y; s2[6] += q1.z; s2[7] += q1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = s1[j] / (float)g.N;                 // mean - K
            mean[j] = piv[j] + d;
            const float var = fmaxf(s2[j] - s1[j] * d, 0.f) / (float)g.Nvar;
            rstd[j] = __frsqrt_rn(var + eps);
        }
        if (chunk == 0 && rs == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                save_mean[(size_t)b * g.C + cs * 8 + j] = mean[j];
                save_rstd[(size_t)b * g.C + cs * 8 + j] = rstd[j];
            }
        }
    }
    float a[8], c[8];
    affine_coeffs(scale, bias, b, sbs, cs * 8, mean, rstd, a, c);
    const int r0 = kFused ? 0 : chunk * g.chunk_rows, r1 = kFused ? g.N : min(g.N, r0 + g.chunk_rows);
    fwd_apply(xb, y + (size_t)b * g.N * g.C + cs * 8, g, r0, r1, rs, a, c, slope);
}

// ---- backward pieces ----------------------------------------------------------------------------------
struct ClStyle {
    float mean[8], rstd[8], a[8], c[8];                 // a = scale * rstd, c = bias - mean * a (the forward's coefficients)
};

__device__ __forceinline__ void load_style(ClStyle &st, const float *scale, const float *bias, const float *save_mean,
                                           const float *save_rstd, int b, int C, int sbs, int c0)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        st.mean[j] = save_mean[(size_t)b * C + c0 + j];
        st.rstd[j] = save_rstd[(size_t)b * C + c0 + j];
    }
    affine_coeffs(scale, bias, b, sbs, c0, st.mean, st.rstd, st.a, st.c);
}

// sg += sum(g), sgx += sum(g * (x - mean)) with g = dy through the activation (mask from the forward's x * a + c)
__device__ __forceinline__ void bwd_accumulate(const __nv_bfloat16 *__restrict__ xb, const __nv_bfloat16 *__restrict__ gb,
                                               const ClGeom &g, int r0, int r1, int rs, const ClStyle &st, float slope,
                                               float (&sg)[8], float (&sgx)[8])
{
    for (int r = r0 + rs; r < r1; r += kClUnrollBwd * g.rows_per_pass) {
        uint4 xr[kClUnrollBwd], gr[kClUnrollBwd];
#pragma unroll
        for (int u = 0; u < kClUnrollBwd; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) {
                xr[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)rr * g.C));
                gr[u] = __ldg(reinterpret_cast<const uint4 *>(gb + (size_t)mapped_row(rr, g) * g.C));
            }
        }
#pragma unroll
        for (int u = 0; u < kClUnrollBwd; ++u) {
            if (r + u * g.rows_per_pass < r1) {
                float xf[8], gf[8];
                unpack8(xr[u], xf);
                unpack8(gr[u], gf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
                    sg[j] += gg;
                    sgx[j] = fmaf(gg, xf[j] - st.mean[j], sgx[j]);
                }
            }
        }
    }
}

// dx = rstd * (g s - s sum(g) / N - xhat s sum(g xhat) / Nvar) = g * a - (x * c2 + c1)
__device__ __forceinline__ void bwd_apply(const __nv_bfloat16 *__restrict__ xb, const __nv_bfloat16 *__restrict__ gb,
                                          __nv_bfloat16 *__restrict__ db, const ClGeom &g, int r0, int r1, int rs,
                                          const ClStyle &st, float slope, const float (&c1)[8], const float (&c2)[8])
{
    for (int r = r0 + rs; r < r1; r += kClUnrollBwd * g.rows_per_pass) {
        uint4 xr[kClUnrollBwd], gr[kClUnrollBwd];
#pragma unroll
        for (int u = 0; u < kClUnrollBwd; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) {
                xr[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)rr * g.C));
                gr[u] = __ldg(reinterpret_cast<const uint4 *>(gb + (size_t)mapped_row(rr, g) * g.C));
            }
        }
#pragma unroll
        for (int u = 0; u < kClUnrollBwd; ++u) {
            const int rr = r + u * g.rows_per_pass;
            if (rr < r1) {
                float xf[8], gf[8];
                unpack8(xr[u], xf);
                unpack8(gr[u], gf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
                    xf[j] = fmaf(gg, st.a[j], -fmaf(xf[j], c2[j], c1[j]));
                }
                st_stream_16(db + (size_t)rr * g.C, pack8(xf));
            }
        }
    }
}

__global__ void __launch_bounds__(kClThreads, 3) adain_cl_bwd_sums_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                       const __nv_bfloat16 *__restrict__ dy,
                                                                       const float *__restrict__ scale,
                                                                       const float *__restrict__ bias,
                                                                       const float *__restrict__ save_mean,
                                                                       const float *__restrict__ save_rstd,
                                                                       float *__restrict__ part, ClGeom g, int sbs, float slope)
{
    __shared__ float red[kClThreads * 16];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    ClStyle st;
    load_style(st, scale, bias, save_mean, save_rstd, b, g.C, sbs, cs * 8);
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    const int r0 = chunk * g.chunk_rows, r1 = min(g.N, r0 + g.chunk_rows);
    bwd_accumulate(x + base, dy + base, g, r0, r1, rs, st, slope, sg, sgx);
    cta_rowslot_sum(sg, sgx, red, g.lanes, g.rows_per_pass, cs, rs);
    if (rs == 0) {
        float *dst = part + ((size_t)b * g.chunks + chunk) * 2 * g.C + cs * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dst[j] = sg[j];
            dst[g.C + j] = sgx[j];
        }
    }
}

template <bool kFused>
__global__ void __launch_bounds__(kClThreads, 3) adain_cl_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                        const __nv_bfloat16 *__restrict__ dy,
                                                                        const float *__restrict__ part,
                                                                        const float *__restrict__ scale,
                                                                        const float *__restrict__ bias,
                                                                        const float *__restrict__ save_mean,
                                                                        const float *__restrict__ save_rstd,
                                                                        __nv_bfloat16 *__restrict__ dx, float *__restrict__ dscale,
                                                                        float *__restrict__ dbias, ClGeom g, int sbs, int dsbs,
                                                                        float slope)
{
    __shared__ float red[kFused ? kClThreads * 16 : 1];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    ClStyle st;
    load_style(st, scale, bias, save_mean, save_rstd, b, g.C, sbs, cs * 8);
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    if (kFused) {
        bwd_accumulate(x + base, dy + base, g, 0, g.N, rs, st, slope, sg, sgx);
        cta_rowslot_sum(sg, sgx, red, g.lanes, g.rows_per_pass, cs, rs);
    } else {                                                    // merge the chunk partials of adain_cl_bwd_sums_kernel
        const float *src = part + (size_t)b * g.chunks * 2 * g.C + cs * 8;
        for (int k = 0; k < g.chunks; ++k) {
            const float4 *a4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C);
            const float4 *q4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C + g.C);
            const float4 a0 = a4[0], a1 = a4[1], q0 = q4[0], q1 = q4[1];
            sg[0] += a0.x; sg[1] += a0.y; sg[2] += a0.z; sg[3] += a0.w; sg[4] += a1.x; sg[5] += a1.y; sg[6] += a1.z; sg[7] += a1.w;
            sgx[0] += q0.x; sgx[1] += q0.y; sgx[2] += q0.z; sgx[3] += q0.w; sgx[4] += q1.x; sgx[5] += q1.y; sgx[6] += q1.z; sgx[7] += q1.w;
        }
    }
    const bool publish = (kFused || chunk == 0) && rs == 0 && dscale && dbias;
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sgxh = sgx[j] * st.rstd[j];                         // sum(g * xhat)
        if (publish) {
            dbias[(size_t)b * dsbs + cs * 8 + j] = sg[j];
            dscale[(size_t)b * dsbs + cs * 8 + j] = sgxh;
        }
        c2[j] = st.a[j] * st.rstd[j] * (sgxh / (float)g.Nvar);          // s rstd^2 sum(g xhat) / Nvar
        c1[j] = st.a[j] * (sg[j] / (float)g.N) - st.mean[j] * c2[j];
    }
    const int r0 = kFused ? 0 : chunk * g.chunk_rows, r1 = kFused ? g.N : min(g.N, r0 + g.chunk_rows);
    bwd_apply(x + base, dy + base, dx + base, g, r0, r1, rs, st, slope, c1, c2);
}


// ---- small instances, register resident -------------------------------------------------------------------
// One CTA per sample and at most kSmallRows rows per thread (the discriminator's three InstanceNorm + LeakyReLU sites:
// 256 x 128, 64 x 256, 16 x 512 at B = 64).  The fused kernels above walk their rows twice in batches of 4-8 loads, i.e.
// 4-8 dependent memory round trips on a grid that cannot hide them (64 CTAs); here every row of the thread is loaded ONCE,
// all loads in flight together, and both the statistics and the normalise / gradient pass run from registers.  Same
// arithmetic in the same order as the fused kernels (bit-identical; tests/test_gpu_channels_last.py).
constexpr int kSmallRows = 16;

__global__ void __launch_bounds__(kClThreads, 1) adain_cl_small_fwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                         const float *__restrict__ scale,
                                                                         const float *__restrict__ bias,
                                                                         __nv_bfloat16 *__restrict__ y, float *__restrict__ save_mean,
                                                                         float *__restrict__ save_rstd, ClGeom g, int sbs, float eps,
                                                                         float slope)
{
    __shared__ float red[kClThreads * 16];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.x;
    const __nv_bfloat16 *xb = x + (size_t)b * g.N * g.C + cs * 8;
    uint4 raw[kSmallRows];
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        const int rr = rs + u * g.rows_per_pass;
        if (rr < g.N) raw[u] = __ldg(reinterpret_cast<const uint4 *>(xb + (size_t)rr * g.C));
    }
    float piv[8], s1[8], s2[8];
    unpack8(__ldg(reinterpret_cast<const uint4 *>(xb)), piv);           // pivot K = the sample's first row
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        if (rs + u * g.rows_per_pass < g.N) {
            float f[8];
            unpack8(raw[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = f[j] - piv[j];
                s1[j] += d;
                s2[j] = fmaf(d, d, s2[j]);
            }
        }
    }
    cta_rowslot_sum(s1, s2, red, g.lanes, g.rows_per_pass, cs, rs);
    float mean[8], rstd[8];
    const float inv_n = 1.f / (float)g.N;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d = s1[j] * inv_n;
        mean[j] = piv[j] + d;
        const float var = fmaxf(s2[j] - s1[j] * d, 0.f) / (float)g.Nvar;
        rstd[j] = __frsqrt_rn(var + eps);
    }
    if (rs == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            save_mean[(size_t)b * g.C + cs * 8 + j] = mean[j];
            save_rstd[(size_t)b * g.C + cs * 8 + j] = rstd[j];
        }
    }
    float a[8], c[8];
    affine_coeffs(scale, bias, b, sbs, cs * 8, mean, rstd, a, c);
    __nv_bfloat16 *yb = y + (size_t)b * g.N * g.C + cs * 8;
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        const int rr = rs + u * g.rows_per_pass;
        if (rr < g.N) {
            float f[8];
            unpack8(raw[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = fmaf(f[j], a[j], c[j]);
                f[j] = p > 0.f ? p : p * slope;
            }
            st_stream_16(yb + (size_t)mapped_row(rr, g) * g.C, pack8(f));
        }
    }
}

__global__ void __launch_bounds__(kClThreads, 1) adain_cl_small_bwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                         const __nv_bfloat16 *__restrict__ dy,
                                                                         const float *__restrict__ scale,
                                                                         const float *__restrict__ bias,
                                                                         const float *__restrict__ save_mean,
                                                                         const float *__restrict__ save_rstd,
                                                                         __nv_bfloat16 *__restrict__ dx, float *__restrict__ dscale,
                                                                         float *__restrict__ dbias, ClGeom g, int sbs, int dsbs,
                                                                         float slope)
{
    __shared__ float red[kClThreads * 16];
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.x;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    uint4 xr[kSmallRows], gr[kSmallRows];
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        const int rr = rs + u * g.rows_per_pass;
        if (rr < g.N) {
            xr[u] = __ldg(reinterpret_cast<const uint4 *>(x + base + (size_t)rr * g.C));
            gr[u] = __ldg(reinterpret_cast<const uint4 *>(dy + base + (size_t)mapped_row(rr, g) * g.C));
        }
    }
    ClStyle st;
    load_style(st, scale, bias, save_mean, save_rstd, b, g.C, sbs, cs * 8);
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        if (rs + u * g.rows_per_pass < g.N) {
            float xf[8], gf[8];
            unpack8(xr[u], xf);
            unpack8(gr[u], gf);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
                sg[j] += gg;
                sgx[j] = fmaf(gg, xf[j] - st.mean[j], sgx[j]);
            }
        }
    }
    cta_rowslot_sum(sg, sgx, red, g.lanes, g.rows_per_pass, cs, rs);
    const bool publish = rs == 0 && dscale && dbias;
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sgxh = sgx[j] * st.rstd[j];                         // sum(g * xhat)
        if (publish) {
            dbias[(size_t)b * dsbs + cs * 8 + j] = sg[j];
            dscale[(size_t)b * dsbs + cs * 8 + j] = sgxh;
        }
        c2[j] = st.a[j] * st.rstd[j] * (sgxh / (float)g.Nvar);
        c1[j] = st.a[j] * (sg[j] / (float)g.N) - st.mean[j] * c2[j];
    }
#pragma unroll
    for (int u = 0; u < kSmallRows; ++u) {
        const int rr = rs + u * g.rows_per_pass;
        if (rr < g.N) {
            float xf[8], gf[8];
            unpack8(xr[u], xf);
            unpack8(gr[u], gf);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
                xf[j] = fmaf(gg, st.a[j], -fmaf(xf[j], c2[j], c1[j]));
            }
            st_stream_16(dx + base + (size_t)rr * g.C, pack8(xf));
        }
    }
}

// ---- ring path (the two backward passes of the chunked kernels) ------------------------------------------
// The register-staged loops above alternate between a load phase and a compute phase: ncu shows both backward kernels
// waiting on long-scoreboard stalls at ~30 % of the issue slots and ~2.3 TB/s (profiles/r02x_ncu_full_summary.txt).
// Here every thread streams its rows through a PRIVATE ring of shared-memory slots filled by cp.async: kRingS - 1 stages
// of kRingUN rows x 2 tensors x 16 bytes stay in flight per thread for the whole loop, with no register cost and no
// cross-thread synchronisation (a thread only reads the slots it filled itself; cp.async.wait_group orders them).
constexpr int kRingUN = 2, kRingS = 4;
constexpr size_t kRingBytes = (size_t)kRingS * kRingUN * 2 * kClThreads * 16;        // 64 KB per CTA, three CTAs per SM

template <typename Body>
__device__ __forceinline__ void ring_stream2(uint4 *ring, const __nv_bfloat16 *__restrict__ xb, const __nv_bfloat16 *__restrict__ gb,
                                             const ClGeom &g, int r0, int r1, int rs, Body body)
{
    uint4 *col = ring + threadIdx.x;                    // [stage][u][tensor][thread]
    const int step = g.rows_per_pass;
    int ri = r0 + rs;
    auto issue = [&](int stage) {
#pragma unroll
        for (int u = 0; u < kRingUN; ++u) {
            const int rr = ri + u * step;
            const bool ok = rr < r1;
            cp_async_16_zfill(col + ((stage * kRingUN + u) * 2 + 0) * kClThreads, ok ? xb + (size_t)rr * g.C : xb, ok);
            cp_async_16_zfill(col + ((stage * kRingUN + u) * 2 + 1) * kClThreads, ok ? gb + (size_t)mapped_row(rr, g) * g.C : gb, ok);
        }
        ri += kRingUN * step;
        cp_async_commit();
    };
#pragma unroll
    for (int sgi = 0; sgi < kRingS - 1; ++sgi) issue(sgi);
    int stage = 0;
    for (int r = r0 + rs; r < r1; r += kRingUN * step) {
        int nxt = stage + kRingS - 1;
        if (nxt >= kRingS) nxt -= kRingS;
        issue(nxt);                                     // past the end: zero-size copies, the group still counts
        cp_async_wait_group<kRingS - 1>();
#pragma unroll
        for (int u = 0; u < kRingUN; ++u) {
            const int rr = r + u * step;
            if (rr < r1) body(rr, col[((stage * kRingUN + u) * 2 + 0) * kClThreads], col[((stage * kRingUN + u) * 2 + 1) * kClThreads]);
        }
        if (++stage == kRingS) stage = 0;
    }
    cp_async_wait_group<0>();
}

__global__ void __launch_bounds__(kClThreads, 3) adain_cl_bwd_sums_ring_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                            const __nv_bfloat16 *__restrict__ dy,
                                                                            const float *__restrict__ scale,
                                                                            const float *__restrict__ bias,
                                                                            const float *__restrict__ save_mean,
                                                                            const float *__restrict__ save_rstd,
                                                                            float *__restrict__ part, ClGeom g, int sbs, float slope)
{
    extern __shared__ __align__(16) unsigned char cl_dyn[];
    uint4 *ring = reinterpret_cast<uint4 *>(cl_dyn);
    float *red = reinterpret_cast<float *>(cl_dyn);     // reused after the stream (cta_rowslot_sum synchronises first)
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    ClStyle st;
    load_style(st, scale, bias, save_mean, save_rstd, b, g.C, sbs, cs * 8);
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    const int r0 = chunk * g.chunk_rows, r1 = min(g.N, r0 + g.chunk_rows);
    ring_stream2(ring, x + base, dy + base, g, r0, r1, rs, [&](int, const uint4 &xr, const uint4 &gr) {
        float xf[8], gf[8];
        unpack8(xr, xf);
        unpack8(gr, gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
            sg[j] += gg;
            sgx[j] = fmaf(gg, xf[j] - st.mean[j], sgx[j]);
        }
    });
    cta_rowslot_sum(sg, sgx, red, g.lanes, g.rows_per_pass, cs, rs);
    if (rs == 0) {
        float *dst = part + ((size_t)b * g.chunks + chunk) * 2 * g.C + cs * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dst[j] = sg[j];
            dst[g.C + j] = sgx[j];
        }
    }
}

__global__ void __launch_bounds__(kClThreads, 3) adain_cl_bwd_apply_ring_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                             const __nv_bfloat16 *__restrict__ dy,
                                                                             const float *__restrict__ part,
                                                                             const float *__restrict__ scale,
                                                                             const float *__restrict__ bias,
                                                                             const float *__restrict__ save_mean,
                                                                             const float *__restrict__ save_rstd,
                                                                             __nv_bfloat16 *__restrict__ dx, float *__restrict__ dscale,
                                                                             float *__restrict__ dbias, ClGeom g, int sbs, int dsbs,
                                                                             float slope)
{
    extern __shared__ __align__(16) unsigned char cl_dyn[];
    uint4 *ring = reinterpret_cast<uint4 *>(cl_dyn);
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    ClStyle st;
    load_style(st, scale, bias, save_mean, save_rstd, b, g.C, sbs, cs * 8);
    float sg[8], sgx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    const float *src = part + (size_t)b * g.chunks * 2 * g.C + cs * 8;      // chunk partials of the sums kernel, fixed order
    for (int k = 0; k < g.chunks; ++k) {
        const float4 *a4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C);
        const float4 *q4 = reinterpret_cast<const float4 *>(src + (size_t)k * 2 * g.C + g.C);
        const float4 a0 = a4[0], a1 = a4[1], q0 = q4[0], q1 = q4[1];
        sg[0] += a0.x; sg[1] += a0.y; sg[2] += a0.z; sg[3] += a0.w; sg[4] += a1.x; sg[5] += a1.y; sg[6] += a1.z; sg[7] += a1.w;
        sgx[0] += q0.x; sgx[1] += q0.y; sgx[2] += q0.z; sgx[3] += q0.w; sgx[4] += q1.x; sgx[5] += q1.y; sgx[6] += q1.z; sgx[7] += q1.w;
    }
    const bool publish = chunk == 0 && rs == 0 && dscale && dbias;
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sgxh = sgx[j] * st.rstd[j];                         // sum(g * xhat)
        if (publish) {
            dbias[(size_t)b * dsbs + cs * 8 + j] = sg[j];
            dscale[(size_t)b * dsbs + cs * 8 + j] = sgxh;
        }
        c2[j] = st.a[j] * st.rstd[j] * (sgxh / (float)g.Nvar);
        c1[j] = st.a[j] * (sg[j] / (float)g.N) - st.mean[j] * c2[j];
    }
    const int r0 = chunk * g.chunk_rows, r1 = min(g.N, r0 + g.chunk_rows);
    __nv_bfloat16 *db = dx + base;
    ring_stream2(ring, x + base, dy + base, g, r0, r1, rs, [&](int rr, const uint4 &xr, const uint4 &gr) {
        float xf[8], gf[8];
        unpack8(xr, xf);
        unpack8(gr, gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gg = fmaf(xf[j], st.a[j], st.c[j]) > 0.f ? gf[j] : gf[j] * slope;
            xf[j] = fmaf(gg, st.a[j], -fmaf(xf[j], c2[j], c1[j]));
        }
        st_stream_16(db + (size_t)rr * g.C, pack8(xf));
    });
}

// ---- cluster path -------------------------------------------------------------------------------------
// grid (CS, B), cluster (CS, 1, 1): CTA `blockIdx.x` of the cluster owns rows [blockIdx.x * U * rpp, +U * rpp) of
// sample blockIdx.y; thread (rs, cs) holds rows r0 + rs + u * rpp, u < U, of channel octet cs in registers.

// Publish this CTA's two 8-float partials per channel octet in `cpart`, sum the cluster's in rank order (every CTA /
// thread ends up with the same bits).  Only C / 2 threads touch remote shared memory (one float4 column each, the
// <= 8 remote loads in flight together); the totals go through the local `ctot` to the rest of the CTA -- a first
// version in which every thread read every peer moved more bytes over the SM-to-SM network than the kernel reads
// from HBM.  The caller must cluster.sync() once more before the CTA exits (peers may still be reading cpart).
__device__ __forceinline__ void cluster_octet_sum(cg::cluster_group &cluster, float (&a)[8], float (&b)[8], float *cpart,
                                                  float *ctot, int C, int cs, int rs)
{
    if (rs == 0) {
        float4 *pa = reinterpret_cast<float4 *>(cpart + cs * 8), *pb = reinterpret_cast<float4 *>(cpart + C + cs * 8);
        pa[0] = make_float4(a[0], a[1], a[2], a[3]); pa[1] = make_float4(a[4], a[5], a[6], a[7]);
        pb[0] = make_float4(b[0], b[1], b[2], b[3]); pb[1] = make_float4(b[4], b[5], b[6], b[7]);
    }
    cluster.sync();
    const unsigned n = cluster.num_blocks();
    for (int i = threadIdx.x; i < C / 2; i += blockDim.x) {             // 2C floats = C / 2 float4 columns
        float4 v[8];
#pragma unroll
        for (unsigned r = 0; r < 8; ++r)
            if (r < n) v[r] = reinterpret_cast<const float4 *>(cluster.map_shared_rank(cpart, r))[i];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (unsigned r = 0; r < 8; ++r)
            if (r < n) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        reinterpret_cast<float4 *>(ctot)[i] = acc;
    }
    __syncthreads();
    const float4 *pa = reinterpret_cast<const float4 *>(ctot + cs * 8), *pb = reinterpret_cast<const float4 *>(ctot + C + cs * 8);
    const float4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only): the rows a thread owns wait in shared memory,
// slot [u][tid], instead of in registers; only the issuing thread reads them back, so cp.async.wait_all is the only
// synchronisation they need.
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(kClThreads, 2) adain_cl_cluster_fwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                             const float *__restrict__ scale,
                                                                             const float *__restrict__ bias,
                                                                             __nv_bfloat16 *__restrict__ y,
                                                                             float *__restrict__ save_mean,
                                                                             float *__restrict__ save_rstd, ClGeom g, int U,
                                                                             int sbs, float eps, float slope)
{
    extern __shared__ __align__(16) unsigned char cl_dyn[];
    uint4 *tile = reinterpret_cast<uint4 *>(cl_dyn) + threadIdx.x;      // [U][kClThreads], this thread's column
    __shared__ float red[kClThreads * 16];
    __shared__ __align__(16) float cpart[2 * kClClusterMaxC], ctot[2 * kClClusterMaxC];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y;
    const __nv_bfloat16 *xb = x + (size_t)b * g.N * g.C + cs * 8;
    const int r0 = blockIdx.x * (U * g.rows_per_pass) + rs;
    for (int u = 0; u < U; ++u) cp_async_16(tile + u * kClThreads, xb + (size_t)(r0 + u * g.rows_per_pass) * g.C);
    float piv[8], s1[8], s2[8];
    unpack8(__ldg(reinterpret_cast<const uint4 *>(xb)), piv);           // pivot K = the sample's first row
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    cp_async_wait_all();
#pragma unroll 4
    for (int u = 0; u < U; ++u) {
        float f[8];
        unpack8(tile[u * kClThreads], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = f[j] - piv[j];
            s1[j] += d;
            s2[j] = fmaf(d, d, s2[j]);
        }
    }
    cta_rowslot_sum<kClThreads>(s1, s2, red, g.lanes, g.rows_per_pass, cs, rs);
    cluster_octet_sum(cluster, s1, s2, cpart, ctot, g.C, cs, rs);
    // y = act(x * a + c),  a = scale * rstd,  c = bias - mean * a; the cluster backward recomputes the same a, c
    // from the saved statistics (same bits, same activation mask)
    float a[8], c[8];
    const float inv_n = 1.f / (float)g.N;
    const bool publish = blockIdx.x == 0 && rs == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d = s1[j] * inv_n;                                  // mean - K
        const float mean = piv[j] + d;
        // Nvar = N - 1: unbiased variance (:338), eps inside the rsqrt (:339); Nvar = N: InstanceNorm2d
        const float var = fmaxf(s2[j] - s1[j] * d, 0.f) / (float)g.Nvar;
        const float rstd = __frsqrt_rn(var + eps);
        if (publish) {
            save_mean[(size_t)b * g.C + cs * 8 + j] = mean;
            save_rstd[(size_t)b * g.C + cs * 8 + j] = rstd;
        }
        a[j] = __fmul_rn(scale ? scale[(size_t)b * sbs + cs * 8 + j] : 1.f, rstd);
        c[j] = __fsub_rn(bias ? bias[(size_t)b * sbs + cs * 8 + j] : 0.f, __fmul_rn(mean, a[j]));
    }
    __nv_bfloat16 *yb = y + (size_t)b * g.N * g.C + cs * 8;
#pragma unroll 4
    for (int u = 0; u < U; ++u) {
        float f[8];
        unpack8(tile[u * kClThreads], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p = fmaf(f[j], a[j], c[j]);
            f[j] = p > 0.f ? p : p * slope;
        }
        st_stream_16(yb + (size_t)mapped_row(r0 + u * g.rows_per_pass, g) * g.C, pack8(f));
    }
    cluster.sync();                                                     // cpart stays alive until every peer has read it
}

__global__ void __launch_bounds__(kClBwdClusterThreads, 1) adain_cl_cluster_bwd_kernel(
    const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ dy, const float *__restrict__ scale,
    const float *__restrict__ bias, const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
    __nv_bfloat16 *__restrict__ dx, float *__restrict__ dscale, float *__restrict__ dbias, ClGeom g, int U, int sbs, int dsbs,
    float slope)
{
    constexpr int NT = kClBwdClusterThreads;
    extern __shared__ __align__(16) unsigned char cl_dyn[];
    uint4 *xt = reinterpret_cast<uint4 *>(cl_dyn) + threadIdx.x;        // [U][NT] rows of x, then [U][NT] rows of dy
    uint4 *gt = xt + U * NT;
    __shared__ float red[NT * 16];
    __shared__ __align__(16) float cpart[2 * kClClusterMaxC], ctot[2 * kClClusterMaxC];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = threadIdx.x % g.lanes, rs = threadIdx.x / g.lanes;
    const int b = blockIdx.y;
    const size_t base = (size_t)b * g.N * g.C + cs * 8;
    const int r0 = blockIdx.x * (U * g.rows_per_pass) + rs;
    for (int u = 0; u < U; ++u) {
        const int rr = r0 + u * g.rows_per_pass;
        cp_async_16(xt + u * NT, x + base + (size_t)rr * g.C);
        cp_async_16(gt + u * NT, dy + base + (size_t)mapped_row(rr, g) * g.C);
    }
    // phase 1 constants: a, c (the forward's pre-activation x * a + c -> activation mask) and the mean
    float a[8], c[8], mean[8];
    const float *pm = save_mean + (size_t)b * g.C + cs * 8, *pr = save_rstd + (size_t)b * g.C + cs * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mean[j] = pm[j];
        a[j] = __fmul_rn(scale ? scale[(size_t)b * sbs + cs * 8 + j] : 1.f, pr[j]);
        c[j] = __fsub_rn(bias ? bias[(size_t)b * sbs + cs * 8 + j] : 0.f, __fmul_rn(mean[j], a[j]));
    }
    float sg[8], sgx[8];                                                // sum(g), sum(g * (x - mean))
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
    cp_async_wait_all();
#pragma unroll 2
    for (int u = 0; u < U; ++u) {
        float xf[8], gf[8];
        unpack8(xt[u * NT], xf);
        unpack8(gt[u * NT], gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gg = fmaf(xf[j], a[j], c[j]) > 0.f ? gf[j] : gf[j] * slope;
            sg[j] += gg;
            sgx[j] = fmaf(gg, xf[j] - mean[j], sgx[j]);
        }
    }
    cta_rowslot_sum<NT>(sg, sgx, red, g.lanes, g.rows_per_pass, cs, rs);
    cluster_octet_sum(cluster, sg, sgx, cpart, ctot, g.C, cs, rs);
    // phase 2 constants: dx = rstd * (g s - s sum(g) / N - xhat s sum(g xhat) / Nvar) = g * a - (x * c2 + c1)
    const bool publish = blockIdx.x == 0 && rs == 0 && dscale && dbias;
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float rstd = pr[j];
        const float sgxh = sgx[j] * rstd;                               // sum(g * xhat)
        if (publish) {
            dbias[(size_t)b * dsbs + cs * 8 + j] = sg[j];
            dscale[(size_t)b * dsbs + cs * 8 + j] = sgxh;
        }
        c2[j] = a[j] * rstd * (sgxh / (float)g.Nvar);                   // s rstd^2 sum(g xhat) / Nvar
        c1[j] = a[j] * (sg[j] / (float)g.N) - mean[j] * c2[j];
    }
#pragma unroll 2
    for (int u = 0; u < U; ++u) {
        float xf[8], gf[8];
        unpack8(xt[u * NT], xf);
        unpack8(gt[u * NT], gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gg = fmaf(xf[j], a[j], c[j]) > 0.f ? gf[j] : gf[j] * slope;
            xf[j] = fmaf(gg, a[j], -fmaf(xf[j], c2[j], c1[j]));
        }
        st_stream_16(dx + base + (size_t)(r0 + u * g.rows_per_pass) * g.C, pack8(xf));
    }
    cluster.sync();
}

static int ilog2(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

}  // namespace hg

using namespace hg;

static int cl_geom(const char *who, int batch, int channels, int ndim, int size, int classes, int biased_var, ClGeom &g)
{
    g.s2d_out = 0;
    if (classes == -4) {        // plain channels-last input, space-to-depth output order (2-D, even size)
        HG_REQUIRE(ndim == 2 && size >= 2, HG_ERR_UNSUPPORTED, "%s: classes == -4 (s2d output) needs ndim 2 and size >= 2", who);
        g.s2d_out = 1;
        classes = 1;
    }
    HG_REQUIRE(batch > 0 && channels > 0 && size > 0 && classes > 0, HG_ERR_INVALID_ARG, "%s: dims must be positive", who);
    HG_REQUIRE(batch <= 65535, HG_ERR_UNSUPPORTED, "%s: batch > 65535", who);
    HG_REQUIRE(ndim == 2 || ndim == 3, HG_ERR_INVALID_ARG, "%s: ndim must be 2 or 3", who);
    g.logS = ilog2(size);
    g.logP = ilog2(classes);
    HG_REQUIRE(g.logS >= 0 && (classes == 1 || classes == (1 << ndim)), HG_ERR_UNSUPPORTED,
               "%s: size must be a power of two and classes 1 or 2^ndim", who);
    const int lanes = channels / 8;
    HG_REQUIRE(channels % 8 == 0 && lanes <= kClThreads && ilog2(lanes) >= 0, HG_ERR_UNSUPPORTED,
               "%s: channels must be 8 * 2^k <= %d (got %d)", who, 8 * kClThreads, channels);
    long long rows = classes;
    for (int i = 0; i < ndim; ++i) rows *= size;
    HG_REQUIRE(rows >= 2 && rows <= (1 << 24), HG_ERR_UNSUPPORTED, "%s: %lld rows per instance not supported", who, rows);
    g.C = channels; g.N = (int)rows; g.Nvar = biased_var ? g.N : g.N - 1; g.ndim = ndim;
    g.lanes = lanes; g.rows_per_pass = kClThreads / lanes;
    // small instance (<= 64 KB): one CTA per sample does both phases; else ~16 chunks per sample
    const long long bytes = rows * channels * 2;
    // else cut every sample into chunks so that the grid is one wave of 3 CTAs per SM, >= kClUnroll rows per thread each
    int chunks = 1;
    if (bytes > 64 * 1024) {
        chunks = (3 * sm_count()) / batch;             // one wave at three resident CTAs per SM
        if (chunks < 2) chunks = 2;
        if (chunks > kClMaxChunks) chunks = kClMaxChunks;
        while (chunks > 1 && rows / chunks < (long long)g.rows_per_pass * kClUnroll) --chunks;
    }
    int cr = (int)((rows + chunks - 1) / chunks);
    cr = (cr + g.rows_per_pass - 1) / g.rows_per_pass * g.rows_per_pass;
    g.chunk_rows = cr;
    g.chunks = (int)((rows + cr - 1) / cr);
    return HG_OK;
}


// Cluster plan for a CTA of `threads`: rows_per_pass = threads / lanes, q = N / rows_per_pass row passes per sample,
// split over the largest cluster size cs in {8, 4, 2} that leaves u = q / cs <= max_u rows per thread.
// Returns false when the instance has no such split, is too small to pay for a cluster launch, or
// HG_ADAIN_CL_NO_CLUSTER is set (A/B runs and tests of the chunked kernels) -> the chunked two-kernel path takes it.
static bool cl_cluster_plan(const ClGeom &g, int threads, int max_u, bool backward, ClGeom &cg_out, int &cs_out, int &u_out)
{
    if (option(kOptAdainClNoCluster)) return false;
    if (g.C > kClClusterMaxC || g.lanes > threads) return false;
    // Measured on B200 (profiles/r01g_microbench_pipeline.txt): a cluster launch has a ~6 us floor per wave, so the
    // forward only wins for the generator's 512 KB instances (24 us vs 32 us); the backward (two staged tensors ->
    // one 164 KB CTA per SM, 3.5 lock-step waves) does not beat the chunked kernels (62 us vs 59 us) and stays
    // opt-in (HG_ADAIN_CL_CLUSTER_BWD=1: tests, experiments).
    if ((long long)g.N * g.C * 2 < 256 * 1024) return false;
    if (backward) {
        if (!option(kOptAdainClClusterBwd)) return false;
    }
    const int rpp = threads / g.lanes;
    if (g.N % rpp) return false;
    const int q = g.N / rpp;
    for (int cs = 8; cs >= 2; cs >>= 1) {
        if (q % cs) continue;
        const int u = q / cs;
        if (u > max_u) continue;
        cg_out = g;
        cg_out.rows_per_pass = rpp;
        cg_out.chunk_rows = u * rpp;
        cg_out.chunks = cs;
        cs_out = cs;
        u_out = u;
        return true;
    }
    return false;
}

template <typename K, typename... Args>
static int cl_cluster_launch(const char *what, K kernel, int cs, int batch, int threads, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, batch);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HG_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    }
    return check_launch(what);
}

// register-resident one-CTA-per-sample kernels: only where a thread owns 9 .. 16 rows (D block 0: backward 16.2 -> 10.8 us);
// with <= 8 rows the fused kernels already need one batch of loads per pass (profiles/r02Q_adain_small.txt)
static bool cl_small_regs(const ClGeom &g)
{
    return option(kOptAdainClSmallRegs) != 0 && g.N <= kSmallRows * g.rows_per_pass && g.N > 8 * g.rows_per_pass;
}

extern "C" long long hg_adain_cl_workspace_bytes(int batch, int channels, int ndim, int size, int classes)
{
    ClGeom g;
    if (cl_geom("hg_adain_cl_workspace_bytes", batch, channels, ndim, size, classes, 0, g)) return -1;
    // per-chunk partials + the merged sums of the backward
    return g.chunks > 1 ? (long long)batch * (g.chunks + 1) * 2 * channels * (long long)sizeof(float) : 0;
}

extern "C" int hg_adain_cl_fwd(const void *x, const float *scale, const float *bias, void *y, float *save_mean,
                               float *save_rstd, void *workspace, long long workspace_bytes, int batch, int channels, int ndim,
                               int size, int classes, int sb_stride, float eps, float neg_slope, int biased_var, void *stream)
{
    HG_REQUIRE(x && y && save_mean && save_rstd, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: null pointer");
    HG_REQUIRE((scale == nullptr) == (bias == nullptr), HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: scale and bias must both be given or both be null");
    ClGeom g;
    int rc = cl_geom("hg_adain_cl_fwd", batch, channels, ndim, size, classes, biased_var, g);
    if (rc) return rc;
    HG_REQUIRE(!scale || sb_stride >= channels, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: sb_stride < channels");
    const __nv_bfloat16 *xp = static_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *yp = static_cast<__nv_bfloat16 *>(y);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ClGeom cgm;
        int cs = 0, u = 0;
        if (cl_cluster_plan(g, kClThreads, kClClusterFwdRows, false, cgm, cs, u)) {
            static bool attr_done = false;      // not a stream operation (graph-capture safe)
            if (!attr_done) {
                cudaFuncSetAttribute(adain_cl_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kClClusterFwdRows * kClThreads * 16);
                attr_done = true;
            }
            return cl_cluster_launch("hg_adain_cl_fwd(cluster)", adain_cl_cluster_fwd_kernel, cs, batch, kClThreads,
                                     (size_t)u * kClThreads * 16, st, xp, scale, bias, yp, save_mean, save_rstd, cgm, u, sb_stride,
                                     eps, neg_slope);
        }
    }
    dim3 grid(g.chunks, batch);
    if (g.chunks == 1 && cl_small_regs(g)) {
        adain_cl_small_fwd_kernel<<<batch, kClThreads, 0, st>>>(xp, scale, bias, yp, save_mean, save_rstd, g, sb_stride, eps, neg_slope);
        return check_launch("hg_adain_cl_fwd(small)");
    }
    if (g.chunks == 1) {
        adain_cl_apply_kernel<true><<<grid, kClThreads, 0, st>>>(xp, nullptr, scale, bias, yp, save_mean, save_rstd, g, sb_stride,
                                                                eps, neg_slope);
        return check_launch("hg_adain_cl_fwd");
    }
    HG_REQUIRE(workspace && workspace_bytes >= (long long)batch * (g.chunks + 1) * 2 * channels * (long long)sizeof(float),
               HG_ERR_INVALID_ARG, "hg_adain_cl_fwd: workspace smaller than hg_adain_cl_workspace_bytes()");
    float *part = static_cast<float *>(workspace);
    adain_cl_stats_kernel<<<grid, kClThreads, 0, st>>>(xp, part, g);
    rc = check_launch("hg_adain_cl_fwd(stats)");
    if (rc) return rc;
    adain_cl_apply_kernel<false><<<grid, kClThreads, 0, st>>>(xp, part, scale, bias, yp, save_mean, save_rstd, g, sb_stride, eps,
                                                             neg_slope);
    return check_launch("hg_adain_cl_fwd(apply)");
}

extern "C" int hg_adain_cl_bwd(const void *x, const void *dy, const float *scale, const float *bias, const float *save_mean,
                               const float *save_rstd, void *dx, float *dscale, float *dbias, void *workspace,
                               long long workspace_bytes, int batch, int channels, int ndim, int size, int classes, int sb_stride,
                               int dsb_stride, float neg_slope, int biased_var, void *stream)
{
    HG_REQUIRE(x && dy && save_mean && save_rstd && dx, HG_ERR_INVALID_ARG, "hg_adain_cl_bwd: null pointer");
    HG_REQUIRE((scale == nullptr) == (bias == nullptr) && (dscale == nullptr) == (dbias == nullptr), HG_ERR_INVALID_ARG,
               "hg_adain_cl_bwd: scale/bias and dscale/dbias come in pairs");
    ClGeom g;
    int rc = cl_geom("hg_adain_cl_bwd", batch, channels, ndim, size, classes, biased_var, g);
    if (rc) return rc;
    HG_REQUIRE((!scale || sb_stride >= channels) && (!dscale || dsb_stride >= channels), HG_ERR_INVALID_ARG,
               "hg_adain_cl_bwd: stride < channels");
    const __nv_bfloat16 *xp = static_cast<const __nv_bfloat16 *>(x), *gp = static_cast<const __nv_bfloat16 *>(dy);
    __nv_bfloat16 *dp = static_cast<__nv_bfloat16 *>(dx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ClGeom cgm;
        int cs = 0, u = 0;
        if (cl_cluster_plan(g, kClBwdClusterThreads, kClClusterBwdRows, true, cgm, cs, u)) {
            static bool attr_done = false;      // not a stream operation (graph-capture safe)
            if (!attr_done) {
                cudaFuncSetAttribute(adain_cl_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     2 * kClClusterBwdRows * kClBwdClusterThreads * 16);
                attr_done = true;
            }
            return cl_cluster_launch("hg_adain_cl_bwd(cluster)", adain_cl_cluster_bwd_kernel, cs, batch, kClBwdClusterThreads,
                                     (size_t)2 * u * kClBwdClusterThreads * 16, st, xp, gp, scale, bias, save_mean, save_rstd, dp,
                                     dscale, dbias, cgm, u, sb_stride, dsb_stride, neg_slope);
        }
    }
    dim3 grid(g.chunks, batch);
    if (g.chunks == 1 && cl_small_regs(g)) {
        adain_cl_small_bwd_kernel<<<batch, kClThreads, 0, st>>>(xp, gp, scale, bias, save_mean, save_rstd, dp, dscale, dbias, g, sb_stride,
                                                               dsb_stride, neg_slope);
        return check_launch("hg_adain_cl_bwd(small)");
    }
    if (g.chunks == 1) {
        adain_cl_bwd_apply_kernel<true><<<grid, kClThreads, 0, st>>>(xp, gp, nullptr, scale, bias, save_mean, save_rstd, dp, dscale,
                                                                    dbias, g, sb_stride, dsb_stride, neg_slope);
        return check_launch("hg_adain_cl_bwd");
    }
    HG_REQUIRE(workspace && workspace_bytes >= (long long)batch * (g.chunks + 1) * 2 * channels * (long long)sizeof(float),
               HG_ERR_INVALID_ARG, "hg_adain_cl_bwd: workspace smaller than hg_adain_cl_workspace_bytes()");
    float *part = static_cast<float *>(workspace);
    if (option(kOptAdainClRing) != 0) {                 // cp.async ring (see ring_stream2)
        static bool ring_attr = false;                  // not a stream operation (graph-capture safe)
        if (!ring_attr) {
            cudaFuncSetAttribute(adain_cl_bwd_sums_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingBytes);
            cudaFuncSetAttribute(adain_cl_bwd_apply_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingBytes);
            ring_attr = true;
        }
        adain_cl_bwd_sums_ring_kernel<<<grid, kClThreads, kRingBytes, st>>>(xp, gp, scale, bias, save_mean, save_rstd, part, g, sb_stride,
                                                                           neg_slope);
        rc = check_launch("hg_adain_cl_bwd(sums, ring)");
        if (rc) return rc;
        adain_cl_bwd_apply_ring_kernel<<<grid, kClThreads, kRingBytes, st>>>(xp, gp, part, scale, bias, save_mean, save_rstd, dp, dscale,
                                                                            dbias, g, sb_stride, dsb_stride, neg_slope);
        return check_launch("hg_adain_cl_bwd(apply, ring)");
    }
    adain_cl_bwd_sums_kernel<<<grid, kClThreads, 0, st>>>(xp, gp, scale, bias, save_mean, save_rstd, part, g, sb_stride, neg_slope);
    rc = check_launch("hg_adain_cl_bwd(sums)");
    if (rc) return rc;
    adain_cl_bwd_apply_kernel<false><<<grid, kClThreads, 0, st>>>(xp, gp, part, scale, bias, save_mean, save_rstd, dp, dscale,
                                                                 dbias, g, sb_stride, dsb_stride, neg_slope);
    return check_launch("hg_adain_cl_bwd(apply)");
}

// AdaIN (+ activation) of a transposed convolution's s2d output whose statistics partials were produced by the GEMM
// epilogue (hg_convt_fwd_stats): a tiny merge kernel + ONE streaming pass (1 read + 1 write of the activation).
extern "C" int hg_adain_cl_fwd_stats(const void *x, const float *stats, const float *scale, const float *bias, void *y,
                                     float *save_mean, float *save_rstd, int batch, int channels, int ndim, int size, int classes,
                                     int sb_stride, float eps, float neg_slope, int biased_var, void *stream)
{
    HG_REQUIRE(x && stats && y && save_mean && save_rstd, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd_stats: null pointer");
    HG_REQUIRE((scale == nullptr) == (bias == nullptr), HG_ERR_INVALID_ARG, "hg_adain_cl_fwd_stats: scale and bias must both be given or both be null");
    HG_REQUIRE(classes >= 1, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd_stats: classes must be 1 or 2^ndim");
    ClGeom g;
    int rc = cl_geom("hg_adain_cl_fwd_stats", batch, channels, ndim, size, classes, biased_var, g);
    if (rc) return rc;
    long long pos = 1;
    for (int i = 0; i < ndim; ++i) pos *= size;
    HG_REQUIRE(pos % 32 == 0, HG_ERR_UNSUPPORTED, "hg_adain_cl_fwd_stats: positions per sample must be a multiple of 32 (got %lld)", pos);
    HG_REQUIRE(!scale || sb_stride >= channels, HG_ERR_INVALID_ARG, "hg_adain_cl_fwd_stats: sb_stride < channels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    adain_cl_stats_finalize_kernel<<<dim3((channels + 31) / 32, batch), 256, 0, st>>>(stats, save_mean, save_rstd, channels, classes,
                                                                                    (int)(pos / 32), g.Nvar, eps);
    rc = check_launch("hg_adain_cl_fwd_stats(finalize)");
    if (rc) return rc;
    // streaming pass: enough chunks for ~3 CTAs per SM, at least kClUnroll rows per thread
    int chunks = (3 * sm_count() + batch - 1) / batch;
    while (chunks > 1 && g.N / chunks < g.rows_per_pass * kClUnroll) --chunks;
    int cr = (g.N + chunks - 1) / chunks;
    cr = (cr + g.rows_per_pass - 1) / g.rows_per_pass * g.rows_per_pass;
    g.chunk_rows = cr;
    g.chunks = (g.N + cr - 1) / cr;
    adain_cl_apply_stats_kernel<<<dim3(g.chunks, batch), kClThreads, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), scale, bias, save_mean,
                                                                             save_rstd, static_cast<__nv_bfloat16 *>(y), g, sb_stride,
                                                                             neg_slope);
    return check_launch("hg_adain_cl_fwd_stats(apply)");
}
