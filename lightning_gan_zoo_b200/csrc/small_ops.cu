// ZMapping: relu(Linear(z)) for the five style mappings (reference hologan_generator.py:7-18), fp32 SIMT
// (0.01 % of the FLOPs; kept in fp32 because every activation of a block is scaled by its output).
#include "hg_common.cuh"

namespace hg {

// -------------------------------------------------------------------------------------------------
// ZMapping: out[b, n] = relu(sum_k z[b,k] * W[n,k] + bias[n])
// -------------------------------------------------------------------------------------------------
// Tiles: 16 output features x 64 samples per CTA, K in chunks of 128 through shared memory (both operands are
// loaded with coalesced rows; the dot products read z as a broadcast and W conflict-free).
constexpr int kLinTn = 16, kLinTb = 64, kLinTk = 128;

__device__ __forceinline__ void linear_relu_fwd_body(const float *__restrict__ z, const float *__restrict__ w,
                                                     const float *__restrict__ bias, float *__restrict__ out, int B, int K,
                                                     int N, int bx, int by)
{
    __shared__ float zs[kLinTb][kLinTk];
    __shared__ float ws[kLinTn][kLinTk + 1];
    const int n0 = bx * kLinTn, b0 = by * kLinTb;
    const int nl = threadIdx.x % kLinTn, bg = threadIdx.x / kLinTn;        // 16 sample groups x 4 samples
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += kLinTk) {
        __syncthreads();
        for (int i = threadIdx.x; i < kLinTb * kLinTk; i += blockDim.x) {
            const int r = i / kLinTk, c = i % kLinTk;
            zs[r][c] = (b0 + r < B && k0 + c < K) ? z[(size_t)(b0 + r) * K + k0 + c] : 0.f;
        }
        for (int i = threadIdx.x; i < kLinTn * kLinTk; i += blockDim.x) {
            const int r = i / kLinTk, c = i % kLinTk;
            ws[r][c] = (n0 + r < N && k0 + c < K) ? w[(size_t)(n0 + r) * K + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kLinTk; ++k) {
            const float wv = ws[nl][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(zs[bg * 4 + j][k], wv, acc[j]);
        }
    }
    if (n0 + nl < N) {
        const float bv = bias[n0 + nl];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + bg * 4 + j;
            if (b < B) {
                const float v = acc[j] + bv;
                out[(size_t)b * N + n0 + nl] = v > 0.f ? v : 0.f;
            }
        }
    }
}

__global__ void __launch_bounds__(256) linear_relu_fwd_kernel(const float *__restrict__ z, const float *__restrict__ w,
                                                              const float *__restrict__ bias, float *__restrict__ out, int B,
                                                              int K, int N)
{
    linear_relu_fwd_body(z, w, bias, out, B, K, N, blockIdx.x, blockIdx.y);
}

// Grouped variant: the generator's five ZMappings read the same z (reference hologan_generator.py:34,54) -- one
// launch walks the feature tiles of all of them (tile_start = prefix sums of ceil(N_l / 16)).
constexpr int kLinMaxGroup = HG_LINEAR_GROUP_MAX;
struct LinGroup {
    const float *w[kLinMaxGroup];
    const float *bias[kLinMaxGroup];
    float *out[kLinMaxGroup];           // forward: outputs; backward: the saved outputs
    const float *dout[kLinMaxGroup];
    float *dw[kLinMaxGroup];
    float *dbias[kLinMaxGroup];
    int n[kLinMaxGroup];
    int tile_start[kLinMaxGroup + 1];
    int layers;
};
__device__ __forceinline__ int lin_group_layer(const LinGroup &g, int tile)
{
    int l = 0;
    while (l + 1 < g.layers && tile >= g.tile_start[l + 1]) ++l;
    return l;
}

__global__ void __launch_bounds__(256) linear_relu_group_fwd_kernel(const __grid_constant__ LinGroup g,
                                                                    const float *__restrict__ z, int B, int K)
{
    const int l = lin_group_layer(g, blockIdx.x);
    linear_relu_fwd_body(z, g.w[l], g.bias[l], g.out[l], B, K, g.n[l], blockIdx.x - g.tile_start[l], blockIdx.y);
}

// dW[n,k] = sum_b g[b,n] z[b,k], dbias[n] = sum_b g[b,n], with g = dout * (out > 0).
// grid = (N/16, K/128); samples in chunks of 64 through shared memory; fixed summation order.
__device__ __forceinline__ void linear_relu_bwd_w_body(const float *__restrict__ z, const float *__restrict__ out,
                                                       const float *__restrict__ dout, float *__restrict__ dw,
                                                       float *__restrict__ dbias, int B, int K, int N, int bx, int by)
{
    __shared__ float zs[kLinTb][kLinTk];
    __shared__ float gs[kLinTb][kLinTn];
    const int n0 = bx * kLinTn, k0 = by * kLinTk;
    const int kl = threadIdx.x % kLinTk, nh = threadIdx.x / kLinTk;        // 2 feature groups x 8 features
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    float bsum = 0.f;
    for (int b0 = 0; b0 < B; b0 += kLinTb) {
        __syncthreads();
        for (int i = threadIdx.x; i < kLinTb * kLinTk; i += blockDim.x) {
            const int r = i / kLinTk, c = i % kLinTk;
            zs[r][c] = (b0 + r < B && k0 + c < K) ? z[(size_t)(b0 + r) * K + k0 + c] : 0.f;
        }
        for (int i = threadIdx.x; i < kLinTb * kLinTn; i += blockDim.x) {
            const int r = i / kLinTn, c = i % kLinTn;
            float g = 0.f;
            if (b0 + r < B && n0 + c < N) {
                const size_t k = (size_t)(b0 + r) * N + n0 + c;
                g = out[k] > 0.f ? dout[k] : 0.f;
            }
            gs[r][c] = g;
        }
        __syncthreads();
#pragma unroll 4
        for (int b = 0; b < kLinTb; ++b) {
            const float zv = zs[b][kl];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(gs[b][nh * 8 + j], zv, acc[j]);
        }
        if (by == 0 && threadIdx.x < kLinTn)
            for (int b = 0; b < kLinTb; ++b) bsum += gs[b][threadIdx.x];
    }
    if (k0 + kl < K) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (n0 + nh * 8 + j < N) dw[(size_t)(n0 + nh * 8 + j) * K + k0 + kl] = acc[j];
    }
    if (by == 0 && threadIdx.x < kLinTn && n0 + threadIdx.x < N) dbias[n0 + threadIdx.x] = bsum;
}

__global__ void __launch_bounds__(256) linear_relu_bwd_w_kernel(const float *__restrict__ z, const float *__restrict__ out,
                                                                const float *__restrict__ dout, float *__restrict__ dw,
                                                                float *__restrict__ dbias, int B, int K, int N)
{
    linear_relu_bwd_w_body(z, out, dout, dw, dbias, B, K, N, blockIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256) linear_relu_group_bwd_w_kernel(const __grid_constant__ LinGroup g,
                                                                      const float *__restrict__ z, int B, int K)
{
    const int l = lin_group_layer(g, blockIdx.x);
    linear_relu_bwd_w_body(z, g.out[l], g.dout[l], g.dw[l], g.dbias[l], B, K, g.n[l], blockIdx.x - g.tile_start[l],
                           blockIdx.y);
}

// dz[b,k] (+)= sum_n g[b,n] W[n,k].  One CTA per sample, threads over k.
__global__ void __launch_bounds__(128) linear_relu_bwd_z_kernel(const float *__restrict__ w, const float *__restrict__ out,
                                                                const float *__restrict__ dout, float *__restrict__ dz, int K,
                                                                int N, int accumulate)
{
    extern __shared__ float g[];                       // g[n] for this sample
    const int b = blockIdx.x;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const size_t i = (size_t)b * N + n;
        g[n] = out[i] > 0.f ? dout[i] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(g[n], w[(size_t)n * K + k], acc);
        float *d = dz + (size_t)b * K + k;
        *d = accumulate ? *d + acc : acc;
    }
}

}  // namespace hg

using namespace hg;

extern "C" int hg_linear_relu_fwd(const float *z, const float *w, const float *bias, float *out, int batch, int k, int n,
                                  void *stream)
{
    HG_REQUIRE(z && w && bias && out, HG_ERR_INVALID_ARG, "hg_linear_relu_fwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && n > 0 && k <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_fwd: bad dims");
    dim3 grid((n + kLinTn - 1) / kLinTn, (batch + kLinTb - 1) / kLinTb);
    linear_relu_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(z, w, bias, out, batch, k, n);
    return check_launch("hg_linear_relu_fwd");
}

extern "C" int hg_linear_relu_bwd(const float *z, const float *w, const float *out, const float *dout, float *dw,
                                  float *dbias, float *dz, int batch, int k, int n, int accumulate_dz, void *stream)
{
    HG_REQUIRE(z && w && out && dout && dw && dbias, HG_ERR_INVALID_ARG, "hg_linear_relu_bwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && n > 0 && batch <= 8192 && n <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_bwd: bad dims");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 wgrid((n + kLinTn - 1) / kLinTn, (k + kLinTk - 1) / kLinTk);
    linear_relu_bwd_w_kernel<<<wgrid, 256, 0, st>>>(z, out, dout, dw, dbias, batch, k, n);
    int rc = check_launch("hg_linear_relu_bwd(w)");
    if (rc || !dz) return rc;
    linear_relu_bwd_z_kernel<<<batch, 128, n * sizeof(float), st>>>(w, out, dout, dz, k, n, accumulate_dz);
    return check_launch("hg_linear_relu_bwd(z)");
}

static int lin_group_fill(LinGroup &g, int layers, const int *n)
{
    g.layers = layers;
    int tiles = 0;
    for (int l = 0; l < layers; ++l) {
        g.n[l] = n[l];
        g.tile_start[l] = tiles;
        tiles += (n[l] + kLinTn - 1) / kLinTn;
    }
    g.tile_start[layers] = tiles;
    return tiles;
}

extern "C" int hg_linear_relu_group_fwd(int layers, const float *z, const float *const *w, const float *const *bias,
                                        float *const *out, const int *n, int batch, int k, void *stream)
{
    HG_REQUIRE(layers >= 1 && layers <= kLinMaxGroup, HG_ERR_INVALID_ARG, "hg_linear_relu_group_fwd: 1..%d layers", kLinMaxGroup);
    HG_REQUIRE(z && w && bias && out && n, HG_ERR_INVALID_ARG, "hg_linear_relu_group_fwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && k <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_group_fwd: bad dims");
    LinGroup g{};
    for (int l = 0; l < layers; ++l) {
        HG_REQUIRE(w[l] && bias[l] && out[l] && n[l] > 0, HG_ERR_INVALID_ARG, "hg_linear_relu_group_fwd: bad layer %d", l);
        g.w[l] = w[l]; g.bias[l] = bias[l]; g.out[l] = out[l];
    }
    const int tiles = lin_group_fill(g, layers, n);
    dim3 grid(tiles, (batch + kLinTb - 1) / kLinTb);
    linear_relu_group_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, z, batch, k);
    return check_launch("hg_linear_relu_group_fwd");
}

extern "C" int hg_linear_relu_group_bwd(int layers, const float *z, const float *const *out, const float *const *dout,
                                        float *const *dw, float *const *dbias, const int *n, int batch, int k, void *stream)
{
    HG_REQUIRE(layers >= 1 && layers <= kLinMaxGroup, HG_ERR_INVALID_ARG, "hg_linear_relu_group_bwd: 1..%d layers", kLinMaxGroup);
    HG_REQUIRE(z && out && dout && dw && dbias && n, HG_ERR_INVALID_ARG, "hg_linear_relu_group_bwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && batch <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_group_bwd: bad dims");
    LinGroup g{};
    for (int l = 0; l < layers; ++l) {
        HG_REQUIRE(out[l] && dout[l] && dw[l] && dbias[l] && n[l] > 0 && n[l] <= 8192, HG_ERR_INVALID_ARG,
                   "hg_linear_relu_group_bwd: bad layer %d", l);
        g.out[l] = const_cast<float *>(out[l]); g.dout[l] = dout[l]; g.dw[l] = dw[l]; g.dbias[l] = dbias[l];
    }
    const int tiles = lin_group_fill(g, layers, n);
    dim3 grid(tiles, (k + kLinTk - 1) / kLinTk);
    linear_relu_group_bwd_w_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, z, batch, k);
    return check_launch("hg_linear_relu_group_bwd");
}
