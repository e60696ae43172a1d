// The two small dense ops at the ends of the generator:
//   * ZMapping: relu(Linear(z)) for the five style mappings (reference hologan_generator.py:7-18), fp32 SIMT
//     (0.01 % of the FLOPs; kept in fp32 because every activation of a block is scaled by its output);
//   * final_layer + tanh: Conv2d(C -> 3, k3, p1) + tanh (reference :69-75,141-142).  N = 3 output channels
//     makes it bandwidth-bound (reads the (B,64,64,C) activation once, 9x reuse from L1), so it is a direct
//     convolution on CUDA cores fused with the tanh, not a tensor-core GEMM.
#include "hg_common.cuh"

namespace hg {

// -------------------------------------------------------------------------------------------------
// ZMapping: out[b, n] = relu(sum_k z[b,k] * W[n,k] + bias[n])
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linear_relu_fwd_kernel(const float *__restrict__ z, const float *__restrict__ w,
                                                              const float *__restrict__ bias, float *__restrict__ out, int B,
                                                              int K, int N)
{
    extern __shared__ float wrow[];                    // W[n, :]
    const int n = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x) wrow[k] = w[(size_t)n * K + k];
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float *zr = z + (size_t)b * K;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(zr[k], wrow[k], acc);
        acc += bias[n];
        out[(size_t)b * N + n] = acc > 0.f ? acc : 0.f;
    }
}

// dW[n,k] = sum_b g[b,n] z[b,k], dbias[n] = sum_b g[b,n], with g = dout * (out > 0).  One CTA per n.
__global__ void __launch_bounds__(128) linear_relu_bwd_w_kernel(const float *__restrict__ z, const float *__restrict__ out,
                                                                const float *__restrict__ dout, float *__restrict__ dw,
                                                                float *__restrict__ dbias, int B, int K, int N)
{
    extern __shared__ float g[];                       // g[b] for this n
    const int n = blockIdx.x;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const size_t i = (size_t)b * N + n;
        g[b] = out[i] > 0.f ? dout[i] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc = fmaf(g[b], z[(size_t)b * K + k], acc);
        dw[(size_t)n * K + k] = acc;
    }
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += g[b];
        dbias[n] = acc;
    }
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

// -------------------------------------------------------------------------------------------------
// final conv (k3, p1, Cout <= 4) + tanh.  x: (B,S,S,C) bf16 NHWC, w: torch (Cout, C, 3, 3) fp32,
// out: (B,Cout,S,S) fp32 NCHW.
// -------------------------------------------------------------------------------------------------
constexpr int kFinalMaxCout = 4;

__device__ __forceinline__ void unpack8_f(const uint4 &u, float *f)
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// weights in smem as ws[tap][ci][co4]
__global__ void __launch_bounds__(256) final_conv_tanh_fwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                  const float *__restrict__ w, const float *__restrict__ bias,
                                                                  float *__restrict__ out, int C, int Cout, int S)
{
    extern __shared__ float ws[];                      // [9][C][4]
    for (int i = threadIdx.x; i < 9 * C * 4; i += blockDim.x) {
        const int co = i & 3, ci = (i >> 2) % C, t = i / (4 * C);
        ws[i] = co < Cout ? w[((size_t)co * C + ci) * 9 + t] : 0.f;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= S * S) return;
    const int py = p / S, px = p - py * S;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16 *xb = x + (size_t)b * S * S * C;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy < 0 || yy >= S || xx < 0 || xx >= S) continue;
        const uint4 *src = reinterpret_cast<const uint4 *>(xb + ((size_t)yy * S + xx) * C);
        const float4 *wt = reinterpret_cast<const float4 *>(ws + t * C * 4);
        for (int c8 = 0; c8 < C / 8; ++c8) {
            float f[8];
            unpack8_f(__ldg(src + c8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 wv = wt[c8 * 8 + j];
                acc[0] = fmaf(f[j], wv.x, acc[0]);
                acc[1] = fmaf(f[j], wv.y, acc[1]);
                acc[2] = fmaf(f[j], wv.z, acc[2]);
                acc[3] = fmaf(f[j], wv.w, acc[3]);
            }
        }
    }
    for (int co = 0; co < Cout; ++co) out[((size_t)b * Cout + co) * S * S + p] = tanhf(acc[co] + bias[co]);
}

// dx[b,y,x,ci] = sum_(co,t) dpre[b,co,y-ty+1,x-tx+1] * w[co,ci,t],  dpre = dout * (1 - out^2)
__global__ void __launch_bounds__(256) final_conv_tanh_bwd_x_kernel(const float *__restrict__ w, const float *__restrict__ out,
                                                                    const float *__restrict__ dout,
                                                                    __nv_bfloat16 *__restrict__ dx, int C, int Cout, int S)
{
    extern __shared__ float ws[];                      // [9][4][C]  (ci fastest)
    for (int i = threadIdx.x; i < 9 * 4 * C; i += blockDim.x) {
        const int ci = i % C, co = (i / C) & 3, t = i / (4 * C);
        ws[i] = co < Cout ? w[((size_t)co * C + ci) * 9 + t] : 0.f;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int lanes = C / 8;                           // threads per pixel, 8 channels each
    const int slot = threadIdx.x / lanes, sub = threadIdx.x % lanes;
    const int p = blockIdx.x * (blockDim.x / lanes) + slot;
    if (p >= S * S) return;
    const int py = p / S, px = p - py * S;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = py - (t / 3 - 1), xx = px - (t % 3 - 1);
        if (yy < 0 || yy >= S || xx < 0 || xx >= S) continue;
        for (int co = 0; co < Cout; ++co) {
            const size_t i = ((size_t)b * Cout + co) * S * S + (size_t)yy * S + xx;
            const float o = out[i];
            const float g = dout[i] * (1.f - o * o);
            const float *wv = ws + (t * 4 + co) * C + sub * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, wv[j], acc[j]);
        }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(dx + ((size_t)b * S * S + p) * C + sub * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// Partial weight / bias gradients of one (sample, 16-row band): part[cta][co][ci][t] and part_b[cta][co].
// Thread t owns weight entries e = t, t+256, ... of the Cout*C*9 (layout [co][t9][ci], ci fastest for
// conflict-free smem reads) and loops over the band's pixels.
__global__ void __launch_bounds__(256) final_conv_tanh_bwd_w_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                    const float *__restrict__ out,
                                                                    const float *__restrict__ dout, float *__restrict__ part,
                                                                    int C, int Cout, int S, int band)
{
    extern __shared__ unsigned char sm_raw[];
    // xs: (band + 2) rows x (S + 2) cols x C bf16 (zero halo) ; gs: Cout x band x S fp32
    __nv_bfloat16 *xs = reinterpret_cast<__nv_bfloat16 *>(sm_raw);
    const int rows = band + 2, cols = S + 2;
    float *gs = reinterpret_cast<float *>(sm_raw + (((size_t)rows * cols * C * 2 + 15) & ~size_t(15)));
    const int b = blockIdx.y, y0 = blockIdx.x * band;
    const __nv_bfloat16 *xb = x + (size_t)b * S * S * C;
    const int vec_per_px = C / 8;
    for (int i = threadIdx.x; i < rows * cols * vec_per_px; i += blockDim.x) {
        const int v = i % vec_per_px, px = (i / vec_per_px) % cols, py = i / (vec_per_px * cols);
        const int yy = y0 + py - 1, xx = px - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (yy >= 0 && yy < S && xx >= 0 && xx < S) val = __ldg(reinterpret_cast<const uint4 *>(xb + ((size_t)yy * S + xx) * C) + v);
        reinterpret_cast<uint4 *>(xs)[i] = val;
    }
    for (int i = threadIdx.x; i < Cout * band * S; i += blockDim.x) {
        const int px = i % S, py = (i / S) % band, co = i / (S * band);
        float g = 0.f;
        if (y0 + py < S) {
            const size_t gi = ((size_t)b * Cout + co) * S * S + (size_t)(y0 + py) * S + px;
            const float o = out[gi];
            g = dout[gi] * (1.f - o * o);
        }
        gs[i] = g;
    }
    __syncthreads();
    const int n_entries = Cout * 9 * C;
    float *pout = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (n_entries + kFinalMaxCout);
    for (int e = threadIdx.x; e < n_entries; e += blockDim.x) {
        const int ci = e % C, t = (e / C) % 9, co = e / (9 * C);
        const int ty = t / 3, tx = t % 3;
        float acc = 0.f;
        for (int py = 0; py < band; ++py) {
            const __nv_bfloat16 *xr = xs + ((size_t)(py + ty) * cols + tx) * C + ci;
            const float *gr = gs + (co * band + py) * S;
            for (int px = 0; px < S; ++px) acc = fmaf(gr[px], __bfloat162float(xr[(size_t)px * C]), acc);
        }
        pout[e] = acc;
    }
    if (threadIdx.x < Cout) {
        float acc = 0.f;
        for (int i = 0; i < band * S; ++i) acc += gs[threadIdx.x * band * S + i];
        pout[n_entries + threadIdx.x] = acc;
    }
}

// dw[co][ci][t] = sum_cta part[cta][co][t][ci]; dbias[co] = sum_cta part_b
__global__ void final_conv_reduce_kernel(const float *__restrict__ part, float *__restrict__ dw, float *__restrict__ dbias,
                                         int C, int Cout, int n_cta)
{
    const int n_entries = Cout * 9 * C;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries + Cout) return;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += part[(size_t)k * (n_entries + kFinalMaxCout) + e];
    if (e < n_entries) {
        const int ci = e % C, t = (e / C) % 9, co = e / (9 * C);
        dw[((size_t)co * C + ci) * 9 + t] = acc;
    } else {
        dbias[e - n_entries] = acc;
    }
}

}  // namespace hg

using namespace hg;

extern "C" int hg_linear_relu_fwd(const float *z, const float *w, const float *bias, float *out, int batch, int k, int n,
                                  void *stream)
{
    HG_REQUIRE(z && w && bias && out, HG_ERR_INVALID_ARG, "hg_linear_relu_fwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && n > 0 && k <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_fwd: bad dims");
    linear_relu_fwd_kernel<<<n, 128, k * sizeof(float), static_cast<cudaStream_t>(stream)>>>(z, w, bias, out, batch, k, n);
    return check_launch("hg_linear_relu_fwd");
}

extern "C" int hg_linear_relu_bwd(const float *z, const float *w, const float *out, const float *dout, float *dw,
                                  float *dbias, float *dz, int batch, int k, int n, int accumulate_dz, void *stream)
{
    HG_REQUIRE(z && w && out && dout && dw && dbias, HG_ERR_INVALID_ARG, "hg_linear_relu_bwd: null pointer");
    HG_REQUIRE(batch > 0 && k > 0 && n > 0 && batch <= 8192 && n <= 8192, HG_ERR_INVALID_ARG, "hg_linear_relu_bwd: bad dims");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    linear_relu_bwd_w_kernel<<<n, 128, batch * sizeof(float), st>>>(z, out, dout, dw, dbias, batch, k, n);
    int rc = check_launch("hg_linear_relu_bwd(w)");
    if (rc || !dz) return rc;
    linear_relu_bwd_z_kernel<<<batch, 128, n * sizeof(float), st>>>(w, out, dout, dz, k, n, accumulate_dz);
    return check_launch("hg_linear_relu_bwd(z)");
}

static int final_check(const char *who, int batch, int cin, int cout, int size)
{
    HG_REQUIRE(batch > 0 && batch <= 65535 && size > 0, HG_ERR_INVALID_ARG, "%s: bad batch / size", who);
    HG_REQUIRE(cin % 8 == 0 && cin >= 8 && cin <= 256 && (256 % (cin / 8)) == 0, HG_ERR_UNSUPPORTED,
               "%s: Cin must be 8 * 2^k <= 256 (got %d)", who, cin);
    HG_REQUIRE(cout >= 1 && cout <= kFinalMaxCout, HG_ERR_UNSUPPORTED, "%s: Cout must be <= %d (got %d)", who, kFinalMaxCout, cout);
    return HG_OK;
}

extern "C" int hg_final_conv_tanh_fwd(const void *x, const float *w, const float *bias, float *out, int batch, int cin,
                                      int cout, int size, void *stream)
{
    HG_REQUIRE(x && w && bias && out, HG_ERR_INVALID_ARG, "hg_final_conv_tanh_fwd: null pointer");
    int rc = final_check("hg_final_conv_tanh_fwd", batch, cin, cout, size);
    if (rc) return rc;
    dim3 grid((size * size + 255) / 256, batch);
    final_conv_tanh_fwd_kernel<<<grid, 256, 9 * cin * 4 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), w, bias, out, cin, cout, size);
    return check_launch("hg_final_conv_tanh_fwd");
}

static int final_band(int size) { return size >= 16 ? 8 : size; }

extern "C" long long hg_final_conv_tanh_bwd_workspace_bytes(int batch, int cin, int cout, int size)
{
    if (batch <= 0 || cin <= 0 || cout <= 0 || size <= 0) return -1;
    const int band = final_band(size);
    const long long ctas = (long long)batch * ((size + band - 1) / band);
    return ctas * (cout * 9LL * cin + kFinalMaxCout) * (long long)sizeof(float);
}

extern "C" int hg_final_conv_tanh_bwd(const void *x, const float *w, const float *out, const float *dout, void *dx, float *dw,
                                      float *dbias, void *workspace, long long workspace_bytes, int batch, int cin, int cout,
                                      int size, void *stream)
{
    HG_REQUIRE(x && w && out && dout && dw && dbias && workspace, HG_ERR_INVALID_ARG, "hg_final_conv_tanh_bwd: null pointer");
    int rc = final_check("hg_final_conv_tanh_bwd", batch, cin, cout, size);
    if (rc) return rc;
    HG_REQUIRE(workspace_bytes >= hg_final_conv_tanh_bwd_workspace_bytes(batch, cin, cout, size), HG_ERR_INVALID_ARG,
               "hg_final_conv_tanh_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dx) {
        const int lanes = cin / 8, px_per_cta = 256 / lanes;
        dim3 grid((size * size + px_per_cta - 1) / px_per_cta, batch);
        final_conv_tanh_bwd_x_kernel<<<grid, 256, 9 * 4 * cin * sizeof(float), st>>>(w, out, dout, static_cast<__nv_bfloat16 *>(dx),
                                                                                  cin, cout, size);
        rc = check_launch("hg_final_conv_tanh_bwd(x)");
        if (rc) return rc;
    }
    const int band = final_band(size);
    const int bands = (size + band - 1) / band;
    const size_t xs_bytes = (((size_t)(band + 2) * (size + 2) * cin * 2) + 15) & ~size_t(15);
    const size_t smem = xs_bytes + (size_t)cout * band * size * sizeof(float);
    HG_REQUIRE(smem <= 200 * 1024, HG_ERR_UNSUPPORTED, "hg_final_conv_tanh_bwd: tile does not fit in shared memory (%zu B)", smem);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(final_conv_tanh_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    dim3 wgrid(bands, batch);
    final_conv_tanh_bwd_w_kernel<<<wgrid, 256, smem, st>>>(static_cast<const __nv_bfloat16 *>(x), out, dout,
                                                         static_cast<float *>(workspace), cin, cout, size, band);
    rc = check_launch("hg_final_conv_tanh_bwd(w)");
    if (rc) return rc;
    const int n = cout * 9 * cin + cout;
    final_conv_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(static_cast<const float *>(workspace), dw, dbias, cin, cout,
                                                            bands * batch);
    return check_launch("hg_final_conv_tanh_bwd(reduce)");
}
