// Backward of the activation + bias that hg_convt_fwd fuses into its epilogue (the 1x1 projection's
// `relu(convTranspose2d1(x))`, reference core/models/hologan_generator.py:135-136):
//     dpre[m, n] = y[m, n] > 0 ? dy[m, n] : slope * dy[m, n]          (bf16, feeds dgrad / wgrad)
//     dbias[n]   = sum_m dpre[m, n]                                     (fp32)
// in ONE pass over y and dy (HBM-bound: algorithmic bytes = 3 * M * N * 2), instead of the compare / scale /
// select / cast / reduce kernels of an eager implementation.  Column sums are deterministic: fixed row
// order inside a CTA, per-CTA partials summed in CTA order by a second tiny kernel.
#include "hg_common.cuh"

namespace hg {

constexpr int kActThreads = 256;
constexpr int kActUnroll = 4;

__device__ __forceinline__ void unpack8_e(const uint4 &u, float *f)
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// thread -> one 8-column vector; the CTA's threads tile (rows_per_pass x cols/8); CTAs stride over row bands
__global__ void __launch_bounds__(kActThreads) act_bwd_bias_kernel(const __nv_bfloat16 *__restrict__ y,
                                                                   const __nv_bfloat16 *__restrict__ dy,
                                                                   __nv_bfloat16 *__restrict__ dpre, float *__restrict__ part,
                                                                   long long rows, int cols, float slope)
{
    const int vec_per_row = cols >> 3;
    // a thread keeps ONE column vector for its whole life so the 8 column sums stay in registers
    const int col_slots = vec_per_row < kActThreads ? vec_per_row : kActThreads;
    const int row_slots = kActThreads / col_slots;
    const int cs = threadIdx.x % col_slots, rs = threadIdx.x / col_slots;
    __shared__ float red[kActThreads * 8];              // [row_slot][col_slot * 8]
    float sum[8];
    for (int cv = cs; cv < vec_per_row; cv += col_slots) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = 0.f;
        if (rs < row_slots) {
            const long long stride = (long long)gridDim.x * row_slots;
            for (long long r0 = (long long)blockIdx.x * row_slots + rs; r0 < rows; r0 += kActUnroll * stride) {
                uint4 yv[kActUnroll], gv[kActUnroll];
#pragma unroll
                for (int u = 0; u < kActUnroll; ++u) {       // all loads of the group in flight before any use
                    const long long r = r0 + u * stride;
                    if (r < rows) {
                        const size_t off = (size_t)r * cols + (size_t)cv * 8;
                        yv[u] = ld_stream_16(y + off);
                        gv[u] = ld_stream_16(dy + off);
                    }
                }
#pragma unroll
                for (int u = 0; u < kActUnroll; ++u) {
                    const long long r = r0 + u * stride;
                    if (r < rows) {
                        float yf[8], gf[8];
                        unpack8_e(yv[u], yf);
                        unpack8_e(gv[u], gf);
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 8; ++j) gf[j] = yf[j] > 0.f ? gf[j] : gf[j] * slope;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(gf[2 * j], gf[2 * j + 1]);
                            pk[j] = *reinterpret_cast<uint32_t *>(&h);
                            // sum what dgrad / wgrad will consume (the rounded values), like the eager reference
                            sum[2 * j] += __uint_as_float(pk[j] << 16);
                            sum[2 * j + 1] += __uint_as_float(pk[j] & 0xffff0000u);
                        }
                        st_stream_16(dpre + (size_t)r * cols + (size_t)cv * 8, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    }
                }
            }
        }
        if (part) {                                     // part[cta][cols]: row slots summed in slot order
            __syncthreads();
            if (rs < row_slots) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red[(rs * col_slots + cs) * 8 + j] = sum[j];
            }
            __syncthreads();
            if (rs == 0) {
                for (int k = 1; k < row_slots; ++k) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sum[j] += red[(k * col_slots + cs) * 8 + j];
                }
                float *dst = part + (size_t)blockIdx.x * cols + (size_t)cv * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = sum[j];
            }
        }
    }
}

// 32 columns x 8 partial groups per CTA; fixed order inside a group, groups summed in group order
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float *__restrict__ part, float *__restrict__ out, int n_part,
                                                            int cols)
{
    __shared__ float red[8][32];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float acc = 0.f;
    if (c < cols) {
        const int per = (n_part + 7) / 8;
        const int lo = grp * per, hi = min(n_part, lo + per);
#pragma unroll 4
        for (int i = lo; i < hi; ++i) acc += part[(size_t)i * cols + c];
    }
    red[grp][cl] = acc;
    __syncthreads();
    if (grp == 0 && c < cols) {
        float s = red[0][cl];
#pragma unroll
        for (int g = 1; g < 8; ++g) s += red[g][cl];
        out[c] = s;
    }
}

static void act_plan(long long rows, int cols, int &grid, int &row_slots)
{
    const int vec_per_row = cols >> 3;
    const int col_slots = vec_per_row < kActThreads ? vec_per_row : kActThreads;
    row_slots = kActThreads / col_slots;
    long long g = (rows + row_slots - 1) / row_slots;
    const long long cap = 2LL * sm_count();
    grid = (int)(g < cap ? g : cap);
    if (grid < 1) grid = 1;
}

}  // namespace hg

using namespace hg;

extern "C" long long hg_act_bwd_bias_workspace_bytes(long long rows, int cols)
{
    if (rows <= 0 || cols <= 0 || cols % 8) return -1;
    int grid, row_slots;
    act_plan(rows, cols, grid, row_slots);
    return (long long)grid * cols * (long long)sizeof(float);
}

extern "C" int hg_act_bwd_bias(const void *y, const void *dy, void *dpre, float *dbias, void *workspace,
                               long long workspace_bytes, long long rows, int cols, float neg_slope, void *stream)
{
    HG_REQUIRE(y && dy && dpre, HG_ERR_INVALID_ARG, "hg_act_bwd_bias: null pointer");
    HG_REQUIRE(rows > 0 && cols > 0, HG_ERR_INVALID_ARG, "hg_act_bwd_bias: dims must be positive");
    HG_REQUIRE(cols % 8 == 0 && (cols <= 8 * kActThreads || cols % (8 * kActThreads) == 0), HG_ERR_UNSUPPORTED,
               "hg_act_bwd_bias: cols must be a multiple of 8, and of 2048 beyond 2048 (got %d)", cols);
    int grid, row_slots;
    act_plan(rows, cols, grid, row_slots);
    if (dbias)
        HG_REQUIRE(workspace && workspace_bytes >= hg_act_bwd_bias_workspace_bytes(rows, cols), HG_ERR_INVALID_ARG,
                   "hg_act_bwd_bias: workspace smaller than hg_act_bwd_bias_workspace_bytes()");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    act_bwd_bias_kernel<<<grid, kActThreads, 0, st>>>(static_cast<const __nv_bfloat16 *>(y),
                                                     static_cast<const __nv_bfloat16 *>(dy),
                                                     static_cast<__nv_bfloat16 *>(dpre),
                                                     dbias ? static_cast<float *>(workspace) : nullptr, rows, cols, neg_slope);
    int rc = check_launch("hg_act_bwd_bias");
    if (rc || !dbias) return rc;
    colsum_reduce_kernel<<<(cols + 31) / 32, 256, 0, st>>>(static_cast<const float *>(workspace), dbias, grid, cols);
    return check_launch("hg_act_bwd_bias(reduce)");
}
