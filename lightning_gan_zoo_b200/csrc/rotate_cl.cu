// Rotate-resample on the channels-last layouts of the bf16 pipeline.
//   forward : vol NDHWC (B,S,S,S,C)  ->  out NDHWC or PROJ [b, z, x, y, c]
//             (PROJ is the A operand of the 1x1 projection GEMM: the depth-into-channels fold of
//              reference hologan_generator.py:130-133 happens in the store addressing, zero passes)
//   backward: grad_out (same layout as out) -> grad_vol NDHWC, formulated as a GATHER:
//             every source voxel s sums w(o, s) * g[o] over the outputs o whose 2x2x2 footprint
//             contains s, read from the per-sample adjoint tables of rotate_il.cu (built once per call
//             from the views, fixed entry order) -> deterministic, no floating-point atomics anywhere.
//             The kernel itself (rotate_cl_bwd_ell_kernel) lives next to the tables in rotate_il.cu.
// With C channels per voxel, C/8 lanes cooperate on one voxel (8 bf16 = 16 bytes each): a corner fetch
// is one coalesced 16*C/8-byte line, corner indices/weights are computed once per voxel for all
// channels, and the 8x reuse of every source line is served by L1 (a CTA walks a compact run of
// output voxels, whose rotated footprint is a slab of a few tens of KB).
#include "hg_common.cuh"
#include "rotate_common.cuh"

namespace hg {

__device__ __forceinline__ void unpack8_cl(const uint4 &u, float *f)
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8_cl(const float *f)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// linear index of lattice point o = (z*S + y)*S + x in the output layout
__device__ __forceinline__ int out_row(int o, int S, int logS, int out_layout)
{
    if (out_layout != HG_PROJ) return o;
    const int x = o & (S - 1), y = (o >> logS) & (S - 1), z = o >> (2 * logS);
    return (((z << logS) + x) << logS) + y;          // [z][x][y]
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
template <bool kZeroBorder>
__global__ void __launch_bounds__(256) rotate_cl_fwd_kernel(const __nv_bfloat16 *__restrict__ vol,
                                                            const float *__restrict__ a_inv,
                                                            __nv_bfloat16 *__restrict__ out, int C, int S, int logS,
                                                            int out_layout, int voxels_per_cta)
{
    __shared__ float m[12];
    const int n = S * S * S;
    const int b = blockIdx.y;
    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    __syncthreads();
    const int lanes = C >> 3;                           // lanes per voxel
    const int sub = threadIdx.x % lanes;                // which 8-channel slice
    const int vslot = threadIdx.x / lanes, vstep = blockDim.x / lanes;
    const __nv_bfloat16 *vb = vol + (size_t)b * n * C + sub * 8;
    __nv_bfloat16 *ob = out + (size_t)b * n * C + sub * 8;
    const int o_end = min(n, (blockIdx.x + 1) * voxels_per_cta);
    for (int o = blockIdx.x * voxels_per_cta + vslot; o < o_end; o += vstep) {
        float x, y, z;
        lattice_coords(m, o, S, logS, x, y, z);
        Corners c;
        make_corners<false>(x, y, z, S, logS, 0, c);
        float acc[8];
        if (kZeroBorder && !c.inside) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        } else {
            uint4 raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = __ldg(reinterpret_cast<const uint4 *>(vb + (size_t)c.idx[k] * C));
            float f[8];
            unpack8_cl(raw[0], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fmul_rn(c.w[0], f[j]);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                unpack8_cl(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (kZeroBorder) acc[j] = fmaf(c.w[k], f[j], acc[j]);                       // fast mode
                    else acc[j] = __fadd_rn(acc[j], __fmul_rn(c.w[k], f[j]));                  // reference order (:320)
                }
            }
        }
        st_stream_16(ob + (size_t)out_row(o, S, logS, out_layout) * C, pack8_cl(acc));
    }
}

}  // namespace hg

using namespace hg;

static int cl_lanes_ok(int channels)
{
    const int lanes = channels / 8;
    return channels % 8 == 0 && lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0;
}

// called from rotate.cu's dispatchers
int hg_rotate_cl_fwd_impl(const void *vol, const float *a_inv, void *out, int batch, int channels, int size, int logS,
                          int out_layout, int border, cudaStream_t st)
{
    HG_REQUIRE(cl_lanes_ok(channels), HG_ERR_UNSUPPORTED,
               "hg_rotate_fwd: channels-last layouts need channels = 8 * 2^k <= 256 (got %d)", channels);
    const int n = size * size * size;
    const int vpc = n >= 4096 ? 512 : n;                       // output voxels per CTA (2 z-planes at 16^3)
    dim3 grid((n + vpc - 1) / vpc, batch);
    const __nv_bfloat16 *v = static_cast<const __nv_bfloat16 *>(vol);
    __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(out);
    if (border == HG_BORDER_ZERO)
        rotate_cl_fwd_kernel<true><<<grid, 256, 0, st>>>(v, a_inv, o, channels, size, logS, out_layout, vpc);
    else
        rotate_cl_fwd_kernel<false><<<grid, 256, 0, st>>>(v, a_inv, o, channels, size, logS, out_layout, vpc);
    return check_launch("rotate_cl_fwd");
}

// rotate_il.cu: per-sample adjoint tables + the channels-last consumer
size_t hg_rotate_il_ws_bytes(int batch, int size);
int hg_rotate_cl_bwd_table_impl(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace, int batch,
                                int channels, int size, int logS, int out_layout, cudaStream_t st);

size_t hg_rotate_cl_ws_bytes(int batch, int size) { return hg_rotate_il_ws_bytes(batch, size); }

int hg_rotate_cl_bwd_impl(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace,
                          long long workspace_bytes, int batch, int channels, int size, int logS, int out_layout,
                          cudaStream_t st)
{
    HG_REQUIRE(cl_lanes_ok(channels), HG_ERR_UNSUPPORTED,
               "hg_rotate_bwd: channels-last layouts need channels = 8 * 2^k <= 256 (got %d)", channels);
    HG_REQUIRE(size <= 16, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: channels-last path supports size 8 and 16 (got %d)", size);
    HG_REQUIRE(workspace && workspace_bytes >= (long long)hg_rotate_cl_ws_bytes(batch, size), HG_ERR_INVALID_ARG,
               "hg_rotate_bwd: channels-last path needs a workspace of hg_rotate_bwd_workspace_bytes() bytes");
    return hg_rotate_cl_bwd_table_impl(grad_out, a_inv, grad_vol, workspace, batch, channels, size, logS, out_layout, st);
}
