// Rotate-resample on the channels-last layouts of the bf16 pipeline.
//   forward : vol NDHWC (B,S,S,S,C)  ->  out NDHWC or PROJ [b, z, x, y, c]
//             (PROJ is the A operand of the 1x1 projection GEMM: the depth-into-channels fold of
//              reference hologan_generator.py:130-133 happens in the store addressing, zero passes)
//   backward: grad_out (same layout as out) -> grad_vol NDHWC, formulated as a GATHER:
//             every source voxel s sums w(o, s) * g[o] over the outputs o whose 2x2x2 footprint
//             contains s.  Those outputs are found through a per-sample table "cell -> outputs whose
//             floor() lands in the cell" (counting sort, built once per sample by a small pre-kernel,
//             entries sorted so the summation order is fixed) -> deterministic, no floating-point
//             atomics anywhere.
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

// -------------------------------------------------------------------------------------------------
// backward: cell table + gather
// -------------------------------------------------------------------------------------------------
// workspace per sample: uint16 start[n + 1] (padded to n + 8), uint16 items[n]
__host__ __device__ inline size_t cells_ws_elems(int n) { return (size_t)(n + 8) + (size_t)n; }

// One CTA per sample.  Counting sort of the in-range output points by the cell their floor() lands in.
// include_outside: also list the out-of-range outputs, keyed by their CLAMPED floor corner (reference border mode).
__global__ void __launch_bounds__(1024) rotate_cells_kernel(const float *__restrict__ a_inv, uint16_t *__restrict__ ws,
                                                            size_t sample_stride_elems, int S, int logS,
                                                            int include_outside)
{
    extern __shared__ uint32_t sm[];
    const int n = S * S * S;
    uint32_t *count = sm;               // [n]  -> later the running cursor
    uint32_t *start = sm + n;           // [n + 1]
    uint32_t *items = sm + 2 * n + 1;   // [n]
    __shared__ float m[12];
    __shared__ uint32_t warp_tot[32];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 12) m[t] = a_inv[b * 16 + t];
    for (int i = t; i < n; i += blockDim.x) count[i] = 0;
    __syncthreads();
    const float lim = (float)(S - 1);
    for (int o = t; o < n; o += blockDim.x) {
        float x, y, z;
        lattice_coords(m, o, S, logS, x, y, z);
        if (include_outside || (x >= 0.f && x < lim && y >= 0.f && y < lim && z >= 0.f && z < lim)) {
            const int q = (((clampi(__float2int_rd(z), S - 1) << logS) + clampi(__float2int_rd(y), S - 1)) << logS) +
                          clampi(__float2int_rd(x), S - 1);
            atomicAdd(&count[q], 1u);
        }
    }
    __syncthreads();
    // exclusive scan of count[0..n) with blockDim.x threads, n / blockDim.x consecutive entries each
    const int per = n / blockDim.x;                   // n = 512, 4096 with 512 / 1024 threads -> 1 or 4
    uint32_t local = 0;
    for (int i = 0; i < per; ++i) local += count[t * per + i];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        uint32_t w = t < (int)(blockDim.x >> 5) ? warp_tot[t] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (t >= o) wi += v;
        }
        warp_tot[t] = wi - w;                         // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t run = warp_tot[t >> 5] + incl - local;
    for (int i = 0; i < per; ++i) {
        const uint32_t c = count[t * per + i];
        start[t * per + i] = run;
        count[t * per + i] = run;                     // cursor
        run += c;
    }
    if (t == (int)blockDim.x - 1) start[n] = run;
    __syncthreads();
    for (int o = t; o < n; o += blockDim.x) {
        float x, y, z;
        lattice_coords(m, o, S, logS, x, y, z);
        if (include_outside || (x >= 0.f && x < lim && y >= 0.f && y < lim && z >= 0.f && z < lim)) {
            const int q = (((clampi(__float2int_rd(z), S - 1) << logS) + clampi(__float2int_rd(y), S - 1)) << logS) +
                          clampi(__float2int_rd(x), S - 1);
            items[atomicAdd(&count[q], 1u)] = (uint32_t)o;
        }
    }
    __syncthreads();
    // fixed order inside every cell (insertion sort of the handful of entries)
    for (int q = t; q < n; q += blockDim.x) {
        const uint32_t lo = start[q], hi = start[q + 1];
        for (uint32_t i = lo + 1; i < hi; ++i) {
            const uint32_t v = items[i];
            uint32_t j = i;
            while (j > lo && items[j - 1] > v) {
                items[j] = items[j - 1];
                --j;
            }
            items[j] = v;
        }
    }
    __syncthreads();
    uint16_t *wb = ws + (size_t)b * sample_stride_elems;
    for (int i = t; i <= n; i += blockDim.x) wb[i] = (uint16_t)start[i];
    for (int i = t; i < n; i += blockDim.x) wb[n + 8 + i] = (uint16_t)items[i];
}

__global__ void __launch_bounds__(256) rotate_cl_bwd_kernel(const __nv_bfloat16 *__restrict__ grad_out,
                                                            const float *__restrict__ a_inv,
                                                            const uint16_t *__restrict__ ws,
                                                            __nv_bfloat16 *__restrict__ grad_vol, int C, int S, int logS,
                                                            int out_layout, int voxels_per_cta)
{
    __shared__ float m[12];
    const int n = S * S * S;
    const int b = blockIdx.y;
    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    __syncthreads();
    const int lanes = C >> 3;
    const int sub = threadIdx.x % lanes;
    const int vslot = threadIdx.x / lanes, vstep = blockDim.x / lanes;
    const uint16_t *start = ws + (size_t)b * cells_ws_elems(n);
    const uint16_t *items = start + n + 8;
    const __nv_bfloat16 *gb = grad_out + (size_t)b * n * C + sub * 8;
    __nv_bfloat16 *db = grad_vol + (size_t)b * n * C + sub * 8;
    const int s_end = min(n, (blockIdx.x + 1) * voxels_per_cta);
    for (int s = blockIdx.x * voxels_per_cta + vslot; s < s_end; s += vstep) {
        const int sx = s & (S - 1), sy = (s >> logS) & (S - 1), sz = s >> (2 * logS);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int dx = d & 1, dy = (d >> 1) & 1, dz = d >> 2;
            const int qx = sx - dx, qy = sy - dy, qz = sz - dz;
            if (qx < 0 || qy < 0 || qz < 0 || qx > S - 2 || qy > S - 2 || qz > S - 2) continue;
            const int q = (((qz << logS) + qy) << logS) + qx;
            const int lo = __ldg(start + q), hi = __ldg(start + q + 1);
            for (int i = lo; i < hi; ++i) {
                const int o = __ldg(items + i);
                float x, y, z;
                lattice_coords(m, o, S, logS, x, y, z);           // same bits as the forward
                // forward weights of this output point (floor == q by construction)
                const float wx = dx ? __fsub_rn(x, (float)qx) : __fsub_rn((float)(qx + 1), x);
                const float wy = dy ? __fsub_rn(y, (float)qy) : __fsub_rn((float)(qy + 1), y);
                const float wz = dz ? __fsub_rn(z, (float)qz) : __fsub_rn((float)(qz + 1), z);
                const float w = __fmul_rn(__fmul_rn(wx, wy), wz);
                float g[8];
                unpack8_cl(__ldg(reinterpret_cast<const uint4 *>(gb + (size_t)out_row(o, S, logS, out_layout) * C)), g);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
            }
        }
        st_stream_16(db + (size_t)s * C, pack8_cl(acc));
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

size_t hg_rotate_cl_ws_bytes(int batch, int size)
{
    const int n = size * size * size;
    return (size_t)batch * cells_ws_elems(n) * sizeof(uint16_t);
}

// Per-sample cell tables (size 8 or 16): also used by the NCDHW gather adjoint of rotate_il.cu.
// sample_stride_bytes: distance between two samples' tables (0 = densely packed).
int hg_rotate_cells_launch(const float *a_inv, void *workspace, size_t sample_stride_bytes, int batch, int size, int logS,
                           int include_outside, cudaStream_t st)
{
    const int n = size * size * size;
    const int threads = n >= 4096 ? 1024 : 512;
    const size_t smem = (size_t)(3 * n + 1) * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(rotate_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done = true;
    }
    const size_t stride = sample_stride_bytes ? sample_stride_bytes / sizeof(uint16_t) : cells_ws_elems(n);
    rotate_cells_kernel<<<batch, threads, smem, st>>>(a_inv, static_cast<uint16_t *>(workspace), stride, size, logS,
                                                      include_outside);
    return check_launch("rotate_cells");
}

int hg_rotate_cl_bwd_impl(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace,
                          long long workspace_bytes, int batch, int channels, int size, int logS, int out_layout,
                          cudaStream_t st)
{
    HG_REQUIRE(cl_lanes_ok(channels), HG_ERR_UNSUPPORTED,
               "hg_rotate_bwd: channels-last layouts need channels = 8 * 2^k <= 256 (got %d)", channels);
    HG_REQUIRE(size <= 16, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: channels-last path supports size 8 and 16 (got %d)", size);
    HG_REQUIRE(workspace && workspace_bytes >= (long long)hg_rotate_cl_ws_bytes(batch, size), HG_ERR_INVALID_ARG,
               "hg_rotate_bwd: channels-last path needs a workspace of hg_rotate_bwd_workspace_bytes() bytes");
    const int n = size * size * size;
    uint16_t *ws = static_cast<uint16_t *>(workspace);
    int rc = hg_rotate_cells_launch(a_inv, workspace, 0, batch, size, logS, 0, st);
    if (rc) return rc;
    const int vpc = n >= 4096 ? 512 : n;
    dim3 grid((n + vpc - 1) / vpc, batch);
    rotate_cl_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(grad_out), a_inv, ws,
                                               static_cast<__nv_bfloat16 *>(grad_vol), channels, size, logS, out_layout, vpc);
    return check_launch("rotate_cl_bwd");
}
