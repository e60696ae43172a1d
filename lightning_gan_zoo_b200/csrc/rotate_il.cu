// Rotate-resample on the torch NCDHW layout, "interleaved tile" kernels (sizes 8^3 and 16^3).
// Replaces core/models/hologan_generator.py:198-331 of the reference (apply_transformation, interpolation,
// meshgrid) and the adjoint autograd derives from :292-320.
//
// Why a second NCDHW kernel: the op is HBM-bound on paper (read the volume once, write it once) but a
// straight shared-memory gather is bound by the LSU instead -- 8 scalar LDS per output per channel, and
// for a general view the 32 lanes of a warp hit pseudo-random banks (conflict degree ~3.5).  This design
// removes both factors:
//   * the CTA's channel group is staged CHANNEL-INTERLEAVED: one 16-byte shared-memory unit per source
//     voxel holds 4 fp32 (8 bf16) channels, so one LDS.128 fetches a corner for all of them and the corner
//     indices / weights are amortised over the group;
//   * a unit's position is hashed, u = v ^ (((y << 1) ^ (z << 2)) & 7): any two voxels that differ by at
//     most 1 in every coordinate land in different 16-byte bank groups, and the 8 lanes of a quarter warp
//     (one LDS.128 phase) are mapped to a 2x2x2 block of output voxels, whose k-th corners form exactly such
//     a compact neighbourhood for a rigid view -> almost conflict-free gathers.  The same hash makes the
//     natural-order staging stores conflict-free (x = 0..7 of a row is a permutation of the 8 groups).
// Backward is the same gather machinery applied to the adjoint: every source voxel sums w(o,s) * g[o] over
// the outputs o whose 2x2x2 footprint contains it, found through the per-sample cell tables of
// rotate_cl.cu (counting sort, fixed order) -> deterministic, no floating-point atomics of any kind.
#include "hg_common.cuh"
#include "rotate_common.cuh"

namespace hg {

__device__ __forceinline__ uint32_t ld_stream_4(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ld_stream_2(const void *p)
{
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}

// hashed unit index of voxel v = (z*S + y)*S + x
__device__ __forceinline__ int il_unit(int v, int logS)
{
    const int y = v >> logS, z = v >> (2 * logS);
    return v ^ (((y << 1) ^ (z << 2)) & 7);
}

// Stage the CT = G * CI channels [c0, c0 + CT) of sample b (NCDHW, n = S^3 voxels each) into tile[G][n] units.
template <typename T, int G>
__device__ __forceinline__ void il_stage(const T *__restrict__ src, uint4 *__restrict__ tile, int n, int logS)
{
    if constexpr (sizeof(T) == 4) {
#pragma unroll 4
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            const int u = il_unit(v, logS);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                uint4 q;
                q.x = ld_stream_4(src + (size_t)(g * 4 + 0) * n + v);
                q.y = ld_stream_4(src + (size_t)(g * 4 + 1) * n + v);
                q.z = ld_stream_4(src + (size_t)(g * 4 + 2) * n + v);
                q.w = ld_stream_4(src + (size_t)(g * 4 + 3) * n + v);
                tile[g * n + u] = q;
            }
        }
    } else {
        // bf16: a thread takes the voxel pair (v, v+1), v even -> 8 four-byte loads fill two units
#pragma unroll 2
        for (int v = 2 * threadIdx.x; v < n; v += 2 * blockDim.x) {
            const int u0 = il_unit(v, logS), u1 = il_unit(v + 1, logS);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                uint32_t w[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) w[c] = ld_stream_4(src + (size_t)(g * 8 + c) * n + v);
                uint4 lo, hi;
                lo.x = __byte_perm(w[0], w[1], 0x5410); hi.x = __byte_perm(w[0], w[1], 0x7632);
                lo.y = __byte_perm(w[2], w[3], 0x5410); hi.y = __byte_perm(w[2], w[3], 0x7632);
                lo.z = __byte_perm(w[4], w[5], 0x5410); hi.z = __byte_perm(w[4], w[5], 0x7632);
                lo.w = __byte_perm(w[6], w[7], 0x5410); hi.w = __byte_perm(w[6], w[7], 0x7632);
                tile[g * n + u0] = lo;
                tile[g * n + u1] = hi;
            }
        }
    }
}

template <typename T> struct IlUnit;
template <> struct IlUnit<float> {
    static constexpr int CI = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
};
template <> struct IlUnit<__nv_bfloat16> {
    static constexpr int CI = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *f)
    {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

// lane -> voxel of the j-th 8x2x2 block: quarter warps (lanes 8q..8q+7) are 2x2x2 sub-blocks
__device__ __forceinline__ void il_block_voxel(int j, int lane, int S, int logS, int &x, int &y, int &z)
{
    const int bxn = S >> 3, byn = S >> 1;
    const int bx = j % bxn, t = j / bxn;
    const int by = t % byn, bz = t / byn;
    x = (bx << 3) + (lane & 1) + ((lane >> 3) << 1);
    y = (by << 1) + ((lane >> 1) & 1);
    z = (bz << 1) + ((lane >> 2) & 1);
}

struct IlCorners {
    int u[8];       // hashed unit index of corners a..h (order of hologan_generator.py:278-287)
    float w[8];     // weights a..h (:309-318)
    bool inside;
};

// Same arithmetic as make_corners<> (bit-exact weights / indices); only the address hash differs.
__device__ __forceinline__ void il_corners(float x, float y, float z, int S, int logS, IlCorners &c)
{
    const int fx = __float2int_rd(x), fy = __float2int_rd(y), fz = __float2int_rd(z);
    const int x0 = clampi(fx, S - 1), x1 = clampi(fx + 1, S - 1);
    const int y0 = clampi(fy, S - 1), y1 = clampi(fy + 1, S - 1);
    const int z0 = clampi(fz, S - 1), z1 = clampi(fz + 1, S - 1);
    const float ux = __fsub_rn((float)x1, x), lx = __fsub_rn(x, (float)x0);
    const float uy = __fsub_rn((float)y1, y), ly = __fsub_rn(y, (float)y0);
    const float uz = __fsub_rn((float)z1, z), lz = __fsub_rn(z, (float)z0);
    const int k00 = ((y0 << 1) ^ (z0 << 2)) & 7, k01 = ((y1 << 1) ^ (z0 << 2)) & 7;
    const int k10 = ((y0 << 1) ^ (z1 << 2)) & 7, k11 = ((y1 << 1) ^ (z1 << 2)) & 7;
    const int r00 = ((z0 << logS) + y0) << logS, r01 = ((z0 << logS) + y1) << logS;
    const int r10 = ((z1 << logS) + y0) << logS, r11 = ((z1 << logS) + y1) << logS;
    c.u[0] = r00 | (x0 ^ k00); c.u[1] = r01 | (x0 ^ k01); c.u[2] = r00 | (x1 ^ k00); c.u[3] = r01 | (x1 ^ k01);
    c.u[4] = r10 | (x0 ^ k10); c.u[5] = r11 | (x0 ^ k11); c.u[6] = r10 | (x1 ^ k10); c.u[7] = r11 | (x1 ^ k11);
    const float uxuy = __fmul_rn(ux, uy), uxly = __fmul_rn(ux, ly), lxuy = __fmul_rn(lx, uy), lxly = __fmul_rn(lx, ly);
    c.w[0] = __fmul_rn(uxuy, uz); c.w[1] = __fmul_rn(uxly, uz); c.w[2] = __fmul_rn(lxuy, uz); c.w[3] = __fmul_rn(lxly, uz);
    c.w[4] = __fmul_rn(uxuy, lz); c.w[5] = __fmul_rn(uxly, lz); c.w[6] = __fmul_rn(lxuy, lz); c.w[7] = __fmul_rn(lxly, lz);
    const float lim = (float)(S - 1);
    c.inside = (x >= 0.f) && (x < lim) && (y >= 0.f) && (y < lim) && (z >= 0.f) && (z < lim);
}

__device__ __forceinline__ void il_coords(const float *__restrict__ m, int ox, int oy, int oz, float &x, float &y, float &z)
{
    const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
    x = row_dot(m, fx, fy, fz);
    y = row_dot(m + 4, fx, fy, fz);
    z = row_dot(m + 8, fx, fy, fz);
}

template <typename T> __device__ __forceinline__ void st_stream_elem(T *p, float v);
template <> __device__ __forceinline__ void st_stream_elem<float>(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
template <> __device__ __forceinline__ void st_stream_elem<__nv_bfloat16>(__nv_bfloat16 *p, float v)
{
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<const uint16_t *>(&h)) : "memory");
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
template <typename T, int G, bool kZeroBorder>
__global__ void __launch_bounds__(256 * G) rotate_fwd_il_kernel(const T *__restrict__ vol, const float *__restrict__ a_inv,
                                                                T *__restrict__ out, int C, int S, int logS)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *tile = reinterpret_cast<uint4 *>(smem_raw);
    __shared__ float m[12];
    constexpr int CI = IlUnit<T>::CI, CT = G * CI;

    const int n = S * S * S;
    const int b = blockIdx.y, c0 = blockIdx.x * CT;
    const T *src = vol + ((size_t)b * C + c0) * n;
    T *dst = out + ((size_t)b * C + c0) * n;
    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    il_stage<T, G>(src, tile, n, logS);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int j = warp; j < (n >> 5); j += nwarps) {
        int ox, oy, oz;
        il_block_voxel(j, lane, S, logS, ox, oy, oz);
        const int o = (((oz << logS) + oy) << logS) + ox;
        float x, y, z;
        il_coords(m, ox, oy, oz, x, y, z);
        IlCorners c;
        il_corners(x, y, z, S, logS, c);
        if (kZeroBorder && !c.inside) {
#pragma unroll
            for (int ci = 0; ci < CT; ++ci) st_stream_elem<T>(dst + (size_t)ci * n + o, 0.f);
            continue;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            uint4 raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = tile[g * n + c.u[k]];
            float acc[CI], f[CI];
            IlUnit<T>::unpack(raw[0], f);
#pragma unroll
            for (int i = 0; i < CI; ++i) acc[i] = __fmul_rn(c.w[0], f[i]);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                IlUnit<T>::unpack(raw[k], f);
#pragma unroll
                for (int i = 0; i < CI; ++i) {
                    if (kZeroBorder) acc[i] = fmaf(c.w[k], f[i], acc[i]);                  // fast mode
                    else acc[i] = __fadd_rn(acc[i], __fmul_rn(c.w[k], f[i]));             // reference order (:320)
                }
            }
#pragma unroll
            for (int i = 0; i < CI; ++i) st_stream_elem<T>(dst + (size_t)(g * CI + i) * n + o, acc[i]);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// backward (gather-formulated adjoint).  ws: per-sample cell tables built by hg_rotate_cells_launch:
// uint16 start[n + 1] (padded to n + 8), uint16 items[n].  Zero border: the table holds the in-range
// outputs keyed by floor(); reference border: ALL outputs keyed by the clamped floor corner, and an
// output whose clamped corners coincide contributes the sum of the coinciding weights.
// -------------------------------------------------------------------------------------------------
template <typename T, int G, bool kZeroBorder>
__global__ void __launch_bounds__(256 * G) rotate_bwd_il_kernel(const T *__restrict__ grad_out, const float *__restrict__ a_inv,
                                                                const uint16_t *__restrict__ ws, T *__restrict__ grad_vol,
                                                                int C, int S, int logS)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *tile = reinterpret_cast<uint4 *>(smem_raw);
    __shared__ float m[12];
    constexpr int CI = IlUnit<T>::CI, CT = G * CI;

    const int n = S * S * S;
    const int b = blockIdx.y, c0 = blockIdx.x * CT;
    const T *src = grad_out + ((size_t)b * C + c0) * n;
    T *dst = grad_vol + ((size_t)b * C + c0) * n;
    const uint16_t *start = ws + (size_t)b * ((size_t)(n + 8) + n);
    const uint16_t *items = start + n + 8;
    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    il_stage<T, G>(src, tile, n, logS);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int qmax = kZeroBorder ? S - 2 : S - 1;
    for (int j = warp; j < (n >> 5); j += nwarps) {
        int sx, sy, sz;
        il_block_voxel(j, lane, S, logS, sx, sy, sz);
        const int s = (((sz << logS) + sy) << logS) + sx;
        float acc[CT];
#pragma unroll
        for (int i = 0; i < CT; ++i) acc[i] = 0.f;
#pragma unroll 1
        for (int d = 0; d < 8; ++d) {
            const int dx = d & 1, dy = (d >> 1) & 1, dz = d >> 2;
            const int qx = sx - dx, qy = sy - dy, qz = sz - dz;
            if (qx < 0 || qy < 0 || qz < 0 || qx > qmax || qy > qmax || qz > qmax) continue;
            const int q = (((qz << logS) + qy) << logS) + qx;
            const int lo = __ldg(start + q), hi = __ldg(start + q + 1);
            for (int i = lo; i < hi; ++i) {
                const int o = __ldg(items + i);
                const int ox = o & (S - 1), oy = (o >> logS) & (S - 1), oz = o >> (2 * logS);
                float x, y, z;
                il_coords(m, ox, oy, oz, x, y, z);                  // same bits as the forward
                float w;
                if (kZeroBorder) {
                    // floor == q by construction
                    const float wx = dx ? __fsub_rn(x, (float)qx) : __fsub_rn((float)(qx + 1), x);
                    const float wy = dy ? __fsub_rn(y, (float)qy) : __fsub_rn((float)(qy + 1), y);
                    const float wz = dz ? __fsub_rn(z, (float)qz) : __fsub_rn((float)(qz + 1), z);
                    w = __fmul_rn(__fmul_rn(wx, wy), wz);
                } else {
                    Corners c;
                    make_corners<false>(x, y, z, S, logS, 0, c);
                    w = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) w += (c.idx[k] == s) ? c.w[k] : 0.f;
                }
                const int uo = il_unit(o, logS);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float f[CI];
                    IlUnit<T>::unpack(tile[g * n + uo], f);
#pragma unroll
                    for (int c = 0; c < CI; ++c) acc[g * CI + c] = fmaf(w, f[c], acc[g * CI + c]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CT; ++i) st_stream_elem<T>(dst + (size_t)i * n + s, acc[i]);
    }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
template <typename T, int G, bool Z>
static int launch_fwd_il(const void *vol, const float *a, void *out, int B, int C, int S, int logS, cudaStream_t st)
{
    const size_t smem = (size_t)G * S * S * S * 16;
    auto k = rotate_fwd_il_kernel<T, G, Z>;
    static bool attr_done = false;      // per template instantiation; not a stream operation (graph-capture safe)
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    dim3 grid(C / (G * IlUnit<T>::CI), B);
    k<<<grid, 256 * G, smem, st>>>(static_cast<const T *>(vol), a, static_cast<T *>(out), C, S, logS);
    return check_launch("rotate_fwd_il");
}

template <typename T, int G, bool Z>
static int launch_bwd_il(const void *g, const float *a, const uint16_t *ws, void *gv, int B, int C, int S, int logS,
                         cudaStream_t st)
{
    const size_t smem = (size_t)G * S * S * S * 16;
    auto k = rotate_bwd_il_kernel<T, G, Z>;
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    dim3 grid(C / (G * IlUnit<T>::CI), B);
    k<<<grid, 256 * G, smem, st>>>(static_cast<const T *>(g), a, ws, static_cast<T *>(gv), C, S, logS);
    return check_launch("rotate_bwd_il");
}

}  // namespace hg

using namespace hg;

// rotate_cl.cu
int hg_rotate_cells_launch(const float *a_inv, void *workspace, int batch, int size, int logS, int include_outside,
                           cudaStream_t st);

// Interleaved-tile kernels cover S in {8, 16} and channel counts that are a multiple of one 16-byte unit
// (4 fp32 / 8 bf16 channels); everything else stays on rotate.cu's per-channel tiles.
bool hg_rotate_il_supported(int channels, int size, int dtype)
{
    const int ci = dtype == HG_F32 ? 4 : 8;
    return (size == 8 || size == 16) && channels % ci == 0;
}

static int g_il_groups = 1;     // channel groups per CTA (tuning knob, see hg_rotate_il_set_groups)
extern "C" void hg_rotate_il_set_groups(int g) { g_il_groups = (g == 2) ? 2 : 1; }

int hg_rotate_il_fwd(const void *vol, const float *a_inv, void *out, int batch, int channels, int size, int logS, int dtype,
                     int border, cudaStream_t st)
{
    const bool z = border == HG_BORDER_ZERO;
    const int ci = dtype == HG_F32 ? 4 : 8;
    const bool two = g_il_groups == 2 && size == 16 && channels % (2 * ci) == 0;
#define HG_IL_FWD(T, G, Z) launch_fwd_il<T, G, Z>(vol, a_inv, out, batch, channels, size, logS, st)
    if (dtype == HG_F32) {
        if (two) return z ? HG_IL_FWD(float, 2, true) : HG_IL_FWD(float, 2, false);
        return z ? HG_IL_FWD(float, 1, true) : HG_IL_FWD(float, 1, false);
    }
    if (two) return z ? HG_IL_FWD(__nv_bfloat16, 2, true) : HG_IL_FWD(__nv_bfloat16, 2, false);
    return z ? HG_IL_FWD(__nv_bfloat16, 1, true) : HG_IL_FWD(__nv_bfloat16, 1, false);
#undef HG_IL_FWD
}

int hg_rotate_il_bwd(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace, int batch, int channels,
                     int size, int logS, int dtype, int border, cudaStream_t st)
{
    const bool z = border == HG_BORDER_ZERO;
    int rc = hg_rotate_cells_launch(a_inv, workspace, batch, size, logS, z ? 0 : 1, st);
    if (rc) return rc;
    const uint16_t *ws = static_cast<const uint16_t *>(workspace);
    const int ci = dtype == HG_F32 ? 4 : 8;
    const bool two = g_il_groups == 2 && size == 16 && channels % (2 * ci) == 0;
#define HG_IL_BWD(T, G, Z) launch_bwd_il<T, G, Z>(grad_out, a_inv, ws, grad_vol, batch, channels, size, logS, st)
    if (dtype == HG_F32) {
        if (two) return z ? HG_IL_BWD(float, 2, true) : HG_IL_BWD(float, 2, false);
        return z ? HG_IL_BWD(float, 1, true) : HG_IL_BWD(float, 1, false);
    }
    if (two) return z ? HG_IL_BWD(__nv_bfloat16, 2, true) : HG_IL_BWD(__nv_bfloat16, 2, false);
    return z ? HG_IL_BWD(__nv_bfloat16, 1, true) : HG_IL_BWD(__nv_bfloat16, 1, false);
#undef HG_IL_BWD
}
