// Rigid-body rotate + trilinear resample of a feature volume (forward gather, backward scatter).
// Replaces core/models/hologan_generator.py:198-331 of the reference (apply_transformation,
// interpolation, meshgrid).  See include/hologan_b200.h for the contract.
//
// HBM-bound op: algorithmic traffic = read the volume once + write it once.  The source slab of a
// (sample, channel-tile) is staged in shared memory with 16-byte coalesced loads, every output voxel
// computes its 8 corner indices / weights ONCE and reuses them for all channels of the tile, and the
// result leaves through coalesced stores.
#include "hg_common.cuh"

namespace hg {

// -------------------------------------------------------------------------------------------------
// Coordinates and corners -- bit-exact with the reference CPU path
// -------------------------------------------------------------------------------------------------
struct Corners {
    int idx[8];     // flat (z*S + y)*S + x of corners a..h (hologan_generator.py:278-287)
    float w[8];     // weights a..h (:309-318)
    bool inside;    // all three coordinates in [0, S-1): the only samples that are not ~0
};

// src = A @ [x y z 1]^T.  torch.matmul on CPU (MKL sgemm, k = 4) is reproduced bit-for-bit by this
// sequential chain: one rounding for m0*x, then three fused multiply-adds.
__device__ __forceinline__ float row_dot(const float *__restrict__ r, float x, float y, float z)
{
    float acc = __fmul_rn(r[0], x);
    acc = __fmaf_rn(r[1], y, acc);
    acc = __fmaf_rn(r[2], z, acc);
    acc = __fmaf_rn(r[3], 1.0f, acc);
    return acc;
}

__device__ __forceinline__ int clampi(int v, int hi) { return min(max(v, 0), hi); }

// Shared-memory placement of voxel (z,y,x): rows keep their order but x is XOR-swizzled with a per-row
// key so that neighbours along y and z (and not only along x) fall into different banks.  The views
// of the hot path are near axis-aligned (azimuth ~270 deg, elevation ~90 deg): without the swizzle a
// warp walking along output x reads source voxels S or S^2 elements apart = one bank, a 16-32-way
// conflict.  ysh = log2(rows per 128-byte bank line).
__device__ __forceinline__ int swz(int z, int y, int x, int S, int logS, int ysh)
{
    return (((z << logS) + y) << logS) + (x ^ ((z ^ (y >> ysh)) & (S - 1)));
}

// Corner indices (floor, +1, clamp: :249-261) and weights from the CLAMPED corner as float against
// the UNCLAMPED coordinate, product order (wx*wy)*wz (:301-318).  Explicit _rn intrinsics keep nvcc
// from contracting / re-associating, so fp32 results carry the reference's bits.
// kSwz: idx[] addresses the swizzled shared-memory tile instead of the linear volume.
template <bool kSwz>
__device__ __forceinline__ void make_corners(float x, float y, float z, int S, int logS, int ysh, Corners &c)
{
    const int fx = __float2int_rd(x), fy = __float2int_rd(y), fz = __float2int_rd(z);
    const int x0 = clampi(fx, S - 1), x1 = clampi(fx + 1, S - 1);
    const int y0 = clampi(fy, S - 1), y1 = clampi(fy + 1, S - 1);
    const int z0 = clampi(fz, S - 1), z1 = clampi(fz + 1, S - 1);
    const float ux = __fsub_rn((float)x1, x), lx = __fsub_rn(x, (float)x0);
    const float uy = __fsub_rn((float)y1, y), ly = __fsub_rn(y, (float)y0);
    const float uz = __fsub_rn((float)z1, z), lz = __fsub_rn(z, (float)z0);
    if (kSwz) {
        const int m = S - 1;
        const int k00 = (z0 ^ (y0 >> ysh)) & m, k01 = (z0 ^ (y1 >> ysh)) & m;
        const int k10 = (z1 ^ (y0 >> ysh)) & m, k11 = (z1 ^ (y1 >> ysh)) & m;
        const int r00 = ((z0 << logS) + y0) << logS, r01 = ((z0 << logS) + y1) << logS;
        const int r10 = ((z1 << logS) + y0) << logS, r11 = ((z1 << logS) + y1) << logS;
        c.idx[0] = r00 + (x0 ^ k00); c.idx[1] = r01 + (x0 ^ k01); c.idx[2] = r00 + (x1 ^ k00); c.idx[3] = r01 + (x1 ^ k01);
        c.idx[4] = r10 + (x0 ^ k10); c.idx[5] = r11 + (x0 ^ k11); c.idx[6] = r10 + (x1 ^ k10); c.idx[7] = r11 + (x1 ^ k11);
    } else {
        const int r00 = (z0 * S + y0) * S, r01 = (z0 * S + y1) * S, r10 = (z1 * S + y0) * S, r11 = (z1 * S + y1) * S;
        c.idx[0] = r00 + x0; c.idx[1] = r01 + x0; c.idx[2] = r00 + x1; c.idx[3] = r01 + x1;
        c.idx[4] = r10 + x0; c.idx[5] = r11 + x0; c.idx[6] = r10 + x1; c.idx[7] = r11 + x1;
    }
    const float uxuy = __fmul_rn(ux, uy), uxly = __fmul_rn(ux, ly), lxuy = __fmul_rn(lx, uy), lxly = __fmul_rn(lx, ly);
    c.w[0] = __fmul_rn(uxuy, uz); c.w[1] = __fmul_rn(uxly, uz); c.w[2] = __fmul_rn(lxuy, uz); c.w[3] = __fmul_rn(lxly, uz);
    c.w[4] = __fmul_rn(uxuy, lz); c.w[5] = __fmul_rn(uxly, lz); c.w[6] = __fmul_rn(lxuy, lz); c.w[7] = __fmul_rn(lxly, lz);
    const float lim = (float)(S - 1);
    c.inside = (x >= 0.f) && (x < lim) && (y >= 0.f) && (y < lim) && (z >= 0.f) && (z < lim);
}

__device__ __forceinline__ void lattice_coords(const float *__restrict__ m, int o, int S, int logS, float &x, float &y,
                                               float &z)
{
    const int ox = o & (S - 1), oy = (o >> logS) & (S - 1), oz = o >> (2 * logS);
    const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
    x = row_dot(m, fx, fy, fz);
    y = row_dot(m + 4, fx, fy, fz);
    z = row_dot(m + 8, fx, fy, fz);
}

// out = ((((((w0*v0 + w1*v1) + w2*v2) + ...) + w7*v7): separate multiply and add, left to right,
// like the reference's `wa*Ia + wb*Ib + ...` tensor expression (:320).
__device__ __forceinline__ float blend8(const Corners &c, const float (&v)[8])
{
    float acc = __fmul_rn(c.w[0], v[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fadd_rn(acc, __fmul_rn(c.w[k], v[k]));
    return acc;
}

// -------------------------------------------------------------------------------------------------
// Debug outputs: "grid coordinates" and "sampling indices" of the reference, for bit-exact tests
// -------------------------------------------------------------------------------------------------
__global__ void rotate_debug_kernel(const float *__restrict__ a_inv, float *__restrict__ coords,
                                    int32_t *__restrict__ idx, int B, int S, int logS)
{
    const int n = S * S * S;
    const int b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    float x, y, z;
    lattice_coords(a_inv + b * 16, o, S, logS, x, y, z);
    if (coords) {
        coords[((size_t)0 * B + b) * n + o] = x;
        coords[((size_t)1 * B + b) * n + o] = y;
        coords[((size_t)2 * B + b) * n + o] = z;
    }
    if (idx) {
        Corners c;
        make_corners<false>(x, y, z, S, logS, 0, c);
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[((size_t)k * B + b) * n + o] = b * n + c.idx[k];
    }
}

// -------------------------------------------------------------------------------------------------
// Forward, NCDHW -> NCDHW.  One CTA = (sample b, CT consecutive channels): the CT*S^3 slab is one
// contiguous run in HBM, staged to smem with 16-byte loads.
// -------------------------------------------------------------------------------------------------
template <typename T, int CT, bool kZeroBorder>
__global__ void __launch_bounds__(256) rotate_fwd_ncdhw_kernel(const T *__restrict__ vol, const float *__restrict__ a_inv,
                                                               T *__restrict__ out, int C, int S, int logS)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);
    __shared__ float m[12];

    const int n = S * S * S;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int ct = min(CT, C - c0);
    const T *src = vol + ((size_t)b * C + c0) * n;
    T *dst = out + ((size_t)b * C + c0) * n;

    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    // n is a multiple of 512 (S >= 8), so every channel slab is 16-byte aligned and sized.
    constexpr int kVec = 16 / sizeof(T);
    const int ysh = max(0, (sizeof(T) == 4 ? 5 : 6) - logS);
    const int nvec = ct * n / kVec;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 u = ld_stream_16(reinterpret_cast<const uint4 *>(src) + i);
        const T *e = reinterpret_cast<const T *>(&u);
        const int lin = i * kVec;                       // element index within the CT-channel slab
        const int ci = lin >> (3 * logS), v = lin & (n - 1);
        const int vx = v & (S - 1), vy = (v >> logS) & (S - 1), vz = v >> (2 * logS);
        T *row = tile + ci * n + (((vz << logS) + vy) << logS);
        const int key = (vz ^ (vy >> ysh)) & (S - 1);
#pragma unroll
        for (int j = 0; j < kVec; ++j) row[(vx + j) ^ key] = e[j];
    }
    __syncthreads();

    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        float x, y, z;
        lattice_coords(m, o, S, logS, x, y, z);
        Corners c;
        make_corners<true>(x, y, z, S, logS, ysh, c);
        if (kZeroBorder && !c.inside) {
#pragma unroll
            for (int ci = 0; ci < CT; ++ci)
                if (ci < ct) dst[(size_t)ci * n + o] = from_f32<T>(0.f);
            continue;
        }
#pragma unroll
        for (int ci = 0; ci < CT; ++ci) {
            if (ci < ct) {
                const T *t = tile + ci * n;
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = to_f32<T>(t[c.idx[k]]);
                dst[(size_t)ci * n + o] = from_f32<T>(blend8(c, v));
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Backward, NCDHW.  Scatter into a shared-memory fp32 accumulator tile owned by the CTA (no global
// atomics), then one coalesced store.  grad_vol is fully overwritten.
// -------------------------------------------------------------------------------------------------
template <typename T, int CT, bool kZeroBorder>
__global__ void __launch_bounds__(256) rotate_bwd_ncdhw_kernel(const T *__restrict__ grad_out,
                                                               const float *__restrict__ a_inv, T *__restrict__ grad_vol,
                                                               int C, int S, int logS)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    __shared__ float m[12];

    const int n = S * S * S;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int ct = min(CT, C - c0);
    const T *g = grad_out + ((size_t)b * C + c0) * n;
    T *dst = grad_vol + ((size_t)b * C + c0) * n;

    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];
    for (int i = threadIdx.x; i < CT * n; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int ysh = max(0, 5 - logS);                    // accumulators are fp32

    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        float x, y, z;
        lattice_coords(m, o, S, logS, x, y, z);
        Corners c;
        make_corners<true>(x, y, z, S, logS, ysh, c);
        if (kZeroBorder && !c.inside) continue;
#pragma unroll
        for (int ci = 0; ci < CT; ++ci) {
            if (ci < ct) {
                const float gv = to_f32<T>(g[(size_t)ci * n + o]);
                float *a = acc + ci * n;
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(a + c.idx[k], c.w[k] * gv);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ct * n; i += blockDim.x) {
        const int ci = i >> (3 * logS), v = i & (n - 1);
        const int vx = v & (S - 1), vy = (v >> logS) & (S - 1), vz = v >> (2 * logS);
        dst[i] = from_f32<T>(acc[ci * n + swz(vz, vy, vx, S, logS, ysh)]);
    }
}

// -------------------------------------------------------------------------------------------------
// Host dispatch
// -------------------------------------------------------------------------------------------------
static int log2_exact(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

template <typename T, int CT, bool Z>
static int launch_fwd_ncdhw(const void *vol, const float *a, void *out, int B, int C, int S, int logS, cudaStream_t st)
{
    const size_t smem = (size_t)CT * S * S * S * sizeof(T);
    auto k = rotate_fwd_ncdhw_kernel<T, CT, Z>;
    static bool attr_done = false;      // per template instantiation; not a stream operation (graph-capture safe)
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    dim3 grid((C + CT - 1) / CT, B);
    k<<<grid, 256, smem, st>>>(static_cast<const T *>(vol), a, static_cast<T *>(out), C, S, logS);
    return check_launch("rotate_fwd_ncdhw");
}

template <typename T, int CT, bool Z>
static int launch_bwd_ncdhw(const void *g, const float *a, void *gv, int B, int C, int S, int logS, cudaStream_t st)
{
    const size_t smem = (size_t)CT * S * S * S * sizeof(float);
    auto k = rotate_bwd_ncdhw_kernel<T, CT, Z>;
    static bool attr_done = false;      // per template instantiation; not a stream operation (graph-capture safe)
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    dim3 grid((C + CT - 1) / CT, B);
    k<<<grid, 256, smem, st>>>(static_cast<const T *>(g), a, static_cast<T *>(gv), C, S, logS);
    return check_launch("rotate_bwd_ncdhw");
}

template <typename T, bool Z>
static int fwd_ncdhw_by_size(const void *vol, const float *a, void *out, int B, int C, int S, int logS, cudaStream_t st)
{
    // channel-tile sized so that >= 2 CTAs fit per SM (227 KB): S=16 -> 64 KB fp32 tile (CT 4)
    const size_t bytes1 = (size_t)S * S * S * sizeof(T);
    if (bytes1 * 8 <= 64 * 1024 && C >= 8) return launch_fwd_ncdhw<T, 8, Z>(vol, a, out, B, C, S, logS, st);
    if (bytes1 * 4 <= 64 * 1024 && C >= 4) return launch_fwd_ncdhw<T, 4, Z>(vol, a, out, B, C, S, logS, st);
    if (bytes1 * 2 <= 64 * 1024 && C >= 2) return launch_fwd_ncdhw<T, 2, Z>(vol, a, out, B, C, S, logS, st);
    if (bytes1 <= 200 * 1024) return launch_fwd_ncdhw<T, 1, Z>(vol, a, out, B, C, S, logS, st);
    return fail(HG_ERR_UNSUPPORTED, "rotate_fwd: one %d^3 channel does not fit in shared memory", S);
}

template <typename T, bool Z>
static int bwd_ncdhw_by_size(const void *g, const float *a, void *gv, int B, int C, int S, int logS, cudaStream_t st)
{
    const size_t bytes1 = (size_t)S * S * S * sizeof(float);
    if (bytes1 * 4 <= 64 * 1024 && C >= 4) return launch_bwd_ncdhw<T, 4, Z>(g, a, gv, B, C, S, logS, st);
    if (bytes1 * 2 <= 64 * 1024 && C >= 2) return launch_bwd_ncdhw<T, 2, Z>(g, a, gv, B, C, S, logS, st);
    if (bytes1 <= 200 * 1024) return launch_bwd_ncdhw<T, 1, Z>(g, a, gv, B, C, S, logS, st);
    return fail(HG_ERR_UNSUPPORTED, "rotate_bwd: one %d^3 channel does not fit in shared memory", S);
}

}  // namespace hg

using namespace hg;

extern "C" int hg_rotate_fwd(const void *vol, const float *a_inv, void *out, float *coords_dbg, int32_t *idx_dbg,
                             int batch, int channels, int size, int in_layout, int out_layout, int dtype, int border,
                             void *stream)
{
    HG_REQUIRE(vol && a_inv && out, HG_ERR_INVALID_ARG, "hg_rotate_fwd: null pointer");
    HG_REQUIRE(batch > 0 && channels > 0, HG_ERR_INVALID_ARG, "hg_rotate_fwd: batch/channels must be positive");
    HG_REQUIRE(batch <= 65535, HG_ERR_UNSUPPORTED, "hg_rotate_fwd: batch > 65535");
    const int logS = log2_exact(size);
    HG_REQUIRE(logS >= 3 && size <= 32, HG_ERR_UNSUPPORTED, "hg_rotate_fwd: size must be 8, 16 or 32 (got %d)", size);
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_rotate_fwd: unknown dtype %d", dtype);
    HG_REQUIRE(border == HG_BORDER_REFERENCE || border == HG_BORDER_ZERO, HG_ERR_INVALID_ARG,
               "hg_rotate_fwd: unknown border mode %d", border);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (coords_dbg || idx_dbg) {
        const int n = size * size * size;
        dim3 grid((n + 255) / 256, batch);
        rotate_debug_kernel<<<grid, 256, 0, st>>>(a_inv, coords_dbg, idx_dbg, batch, size, logS);
        int rc = check_launch("rotate_debug");
        if (rc) return rc;
    }
    if (in_layout == HG_NCDHW && out_layout == HG_NCDHW) {
        const bool z = border == HG_BORDER_ZERO;
        if (dtype == HG_F32)
            return z ? fwd_ncdhw_by_size<float, true>(vol, a_inv, out, batch, channels, size, logS, st)
                     : fwd_ncdhw_by_size<float, false>(vol, a_inv, out, batch, channels, size, logS, st);
        return z ? fwd_ncdhw_by_size<__nv_bfloat16, true>(vol, a_inv, out, batch, channels, size, logS, st)
                 : fwd_ncdhw_by_size<__nv_bfloat16, false>(vol, a_inv, out, batch, channels, size, logS, st);
    }
    return fail(HG_ERR_UNSUPPORTED, "hg_rotate_fwd: layout pair (%d -> %d) not implemented", in_layout, out_layout);
}

extern "C" int hg_rotate_bwd(const void *grad_out, const float *a_inv, void *grad_vol, int batch, int channels, int size,
                             int in_layout, int out_layout, int dtype, int border, void *stream)
{
    HG_REQUIRE(grad_out && a_inv && grad_vol, HG_ERR_INVALID_ARG, "hg_rotate_bwd: null pointer");
    HG_REQUIRE(batch > 0 && channels > 0, HG_ERR_INVALID_ARG, "hg_rotate_bwd: batch/channels must be positive");
    HG_REQUIRE(batch <= 65535, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: batch > 65535");
    const int logS = log2_exact(size);
    HG_REQUIRE(logS >= 3 && size <= 32, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: size must be 8, 16 or 32 (got %d)", size);
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_rotate_bwd: unknown dtype %d", dtype);
    HG_REQUIRE(border == HG_BORDER_REFERENCE || border == HG_BORDER_ZERO, HG_ERR_INVALID_ARG,
               "hg_rotate_bwd: unknown border mode %d", border);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_layout == HG_NCDHW && out_layout == HG_NCDHW) {
        const bool z = border == HG_BORDER_ZERO;
        if (dtype == HG_F32)
            return z ? bwd_ncdhw_by_size<float, true>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st)
                     : bwd_ncdhw_by_size<float, false>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st);
        return z ? bwd_ncdhw_by_size<__nv_bfloat16, true>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st)
                 : bwd_ncdhw_by_size<__nv_bfloat16, false>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st);
    }
    return fail(HG_ERR_UNSUPPORTED, "hg_rotate_bwd: layout pair (%d -> %d) not implemented", in_layout, out_layout);
}
