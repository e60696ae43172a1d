// Rigid-body rotate + trilinear resample of a feature volume (forward gather, backward scatter).
// Replaces core/models/hologan_generator.py:198-331 of the reference (apply_transformation,
// interpolation, meshgrid).  See include/hologan_b200.h for the contract.
//
// HBM-bound op: algorithmic traffic = read the volume once + write it once.  The source slab of a
// (sample, channel-tile) is staged in shared memory with 16-byte coalesced loads, every output voxel
// computes its 8 corner indices / weights ONCE and reuses them for all channels of the tile, and the
// result leaves through coalesced stores.
#include "hg_common.cuh"
#include "rotate_common.cuh"

namespace hg {

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

// rotate_cl.cu
int hg_rotate_cl_fwd_impl(const void *vol, const float *a_inv, void *out, int batch, int channels, int size, int logS,
                          int out_layout, int border, cudaStream_t st);
size_t hg_rotate_cl_ws_bytes(int batch, int size);
int hg_rotate_cl_bwd_impl(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace,
                          long long workspace_bytes, int batch, int channels, int size, int logS, int out_layout,
                          cudaStream_t st);

// rotate_slab.cu (32^3 forward on source-slab tiles, opt-in)
bool hg_rotate_slab32_enabled(int channels, int size, int dtype, int batch);
int hg_rotate_slab32_fwd(const void *vol, const float *a_inv, void *out, int batch, int channels, int dtype, int border,
                         cudaStream_t st);

bool hg_rotate_gather_bwd_enabled(int size, int batch);
int hg_rotate_gather_bwd(const void *grad_out, const float *a_inv, void *grad_vol, int batch, int channels, int size, int logS,
                         int dtype, cudaStream_t st);

// rotate_il.cu
bool hg_rotate_il_supported(int channels, int size, int dtype);
size_t hg_rotate_il_ws_bytes(int batch, int size);
int hg_rotate_il_fwd(const void *vol, const float *a_inv, void *out, int batch, int channels, int size, int logS, int dtype,
                     int border, cudaStream_t st);
int hg_rotate_il_bwd(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace, int batch, int channels,
                     int size, int logS, int dtype, int border, cudaStream_t st);

extern "C" long long hg_rotate_bwd_workspace_bytes(int batch, int size, int in_layout)
{
    if (batch <= 0 || size <= 0) return 0;
    if (in_layout == HG_NDHWC) return (long long)hg_rotate_cl_ws_bytes(batch, size);
    // NCDHW: the gather adjoint (sizes 8, 16) needs per-sample cell + adjoint tables; 32^3 scatters in shared memory
    return size <= 16 ? (long long)hg_rotate_il_ws_bytes(batch, size) : 0;
}

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
    const int tune = border & HG_TUNE_CTA1024;      // launch-shape tuning flag, results unaffected
    border &= ~HG_TUNE_CTA1024;
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
        if (hg_rotate_slab32_enabled(channels, size, dtype, batch))
            return hg_rotate_slab32_fwd(vol, a_inv, out, batch, channels, dtype, border, st);
        if (hg_rotate_il_supported(channels, size, dtype))
            return hg_rotate_il_fwd(vol, a_inv, out, batch, channels, size, logS, dtype, border | tune, st);
        const bool z = border == HG_BORDER_ZERO;
        if (dtype == HG_F32)
            return z ? fwd_ncdhw_by_size<float, true>(vol, a_inv, out, batch, channels, size, logS, st)
                     : fwd_ncdhw_by_size<float, false>(vol, a_inv, out, batch, channels, size, logS, st);
        return z ? fwd_ncdhw_by_size<__nv_bfloat16, true>(vol, a_inv, out, batch, channels, size, logS, st)
                 : fwd_ncdhw_by_size<__nv_bfloat16, false>(vol, a_inv, out, batch, channels, size, logS, st);
    }
    if (in_layout == HG_NDHWC && (out_layout == HG_NDHWC || out_layout == HG_PROJ)) {
        HG_REQUIRE(dtype == HG_BF16, HG_ERR_UNSUPPORTED, "hg_rotate_fwd: channels-last layouts are bf16 only");
        return hg_rotate_cl_fwd_impl(vol, a_inv, out, batch, channels, size, logS, out_layout, border, st);
    }
    return fail(HG_ERR_UNSUPPORTED, "hg_rotate_fwd: layout pair (%d -> %d) not implemented", in_layout, out_layout);
}

extern "C" int hg_rotate_bwd(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace,
                             long long workspace_bytes, int batch, int channels, int size, int in_layout, int out_layout,
                             int dtype, int border, void *stream)
{
    HG_REQUIRE(grad_out && a_inv && grad_vol, HG_ERR_INVALID_ARG, "hg_rotate_bwd: null pointer");
    HG_REQUIRE(batch > 0 && channels > 0, HG_ERR_INVALID_ARG, "hg_rotate_bwd: batch/channels must be positive");
    HG_REQUIRE(batch <= 65535, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: batch > 65535");
    const int logS = log2_exact(size);
    HG_REQUIRE(logS >= 3 && size <= 32, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: size must be 8, 16 or 32 (got %d)", size);
    HG_REQUIRE(dtype == HG_F32 || dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_rotate_bwd: unknown dtype %d", dtype);
    const int tune = border & HG_TUNE_CTA1024;      // launch-shape tuning flag, results unaffected
    border &= ~HG_TUNE_CTA1024;
    HG_REQUIRE(border == HG_BORDER_REFERENCE || border == HG_BORDER_ZERO, HG_ERR_INVALID_ARG,
               "hg_rotate_bwd: unknown border mode %d", border);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_layout == HG_NCDHW && out_layout == HG_NCDHW) {
        if (hg_rotate_gather_bwd_enabled(size, batch))
            return hg_rotate_gather_bwd(grad_out, a_inv, grad_vol, batch, channels, size, logS, dtype, st);
        if (hg_rotate_il_supported(channels, size, dtype) && workspace &&
            workspace_bytes >= (long long)hg_rotate_il_ws_bytes(batch, size))
            return hg_rotate_il_bwd(grad_out, a_inv, grad_vol, workspace, batch, channels, size, logS, dtype, border | tune, st);
        const bool z = border == HG_BORDER_ZERO;
        if (dtype == HG_F32)
            return z ? bwd_ncdhw_by_size<float, true>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st)
                     : bwd_ncdhw_by_size<float, false>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st);
        return z ? bwd_ncdhw_by_size<__nv_bfloat16, true>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st)
                 : bwd_ncdhw_by_size<__nv_bfloat16, false>(grad_out, a_inv, grad_vol, batch, channels, size, logS, st);
    }
    if (in_layout == HG_NDHWC && (out_layout == HG_NDHWC || out_layout == HG_PROJ)) {
        HG_REQUIRE(dtype == HG_BF16, HG_ERR_UNSUPPORTED, "hg_rotate_bwd: channels-last layouts are bf16 only");
        return hg_rotate_cl_bwd_impl(grad_out, a_inv, grad_vol, workspace, workspace_bytes, batch, channels, size, logS,
                                     out_layout, st);
    }
    return fail(HG_ERR_UNSUPPORTED, "hg_rotate_bwd: layout pair (%d -> %d) not implemented", in_layout, out_layout);
}
