// Spectral normalisation of the discriminator's three 5x5 convolutions (reference
// core/models/hologan_discriminator.py:15 -> torch.nn.utils.spectral_norm, n_power_iterations = 1, eps = 1e-12,
// dim = 0), grouped over the layers:
//     t = W^T u;  v = t / max(|t|, eps);  s = W v;  u = s / max(|s|, eps);  sigma = u . s;  W_sn = W / sigma
// with W = weight_orig viewed as a (Cout, K = Cin * taps) matrix.  The stock path is ~16 small launches per layer
// and call (reshape copy, 3 gemv, 2 norms, clamps, divisions, clones, dot, the fp32 division and the bf16 cast);
// here it is four launches for ALL layers, and the last pass writes the conv operand (bf16 or fp32) directly.
// Backward (u, v are constants, as in torch):  dW_orig = dW / sigma - (sum(dW * W_orig) / sigma^2) * u v^T.
//
// Everything runs in the weight's PHYSICAL element order (row o, position j): for a channels_last parameter
// j = tap * Cin + ci, for a contiguous one j = ci * taps + tap = the logical index k.  Norms and dot products do
// not care about the order; only the `v` buffer (logical order, state_dict-compatible) is permuted on access.
// Fixed summation orders: deterministic, no atomics.
#include "hg_common.cuh"

namespace hg {

constexpr int kSnMaxLayers = HG_SN_MAX_LAYERS;
constexpr int kSnOSplit = 32;       // row splits of the W^T u pass (<= 16 rows per thread: every load in flight at once)
constexpr int kSnDotBlocks = 256;   // partial sums of the backward dot product (one per thread of the consumer CTA)

struct SnLayer {
    const float *w;       // (Cout, K) physical order
    float *u, *v;         // module buffers (updated when iterating)
    void *out;            // W / sigma, same physical order
    float *state;         // per-call: [0] sigma, [1] 1/sigma, [4 .. 4+Cout) u, then v in physical order
    float *t_part;        // workspace [kSnOSplit][K]
    float *s;             // workspace [Cout]
    int cout, cin, taps, K;
};
struct SnParams {
    SnLayer l[kSnMaxLayers];
    int channels_last, iterate, out_dtype;
    float eps;
};
struct SnBwdLayer {
    const void *dw;       // gradient w.r.t. W / sigma, physical order, dtype dw_dtype
    const float *w;
    const float *state;
    float *dw_orig;
    float *partial;       // workspace [kSnDotBlocks]
    int cout, K;
};
struct SnBwdParams {
    SnBwdLayer l[kSnMaxLayers];
    int dw_dtype, accumulate;
};

__device__ __forceinline__ int sn_logical(int j, int cin, int taps, int channels_last)
{
    return channels_last ? (j % cin) * taps + j / cin : j;
}

// 16-byte alignment of every pointer of a vectorised path (parameters and gradient views usually are; a view into a
// flat buffer behind an odd-sized tensor is not -> scalar path)
__device__ __forceinline__ bool sn_aligned16(const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr)
{
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

template <int NT> __device__ __forceinline__ float sn_block_sum(float v, float *sh)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) r += sh[i];      // same order in every thread / CTA
    return r;
}

// t_part[os][j] = sum over the os-th slice of rows of W[o][j] * u[o]
__global__ void __launch_bounds__(128) sn_tu_kernel(const __grid_constant__ SnParams p)
{
    const SnLayer &L = p.l[blockIdx.z];
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= L.K) return;
    const int per = (L.cout + kSnOSplit - 1) / kSnOSplit;
    const int o0 = blockIdx.y * per, o1 = min(L.cout, o0 + per);
    float acc = 0.f;
#pragma unroll 8
    for (int o = o0; o < o1; ++o) acc = fmaf(__ldg(L.w + (size_t)o * L.K + j), __ldg(L.u + o), acc);
    L.t_part[(size_t)blockIdx.y * L.K + j] = acc;
}

// v = normalize(t) (or the stored v when not iterating), kept in physical order in the state
__global__ void __launch_bounds__(1024) sn_v_kernel(const __grid_constant__ SnParams p)
{
    __shared__ float sh[32];
    const SnLayer &L = p.l[blockIdx.x];
    float *vphys = L.state + 4 + L.cout;
    if (!p.iterate) {
        for (int j = threadIdx.x; j < L.K; j += 1024) vphys[j] = L.v[sn_logical(j, L.cin, L.taps, p.channels_last)];
        return;
    }
    float ss = 0.f;
    for (int j = threadIdx.x; j < L.K; j += 1024) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < kSnOSplit; ++s) t += L.t_part[(size_t)s * L.K + j];
        vphys[j] = t;
        ss = fmaf(t, t, ss);
    }
    const float nrm = fmaxf(sqrtf(sn_block_sum<1024>(ss, sh)), p.eps);
    for (int j = threadIdx.x; j < L.K; j += 1024) {
        const float v = vphys[j] / nrm;
        vphys[j] = v;
        L.v[sn_logical(j, L.cin, L.taps, p.channels_last)] = v;
    }
}

// s[o] = W[o] . v  (one CTA per row: 16-byte loads, all of a thread's loads independent, fixed-order block sum)
__global__ void __launch_bounds__(256) sn_s_kernel(const __grid_constant__ SnParams p)
{
    __shared__ float sh[8];
    const SnLayer &L = p.l[blockIdx.y];
    const int o = blockIdx.x;
    if (o >= L.cout) return;
    const float *row = L.w + (size_t)o * L.K, *vphys = L.state + 4 + L.cout;
    float acc = 0.f;
    if (((L.K | L.cout) & 3) == 0 && sn_aligned16(L.w, L.state)) {   // rows and the v vector are 16-byte aligned
#pragma unroll 4
        for (int j = threadIdx.x * 4; j < L.K; j += 1024) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(row + j));
            const float4 b = *reinterpret_cast<const float4 *>(vphys + j);
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
    } else {
        for (int j = threadIdx.x; j < L.K; j += 256) acc = fmaf(__ldg(row + j), vphys[j], acc);
    }
    const float tot = sn_block_sum<256>(acc, sh);
    if (threadIdx.x == 0) L.s[o] = tot;
}

// sigma (every CTA recomputes it from s, identically), u, and out = W / sigma
template <typename TO> __global__ void __launch_bounds__(256) sn_scale_kernel(const __grid_constant__ SnParams p)
{
    __shared__ float sh[8];
    const SnLayer &L = p.l[blockIdx.y];
    float *u_saved = L.state + 4;
    float part = 0.f;
    if (p.iterate) {
        for (int o = threadIdx.x; o < L.cout; o += 256) part = fmaf(L.s[o], L.s[o], part);
    } else {
        for (int o = threadIdx.x; o < L.cout; o += 256) part = fmaf(__ldg(L.u + o), L.s[o], part);
    }
    const float tot = sn_block_sum<256>(part, sh);
    float sigma, nrm = 1.f;
    if (p.iterate) {
        nrm = fmaxf(sqrtf(tot), p.eps);
        sigma = tot / nrm;                               // u . s with u = s / max(|s|, eps)
    } else {
        sigma = tot;
    }
    if (blockIdx.x == 0) {
        for (int o = threadIdx.x; o < L.cout; o += 256) {
            const float un = p.iterate ? L.s[o] / nrm : L.u[o];
            u_saved[o] = un;
            if (p.iterate) L.u[o] = un;
        }
        if (threadIdx.x == 0) {
            L.state[0] = sigma;
            L.state[1] = 1.f / sigma;
        }
    }
    const size_t n = (size_t)L.cout * L.K;
    TO *out = static_cast<TO *>(L.out);
    if (out == nullptr) return;                          // sigma only: the caller folds 1 / sigma into its own weight pack
    if ((L.K & 3) == 0 && sn_aligned16(L.w, out)) {      // four elements per thread and trip: 16-byte loads, 8/16-byte stores
        for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * 1024) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(L.w + i));
            const float q0 = __fdiv_rn(a.x, sigma), q1 = __fdiv_rn(a.y, sigma), q2 = __fdiv_rn(a.z, sigma), q3 = __fdiv_rn(a.w, sigma);
            if constexpr (sizeof(TO) == 4) {
                *reinterpret_cast<float4 *>(out + i) = make_float4(q0, q1, q2, q3);
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(q0, q1), hi = __floats2bfloat162_rn(q2, q3);
                *reinterpret_cast<uint2 *>(out + i) = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
            }
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = from_f32<TO>(__fdiv_rn(__ldg(L.w + i), sigma));
}

// ---- backward --------------------------------------------------------------------------------------------------
template <typename TG> __global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const __grid_constant__ SnBwdParams p)
{
    __shared__ float sh[8];
    const SnBwdLayer &L = p.l[blockIdx.y];
    const TG *dw = static_cast<const TG *>(L.dw);
    const size_t n = (size_t)L.cout * L.K;
    float acc = 0.f;
    if ((L.K & 3) == 0 && sn_aligned16(L.w, dw)) {
#pragma unroll 4
        for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (size_t)kSnDotBlocks * 1024) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(L.w + i));
            float g[4];
            if constexpr (sizeof(TG) == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(dw + i));
                g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
            } else {
                const uint2 q = __ldg(reinterpret_cast<const uint2 *>(dw + i));
                g[0] = __uint_as_float(q.x << 16); g[1] = __uint_as_float(q.x & 0xffff0000u);
                g[2] = __uint_as_float(q.y << 16); g[3] = __uint_as_float(q.y & 0xffff0000u);
            }
            acc = fmaf(g[0], a.x, acc); acc = fmaf(g[1], a.y, acc); acc = fmaf(g[2], a.z, acc); acc = fmaf(g[3], a.w, acc);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)kSnDotBlocks * 256)
            acc = fmaf(to_f32<TG>(dw[i]), __ldg(L.w + i), acc);
    }
    const float tot = sn_block_sum<256>(acc, sh);
    if (threadIdx.x == 0) L.partial[blockIdx.x] = tot;
}

template <typename TG> __global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const __grid_constant__ SnBwdParams p)
{
    __shared__ float sh[8];
    const SnBwdLayer &L = p.l[blockIdx.y];
    const TG *dw = static_cast<const TG *>(L.dw);
    static_assert(kSnDotBlocks == 256, "one partial per thread");
    const float dot = sn_block_sum<256>(L.partial[threadIdx.x], sh);      // fixed order: same bits in every CTA
    const float inv = L.state[1];
    const float c = dot * inv * inv;
    const float *u = L.state + 4, *vphys = L.state + 4 + L.cout;
    const size_t n = (size_t)L.cout * L.K;
    if (((L.K | L.cout) & 3) == 0 && n < (1ull << 31) && sn_aligned16(dw, L.state, L.dw_orig)) {   // four elements of one row per thread and trip
        const unsigned K = (unsigned)L.K;
        for (unsigned i = (blockIdx.x * 256u + threadIdx.x) * 4u; i < (unsigned)n; i += gridDim.x * 1024u) {
            const unsigned o = i / K, j = i - o * K;
            float g[4];
            if constexpr (sizeof(TG) == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(dw + i));
                g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
            } else {
                const uint2 q = __ldg(reinterpret_cast<const uint2 *>(dw + i));
                g[0] = __uint_as_float(q.x << 16); g[1] = __uint_as_float(q.x & 0xffff0000u);
                g[2] = __uint_as_float(q.y << 16); g[3] = __uint_as_float(q.y & 0xffff0000u);
            }
            const float cu = c * u[o];
            const float4 v = *reinterpret_cast<const float4 *>(vphys + j);
            float4 r = make_float4(g[0] * inv - cu * v.x, g[1] * inv - cu * v.y, g[2] * inv - cu * v.z, g[3] * inv - cu * v.w);
            float4 *dst = reinterpret_cast<float4 *>(L.dw_orig + i);
            if (p.accumulate) {
                const float4 old = *dst;
                r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
            }
            *dst = r;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int o = (int)(i / L.K), j = (int)(i - (size_t)o * L.K);
        const float g = to_f32<TG>(dw[i]) * inv - c * u[o] * vphys[j];
        L.dw_orig[i] = p.accumulate ? L.dw_orig[i] + g : g;
    }
}

static size_t sn_align(size_t v) { return (v + 63) / 64 * 64; }

}  // namespace hg

using namespace hg;

extern "C" long long hg_spectral_norm_state_floats(int cout, int cin, int taps) { return 4ll + cout + (long long)cin * taps; }

extern "C" long long hg_spectral_norm_workspace_bytes(int layers, const int *cout, const int *cin, const int *taps)
{
    if (layers < 1 || layers > kSnMaxLayers || !cout || !cin || !taps) return -1;
    size_t total = 0;
    for (int i = 0; i < layers; ++i) {
        const size_t K = (size_t)cin[i] * taps[i];
        total += sn_align((size_t)kSnOSplit * K * 4) + sn_align((size_t)cout[i] * 4) + sn_align(kSnDotBlocks * 4);
    }
    return (long long)total;
}

extern "C" int hg_spectral_norm_fwd(int layers, const float *const *w, float *const *u, float *const *v, void *const *w_out,
                                    float *const *state, const int *cout, const int *cin, const int *taps, int channels_last,
                                    int power_iteration, float eps, int out_dtype, void *workspace, long long workspace_bytes,
                                    void *stream)
{
    HG_REQUIRE(layers >= 1 && layers <= kSnMaxLayers, HG_ERR_INVALID_ARG, "hg_spectral_norm_fwd: 1..%d layers", kSnMaxLayers);
    HG_REQUIRE(w && u && v && w_out && state && cout && cin && taps && workspace, HG_ERR_INVALID_ARG,
               "hg_spectral_norm_fwd: null pointer");
    HG_REQUIRE(out_dtype == HG_F32 || out_dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_spectral_norm_fwd: dtype");
    HG_REQUIRE(workspace_bytes >= hg_spectral_norm_workspace_bytes(layers, cout, cin, taps), HG_ERR_INVALID_ARG,
               "hg_spectral_norm_fwd: workspace too small");
    SnParams p{};
    p.channels_last = channels_last != 0;
    p.iterate = power_iteration != 0;
    p.out_dtype = out_dtype;
    p.eps = eps;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    int max_k = 0, max_cout = 0;
    size_t max_n = 0;
    for (int i = 0; i < layers; ++i) {
        HG_REQUIRE(w[i] && u[i] && v[i] && state[i], HG_ERR_INVALID_ARG, "hg_spectral_norm_fwd: null layer pointer");
        HG_REQUIRE(cout[i] > 0 && cin[i] > 0 && taps[i] > 0, HG_ERR_INVALID_ARG, "hg_spectral_norm_fwd: non-positive dims");
        SnLayer &L = p.l[i];
        L.w = w[i]; L.u = u[i]; L.v = v[i]; L.out = w_out[i]; L.state = state[i];
        L.cout = cout[i]; L.cin = cin[i]; L.taps = taps[i]; L.K = cin[i] * taps[i];
        L.t_part = reinterpret_cast<float *>(ws); ws += sn_align((size_t)kSnOSplit * L.K * 4);
        L.s = reinterpret_cast<float *>(ws); ws += sn_align((size_t)L.cout * 4);
        ws += sn_align(kSnDotBlocks * 4);
        if (L.K > max_k) max_k = L.K;
        if (L.cout > max_cout) max_cout = L.cout;
        if ((size_t)L.cout * L.K > max_n) max_n = (size_t)L.cout * L.K;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.iterate) sn_tu_kernel<<<dim3((max_k + 127) / 128, kSnOSplit, layers), 128, 0, st>>>(p);
    sn_v_kernel<<<layers, 1024, 0, st>>>(p);
    sn_s_kernel<<<dim3(max_cout, layers), 256, 0, st>>>(p);
    const size_t want = (max_n + 256 * 8 - 1) / (256 * 8), cap = (size_t)sm_count() * 2;
    const int gx = (int)(want < cap ? want : cap);
    if (out_dtype == HG_F32) sn_scale_kernel<float><<<dim3(gx, layers), 256, 0, st>>>(p);
    else sn_scale_kernel<__nv_bfloat16><<<dim3(gx, layers), 256, 0, st>>>(p);
    return check_launch("spectral_norm_fwd");
}

extern "C" int hg_spectral_norm_bwd(int layers, const void *const *dw, const float *const *w, const float *const *state,
                                    float *const *dw_orig, const int *cout, const int *cin, const int *taps, int accumulate,
                                    int dw_dtype, void *workspace, long long workspace_bytes, void *stream)
{
    HG_REQUIRE(layers >= 1 && layers <= kSnMaxLayers, HG_ERR_INVALID_ARG, "hg_spectral_norm_bwd: 1..%d layers", kSnMaxLayers);
    HG_REQUIRE(dw && w && state && dw_orig && cout && cin && taps && workspace, HG_ERR_INVALID_ARG,
               "hg_spectral_norm_bwd: null pointer");
    HG_REQUIRE(dw_dtype == HG_F32 || dw_dtype == HG_BF16, HG_ERR_INVALID_ARG, "hg_spectral_norm_bwd: dtype");
    HG_REQUIRE(workspace_bytes >= hg_spectral_norm_workspace_bytes(layers, cout, cin, taps), HG_ERR_INVALID_ARG,
               "hg_spectral_norm_bwd: workspace too small");
    SnBwdParams p{};
    p.dw_dtype = dw_dtype;
    p.accumulate = accumulate != 0;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    size_t max_n = 0;
    for (int i = 0; i < layers; ++i) {
        HG_REQUIRE(dw[i] && w[i] && state[i] && dw_orig[i], HG_ERR_INVALID_ARG, "hg_spectral_norm_bwd: null layer pointer");
        HG_REQUIRE(cout[i] > 0 && cin[i] > 0 && taps[i] > 0, HG_ERR_INVALID_ARG, "hg_spectral_norm_bwd: non-positive dims");
        SnBwdLayer &L = p.l[i];
        L.dw = dw[i]; L.w = w[i]; L.state = state[i]; L.dw_orig = dw_orig[i];
        L.cout = cout[i]; L.K = cin[i] * taps[i];
        ws += sn_align((size_t)kSnOSplit * L.K * 4) + sn_align((size_t)L.cout * 4);
        L.partial = reinterpret_cast<float *>(ws); ws += sn_align(kSnDotBlocks * 4);
        if ((size_t)L.cout * L.K > max_n) max_n = (size_t)L.cout * L.K;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t want = (max_n + 256 * 8 - 1) / (256 * 8), cap = (size_t)sm_count() * 2;
    const int gx = (int)(want < cap ? want : cap);
    if (dw_dtype == HG_F32) {
        sn_bwd_dot_kernel<float><<<dim3(kSnDotBlocks, layers), 256, 0, st>>>(p);
        sn_bwd_apply_kernel<float><<<dim3(gx, layers), 256, 0, st>>>(p);
    } else {
        sn_bwd_dot_kernel<__nv_bfloat16><<<dim3(kSnDotBlocks, layers), 256, 0, st>>>(p);
        sn_bwd_apply_kernel<__nv_bfloat16><<<dim3(gx, layers), 256, 0, st>>>(p);
    }
    return check_launch("spectral_norm_bwd");
}
