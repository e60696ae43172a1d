// final_layer + tanh: Conv2d(C -> Cout <= 4, k3, p1) + tanh (reference core/models/hologan_generator.py:69-75,
// 141-142), forward and backward.  x: (B,S,S,C) bf16 NHWC, w: torch (Cout,C,3,3) fp32, out: (B,Cout,S,S) fp32 NCHW.
//
// N = 3 output channels: not tensor-core work.  Algorithmic traffic = one read of the
// (B,S,S,C) activation (32 MiB at B = 64) + the 3 MiB image; 0.9 GFLOP of FP32 FMAs per pass (forward, dx, dw
// each) -- the FFMA pipe (~12 us at B = 64) is the binding limit, HBM is ~5 us.  All three kernels use the same
// thread map: L = C/8 lanes cooperate on one pixel (8 channels = one 16-byte load each, so a pixel is one
// coalesced 128-byte line at C = 64) and every thread register-blocks kFcR vertically adjacent pixels, so a
// weight (forward / dx) or a gradient value (dw) fetched from shared memory feeds 8 * kFcR FMAs.
//   forward : partial dot products per lane, xor-shuffle reduction over the L lanes, tanh, store
//   dx      : g = dout * (1 - out^2) staged once per tile in smem (with halo), 27 (tap, co) x 8 channels
//   dw      : 72 accumulators (9 taps x 8 channels) per thread for one co; CTAs stride over tiles, fixed-order
//             reduction (shuffles -> smem -> per-CTA partial -> second kernel): deterministic, no atomics
#include <stdlib.h>

#include "hg_common.cuh"

namespace hg {

constexpr int kFinalMaxCout = 4;
constexpr int kFcThreads = 256;
constexpr int kFcR = 4;              // rows per thread
constexpr int kFcDwCtas = 148;       // CTAs per output channel in the dw kernel (= CTAs of the MMA dw kernel)

struct FcGeom {
    int L, slots, tw, th, TH, tiles_x, tiles_y;
};

static FcGeom fc_geom(int C, int S)
{
    FcGeom g;
    g.L = C / 8;
    g.slots = kFcThreads / g.L;
    g.tw = 1;
    while (g.tw * 2 <= g.slots && g.tw * 2 <= S) g.tw *= 2;
    g.th = g.slots / g.tw;
    g.TH = g.th * kFcR;
    g.tiles_x = (S + g.tw - 1) / g.tw;
    g.tiles_y = (S + g.TH - 1) / g.TH;
    return g;
}

__device__ __forceinline__ void unpack8_f(const uint4 &u, float *f)
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// smem weight layout: [sub][tap*COUT + co][8 channels] with a padded per-sub stride: a thread keeps ONE base
// pointer (its channel octet) and every (tap, co) access is a compile-time offset from it -- no per-access index
// arithmetic.  The stride (72*COUT + 4 floats = an odd number of 16-byte groups) makes the L lanes of a pixel hit
// distinct bank groups.
template <int COUT>
__host__ __device__ constexpr int fc_wstride() { return 9 * COUT * 8 + 4; }

template <int COUT>
__device__ __forceinline__ void load_weights_smem(float *ws, const float *__restrict__ w, int C)
{
    for (int i = threadIdx.x; i < 9 * COUT * C; i += blockDim.x) {
        const int t = i % 9, ci = (i / 9) % C, co = i / (9 * C);          // torch order (co, ci, t): coalesced read
        ws[(ci >> 3) * fc_wstride<COUT>() + (t * COUT + co) * 8 + (ci & 7)] = w[i];
    }
}

// wsub = ws + sub * fc_wstride<COUT>();  tc = tap * COUT + co (compile-time at every call site)
__device__ __forceinline__ void lane_weights(const float *wsub, int tc, float (&wv)[8])
{
    const float4 a = *reinterpret_cast<const float4 *>(wsub + tc * 8);
    const float4 b = *reinterpret_cast<const float4 *>(wsub + tc * 8 + 4);
    wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------------
// forward.  CTA = 8 warps on a tile of 32 pixels x kFcR rows.  The input tile (with halo) is staged in shared
// memory by coalesced 16-byte loads (pixel pitch C*2 + 16 bytes: conflict-free for the reads below); lane =
// pixel column, WARP = channel octet, so every weight fetch is a warp-uniform LDS.128 (one wavefront) feeding
// 8 * kFcR FMAs per thread, and the FP32 pipe -- not shared memory -- is the limit.  Partial sums of the
// octets meet in shared memory; bias + tanh + one coalesced 128-byte store per (co, row).
// -------------------------------------------------------------------------------------------------
constexpr int kFcTw = 32;

__host__ __device__ inline int fc_pixel_pitch(int C) { return C * 2 + 16; }      // bytes

template <int COUT>
__global__ void __launch_bounds__(kFcThreads, 2) final_conv_tanh_fwd_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                            const float *__restrict__ w,
                                                                            const float *__restrict__ bias,
                                                                            float *__restrict__ out, int C, int S, int tiles_x)
{
    extern __shared__ __align__(16) unsigned char smem_fc[];
    const int L = C >> 3, pitch = fc_pixel_pitch(C);
    float *ws = reinterpret_cast<float *>(smem_fc);                           // [L][9*COUT*8 + 4]
    unsigned char *xs = smem_fc + (size_t)L * fc_wstride<COUT>() * sizeof(float);   // [(R+2)][(32+2)] pixels
    float *red = reinterpret_cast<float *>(xs + (size_t)(kFcR + 2) * (kFcTw + 2) * pitch);   // [8][R*COUT][32]
    const int b = blockIdx.y, tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int x0 = tile_x * kFcTw, y0 = tile_y * kFcR;
    load_weights_smem<COUT>(ws, w, C);
    const __nv_bfloat16 *xb = x + (size_t)b * S * S * C;
    for (int i = threadIdx.x; i < (kFcR + 2) * (kFcTw + 2) * L; i += kFcThreads) {
        const int v = i % L, pix = i / L, gx = pix % (kFcTw + 2), gy = pix / (kFcTw + 2);
        const int yy = y0 + gy - 1, xx = x0 + gx - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (yy >= 0 && yy < S && xx >= 0 && xx < S) val = __ldg(reinterpret_cast<const uint4 *>(xb + ((size_t)yy * S + xx) * C) + v);
        *reinterpret_cast<uint4 *>(xs + (size_t)pix * pitch + v * 16) = val;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[kFcR][COUT];
#pragma unroll
    for (int r = 0; r < kFcR; ++r)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[r][co] = 0.f;
    for (int sub = warp; sub < L; sub += kFcThreads / 32) {
        const float *wsub = ws + sub * fc_wstride<COUT>();
        const unsigned char *xl = xs + (size_t)lane * pitch + sub * 16;
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                float f[kFcR][8];
#pragma unroll
                for (int r = 0; r < kFcR; ++r)
                    unpack8_f(*reinterpret_cast<const uint4 *>(xl + (size_t)((r + ty) * (kFcTw + 2) + tx) * pitch), f[r]);
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    float wv[8];
                    lane_weights(wsub, (ty * 3 + tx) * COUT + co, wv);
#pragma unroll
                    for (int r = 0; r < kFcR; ++r)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[r][co] = fmaf(f[r][j], wv[j], acc[r][co]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kFcR; ++r)
#pragma unroll
        for (int co = 0; co < COUT; ++co) red[(warp * kFcR * COUT + r * COUT + co) * 32 + lane] = acc[r][co];
    __syncthreads();
    for (int k = warp; k < kFcR * COUT; k += kFcThreads / 32) {      // k = r * COUT + co
        float sum = 0.f;
#pragma unroll
        for (int wgt = 0; wgt < kFcThreads / 32; ++wgt) sum += red[(wgt * kFcR * COUT + k) * 32 + lane];
        const int r = k / COUT, co = k - r * COUT;
        if (y0 + r < S && x0 + lane < S) out[(((size_t)b * COUT + co) * S + y0 + r) * S + x0 + lane] = tanhf(sum + bias[co]);
    }
}

// g tile of one CTA: gs[co][(TH + 2)][(tw + 2)] = dout * (1 - out^2) with a zero halo outside the image
template <int COUT>
__device__ __forceinline__ float stage_g_tile(float *gs, const float *__restrict__ out, const float *__restrict__ dout, int b,
                                              int S, int x0, int y0, int tw, int TH, int co_first, int n_co)
{
    const int gw = tw + 2, gh = TH + 2;
    float interior = 0.f;
    for (int i = threadIdx.x; i < n_co * gh * gw; i += blockDim.x) {
        const int gx = i % gw, gy = (i / gw) % gh, c = i / (gw * gh);
        const int yy = y0 + gy - 1, xx = x0 + gx - 1;
        float g = 0.f;
        if (yy >= 0 && yy < S && xx >= 0 && xx < S) {
            const size_t k = (((size_t)b * COUT + co_first + c) * S + yy) * S + xx;
            const float o = out[k];
            g = dout[k] * (1.f - o * o);
            if (gy >= 1 && gy <= TH && gx >= 1 && gx <= tw) interior += g;
        }
        gs[i] = g;
    }
    return interior;
}

// -------------------------------------------------------------------------------------------------
// dx[b,y,x,ci] = sum_(co,ty,tx) g[b,co,y-ty+1,x-tx+1] * w[co,ci,ty,tx].  Same thread map as the forward (lane =
// pixel column, warp = channel octet: warp-uniform weight fetches, conflict-free g reads); results are staged
// in shared memory and leave as coalesced 16-byte stores.
// -------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(kFcThreads, 2) final_conv_tanh_bwd_x_kernel(const float *__restrict__ w,
                                                                              const float *__restrict__ out,
                                                                              const float *__restrict__ dout,
                                                                              __nv_bfloat16 *__restrict__ dx, int C, int S,
                                                                              int tiles_x)
{
    extern __shared__ __align__(16) unsigned char smem_fc[];
    const int L = C >> 3, pitch = fc_pixel_pitch(C);
    constexpr int gw = kFcTw + 2, gh = kFcR + 2;
    float *ws = reinterpret_cast<float *>(smem_fc);
    float *gs = ws + (size_t)L * fc_wstride<COUT>();                            // [COUT][gh][gw]
    unsigned char *os = reinterpret_cast<unsigned char *>(gs + COUT * gh * gw + 2);   // [kFcR * 32] pixels, 16-byte aligned below
    os = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(os) + 15) & ~uintptr_t(15));
    const int b = blockIdx.y, tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int x0 = tile_x * kFcTw, y0 = tile_y * kFcR;
    load_weights_smem<COUT>(ws, w, C);
    stage_g_tile<COUT>(gs, out, dout, b, S, x0, y0, kFcTw, kFcR, 0, COUT);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sub = warp; sub < L; sub += kFcThreads / 32) {
        const float *wsub = ws + sub * fc_wstride<COUT>();
        float acc[kFcR][8];
#pragma unroll
        for (int r = 0; r < kFcR; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    float wv[8];
                    lane_weights(wsub, (ty * 3 + tx) * COUT + co, wv);
                    const float *gp = gs + (co * gh + 2 - ty) * gw + lane + 2 - tx;
#pragma unroll
                    for (int r = 0; r < kFcR; ++r) {
                        const float g = gp[r * gw];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(g, wv[j], acc[r][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kFcR; ++r) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(acc[r][2 * j], acc[r][2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t *>(&h);
            }
            *reinterpret_cast<uint4 *>(os + (size_t)(r * kFcTw + lane) * pitch + sub * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kFcR * kFcTw * L; i += kFcThreads) {
        const int v = i % L, pix = i / L, lx = pix % kFcTw, r = pix / kFcTw;
        if (y0 + r < S && x0 + lx < S)
            st_stream_16(dx + (((size_t)b * S + y0 + r) * S + x0 + lx) * C + v * 8,
                         *reinterpret_cast<const uint4 *>(os + (size_t)pix * pitch + v * 16));
    }
}

// -------------------------------------------------------------------------------------------------
// dw[co,ci,ty,tx] = sum_(b,y,x) g[b,co,y,x] * x[b,y+ty-1,x+tx-1,ci];  dbias[co] = sum g.
// grid = (kFcDwCtas, COUT): a CTA owns one co and strides over (sample, tile); partials per CTA:
// part[(cta * COUT + co) * (9*C + 1) + tap*C + ci], bias sum at index 9*C.
// -------------------------------------------------------------------------------------------------
template <int COUT, int LT>
__global__ void __launch_bounds__(kFcThreads, 2) final_conv_tanh_bwd_w_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                              const float *__restrict__ out,
                                                                              const float *__restrict__ dout,
                                                                              float *__restrict__ part, int C, int S, int B,
                                                                              int tw_arg, int tiles_x, int tiles_y)
{
    extern __shared__ __align__(16) float gs[];        // g tile [(TH+2)][(tw+2)], later the reduction scratch
    __shared__ float bsum[kFcThreads / 32];
    const int L = LT > 0 ? LT : (C >> 3), tw = LT > 0 ? kFcThreads / LT : tw_arg;
    const int sub = threadIdx.x % L, slot = threadIdx.x / L;
    const int lx = slot % tw, ly = slot / tw, th = (kFcThreads / L) / tw, TH = th * kFcR;
    const int co = blockIdx.y;
    const int gw = tw + 2;
    const int n_tiles = B * tiles_x * tiles_y;

    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    float gsum = 0.f;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tile_x = tile % tiles_x, tile_y = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        __syncthreads();                                // previous tile's readers are done with gs
        gsum += stage_g_tile<COUT>(gs, out, dout, b, S, tile_x * tw, tile_y * TH, tw, TH, co, 1);
        __syncthreads();
        const int px = tile_x * tw + lx, y0 = tile_y * TH + ly * kFcR;
        const __nv_bfloat16 *xb = x + (size_t)b * S * S * C + sub * 8;
        uint4 raw[kFcR];
#pragma unroll
        for (int r = 0; r < kFcR; ++r) {
            raw[r] = make_uint4(0, 0, 0, 0);
            if (px < S && y0 + r < S) raw[r] = __ldg(reinterpret_cast<const uint4 *>(xb + ((size_t)(y0 + r) * S + px) * C));
        }
#pragma unroll
        for (int r = 0; r < kFcR; ++r) {
            float f[8];
            unpack8_f(raw[r], f);
            // input pixel (y0+r, px) meets g at (y - ty + 1, x - tx + 1); tile-local +1 for the halo
            const float *gp = gs + (ly * kFcR + r + 2) * gw + lx + 2;
#pragma unroll
            for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                for (int tx = 0; tx < 3; ++tx) {
                    const float g = gp[-ty * gw - tx];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[ty * 3 + tx][j] = fmaf(g, f[j], acc[ty * 3 + tx][j]);
                }
        }
    }
    // ---- fixed-order reduction over the CTA's pixel slots ----
    for (int o = L; o < 32; o <<= 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], o);
    }
    gsum = warp_sum(gsum);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = L >= 32 ? kFcThreads / L : kFcThreads / 32;     // partial sums per channel octet after the shuffles
    float *pout = part + ((size_t)blockIdx.x * COUT + co) * (9 * C + 1);
    if (lane == 0) bsum[warp] = gsum;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        __syncthreads();
        if (L >= 32 || lane < L) {
            const int grp = L >= 32 ? slot : warp;
#pragma unroll
            for (int j = 0; j < 8; ++j) gs[grp * C + sub * 8 + j] = acc[t][j];
        }
        __syncthreads();
        for (int ci = threadIdx.x; ci < C; ci += blockDim.x) {
            float s = 0.f;
            for (int g = 0; g < groups; ++g) s += gs[g * C + ci];
            pout[t * C + ci] = s;
        }
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kFcThreads / 32; ++i) s += bsum[i];
        pout[9 * C] = s;
    }
}

// one warp per entry: dw[co][ci][t] = sum_cta part[cta][co][t*C + ci]; dbias[co] = sum_cta part[cta][co][9*C]
__global__ void __launch_bounds__(256) final_conv_reduce_kernel(const float *__restrict__ part, float *__restrict__ dw,
                                                                float *__restrict__ dbias, int C, int Cout, int n_cta)
{
    const int per = 9 * C + 1;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= Cout * per) return;
    const int co = e / per, k = e - co * per;
    float acc = 0.f;
    for (int i = lane; i < n_cta; i += 32) acc += part[((size_t)i * Cout + co) * per + k];
    acc = warp_sum(acc);
    if (lane == 0) {
        if (k < 9 * C) {
            const int t = k / C, ci = k - t * C;
            dw[((size_t)co * C + ci) * 9 + t] = acc;
        } else {
            dbias[co] = acc;
        }
    }
}

}  // namespace hg

using namespace hg;

// final_conv_mma.cu: the same three passes on mma.sync (Cin = 64, Cout = 3, S % 32 == 0)
bool hg_final_conv_mma_supported(int cin, int cout, int size);
int hg_final_conv_fwd_mma(const void *x, const float *w, const float *bias, float *out, int batch, int size, cudaStream_t st);
int hg_final_conv_bwd_x_mma(const float *w, const float *out, const float *dout, void *dx, int batch, int size, cudaStream_t st);
int hg_final_conv_bwd_w_mma(const void *x, const float *out, const float *dout, float *part, int n_cta, int batch, int size,
                            cudaStream_t st);

// Which passes run on the tensor-core kernels: bit 0 forward, bit 1 dx, bit 2 dw.  HG_FINAL_CONV_MMA overrides the
// default (A/B measurements, tests of both implementations; hg_set_option).  Measured on B200 at B = 64
// (profiles/r01j_*, r01k_*): training step 1.916 ms with all three (mask 7), 1.931 ms with the forward only (mask 1);
// the forward kernel alone 43.5 us vs 56.7 us for the SIMT kernel.
static int fc_mma_mask(int cin, int cout, int size)
{
    if (!hg_final_conv_mma_supported(cin, cout, size)) return 0;
    return option(kOptFinalConvMma);
}

static int final_check(const char *who, int batch, int cin, int cout, int size)
{
    HG_REQUIRE(batch > 0 && batch <= 65535 && size > 0, HG_ERR_INVALID_ARG, "%s: bad batch / size", who);
    HG_REQUIRE(cin % 8 == 0 && cin >= 8 && cin <= 256 && (256 % (cin / 8)) == 0, HG_ERR_UNSUPPORTED,
               "%s: Cin must be 8 * 2^k <= 256 (got %d)", who, cin);
    HG_REQUIRE(cout >= 1 && cout <= kFinalMaxCout, HG_ERR_UNSUPPORTED, "%s: Cout must be <= %d (got %d)", who, kFinalMaxCout, cout);
    return HG_OK;
}

// CO = Cout, LT = compile-time lanes per pixel (8 for the hot path Cout 3 / Cin 64 / S >= 32, else 0 = runtime)
#define HG_FC_DISPATCH(COUT_VAR, HOT, CALL)                         \
    switch (COUT_VAR) {                                             \
    case 1: { constexpr int CO = 1, LT = 0; (void)LT; CALL; } break;          \
    case 2: { constexpr int CO = 2, LT = 0; (void)LT; CALL; } break;          \
    case 3:                                                         \
        if (HOT) { constexpr int CO = 3, LT = 8; (void)LT; CALL; }            \
        else { constexpr int CO = 3, LT = 0; (void)LT; CALL; }                \
        break;                                                      \
    default: { constexpr int CO = 4, LT = 0; (void)LT; CALL; } break;         \
    }

template <int COUT>
static size_t fc_weight_smem_floats(int cin) { return (size_t)(cin / 8) * fc_wstride<COUT>(); }
static size_t fc_weight_smem_bytes(int cin, int cout)
{
    return (size_t)(cin / 8) * (9 * cout * 8 + 4) * sizeof(float);
}

extern "C" int hg_final_conv_tanh_fwd(const void *x, const float *w, const float *bias, float *out, int batch, int cin,
                                      int cout, int size, void *stream)
{
    HG_REQUIRE(x && w && bias && out, HG_ERR_INVALID_ARG, "hg_final_conv_tanh_fwd: null pointer");
    int rc = final_check("hg_final_conv_tanh_fwd", batch, cin, cout, size);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fc_mma_mask(cin, cout, size) & 1) return hg_final_conv_fwd_mma(x, w, bias, out, batch, size, st);
    const int tiles_x = (size + kFcTw - 1) / kFcTw, tiles_y = (size + kFcR - 1) / kFcR;
    dim3 grid(tiles_x * tiles_y, batch);
    const size_t smem = fc_weight_smem_bytes(cin, cout) + (size_t)(kFcR + 2) * (kFcTw + 2) * fc_pixel_pitch(cin) +
                        (size_t)(kFcThreads / 32) * kFcR * cout * 32 * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(final_conv_tanh_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(final_conv_tanh_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(final_conv_tanh_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(final_conv_tanh_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    HG_FC_DISPATCH(cout, false, (final_conv_tanh_fwd_kernel<CO><<<grid, kFcThreads, smem, st>>>(
                                    static_cast<const __nv_bfloat16 *>(x), w, bias, out, cin, size, tiles_x)));
    return check_launch("hg_final_conv_tanh_fwd");
}

extern "C" long long hg_final_conv_tanh_bwd_workspace_bytes(int batch, int cin, int cout, int size)
{
    if (batch <= 0 || cin <= 0 || cout <= 0 || size <= 0) return -1;
    return (long long)kFcDwCtas * cout * (9LL * cin + 1) * (long long)sizeof(float);
}

extern "C" int hg_final_conv_tanh_bwd(const void *x, const float *w, const float *out, const float *dout, void *dx, float *dw,
                                      float *dbias, void *workspace, long long workspace_bytes, int batch, int cin, int cout,
                                      int size, void *stream)
{
    HG_REQUIRE(x && w && out && dout && workspace, HG_ERR_INVALID_ARG, "hg_final_conv_tanh_bwd: null pointer");
    HG_REQUIRE((dw == nullptr) == (dbias == nullptr), HG_ERR_INVALID_ARG, "hg_final_conv_tanh_bwd: dw and dbias must both be given or both be null");
    HG_REQUIRE(dx || dw, HG_ERR_INVALID_ARG, "hg_final_conv_tanh_bwd: nothing to compute");
    int rc = final_check("hg_final_conv_tanh_bwd", batch, cin, cout, size);
    if (rc) return rc;
    HG_REQUIRE(workspace_bytes >= hg_final_conv_tanh_bwd_workspace_bytes(batch, cin, cout, size), HG_ERR_INVALID_ARG,
               "hg_final_conv_tanh_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const FcGeom g = fc_geom(cin, size);
    const size_t tile_floats = (size_t)(g.TH + 2) * (g.tw + 2);
    const bool hot = cin == 64 && size >= 32;
    const int mma = fc_mma_mask(cin, cout, size);
    if (dx && (mma & 2)) {
        rc = hg_final_conv_bwd_x_mma(w, out, dout, dx, batch, size, st);
        if (rc) return rc;
    } else if (dx) {
        const int tiles_x = (size + kFcTw - 1) / kFcTw, tiles_y = (size + kFcR - 1) / kFcR;
        dim3 grid(tiles_x * tiles_y, batch);
        const size_t smem = fc_weight_smem_bytes(cin, cout) + (size_t)(cout * (kFcR + 2) * (kFcTw + 2) + 2) * sizeof(float) + 16 +
                            (size_t)kFcR * kFcTw * fc_pixel_pitch(cin);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(final_conv_tanh_bwd_x_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(final_conv_tanh_bwd_x_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(final_conv_tanh_bwd_x_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(final_conv_tanh_bwd_x_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_done = true;
        }
        HG_FC_DISPATCH(cout, false, (final_conv_tanh_bwd_x_kernel<CO><<<grid, kFcThreads, smem, st>>>(
                                        w, out, dout, static_cast<__nv_bfloat16 *>(dx), cin, size, tiles_x)));
        rc = check_launch("hg_final_conv_tanh_bwd(x)");
        if (rc) return rc;
    }
    if (!dw) return HG_OK;                              // input gradient only (the caller runs the weight part elsewhere)
    const int groups = g.L >= 32 ? kFcThreads / g.L : kFcThreads / 32;
    size_t wsmem = tile_floats > (size_t)groups * cin ? tile_floats : (size_t)groups * cin;
    wsmem *= sizeof(float);
    if (mma & 4) {
        rc = hg_final_conv_bwd_w_mma(x, out, dout, static_cast<float *>(workspace), kFcDwCtas, batch, size, st);
        if (rc) return rc;
    } else {
        dim3 wgrid(kFcDwCtas, cout);
        HG_FC_DISPATCH(cout, hot, (final_conv_tanh_bwd_w_kernel<CO, LT><<<wgrid, kFcThreads, wsmem, st>>>(
                                 static_cast<const __nv_bfloat16 *>(x), out, dout, static_cast<float *>(workspace), cin, size, batch,
                                 g.tw, g.tiles_x, g.tiles_y)));
        rc = check_launch("hg_final_conv_tanh_bwd(w)");
        if (rc) return rc;
    }
    const int entries = cout * (9 * cin + 1);
    final_conv_reduce_kernel<<<(entries * 32 + 255) / 256, 256, 0, st>>>(static_cast<const float *>(workspace), dw, dbias, cin,
                                                                        cout, kFcDwCtas);
    return check_launch("hg_final_conv_tanh_bwd(reduce)");
}
