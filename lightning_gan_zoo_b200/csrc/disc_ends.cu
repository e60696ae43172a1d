// The two ends of the HoloGAN discriminator (reference core/models/hologan_discriminator.py):
//   conv0 : Conv2d(3 -> 64, k5, s2, p2) + bias + LeakyReLU(0.2)                         (:30, :58)
//   heads : logits = linear1(h);  z = tanh(linear3(leaky_relu(linear2(h), 0.2)))        (:41-51, :64-68)
// forward and backward, in the layouts of the bf16 pipeline: conv0 reads the fp32 NCHW image (real batch or the
// generator's output) and writes its activation in the 2x2 space-to-depth order that the first spectral-norm block's
// tap GEMM consumes (hg_conv5s2_fwd); the heads read the last block's channels-last activation (B, H, W, 512) -- the
// reference flattens (c, h, w), so the kernels address the torch-layout weights with feature f = c * HW + hw.
//
// conv0 has K = 75 and three input channels: it is bandwidth-bound (3 MB in, 8 MB out at B = 64) but needs 0.3 GFMA,
// which on the FP32 pipe costs more than the memory traffic, so it runs on the warp-level tensor cores
// (mma.sync.m16n8k16, bf16 operands, fp32 accumulation) as an im2col GEMM whose A fragments are gathered straight
// from a staged bf16 input patch.  The heads are a 64 x 129 x 8192 problem (0.07 GFMA): fp32 SIMT, split over the
// feature axis across CTAs, partials summed in a fixed order.  Everything is deterministic; no atomics.
#include "hg_common.cuh"
#include "mma_sync.cuh"

namespace hg {

// -------------------------------------------------------------------------------------------------
// 3 <-> 64 channel stride-2 convolutions on mma.sync: the discriminator's first convolution (kernel 5, padding 2) and
// the generator's patched 128 x 128 head ConvTranspose2d(64 -> 3, k4, s2, p1) + tanh (reference
// core/models/hologan_generator.py:71-72 with the missing stride) share three kernels, templated on <KS, PAD, HEAD>:
//   img2map : 3-channel image (2S x 2S) -> 64-channel map (S x S)      conv0 forward        | head dx
//   map2img : 64-channel map (S x S)    -> 3-channel image (2S x 2S)   conv0 dx             | head forward (+ bias, tanh)
//   dw      : sum over pixels of map (x) image patches                 conv0 dw (+ dbias)   | head dw
// The weight element that couples map channel n (of 64), image channel c (of 3) and tap (ky, kx) sits at
// w[((n * 3 + c) * KS + ky) * KS + kx] in BOTH torch layouts: Conv2d (64, 3, 5, 5) and ConvTranspose2d (64, 3, 4, 4).
// Geometry: map pixel i and tap k touch image pixel 2i - PAD + k.
// -------------------------------------------------------------------------------------------------
constexpr int kC0Threads = 256;
constexpr int kC0Cout = 64, kC0Cin = 3;
constexpr int kC0TH = 8, kC0TW = 16;                    // map tile: one row of 16 pixels per warp

template <int KS> struct C0Geom {
    static constexpr int PH = 2 * kC0TH + KS - 2, PW = 2 * kC0TW + KS - 2;      // image patch of a tile
    static constexpr int PPitch = (PW + 2) & ~1;                                // even: bf16 pairs stay 4-byte aligned
    static constexpr int KX = (KS + 1) & ~1;                                    // taps per row padded to even (kx == KS: zero weight)
    static constexpr int KReal = kC0Cin * KS * KX;                              // im2col columns k = (c * KS + ky) * KX + kx
    static constexpr int KSteps = (KReal + 15) / 16;
    static constexpr int KPad = KSteps * 16;
};
constexpr int kC0OnesCol = 90;                          // conv0 dw: the unused im2col column 90 is all ones -> its "dw" is dbias

// element offset of im2col column k inside the patch (relative to the pixel's top-left tap); columns >= KReal map to 0
template <int KS> __device__ __forceinline__ int c0_col_offset(int k)
{
    using G = C0Geom<KS>;
    if (k >= G::KReal) return 0;
    const int kx = k % G::KX, r = k / G::KX, ky = r % KS, ci = r / KS;
    return (ci * G::PH + ky) * G::PPitch + kx;
}

// The 3-channel image the kernels read: conv0 -> x (fp32 NCHW); head -> g = dout * (1 - out^2) from two fp32 NCHW tensors
struct C0Image {
    const float *a;       // x, or dout
    const float *b;       // null, or out (tanh output)
    __device__ __forceinline__ float at(size_t i) const
    {
        const float v = __ldg(a + i);
        if (!b) return v;
        const float o = __ldg(b + i);
        return v * (1.f - o * o);
    }
};

// stage the image patch of the tile as bf16 [3][PH][PPitch]; out-of-image elements are zero
template <int KS, int PAD>
__device__ __forceinline__ void c0_stage_patch(__nv_bfloat16 *patch, const C0Image &img, size_t img_off, int S, int oy0, int ox0)
{
    using G = C0Geom<KS>;
    const int iy0 = 2 * oy0 - PAD, ix0 = 2 * ox0 - PAD;
    // batches of 4 elements per thread: the loads of a batch are issued together (one memory round trip per batch, not
    // per element -- the loop body is a conditional load followed by a shared-memory store, which the compiler keeps in order)
    constexpr int N = kC0Cin * G::PH * G::PPitch, U = 4;
    for (int i0 = threadIdx.x; i0 < N; i0 += U * kC0Threads) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * kC0Threads;
            const int px = i % G::PPitch, r = i / G::PPitch, py = r % G::PH, ci = r / G::PH;
            const int yy = iy0 + py, xx = ix0 + px;
            const bool ok = i < N && px < G::PW && yy >= 0 && yy < S && xx >= 0 && xx < S;
            v[u] = ok ? img.at(img_off + ((size_t)ci * S + yy) * S + xx) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * kC0Threads;
            if (i < N) patch[i] = __float2bfloat16_rn(v[u]);
        }
    }
}

// row of the 64-channel map that holds pixel (oy, ox): space-to-depth order (conv0, S4 = S2 / 2) or plain (head)
template <bool HEAD> __device__ __forceinline__ int c0_map_row(int oy, int ox, int S2)
{
    if (HEAD) return oy * S2 + ox;
    return (((oy >> 1) * (S2 >> 1) + (ox >> 1)) << 2) + ((oy & 1) << 1) + (ox & 1);
}

// ---- image -> map --------------------------------------------------------------------------------------
// im2col GEMM  M = 16 map pixels per warp, K = KPad, N = 64: A fragments gathered from the staged patch, B fragments
// (weights) fragment-ready in shared memory.  conv0: + bias, LeakyReLU; head dx: raw.
template <int KS, int PAD, bool HEAD>
__global__ void __launch_bounds__(kC0Threads, 3) c0_img2map_kernel(C0Image img, const float *__restrict__ w, const float *__restrict__ bias,
                                                                __nv_bfloat16 *__restrict__ y, int S, int total_tiles, float slope)
{
    using G = C0Geom<KS>;
    __shared__ __align__(16) __nv_bfloat16 patch[kC0Cin * G::PH * G::PPitch];
    // B fragments (weights) of all k-steps and n-tiles, fragment-ready: one conflict-free 8-byte load per MMA.  They live
    // in shared memory, not registers (a register-resident version needed 255 registers = one CTA per SM, and the kernel
    // is latency-bound: patch loads, shared-memory gathers)
    __shared__ __align__(8) uint2 bws[G::KSteps * 8 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int S2 = S / 2, tiles_x = S2 / kC0TW, tiles_img = tiles_x * (S2 / kC0TH);
    for (int i = threadIdx.x; i < G::KSteps * 8 * 32; i += kC0Threads) {      // B[k][n] = w[n][c][ky][kx]
        const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8, gg = ln >> 2, qq = ln & 3;
        uint32_t frag[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = ks * 16 + 2 * qq + 8 * h, n = nt * 8 + gg;
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = k + e, kx = kk % G::KX, r = kk / G::KX, ky = r % KS, ci = r / KS;
                v[e] = (kk < G::KReal && kx < KS) ? __ldg(w + ((size_t)(n * kC0Cin + ci) * KS + ky) * KS + kx) : 0.f;
            }
            frag[h] = pack_bf16x2(v[0], v[1]);
        }
        bws[i] = make_uint2(frag[0], frag[1]);
    }
    int aoff[G::KSteps][2];
#pragma unroll
    for (int ks = 0; ks < G::KSteps; ++ks) {
        aoff[ks][0] = c0_col_offset<KS>(ks * 16 + 2 * q);
        aoff[ks][1] = c0_col_offset<KS>(ks * 16 + 2 * q + 8);
    }
    float bv[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        bv[nt][0] = bias ? __ldg(bias + nt * 8 + 2 * q) : 0.f;
        bv[nt][1] = bias ? __ldg(bias + nt * 8 + 2 * q + 1) : 0.f;
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, t = tile - b * tiles_img;
        const int oy0 = (t / tiles_x) * kC0TH, ox0 = (t % tiles_x) * kC0TW;
        __syncthreads();
        c0_stage_patch<KS, PAD>(patch, img, (size_t)b * kC0Cin * S * S, S, oy0, ox0);
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
        const unsigned char *prow = reinterpret_cast<const unsigned char *>(patch + (2 * warp) * G::PPitch);
#pragma unroll
        for (int ks = 0; ks < G::KSteps; ++ks) {
            uint32_t a[4];
            a[0] = lds32(prow + (aoff[ks][0] + 2 * g) * 2);
            a[1] = lds32(prow + (aoff[ks][0] + 2 * (g + 8)) * 2);
            a[2] = lds32(prow + (aoff[ks][1] + 2 * g) * 2);
            a[3] = lds32(prow + (aoff[ks][1] + 2 * (g + 8)) * 2);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint2 bf = bws[(ks * 8 + nt) * 32 + lane];
                mma_bf16_16816(acc[nt], a, bf.x, bf.y);
            }
        }
        const int oy = oy0 + warp;
        __nv_bfloat16 *yb = y + (size_t)b * S2 * S2 * kC0Cout;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ox = ox0 + g + 8 * h;
            __nv_bfloat16 *dst = yb + (size_t)c0_map_row<HEAD>(oy, ox, S2) * kC0Cout + 2 * q;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float v0 = acc[nt][2 * h] + bv[nt][0], v1 = acc[nt][2 * h + 1] + bv[nt][1];
                v0 = v0 > 0.f ? v0 : v0 * slope;
                v1 = v1 > 0.f ? v1 : v1 * slope;
                *reinterpret_cast<uint32_t *>(dst + nt * 8) = pack_bf16x2(v0, v1);
            }
        }
    }
}

// The 64-channel map operand of the other two kernels: conv0 -> dpre = dy * (y > 0 ? 1 : slope) (gradient through the
// LeakyReLU, both tensors in the s2d activation layout); head -> the activation x itself (gate == null).
struct C0Map {
    const __nv_bfloat16 *v;       // dy, or x
    const __nv_bfloat16 *gate;    // y (LeakyReLU output), or null
    float slope;
    // 8 consecutive channels of one map row as fp32
    __device__ __forceinline__ void load8(size_t off, float (&f)[8]) const
    {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(v + off));
        const __nv_bfloat16 *h = reinterpret_cast<const __nv_bfloat16 *>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(h[j]);
        if (gate) {
            const uint4 yraw = __ldg(reinterpret_cast<const uint4 *>(gate + off));
            const __nv_bfloat16 *yh = reinterpret_cast<const __nv_bfloat16 *>(&yraw);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (!(__bfloat162float(yh[j]) > 0.f)) f[j] *= slope;
        }
    }
};

// ---- weight (/ bias) gradient -----------------------------------------------------------------------
// dW[n][k] = sum_pix map[pix][n] * im2col[pix][k]  as  M = n (64), N = k (KPad), K = pixels.  A CTA walks tiles, its 8
// warps = 4 row groups x 2 channel halves hold a 32 x KPad accumulator each; the row groups are summed through shared
// memory and every CTA writes one partial [64][KPad]; c0_dw_reduce_kernel sums the CTA partials in order.
constexpr int kC0DtPitch = kC0TH * kC0TW + 8;           // 136 bf16 per channel row of the transposed map tile

template <int KS, int PAD, bool HEAD>
__global__ void __launch_bounds__(kC0Threads) c0_dw_kernel(C0Image img, C0Map map, float *__restrict__ part, int S, int total_tiles)
{
    using G = C0Geom<KS>;
    constexpr int NT = G::KPad / 8;
    __shared__ __align__(16) __nv_bfloat16 patch[kC0Cin * G::PH * G::PPitch];
    __shared__ __align__(16) __nv_bfloat16 dt[kC0Cout * kC0DtPitch];             // map tile transposed: [channel][pixel]
    __shared__ float red[kC0Cout * (G::KPad + 1)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int rg = warp & 3, ch = warp >> 2;
    const int S2 = S / 2, tiles_x = S2 / kC0TW, tiles_img = tiles_x * (S2 / kC0TH);
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
    int boff[NT];                                       // patch offset of im2col column nt * 8 + g
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) boff[nt] = c0_col_offset<KS>(nt * 8 + g);
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, t = tile - b * tiles_img;
        const int oy0 = (t / tiles_x) * kC0TH, ox0 = (t % tiles_x) * kC0TW;
        __syncthreads();
        c0_stage_patch<KS, PAD>(patch, img, (size_t)b * kC0Cin * S * S, S, oy0, ox0);
        {   // map tile, transposed: thread -> (pixel, 8 channels)
            const size_t mimg = (size_t)b * S2 * S2 * kC0Cout;
            constexpr int NV = kC0TH * kC0TW * 8 / kC0Threads;      // 4 vectors per thread: loaded together, then scattered
            float f[NV][8];
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int i = threadIdx.x + u * kC0Threads;
                const int c8 = i & 7, pix = i >> 3, py = pix / kC0TW, px = pix % kC0TW;
                map.load8(mimg + (size_t)c0_map_row<HEAD>(oy0 + py, ox0 + px, S2) * kC0Cout + c8 * 8, f[u]);
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int i = threadIdx.x + u * kC0Threads;
                const int c8 = i & 7, pix = i >> 3;
#pragma unroll
                for (int j = 0; j < 8; ++j) dt[(c8 * 8 + j) * kC0DtPitch + pix] = __float2bfloat16_rn(f[u][j]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int row = rg + 4 * rr;                // tile row = one 16-pixel K step
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const unsigned char *base = reinterpret_cast<const unsigned char *>(dt + (ch * 32 + mt * 16 + g) * kC0DtPitch + row * kC0TW + 2 * q);
                a[mt][0] = lds32(base);
                a[mt][1] = lds32(base + 8 * kC0DtPitch * 2);
                a[mt][2] = lds32(base + 16);
                a[mt][3] = lds32(base + 8 * kC0DtPitch * 2 + 16);
            }
            const __nv_bfloat16 *prow = patch + (2 * row) * G::PPitch;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                // B[pix][kcol]: pixels 2q, 2q+1 (b0) and 2q+8, 2q+9 (b1) of im2col column kcol = nt * 8 + g
                const int kcol = nt * 8 + g;
                __nv_bfloat16 e[4];
                if (kcol < G::KReal) {
                    const __nv_bfloat16 *src = prow + boff[nt];
                    e[0] = src[2 * (2 * q)]; e[1] = src[2 * (2 * q + 1)]; e[2] = src[2 * (2 * q + 8)]; e[3] = src[2 * (2 * q + 9)];
                } else {
                    e[0] = e[1] = e[2] = e[3] = (!HEAD && kcol == kC0OnesCol) ? one : zero;
                }
                const uint32_t b0 = (uint32_t)__bfloat16_as_ushort(e[0]) | ((uint32_t)__bfloat16_as_ushort(e[1]) << 16);
                const uint32_t b1 = (uint32_t)__bfloat16_as_ushort(e[2]) | ((uint32_t)__bfloat16_as_ushort(e[3]) << 16);
                mma_bf16_16816(acc[0][nt], a[0], b0, b1);
                mma_bf16_16816(acc[1][nt], a[1], b0, b1);
            }
        }
    }
    // sum the four row groups of each channel half in a fixed order, then one partial per CTA
    for (int r = 0; r < 4; ++r) {
        __syncthreads();
        if (rg == r) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int co = ch * 32 + mt * 16 + g + 8 * (j >> 1), k = nt * 8 + 2 * q + (j & 1);
                        float *d = red + co * (G::KPad + 1) + k;
                        *d = r == 0 ? acc[mt][nt][j] : *d + acc[mt][nt][j];
                    }
        }
    }
    __syncthreads();
    float *dst = part + (size_t)blockIdx.x * kC0Cout * G::KPad;
    for (int i = threadIdx.x; i < kC0Cout * G::KPad; i += kC0Threads) dst[i] = red[(i / G::KPad) * (G::KPad + 1) + i % G::KPad];
}

// dw (64, 3, KS, KS) (and, conv0 only, dbias (64) from the ones column) from the CTA partials; accumulate != 0 adds
template <int KS, bool HEAD>
__global__ void __launch_bounds__(256) c0_dw_reduce_kernel(const float *__restrict__ part, int nparts, float *__restrict__ dw,
                                                           float *__restrict__ dbias, int accumulate)
{
    using G = C0Geom<KS>;
    const int i = blockIdx.x * 256 + threadIdx.x;       // (n, k) with k < KPad
    if (i >= kC0Cout * G::KPad) return;
    const int co = i / G::KPad, k = i % G::KPad;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * kC0Cout * G::KPad + i];
    if (!HEAD && k == kC0OnesCol) {
        if (dbias) dbias[co] = accumulate ? dbias[co] + s : s;
        return;
    }
    const int kx = k % G::KX, r = k / G::KX, ky = r % KS, ci = r / KS;
    if (k >= G::KReal || kx >= KS) return;
    float *d = dw + ((size_t)(co * kC0Cin + ci) * KS + ky) * KS + kx;
    *d = accumulate ? *d + s : s;
}

// head only: dbias[c] = sum over the batch and all pixels of g = dout * (1 - out^2).  Two stages, fixed order:
// grid (3, kC0BiasChunks) partial sums, then one warp per channel.
constexpr int kC0BiasChunks = 32;
__global__ void __launch_bounds__(256) c0_head_dbias_kernel(C0Image img, float *__restrict__ part, int batch, int S)
{
    __shared__ float sh[8];
    const int c = blockIdx.x, chunk = blockIdx.y;
    const size_t plane = (size_t)S * S, total = (size_t)batch * plane;
    const size_t per = (total + kC0BiasChunks - 1) / kC0BiasChunks, i0 = chunk * per, i1 = i0 + per < total ? i0 + per : total;
    float s = 0.f;
#pragma unroll 4
    for (size_t i = i0 + threadIdx.x; i < i1; i += 256) {
        const size_t b = i / plane, r = i - b * plane;
        s += img.at((b * kC0Cin + c) * plane + r);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sh[k];
        part[c * kC0BiasChunks + chunk] = t;
    }
}
__global__ void __launch_bounds__(96) c0_head_dbias_final_kernel(const float *__restrict__ part, float *__restrict__ dbias)
{
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float s = warp_sum(part[c * kC0BiasChunks + lane]);
    if (lane == 0) dbias[c] = s;
}

// ---- map -> image ---------------------------------------------------------------------------------------
// img[b][c][2i + p_y][2j + p_x] = sum_{n, taps of the parity class} map[b][i + d_y][j + d_x][n] * w[n][c][ky][kx], with
// p = (k + PAD) & 1 and d = (p + PAD - k) / 2 per dimension: per class a GEMM  M = pixels (i, j), K = taps x 64, N = 3 -> 8.
// A CTA owns an 8 x 16 block of (i, j): the map tile with a one-pixel halo sits in shared memory (pixel pitch 144 B:
// conflict-free fragment loads), the weights as ready-made B fragments; the 16 x 32 x 3 result goes through shared memory
// so that global rows are written contiguously.  conv0 dx: raw; head forward: tanh(. + bias).
constexpr int kC0XPitch = kC0Cout * 2 + 16;             // bytes per staged map pixel
constexpr int kC0HaloW = kC0TW + 2, kC0HaloH = kC0TH + 2;
template <int KS> constexpr size_t c0_map2img_smem()
{
    return (size_t)kC0HaloH * kC0HaloW * kC0XPitch + (size_t)KS * KS * 4 * 32 * sizeof(uint2) +
           (size_t)kC0Cin * (2 * kC0TH) * (2 * kC0TW + 1) * sizeof(float);
}

template <int KS, int PAD, bool HEAD>
__global__ void __launch_bounds__(kC0Threads) c0_map2img_kernel(C0Map map, const float *__restrict__ w, const float *__restrict__ bias,
                                                                float *__restrict__ out, int S, int total_tiles)
{
    extern __shared__ __align__(16) unsigned char c0x_smem[];
    unsigned char *ds = c0x_smem;                                                              // [10 x 18 pixels][144 B]
    uint2 *wfrag = reinterpret_cast<uint2 *>(ds + kC0HaloH * kC0HaloW * kC0XPitch);            // [taps][4 k-steps][32 lanes]
    float *outs = reinterpret_cast<float *>(wfrag + KS * KS * 4 * 32);                         // [3][16][33]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int S2 = S / 2, tiles_x = S2 / kC0TW, tiles_img = tiles_x * (S2 / kC0TH);
    for (int i = threadIdx.x; i < KS * KS * 4 * 32; i += kC0Threads) {
        const int ln = i & 31, ks = (i >> 5) & 3, tap = i >> 7, gg = ln >> 2, qq = ln & 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gg < kC0Cin) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = ks * 16 + 2 * qq + (e & 1) + 8 * (e >> 1);
                v[e] = __ldg(w + ((size_t)n * kC0Cin + gg) * (KS * KS) + tap);
            }
        }
        wfrag[i] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, t = tile - b * tiles_img;
        const int i0 = (t / tiles_x) * kC0TH, j0 = (t % tiles_x) * kC0TW;
        __syncthreads();
        {
            const size_t mimg = (size_t)b * S2 * S2 * kC0Cout;
            constexpr int NH = kC0HaloH * kC0HaloW * 8, U = 3;      // 1440 vectors: two batches of 3 per thread, loads together
            for (int i0v = threadIdx.x; i0v < NH; i0v += U * kC0Threads) {
                float f[U][8];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0v + u * kC0Threads;
                    const int c8 = i & 7, pix = i >> 3, hy = pix / kC0HaloW, hx = pix % kC0HaloW;
                    const int oy = i0 + hy - 1, ox = j0 + hx - 1;
                    ok[u] = i < NH && oy >= 0 && oy < S2 && ox >= 0 && ox < S2;
                    map.load8(ok[u] ? mimg + (size_t)c0_map_row<HEAD>(oy, ox, S2) * kC0Cout + c8 * 8 : mimg, f[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0v + u * kC0Threads;
                    const int c8 = i & 7, pix = i >> 3;
                    const uint4 o = ok[u] ? make_uint4(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]),
                                                       pack_bf16x2(f[u][4], f[u][5]), pack_bf16x2(f[u][6], f[u][7]))
                                          : make_uint4(0, 0, 0, 0);
                    if (i < NH) *reinterpret_cast<uint4 *>(ds + (size_t)pix * kC0XPitch + c8 * 16) = o;
                }
            }
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                constexpr int dummy = 0; (void)dummy;
                const int py = (ky + PAD) & 1, px = (kx + PAD) & 1;
                const int cls = py * 2 + px, dyy = (py + PAD - ky) / 2, dxx = (px + PAD - kx) / 2;
                // tile-local halo coordinates of pixel (i = i0 + warp, j = j0 + g [+ 8]) shifted by (dyy, dxx)
                const unsigned char *p0 = ds + (size_t)((warp + 1 + dyy) * kC0HaloW + (g + 1 + dxx)) * kC0XPitch + 4 * q;
                const unsigned char *p1 = p0 + 8 * kC0XPitch;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a[4];
                    a[0] = lds32(p0 + ks * 32); a[1] = lds32(p1 + ks * 32);
                    a[2] = lds32(p0 + ks * 32 + 16); a[3] = lds32(p1 + ks * 32 + 16);
                    const uint2 bf = wfrag[((ky * KS + kx) * 4 + ks) * 32 + lane];
                    mma_bf16_16816(acc[cls], a, bf.x, bf.y);
                }
            }
        // C[pix g (+8)][c = 2q, 2q+1] -> outs[c][row 2*warp + py][col 2*(g [+8]) + px]
        constexpr int OW = 2 * kC0TW + 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int py = c >> 1, px = c & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = 2 * q + (j & 1), col = 2 * (g + 8 * (j >> 1)) + px;
                if (ci < kC0Cin) outs[(ci * 2 * kC0TH + 2 * warp + py) * OW + col] = acc[c][j];
            }
        }
        __syncthreads();
        float *ob = out + (size_t)b * kC0Cin * S * S;
        for (int i = threadIdx.x; i < kC0Cin * 2 * kC0TH * 2 * kC0TW; i += kC0Threads) {
            const int col = i % (2 * kC0TW), r = i / (2 * kC0TW), row = r % (2 * kC0TH), ci = r / (2 * kC0TH);
            float v = outs[(ci * 2 * kC0TH + row) * OW + col];
            if (HEAD) v = tanhf(v + __ldg(bias + ci));
            ob[((size_t)ci * S + 2 * i0 + row) * S + 2 * j0 + col] = v;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// heads
// -------------------------------------------------------------------------------------------------
constexpr int kHdThreads = 256;
constexpr int kHdSlab = 64;                             // features per CTA
constexpr int kHdNPad = 160;                            // 129 output columns (128 of linear2 + the logit) padded to 5 x 32
constexpr int kHdN = 129;
constexpr int kHdLd = 132;                              // row pitch of the (B, 129) partial / gradient matrices
constexpr int kHdMaxB = 64;

// feature slab `slab`: channels [c0, c0 + cs) x all HW positions; local index k = c_local * HW + hw is the torch feature
// c0 * HW + k (contiguous in the weight rows) and sits at h[b][hw * C + c0 + c_local] in the channels-last activation
struct HeadGeom { int C, HW, cs, slabs, F; };

// All staging loops issue their global loads in register batches (unrolled, independent) before touching shared memory:
// a first version with one dependent load per trip was latency-bound (28 us for 58 KB per CTA).
__device__ __forceinline__ void hd_stage_h(float *hs, const __nv_bfloat16 *__restrict__ h, const HeadGeom &g, int batch, int c0)
{
    // item i = b * 64 + hw * cs + cl (cl fastest: cs contiguous channels in memory); shared index k = cl * HW + hw
    constexpr int kPer = kHdMaxB * kHdSlab / kHdThreads;       // 16
    __nv_bfloat16 v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = threadIdx.x + j * kHdThreads, b = i / kHdSlab, r = i % kHdSlab, hw = r / g.cs, cl = r - hw * g.cs;
        v[j] = b < batch ? h[(size_t)b * g.F + (size_t)hw * g.C + c0 + cl] : __float2bfloat16_rn(0.f);
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = threadIdx.x + j * kHdThreads, b = i / kHdSlab, r = i % kHdSlab, hw = r / g.cs, cl = r - hw * g.cs;
        hs[b * kHdSlab + cl * g.HW + hw] = __bfloat162float(v[j]);
    }
}
// ws[n][k] (pitch kHdSlab + 1): rows 0..127 = linear2, row 128 = linear1, rows 129.. zero
__device__ __forceinline__ void hd_stage_w(float *ws, const float *__restrict__ w1, const float *__restrict__ w2, const HeadGeom &g,
                                           int f0, int rows)
{
    constexpr int kVecRow = kHdSlab / 4, kVecs = kHdN * kVecRow, kPer = (kVecs + kHdThreads - 1) / kHdThreads;      // 16, 2064, 9
    float4 v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = threadIdx.x + j * kHdThreads, n = i / kVecRow, q = i % kVecRow;
        if (i < kVecs) v[j] = __ldg(reinterpret_cast<const float4 *>((n < 128 ? w2 + (size_t)n * g.F : w1) + f0) + q);
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = threadIdx.x + j * kHdThreads, n = i / kVecRow, q = i % kVecRow;
        if (i < kVecs) {
            float *d = ws + n * (kHdSlab + 1) + q * 4;
            d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
        }
    }
    for (int i = kHdN * (kHdSlab + 1) + threadIdx.x; i < rows * (kHdSlab + 1); i += kHdThreads) ws[i] = 0.f;
}

// partial[slab][b][n] = sum_{k in slab} h[b][k] * W[n][k]
__global__ void __launch_bounds__(kHdThreads) dheads_fwd_partial_kernel(const __nv_bfloat16 *__restrict__ h, const float *__restrict__ w1,
                                                                        const float *__restrict__ w2, float *__restrict__ part,
                                                                        HeadGeom g, int batch)
{
    extern __shared__ __align__(16) float hd_smem[];
    float *hs = hd_smem, *ws = hd_smem + kHdMaxB * kHdSlab;
    const int slab = blockIdx.x, c0 = slab * g.cs, f0 = c0 * g.HW;
    hd_stage_h(hs, h, g, batch, c0);
    hd_stage_w(ws, w1, w2, g, f0, kHdNPad);
    __syncthreads();
    const int tn = threadIdx.x & 31, tb = threadIdx.x >> 5;
    float acc[8][5];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < kHdSlab; ++k) {
        float hv[8], wv[5];
#pragma unroll
        for (int i = 0; i < 8; ++i) hv[i] = hs[(tb * 8 + i) * kHdSlab + k];
#pragma unroll
        for (int j = 0; j < 5; ++j) wv[j] = ws[(tn + 32 * j) * (kHdSlab + 1) + k];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(hv[i], wv[j], acc[i][j]);
    }
    float *dst = part + (size_t)slab * kHdMaxB * kHdLd;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int n = tn + 32 * j;
            if (n < kHdN) dst[(tb * 8 + i) * kHdLd + n] = acc[i][j];
        }
}

// one CTA per sample: four thread groups split the slabs of every column (8 independent loads in flight each, the
// groups meet in shared memory in a fixed order), bias / LeakyReLU, then linear3 + tanh with one warp per output row
// (coalesced weight rows, shuffle reduction)
constexpr int kHdFinThreads = 512;
__global__ void __launch_bounds__(kHdFinThreads) dheads_fwd_finish_kernel(const float *__restrict__ part, int slabs,
                                                                         const float *__restrict__ b1, const float *__restrict__ b2,
                                                                         const float *__restrict__ w3, const float *__restrict__ b3,
                                                                         float *__restrict__ logits, float *__restrict__ t2,
                                                                         float *__restrict__ zp, int zdim, float slope)
{
    __shared__ float red[4][132];
    __shared__ float ts[128];
    const int b = blockIdx.x, n = threadIdx.x & 127, grp = threadIdx.x >> 7, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr size_t kSlabStride = (size_t)kHdMaxB * kHdLd;
    const float *col = part + (size_t)b * kHdLd + n;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    for (int sl = grp; sl < slabs; sl += 32) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (sl + 4 * u < slabs) acc[u] += col[(size_t)(sl + 4 * u) * kSlabStride];
    }
    red[grp][n] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    if (warp == 15) {                                   // the logit column: lanes stride over the slabs
        float l = 0.f;
        for (int sl = lane; sl < slabs; sl += 32) l += part[(size_t)sl * kSlabStride + (size_t)b * kHdLd + 128];
        l = warp_sum(l);
        if (lane == 0) logits[b] = l + b1[0];
    }
    __syncthreads();
    if (grp == 0) {
        float s = ((red[0][n] + red[1][n]) + (red[2][n] + red[3][n])) + b2[n];
        s = s > 0.f ? s : s * slope;
        ts[n] = s;
        t2[(size_t)b * 128 + n] = s;
    }
    __syncthreads();
    const float t0 = ts[lane], t1 = ts[lane + 32], t2v = ts[lane + 64], t3 = ts[lane + 96];
#pragma unroll 4
    for (int j = warp; j < zdim; j += kHdFinThreads / 32) {
        const float *wr = w3 + (size_t)j * 128 + lane;
        float a = __ldg(wr) * t0 + __ldg(wr + 32) * t1 + __ldg(wr + 64) * t2v + __ldg(wr + 96) * t3;
        a = warp_sum(a);
        if (lane == 0) zp[(size_t)b * zdim + j] = tanhf(a + b3[j]);
    }
}

// backward, per sample: dt3 = dzp * (1 - zp^2);  dO[b][n] = (sum_j dt3[j] * W3[j][n]) * lrelu'(t2[n]);  dO[b][128] = dlogits[b]
__global__ void __launch_bounds__(128) dheads_bwd_small_kernel(const float *__restrict__ dlogits, const float *__restrict__ dzp,
                                                               const float *__restrict__ zp, const float *__restrict__ t2,
                                                               const float *__restrict__ w3, float *__restrict__ dt3,
                                                               float *__restrict__ dO, int zdim, float slope)
{
    extern __shared__ float d3[];
    const int b = blockIdx.x, n = threadIdx.x;
    for (int j = n; j < zdim; j += 128) {
        const float z = zp[(size_t)b * zdim + j];
        const float v = (dzp ? dzp[(size_t)b * zdim + j] : 0.f) * (1.f - z * z);
        d3[j] = v;
        dt3[(size_t)b * zdim + j] = v;
    }
    __syncthreads();
    float s = 0.f;
#pragma unroll 8
    for (int j = 0; j < zdim; ++j) s = fmaf(d3[j], __ldg(w3 + (size_t)j * 128 + n), s);
    dO[(size_t)b * kHdLd + n] = t2[(size_t)b * 128 + n] > 0.f ? s : s * slope;
    if (n == 0) dO[(size_t)b * kHdLd + 128] = dlogits ? dlogits[b] : 0.f;
}

// small parameter gradients: dW3[j][n] = sum_b dt3[b][j] * t2[b][n], db3[j], db2[n], db1
__global__ void __launch_bounds__(128) dheads_bwd_params_kernel(const float *__restrict__ dt3, const float *__restrict__ t2,
                                                                const float *__restrict__ dO, float *__restrict__ dw3,
                                                                float *__restrict__ db3, float *__restrict__ db2, float *__restrict__ db1,
                                                                int batch, int zdim)
{
    const int j = blockIdx.x, n = threadIdx.x;
    if (j < zdim) {
        float s = 0.f, sb = 0.f;
        for (int b = 0; b < batch; ++b) {
            const float d = dt3[(size_t)b * zdim + j];
            s = fmaf(d, t2[(size_t)b * 128 + n], s);
            sb += d;
        }
        dw3[(size_t)j * 128 + n] = s;
        if (n == 0) db3[j] = sb;
    } else {                                            // one extra block: biases of linear2 / linear1
        float s = 0.f, l = 0.f;
        for (int b = 0; b < batch; ++b) {
            s += dO[(size_t)b * kHdLd + n];
            if (n == 0) l += dO[(size_t)b * kHdLd + 128];
        }
        db2[n] = s;
        if (n == 0) db1[0] = l;
    }
}

// big gradients per feature slab: dW[n][k] = sum_b dO[b][n] * h[b][k] (torch layout rows of W1 / W2, optional) and
// dh[b][k] = sum_n dO[b][n] * W[n][k] (bf16, channels-last activation layout)
__global__ void __launch_bounds__(kHdThreads) dheads_bwd_big_kernel(const __nv_bfloat16 *__restrict__ h, const float *__restrict__ w1,
                                                                    const float *__restrict__ w2, const float *__restrict__ dO,
                                                                    float *__restrict__ dw1, float *__restrict__ dw2,
                                                                    __nv_bfloat16 *__restrict__ dh, HeadGeom g, int batch)
{
    extern __shared__ __align__(16) float hd_smem[];
    float *hs = hd_smem, *ws = hs + kHdMaxB * kHdSlab, *os = ws + kHdNPad * (kHdSlab + 1);     // os[b][n], pitch kHdNPad
    const int slab = blockIdx.x, c0 = slab * g.cs, f0 = c0 * g.HW;
    hd_stage_h(hs, h, g, batch, c0);
    hd_stage_w(ws, w1, w2, g, f0, kHdNPad);
    {
        constexpr int kPer = kHdMaxB * kHdNPad / kHdThreads;        // 40 elements per thread, loaded in two batches of 20
        static_assert(kPer % 2 == 0, "batching");
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[kPer / 2];
#pragma unroll
            for (int j = 0; j < kPer / 2; ++j) {
                const int i = threadIdx.x + (half * (kPer / 2) + j) * kHdThreads, n = i % kHdNPad, b = i / kHdNPad;
                v[j] = (b < batch && n < kHdN) ? __ldg(dO + (size_t)b * kHdLd + n) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kPer / 2; ++j) os[threadIdx.x + (half * (kPer / 2) + j) * kHdThreads] = v[j];
        }
    }
    __syncthreads();
    // blockIdx.y selects the role when the launch has two (weight gradients and dh both wanted): halves the work per CTA
    const bool do_dw = dw2 != nullptr && (gridDim.y == 1 || blockIdx.y == 0);
    const bool do_dh = dh != nullptr && (gridDim.y == 1 || blockIdx.y == 1);
    if (do_dw) {
        const int tn = threadIdx.x & 31, tk = threadIdx.x >> 5;
        float acc[5][8];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
        for (int b = 0; b < batch; ++b) {
            float ov[5], hv[8];
#pragma unroll
            for (int j = 0; j < 5; ++j) ov[j] = os[b * kHdNPad + tn + 32 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[i] = hs[b * kHdSlab + tk * 8 + i];
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(ov[j], hv[i], acc[j][i]);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int n = tn + 32 * j;
            if (n < kHdN) {
                float *dst = (n < 128 ? dw2 + (size_t)n * g.F : dw1) + f0 + tk * 8;
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
            }
        }
    }
    if (do_dh) {
        const int k = threadIdx.x & 63, tb = threadIdx.x >> 6;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int n = 0; n < kHdN; ++n) {
            const float wv = ws[n * (kHdSlab + 1) + k];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(os[(tb * 16 + i) * kHdNPad + n], wv, acc[i]);
        }
        const int cl = k / g.HW, hw = k - cl * g.HW;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int b = tb * 16 + i;
            if (b < batch) dh[(size_t)b * g.F + (size_t)hw * g.C + c0 + cl] = __float2bfloat16_rn(acc[i]);
        }
    }
}

static int head_geom(const char *who, int batch, int channels, int hw, HeadGeom &g)
{
    HG_REQUIRE(batch > 0 && batch <= kHdMaxB, HG_ERR_UNSUPPORTED, "%s: batch must be 1..%d per call (got %d)", who, kHdMaxB, batch);
    HG_REQUIRE(channels > 0 && hw > 0 && (hw <= kHdSlab ? kHdSlab % hw == 0 : false), HG_ERR_UNSUPPORTED,
               "%s: H*W must divide %d (got %d)", who, kHdSlab, hw);
    g.C = channels; g.HW = hw; g.cs = kHdSlab / hw; g.F = channels * hw;
    HG_REQUIRE(channels % g.cs == 0, HG_ERR_UNSUPPORTED, "%s: channels %% %d != 0", who, g.cs);
    g.slabs = channels / g.cs;
    return HG_OK;
}

static int c0_geom(const char *who, int batch, int cin, int cout, int size, int &tiles, int &grid)
{
    HG_REQUIRE(batch > 0 && size > 0, HG_ERR_INVALID_ARG, "%s: dims must be positive", who);
    HG_REQUIRE(cin == kC0Cin && cout == kC0Cout, HG_ERR_UNSUPPORTED, "%s: only 3 <-> 64 channels are built (got %d, %d)", who, cin, cout);
    HG_REQUIRE(size % (2 * kC0TW) == 0, HG_ERR_UNSUPPORTED, "%s: image size must be a multiple of %d (got %d)", who, 2 * kC0TW, size);
    const int S2 = size / 2;
    tiles = batch * (S2 / kC0TH) * (S2 / kC0TW);
    const int per_cta = (tiles + 3 * sm_count() - 1) / (3 * sm_count());     // balanced: every CTA walks the same number of tiles
    grid = (tiles + per_cta - 1) / per_cta;
    return HG_OK;
}

template <int KS, int PAD, bool HEAD>
static int c0_launch_dw(const char *who, const C0Image &img, const C0Map &map, float *dw, float *dbias, void *workspace,
                        long long workspace_bytes, int size, int tiles, int grid, int accumulate, cudaStream_t st)
{
    using G = C0Geom<KS>;
    const int g = grid < sm_count() ? grid : sm_count();          // one CTA per SM: each writes one partial
    HG_REQUIRE(workspace && workspace_bytes >= (long long)g * kC0Cout * G::KPad * (long long)sizeof(float), HG_ERR_INVALID_ARG,
               "%s: workspace too small", who);
    float *part = static_cast<float *>(workspace);
    c0_dw_kernel<KS, PAD, HEAD><<<g, kC0Threads, 0, st>>>(img, map, part, size, tiles);
    int rc = check_launch(who);
    if (rc) return rc;
    c0_dw_reduce_kernel<KS, HEAD><<<(kC0Cout * G::KPad + 255) / 256, 256, 0, st>>>(part, g, dw, dbias, accumulate);
    return check_launch(who);
}

template <int KS, int PAD, bool HEAD>
static int c0_launch_map2img(const char *who, const C0Map &map, const float *w, const float *bias, float *out, int size, int tiles,
                             int grid, cudaStream_t st)
{
    constexpr size_t smem = c0_map2img_smem<KS>();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(c0_map2img_kernel<KS, PAD, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    c0_map2img_kernel<KS, PAD, HEAD><<<grid, kC0Threads, smem, st>>>(map, w, bias, out, size, tiles);
    return check_launch(who);
}

}  // namespace hg

using namespace hg;

extern "C" int hg_dconv0_fwd(const float *x, const float *w, const float *bias, void *y_s2d, int batch, int cin, int cout, int size,
                             float neg_slope, void *stream)
{
    HG_REQUIRE(x && w && bias && y_s2d, HG_ERR_INVALID_ARG, "hg_dconv0_fwd: null pointer");
    int tiles, grid;
    int rc = c0_geom("hg_dconv0_fwd", batch, cin, cout, size, tiles, grid);
    if (rc) return rc;
    c0_img2map_kernel<5, 2, false><<<grid, kC0Threads, 0, static_cast<cudaStream_t>(stream)>>>(C0Image{x, nullptr}, w, bias,
                                                                                               static_cast<__nv_bfloat16 *>(y_s2d), size, tiles,
                                                                                               neg_slope);
    return check_launch("hg_dconv0_fwd");
}

extern "C" long long hg_dconv0_bwd_workspace_bytes(int batch, int size)
{
    if (batch <= 0 || size <= 0 || size % (2 * kC0TW)) return -1;
    return (long long)sm_count() * kC0Cout * C0Geom<5>::KPad * (long long)sizeof(float);
}

extern "C" int hg_dconv0_bwd(const float *x, const float *w, const void *y_s2d, const void *dy_s2d, float *dx, float *dw, float *dbias,
                             void *workspace, long long workspace_bytes, int batch, int cin, int cout, int size, float neg_slope,
                             int accumulate, void *stream)
{
    HG_REQUIRE(x && w && y_s2d && dy_s2d, HG_ERR_INVALID_ARG, "hg_dconv0_bwd: null pointer");
    HG_REQUIRE((dw == nullptr) == (dbias == nullptr), HG_ERR_INVALID_ARG, "hg_dconv0_bwd: dw and dbias come as a pair");
    int tiles, grid;
    int rc = c0_geom("hg_dconv0_bwd", batch, cin, cout, size, tiles, grid);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const C0Map map{static_cast<const __nv_bfloat16 *>(dy_s2d), static_cast<const __nv_bfloat16 *>(y_s2d), neg_slope};
    if (dw) {
        rc = c0_launch_dw<5, 2, false>("hg_dconv0_bwd(dw)", C0Image{x, nullptr}, map, dw, dbias, workspace, workspace_bytes, size, tiles,
                                       grid, accumulate, st);
        if (rc) return rc;
    }
    if (dx) rc = c0_launch_map2img<5, 2, false>("hg_dconv0_bwd(dx)", map, w, nullptr, dx, size, tiles, grid, st);
    return rc;
}

// The generator's patched 128 x 128 head: out = tanh(ConvTranspose2d(64 -> 3, k4, s2, p1)(x) + bias)
//   x (B, S/2, S/2, 64) bf16 NHWC, w (64, 3, 4, 4) fp32 torch layout, out (B, 3, S, S) fp32 NCHW, S = size (the OUTPUT extent)
extern "C" int hg_head128_fwd(const void *x, const float *w, const float *bias, float *out, int batch, int cin, int cout, int size,
                              void *stream)
{
    HG_REQUIRE(x && w && bias && out, HG_ERR_INVALID_ARG, "hg_head128_fwd: null pointer");
    int tiles, grid;
    int rc = c0_geom("hg_head128_fwd", batch, cout, cin, size, tiles, grid);
    if (rc) return rc;
    const C0Map map{static_cast<const __nv_bfloat16 *>(x), nullptr, 1.f};
    return c0_launch_map2img<4, 1, true>("hg_head128_fwd", map, w, bias, out, size, tiles, grid, static_cast<cudaStream_t>(stream));
}

extern "C" long long hg_head128_bwd_workspace_bytes(int batch, int size)
{
    if (batch <= 0 || size <= 0 || size % (2 * kC0TW)) return -1;
    return ((long long)sm_count() * kC0Cout * C0Geom<4>::KPad + kC0Cin * kC0BiasChunks) * (long long)sizeof(float);
}

// backward: g = dout * (1 - out^2); dx (B, S/2, S/2, 64) bf16 (may be null), dw (64, 3, 4, 4) + dbias (3) fp32 (both or neither)
extern "C" int hg_head128_bwd(const void *x, const float *w, const float *out, const float *dout, void *dx, float *dw, float *dbias,
                              void *workspace, long long workspace_bytes, int batch, int cin, int cout, int size, void *stream)
{
    HG_REQUIRE(x && w && out && dout, HG_ERR_INVALID_ARG, "hg_head128_bwd: null pointer");
    HG_REQUIRE((dw == nullptr) == (dbias == nullptr), HG_ERR_INVALID_ARG, "hg_head128_bwd: dw and dbias come as a pair");
    int tiles, grid;
    int rc = c0_geom("hg_head128_bwd", batch, cout, cin, size, tiles, grid);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const C0Image g{dout, out};
    if (dx) {
        c0_img2map_kernel<4, 1, true><<<grid, kC0Threads, 0, st>>>(g, w, nullptr, static_cast<__nv_bfloat16 *>(dx), size, tiles, 1.f);
        rc = check_launch("hg_head128_bwd(dx)");
        if (rc) return rc;
    }
    if (dw) {
        const C0Map map{static_cast<const __nv_bfloat16 *>(x), nullptr, 1.f};
        HG_REQUIRE(workspace_bytes >= hg_head128_bwd_workspace_bytes(batch, size), HG_ERR_INVALID_ARG,
                   "hg_head128_bwd: workspace smaller than hg_head128_bwd_workspace_bytes()");
        rc = c0_launch_dw<4, 1, true>("hg_head128_bwd(dw)", g, map, dw, nullptr, workspace, workspace_bytes, size, tiles, grid, 0, st);
        if (rc) return rc;
        float *bpart = static_cast<float *>(workspace) + (size_t)sm_count() * kC0Cout * C0Geom<4>::KPad;
        c0_head_dbias_kernel<<<dim3(kC0Cin, kC0BiasChunks), 256, 0, st>>>(g, bpart, batch, size);
        c0_head_dbias_final_kernel<<<1, 96, 0, st>>>(bpart, dbias);
        rc = check_launch("hg_head128_bwd(dbias)");
    }
    return rc;
}

static size_t hd_fwd_smem() { return (size_t)(kHdMaxB * kHdSlab + kHdNPad * (kHdSlab + 1)) * sizeof(float); }
static size_t hd_bwd_smem() { return hd_fwd_smem() + (size_t)kHdMaxB * kHdNPad * sizeof(float); }

extern "C" long long hg_dheads_workspace_bytes(int batch, int channels, int hw, int zdim)
{
    HeadGeom g;
    if (head_geom("hg_dheads_workspace_bytes", batch, channels, hw, g) || zdim <= 0) return -1;
    // forward: slab partials; backward: dt3 (B, zdim) + dO (B, 132)
    const long long fwd = (long long)g.slabs * kHdMaxB * kHdLd * 4, bwd = (long long)batch * (zdim + kHdLd) * 4;
    return fwd > bwd ? fwd : bwd;
}

extern "C" int hg_dheads_fwd(const void *h, const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                             const float *b3, float *logits, float *t2, float *z_pred, void *workspace, long long workspace_bytes,
                             int batch, int channels, int hw, int zdim, float neg_slope, void *stream)
{
    HG_REQUIRE(h && w1 && b1 && w2 && b2 && w3 && b3 && logits && t2 && z_pred && workspace, HG_ERR_INVALID_ARG, "hg_dheads_fwd: null pointer");
    HeadGeom g;
    int rc = head_geom("hg_dheads_fwd", batch, channels, hw, g);
    if (rc) return rc;
    HG_REQUIRE(zdim > 0 && workspace_bytes >= hg_dheads_workspace_bytes(batch, channels, hw, zdim), HG_ERR_INVALID_ARG,
               "hg_dheads_fwd: workspace smaller than hg_dheads_workspace_bytes()");
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(dheads_fwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hd_fwd_smem());
        cudaFuncSetAttribute(dheads_bwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hd_bwd_smem());
        attr = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *part = static_cast<float *>(workspace);
    dheads_fwd_partial_kernel<<<g.slabs, kHdThreads, hd_fwd_smem(), st>>>(static_cast<const __nv_bfloat16 *>(h), w1, w2, part, g, batch);
    rc = check_launch("hg_dheads_fwd(partial)");
    if (rc) return rc;
    dheads_fwd_finish_kernel<<<batch, kHdFinThreads, 0, st>>>(part, g.slabs, b1, b2, w3, b3, logits, t2, z_pred, zdim, neg_slope);
    return check_launch("hg_dheads_fwd(finish)");
}

extern "C" int hg_dheads_bwd(const void *h, const float *w1, const float *w2, const float *w3, const float *t2, const float *z_pred,
                             const float *dlogits, const float *dz_pred, void *dh, float *dw1, float *db1, float *dw2, float *db2,
                             float *dw3, float *db3, void *workspace, long long workspace_bytes, int batch, int channels, int hw,
                             int zdim, float neg_slope, void *stream)
{
    HG_REQUIRE(h && w1 && w2 && w3 && t2 && z_pred && workspace, HG_ERR_INVALID_ARG, "hg_dheads_bwd: null pointer");
    const bool params = dw1 != nullptr;
    HG_REQUIRE(params == (db1 != nullptr) && params == (dw2 != nullptr) && params == (db2 != nullptr) && params == (dw3 != nullptr) &&
                   params == (db3 != nullptr), HG_ERR_INVALID_ARG, "hg_dheads_bwd: the six parameter gradients come together or not at all");
    HeadGeom g;
    int rc = head_geom("hg_dheads_bwd", batch, channels, hw, g);
    if (rc) return rc;
    HG_REQUIRE(zdim > 0 && workspace_bytes >= hg_dheads_workspace_bytes(batch, channels, hw, zdim), HG_ERR_INVALID_ARG,
               "hg_dheads_bwd: workspace smaller than hg_dheads_workspace_bytes()");
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(dheads_bwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hd_bwd_smem());
        attr = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *dt3 = static_cast<float *>(workspace), *dO = dt3 + (size_t)batch * zdim;
    dheads_bwd_small_kernel<<<batch, 128, zdim * sizeof(float), st>>>(dlogits, dz_pred, z_pred, t2, w3, dt3, dO, zdim, neg_slope);
    rc = check_launch("hg_dheads_bwd(small)");
    if (rc) return rc;
    if (params) {
        dheads_bwd_params_kernel<<<zdim + 1, 128, 0, st>>>(dt3, t2, dO, dw3, db3, db2, db1, batch, zdim);
        rc = check_launch("hg_dheads_bwd(params)");
        if (rc) return rc;
    }
    if (params || dh) {
        dheads_bwd_big_kernel<<<dim3(g.slabs, (params && dh) ? 2 : 1), kHdThreads, hd_bwd_smem(), st>>>(static_cast<const __nv_bfloat16 *>(h), w1, w2, dO, dw1, dw2,
                                                                          static_cast<__nv_bfloat16 *>(dh), g, batch);
        rc = check_launch("hg_dheads_bwd(big)");
    }
    return rc;
}
