// final_layer + tanh (reference core/models/hologan_generator.py:69-75,141-142) on the warp-level tensor-core MMA
// (mma.sync.m16n8k16, bf16 operands, fp32 accumulation) for the hot-path shape Cin = 64, Cout = 3, S % 32 == 0.
//
// Why: with 3 output channels the op is bandwidth-bound on paper (one read of the 32 MiB activation at B = 64 = 5 us),
// but as a direct convolution on the FP32 pipe it needs 0.45 G FMAs per pass (>= 12 us) and measured 51-57 us per pass
// (profiles/r01i_microbench_pipeline.txt).  The contraction is re-associated so that it is a small dense GEMM with NO
// padding waste, and the 3x3 stencil becomes a cheap gather on fp32 partial results:
//   forward : T[q][(tap, co)] = sum_ci x[q][ci] * w[co][ci][tap]          GEMM  M = pixels (with halo), K = 64, N = 27 -> 32
//             out[p][co]      = tanh(bias[co] + sum_tap T[p + shift(tap)][(tap, co)])      9-term gather from shared memory
//   dx      : G[p][(tap, co)] = g[co][p - shift(tap)]  (im2col of the 3-channel g = dout * (1 - out^2))
//             dx[p][ci]       = sum_(tap,co) G[p][(tap, co)] * w[co][ci][tap]   GEMM  M = pixels, K = 27 -> 32, N = 64
//   dw      : dw[(tap, co)][ci] = sum_p G[p][(tap, co)] * x[p][ci]              GEMM  M = 27 -> 32, K = pixels, N = 64
// fp32 quantities that enter an MMA as bf16 (the weights, g) are split into bf16 terms (hi + lo = 16 mantissa bits; the
// forward's weights hi + mid + lo = 24 bits, exact), one MMA per term into the same fp32 accumulator: the results stay
// within 1e-5 / 2e-5 of the fp32 reference (tests/test_gpu_small_ops.py); the activation x is bf16 already (exact).
// Fragment layouts (PTX ISA, mma.m16n8k16 .row.col): with g = lane / 4, q = lane % 4
//   A (16 x 16): a0 = A[g][2q, 2q+1]   a1 = A[g+8][2q, 2q+1]   a2 = A[g][2q+8, 2q+9]   a3 = A[g+8][2q+8, 2q+9]
//   B (16 x  8): b0 = B[2q, 2q+1][g]   b1 = B[2q+8, 2q+9][g]
//   C (16 x  8): c0, c1 = C[g][2q, 2q+1]   c2, c3 = C[g+8][2q, 2q+1]
// Operands live in shared memory with the K index contiguous, so every fragment register is one 32-bit load; row
// pitches are chosen = 4 (mod 32) words, which makes the 8 x 4 (g, q) pattern of a warp hit 32 distinct banks.
// Deterministic (fixed summation order), no atomics.  Selected by final_conv.cu (HG_FINAL_CONV_MMA).
#include "hg_common.cuh"
#include "mma_sync.cuh"

namespace hg {

constexpr int kFmThreads = 256;
constexpr int kFmTH = 8, kFmTW = 32;                    // output tile of one CTA / one trip
constexpr int kFmC = 64, kFmCout = 3, kFmN = 32;        // N = 27 (tap, co) columns padded to 32
constexpr int kFmHaloPix = (kFmTH + 2) * (kFmTW + 2);   // 340
constexpr int kFmHaloPad = (kFmHaloPix + 15) / 16 * 16; // 352 = 22 m16 tiles
constexpr int kFmXPitch = kFmC * 2 + 16;                // bytes per staged pixel (36 words)
constexpr int kFmTPitch = 29;                           // floats per T row (odd: conflict-free gather)


__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo)
{
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}


// Stage the (TH+2) x (TW+2) halo tile of x (NHWC bf16, C = 64) as [pixel][C] with kFmXPitch bytes per pixel; pixels
// outside the image and the pad rows up to kFmHaloPad are zero.  Asynchronous copies (one commit group): all 11 copies of
// a thread are in flight together; the caller waits (cp_async_wait_group) and synchronises the CTA before reading.
__device__ __forceinline__ void fm_stage_x_halo_async(unsigned char *xs, const __nv_bfloat16 *__restrict__ xb, int S, int x0, int y0)
{
    constexpr int L = kFmC / 8;
#pragma unroll
    for (int i0 = 0; i0 < kFmHaloPad * L; i0 += kFmThreads) {
        const int i = i0 + threadIdx.x;
        const int v = i % L, pix = i / L;
        const int gx = pix % (kFmTW + 2), gy = pix / (kFmTW + 2);
        const int yy = y0 + gy - 1, xx = x0 + gx - 1;
        const bool ok = pix < kFmHaloPix && yy >= 0 && yy < S && xx >= 0 && xx < S;
        cp_async_16_zfill(xs + (size_t)pix * kFmXPitch + v * 16, ok ? xb + ((size_t)yy * S + xx) * kFmC + v * 8 : xb, ok);
    }
    cp_async_commit();
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
// shared memory: xs [352][144 B] | wh, wm, wl [32][144 B] (rows n = tap * 3 + co, K = ci contiguous) | T [352][29] fp32
// The forward keeps the fp32 tolerance of the SIMT kernel (1e-5 of max|out|): two bf16 terms per weight leave 1.2e-5,
// so the weights are split three ways (24 mantissa bits: exact); hi and mid fragments stay in registers, lo is re-read.
constexpr size_t kFmFwdSmem = (size_t)kFmHaloPad * kFmXPitch + 3 * (size_t)kFmN * kFmXPitch + (size_t)kFmHaloPad * kFmTPitch * 4;

__global__ void __launch_bounds__(kFmThreads, 2) final_conv_tanh_fwd_mma_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                                const float *__restrict__ w,
                                                                                const float *__restrict__ bias,
                                                                                float *__restrict__ out, int S, int tiles_x)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    unsigned char *xs = fm_smem;
    unsigned char *wh = xs + (size_t)kFmHaloPad * kFmXPitch;
    unsigned char *wm = wh + (size_t)kFmN * kFmXPitch;
    unsigned char *wl = wm + (size_t)kFmN * kFmXPitch;
    float *T = reinterpret_cast<float *>(wl + (size_t)kFmN * kFmXPitch);
    const int b = blockIdx.y, tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int x0 = tile_x * kFmTW, y0 = tile_y * kFmTH;

    fm_stage_x_halo_async(xs, x + (size_t)b * S * S * kFmC, S, x0, y0);      // in flight while the weights are prepared
    // weights: torch (co, ci, tap) fp32 -> rows n = tap * 3 + co of bf16 hi / mid / lo, rows 27..31 zero.  All 8 loads of a
    // thread are issued before the first is used.
    {
        float wv[kFmN * kFmC / kFmThreads];
#pragma unroll
        for (int k = 0; k < kFmN * kFmC / kFmThreads; ++k) {
            const int i = threadIdx.x + k * kFmThreads, ci = i % kFmC, n = i / kFmC;
            const int tap = n / kFmCout, co = n - tap * kFmCout;
            wv[k] = n < 9 * kFmCout ? __ldg(w + ((size_t)co * kFmC + ci) * 9 + tap) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kFmN * kFmC / kFmThreads; ++k) {
            const int i = threadIdx.x + k * kFmThreads, ci = i % kFmC, n = i / kFmC;
            __nv_bfloat16 hi, mid;
            split_bf16(wv[k], hi, mid);
            const __nv_bfloat16 lo = __float2bfloat16_rn((wv[k] - __bfloat162float(hi)) - __bfloat162float(mid));
            *reinterpret_cast<__nv_bfloat16 *>(wh + (size_t)n * kFmXPitch + ci * 2) = hi;
            *reinterpret_cast<__nv_bfloat16 *>(wm + (size_t)n * kFmXPitch + ci * 2) = mid;
            *reinterpret_cast<__nv_bfloat16 *>(wl + (size_t)n * kFmXPitch + ci * 2) = lo;
        }
    }
    cp_async_wait_group<0>();
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    // B fragments of all 4 n-tiles x 4 k-chunks, hi and mid: 64 registers, reused for every m-tile of the warp
    uint32_t bh[4][4][2], bm[4][4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            const size_t off = (size_t)(nt * 8 + g) * kFmXPitch + (kc * 16 + 2 * q) * 2;
            bh[nt][kc][0] = lds32(wh + off); bh[nt][kc][1] = lds32(wh + off + 16);
            bm[nt][kc][0] = lds32(wm + off); bm[nt][kc][1] = lds32(wm + off + 16);
        }
    for (int mt = warp; mt < kFmHaloPad / 16; mt += kFmThreads / 32) {
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
        const unsigned char *arow = xs + (size_t)(mt * 16 + g) * kFmXPitch + 2 * q * 2;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t a[4];
            a[0] = lds32(arow + kc * 32);
            a[1] = lds32(arow + 8 * kFmXPitch + kc * 32);
            a[2] = lds32(arow + kc * 32 + 16);
            a[3] = lds32(arow + 8 * kFmXPitch + kc * 32 + 16);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const size_t off = (size_t)(nt * 8 + g) * kFmXPitch + (kc * 16 + 2 * q) * 2;
                mma_bf16_16816(acc[nt], a, lds32(wl + off), lds32(wl + off + 16));      // smallest terms first
                mma_bf16_16816(acc[nt], a, bm[nt][kc][0], bm[nt][kc][1]);
                mma_bf16_16816(acc[nt], a, bh[nt][kc][0], bh[nt][kc][1]);
            }
        }
        float *t0 = T + (size_t)(mt * 16 + g) * kFmTPitch, *t1 = t0 + 8 * kFmTPitch;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int col = nt * 8 + 2 * q;
            if (col < 9 * kFmCout) { t0[col] = acc[nt][0]; t1[col] = acc[nt][2]; }
            if (col + 1 < 9 * kFmCout) { t0[col + 1] = acc[nt][1]; t1[col + 1] = acc[nt][3]; }
        }
    }
    __syncthreads();
    // stencil gather: thread = output pixel (ly, lx); input halo pixel of tap (ty, tx) is (ly + ty, lx + tx)
    const int lx = threadIdx.x % kFmTW, ly = threadIdx.x / kFmTW;
    float sum[kFmCout];
#pragma unroll
    for (int co = 0; co < kFmCout; ++co) sum[co] = bias[co];
#pragma unroll
    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
            const float *tp = T + (size_t)((ly + ty) * (kFmTW + 2) + lx + tx) * kFmTPitch + (ty * 3 + tx) * kFmCout;
#pragma unroll
            for (int co = 0; co < kFmCout; ++co) sum[co] += tp[co];
        }
#pragma unroll
    for (int co = 0; co < kFmCout; ++co) out[(((size_t)b * kFmCout + co) * S + y0 + ly) * S + x0 + lx] = tanhf(sum[co]);
}

// -------------------------------------------------------------------------------------------------
// g = dout * (1 - out^2) of a tile with a one-pixel halo (zero outside the image): gs[co][(TH+2)][(TW+2)] fp32
// -------------------------------------------------------------------------------------------------
constexpr int kFmGsN = kFmCout * (kFmTH + 2) * (kFmTW + 2);               // 1020 values
constexpr int kFmGsPer = (kFmGsN + kFmThreads - 1) / kFmThreads;         // 4 per thread
struct FmGRegs { float v[kFmGsPer]; };
// loads of a tile's g into registers: all 8 (out, dout) loads of a thread are issued together; fm_store_g writes them to
// shared memory later (the dw kernel fetches the next tile's g while it computes on the current one)
__device__ __forceinline__ void fm_load_g(FmGRegs &r, const float *__restrict__ out, const float *__restrict__ dout, int b, int S, int x0,
                                          int y0)
{
    constexpr int gw = kFmTW + 2, gh = kFmTH + 2;
    float o[kFmGsPer], d[kFmGsPer];
#pragma unroll
    for (int k = 0; k < kFmGsPer; ++k) {
        const int i = threadIdx.x + k * kFmThreads;
        const int gx = i % gw, gy = (i / gw) % gh, co = i / (gw * gh);
        const int yy = y0 + gy - 1, xx = x0 + gx - 1;
        const bool ok = i < kFmGsN && yy >= 0 && yy < S && xx >= 0 && xx < S;
        const size_t idx = ok ? (((size_t)b * kFmCout + co) * S + yy) * S + xx : 0;
        o[k] = ok ? __ldg(out + idx) : 0.f;
        d[k] = ok ? __ldg(dout + idx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kFmGsPer; ++k) r.v[k] = d[k] * (1.f - o[k] * o[k]);
}
__device__ __forceinline__ void fm_store_g(float *gs, const FmGRegs &r)
{
#pragma unroll
    for (int k = 0; k < kFmGsPer; ++k) {
        const int i = threadIdx.x + k * kFmThreads;
        if (i < kFmGsN) gs[i] = r.v[k];
    }
}
__device__ __forceinline__ void fm_stage_g(float *gs, const float *__restrict__ out, const float *__restrict__ dout, int b, int S,
                                           int x0, int y0)
{
    FmGRegs r;
    fm_load_g(r, out, dout, b, S, x0, y0);
    fm_store_g(gs, r);
}

// -------------------------------------------------------------------------------------------------
// dx
// -------------------------------------------------------------------------------------------------
constexpr int kFmPix = kFmTH * kFmTW;                   // 256 output pixels = 16 m16 tiles
constexpr int kFmGPitch = kFmN * 2 + 16;                // bytes per im2col row [pixel][32 k]   (20 words)
// shared memory: gs [3][10][34] fp32 | Gh, Gl [256][80 B] | wh, wl [64 ci][80 B] (K = (tap, co) contiguous)
constexpr size_t kFmGsBytes = ((size_t)kFmCout * (kFmTH + 2) * (kFmTW + 2) * 4 + 15) / 16 * 16;
constexpr size_t kFmDxSmem = kFmGsBytes + 2 * (size_t)kFmPix * kFmGPitch + 2 * (size_t)kFmC * kFmGPitch;

__global__ void __launch_bounds__(kFmThreads, 2) final_conv_tanh_bwd_x_mma_kernel(const float *__restrict__ w,
                                                                                  const float *__restrict__ out,
                                                                                  const float *__restrict__ dout,
                                                                                  __nv_bfloat16 *__restrict__ dx, int S, int tiles_x)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    float *gs = reinterpret_cast<float *>(fm_smem);
    unsigned char *Gh = fm_smem + kFmGsBytes;
    unsigned char *Gl = Gh + (size_t)kFmPix * kFmGPitch;
    unsigned char *wh = Gl + (size_t)kFmPix * kFmGPitch;
    unsigned char *wl = wh + (size_t)kFmC * kFmGPitch;
    const int b = blockIdx.y, tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int x0 = tile_x * kFmTW, y0 = tile_y * kFmTH;
    constexpr int gw = kFmTW + 2, gh = kFmTH + 2;

    // g and the weights: every global load of the thread is issued before the first result is used
    FmGRegs greg;
    fm_load_g(greg, out, dout, b, S, x0, y0);
    // B operand: rows n = ci, K index k = tap * 3 + co (27 -> 32, zero padded), hi / lo
    {
        float wv[kFmC * kFmN / kFmThreads];
#pragma unroll
        for (int j = 0; j < kFmC * kFmN / kFmThreads; ++j) {
            const int i = threadIdx.x + j * kFmThreads, k = i % kFmN, ci = i / kFmN;
            const int tap = k / kFmCout, co = k - tap * kFmCout;
            wv[j] = k < 9 * kFmCout ? __ldg(w + ((size_t)co * kFmC + ci) * 9 + tap) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kFmC * kFmN / kFmThreads; ++j) {
            const int i = threadIdx.x + j * kFmThreads, k = i % kFmN, ci = i / kFmN;
            __nv_bfloat16 hi, lo;
            split_bf16(wv[j], hi, lo);
            *reinterpret_cast<__nv_bfloat16 *>(wh + (size_t)ci * kFmGPitch + k * 2) = hi;
            *reinterpret_cast<__nv_bfloat16 *>(wl + (size_t)ci * kFmGPitch + k * 2) = lo;
        }
    }
    fm_store_g(gs, greg);
    __syncthreads();
    // A operand: im2col of g.  dx[y][x] takes tap (ty, tx) from g[y - ty + 1][x - tx + 1] = halo (ly + 2 - ty, lx + 2 - tx)
    for (int i = threadIdx.x; i < kFmPix * kFmN; i += kFmThreads) {
        const int k = i % kFmN, p = i / kFmN;
        __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
        if (k < 9 * kFmCout) {
            const int tap = k / kFmCout, co = k - tap * kFmCout, ty = tap / 3, tx = tap - ty * 3;
            const int lx = p % kFmTW, ly = p / kFmTW;
            split_bf16(gs[(co * gh + ly + 2 - ty) * gw + lx + 2 - tx], hi, lo);
        }
        *reinterpret_cast<__nv_bfloat16 *>(Gh + (size_t)p * kFmGPitch + k * 2) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(Gl + (size_t)p * kFmGPitch + k * 2) = lo;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    uint32_t bh[8][2][2], bl[8][2][2];                  // 8 n-tiles (64 ci) x 2 k-chunks
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
            const size_t off = (size_t)(nt * 8 + g) * kFmGPitch + (kc * 16 + 2 * q) * 2;
            bh[nt][kc][0] = lds32(wh + off); bh[nt][kc][1] = lds32(wh + off + 16);
            bl[nt][kc][0] = lds32(wl + off); bl[nt][kc][1] = lds32(wl + off + 16);
        }
    __nv_bfloat16 *dxb = dx + (size_t)b * S * S * kFmC;
    for (int mt = warp; mt < kFmPix / 16; mt += kFmThreads / 32) {
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
            const size_t off = (size_t)(mt * 16 + g) * kFmGPitch + (kc * 16 + 2 * q) * 2;
            uint32_t ah[4], al[4];
            ah[0] = lds32(Gh + off); ah[1] = lds32(Gh + off + 8 * kFmGPitch); ah[2] = lds32(Gh + off + 16); ah[3] = lds32(Gh + off + 8 * kFmGPitch + 16);
            al[0] = lds32(Gl + off); al[1] = lds32(Gl + off + 8 * kFmGPitch); al[2] = lds32(Gl + off + 16); al[3] = lds32(Gl + off + 8 * kFmGPitch + 16);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mma_bf16_16816(acc[nt], ah, bh[nt][kc][0], bh[nt][kc][1]);
                mma_bf16_16816(acc[nt], al, bh[nt][kc][0], bh[nt][kc][1]);
                mma_bf16_16816(acc[nt], ah, bl[nt][kc][0], bl[nt][kc][1]);
            }
        }
        // rows p = mt * 16 + g and + 8: tile pixel (ly, lx) -> dx[y0 + ly][x0 + lx][ci = nt * 8 + 2q, +1]
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int p = mt * 16 + g + half * 8, lx = p % kFmTW, ly = p / kFmTW;
            __nv_bfloat16 *row = dxb + ((size_t)(y0 + ly) * S + x0 + lx) * kFmC + 2 * q;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<__nv_bfloat162 *>(row + nt * 8) = __floats2bfloat162_rn(acc[nt][half * 2], acc[nt][half * 2 + 1]);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// dw (+ dbias): per-CTA partials in the layout of final_conv_reduce_kernel,
//   part[(cta * Cout + co) * (9 * C + 1) + tap * C + ci],  bias sum at index 9 * C
// -------------------------------------------------------------------------------------------------
constexpr int kFmGtPitch = kFmPix * 2 + 16;             // bytes per row of Gt [n][pixel]   (132 words)
// shared memory: gs [3][10][34] fp32 | Gth, Gtl [32][528 B] | xs0, xs1 [256][144 B] | red [8 warps][32][64] fp32 (alias)
constexpr size_t kFmDwSmem = kFmGsBytes + 2 * (size_t)kFmN * kFmGtPitch + 2 * (size_t)kFmPix * kFmXPitch;
constexpr size_t kFmDwRedBytes = (size_t)(kFmThreads / 32) * kFmN * kFmC * 4;      // 64 KB, aliases the staging buffers

__global__ void __launch_bounds__(kFmThreads, 1) final_conv_tanh_bwd_w_mma_kernel(const __nv_bfloat16 *__restrict__ x,
                                                                                  const float *__restrict__ out,
                                                                                  const float *__restrict__ dout,
                                                                                  float *__restrict__ part, int S, int B,
                                                                                  int tiles_x, int tiles_y)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    __shared__ float bsum[kFmCout][kFmThreads / 32];
    float *gs = reinterpret_cast<float *>(fm_smem);
    unsigned char *Gth = fm_smem + kFmGsBytes;
    unsigned char *Gtl = Gth + (size_t)kFmN * kFmGtPitch;
    unsigned char *xs0 = Gtl + (size_t)kFmN * kFmGtPitch;
    unsigned char *xs1 = xs0 + (size_t)kFmPix * kFmXPitch;
    constexpr int gw = kFmTW + 2, gh = kFmTH + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int n_tiles = B * tiles_x * tiles_y;

    float acc[2][8][4];                                 // 2 m-tiles (32 rows n) x 8 n-tiles (64 ci)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
    float gsum[kFmCout] = {0.f, 0.f, 0.f};

    // Software pipeline over the CTA's tiles: the x tile of trip i + 1 streams into the other staging buffer (cp.async) and
    // its g values wait in registers while trip i runs its im2col and MMAs.
    auto fetch = [&](int tile, unsigned char *xbuf, FmGRegs &gr) {
        const int tile_x = tile % tiles_x, tile_y = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int x0 = tile_x * kFmTW, y0 = tile_y * kFmTH;
        // x tile (no halo): the K index of this GEMM is the tile pixel p = ly * TW + lx
        const __nv_bfloat16 *xb = x + (size_t)b * S * S * kFmC;
#pragma unroll
        for (int i0 = 0; i0 < kFmPix * (kFmC / 8); i0 += kFmThreads) {
            const int i = i0 + threadIdx.x;
            const int v = i % (kFmC / 8), p = i / (kFmC / 8), lx = p % kFmTW, ly = p / kFmTW;
            cp_async_16_zfill(xbuf + (size_t)p * kFmXPitch + v * 16, xb + ((size_t)(y0 + ly) * S + x0 + lx) * kFmC + v * 8, true);
        }
        fm_load_g(gr, out, dout, b, S, x0, y0);
    };
    FmGRegs greg;
    int buf = 0;
    if ((int)blockIdx.x < n_tiles) fetch(blockIdx.x, xs0, greg);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        unsigned char *xs = buf ? xs1 : xs0;
        __syncthreads();                                // the previous trip's MMAs are done with gs, Gt and the other x buffer
        fm_store_g(gs, greg);
        __syncthreads();
        const int next = tile + gridDim.x;
        if (next < n_tiles) fetch(next, buf ? xs0 : xs1, greg);
        cp_async_commit();                              // (possibly empty) group of the next tile
        // Gt[n = tap * 3 + co][p]: input pixel p = (ly, lx) meets g at (ly - ty + 1, lx - tx + 1) = halo (ly + 2 - ty, lx + 2 - tx)
        for (int i = threadIdx.x; i < kFmN * kFmPix; i += kFmThreads) {
            const int p = i % kFmPix, n = i / kFmPix;
            __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
            if (n < 9 * kFmCout) {
                const int tap = n / kFmCout, co = n - tap * kFmCout, ty = tap / 3, tx = tap - ty * 3;
                const int lx = p % kFmTW, ly = p / kFmTW;
                split_bf16(gs[(co * gh + ly + 2 - ty) * gw + lx + 2 - tx], hi, lo);
            }
            *reinterpret_cast<__nv_bfloat16 *>(Gth + (size_t)n * kFmGtPitch + p * 2) = hi;
            *reinterpret_cast<__nv_bfloat16 *>(Gtl + (size_t)n * kFmGtPitch + p * 2) = lo;
        }
        // dbias: interior g of this tile, thread = pixel
        {
            const int lx = threadIdx.x % kFmTW, ly = threadIdx.x / kFmTW;
#pragma unroll
            for (int co = 0; co < kFmCout; ++co) gsum[co] += gs[(co * gh + ly + 1) * gw + lx + 1];
        }
        cp_async_wait_group<1>();                       // this trip's x tile has landed (the next one may still be in flight)
        __syncthreads();
        // K = 256 pixels = 16 k-chunks; warp w takes chunks w, w + 8
        for (int kc = warp; kc < kFmPix / 16; kc += kFmThreads / 32) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const size_t off = (size_t)(mt * 16 + g) * kFmGtPitch + (kc * 16 + 2 * q) * 2;
                ah[mt][0] = lds32(Gth + off); ah[mt][1] = lds32(Gth + off + 8 * kFmGtPitch);
                ah[mt][2] = lds32(Gth + off + 16); ah[mt][3] = lds32(Gth + off + 8 * kFmGtPitch + 16);
                al[mt][0] = lds32(Gtl + off); al[mt][1] = lds32(Gtl + off + 8 * kFmGtPitch);
                al[mt][2] = lds32(Gtl + off + 16); al[mt][3] = lds32(Gtl + off + 8 * kFmGtPitch + 16);
            }
            // B[k = pixel][n = ci]: b0 = {x[p0 + 2q][ci], x[p0 + 2q + 1][ci]}, b1 = the same 8 pixels further, ci = nt * 8 + g
            const unsigned char *xr = xs + (size_t)(kc * 16 + 2 * q) * kFmXPitch + g * 2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const unsigned char *xp = xr + nt * 16;
                const uint32_t e0 = *reinterpret_cast<const uint16_t *>(xp), e1 = *reinterpret_cast<const uint16_t *>(xp + kFmXPitch);
                const uint32_t e2 = *reinterpret_cast<const uint16_t *>(xp + 8 * kFmXPitch),
                               e3 = *reinterpret_cast<const uint16_t *>(xp + 9 * kFmXPitch);
                const uint32_t b0 = e0 | (e1 << 16), b1 = e2 | (e3 << 16);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16_16816(acc[mt][nt], ah[mt], b0, b1);
                    mma_bf16_16816(acc[mt][nt], al[mt], b0, b1);
                }
            }
        }
        buf ^= 1;
    }
    cp_async_wait_group<0>();
    // ---- fixed-order reduction over the 8 warps, then the per-CTA partial -------------------------------------
#pragma unroll
    for (int co = 0; co < kFmCout; ++co) gsum[co] = warp_sum(gsum[co]);
    __syncthreads();                                    // staging buffers are free: red aliases them
    float *red = reinterpret_cast<float *>(fm_smem);   // [warp][n (32)][ci (64)]
    if (lane == 0) {
#pragma unroll
        for (int co = 0; co < kFmCout; ++co) bsum[co][warp] = gsum[co];
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float *r0 = red + ((size_t)warp * kFmN + mt * 16 + g) * kFmC + nt * 8 + 2 * q;
            r0[0] = acc[mt][nt][0]; r0[1] = acc[mt][nt][1];
            r0[8 * kFmC] = acc[mt][nt][2]; r0[8 * kFmC + 1] = acc[mt][nt][3];
        }
    __syncthreads();
    constexpr int per = 9 * kFmC + 1;
    float *pout = part + (size_t)blockIdx.x * kFmCout * per;
    for (int i = threadIdx.x; i < 9 * kFmCout * kFmC; i += kFmThreads) {
        const int ci = i % kFmC, n = i / kFmC, tap = n / kFmCout, co = n - tap * kFmCout;
        float s = 0.f;
#pragma unroll
        for (int wgt = 0; wgt < kFmThreads / 32; ++wgt) s += red[((size_t)wgt * kFmN + n) * kFmC + ci];
        pout[(size_t)co * per + tap * kFmC + ci] = s;
    }
    if (threadIdx.x < kFmCout) {
        float s = 0.f;
        for (int i = 0; i < kFmThreads / 32; ++i) s += bsum[threadIdx.x][i];
        pout[(size_t)threadIdx.x * per + 9 * kFmC] = s;
    }
}

static_assert(kFmDwRedBytes <= kFmDwSmem, "the reduction scratch must fit in the staging buffers it aliases");

}  // namespace hg

using namespace hg;

// Called from final_conv.cu.  Shapes: Cin = 64, Cout = 3, S % 32 == 0 (checked by the caller through
// hg_final_conv_mma_supported).
bool hg_final_conv_mma_supported(int cin, int cout, int size) { return cin == kFmC && cout == kFmCout && size >= 32 && size % kFmTW == 0; }

int hg_final_conv_fwd_mma(const void *x, const float *w, const float *bias, float *out, int batch, int size, cudaStream_t st)
{
    static bool attr_done = false;      // not a stream operation (graph-capture safe)
    if (!attr_done) {
        cudaFuncSetAttribute(final_conv_tanh_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmFwdSmem);
        attr_done = true;
    }
    const int tiles_x = size / kFmTW, tiles_y = size / kFmTH;
    final_conv_tanh_fwd_mma_kernel<<<dim3(tiles_x * tiles_y, batch), kFmThreads, kFmFwdSmem, st>>>(
        static_cast<const __nv_bfloat16 *>(x), w, bias, out, size, tiles_x);
    return check_launch("hg_final_conv_tanh_fwd(mma)");
}

int hg_final_conv_bwd_x_mma(const float *w, const float *out, const float *dout, void *dx, int batch, int size, cudaStream_t st)
{
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(final_conv_tanh_bwd_x_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmDxSmem);
        attr_done = true;
    }
    const int tiles_x = size / kFmTW, tiles_y = size / kFmTH;
    final_conv_tanh_bwd_x_mma_kernel<<<dim3(tiles_x * tiles_y, batch), kFmThreads, kFmDxSmem, st>>>(
        w, out, dout, static_cast<__nv_bfloat16 *>(dx), size, tiles_x);
    return check_launch("hg_final_conv_tanh_bwd(x, mma)");
}

// part: n_cta * Cout * (9 * Cin + 1) floats (the caller's workspace); reduced by final_conv_reduce_kernel
int hg_final_conv_bwd_w_mma(const void *x, const float *out, const float *dout, float *part, int n_cta, int batch, int size,
                            cudaStream_t st)
{
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(final_conv_tanh_bwd_w_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmDwSmem);
        attr_done = true;
    }
    const int tiles_x = size / kFmTW, tiles_y = size / kFmTH;
    final_conv_tanh_bwd_w_mma_kernel<<<n_cta, kFmThreads, kFmDwSmem, st>>>(static_cast<const __nv_bfloat16 *>(x), out, dout, part,
                                                                          size, batch, tiles_x, tiles_y);
    return check_launch("hg_final_conv_tanh_bwd(w, mma)");
}
