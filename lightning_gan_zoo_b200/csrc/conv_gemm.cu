// tcgen05 implicit-GEMM kernels for the dense contractions of the HoloGAN generator:
// ConvTranspose3d(k3,s2,p1,op1), ConvTranspose2d(k4,s2,p1) and the 1x1 projection -- forward, dgrad and
// wgrad (reference core/models/hologan_generator.py:25-30,60,135,38).
//
// Formulation.  A stride-2 transposed convolution is split by output parity class: inside one class
// every output position (b, i) is a plain sum over a few "taps", each tap being the input window
// shifted by (sz,sy,sx) in {-1,0,+1}^d times one [Cout x Cin] slice of the weight.  Activations are
// channels-last bf16, so the A operand of a tap is a TMA box of the (C, X, Y, Z, B) tensor at shifted
// coordinates -- out-of-range rows are zero-filled by the TMA unit (the padding), nothing is ever
// materialised (no im2col, no zero-insertion).  Conv outputs and their gradients live in a
// "space-to-depth" layout (B, [Z,] Y, X, P*Cout), P = 2^d parity classes, which makes
//   forward : Ys2d[pos, cls*Cout + co] = sum_taps X[pos + s][ci] * W[ci, co, tap]        (K-major A, B)
//   dgrad   : dX[pos, ci] = sum_(cls,tap) dYs2d[pos - s, cls*Cout + co] * W[ci, co, tap]  (K-major A, B)
//   wgrad   : dW[tap][ci, co] = sum_pos X[pos + s][ci] * dYs2d[pos, cls*Cout + co]        (MN-major A, B)
// three instances of one warp-specialised pipeline:  TMA producer warp -> 128B-swizzled smem ring ->
// single-thread tcgen05.mma issue (fp32 accumulators in TMEM) -> 4 epilogue warps (tcgen05.ld).
#include <cuda.h>

#include "hg_common.cuh"
#include <algorithm>

#include "sm100_ptx.cuh"

namespace hg {

constexpr int kBM = 128;              // rows of the accumulator tile = TMEM lanes
constexpr int kBK = 64;               // bf16 elements per 128-byte swizzle row
constexpr int kThreads = 192;         // warp 0: TMA, warp 1: TMEM alloc + MMA, warps 2-5: epilogue
constexpr int kMaxStages = 8;
constexpr int kMaxTaps = 32;
constexpr int kMaxGroups = 16;
constexpr int kSmemBudget = 200 * 1024;
constexpr int kDualSmemBudget = 100 * 1024;       // "dual" launches: two tap-GEMM CTAs per SM (<= 256 TMEM columns each)
constexpr long long kDualMaxBytes = 2ll << 20;    // ... chosen when a CTA streams at most this many operand bytes
constexpr int kWgradSmemBudget = 100 * 1024;      // two wgrad CTAs per SM

struct Tap {
    int16_t sx, sy, sz;
    int16_t acc;          // which accumulator (TMEM column block of BN columns) this tap adds into
    int32_t a_c_off;      // channel offset added to the A box (selects the parity class in s2d tensors)
    int32_t b_row_off;    // first row of this tap's slice in the packed weight matrix
};
// Shift-major work item (forward of the narrow layers, several parity classes per CTA): all (class, tap) pairs of the
// CTA's classes that read the SAME shifted A box.  The box is loaded once per K chunk, the weight boxes of the pairs are
// stacked behind each other in the stage, and one MMA per run of TMEM-adjacent classes (N = run * Cout <= 256) adds into
// their accumulators -- 9 A boxes instead of 16 (k4 2-D) or 8 instead of 27 (k3 3-D), and N up to 256 instead of Cout.
constexpr int kItemBoxes = 4;
struct Item {
    int16_t sx, sy, sz;
    uint8_t nb, nseg;                 // weight boxes (<= kItemBoxes, nb * BN <= 256 rows) and MMA segments
    int32_t a_c_off;
    int32_t b_row[kItemBoxes];        // first row of every box in the packed weight matrix
    struct Seg {
        uint16_t col;                 // first accumulator column
        uint16_t b_units;             // offset of the segment's first weight row inside the stage's B region, 16-byte units
        uint32_t idesc;               // instruction descriptor with N = run * BN
        uint32_t first;               // 1: no earlier MMA of this tile wrote these columns (the first K step overwrites)
    } seg[kItemBoxes];
};
struct Group {
    int32_t tap_begin, tap_count;
    int32_t out_col_off;  // column offset of this group's output (parity class) in the output row
    int32_t part;         // split-K: which fp32 partial buffer this group writes (TapGemmParams::partial)
};

enum EpilogueMode { kEpiBf16 = 0, kEpiF32Atomic = 1 };

struct TapGemmParams {
    CUtensorMap tmA;      // K-major kernel: activations (C, X, Y, Z, B), box (64, bx, by, bz, bb), prod = 128
    CUtensorMap tmB;      // K-major kernel: packed weights (Kcols, rows), box (64, BN)
    int X, Y, Z, Bn;      // spatial extents / batch of the A tensor (for the tile -> coordinate decode)
    int BN;               // columns of one accumulator = N of one MMA = rows of one weight box
    int epi_cols;         // columns the epilogue drains = accumulators per CTA * BN (<= 512)
    int bias_mod;         // bias index = output column % bias_mod
    int k_chunks;         // 64-wide K chunks per tap
    int stages;
    int m_total;          // valid rows
    long long ld_out;     // elements per output row
    void *out;
    const float *bias;    // per output column (without group offset) or null
    float slope;          // act(v) = v > 0 ? v : slope * v   (1 = identity)
    // split-K over taps (small-M layers: the discriminator's convolutions): non-null -> the epilogue stores raw fp32
    // accumulators to partial + group.part * partial_stride + m * ld_out + col (plain stores, no bias / activation);
    // splitk_reduce_kernel sums the partials of every column block in a fixed order and writes the bf16 result
    float *partial;
    long long partial_stride;
    float *stats;               // may be null.  AdaIN statistics fused into the epilogue (north_star "AdaIN fused into the GEMM
                                // epilogues"): per 32-row group g = m / 32 and output column, the sum and the sum of squares of
                                // the fp32 results (after bias, before the activation): stats[(g * ld_out + col) * 2 + {0, 1}].
                                // hg_adain_cl_fwd_stats merges the groups of a sample (Chan) -- the conv output is not re-read.
    const float *out_scale;     // device scalar (may be null): accumulators are multiplied by it before bias / activation
                                // (1 / sigma of the spectral norm: conv(x, W / sigma) = conv(x, W) / sigma)
    // tile scheduler: a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the (group, n tile, m tile) space
    // (m fastest, groups in the order of `order`: longest K loops first).  A one-tile-per-CTA launch is the special case
    // gridDim.x == total tiles.  nacc accumulator stages of acc_stride TMEM columns each: with two, the epilogue of
    // tile i overlaps the main loop of tile i + 1.
    int m_tiles, n_tiles;
    int m_sub;                  // 128-row sub-tiles per CTA tile (1 or 2).  With 2, both share every B (weight) tile of the K
                                // loop: 64 KB per 256 x 256 x 64 MMA block instead of 2 x 48 KB -- the wide GEMMs are bound by
                                // the bytes they must keep in flight from L2, not by the tensor pipe (profiles/r02o_*)
    int nacc, acc_stride;       // acc_stride covers all sub-tiles of one stage (m_sub * sub_stride)
    int sub_stride;
    int stats_scratch;          // 1: the epilogue's statistics scratch sits behind the operand ring (persistent launches)
    int num_groups;
    uint8_t order[kMaxGroups];
    Group groups[kMaxGroups];
    Tap taps[kMaxTaps];
    int item_rows;              // > 0: the groups index `items` (tap_begin / tap_count count items); rows of the stage's B region
    Item items[kMaxTaps];
};

struct WgradParams {
    CUtensorMap tmX;      // activations (Cin, X, Y, Z, B), box (64, px, py, pz, pb), prod = 64 positions
    CUtensorMap tmDY;     // output gradients, s2d (P*Cout, X, Y, Z, B), same box
    int X, Y, Z, Bn;
    int BN;               // Cout tile
    int stages;
    int pos_tiles;        // number of 64-position tiles
    int splits;           // split-K factor (grid.z = pairs * splits)
    int cin, cout;
    int num_taps_total;   // k^d slices per split in the partial buffer
    float *dw;            // split-K partials, fp32 [split][tap][Cin][Cout] (plain stores; reduced by wgrad_reduce_kernel)
    int num_pairs;
    struct Pair {
        int16_t sx, sy, sz, pad;
        int32_t dy_c_off;     // cls * Cout
        int32_t tap_flat;     // slice index in dw
    } pairs[kMaxTaps];
};

struct SharedCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_ready;         // wgrad kernel (one tile per CTA)
    uint64_t acc_full[2];       // tap kernel: MMA warp -> epilogue warps, per accumulator stage
    uint64_t acc_empty[2];      // epilogue warps -> MMA warp
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t tmem_cols_for(int bn)
{
    return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : bn <= 256 ? 256u : 512u;
}

__device__ __forceinline__ uint8_t *align_1024(uint8_t *p)
{
    return reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// -------------------------------------------------------------------------------------------------
// K-major kernel: forward and dgrad (and plain GEMMs).  1-D grid over the (group, n tile, m tile) list: one tile per CTA,
// or one persistent CTA per SM walking the list (TapGemmParams: tile scheduler)
// -------------------------------------------------------------------------------------------------
template <int MSUB, bool ITEMS>
__global__ void __launch_bounds__(kThreads, 1) tap_gemm_kernel(const __grid_constant__ TapGemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedCtl ctl;
    __shared__ __align__(16) float bias_stage[4][512 + 16];        // one copy of the tile's bias columns per epilogue warp
    uint8_t *tiles = align_1024(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = p.BN;
    const uint32_t a_bytes = kBM * 128, b_bytes = (uint32_t)BN * 128;
    const uint32_t stage_bytes = ITEMS ? a_bytes + (uint32_t)p.item_rows * 128 : (uint32_t)MSUB * a_bytes + b_bytes;
    const uint32_t b_off = (uint32_t)MSUB * a_bytes;               // stage layout: [A sub 0][A sub 1]?[B]
    const int tiles_mn = p.m_tiles * p.n_tiles, total_tiles = tiles_mn * p.num_groups;
    const uint32_t tmem_cols = tmem_cols_for(p.nacc * p.acc_stride);

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&p.tmA);
        ptx::prefetch_tensormap(&p.tmB);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&ctl.full[s], 1);
            ptx::mbar_init(&ctl.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&ctl.acc_full[a], 1);
            ptx::mbar_init(&ctl.acc_empty[a], 4);       // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&ctl.tmem_base, tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl.tmem_base;
    // The producer and the MMA issuer are ONE thread each and the K loops are short: ring slot / phase are advanced
    // incrementally (no division), barrier and tile addresses stay in the shared window, descriptors are derived from one
    // base by adds (same-box A/B, profiles/r02v_tap_gemm_ab.txt: a few % on the narrow layers; the large term was the
    // epilogue's per-element bias branches, see below).
    const uint32_t tiles_a = ptx::smem_u32(tiles), full_a = ptx::smem_u32(&ctl.full[0]), empty_a = ptx::smem_u32(&ctl.empty[0]);
    const uint32_t nstages = (uint32_t)p.stages;

    if (warp == 0) {
        if (ptx::elect_one()) {
            // ===== TMA producer: runs ahead of the MMA warp across tile boundaries =====
            uint32_t s = 0, ph = 1;                     // waits on empty[s] with the inverted phase: passes on a fresh barrier
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int gi = t / tiles_mn, r = t - gi * tiles_mn, n_idx = r / p.m_tiles, m_idx = r - n_idx * p.m_tiles;
                const Group grp = p.groups[p.order[gi]];
                const int n0 = n_idx * BN;
                int x0[MSUB], y0[MSUB], z0[MSUB], b0[MSUB];
#pragma unroll
                for (int sub = 0; sub < MSUB; ++sub) {      // a sub-tile past the end of the tensor lands out of range: zero fill
                    uint32_t pos = ((uint32_t)m_idx * MSUB + sub) * kBM;        // 32-bit: m_total fits an int (64-bit divisions
                    x0[sub] = (int)(pos % (uint32_t)p.X); pos /= (uint32_t)p.X;  // cost ~100 instructions each in this one thread)
                    y0[sub] = (int)(pos % (uint32_t)p.Y); pos /= (uint32_t)p.Y;
                    z0[sub] = (int)(pos % (uint32_t)p.Z);
                    b0[sub] = (int)(pos / (uint32_t)p.Z);
                }
                if constexpr (ITEMS) {
                    for (int ii = 0; ii < grp.tap_count; ++ii) {
                        const Item &it = p.items[grp.tap_begin + ii];
                        const int cx = x0[0] + it.sx, cy = y0[0] + it.sy, cz = z0[0] + it.sz, nb = it.nb;
                        const uint32_t tx = a_bytes + (uint32_t)nb * b_bytes;
                        const int r0 = it.b_row[0] + n0, r1 = it.b_row[1] + n0, r2 = it.b_row[2] + n0, r3 = it.b_row[3] + n0;
                        for (int kc = 0, kcol = 0; kc < p.k_chunks; ++kc, kcol += kBK) {
                            const uint32_t fb = full_a + 8u * s, dst = tiles_a + s * stage_bytes;
                            ptx::mbar_wait_a(empty_a + 8u * s, ph);
                            ptx::mbar_arrive_expect_tx_a(fb, tx);
                            ptx::tma_load_5d_a(dst, &p.tmA, fb, it.a_c_off + kcol, cx, cy, cz, b0[0]);
                            ptx::tma_load_2d_a(dst + a_bytes, &p.tmB, fb, kcol, r0);
                            if (nb > 1) ptx::tma_load_2d_a(dst + a_bytes + b_bytes, &p.tmB, fb, kcol, r1);
                            if (nb > 2) ptx::tma_load_2d_a(dst + a_bytes + 2 * b_bytes, &p.tmB, fb, kcol, r2);
                            if (nb > 3) ptx::tma_load_2d_a(dst + a_bytes + 3 * b_bytes, &p.tmB, fb, kcol, r3);
                            if (++s == nstages) { s = 0; ph ^= 1u; }
                        }
                    }
                } else
                for (int tp = 0; tp < grp.tap_count; ++tp) {
                    const Tap tap = p.taps[grp.tap_begin + tp];
                    const int brow = tap.b_row_off + n0;
                    int cx[MSUB], cy[MSUB], cz[MSUB];
#pragma unroll
                    for (int sub = 0; sub < MSUB; ++sub) { cx[sub] = x0[sub] + tap.sx; cy[sub] = y0[sub] + tap.sy; cz[sub] = z0[sub] + tap.sz; }
                    for (int kc = 0, kcol = 0; kc < p.k_chunks; ++kc, kcol += kBK) {
                        const uint32_t fb = full_a + 8u * s, dst = tiles_a + s * stage_bytes;
                        ptx::mbar_wait_a(empty_a + 8u * s, ph);
                        ptx::mbar_arrive_expect_tx_a(fb, stage_bytes);
#pragma unroll
                        for (int sub = 0; sub < MSUB; ++sub)
                            ptx::tma_load_5d_a(dst + sub * a_bytes, &p.tmA, fb, tap.a_c_off + kcol, cx[sub], cy[sub], cz[sub], b0[sub]);
                        ptx::tma_load_2d_a(dst + b_off, &p.tmB, fb, kcol, brow);
                        if (++s == nstages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            // ===== MMA issuer =====
            const uint32_t idesc = ptx::idesc_bf16(kBM, BN, false, false);
            // descriptor of a K-major 128B-swizzled tile at shared address 0; the address field (bits 0-13, 16-byte units)
            // is added per stage -- shared addresses stay below 256 KB, so the add never carries out of the field
            const uint64_t desc0 = ptx::smem_desc_sw128(0, 16, 1024);
            const uint32_t stage_units = stage_bytes >> 4, b_units = b_off >> 4, a_units = a_bytes >> 4, tiles_units = tiles_a >> 4;
            uint32_t s = 0, ph = 0;
            int i = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
                const int gi = t / tiles_mn;
                const Group grp = p.groups[p.order[gi]];
                const int acc_stage = i % p.nacc;
                const uint32_t aph = (uint32_t)(i / p.nacc) & 1u;
                ptx::mbar_wait(&ctl.acc_empty[acc_stage], aph ^ 1u);       // the epilogue has drained this stage
                ptx::tc_fence_after();
                if constexpr (ITEMS) {
                    const uint32_t acc0 = tmem_base + (uint32_t)(acc_stage * p.acc_stride);
                    for (int ii = 0; ii < grp.tap_count; ++ii) {
                        const Item &it = p.items[grp.tap_begin + ii];
                        const int nseg = it.nseg;
                        for (int kc = 0; kc < p.k_chunks; ++kc) {
                            ptx::mbar_wait_a(full_a + 8u * s, ph);
                            ptx::tc_fence_after();
                            const uint64_t a_desc = desc0 + (uint64_t)(tiles_units + s * stage_units);
                            const uint64_t b_base = a_desc + a_units;
#pragma unroll
                            for (int sg = 0; sg < kItemBoxes; ++sg) {
                                if (sg < nseg) {
                                    const Item::Seg seg = it.seg[sg];
                                    const uint32_t d = acc0 + seg.col;
                                    const uint64_t b_desc = b_base + seg.b_units;
                                    ptx::umma_bf16(d, a_desc, b_desc, seg.idesc, (seg.first == 0u) || kc != 0);
                                    ptx::umma_bf16(d, a_desc + 2, b_desc + 2, seg.idesc, true);
                                    ptx::umma_bf16(d, a_desc + 4, b_desc + 4, seg.idesc, true);
                                    ptx::umma_bf16(d, a_desc + 6, b_desc + 6, seg.idesc, true);
                                }
                            }
                            ptx::umma_commit_a(empty_a + 8u * s);
                            if (++s == nstages) { s = 0; ph ^= 1u; }
                        }
                    }
                } else {
                uint32_t started = 0;                       // accumulators that already hold a partial sum
                for (int tp = 0; tp < grp.tap_count; ++tp) {
                    const int acc = p.taps[grp.tap_begin + tp].acc;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc_stage * p.acc_stride + acc * BN);
                    uint32_t accumulate = (started >> acc) & 1u;
                    started |= 1u << acc;
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        ptx::mbar_wait_a(full_a + 8u * s, ph);
                        ptx::tc_fence_after();
                        const uint64_t a_desc = desc0 + (uint64_t)(tiles_units + s * stage_units);
                        const uint64_t b_desc = a_desc + b_units;
#pragma unroll
                        for (int sub = 0; sub < MSUB; ++sub) {          // the sub-tiles share this stage's B tile
                            const uint64_t as = a_desc + (uint64_t)(sub * a_units);
                            const uint32_t d = d_tmem + (uint32_t)(sub * p.sub_stride);
                            // +32 bytes (2 units) per 16-element K step inside the swizzle row
                            ptx::umma_bf16(d, as, b_desc, idesc, accumulate != 0);
                            ptx::umma_bf16(d, as + 2, b_desc + 2, idesc, true);
                            ptx::umma_bf16(d, as + 4, b_desc + 4, idesc, true);
                            ptx::umma_bf16(d, as + 6, b_desc + 6, idesc, true);
                        }
                        accumulate = 1u;
                        ptx::umma_commit_a(empty_a + 8u * s);   // frees the smem slot when these MMAs retire
                        if (++s == nstages) { s = 0; ph ^= 1u; }
                    }
                }
                }
                ptx::umma_commit(&ctl.acc_full[acc_stage]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias / activation -> bf16 -> global =====
        const int quad = warp & 3;                      // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int row = quad * 32 + lane;
        // statistics scratch: behind the ring when tiles overlap (persistent), else the idle ring itself
        float *bias_s = bias_stage[quad];               // per-warp copy: no cross-warp synchronisation
        float *scr = reinterpret_cast<float *>(p.stats_scratch ? tiles + (size_t)p.stages * stage_bytes : tiles) + quad * (32 * 33);
        int i = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
            const int gi = t / tiles_mn, r = t - gi * tiles_mn, n_idx = r / p.m_tiles, m_idx = r - n_idx * p.m_tiles;
            const Group grp = p.groups[p.order[gi]];
            const int n0 = n_idx * BN;
            const int acc_stage = i % p.nacc;
            const uint32_t aph = (uint32_t)(i / p.nacc) & 1u;
            if (p.bias) {                               // this tile's bias columns, staged while the main loop runs
                __syncwarp();
                for (int c = lane; c < p.epi_cols; c += 32) bias_s[c] = __ldg(p.bias + (n0 + c) % p.bias_mod);
                __syncwarp();
            }
            ptx::mbar_wait(&ctl.acc_full[acc_stage], aph);
            ptx::tc_fence_after();
          for (int sub = 0; sub < MSUB; ++sub) {
            const long long m128 = (long long)m_idx * MSUB + sub;           // index of this 128-row sub-tile
            const long long m = m128 * kBM + row;
            const uint32_t taddr = tmem_base + (uint32_t)(acc_stage * p.acc_stride + sub * p.sub_stride) + ((uint32_t)(quad * 32) << 16);
            __nv_bfloat16 *orow = static_cast<__nv_bfloat16 *>(p.out) + m * p.ld_out + grp.out_col_off + n0;
            float *prow = p.partial ? p.partial + (size_t)grp.part * p.partial_stride + m * p.ld_out + grp.out_col_off + n0 : nullptr;
            const float oscale = (p.out_scale && !prow) ? __ldg(p.out_scale) : 1.f;
            for (int c0 = 0; c0 < p.epi_cols; c0 += 32) {
                float v[32];
                if (c0 + 32 <= p.epi_cols) {
                    ptx::tmem_ld_32(taddr + (uint32_t)c0, v);
                } else {                                    // 16-column tail (epi_cols is a multiple of 16)
                    float tt[16];
                    ptx::tmem_ld_16(taddr + (uint32_t)c0, tt);
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v[j] = tt[j]; v[16 + j] = 0.f; }
                }
                if (prow) {                                 // split-K partial: raw fp32 accumulators
                    if (m < p.m_total) {
                        const int nvec = min(32, p.epi_cols - c0) / 4;
                        float4 *dst = reinterpret_cast<float4 *>(prow + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nvec) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    continue;
                }
                const int ncol = min(32, p.epi_cols - c0);
                // ONE uniform branch per chunk: per-element tests (`if (bias && j < ncol)`) compile to 32 divergence
                // regions and cost ~1.5 k cycles per chunk -- a quarter of a tile's life on the narrow layers
                if (p.bias) {
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias_s + c0);      // broadcast reads
#pragma unroll
                    for (int j = 0; j < 8; ++j) {       // columns past a 16-wide tail pick up stale shared memory: never stored
                        const float4 b = b4[j];
                        v[4 * j] = fmaf(v[4 * j], oscale, b.x);
                        v[4 * j + 1] = fmaf(v[4 * j + 1], oscale, b.y);
                        v[4 * j + 2] = fmaf(v[4 * j + 2], oscale, b.z);
                        v[4 * j + 3] = fmaf(v[4 * j + 3], oscale, b.w);
                    }
                } else if (p.out_scale) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= oscale;
                }
                if (p.stats) {
                    // column sums over this warp's 32 rows through a padded shared-memory transpose; rows past the end
                    // of the tensor count as zero
                    const bool live = m < p.m_total;
#pragma unroll
                    for (int j = 0; j < 32; ++j) scr[lane * 33 + j] = live ? v[j] : 0.f;
                    __syncwarp();
                    float sum = 0.f, sq = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        const float tv = scr[rr * 33 + lane];
                        sum += tv;
                        sq = fmaf(tv, tv, sq);
                    }
                    __syncwarp();
                    if (lane < ncol) {
                        const long long grow = m128 * 4 + quad;
                        *reinterpret_cast<float2 *>(p.stats + (grow * p.ld_out + grp.out_col_off + n0 + c0 + lane) * 2) = make_float2(sum, sq);
                    }
                }
                if (m < p.m_total) {
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float a = v[j], b = v[j + 1];
                        a = a > 0.f ? a : a * p.slope;
                        b = b > 0.f ? b : b * p.slope;
                        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        packed[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(orow + c0);
                    dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                    if (ncol > 16) {
                        dst[2] = make_uint4(packed[8], packed[9], packed[10], packed[11]);
                        dst[3] = make_uint4(packed[12], packed[13], packed[14], packed[15]);
                    }
                }
            }
          }
            // this warp is done with the accumulator stage: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl.acc_empty[acc_stage]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// Sum of the split-K partials written by tap_gemm_kernel: out[m][col] = bf16(sum_{s < nsplit[col / block_cols]}
// partial[s][m][col]), fixed order.  One thread per 8 consecutive columns.
struct SplitReduceParams {
    const float *partial;
    const float *scale;                 // device scalar (may be null) applied to the sum
    __nv_bfloat16 *out;
    long long partial_stride, vecs;     // vecs = rows * ld / 8
    int ld, block_cols;
    int nsplit[kMaxGroups];
};
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitReduceParams p)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= p.vecs) return;
    const int col = (int)((i * 8) % p.ld);
    const int n = p.nsplit[col / p.block_cols];
    const float4 *src = reinterpret_cast<const float4 *>(p.partial + i * 8);
    float4 a = src[0], b = src[1];
    for (int s = 1; s < n; ++s) {
        const float4 *q = reinterpret_cast<const float4 *>(p.partial + (size_t)s * p.partial_stride + i * 8);
        const float4 c = q[0], d = q[1];
        a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
        b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    if (p.scale) {
        const float sc = __ldg(p.scale);
        a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
        b.x *= sc; b.y *= sc; b.z *= sc; b.w *= sc;
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t *>(&h0); o.y = *reinterpret_cast<uint32_t *>(&h1);
    o.z = *reinterpret_cast<uint32_t *>(&h2); o.w = *reinterpret_cast<uint32_t *>(&h3);
    *reinterpret_cast<uint4 *>(p.out + i * 8) = o;
}

// -------------------------------------------------------------------------------------------------
// MN-major kernel: wgrad.  grid = (Cin/128, Cout/BN, pairs * splits).  K runs over positions.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedCtl ctl;
    uint8_t *tiles = align_1024(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = p.BN;
    constexpr int kPos = 64;                                   // positions (K) per stage
    const uint32_t box_bytes = kPos * 128;                     // one (64 channels x 64 positions) box
    const uint32_t a_bytes = 2 * box_bytes, b_bytes = (uint32_t)(BN / 64) * box_bytes, stage_bytes = a_bytes + b_bytes;
    const int pair_idx = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
    const WgradParams::Pair pr = p.pairs[pair_idx];
    const int ci0 = blockIdx.x * kBM, co0 = blockIdx.y * BN;
    const int per = (p.pos_tiles + p.splits - 1) / p.splits;
    const int t_begin = split * per, t_end = min(p.pos_tiles, t_begin + per);
    const int total_iters = max(0, t_end - t_begin);

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&p.tmX);
        ptx::prefetch_tensormap(&p.tmDY);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&ctl.full[s], 1);
            ptx::mbar_init(&ctl.empty[s], 1);
        }
        ptx::mbar_init(&ctl.acc_ready, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&ctl.tmem_base, tmem_cols_for(BN));
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl.tmem_base;

    if (total_iters == 0 && warp >= 2) {                       // empty K range: this split's partial tile is all zero
        const int row = (warp & 3) * 32 + lane;
        if (ci0 + row < p.cin) {
            float *orow = p.dw + (((size_t)split * p.num_taps_total + pr.tap_flat) * p.cin + ci0 + row) * p.cout + co0;
            for (int c = 0; c < BN; c += 4) *reinterpret_cast<float4 *>(orow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (total_iters > 0) {
        // one producer thread and one MMA thread: keep their per-iteration instruction streams short (no divisions --
        // the position of the next 64-position tile is a mixed-radix increment -- shared-window addresses, descriptor adds)
        const uint32_t tiles_a = ptx::smem_u32(tiles), full_a = ptx::smem_u32(&ctl.full[0]), empty_a = ptx::smem_u32(&ctl.empty[0]);
        const uint32_t nstages = (uint32_t)p.stages;
        if (warp == 0) {
            if (ptx::elect_one()) {
                const uint32_t X = (uint32_t)p.X, Y = (uint32_t)p.Y, Z = (uint32_t)p.Z;
                uint32_t q = (uint32_t)t_begin * kPos;          // < B * X * Y * Z, which fits an int (m_total)
                uint32_t x0 = q % X; q /= X;
                uint32_t y0 = q % Y; q /= Y;
                uint32_t z0 = q % Z;
                uint32_t b0 = q / Z;
                const uint32_t dx = kPos % X, c1 = kPos / X, dy = c1 % Y, c2 = c1 / Y, dz = c2 % Z, db = c2 / Z;
                const int cA0 = ci0, cA1 = ci0 + 64, cB0 = pr.dy_c_off + co0, nb = BN / 64;
                uint32_t s = 0, ph = 1;
                for (int it = 0; it < total_iters; ++it) {
                    const uint32_t fb = full_a + 8u * s, dst = tiles_a + s * stage_bytes;
                    ptx::mbar_wait_a(empty_a + 8u * s, ph);
                    ptx::mbar_arrive_expect_tx_a(fb, stage_bytes);
                    const int xs = (int)x0 + pr.sx, ys = (int)y0 + pr.sy, zs = (int)z0 + pr.sz;
                    ptx::tma_load_5d_a(dst, &p.tmX, fb, cA0, xs, ys, zs, (int)b0);
                    ptx::tma_load_5d_a(dst + box_bytes, &p.tmX, fb, cA1, xs, ys, zs, (int)b0);
                    for (int j = 0; j < nb; ++j)
                        ptx::tma_load_5d_a(dst + a_bytes + j * box_bytes, &p.tmDY, fb, cB0 + j * 64, (int)x0, (int)y0, (int)z0, (int)b0);
                    if (++s == nstages) { s = 0; ph ^= 1u; }
                    x0 += dx; uint32_t carry = x0 >= X ? 1u : 0u; x0 -= carry ? X : 0u;
                    y0 += dy + carry; carry = y0 >= Y ? 1u : 0u; y0 -= carry ? Y : 0u;
                    z0 += dz + carry; carry = z0 >= Z ? 1u : 0u; z0 -= carry ? Z : 0u;
                    b0 += db + carry;
                }
            }
        } else if (warp == 1) {
            if (ptx::elect_one()) {
                const uint32_t idesc = ptx::idesc_bf16(kBM, BN, true, true);
                // MN-major: LBO = distance between 64-channel groups (one box), SBO = 8 positions
                const uint64_t desc0 = ptx::smem_desc_sw128(0, box_bytes, 1024);
                const uint32_t stage_units = stage_bytes >> 4, a_units = a_bytes >> 4, tiles_units = tiles_a >> 4;
                uint32_t s = 0, ph = 0;
                for (int it = 0; it < total_iters; ++it) {
                    ptx::mbar_wait_a(full_a + 8u * s, ph);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = desc0 + (uint64_t)(tiles_units + s * stage_units);
                    const uint64_t b_desc = a_desc + a_units;
                    // 16 positions = 16 rows x 128 B = 2048 B (>>4 = 128)
                    ptx::umma_bf16(tmem_base, a_desc, b_desc, idesc, it != 0);
                    ptx::umma_bf16(tmem_base, a_desc + 128, b_desc + 128, idesc, true);
                    ptx::umma_bf16(tmem_base, a_desc + 256, b_desc + 256, idesc, true);
                    ptx::umma_bf16(tmem_base, a_desc + 384, b_desc + 384, idesc, true);
                    ptx::umma_commit_a(empty_a + 8u * s);
                    if (++s == nstages) { s = 0; ph ^= 1u; }
                }
                ptx::umma_commit(&ctl.acc_ready);
            }
        } else {
            ptx::mbar_wait(&ctl.acc_ready, 0);
            ptx::tc_fence_after();
            const int quad = warp & 3;
            const int row = quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
            float *orow = p.dw + (((size_t)split * p.num_taps_total + pr.tap_flat) * p.cin + ci0 + row) * p.cout + co0;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                ptx::tmem_ld_16(taddr + (uint32_t)c0, v);
                if (ci0 + row < p.cin) {
                    float4 *dst = reinterpret_cast<float4 *>(orow + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols_for(BN));
}

// -------------------------------------------------------------------------------------------------
// Weight packing: torch ConvTranspose layout (Cin, Cout, k^d) fp32  <->  GEMM operand layouts
// -------------------------------------------------------------------------------------------------
// Channel permutation of the projection operand (HG_PROJ): GEMM K index k' = y*C + c pairs with the
// reference's folded input channel c*S + (S-1-y) (hologan_generator.py:130-133).  perm_s == 0: identity.
__device__ __forceinline__ int torch_cin(int ci, int perm_c, int perm_s)
{
    if (perm_s == 0) return ci;
    const int y = ci / perm_c, c = ci - y * perm_c;
    return c * perm_s + (perm_s - 1 - y);
}

// Bricks of kBrickCi x kBrickCo x T weights go through shared memory so that both sides of the transposition
// are coalesced: the torch layout (Cin, Cout, T) is read / written in runs of kBrickCo*T floats, the GEMM
// layouts in runs of kBrickCo (or kBrickCi) consecutive elements.  The tap index is padded to an odd pitch
// (bank-conflict-free for both access directions).
constexpr int kBrickCi = 16;
constexpr int kBrickCo = 32;
__host__ __device__ inline int brick_tpitch(int taps) { return taps | 1; }
__host__ __device__ inline int brick_row_pitch(int taps) { return kBrickCo * brick_tpitch(taps) + 1; }

// torch (Cin, Cout, T) fp32 -> w_fwd[t][co][ci] (B operand of forward: rows = Cout, K = Cin) and
// w_dgrad[t][ci][co] (rows = Cin, K = Cout), bf16, written as bf16x2 pairs.  T is a template parameter (1, 16,
// 27 on the hot path) so the index arithmetic has no runtime divisions; global reads are float4 with four
// independent loads in flight per thread.
// `divisor` (device pointer, may be null): every weight is divided by *divisor first -- the spectral norm sigma of the
// discriminator's convolutions (W / sigma, core/models/hologan_discriminator.py:15), so no normalised fp32 copy exists.
// BCI = input channels per brick: 16 normally, 4 for small weights (more bricks than SMs; see wgrad_reduce_kernel).
template <int T, int BCI>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ w_fwd,
                                                          __nv_bfloat16 *__restrict__ w_dgrad, int cin, int cout, int perm_c,
                                                          int perm_s, const float *__restrict__ divisor)
{
    extern __shared__ float brick[];                    // [kBrickCi][kBrickCo][tpitch] (+1 pad per ci row)
    constexpr int kBrickCi = BCI;
    constexpr int row_len = kBrickCo * T, tp = T | 1, pitch = kBrickCo * tp + 1, vec_per_row = row_len / 4;
    const int ci0 = blockIdx.x * kBrickCi, co0 = blockIdx.y * kBrickCo;
    const float div = divisor ? __ldg(divisor) : 1.f;
#pragma unroll 4
    for (int i = threadIdx.x; i < kBrickCi * vec_per_row; i += 256) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci0 + r < cin)      // cout % kBrickCo == 0 (checked on the host): the whole row is in range
            v = __ldg(reinterpret_cast<const float4 *>(w + ((size_t)torch_cin(ci0 + r, perm_c, perm_s) * cout + co0) * T + c));
        float e[4] = {v.x, v.y, v.z, v.w};
        if (divisor) {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = __fdiv_rn(e[j], div);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = (c + j) / T, t = (c + j) - co * T;
            brick[r * pitch + co * tp + t] = e[j];
        }
    }
    __syncthreads();
    if (w_dgrad) {      // [t][ci][co]: a half-warp writes one (t, ci) row of 32 co as 16 bf16x2
        for (int i = threadIdx.x; i < T * kBrickCi * (kBrickCo / 2); i += 256) {
            const int cp = i % (kBrickCo / 2), ci = (i / (kBrickCo / 2)) % kBrickCi, t = i / (kBrickCi * (kBrickCo / 2));
            if (ci0 + ci < cin) {
                const float *src = brick + ci * pitch + (2 * cp) * tp + t;
                *reinterpret_cast<__nv_bfloat162 *>(w_dgrad + ((size_t)t * cin + ci0 + ci) * cout + co0 + 2 * cp) =
                    __floats2bfloat162_rn(src[0], src[tp]);
            }
        }
    }
    if (w_fwd) {        // [t][co][ci]: 8 threads write one (t, co) row of 16 ci as bf16x2
        for (int i = threadIdx.x; i < T * kBrickCo * (kBrickCi / 2); i += 256) {
            const int cp = i % (kBrickCi / 2), co = (i / (kBrickCi / 2)) % kBrickCo, t = i / (kBrickCo * (kBrickCi / 2));
            if (ci0 + 2 * cp < cin) {
                const float *src = brick + (2 * cp) * pitch + co * tp + t;
                *reinterpret_cast<__nv_bfloat162 *>(w_fwd + ((size_t)t * cout + co0 + co) * cin + ci0 + 2 * cp) =
                    __floats2bfloat162_rn(src[0], src[pitch]);
            }
        }
    }
}

// Sum the split-K partials [split][t][ci][co] and write the torch layout (Cin, Cout, T) fp32.  A thread reads
// float4 (4 co) of one (t, ci) row for every split, two rows in flight; writes are float4 runs of the torch
// rows.  accumulate != 0: dw += result (lets the caller target a live .grad buffer).  Fixed summation order.
// BCI = input channels per brick: 16 normally, 2 for small weights (a grid of Cin/16 x Cout/32 bricks would leave most
// SMs idle and serialise `splits` dependent load rounds in a handful of CTAs: 16 CTAs x 11 splits took ~25 us for the
// discriminator's first block, profiles/r02b_profile_d_tcgen05.txt).
template <int T, int BCI>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int cin,
                                                           int cout, int splits, int perm_c, int perm_s, int accumulate)
{
    extern __shared__ float brick[];
    constexpr int row_len = kBrickCo * T, tp = T | 1, pitch = kBrickCo * tp + 1, vec_per_row = row_len / 4;
    constexpr int kBrickCi = BCI;
    const int ci0 = blockIdx.x * kBrickCi, co0 = blockIdx.y * kBrickCo;
    const size_t split_stride = (size_t)T * cin * cout;
#pragma unroll 2
    for (int i = threadIdx.x; i < T * kBrickCi * (kBrickCo / 4); i += 256) {
        const int cq = i % (kBrickCo / 4), ci = (i / (kBrickCo / 4)) % kBrickCi, t = i / (kBrickCi * (kBrickCo / 4));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci0 + ci < cin) {
            const float *src = partial + ((size_t)t * cin + ci0 + ci) * cout + co0 + cq * 4;
            int sp = 0;
            for (; sp + 4 <= splits; sp += 4) {
                const float4 a0 = __ldg(reinterpret_cast<const float4 *>(src + (size_t)sp * split_stride));
                const float4 a1 = __ldg(reinterpret_cast<const float4 *>(src + (size_t)(sp + 1) * split_stride));
                const float4 a2 = __ldg(reinterpret_cast<const float4 *>(src + (size_t)(sp + 2) * split_stride));
                const float4 a3 = __ldg(reinterpret_cast<const float4 *>(src + (size_t)(sp + 3) * split_stride));
                v.x += a0.x; v.y += a0.y; v.z += a0.z; v.w += a0.w;
                v.x += a1.x; v.y += a1.y; v.z += a1.z; v.w += a1.w;
                v.x += a2.x; v.y += a2.y; v.z += a2.z; v.w += a2.w;
                v.x += a3.x; v.y += a3.y; v.z += a3.z; v.w += a3.w;
            }
            for (; sp < splits; ++sp) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src + (size_t)sp * split_stride));
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
        }
        float *d = brick + ci * pitch + (cq * 4) * tp + t;
        d[0] = v.x; d[tp] = v.y; d[2 * tp] = v.z; d[3 * tp] = v.w;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = threadIdx.x; i < kBrickCi * vec_per_row; i += 256) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
        if (ci0 + r < cin) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = (c + j) / T, t = (c + j) - co * T;
                e[j] = brick[r * pitch + co * tp + t];
            }
            float4 *dst = reinterpret_cast<float4 *>(dw + ((size_t)torch_cin(ci0 + r, perm_c, perm_s) * cout + co0) * T + c);
            float4 o = make_float4(e[0], e[1], e[2], e[3]);
            if (accumulate) {
                const float4 old = *dst;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *dst = o;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Host side
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// bf16 channels-last tensor (C, X, Y, Z, B) with box (64, bx, by, bz, bb), 128B swizzle, zero OOB fill
static int make_map_5d(CUtensorMap *m, const void *base, long long C, int X, int Y, int Z, int B, int bx, int by, int bz,
                       int bb)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(HG_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point unavailable");
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * X, (cuuint64_t)C * 2 * X * Y, (cuuint64_t)C * 2 * X * Y * Z};
    cuuint32_t box[5] = {(cuuint32_t)kBK, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bb};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ERR_INVALID_ARG, "cuTensorMapEncodeTiled(5d) failed with %d", (int)r);
    return HG_OK;
}

static int make_map_2d(CUtensorMap *m, const void *base, long long cols, long long rows, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(HG_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ERR_INVALID_ARG, "cuTensorMapEncodeTiled(2d) failed with %d", (int)r);
    return HG_OK;
}

// split `count` positions into a box over (X, Y, Z, B): fill x, then y, then z, then batch
static bool box_for(int count, int X, int Y, int Z, int &bx, int &by, int &bz, int &bb)
{
    bx = by = bz = bb = 1;
    int rem = count;
    bx = rem < X ? rem : X; if (X % bx) return false; rem /= bx; if (bx < X) return rem == 1;
    by = rem < Y ? rem : Y; if (Y % by) return false; rem /= by; if (by < Y) return rem == 1;
    bz = rem < Z ? rem : Z; if (Z % bz) return false; rem /= bz; if (bz < Z) return rem == 1;
    bb = rem;
    return bb <= 256;
}

static int pick_bn(int n)
{
    if (n % 256 == 0) return 256;
    if (n % 128 == 0) return 128;
    if (n % 64 == 0) return 64;
    if (n % 32 == 0) return 32;
    if (n % 16 == 0) return 16;
    return 0;
}

static int pick_stages(int stage_bytes, int iters, int budget = kSmemBudget)
{
    int s = (budget - 1024) / stage_bytes;
    if (s > kMaxStages) s = kMaxStages;
    if (s > iters) s = iters < 1 ? 1 : iters;
    return s;
}

// per-dimension tap table of a stride-2 transposed convolution: for output parity p, the kernel
// indices k and input shifts s with  out[2*i + p] += in[i + s] * w[k]
struct DimTaps { int n[2]; int k[2][3]; int s[2][3]; int classes; };
static DimTaps dim_taps(int kernel)
{
    DimTaps d{};
    if (kernel == 5) {          // k5 s2 p2 op1: o = 2i - 2 + k.  Its dgrad is Conv2d(k5, s2, p2) -- the discriminator's
        d.classes = 2;          // convolution (core/models/hologan_discriminator.py:12) -- on a space-to-depth input.
        d.n[0] = 3; d.k[0][0] = 0; d.s[0][0] = 1; d.k[0][1] = 2; d.s[0][1] = 0; d.k[0][2] = 4; d.s[0][2] = -1;
        d.n[1] = 2; d.k[1][0] = 1; d.s[1][0] = 1; d.k[1][1] = 3; d.s[1][1] = 0;
    } else if (kernel == 4) {   // k4 s2 p1: o = 2i - 1 + k
        d.classes = 2;
        d.n[0] = 2; d.k[0][0] = 1; d.s[0][0] = 0; d.k[0][1] = 3; d.s[0][1] = -1;
        d.n[1] = 2; d.k[1][0] = 0; d.s[1][0] = 1; d.k[1][1] = 2; d.s[1][1] = 0;
    } else if (kernel == 3) {   // k3 s2 p1 op1: o = 2i - 1 + k
        d.classes = 2;
        d.n[0] = 1; d.k[0][0] = 1; d.s[0][0] = 0;
        d.n[1] = 2; d.k[1][0] = 0; d.s[1][0] = 1; d.k[1][1] = 2; d.s[1][1] = 0;
    } else {                    // k1 s1: plain per-pixel GEMM
        d.classes = 1;
        d.n[0] = 1; d.k[0][0] = 0; d.s[0][0] = 0;
    }
    return d;
}

struct ConvShape {
    int batch, cin, cout, ndim, size, kernel;
    int X, Y, Z, P, taps;
};
static int conv_shape(ConvShape &c, const char *who)
{
    HG_REQUIRE(c.batch > 0 && c.cin > 0 && c.cout > 0 && c.size > 0, HG_ERR_INVALID_ARG, "%s: dims must be positive", who);
    HG_REQUIRE((c.ndim == 2 && (c.kernel == 4 || c.kernel == 5 || c.kernel == 1)) || (c.ndim == 3 && c.kernel == 3), HG_ERR_UNSUPPORTED,
               "%s: supported: ndim 2 with kernel 4 (s2,p1), 5 (s2,p2,op1) or 1, ndim 3 with kernel 3 (s2,p1,op1)", who);
    c.X = c.size; c.Y = c.size; c.Z = c.ndim == 3 ? c.size : 1;
    c.P = c.kernel == 1 ? 1 : (c.ndim == 3 ? 8 : 4);
    c.taps = c.kernel == 1 ? 1 : (c.ndim == 3 ? 27 : c.kernel * c.kernel);
    return HG_OK;
}

template <typename Fn>
static void for_each_class_tap(const ConvShape &c, Fn fn)
{
    const DimTaps d = dim_taps(c.kernel);
    const int nz = c.ndim == 3 ? d.classes : 1;
    for (int pz = 0; pz < nz; ++pz)
        for (int py = 0; py < d.classes; ++py)
            for (int px = 0; px < d.classes; ++px) {
                const int cls = (pz * d.classes + py) * d.classes + px;
                const int tz_n = c.ndim == 3 ? d.n[pz] : 1;
                for (int tz = 0; tz < tz_n; ++tz)
                    for (int ty = 0; ty < d.n[py]; ++ty)
                        for (int tx = 0; tx < d.n[px]; ++tx) {
                            const int kz = c.ndim == 3 ? d.k[pz][tz] : 0, ky = d.k[py][ty], kx = d.k[px][tx];
                            const int sz = c.ndim == 3 ? d.s[pz][tz] : 0, sy = d.s[py][ty], sx = d.s[px][tx];
                            const int flat = c.ndim == 3 ? (kz * c.kernel + ky) * c.kernel + kx : ky * c.kernel + kx;
                            fn(cls, flat, sx, sy, sz);
                        }
            }
}

// Launch shape of the K-major kernel.  "single": one CTA per SM with a deep ring (kSmemBudget) and up to 512
// accumulator columns.  "dual": two co-resident CTAs per SM (kDualSmemBudget each, <= 256 columns each), so that a
// CTA's prologue (tensor-map fetch, first TMA round trip) and epilogue (TMEM drain + stores) overlap its
// neighbour's main loop -- the ~11 k cycles of fixed cost per CTA are half the life of a CTA on the narrow layers
// (profiles/r01c_ncu_full_summary.txt).  Option TAPGEMM_DUAL = 0 / 1 overrides the choice (tuning only).
static int dual_override() { return option(kOptTapGemmDual); }
static bool want_dual(int stage_bytes, int iters_per_cta, int epi_cols)
{
    if (epi_cols > 256 || 2 * stage_bytes + 1024 > kDualSmemBudget) return false;
    const int ov = dual_override();
    if (ov >= 0) return ov == 1;
    return (long long)stage_bytes * iters_per_cta <= kDualMaxBytes;
}

// Regroup the (class, tap) list of every group into shift-major items (struct Item).  Taps of one group: `acc` is the class
// index inside the group, accumulator columns acc * BN.  The zero-shift bucket goes first: every class has a tap with
// shift 0 (both the k3 and the k4 stride-2 tables), so its MMAs initialise all accumulators and everything after
// accumulates.  Kept general anyway: a segment never mixes fresh and started accumulators.
static void build_shift_items(TapGemmParams &p, int cpc)
{
    const int max_boxes = std::min(std::min(kItemBoxes, option(kOptTapGemmShareA)), 256 / p.BN);
    if (max_boxes < 2) return;
    int nitems = 0, item_rows = 0;
    Item items[kMaxTaps];
    Group groups[kMaxGroups];
    for (int g = 0; g < p.num_groups; ++g) {
        const Group &grp = p.groups[g];
        groups[g] = grp;
        groups[g].tap_begin = nitems;
        bool used[kMaxTaps] = {};
        uint32_t started = 0;
        for (int pass = 0; pass < 2; ++pass)                    // pass 0: the zero shift; pass 1: the rest in tap order
            for (int t0 = 0; t0 < grp.tap_count; ++t0) {
                const Tap &lead = p.taps[grp.tap_begin + t0];
                if (used[t0] || (pass == 0 && (lead.sx | lead.sy | lead.sz) != 0)) continue;
                // bucket: all taps of the group with lead's shift, by class
                int idx[kMaxTaps], n = 0;
                for (int t = t0; t < grp.tap_count; ++t) {
                    const Tap &tp = p.taps[grp.tap_begin + t];
                    if (!used[t] && tp.sx == lead.sx && tp.sy == lead.sy && tp.sz == lead.sz && tp.a_c_off == lead.a_c_off) {
                        used[t] = true;
                        idx[n++] = t;
                    }
                }
                std::sort(idx, idx + n, [&](int a, int b) { return p.taps[grp.tap_begin + a].acc < p.taps[grp.tap_begin + b].acc; });
                for (int b0 = 0; b0 < n; b0 += max_boxes) {
                    Item it{};
                    it.sx = lead.sx; it.sy = lead.sy; it.sz = lead.sz; it.a_c_off = lead.a_c_off;
                    it.nb = (uint8_t)std::min(max_boxes, n - b0);
                    for (int j = 0; j < kItemBoxes; ++j) it.b_row[j] = p.taps[grp.tap_begin + idx[b0 + std::min<int>(j, it.nb - 1)]].b_row_off;
                    int j = 0;
                    while (j < it.nb) {                          // maximal runs of adjacent classes with one started state
                        const int acc0 = p.taps[grp.tap_begin + idx[b0 + j]].acc;
                        const uint32_t st0 = (started >> acc0) & 1u;
                        int run = 1;
                        while (j + run < it.nb && p.taps[grp.tap_begin + idx[b0 + j + run]].acc == acc0 + run &&
                               ((started >> (acc0 + run)) & 1u) == st0)
                            ++run;
                        Item::Seg &sg = it.seg[it.nseg++];
                        sg.col = (uint16_t)(acc0 * p.BN);
                        sg.b_units = (uint16_t)((j * p.BN * 128) >> 4);
                        sg.idesc = ptx::idesc_bf16(kBM, run * p.BN, false, false);
                        sg.first = st0 ? 0u : 1u;
                        for (int r = 0; r < run; ++r) started |= 1u << (acc0 + r);
                        j += run;
                    }
                    item_rows = std::max(item_rows, it.nb * p.BN);
                    items[nitems++] = it;
                }
            }
        groups[g].tap_count = nitems - groups[g].tap_begin;
    }
    for (int g = 0; g < p.num_groups; ++g) p.groups[g] = groups[g];
    for (int i = 0; i < nitems; ++i) p.items[i] = items[i];
    p.item_rows = item_rows;
}

static int launch_tap_gemm(TapGemmParams &p, int m_tiles, int n_tiles, cudaStream_t st, const char *who)
{
    int max_iters = 0;
    for (int g = 0; g < p.num_groups; ++g) max_iters = p.groups[g].tap_count * p.k_chunks > max_iters ? p.groups[g].tap_count * p.k_chunks : max_iters;
    // two 128-row sub-tiles per CTA tile for the wide GEMMs (one 256-column accumulator each = all 512 TMEM columns), when
    // that still leaves at least ~1.5 tiles per SM
    p.m_sub = 1;
    if (option(kOptTapGemmMsub) != 0 && p.item_rows == 0 && p.BN == 256 && p.epi_cols == 256 && !p.partial &&
        (long long)(m_tiles / 2) * n_tiles * p.num_groups >= (3ll * sm_count()) / 2) {
        p.m_sub = 2;
        m_tiles = (m_tiles + 1) / 2;
    }
    p.sub_stride = p.m_sub == 2 ? 256 : 0;
    const int stage_bytes = p.item_rows > 0 ? kBM * 128 + p.item_rows * 128 : p.m_sub * kBM * 128 + p.BN * 128;
    p.m_tiles = m_tiles; p.n_tiles = n_tiles;
    // longest K loops first (tiles of one group are equally long)
    for (int g = 0; g < p.num_groups; ++g) p.order[g] = (uint8_t)g;
    for (int a = 1; a < p.num_groups; ++a)
        for (int b = a; b > 0 && p.groups[p.order[b]].tap_count > p.groups[p.order[b - 1]].tap_count; --b) {
            const uint8_t tmp = p.order[b]; p.order[b] = p.order[b - 1]; p.order[b - 1] = tmp;
        }
    const long long total = (long long)m_tiles * n_tiles * p.num_groups;
    const int scratch = p.stats ? 4 * 32 * 33 * (int)sizeof(float) : 0;
    // persistent (one CTA per SM walks the tiles, two accumulator stages): auto = the wide tiles (BN == 256: the projection,
    // block3, block4 dgrad), whose one-tile launches get two shallow stages per co-resident CTA or no epilogue overlap at
    // all; the narrow layers keep two co-resident one-tile CTAs (profiles/r02z_tap_gemm_options.txt)
    const int popt = option(kOptTapGemmPersistent);
    const bool persistent = (popt > 0 || (popt < 0 && p.BN == 256 && p.item_rows == 0 && !p.stats)) && total > sm_count();
    unsigned grid;
    size_t smem;
    if (persistent) {
        // one CTA per SM walks the tile list; two accumulator stages when they fit (<= 256 columns each): the epilogue of
        // a tile overlaps the main loop of the next, and prologue / TMEM allocation / tensor-map fetch happen once
        p.nacc = (p.epi_cols <= 256 && p.m_sub == 1) ? 2 : 1;
        p.acc_stride = p.m_sub == 2 ? 512
                                    : p.epi_cols <= 256 ? (int)(p.epi_cols <= 32 ? 32 : p.epi_cols <= 64 ? 64 : p.epi_cols <= 128 ? 128 : 256) : 512;
        p.stats_scratch = p.stats ? 1 : 0;
        p.stages = pick_stages(stage_bytes, 1 << 20, kSmemBudget - scratch);
        smem = (size_t)p.stages * stage_bytes + 1024 + scratch;
        grid = (unsigned)sm_count();
    } else {
        const bool dual = p.m_sub == 1 && want_dual(stage_bytes, max_iters, p.epi_cols);
        p.nacc = 1;
        p.acc_stride = p.m_sub == 2 ? 512 : p.epi_cols;
        p.stats_scratch = 0;
        p.stages = pick_stages(stage_bytes, max_iters, dual ? kDualSmemBudget : kSmemBudget);
        if (p.stats && (size_t)p.stages * stage_bytes < (size_t)scratch) p.stages = (scratch + stage_bytes - 1) / stage_bytes;
        smem = (size_t)p.stages * stage_bytes + 1024;
        grid = (unsigned)total;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tap_gemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 2048);
        cudaFuncSetAttribute(tap_gemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 2048);
        cudaFuncSetAttribute(tap_gemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 2048);
        attr_set = true;
    }
    if (p.item_rows > 0) tap_gemm_kernel<1, true><<<grid, kThreads, smem, st>>>(p);
    else if (p.m_sub == 2) tap_gemm_kernel<2, false><<<grid, kThreads, smem, st>>>(p);
    else tap_gemm_kernel<1, false><<<grid, kThreads, smem, st>>>(p);
    return check_launch(who);
}

// N tile of a K-major GEMM: the widest MMA that divides n, narrowed while the grid would leave SMs idle.
static int pick_bn_for_grid(int n, int m_tiles)
{
    int bn = pick_bn(n);
    while (bn > 64 && 4ll * m_tiles * (n / bn) < 3ll * sm_count()) bn /= 2;
    return bn;
}

template <int T, int BCI>
static void launch_wgrad_reduce_t(const float *part, float *dw, int cin, int cout, int splits, int perm_c, int perm_s,
                                  int accumulate, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wgrad_reduce_kernel<T, BCI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    dim3 grid((cin + BCI - 1) / BCI, (cout + kBrickCo - 1) / kBrickCo);
    const size_t smem = (size_t)BCI * brick_row_pitch(T) * sizeof(float);
    wgrad_reduce_kernel<T, BCI><<<grid, 256, smem, st>>>(part, dw, cin, cout, splits, perm_c, perm_s, accumulate);
}

static int launch_wgrad_reduce(const float *part, float *dw, int cin, int cout, int taps, int splits, int perm_c, int perm_s,
                               int accumulate, cudaStream_t st)
{
    // enough bricks to cover the SMs: small weights use 2-channel bricks
    const bool small = (long long)((cin + 15) / 16) * ((cout + kBrickCo - 1) / kBrickCo) < sm_count();
#define HG_REDUCE(T)                                                                                         \
    do {                                                                                                     \
        if (small) launch_wgrad_reduce_t<T, 2>(part, dw, cin, cout, splits, perm_c, perm_s, accumulate, st);  \
        else launch_wgrad_reduce_t<T, 16>(part, dw, cin, cout, splits, perm_c, perm_s, accumulate, st);       \
    } while (0)
    if (taps == 1) HG_REDUCE(1);
    else if (taps == 16) HG_REDUCE(16);
    else if (taps == 25) HG_REDUCE(25);
    else HG_REDUCE(27);
#undef HG_REDUCE
    return check_launch("hg_convt_wgrad(reduce)");
}

}  // namespace hg

using namespace hg;

// D[M,N] = act(A[M,K] @ B[N,K]^T + bias), bf16 in / bf16 out, fp32 accumulate.
extern "C" int hg_gemm_bf16_nt(const void *a, const void *b, const float *bias, void *d, int m, int n, int k, long long ldd,
                               float neg_slope, void *stream)
{
    HG_REQUIRE(a && b && d, HG_ERR_INVALID_ARG, "hg_gemm_bf16_nt: null pointer");
    HG_REQUIRE(m > 0 && n > 0 && k > 0, HG_ERR_INVALID_ARG, "hg_gemm_bf16_nt: dims must be positive");
    const int bn = pick_bn_for_grid(n, (m + kBM - 1) / kBM);
    HG_REQUIRE(k % kBK == 0 && bn >= 16 && ldd % 8 == 0, HG_ERR_UNSUPPORTED,
               "hg_gemm_bf16_nt: need K %% 64 == 0, N %% 16 == 0, ldd %% 8 == 0 (got M=%d N=%d K=%d)", m, n, k);
    TapGemmParams p{};
    int rc = make_map_5d(&p.tmA, a, k, m, 1, 1, 1, kBM, 1, 1, 1);
    if (rc) return rc;
    rc = make_map_2d(&p.tmB, b, k, n, bn);
    if (rc) return rc;
    p.X = m; p.Y = 1; p.Z = 1; p.Bn = 1;
    p.BN = bn; p.epi_cols = bn; p.bias_mod = n; p.k_chunks = k / kBK; p.m_total = m; p.ld_out = ldd; p.out = d; p.bias = bias;
    p.slope = neg_slope;
    p.num_groups = 1;
    p.groups[0] = Group{0, 1, 0, 0};
    p.taps[0] = Tap{0, 0, 0, 0, 0, 0};
    return launch_tap_gemm(p, (m + kBM - 1) / kBM, n / bn, static_cast<cudaStream_t>(stream), "hg_gemm_bf16_nt");
}

static int perm_ok(const char *who, int cin, int perm_c, int perm_s)
{
    HG_REQUIRE(perm_s == 0 || (perm_c > 0 && perm_s > 0 && perm_c * perm_s == cin), HG_ERR_INVALID_ARG,
               "%s: channel permutation (%d, %d) does not factor Cin = %d", who, perm_c, perm_s, cin);
    return HG_OK;
}

static int pack_weight_impl(const float *w, void *w_fwd, void *w_dgrad, int cin, int cout, int taps, int perm_c, int perm_s,
                            const float *divisor, void *stream);

template <int T, int BCI>
static void launch_pack_t(const float *w, __nv_bfloat16 *wf, __nv_bfloat16 *wd, int cin, int cout, int perm_c, int perm_s,
                          const float *divisor, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(pack_weight_kernel<T, BCI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    dim3 grid((cin + BCI - 1) / BCI, cout / kBrickCo);
    const size_t smem = (size_t)BCI * brick_row_pitch(T) * sizeof(float);
    pack_weight_kernel<T, BCI><<<grid, 256, smem, st>>>(w, wf, wd, cin, cout, perm_c, perm_s, divisor);
}

extern "C" int hg_convt_pack_weight(const float *w, void *w_fwd, void *w_dgrad, int cin, int cout, int taps, int perm_c,
                                    int perm_s, void *stream)
{
    return pack_weight_impl(w, w_fwd, w_dgrad, cin, cout, taps, perm_c, perm_s, nullptr, stream);
}

static int pack_weight_impl(const float *w, void *w_fwd, void *w_dgrad, int cin, int cout, int taps, int perm_c, int perm_s,
                            const float *divisor, void *stream)
{
    HG_REQUIRE(w && (w_fwd || w_dgrad), HG_ERR_INVALID_ARG, "hg_convt_pack_weight: null pointer");
    HG_REQUIRE(cin > 0 && cout > 0 && taps > 0 && taps <= 27, HG_ERR_INVALID_ARG, "hg_convt_pack_weight: bad dims");
    int rc = perm_ok("hg_convt_pack_weight", cin, perm_c, perm_s);
    if (rc) return rc;
    HG_REQUIRE(cin % 2 == 0 && cout % kBrickCo == 0, HG_ERR_UNSUPPORTED,
               "hg_convt_pack_weight: Cin must be even and Cout a multiple of %d (got %d, %d)", kBrickCo, cin, cout);
    HG_REQUIRE(taps == 1 || taps == 16 || taps == 25 || taps == 27, HG_ERR_UNSUPPORTED,
               "hg_convt_pack_weight: taps must be 1 (k1), 16 (2-D k4), 25 (2-D k5) or 27 (3-D k3), got %d", taps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16 *wf = static_cast<__nv_bfloat16 *>(w_fwd), *wd = static_cast<__nv_bfloat16 *>(w_dgrad);
    const bool small = (long long)((cin + 15) / 16) * (cout / kBrickCo) < 2ll * sm_count();
#define HG_PACK(T)                                                                                           \
    do {                                                                                                     \
        if (small) launch_pack_t<T, 4>(w, wf, wd, cin, cout, perm_c, perm_s, divisor, st);                    \
        else launch_pack_t<T, 16>(w, wf, wd, cin, cout, perm_c, perm_s, divisor, st);                         \
    } while (0)
    if (taps == 1) HG_PACK(1);
    else if (taps == 16) HG_PACK(16);
    else if (taps == 25) HG_PACK(25);
    else HG_PACK(27);
#undef HG_PACK
    return check_launch("hg_convt_pack_weight");
}

static int convt_fwd_impl(const void *x, const void *w_fwd, const float *bias, const float *out_scale, void *y_s2d, int batch,
                          int cin, int cout, int ndim, int size, int kernel, float neg_slope, void *stream, float *stats = nullptr);

extern "C" int hg_convt_fwd(const void *x, const void *w_fwd, const float *bias, void *y_s2d, int batch, int cin, int cout,
                            int ndim, int size, int kernel, float neg_slope, void *stream)
{
    return convt_fwd_impl(x, w_fwd, bias, nullptr, y_s2d, batch, cin, cout, ndim, size, kernel, neg_slope, stream);
}

// Number of floats hg_convt_fwd_stats writes: 2 per (32-row group, output column), groups padded to whole 128-row tiles.
extern "C" long long hg_convt_stats_floats(int batch, int cout, int ndim, int size, int kernel)
{
    ConvShape c{batch, 64, cout, ndim, size, kernel};
    if (conv_shape(c, "hg_convt_stats_floats")) return -1;
    const long long m_total = (long long)batch * c.X * c.Y * c.Z;
    return ((m_total + kBM - 1) / kBM) * 4 * (long long)c.P * cout * 2;
}

// hg_convt_fwd that also emits the AdaIN statistics partials of its output (TapGemmParams::stats)
extern "C" int hg_convt_fwd_stats(const void *x, const void *w_fwd, const float *bias, void *y_s2d, float *stats, int batch, int cin,
                                  int cout, int ndim, int size, int kernel, float neg_slope, void *stream)
{
    HG_REQUIRE(stats, HG_ERR_INVALID_ARG, "hg_convt_fwd_stats: null stats pointer");
    HG_REQUIRE(cout % 32 == 0, HG_ERR_UNSUPPORTED, "hg_convt_fwd_stats: Cout must be a multiple of 32 (got %d)", cout);
    return convt_fwd_impl(x, w_fwd, bias, nullptr, y_s2d, batch, cin, cout, ndim, size, kernel, neg_slope, stream, stats);
}

static int convt_fwd_impl(const void *x, const void *w_fwd, const float *bias, const float *out_scale, void *y_s2d, int batch,
                          int cin, int cout, int ndim, int size, int kernel, float neg_slope, void *stream, float *stats)
{
    HG_REQUIRE(x && w_fwd && y_s2d, HG_ERR_INVALID_ARG, "hg_convt_fwd: null pointer");
    ConvShape c{batch, cin, cout, ndim, size, kernel};
    int rc = conv_shape(c, "hg_convt_fwd");
    if (rc) return rc;
    const int bn = pick_bn(cout);
    HG_REQUIRE(cin % kBK == 0 && bn >= 16, HG_ERR_UNSUPPORTED, "hg_convt_fwd: need Cin %% 64 == 0 and Cout %% 16 == 0 (got %d, %d)", cin, cout);
    TapGemmParams p{};
    int bx, by, bz, bb;
    HG_REQUIRE(box_for(kBM, c.X, c.Y, c.Z, bx, by, bz, bb), HG_ERR_UNSUPPORTED, "hg_convt_fwd: spatial size %d does not tile into 128-row boxes", size);
    rc = make_map_5d(&p.tmA, x, cin, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    rc = make_map_2d(&p.tmB, w_fwd, cin, (long long)c.taps * cout, bn);
    if (rc) return rc;
    p.X = c.X; p.Y = c.Y; p.Z = c.Z; p.Bn = batch;
    p.BN = bn; p.k_chunks = cin / kBK; p.bias_mod = cout;
    const long long m_total = (long long)batch * c.X * c.Y * c.Z;
    p.m_total = (int)m_total; p.ld_out = (long long)c.P * cout; p.out = y_s2d; p.bias = bias; p.slope = neg_slope;
    p.out_scale = out_scale; p.stats = stats;
    // Narrow layers (Cout <= 128 == one N tile): several parity classes share a CTA, one TMEM accumulator
    // each, so a CTA does classes_per_cta x the work per prologue/epilogue and writes one contiguous run of the
    // s2d output row.  Single launches may fill all 512 TMEM columns; dual launches (see want_dual) stop at 256,
    // and split further while the longest CTA (classes are unequal: 1..2^d taps) would outlast an even share of
    // the whole layer over the 2 x SM slots.
    const int m_tiles = (int)((m_total + kBM - 1) / kBM);
    auto classes_per_cta = [&](int max_cols) {
        int v = 1;
        if (bn == cout && c.P > 1) {
            v = max_cols / cout;
            if (v > c.P) v = c.P;
            if (v < 1) v = 1;
            while (c.P % v) --v;
        }
        return v;
    };
    auto longest_taps = [&](int per) {          // taps of the heaviest group when `per` consecutive classes share a CTA
        int best = 0, cur_g = -1, cur_n = 0;
        for_each_class_tap(c, [&](int cls, int, int, int, int) {
            const int g = cls / per;
            if (g != cur_g) { cur_g = g; cur_n = 0; }
            ++cur_n;
            if (cur_n > best) best = cur_n;
        });
        return best;
    };
    int cpc = classes_per_cta(512);
    {
        int cd = classes_per_cta(256);
        const int stage_bytes = kBM * 128 + bn * 128;
        const double fair = (double)m_tiles * (cout / bn) * c.taps / (2.0 * sm_count());     // taps per slot, evenly spread
        while (cd > 1 && longest_taps(cd) > fair && c.P % (cd / 2) == 0) cd /= 2;
        if (want_dual(stage_bytes, longest_taps(cd) * (cin / kBK), cd * bn)) cpc = cd;
    }
    p.epi_cols = cpc * bn;
    p.num_groups = c.P / cpc;
    int ntap = 0;
    for (int g = 0; g < p.num_groups; ++g) p.groups[g] = Group{0, 0, g * cpc * cout, 0};
    int cur = -1;
    for_each_class_tap(c, [&](int cls, int flat, int sx, int sy, int sz) {
        const int g = cls / cpc;
        if (cls != cur) {
            cur = cls;
            if (cls % cpc == 0) p.groups[g].tap_begin = ntap;
        }
        p.taps[ntap] = Tap{(int16_t)sx, (int16_t)sy, (int16_t)sz, (int16_t)(cls % cpc), 0, flat * cout};
        p.groups[g].tap_count++;
        ntap++;
    });
    if (cpc > 1 && option(kOptTapGemmShareA) != 0) build_shift_items(p, cpc);
    return launch_tap_gemm(p, m_tiles, cout / bn, static_cast<cudaStream_t>(stream), "hg_convt_fwd");
}

extern "C" int hg_convt_dgrad(const void *dy_s2d, const void *w_dgrad, void *dx, int batch, int cin, int cout, int ndim,
                              int size, int kernel, void *stream)
{
    HG_REQUIRE(dy_s2d && w_dgrad && dx, HG_ERR_INVALID_ARG, "hg_convt_dgrad: null pointer");
    ConvShape c{batch, cin, cout, ndim, size, kernel};
    int rc = conv_shape(c, "hg_convt_dgrad");
    if (rc) return rc;
    const long long m_total = (long long)batch * c.X * c.Y * c.Z;
    const int m_tiles = (int)((m_total + kBM - 1) / kBM);
    const int bn = pick_bn_for_grid(cin, m_tiles);
    HG_REQUIRE(cout % kBK == 0 && bn >= 16, HG_ERR_UNSUPPORTED, "hg_convt_dgrad: need Cout %% 64 == 0 and Cin %% 16 == 0 (got %d, %d)", cout, cin);
    TapGemmParams p{};
    int bx, by, bz, bb;
    HG_REQUIRE(box_for(kBM, c.X, c.Y, c.Z, bx, by, bz, bb), HG_ERR_UNSUPPORTED, "hg_convt_dgrad: spatial size %d does not tile into 128-row boxes", size);
    rc = make_map_5d(&p.tmA, dy_s2d, (long long)c.P * cout, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    rc = make_map_2d(&p.tmB, w_dgrad, cout, (long long)c.taps * cin, bn);
    if (rc) return rc;
    p.X = c.X; p.Y = c.Y; p.Z = c.Z; p.Bn = batch;
    p.BN = bn; p.epi_cols = bn; p.bias_mod = cin; p.k_chunks = cout / kBK;
    p.m_total = (int)m_total; p.ld_out = cin; p.out = dx; p.bias = nullptr; p.slope = 1.0f;
    p.num_groups = 1;
    int ntap = 0;
    for_each_class_tap(c, [&](int cls, int flat, int sx, int sy, int sz) {
        // forward: Y[2i+p] += X[i+s] W[k]   =>   dX[i'] += dY_s2d[i'-s, class p] W[k]
        p.taps[ntap++] = Tap{(int16_t)-sx, (int16_t)-sy, (int16_t)-sz, 0, cls * cout, flat * cin};   // one accumulator
    });
    p.groups[0] = Group{0, ntap, 0, 0};
    return launch_tap_gemm(p, m_tiles, cin / bn, static_cast<cudaStream_t>(stream), "hg_convt_dgrad");
}

struct WgradPlan { int bn, splits, pairs, pos_tiles; };

static int wgrad_plan(const ConvShape &c, WgradPlan &pl, const char *who)
{
    pl.bn = c.cout % 256 == 0 ? 256 : c.cout % 128 == 0 ? 128 : c.cout % 64 == 0 ? 64 : 0;
    HG_REQUIRE(c.cin % kBM == 0 && pl.bn > 0, HG_ERR_UNSUPPORTED, "%s: need Cin %% 128 == 0 and Cout %% 64 == 0 (got %d, %d)", who,
               c.cin, c.cout);
    const long long positions = (long long)c.batch * c.X * c.Y * c.Z;
    pl.pos_tiles = (int)((positions + 63) / 64);
    pl.pairs = c.taps;                                          // every kernel tap belongs to exactly one parity class
    const int out_tiles = (c.cin / kBM) * (c.cout / pl.bn) * pl.pairs;
    // Split K so that all CTAs are resident at once (two per SM, kWgradSmemBudget each): no tail wave, and a
    // CTA's prologue / epilogue overlaps its neighbour's main loop.
    int splits = (2 * sm_count()) / out_tiles;
    if (splits > pl.pos_tiles / 8) splits = pl.pos_tiles / 8;    // keep >= 8 K-iterations per CTA
    if (splits < 1) splits = 1;
    pl.splits = splits;
    return HG_OK;
}

extern "C" long long hg_convt_wgrad_workspace_bytes(int batch, int cin, int cout, int ndim, int size, int kernel)
{
    ConvShape c{batch, cin, cout, ndim, size, kernel};
    if (conv_shape(c, "hg_convt_wgrad_workspace_bytes")) return -1;
    WgradPlan pl;
    if (wgrad_plan(c, pl, "hg_convt_wgrad_workspace_bytes")) return -1;
    return (long long)pl.splits * c.taps * cin * cout * (long long)sizeof(float);
}

extern "C" int hg_convt_wgrad(const void *x, const void *dy_s2d, float *dw, void *workspace, long long workspace_bytes,
                              int batch, int cin, int cout, int ndim, int size, int kernel, int perm_c, int perm_s,
                              int accumulate, void *stream)
{
    HG_REQUIRE(x && dy_s2d && dw && workspace, HG_ERR_INVALID_ARG, "hg_convt_wgrad: null pointer");
    ConvShape c{batch, cin, cout, ndim, size, kernel};
    int rc = conv_shape(c, "hg_convt_wgrad");
    if (rc) return rc;
    WgradPlan pl;
    rc = wgrad_plan(c, pl, "hg_convt_wgrad");
    if (rc) return rc;
    rc = perm_ok("hg_convt_wgrad", cin, perm_c, perm_s);
    if (rc) return rc;
    HG_REQUIRE(workspace_bytes >= (long long)pl.splits * c.taps * cin * cout * (long long)sizeof(float), HG_ERR_INVALID_ARG,
               "hg_convt_wgrad: workspace smaller than hg_convt_wgrad_workspace_bytes()");
    const int bn = pl.bn;
    WgradParams p{};
    int bx, by, bz, bb;
    HG_REQUIRE(box_for(64, c.X, c.Y, c.Z, bx, by, bz, bb), HG_ERR_UNSUPPORTED, "hg_convt_wgrad: spatial size %d does not tile into 64-position boxes", size);
    rc = make_map_5d(&p.tmX, x, cin, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    rc = make_map_5d(&p.tmDY, dy_s2d, (long long)c.P * cout, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    p.X = c.X; p.Y = c.Y; p.Z = c.Z; p.Bn = batch;
    p.BN = bn; p.cin = cin; p.cout = cout; p.dw = static_cast<float *>(workspace); p.num_taps_total = c.taps;
    p.pos_tiles = pl.pos_tiles;
    int np = 0;
    for_each_class_tap(c, [&](int cls, int flat, int sx, int sy, int sz) {
        p.pairs[np++] = WgradParams::Pair{(int16_t)sx, (int16_t)sy, (int16_t)sz, 0, cls * cout, flat};
    });
    p.num_pairs = np;
    p.splits = pl.splits;
    const int stage_bytes = 2 * 64 * 128 + (bn / 64) * 64 * 128;
    p.stages = pick_stages(stage_bytes, (p.pos_tiles + pl.splits - 1) / pl.splits, kWgradSmemBudget);
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmemBudget);
        attr_set = true;
    }
    dim3 grid(cin / kBM, cout / bn, np * pl.splits);
    wgrad_gemm_kernel<<<grid, kThreads, smem, st>>>(p);
    rc = check_launch("hg_convt_wgrad");
    if (rc) return rc;
    return launch_wgrad_reduce(static_cast<const float *>(workspace), dw, cin, cout, c.taps, pl.splits, perm_c, perm_s, accumulate, st);
}

// -------------------------------------------------------------------------------------------------
// The discriminator's Conv2d(k5, s2, p2) (reference core/models/hologan_discriminator.py:12) on the tap GEMMs.
// On a space-to-depth copy of its input, x_s2d[b, i, j, (py, px), c] = x[b, c, 2i + py, 2j + px], every one of the
// 25 kernel taps is a shift in {-1, 0, 1}^2 of one parity class -- the convolution is the dgrad of the dual
// ConvTranspose2d(k5, s2, p2, op1), its dx that transposed convolution's forward, its dw the dual's wgrad.  These
// layers have few output positions (B x 16^2 .. 4^2) and a long K (25 taps x Cin): the K loop is split over taps
// across CTAs (fp32 partials + splitk_reduce_kernel) so that all SMs work on them.
// -------------------------------------------------------------------------------------------------
namespace hg {

struct Conv5Plan {
    int m_tiles, bn, n_tiles;
    int nsplit[4];        // per parity class (dx) or [0] only (fwd)
    int max_split;
};

static void conv5_fwd_plan(long long m_total, int cout, Conv5Plan &pl)
{
    pl.m_tiles = (int)((m_total + kBM - 1) / kBM);
    int bn = pick_bn(cout);
    if (bn == 256 && (long long)pl.m_tiles * (cout / 256) < 32) bn = 128;
    pl.bn = bn;
    pl.n_tiles = cout / bn;
    const int tiles = pl.m_tiles * pl.n_tiles;
    int ks = sm_count() / tiles;                        // never more CTAs than SMs: a second partial wave doubles the time
    // Re-measured after the epilogue / producer fixes (profiles/r02M_dconv_split.txt): the fp32 partials + reduce pass only
    // pay when there are very few row tiles (B = 64: block 2, 8 tiles: 16.6 vs 20.9 us); with >= 16 the unsplit kernel wins
    // (block 1: 14.2 vs 20.5 us; at 128 x 128: 17.9 vs 30.4 and 21.8 vs 29.1 us)
    if (pl.m_tiles > 8) ks = 1;
    if (ks <= 1) {                                      // unsplit: narrow the N tile until the grid fills the SMs instead
        pl.bn = pick_bn_for_grid(cout, pl.m_tiles);
        pl.n_tiles = cout / pl.bn;
    }
    if (ks < 1) ks = 1;
    if (ks > 12) ks = 12;
    pl.nsplit[0] = pl.max_split = ks;
}

static void conv5_dx_plan(long long m_total, int cin, int cout, Conv5Plan &pl)
{
    pl.m_tiles = (int)((m_total + kBM - 1) / kBM);
    int bn = pick_bn(cin);
    if (bn > 128 && pl.m_tiles < 32) bn = 128;
    pl.bn = bn;
    pl.n_tiles = cin / bn;
    const int taps_c[4] = {9, 6, 6, 4};                         // (py, px) = (0,0), (0,1), (1,0), (1,1): 3x3, 3x2, 2x3, 2x2
    const int kc = cout / kBK;
    const long long total = (long long)pl.m_tiles * pl.n_tiles * 25 * kc;
    long long target = (total + sm_count() - 1) / sm_count();                       // iterations per CTA at one CTA per SM
    if (target < 8) target = 8;
    pl.max_split = 1;
    for (int c = 0; c < 4; ++c) {
        int sp = (int)((taps_c[c] * kc + target / 2) / target);
        if (sp < 1) sp = 1;
        if (sp > 4) sp = 4;
        if (sp > taps_c[c]) sp = taps_c[c];
        pl.nsplit[c] = sp;
    }
    // one wave: while the grid exceeds the SM count, undo the split of the class with the shortest runs
    for (;;) {
        int groups = 0, worst = -1;
        for (int c = 0; c < 4; ++c) {
            groups += pl.nsplit[c];
            if (pl.nsplit[c] > 1 && (worst < 0 || taps_c[c] * pl.nsplit[worst] < taps_c[worst] * pl.nsplit[c])) worst = c;
        }
        if ((long long)pl.m_tiles * pl.n_tiles * groups <= sm_count() || worst < 0) break;
        pl.nsplit[worst]--;
    }
    for (int c = 0; c < 4; ++c)
        if (pl.nsplit[c] > pl.max_split) pl.max_split = pl.nsplit[c];
}

static int launch_splitk_reduce(const float *partial, void *out, long long rows, int ld, int block_cols, const int *nsplit,
                                int nblocks, long long partial_stride, const float *scale, cudaStream_t st)
{
    SplitReduceParams rp{};
    rp.partial = partial; rp.scale = scale; rp.out = static_cast<__nv_bfloat16 *>(out); rp.partial_stride = partial_stride;
    rp.vecs = rows * ld / 8; rp.ld = ld; rp.block_cols = block_cols;
    for (int i = 0; i < nblocks; ++i) rp.nsplit[i] = nsplit[i];
    splitk_reduce_kernel<<<(unsigned)((rp.vecs + 255) / 256), 256, 0, st>>>(rp);
    return check_launch("splitk_reduce");
}

}  // namespace hg

static int conv5_shape(ConvShape &c, const char *who, int batch, int cin, int cout, int size_out)
{
    // the dual transposed convolution: Cin_T = Cout, Cout_T = Cin, input extent = size_out
    c = ConvShape{batch, cout, cin, 2, size_out, 5};
    return conv_shape(c, who);
}

extern "C" long long hg_conv5s2_workspace_bytes(int batch, int cin, int cout, int size_out)
{
    ConvShape c;
    if (conv5_shape(c, "hg_conv5s2_workspace_bytes", batch, cin, cout, size_out)) return -1;
    const long long m_total = (long long)batch * size_out * size_out;
    Conv5Plan pf, pd;
    conv5_fwd_plan(m_total, cout, pf);
    conv5_dx_plan(m_total, cin, cout, pd);
    long long fwd = pf.max_split > 1 ? (long long)pf.max_split * m_total * cout * 4 : 0;
    long long dx = pd.max_split > 1 ? (long long)pd.max_split * m_total * 4 * cin * 4 : 0;
    long long dw = hg_convt_wgrad_workspace_bytes(batch, cout, cin, 2, size_out, 5);
    if (dw < 0) dw = 0;
    long long m = fwd > dx ? fwd : dx;
    return m > dw ? m : dw;
}

// y[b, i, j, co] = sum_{ci, ky, kx} x[b, ci, 2i + ky - 2, 2j + kx - 2] * w[co, ci, ky, kx]    (no bias)
//   x_s2d (B, S, S, 4, Cin) bf16, w_k [25][Cout][Cin] bf16 (tap = ky * 5 + kx), y (B, S, S, Cout) bf16
extern "C" int hg_conv5s2_fwd(const void *x_s2d, const void *w_k, const float *out_scale, void *y, void *workspace,
                              long long workspace_bytes, int batch, int cin, int cout, int size_out, void *stream)
{
    HG_REQUIRE(x_s2d && w_k && y, HG_ERR_INVALID_ARG, "hg_conv5s2_fwd: null pointer");
    ConvShape c;
    int rc = conv5_shape(c, "hg_conv5s2_fwd", batch, cin, cout, size_out);
    if (rc) return rc;
    HG_REQUIRE(cin % kBK == 0 && cout % 16 == 0, HG_ERR_UNSUPPORTED, "hg_conv5s2_fwd: need Cin %% 64 == 0 and Cout %% 16 == 0 (got %d, %d)", cin, cout);
    const long long m_total = (long long)batch * size_out * size_out;
    Conv5Plan pl;
    conv5_fwd_plan(m_total, cout, pl);
    const int ks = pl.max_split;
    HG_REQUIRE(ks == 1 || (workspace && workspace_bytes >= (long long)ks * m_total * cout * 4), HG_ERR_INVALID_ARG,
               "hg_conv5s2_fwd: workspace smaller than hg_conv5s2_workspace_bytes()");
    TapGemmParams p{};
    int bx, by, bz, bb;
    HG_REQUIRE(box_for(kBM, c.X, c.Y, c.Z, bx, by, bz, bb), HG_ERR_UNSUPPORTED, "hg_conv5s2_fwd: output size %d does not tile into 128-row boxes", size_out);
    rc = make_map_5d(&p.tmA, x_s2d, 4ll * cin, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    rc = make_map_2d(&p.tmB, w_k, cin, 25ll * cout, pl.bn);
    if (rc) return rc;
    p.X = c.X; p.Y = c.Y; p.Z = c.Z; p.Bn = batch;
    p.BN = pl.bn; p.epi_cols = pl.bn; p.bias_mod = cout; p.k_chunks = cin / kBK;
    p.m_total = (int)m_total; p.ld_out = cout; p.out = y; p.bias = nullptr; p.slope = 1.0f;
    if (ks > 1) { p.partial = static_cast<float *>(workspace); p.partial_stride = m_total * cout; }
    p.out_scale = out_scale;
    int ntap = 0;
    for_each_class_tap(c, [&](int cls, int flat, int sx, int sy, int sz) {
        p.taps[ntap++] = Tap{(int16_t)-sx, (int16_t)-sy, (int16_t)-sz, 0, cls * cin, flat * cout};
    });
    p.num_groups = ks;
    for (int s = 0; s < ks; ++s) {
        const int t0 = s * ntap / ks, t1 = (s + 1) * ntap / ks;
        p.groups[s] = Group{t0, t1 - t0, 0, s};
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_tap_gemm(p, pl.m_tiles, pl.n_tiles, st, "hg_conv5s2_fwd");
    if (rc || ks == 1) return rc;
    return launch_splitk_reduce(p.partial, y, m_total, cout, cout, pl.nsplit, 1, p.partial_stride, out_scale, st);
}

// dx_s2d[b, i, j, (py, px), ci] = d loss / d x[b, ci, 2i + py, 2j + px]
//   dy (B, S, S, Cout) bf16, w_t [25][Cin][Cout] bf16, dx_s2d (B, S, S, 4, Cin) bf16
extern "C" int hg_conv5s2_dx(const void *dy, const void *w_t, const float *out_scale, void *dx_s2d, void *workspace,
                             long long workspace_bytes, int batch, int cin, int cout, int size_out, void *stream)
{
    HG_REQUIRE(dy && w_t && dx_s2d, HG_ERR_INVALID_ARG, "hg_conv5s2_dx: null pointer");
    ConvShape c;
    int rc = conv5_shape(c, "hg_conv5s2_dx", batch, cin, cout, size_out);
    if (rc) return rc;
    HG_REQUIRE(cout % kBK == 0 && cin % 16 == 0, HG_ERR_UNSUPPORTED, "hg_conv5s2_dx: need Cout %% 64 == 0 and Cin %% 16 == 0 (got %d, %d)", cout, cin);
    const long long m_total = (long long)batch * size_out * size_out;
    Conv5Plan pl;
    conv5_dx_plan(m_total, cin, cout, pl);
    if (pl.max_split == 1)      // enough CTAs without splitting: the transposed convolution's own launch plan
        return convt_fwd_impl(dy, w_t, nullptr, out_scale, dx_s2d, batch, cout, cin, 2, size_out, 5, 1.0f, stream);
    HG_REQUIRE(workspace && workspace_bytes >= (long long)pl.max_split * m_total * 4 * cin * 4, HG_ERR_INVALID_ARG,
               "hg_conv5s2_dx: workspace smaller than hg_conv5s2_workspace_bytes()");
    TapGemmParams p{};
    int bx, by, bz, bb;
    HG_REQUIRE(box_for(kBM, c.X, c.Y, c.Z, bx, by, bz, bb), HG_ERR_UNSUPPORTED, "hg_conv5s2_dx: output size %d does not tile into 128-row boxes", size_out);
    rc = make_map_5d(&p.tmA, dy, cout, c.X, c.Y, c.Z, batch, bx, by, bz, bb);
    if (rc) return rc;
    rc = make_map_2d(&p.tmB, w_t, cout, 25ll * cin, pl.bn);
    if (rc) return rc;
    p.X = c.X; p.Y = c.Y; p.Z = c.Z; p.Bn = batch;
    p.BN = pl.bn; p.epi_cols = pl.bn; p.bias_mod = cin; p.k_chunks = cout / kBK;
    p.m_total = (int)m_total; p.ld_out = 4ll * cin; p.out = dx_s2d; p.bias = nullptr; p.slope = 1.0f;
    p.partial = static_cast<float *>(workspace); p.partial_stride = m_total * 4 * cin;
    // taps in class-major order (for_each_class_tap), a class's taps cut into nsplit[cls] contiguous runs
    int ntap = 0, class_begin[5] = {0, 0, 0, 0, 0};
    for_each_class_tap(c, [&](int cls, int flat, int sx, int sy, int sz) {
        p.taps[ntap++] = Tap{(int16_t)sx, (int16_t)sy, (int16_t)sz, 0, 0, flat * cin};
        class_begin[cls + 1] = ntap;
    });
    int ng = 0;
    for (int cls = 0; cls < 4; ++cls) {
        const int n = class_begin[cls + 1] - class_begin[cls], sp = pl.nsplit[cls];
        for (int s = 0; s < sp; ++s) {
            const int t0 = class_begin[cls] + s * n / sp, t1 = class_begin[cls] + (s + 1) * n / sp;
            p.groups[ng++] = Group{t0, t1 - t0, cls * cin, s};
        }
    }
    p.num_groups = ng;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_tap_gemm(p, pl.m_tiles, pl.n_tiles, st, "hg_conv5s2_dx");
    if (rc) return rc;
    return launch_splitk_reduce(p.partial, dx_s2d, m_total, 4 * cin, cin, pl.nsplit, 4, p.partial_stride, out_scale, st);
}

// dw[co, ci, ky, kx] (torch Conv2d layout, fp32) = sum over positions; accumulate != 0 adds into dw.
extern "C" int hg_conv5s2_dw(const void *dy, const void *x_s2d, float *dw, void *workspace, long long workspace_bytes,
                             int batch, int cin, int cout, int size_out, int accumulate, void *stream)
{
    return hg_convt_wgrad(dy, x_s2d, dw, workspace, workspace_bytes, batch, cout, cin, 2, size_out, 5, 0, 0, accumulate, stream);
}

// Conv2d weight (Cout, Cin, 5, 5) fp32 contiguous, divided by *sigma (device pointer, may be null) ->
//   w_k [25][Cout][Cin] bf16 (hg_conv5s2_fwd) and w_t [25][Cin][Cout] bf16 (hg_conv5s2_dx); either may be null.
extern "C" int hg_conv5s2_pack_weight(const float *w, const float *sigma, void *w_k, void *w_t, int cin, int cout, void *stream)
{
    // as the dual transposed convolution's weight: Cin_T = Cout, Cout_T = Cin -> w_dgrad [t][Cin_T][Cout_T] = w_k, w_fwd = w_t
    return pack_weight_impl(w, w_t, w_k, cout, cin, 25, 0, 0, sigma, stream);
}
