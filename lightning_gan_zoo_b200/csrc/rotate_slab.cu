// Rotate-resample forward on the torch NCDHW layout at 32^3: source-slab tiles.
// Replaces core/models/hologan_generator.py:198-331 of the reference for the 32^3 volumes of BASELINE cfg 3.
//
// The interleaved-tile design of rotate_il.cu (one 16-byte unit per voxel = 4 fp32 / 8 bf16 channels, hashed
// placement, one LDS.128 per corner for the whole channel group) needs the source volume of a channel group in shared
// memory: 64 KB at 16^3, 512 KB at 32^3.  Here a CTA stages only a SLAB of the source -- kSlabZ z-planes plus the one
// plane of halo the trilinear footprint needs (5 x 1024 units = 80 KB, two CTAs per SM) -- and produces exactly the
// outputs whose floor(z) falls into its slab:
//   * z_src is an affine function of the output lattice point, so whether an 8x2x2 output block can touch the slab at
//     all is a warp-uniform test on the block's z-range (conservative by kSlabEps); 7 of 8 blocks are skipped with ~20
//     instructions per warp;
//   * inside a candidate block every lane recomputes its coordinates with the bit-exact chain of the other kernels
//     and keeps the output iff  clamp(floor(z), 0, 31) / kSlabZ == slab  -- every output is owned by exactly one slab,
//     the clamped corner planes z0, z1 = z0 + 1 (clamped) always lie inside the staged planes;
//   * corner indices / weights / summation order are those of rotate_il.cu (fp32 results carry the reference's bits).
// HBM traffic: the volume is read (kSlabZ + 1) / kSlabZ = 1.25 times, written once.
// Status: parity-tested against the per-channel kernels when enabled (HG_ROTATE_SLAB32=1, tests/test_gpu_rotate.py);
// opt-in until it has been measured on a B200.
#include "hg_common.cuh"
#include "rotate_common.cuh"
#include "rotate_il.cuh"

namespace hg {

constexpr int kSlabLogS = 5, kSlabS = 32, kSlabN = kSlabS * kSlabS * kSlabS;
constexpr int kSlabZ = 4;                                  // source planes owned by a CTA
constexpr int kSlabs = kSlabS / kSlabZ;                    // 8
constexpr int kSlabPlanes = kSlabZ + 1;                    // + halo
constexpr int kSlabUnits = kSlabPlanes * kSlabS * kSlabS;  // 5120 units of 16 bytes = 80 KB
constexpr int kSlabThreads = 512;
constexpr float kSlabEps = 0.01f;                          // slack of the warp-uniform block test (plain fp32 vs the exact chain)

template <typename T, bool kZeroBorder>
__global__ void __launch_bounds__(kSlabThreads, 2) rotate_fwd_slab32_kernel(const T *__restrict__ vol, const float *__restrict__ a_inv,
                                                                            T *__restrict__ out, int groups)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int S = kSlabS, LOGS = kSlabLogS, N = kSlabN, CI = IlUnit<T>::CI;
    uint4 *tile = reinterpret_cast<uint4 *>(smem_raw);     // [kSlabUnits], local voxel v' = ((z - z_base) * S + y) * S + x, hashed
    __shared__ float m[12];
    const int slab = blockIdx.x, grp = blockIdx.y, b = blockIdx.z;
    const int z_base = slab * kSlabZ;
    const int planes = min(kSlabPlanes, S - z_base);       // the last slab has no halo plane (z1 clamps to S - 1)
    const T *src = vol + ((size_t)b * groups + grp) * CI * N + (size_t)z_base * S * S;
    T *dst = out + ((size_t)b * groups + grp) * CI * N;
    if (threadIdx.x < 12) m[threadIdx.x] = a_inv[b * 16 + threadIdx.x];

    // ---- stage the slab, channel-interleaved ----------------------------------------------------------------
    const int n_stage = planes * S * S;
    if constexpr (sizeof(T) == 4) {
#pragma unroll 4
        for (int v = threadIdx.x; v < n_stage; v += kSlabThreads) {
            uint4 q;
            q.x = ld_stream_4(src + v);
            q.y = ld_stream_4(src + N + v);
            q.z = ld_stream_4(src + 2 * N + v);
            q.w = ld_stream_4(src + 3 * N + v);
            tile[il_unit(v, LOGS)] = q;
        }
    } else {
#pragma unroll 2
        for (int v = 2 * threadIdx.x; v < n_stage; v += 2 * kSlabThreads) {     // voxel pair (v, v + 1), v even
            uint32_t w[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) w[c] = ld_stream_4(src + (size_t)c * N + v);
            uint4 lo, hi;
            lo.x = __byte_perm(w[0], w[1], 0x5410); hi.x = __byte_perm(w[0], w[1], 0x7632);
            lo.y = __byte_perm(w[2], w[3], 0x5410); hi.y = __byte_perm(w[2], w[3], 0x7632);
            lo.z = __byte_perm(w[4], w[5], 0x5410); hi.z = __byte_perm(w[4], w[5], 0x7632);
            lo.w = __byte_perm(w[6], w[7], 0x5410); hi.w = __byte_perm(w[6], w[7], 0x7632);
            tile[il_unit(v, LOGS)] = lo;
            tile[il_unit(v + 1, LOGS)] = hi;
        }
    }
    __syncthreads();

    // ---- gather -----------------------------------------------------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float lim = (float)(S - 1);
    const float zlo = (float)z_base - kSlabEps, zhi = (float)(z_base + kSlabZ) + kSlabEps;
    for (int j = warp; j < N / 32; j += kSlabThreads / 32) {
        // warp-uniform reject: z range of the 8x2x2 block whose low corner is lane 0's voxel
        int bx, by, bz;
        il_block_voxel(j, 0, S, LOGS, bx, by, bz);
        const float zc = m[8] * (float)bx + m[9] * (float)by + m[10] * (float)bz + m[11];
        const float zmin = zc + fminf(0.f, 7.f * m[8]) + fminf(0.f, m[9]) + fminf(0.f, m[10]);
        const float zmax = zc + fmaxf(0.f, 7.f * m[8]) + fmaxf(0.f, m[9]) + fmaxf(0.f, m[10]);
        if (!((slab == 0 || zmax >= zlo) && (slab == kSlabs - 1 || zmin < zhi))) continue;

        int ox, oy, oz;
        il_block_voxel(j, lane, S, LOGS, ox, oy, oz);
        float x, y, z;
        il_coords(m, ox, oy, oz, x, y, z);
        const int fz = __float2int_rd(z);
        if ((clampi(fz, S - 1) >> 2) != slab) continue;    // owned by another slab's CTA  (kSlabZ == 4)
        const int o = (((oz << LOGS) + oy) << LOGS) + ox;
        const bool inside = (x >= 0.f) && (x < lim) && (y >= 0.f) && (y < lim) && (z >= 0.f) && (z < lim);
        if (kZeroBorder && !inside) {
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) st_stream_elem<T>(dst + (size_t)ci * N + o, 0.f);
            continue;
        }
        // corners and weights: the arithmetic of il_corners / make_corners, rows relative to the slab
        const int fx = __float2int_rd(x), fy = __float2int_rd(y);
        const int x0 = clampi(fx, S - 1), x1 = clampi(fx + 1, S - 1);
        const int y0 = clampi(fy, S - 1), y1 = clampi(fy + 1, S - 1);
        const int z0 = clampi(fz, S - 1), z1 = clampi(fz + 1, S - 1);
        const float ux = __fsub_rn((float)x1, x), lx = __fsub_rn(x, (float)x0);
        const float uy = __fsub_rn((float)y1, y), ly = __fsub_rn(y, (float)y0);
        const float uz = __fsub_rn((float)z1, z), lz = __fsub_rn(z, (float)z0);
        const int z0l = z0 - z_base, z1l = z1 - z_base;    // 0 .. kSlabZ (z_base is a multiple of 4: same hash bits)
        const int k00 = ((y0 << 1) ^ (z0l << 2)) & 7, k01 = ((y1 << 1) ^ (z0l << 2)) & 7;
        const int k10 = ((y0 << 1) ^ (z1l << 2)) & 7, k11 = ((y1 << 1) ^ (z1l << 2)) & 7;
        const int r00 = ((z0l << LOGS) + y0) << LOGS, r01 = ((z0l << LOGS) + y1) << LOGS;
        const int r10 = ((z1l << LOGS) + y0) << LOGS, r11 = ((z1l << LOGS) + y1) << LOGS;
        int u[8];
        u[0] = r00 | (x0 ^ k00); u[1] = r01 | (x0 ^ k01); u[2] = r00 | (x1 ^ k00); u[3] = r01 | (x1 ^ k01);
        u[4] = r10 | (x0 ^ k10); u[5] = r11 | (x0 ^ k11); u[6] = r10 | (x1 ^ k10); u[7] = r11 | (x1 ^ k11);
        const float uxuy = __fmul_rn(ux, uy), uxly = __fmul_rn(ux, ly), lxuy = __fmul_rn(lx, uy), lxly = __fmul_rn(lx, ly);
        float w[8];
        w[0] = __fmul_rn(uxuy, uz); w[1] = __fmul_rn(uxly, uz); w[2] = __fmul_rn(lxuy, uz); w[3] = __fmul_rn(lxly, uz);
        w[4] = __fmul_rn(uxuy, lz); w[5] = __fmul_rn(uxly, lz); w[6] = __fmul_rn(lxuy, lz); w[7] = __fmul_rn(lxly, lz);
        uint4 raw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = tile[u[k]];
        float acc[CI], f[CI];
        IlUnit<T>::unpack(raw[0], f);
#pragma unroll
        for (int i = 0; i < CI; ++i) acc[i] = __fmul_rn(w[0], f[i]);
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            IlUnit<T>::unpack(raw[k], f);
#pragma unroll
            for (int i = 0; i < CI; ++i) {
                if (kZeroBorder) acc[i] = fmaf(w[k], f[i], acc[i]);                      // fast mode
                else acc[i] = __fadd_rn(acc[i], __fmul_rn(w[k], f[i]));                 // reference order (:320)
            }
        }
#pragma unroll
        for (int i = 0; i < CI; ++i) st_stream_elem<T>(dst + (size_t)i * N + o, acc[i]);
    }
}

// -------------------------------------------------------------------------------------------------
// Backward for volumes whose channel group does not fit in shared memory (32^3): per-voxel gather with an analytic
// candidate search -- no tables, no workspace, no atomics (shared-memory fp32 atomicAdd is a CAS loop on sm_100,
// ATOMS.CAST.SPIN in the SASS: that is what makes the scatter backward of rotate.cu take 5.5 ms at (64,64,32^3)).
//   grad_vol[s] = sum over outputs o whose 2x2x2 footprint contains s of w(o, s) * grad_out[o]
// The outputs that can touch source voxel s satisfy |src(o) - s|_inf < 1, i.e. o lies in the image of the cube
// (s-1, s+1)^3 under the inverse affine map: |o_a - oc_a| < h_a = sum_j |Linv[a][j]| with oc = Linv (s - t).  Every
// lane scans that lattice box (same trip counts for the whole sample), decides hits with the forward's exact
// coordinate chain (same bits -> same floor cell as the forward) and keeps (output index, weight) pairs in a
// per-thread shared-memory list (phase 1); phase 2 walks the list once per chunk of channels, so the search is paid
// once per voxel and the gathers of a warp (32 neighbouring source voxels) hit the same few lines of grad_out.
// Hits beyond the list capacity (strongly shrinking views) are re-enumerated per chunk.  Fixed order: deterministic.
// Out-of-range outputs contribute exactly 0, as in the other backward kernels (include/hologan_b200.h).
// Status: opt-in (HG_ROTATE_GATHER_BWD=1), enumeration emulated on CPU against the scatter definition; not yet run on
// a B200.
// -------------------------------------------------------------------------------------------------
constexpr int kGbThreads = 256;
constexpr int kGbHits = 20;                                // list entries per source voxel (8 cells x ~1-2 outputs at scale 1)
constexpr int kGbChunk = 8;                                // channels per pass over the list
constexpr float kGbEps = 0.01f;

struct GbSearch {
    float m[12];        // rows 0..2 of a_inv: src = L o + t
    float li[9];        // L^-1
    float h[3];         // half extents of the candidate box
    int n[3];           // trip counts per axis
};

__device__ __forceinline__ bool gb_hit(const float *__restrict__ m, int ox, int oy, int oz, int sx, int sy, int sz, float lim,
                                       float &w)
{
    float x, y, z;
    il_coords(m, ox, oy, oz, x, y, z);                     // the forward's bits
    if (!((x >= 0.f) && (x < lim) && (y >= 0.f) && (y < lim) && (z >= 0.f) && (z < lim))) return false;
    const int qx = __float2int_rd(x), qy = __float2int_rd(y), qz = __float2int_rd(z);
    const int dx = sx - qx, dy = sy - qy, dz = sz - qz;
    if ((unsigned)dx > 1u || (unsigned)dy > 1u || (unsigned)dz > 1u) return false;
    const float wx = dx ? __fsub_rn(x, (float)qx) : __fsub_rn((float)(qx + 1), x);
    const float wy = dy ? __fsub_rn(y, (float)qy) : __fsub_rn((float)(qy + 1), y);
    const float wz = dz ? __fsub_rn(z, (float)qz) : __fsub_rn((float)(qz + 1), z);
    w = __fmul_rn(__fmul_rn(wx, wy), wz);
    return true;
}

template <typename T>
__global__ void __launch_bounds__(kGbThreads) rotate_bwd_gather_kernel(const T *__restrict__ grad_out, const float *__restrict__ a_inv,
                                                                       T *__restrict__ grad_vol, int C, int S, int logS)
{
    extern __shared__ uint2 gb_hits[];                      // [kGbHits][kGbThreads]: (output index, weight bits)
    __shared__ GbSearch q;
    const int N = S * S * S;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        const float *a = a_inv + b * 16;
        for (int i = 0; i < 12; ++i) q.m[i] = a[i];
        const float a0 = a[0], a1 = a[1], a2 = a[2], b0 = a[4], b1 = a[5], b2 = a[6], c0 = a[8], c1 = a[9], c2 = a[10];
        const float A = b1 * c2 - b2 * c1, B = a2 * c1 - a1 * c2, Cc = a1 * b2 - a2 * b1;
        const float D = b2 * c0 - b0 * c2, E = a0 * c2 - a2 * c0, F = a2 * b0 - a0 * b2;
        const float G = b0 * c1 - b1 * c0, H = a1 * c0 - a0 * c1, I = a0 * b1 - a1 * b0;
        const float det = a0 * A + a1 * D + a2 * G;
        const bool ok = fabsf(det) > 1e-12f;
        const float r = ok ? 1.0f / det : 0.f;
        const float li[9] = {A * r, B * r, Cc * r, D * r, E * r, F * r, G * r, H * r, I * r};
        for (int i = 0; i < 9; ++i) q.li[i] = li[i];
        for (int ax = 0; ax < 3; ++ax) {
            const float h = fabsf(li[3 * ax]) + fabsf(li[3 * ax + 1]) + fabsf(li[3 * ax + 2]) + kGbEps;
            q.h[ax] = h;
            const float span = 2.f * h + 2.f;
            q.n[ax] = (!ok || span > (float)S) ? S : (int)span;       // degenerate / strongly shrinking: scan the axis
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * (kGbThreads / 32) + warp;   // 8x2x2 block of source voxels
    int sx, sy, sz;
    il_block_voxel(j, lane, S, logS, sx, sy, sz);
    const int s = (((sz << logS) + sy) << logS) + sx;
    const float lim = (float)(S - 1);
    // candidate box: start_a = ceil(oc_a - h_a), n_a points (or the whole axis)
    const float tx = (float)sx - q.m[3], ty = (float)sy - q.m[7], tz = (float)sz - q.m[11];
    int st[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float oc = q.li[3 * ax] * tx + q.li[3 * ax + 1] * ty + q.li[3 * ax + 2] * tz;
        st[ax] = q.n[ax] >= S ? 0 : (int)ceilf(oc - q.h[ax]);
    }
    const int nx = q.n[0], ny = q.n[1], nz = q.n[2];
    // ---- phase 1: enumerate, keep the first kGbHits hits ----------------------------------------------------
    int count = 0;
    for (int iz = 0; iz < nz; ++iz) {
        const int oz = st[2] + iz;
        for (int iy = 0; iy < ny; ++iy) {
            const int oy = st[1] + iy;
            for (int ix = 0; ix < nx; ++ix) {
                const int ox = st[0] + ix;
                float w;
                if ((unsigned)ox < (unsigned)S && (unsigned)oy < (unsigned)S && (unsigned)oz < (unsigned)S &&
                    gb_hit(q.m, ox, oy, oz, sx, sy, sz, lim, w)) {
                    if (count < kGbHits)
                        gb_hits[count * kGbThreads + threadIdx.x] = make_uint2((uint32_t)((((oz << logS) + oy) << logS) + ox), __float_as_uint(w));
                    ++count;
                }
            }
        }
    }
    const int listed = min(count, kGbHits);
    const int nmax = __reduce_max_sync(0xffffffffu, listed);
    // ---- phase 2: one pass over the list per chunk of channels ------------------------------------------------
    for (int c0 = 0; c0 < C; c0 += kGbChunk) {
        const int nc = min(kGbChunk, C - c0);
        const T *g = grad_out + ((size_t)b * C + c0) * N;
        float acc[kGbChunk];
#pragma unroll
        for (int c = 0; c < kGbChunk; ++c) acc[c] = 0.f;
        for (int i = 0; i < nmax; ++i) {
            if (i < listed) {
                const uint2 e = gb_hits[i * kGbThreads + threadIdx.x];
                const float w = __uint_as_float(e.y);
#pragma unroll
                for (int c = 0; c < kGbChunk; ++c)
                    if (c < nc) acc[c] = fmaf(w, to_f32<T>(g[(size_t)c * N + e.x]), acc[c]);
            }
        }
        if (count > kGbHits) {                              // rare: re-enumerate the hits that did not fit
            int seen = 0;
            for (int iz = 0; iz < nz; ++iz)
                for (int iy = 0; iy < ny; ++iy)
                    for (int ix = 0; ix < nx; ++ix) {
                        const int ox = st[0] + ix, oy = st[1] + iy, oz = st[2] + iz;
                        float w;
                        if ((unsigned)ox < (unsigned)S && (unsigned)oy < (unsigned)S && (unsigned)oz < (unsigned)S &&
                            gb_hit(q.m, ox, oy, oz, sx, sy, sz, lim, w)) {
                            if (seen >= kGbHits) {
                                const int o = (((oz << logS) + oy) << logS) + ox;
                                for (int c = 0; c < nc; ++c) acc[c] = fmaf(w, to_f32<T>(g[(size_t)c * N + o]), acc[c]);
                            }
                            ++seen;
                        }
                    }
        }
        T *dst = grad_vol + ((size_t)b * C + c0) * N + s;
#pragma unroll
        for (int c = 0; c < kGbChunk; ++c)
            if (c < nc) st_stream_elem<T>(dst + (size_t)c * N, acc[c]);
    }
}

template <typename T, bool Z>
static int launch_fwd_slab32(const void *vol, const float *a, void *out, int B, int C, cudaStream_t st)
{
    const size_t smem = (size_t)kSlabUnits * 16;
    auto k = rotate_fwd_slab32_kernel<T, Z>;
    static bool attr_done = false;      // per instantiation; not a stream operation (graph-capture safe)
    if (!attr_done) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    const int groups = C / IlUnit<T>::CI;
    k<<<dim3(kSlabs, groups, B), kSlabThreads, smem, st>>>(static_cast<const T *>(vol), a, static_cast<T *>(out), groups);
    return check_launch("rotate_fwd_slab32");
}

}  // namespace hg

using namespace hg;

// 32^3, channel counts that are a multiple of one 16-byte unit, opt-in (see the header comment)
bool hg_rotate_slab32_enabled(int channels, int size, int dtype, int batch)
{
    const int ci = dtype == HG_F32 ? 4 : 8;
    if (size != kSlabS || channels % ci != 0 || channels / ci > 65535 || batch > 65535) return false;
    return option(kOptRotateSlab32) != 0;
}

int hg_rotate_slab32_fwd(const void *vol, const float *a_inv, void *out, int batch, int channels, int dtype, int border,
                         cudaStream_t st)
{
    const bool z = (border & 0xFF) == HG_BORDER_ZERO;
    if (dtype == HG_F32)
        return z ? launch_fwd_slab32<float, true>(vol, a_inv, out, batch, channels, st)
                 : launch_fwd_slab32<float, false>(vol, a_inv, out, batch, channels, st);
    return z ? launch_fwd_slab32<__nv_bfloat16, true>(vol, a_inv, out, batch, channels, st)
             : launch_fwd_slab32<__nv_bfloat16, false>(vol, a_inv, out, batch, channels, st);
}

// size 32 (any channel count), opt-in
bool hg_rotate_gather_bwd_enabled(int size, int batch)
{
    if (size != 32 || batch > 65535) return false;
    return option(kOptRotateGatherBwd) != 0;
}

int hg_rotate_gather_bwd(const void *grad_out, const float *a_inv, void *grad_vol, int batch, int channels, int size, int logS,
                         int dtype, cudaStream_t st)
{
    const int n = size * size * size;
    const size_t smem = (size_t)kGbHits * kGbThreads * sizeof(uint2);
    dim3 grid(n / kGbThreads, batch);
    if (dtype == HG_F32)
        rotate_bwd_gather_kernel<float><<<grid, kGbThreads, smem, st>>>(static_cast<const float *>(grad_out), a_inv,
                                                                       static_cast<float *>(grad_vol), channels, size, logS);
    else
        rotate_bwd_gather_kernel<__nv_bfloat16><<<grid, kGbThreads, smem, st>>>(static_cast<const __nv_bfloat16 *>(grad_out), a_inv,
                                                                               static_cast<__nv_bfloat16 *>(grad_vol), channels, size, logS);
    return check_launch("rotate_bwd_gather");
}
