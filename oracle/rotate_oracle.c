/* CPU oracle for the HoloGAN voxel rotate-resample  --  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of core/models/hologan_generator.py:225-321 (reference commit 33c7f1b):
 * apply_transformation's lattice x inverse-matrix product and `interpolation`'s clamped-corner
 * trilinear gather, plus the adjoint (what autograd's index_put_(accumulate=True) produces).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product never does.
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/rotate_*.npz, which hold outputs of
 * the reference itself (oracle/gen_golden.py): coordinates and forward output bit-exact.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit FMA contraction; the one FMA
 * chain below is explicit).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* Source coordinate of output lattice point (x,y,z) under one row of the inverse matrix.
 * hologan_generator.py:229 computes `A @ grid` with torch.matmul (MKL sgemm, k=4); its fp32 bits
 * are reproduced by this sequential chain (SURVEY.md section 7 "Bit-exact coordinates"). */
static inline float row_dot(const float *r, float x, float y, float z)
{
    float acc = r[0] * x;
    acc = fmaf(r[1], y, acc);
    acc = fmaf(r[2], z, acc);
    acc = fmaf(r[3], 1.0f, acc);
    return acc;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* a_inv: (B,4,4) row-major fp32 (only rows 0..2 are used); coords: (3,B,S^3) x,y,z planes;
 * lattice order z-major, x fastest (hologan_generator.py:323-331). */
void orc_rotate_coords(const float *a_inv, float *coords, int batch, int size)
{
    const size_t n = (size_t)size * size * size;
    for (int b = 0; b < batch; ++b) {
        const float *m = a_inv + (size_t)b * 16;
        size_t i = 0;
        for (int z = 0; z < size; ++z)
            for (int y = 0; y < size; ++y)
                for (int x = 0; x < size; ++x, ++i)
                    for (int r = 0; r < 3; ++r)
                        coords[((size_t)r * batch + b) * n + i] = row_dot(m + 4 * r, (float)x, (float)y, (float)z);
    }
}

typedef struct {
    int idx[8];     /* flat (z*S+y)*S+x of corners a..h */
    float w[8];     /* weights a..h */
} corners_t;

/* hologan_generator.py:249-261 (floor, +1, clamp) and :301-318 (weights from the clamped corner as
 * float against the unclamped coordinate; product order (wx*wy)*wz). */
static inline void corners(float x, float y, float z, int s, corners_t *c)
{
    const int fx = (int)floorf(x), fy = (int)floorf(y), fz = (int)floorf(z);
    const int x0 = clampi(fx, 0, s - 1), x1 = clampi(fx + 1, 0, s - 1);
    const int y0 = clampi(fy, 0, s - 1), y1 = clampi(fy + 1, 0, s - 1);
    const int z0 = clampi(fz, 0, s - 1), z1 = clampi(fz + 1, 0, s - 1);
    const float ux = (float)x1 - x, lx = x - (float)x0;
    const float uy = (float)y1 - y, ly = y - (float)y0;
    const float uz = (float)z1 - z, lz = z - (float)z0;
    const int zi[2] = {z0, z1}, yi[2] = {y0, y1}, xi[2] = {x0, x1};
    const float wz[2] = {uz, lz}, wy[2] = {uy, ly}, wx[2] = {ux, lx};
    /* order a..h = (z0,y0,x0) (z0,y1,x0) (z0,y0,x1) (z0,y1,x1) (z1,...) : :278-287 */
    static const int ord[8][3] = {{0,0,0},{0,1,0},{0,0,1},{0,1,1},{1,0,0},{1,1,0},{1,0,1},{1,1,1}};
    for (int k = 0; k < 8; ++k) {
        const int kz = ord[k][0], ky = ord[k][1], kx = ord[k][2];
        c->idx[k] = (zi[kz] * s + yi[ky]) * s + xi[kx];
        c->w[k] = (wx[kx] * wy[ky]) * wz[kz];
    }
}

/* Coordinates far outside int range never occur for 16^3/32^3 lattices with sane views; the
 * reference's .long() of a huge float is undefined-ish too, so no special casing. */

/* vol, out: (B,C,S,S,S) fp32 contiguous.  out[b,c,o] = sum_k w_k * vol[b,c,idx_k] summed a..h left
 * to right with separate multiply and add (:320). */
void orc_rotate_fwd(const float *vol, const float *a_inv, float *out, int batch, int ch, int size)
{
    const size_t n = (size_t)size * size * size;
    for (int b = 0; b < batch; ++b) {
        const float *m = a_inv + (size_t)b * 16;
        size_t o = 0;
        for (int z = 0; z < size; ++z)
            for (int y = 0; y < size; ++y)
                for (int x = 0; x < size; ++x, ++o) {
                    corners_t c;
                    corners(row_dot(m, (float)x, (float)y, (float)z), row_dot(m + 4, (float)x, (float)y, (float)z),
                            row_dot(m + 8, (float)x, (float)y, (float)z), size, &c);
                    for (int ci = 0; ci < ch; ++ci) {
                        const float *v = vol + ((size_t)b * ch + ci) * n;
                        float acc = c.w[0] * v[c.idx[0]];
                        for (int k = 1; k < 8; ++k) {
                            const float t = c.w[k] * v[c.idx[k]];
                            acc = acc + t;
                        }
                        out[((size_t)b * ch + ci) * n + o] = acc;
                    }
                }
    }
}

/* Adjoint w.r.t. vol: grad_vol[b,c,idx_k] += w_k * grad_out[b,c,o].  (autograd of :292-320; the
 * reference's accumulation order inside index_put_ is unspecified -> compare with a tolerance.) */
void orc_rotate_bwd(const float *grad_out, const float *a_inv, float *grad_vol, int batch, int ch, int size)
{
    const size_t n = (size_t)size * size * size;
    memset(grad_vol, 0, sizeof(float) * (size_t)batch * ch * n);
    for (int b = 0; b < batch; ++b) {
        const float *m = a_inv + (size_t)b * 16;
        size_t o = 0;
        for (int z = 0; z < size; ++z)
            for (int y = 0; y < size; ++y)
                for (int x = 0; x < size; ++x, ++o) {
                    corners_t c;
                    corners(row_dot(m, (float)x, (float)y, (float)z), row_dot(m + 4, (float)x, (float)y, (float)z),
                            row_dot(m + 8, (float)x, (float)y, (float)z), size, &c);
                    for (int ci = 0; ci < ch; ++ci) {
                        const float g = grad_out[((size_t)b * ch + ci) * n + o];
                        float *gv = grad_vol + ((size_t)b * ch + ci) * n;
                        for (int k = 0; k < 8; ++k)
                            gv[c.idx[k]] += c.w[k] * g;
                    }
                }
    }
}
