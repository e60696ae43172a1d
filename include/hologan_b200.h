/* hologan_b200.h -- C ABI of the B200-native HoloGAN generator hot path.
 *
 * The reference (ebartrum/lightning_gan_zoo @ 33c7f1b) is 100 % Python on stock torch ops and has no
 * FFI of its own; each entry point below names the reference lines it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions only enqueue work: no allocation, no synchronisation, safe to call from several host
 *     threads and inside CUDA-graph capture.  The only process-wide state is the table of tuning options
 *     below (hg_set_option), which selects between equivalent kernels and is never read from the
 *     environment at launch time;
 *   - return 0 on success, a negative hg_status_t otherwise; hg_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 */
#ifndef HOLOGAN_B200_H
#define HOLOGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_ABI_VERSION 2
#define HG_SN_MAX_LAYERS 4
#define HG_LINEAR_GROUP_MAX 8

typedef enum {
    HG_OK = 0,
    HG_ERR_INVALID_ARG = -1,   /* null pointer, non-positive dim, unknown enum value     */
    HG_ERR_UNSUPPORTED = -2,   /* shape / dtype / layout combination has no kernel        */
    HG_ERR_LAUNCH = -3,        /* CUDA reported an error at launch                        */
    HG_ERR_NO_DEVICE = -4      /* no sm_100 device / driver entry point unavailable       */
} hg_status_t;

typedef enum { HG_F32 = 0, HG_BF16 = 1 } hg_dtype_t;

/* Feature-volume layouts.  NCDHW is the reference's (torch) layout. */
typedef enum {
    HG_NCDHW = 0,   /* (B, C, S, S, S)                                                     */
    HG_NDHWC = 1,   /* (B, S, S, S, C)   channels-last                                     */
    HG_PROJ = 2     /* (B, S, S, S, C) ordered [b, z, x, y, c]: the depth-into-channels fold of  *
                     * hologan_generator.py:130-133 as the A operand of the 1x1 projection  *
                     * GEMM -- row (b, z, x), K index y*C + c, which pairs with the          *
                     * reference's folded channel c*S + j at y = S-1-j                       */
} hg_layout_t;

/* How samples whose source coordinate leaves [0, S-1) are produced. */
typedef enum {
    HG_BORDER_REFERENCE = 0,  /* reference arithmetic: clamped corners, weights from the clamped *
                               * corner vs the unclamped coordinate (hologan_generator.py:256-318) *
                               * -> bit-identical to the reference CPU path in fp32                */
    HG_BORDER_ZERO = 1        /* write exact 0 where the reference's terms cancel to ~1e-7         */
} hg_border_t;
/* Launch-shape tuning flag, OR-ed into a `border` argument of hg_rotate_fwd / hg_rotate_bwd (NCDHW,
 * S = 16): use 1024-thread CTAs instead of 512.  Results are identical; the library keeps no state. */
#define HG_TUNE_CTA1024 0x100

int hg_abi_version(void);
const char *hg_last_error(void);

/* Tuning options: switches between EQUIVALENT implementations of a pass (A/B measurements, tests of both kernels).
 * Each option NAME is initialised once from the environment variable HG_<NAME> (when the library first needs an
 * option) and changes afterwards only through hg_set_option; `name` may carry the HG_ prefix.
 *   ADAIN_CL_NO_CLUSTER  (0)   1: channels-last AdaIN on the chunked two-kernel path only
 *   ADAIN_CL_CLUSTER_BWD (0)   1: cluster single-pass kernel for the channels-last AdaIN backward
 *   ADAIN_CL_RING        (1)   chunked channels-last AdaIN backward: rows stream through per-thread cp.async rings in shared
 *                              memory (3 stages of 2 rows x 2 tensors in flight per thread) instead of register-staged loads;
 *                              bit-identical results; B200: 45.5 -> 40.4 us (block2), 49.8 -> 44.4 (block3), +1.2 % step
 *   ADAIN_CL_SMALL_REGS  (1)   one-CTA-per-sample channels-last norm kernels with 9-16 rows per thread keep their rows in
 *                              registers: one load pass for statistics + normalise / gradient (D block 0 backward 16.2 -> 10.8 us)
 *   TAPGEMM_DUAL         (-1)  -1 auto; 0 / 1: force one / two tap-GEMM CTAs per SM
 *   FINAL_CONV_MMA       (7)   bit mask of final-layer passes on the mma.sync kernels (1 fwd, 2 dx, 4 dw)
 *   ROTATE_SLAB32        (1)   32^3 rotate forward on source-slab tiles (0: per-channel slab kernel)
 *   ROTATE_GATHER_BWD    (1)   32^3 rotate backward as a table-free per-voxel gather (0: shared-memory scatter)
 *   ADAIN_GEMM_STATS     (0)   host layer: the generator's AdaIN statistics come from the tap-GEMM epilogue
 *                              (hg_convt_fwd_stats + hg_adain_cl_fwd_stats) instead of the single-pass cluster kernel;
 *                              measured break-even on B200 (profiles/r02f_microbench_conv.txt), hence off
 *   TAPGEMM_PERSISTENT   (-1)  tap GEMMs with more tiles than SMs as one persistent CTA per SM that walks the tiles with two
 *                              TMEM accumulator stages (epilogue of tile i overlaps the main loop of tile i + 1).  -1 auto: the
 *                              wide tiles only (256-column MMAs: projection, block3, block4 dgrad); 0 never; 1 always.
 *                              B200 (profiles/r02z_tap_gemm_options.txt): auto is +1.0 % step throughput over 0; forcing it on
 *                              the narrow layers loses (two co-resident one-tile CTAs issue MMAs from two threads)
 *   TAPGEMM_MSUB         (0)   1: wide tap GEMMs (256-column tiles) process two 128-row sub-tiles per CTA that share every
 *                              weight tile of the K loop (a third less L2 traffic per FLOP).  B200 (same file): block3 dgrad
 *                              107 -> 99 us, everything else unchanged or slower (256 tiles on 148 SMs quantise worse than
 *                              512); +0.4 % alone, -0.4 % on top of the persistent default, hence off
 *   TAPGEMM_SHARE_A      (2)   forward of the narrow layers (several parity classes per CTA): every shifted activation box is
 *                              loaded ONCE per K chunk for all classes of the CTA that use it, their weight boxes (at most this
 *                              many, <= 256 rows) are stacked behind each other and one MMA per run of adjacent classes adds
 *                              into their accumulators -- 9 activation boxes instead of 16 (k4, 2-D), 8 instead of 27 (k3, 3-D).
 *                              0: one activation box and one weight box per (class, tap).  B200, B=64 (profiles/r02x_*): block4
 *                              54.9 -> 48.1 us with 2 boxes (50.8 with 4: 48 KB stages leave two per co-resident CTA), block2
 *                              31.7 -> 28.6; the forward of these layers sits at the L2 -> SM throughput cap (~12-14 TB/s) */
int hg_set_option(const char *name, int value);
int hg_get_option(const char *name, int *value);

/* ---- a6 + a7 (+ a8): rigid-body rotate + trilinear resample ------------------------------------
 * Replaces Generator.apply_transformation / interpolation / meshgrid
 * (core/models/hologan_generator.py:198-331) and, with out_layout == HG_PROJ, the projection
 * reshape of :130-133.
 *   vol      (B,C,S,S,S) in `in_layout`, element type `dtype`
 *   a_inv    (B,4,4) fp32 row-major: inverse(Tn @ M @ Tc) of :219-221 (rows 0..2 are used)
 *   out      same element type, `out_layout`, S' == S.  Supported layout pairs: NCDHW -> NCDHW
 *            (fp32 / bf16, S in {8,16,32}) and NDHWC -> NDHWC | PROJ (bf16, channels = 8 * 2^k <= 256)
 *   coords_dbg (optional, may be NULL) (3,B,S^3) fp32: x,y,z source coordinates  ("grid coordinates")
 *   idx_dbg    (optional, may be NULL) (8,B,S^3) int32: flat corner indices a..h incl. the batch
 *              base b*S^3, i.e. idx_a..idx_h of :278-287                        ("sampling indices")
 */
int hg_rotate_fwd(const void *vol, const float *a_inv, void *out, float *coords_dbg, int32_t *idx_dbg,
                  int batch, int channels, int size, int in_layout, int out_layout, int dtype,
                  int border, void *stream);

/* Adjoint of hg_rotate_fwd w.r.t. `vol` (what autograd derives from :292-320).  grad_out is in the
 * forward's out_layout, grad_vol in the forward's in_layout; grad_vol is fully overwritten.  No global
 * atomics.  The channels-last path (in_layout == HG_NDHWC) is a deterministic gather and needs a
 * scratch buffer of hg_rotate_bwd_workspace_bytes() bytes (per-sample cell tables); the NCDHW path
 * needs none (workspace may be NULL).  Out-of-range samples contribute exactly 0 in both border
 * modes (the reference's pairwise-cancelling terms, |residue| ~1e-7). */
long long hg_rotate_bwd_workspace_bytes(int batch, int size, int in_layout);
int hg_rotate_bwd(const void *grad_out, const float *a_inv, void *grad_vol, void *workspace,
                  long long workspace_bytes, int batch, int channels, int size, int in_layout, int out_layout,
                  int dtype, int border, void *stream);

/* ---- a1 + a3 (+ activation): adaptive instance norm ---------------------------------------------
 * Replaces AdaIn (core/models/hologan_generator.py:333-345) fused with the ReLU that follows every
 * call site (:41, :124):  y = act(scale * (x - mean) * rsqrt(var_unbiased + eps) + bias),
 * act(v) = v > 0 ? v : neg_slope * v   (neg_slope = 0 -> ReLU, 1 -> identity / plain AdaIn).
 *   x        (B, C, N) contiguous in N (torch NC* layout); x_batch_stride is the element stride
 *            between samples: C*N normally, 0 for the learned constant of :49-51,121 (never
 *            materialised B times)
 *   scale, bias (B, C) fp32 with row stride `sb_stride` elements (lets both halves of the
 *            ZMapping output, :18, be used in place)
 *   y        (B, C, N) same dtype as x
 *   save_mean, save_rstd (B, C) fp32, written for the backward
 *   biased_var: 0 = unbiased variance (N-1, the generator's AdaIn), 1 = biased (N): with scale = 1,
 *            bias = 0, eps = 1e-5, neg_slope = 0.2 this is the discriminator's
 *            InstanceNorm2d + LeakyReLU (core/models/hologan_discriminator.py:16-17,21-22)
 */
int hg_adain_act_fwd(const void *x, const float *scale, const float *bias, void *y, float *save_mean,
                     float *save_rstd, int batch, int channels, int n, long long x_batch_stride,
                     int sb_stride, float eps, float neg_slope, int biased_var, int dtype, void *stream);

/* Backward of the above.  dy is the gradient w.r.t. the activated output.
 *   dx       (B, C, N), or (C, N) when x_batch_stride == 0 (summed over the batch)
 *   dscale, dbias (B, C) fp32 with row stride `dsb_stride`
 */
int hg_adain_act_bwd(const void *x, const void *dy, const float *scale, const float *bias,
                     const float *save_mean, const float *save_rstd, void *dx, float *dscale, float *dbias,
                     int batch, int channels, int n, long long x_batch_stride, int sb_stride, int dsb_stride,
                     float neg_slope, int biased_var, int dtype, void *stream);


/* Channels-last AdaIN(+activation) for the bf16 pipeline (same arithmetic as hg_adain_act_*).
 *   x  (B, S^ndim, classes, C) bf16: the space-to-depth output of hg_convt_fwd (classes = 2^ndim), or a
 *      plain channels-last tensor (classes = 1)
 *   y  (B, (2S)^ndim, C) bf16 plain channels-last (the depth-to-space shuffle happens in the store);
 *      dy has y's layout, dx has x's layout.  C = 8 * 2^k <= 2048.
 *   scale / bias may both be NULL (= 1 / 0) and dscale / dbias may both be NULL (not computed); with
 *   biased_var = 1, eps = 1e-5, neg_slope = 0.2 that is the discriminator's InstanceNorm2d + LeakyReLU
 *   (core/models/hologan_discriminator.py:16-17,21-22) on channels-last activations.
 *   classes = -4 (ndim 2): x is a plain channels-last (B, S, S, C) tensor and y (and dy) are stored in 2x2
 *   space-to-depth order (B, S/2, S/2, 4, C), y[b, i, j, (py, px), c] = pixel (2i+py, 2j+px) -- the input layout of
 *   hg_conv5s2_fwd, so the discriminator's blocks chain without a layout pass.
 *   workspace: hg_adain_cl_workspace_bytes() bytes of scratch for the per-chunk partial sums (0 for small
 *   instances, then it may be NULL); deterministic (fixed summation order). */
long long hg_adain_cl_workspace_bytes(int batch, int channels, int ndim, int size, int classes);
int hg_adain_cl_fwd(const void *x, const float *scale, const float *bias, void *y, float *save_mean, float *save_rstd,
                    void *workspace, long long workspace_bytes, int batch, int channels, int ndim, int size, int classes,
                    int sb_stride, float eps, float neg_slope, int biased_var, void *stream);
int hg_adain_cl_bwd(const void *x, const void *dy, const float *scale, const float *bias, const float *save_mean,
                    const float *save_rstd, void *dx, float *dscale, float *dbias, void *workspace, long long workspace_bytes,
                    int batch, int channels, int ndim, int size, int classes, int sb_stride, int dsb_stride, float neg_slope,
                    int biased_var, void *stream);
/* hg_adain_cl_fwd with the statistics partials of hg_convt_fwd_stats (same geometry: x = that call's y_s2d): a merge
 * kernel (Chan's update over the 32-row groups x classes of each instance, fixed order) + one streaming pass.  The
 * statistics are those of the fp32 GEMM results, i.e. before x was rounded to bf16.  S^ndim % 32 == 0. */
int hg_adain_cl_fwd_stats(const void *x, const float *stats, const float *scale, const float *bias, void *y, float *save_mean,
                          float *save_rstd, int batch, int channels, int ndim, int size, int classes, int sb_stride, float eps,
                          float neg_slope, int biased_var, void *stream);

/* ---- a4 / a9 / a10: transposed convolutions and the 1x1 projection as tcgen05 implicit GEMMs -------
 * Replace nn.ConvTranspose3d(k3,s2,p1,op1) / nn.ConvTranspose2d(k4,s2,p1) / nn.ConvTranspose2d(k1)
 * forward and backward (core/models/hologan_generator.py:25-30, 38, 60, 135).  kernel = 5 (2-D: s2, p2, op1) is the
 * transposed convolution whose dgrad / forward / wgrad are the forward / dgrad / wgrad of the discriminator's
 * Conv2d(k5, s2, p2) (core/models/hologan_discriminator.py:12) on a space-to-depth input.  bf16 operands, fp32
 * accumulation in TMEM.  Layouts (all channels-last, bf16):
 *   x      (B, [S,] S, S, Cin)                       `ndim` spatial dims of extent `size`
 *   y_s2d  (B, [S,] S, S, P, Cout)   P = 2^ndim parity classes (P = 1 for kernel 1):
 *          y_s2d[b, i.., (pz,py,px), co] == y[b, co, 2*iz+pz, 2*iy+py, 2*ix+px] of the torch op
 *   packed weights (hg_convt_pack_weight) from the torch parameter (Cin, Cout, k..) fp32:
 *          w_fwd [t][Cout][Cin], w_dgrad [t][Cin][Cout], t = flat kernel index (kz*k + ky)*k + kx
 * Supported: Cin % 64 == 0, Cout % 16 == 0 (fwd); Cout % 64 == 0 (dgrad); Cin % 128 == 0 and
 * Cout % 64 == 0 (wgrad); size in {4, 8, 16, 32, ...} tiling into 128-row boxes.
 */
/* perm_c / perm_s: optional permutation of the input-channel (GEMM K) index for the projection operand
 * in HG_PROJ layout: packed channel y*perm_c + c <-> torch channel c*perm_s + (perm_s-1-y)
 * (reference :130-133); perm_s = 0 means identity.  Cin even, Cout % 32 == 0, taps in {1, 16, 25, 27}. */
int hg_convt_pack_weight(const float *w, void *w_fwd, void *w_dgrad, int cin, int cout, int taps, int perm_c, int perm_s,
                         void *stream);
/* y_s2d = act(convT(x) + bias); bias (Cout) fp32 or NULL; act: v > 0 ? v : neg_slope * v (1 = none) */
int hg_convt_fwd(const void *x, const void *w_fwd, const float *bias, void *y_s2d, int batch, int cin, int cout,
                 int ndim, int size, int kernel, float neg_slope, void *stream);
/* hg_convt_fwd that also emits the AdaIN statistics of its output from the epilogue (north_star: "AdaIN fused into
 * the GEMM epilogues"; SURVEY 8b epilogue flag adain-stats): for every group of 32 consecutive output positions g and
 * every s2d column col = cls * Cout + co, stats[(g * P * Cout + col) * 2 + {0, 1}] = sum / sum of squares of the fp32
 * results (after bias, before the activation).  `stats` holds hg_convt_stats_floats() floats.  hg_adain_cl_fwd_stats
 * merges them per instance, so the convolution output is read once (by the normalise pass) instead of twice.
 * Cout % 32 == 0. */
long long hg_convt_stats_floats(int batch, int cout, int ndim, int size, int kernel);
int hg_convt_fwd_stats(const void *x, const void *w_fwd, const float *bias, void *y_s2d, float *stats, int batch, int cin,
                       int cout, int ndim, int size, int kernel, float neg_slope, void *stream);
/* dx (B, .., Cin) = adjoint of the forward w.r.t. x */
int hg_convt_dgrad(const void *dy_s2d, const void *w_dgrad, void *dx, int batch, int cin, int cout, int ndim, int size,
                   int kernel, void *stream);
/* dw (Cin, Cout, k..) fp32 in the torch parameter layout = adjoint w.r.t. the weight (accumulate != 0:
 * added to dw, else overwritten).  Split-K partial tiles go to `workspace` with plain stores and are
 * summed in a fixed order by a second kernel: deterministic, no atomics. */
long long hg_convt_wgrad_workspace_bytes(int batch, int cin, int cout, int ndim, int size, int kernel);
int hg_convt_wgrad(const void *x, const void *dy_s2d, float *dw, void *workspace, long long workspace_bytes, int batch,
                   int cin, int cout, int ndim, int size, int kernel, int perm_c, int perm_s, int accumulate, void *stream);

/* Backward of the bias + activation fused into hg_convt_fwd's epilogue (the projection's relu(conv1x1(x)),
 * core/models/hologan_generator.py:135-136), one pass:
 *   dpre[m,n] = y[m,n] > 0 ? dy[m,n] : neg_slope * dy[m,n]   (rows x cols, bf16, row-major, cols % 8 == 0)
 *   dbias[n]  = sum_m dpre[m,n]  (fp32; NULL = not needed, then workspace may be NULL)
 * Deterministic two-stage column sum through `workspace`. */
long long hg_act_bwd_bias_workspace_bytes(long long rows, int cols);
int hg_act_bwd_bias(const void *y, const void *dy, void *dpre, float *dbias, void *workspace, long long workspace_bytes,
                    long long rows, int cols, float neg_slope, void *stream);

/* Plain GEMM on the same pipeline: D[M,N] = act(A[M,K] @ B[N,K]^T + bias[N]); bf16 row-major A, B, D
 * (row stride of D = ldd elements).  Used for the batched ZMapping (a2). */
int hg_gemm_bf16_nt(const void *a, const void *b, const float *bias, void *d, int m, int n, int k, long long ldd,
                    float neg_slope, void *stream);

/* ---- a2: ZMapping  relu(Linear(z))  (core/models/hologan_generator.py:7-18), fp32 ------------------
 *   z (B,K), w (N,K) torch nn.Linear layout, bias (N), out (B,N) = relu(z @ w^T + bias)
 *   backward: dw (N,K), dbias (N), dz (B,K) (+= when accumulate_dz, may be NULL) from dout (B,N) */
int hg_linear_relu_fwd(const float *z, const float *w, const float *bias, float *out, int batch, int k, int n, void *stream);
int hg_linear_relu_bwd(const float *z, const float *w, const float *out, const float *dout, float *dw, float *dbias,
                       float *dz, int batch, int k, int n, int accumulate_dz, void *stream);

/* Grouped form: `layers` (<= HG_LINEAR_GROUP_MAX) ZMappings that read the same z (the generator's five, reference
 * :34,54) in one launch each way.  w / bias / out / dout / dw / dbias and n are HOST arrays of `layers` entries
 * (device pointers / feature counts); out[l] is (B, n[l]).  The backward gives dw, dbias only (z carries no gradient
 * in training). */
int hg_linear_relu_group_fwd(int layers, const float *z, const float *const *w, const float *const *bias, float *const *out,
                             const int *n, int batch, int k, void *stream);
int hg_linear_relu_group_bwd(int layers, const float *z, const float *const *out, const float *const *dout, float *const *dw,
                             float *const *dbias, const int *n, int batch, int k, void *stream);

/* ---- a11: final_layer + tanh  (core/models/hologan_generator.py:69-75,141-142, img_size 64) ---------
 *   out (B,Cout,S,S) fp32 NCHW = tanh(conv2d(x, w, bias, k3, p1)); x (B,S,S,Cin) bf16 NHWC,
 *   w torch (Cout,Cin,3,3) fp32.  Cout <= 4, Cin = 8 * 2^k <= 256.  Direct convolution (bandwidth-bound).
 *   backward: dx (B,S,S,Cin) bf16 (may be NULL), dw, dbias (both NULL: input gradient only -- the two halves are
 *   independent and may be issued as two calls on two streams, each with its own workspace) from dout (B,Cout,S,S) fp32;
 *   deterministic two-stage reduction through `workspace`. */
int hg_final_conv_tanh_fwd(const void *x, const float *w, const float *bias, float *out, int batch, int cin, int cout,
                           int size, void *stream);
long long hg_final_conv_tanh_bwd_workspace_bytes(int batch, int cin, int cout, int size);
int hg_final_conv_tanh_bwd(const void *x, const float *w, const float *out, const float *dout, void *dx, float *dw,
                           float *dbias, void *workspace, long long workspace_bytes, int batch, int cin, int cout, int size,
                           void *stream);

/* The patched 128 x 128 head (the reference's img_size == 128 branch, core/models/hologan_generator.py:71-72,
 * lacks stride = 2 and yields 65 x 65; the working variant is ConvTranspose2d(64 -> 3, k4, s2, p1)):
 *   out (B, 3, S, S) fp32 NCHW = tanh(convT(x) + bias); x (B, S/2, S/2, 64) bf16 NHWC; w (64, 3, 4, 4) fp32 torch layout.
 * Warp-level tensor cores (mma.sync) -- the kernels are shared with the discriminator's first convolution (disc_ends.cu).
 * backward: g = dout * (1 - out^2); dx (B, S/2, S/2, 64) bf16 (NULL = not needed); dw (64, 3, 4, 4) + dbias (3) fp32 (both
 * or neither).  S % 32 == 0.  Deterministic. */
int hg_head128_fwd(const void *x, const float *w, const float *bias, float *out, int batch, int cin, int cout, int size,
                   void *stream);
long long hg_head128_bwd_workspace_bytes(int batch, int size);
int hg_head128_bwd(const void *x, const float *w, const float *out, const float *dout, void *dx, float *dw, float *dbias,
                   void *workspace, long long workspace_bytes, int batch, int cin, int cout, int size, void *stream);

/* ---- a14: spectral normalisation of the discriminator's convolutions, grouped over the layers ----------
 * Replaces torch.nn.utils.spectral_norm (n_power_iterations = 1, eps = 1e-12, dim = 0) as used at
 * core/models/hologan_discriminator.py:15 -- per layer i, with W = w[i] viewed as (cout, K = cin * taps):
 *     t = W^T u;  v = t / max(|t|, eps);  s = W v;  u = s / max(|s|, eps);  sigma = u . s;  w_out = W / sigma
 * u[i] (cout) and v[i] (K, LOGICAL order (ci, tap), as in the reference's state_dict) are updated in place when
 * power_iteration != 0 (training-mode forward) and only read otherwise (eval).  W, w_out, dw, dw_orig are in the
 * parameter's PHYSICAL element order: channels_last != 0 means (cout, tap, cin), else (cout, cin, tap).
 * state[i]: hg_spectral_norm_state_floats() floats written by the forward ([0] sigma, [1] 1/sigma, the u and v used)
 * -- the backward reads it, so the buffers may be overwritten by the next forward in between.
 * backward: dw_orig = dw / sigma - (sum(dw * W) / sigma^2) * u v^T   (+= when accumulate != 0); dw has dw_dtype.
 * The arrays of pointers / dims are HOST arrays of `layers` (<= HG_SN_MAX_LAYERS) entries.  Deterministic. */
long long hg_spectral_norm_state_floats(int cout, int cin, int taps);
long long hg_spectral_norm_workspace_bytes(int layers, const int *cout, const int *cin, const int *taps);
int hg_spectral_norm_fwd(int layers, const float *const *w, float *const *u, float *const *v, void *const *w_out,
                         float *const *state, const int *cout, const int *cin, const int *taps, int channels_last,
                         int power_iteration, float eps, int out_dtype, void *workspace, long long workspace_bytes, void *stream);
int hg_spectral_norm_bwd(int layers, const void *const *dw, const float *const *w, const float *const *state,
                         float *const *dw_orig, const int *cout, const int *cin, const int *taps, int accumulate, int dw_dtype,
                         void *workspace, long long workspace_bytes, void *stream);

/* w_out[i] may be NULL: only the power iteration and sigma (state[i][0], a device float) are produced -- the caller
 * then folds 1 / sigma into its own weight pack (hg_conv5s2_pack_weight). */

/* ---- a14 (f1): the discriminator's convolutions and heads on hand-written kernels -----------------------
 * Conv2d(k5, s2, p2) of the three spectral-norm blocks (core/models/hologan_discriminator.py:12,20) on the tcgen05 tap
 * GEMMs.  On a 2x2 space-to-depth input every kernel tap is a shift in {-1,0,1}^2 of one parity class, i.e. the op is
 * the dgrad of the dual ConvTranspose2d(k5,s2,p2,op1) (kernel = 5 of hg_convt_*).  These layers have few output
 * positions and a long K, so the K loop is split over taps across CTAs: fp32 partial tiles in `workspace`, summed in a
 * fixed order by a second kernel (deterministic, no atomics).  No bias: InstanceNorm follows (a per-channel constant
 * cancels).  bf16, channels-last:
 *   x_s2d  (B, S, S, 4, Cin)  x_s2d[b, i, j, (py, px), c] = x[b, c, 2i+py, 2j+px];  y / dy (B, S, S, Cout), S = size_out
 *   hg_conv5s2_pack_weight: torch Conv2d weight (Cout, Cin, 5, 5) fp32 contiguous, divided by *sigma (DEVICE float,
 *   NULL = 1: the spectral norm of :15) -> w_k [25][Cout][Cin] (forward), w_t [25][Cin][Cout] (dx); either may be NULL
 *   out_scale: DEVICE float or NULL -- the result of forward / dx is multiplied by it in the epilogue.  With 1 / sigma
 *   there (state[1] of hg_spectral_norm_fwd) the packed weights can stay the UN-normalised weight_orig (packed once per
 *   optimizer step, sigma = NULL) although sigma moves with every forward: conv(x, W / sigma) = conv(x, W) / sigma.
 *   hg_conv5s2_dw: dw (Cout, Cin, 5, 5) fp32 w.r.t. the (normalised) weight; accumulate != 0 adds.
 * Supported: Cin % 64 == 0, Cout % 128 == 0, B * S * S % 128 == 0 with S a power of two. */
long long hg_conv5s2_workspace_bytes(int batch, int cin, int cout, int size_out);
int hg_conv5s2_pack_weight(const float *w, const float *sigma, void *w_k, void *w_t, int cin, int cout, void *stream);
int hg_conv5s2_fwd(const void *x_s2d, const void *w_k, const float *out_scale, void *y, void *workspace,
                   long long workspace_bytes, int batch, int cin, int cout, int size_out, void *stream);
int hg_conv5s2_dx(const void *dy, const void *w_t, const float *out_scale, void *dx_s2d, void *workspace,
                  long long workspace_bytes, int batch, int cin, int cout, int size_out, void *stream);
int hg_conv5s2_dw(const void *dy, const void *x_s2d, float *dw, void *workspace, long long workspace_bytes, int batch, int cin,
                  int cout, int size_out, int accumulate, void *stream);

/* First convolution of the discriminator: y = leaky_relu(Conv2d(3 -> 64, k5, s2, p2)(x) + bias, neg_slope)
 * (core/models/hologan_discriminator.py:30,58).  x (B, 3, S, S) fp32 NCHW (the real batch, or the generator's output);
 * y_s2d (B, S/4, S/4, 4, 64) bf16: the (S/2 x S/2) activation in the space-to-depth order hg_conv5s2_fwd reads.
 * Warp-level tensor cores (mma.sync, bf16 operands, fp32 accumulation).  S % 32 == 0.
 * backward: dy_s2d in y's layout; dx (B, 3, S, S) fp32 (NULL = not needed: real images), dw (64, 3, 5, 5) + dbias (64)
 * fp32 (both NULL = not needed: generator step; accumulate != 0 adds to them).  workspace: hg_dconv0_bwd_workspace_bytes(). */
int hg_dconv0_fwd(const float *x, const float *w, const float *bias, void *y_s2d, int batch, int cin, int cout, int size,
                  float neg_slope, void *stream);
long long hg_dconv0_bwd_workspace_bytes(int batch, int size);
int hg_dconv0_bwd(const float *x, const float *w, const void *y_s2d, const void *dy_s2d, float *dx, float *dw, float *dbias,
                  void *workspace, long long workspace_bytes, int batch, int cin, int cout, int size, float neg_slope,
                  int accumulate, void *stream);

/* The two heads (core/models/hologan_discriminator.py:41-51,64-68) on the last block's channels-last activation
 * h (B, H*W, C) bf16 (the reference flattens (c, h, w): weights are addressed with feature f = c * HW + hw):
 *   logits (B) = linear1(h);  t2 (B, 128) = leaky_relu(linear2(h), neg_slope);  z_pred (B, zdim) = tanh(linear3(t2))
 * w1 (1, F), w2 (128, F), w3 (zdim, 128) fp32 torch layouts, F = C * HW; outputs fp32.  batch <= 64 per call, HW | 64.
 * backward: dlogits (B) / dz_pred (B, zdim) fp32 (either may be NULL = zero); dh (B, HW, C) bf16 (may be NULL); the six
 * parameter gradients are overwritten, all given or all NULL (generator step).  workspace: hg_dheads_workspace_bytes(). */
long long hg_dheads_workspace_bytes(int batch, int channels, int hw, int zdim);
int hg_dheads_fwd(const void *h, const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                  const float *b3, float *logits, float *t2, float *z_pred, void *workspace, long long workspace_bytes,
                  int batch, int channels, int hw, int zdim, float neg_slope, void *stream);
int hg_dheads_bwd(const void *h, const float *w1, const float *w2, const float *w3, const float *t2, const float *z_pred,
                  const float *dlogits, const float *dz_pred, void *dh, float *dw1, float *db1, float *dw2, float *db2,
                  float *dw3, float *db3, void *workspace, long long workspace_bytes, int batch, int channels, int hw, int zdim,
                  float neg_slope, void *stream);

/* ---- a13: the optimizer step.  torch.optim.Adam(lr, betas) of core/lightning_module.py:75-87 (no weight decay, no
 * amsgrad) over ONE flat fp32 buffer holding every parameter of a network (param / grad / exp_avg / exp_avg_sq are
 * parallel buffers of n floats, n % 4 == 0, 16-byte aligned).  grad is multiplied by grad_scale first (1 / world size
 * of the data-parallel gradient SUM).  state: 4 device floats owned by the caller, state[0] = steps taken so far (the
 * call increments it); lr: device float.  Both live on the device so that the step replays from a CUDA graph.
 * Arithmetic as torch's fused kernel: m += (g - m)(1 - b1); v = b2 v + (1 - b2) g^2;
 * p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps). */
int hg_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, float *state,
                 const float *lr, float beta1, float beta2, float eps, float grad_scale, void *stream);
/* The same step in two parts, for an optimizer that runs bucket by bucket behind the gradient exchange (the data-parallel
 * harness applies a bucket on a side stream as soon as its all-reduce has finished, while the backward pass continues --
 * run_network.py:66-71's DDP overlap, extended to the update): hg_adam_tick advances state[0] and the bias corrections
 * ONCE per step; hg_adam_apply updates any 16-byte aligned sub-range (n % 4 == 0) of the four flat buffers with them. */
int hg_adam_tick(float *state, const float *lr, float beta1, float beta2, void *stream);
int hg_adam_apply(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, const float *state,
                  float beta1, float beta2, float eps, float grad_scale, void *stream);

/* ---- a13: the losses of HOLOGAN.training_step  (core/lightning_module.py:217-237) -------------------
 *   adv = wa * mean_i BCEWithLogits(a[i], ta) + wb * mean_j BCEWithLogits(b[j], tb)   (b may be NULL, nb = 0)
 *   q   = mean_k (zp[k] - z[k])^2
 *   total[0] = adv + q;  parts[0] = adv, parts[1] = q.
 * D step (:219-229): a = D(real) logits, ta = 1, wa = 0.5, b = D(fake) logits, tb = 0, wb = 0.5.
 * G step (:231-237): a = D(fake) logits, ta = 1, wa = 1, nb = 0.
 * a, b, zp have `dtype` (HG_F32 / HG_BF16), z is fp32; arithmetic in fp32, fixed summation order.
 * backward: da, db, dzp (same dtype / sizes as a, b, zp) = d total / d(.) * gout[0]  (gout: device fp32 scalar,
 * NULL = 1). */
int hg_gan_loss_fwd(const void *a, int na, float ta, float wa, const void *b, int nb, float tb, float wb, const void *zp,
                    const float *z, int nz, int dtype, float *total, float *parts, void *stream);
int hg_gan_loss_bwd(const float *gout, const void *a, int na, float ta, float wa, const void *b, int nb, float tb, float wb,
                    const void *zp, const float *z, int nz, int dtype, void *da, void *db, void *dzp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HOLOGAN_B200_H */
