// Adam over ONE flat fp32 parameter buffer (reference: torch.optim.Adam(lr 1e-4, betas (0.9, 0.999)) for both
// networks, core/lightning_module.py:75-87 + conf/expt/hologan.yaml:14-23; no weight decay, no amsgrad).
// The HoloGAN trainer keeps every parameter of a network, its gradient and both moments as views of four flat buffers,
// so the whole optimizer step is one streaming kernel (28 bytes per parameter: 4 reads + 3 writes) instead of a
// multi-tensor library launch, and the data-parallel gradient average 1 / world is folded in (`grad_scale`) instead of
// a separate pass over the gradients.  The step counter and the learning rate live in device memory so that the step
// can be replayed from a CUDA graph; arithmetic follows torch's fused kernel:
//     m = m + (g - m) * (1 - b1);  v = b2 * v + (1 - b2) * g * g
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "hg_common.cuh"

namespace hg {

// state[0] = step t (as float, exact up to 2^24), state[1] = lr / (1 - b1^t), state[2] = 1 / sqrt(1 - b2^t)
__global__ void adam_tick_kernel(float *__restrict__ state, const float *__restrict__ lr, float b1, float b2)
{
    const float t = state[0] + 1.f;
    state[0] = t;
    state[1] = __ldg(lr) / (1.f - powf(b1, t));
    state[2] = rsqrtf(1.f - powf(b2, t));
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, long long n4, const float *__restrict__ state, float b1,
                                                   float b2, float eps, float grad_scale)
{
    const float step_size = __ldg(state + 1), inv_bc2_sqrt = __ldg(state + 2);
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        float *pf = &pp.x, *mf = &mm.x, *vf = &vv.x;
        const float *gf = &gg.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gf[j] * grad_scale;
            mf[j] = mf[j] + (gr - mf[j]) * (1.f - b1);
            vf[j] = b2 * vf[j] + (1.f - b2) * gr * gr;
            pf[j] -= step_size * mf[j] / (sqrtf(vf[j]) * inv_bc2_sqrt + eps);
        }
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
}

}  // namespace hg

using namespace hg;

// state: 4 device floats owned by the caller ([0] = number of steps taken so far; zero it to reset).
// hg_adam_tick advances the step counter and the bias-correction factors ONCE per optimizer step; hg_adam_apply then updates
// any 16-byte aligned sub-range of the flat buffers (the trainer applies a gradient bucket as soon as its all-reduce has
// finished, on a side stream, while the rest of the backward still runs).  hg_adam_step = tick + apply on one range.
extern "C" int hg_adam_tick(float *state, const float *lr, float beta1, float beta2, void *stream)
{
    HG_REQUIRE(state && lr, HG_ERR_INVALID_ARG, "hg_adam_tick: null pointer");
    adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state, lr, beta1, beta2);
    return check_launch("hg_adam_tick");
}

// n must be a multiple of 4 and the four buffers 16-byte aligned (the trainer pads its flat buffers).
extern "C" int hg_adam_apply(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, const float *state,
                             float beta1, float beta2, float eps, float grad_scale, void *stream)
{
    HG_REQUIRE(param && grad && exp_avg && exp_avg_sq && state, HG_ERR_INVALID_ARG, "hg_adam_apply: null pointer");
    HG_REQUIRE(n > 0 && n % 4 == 0, HG_ERR_INVALID_ARG, "hg_adam_apply: n must be a positive multiple of 4 (got %lld)", n);
    HG_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, HG_ERR_INVALID_ARG, "hg_adam_apply: buffers must be 16-byte aligned");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n4, state, beta1, beta2, eps,
                                                                               grad_scale);
    return check_launch("hg_adam_apply");
}

extern "C" int hg_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, float *state,
                            const float *lr, float beta1, float beta2, float eps, float grad_scale, void *stream)
{
    HG_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && lr, HG_ERR_INVALID_ARG, "hg_adam_step: null pointer");
    int rc = hg_adam_tick(state, lr, beta1, beta2, stream);
    if (rc) return rc;
    return hg_adam_apply(param, grad, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, grad_scale, stream);
}
