"""The discriminator on hand-written kernels: first convolution, spectral-norm 5x5 stride-2 convolutions
on the tcgen05 tap GEMMs (split-K over taps), InstanceNorm + LeakyReLU with space-to-depth store, the two heads -- each
op against torch's fp32 arithmetic on the same (bf16-rounded) operands, then the whole network against the fp32 oracle
with the bf16 storage floor (oracle.discriminator_forward(bf16_storage=True)) as the bar for the gradients."""
import copy
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from lightning_gan_zoo_b200 import _lib, ops
from lightning_gan_zoo_b200.core.models.hologan_discriminator import Discriminator
from oracle import hologan_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def s2d(x_nchw):
    """(B, C, H, W) -> (B, H/2, W/2, 4, C): x_s2d[b, i, j, (py, px), c] = x[b, c, 2i+py, 2j+px]."""
    b, c, h, w = x_nchw.shape
    return x_nchw.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(b, h // 2, w // 2, 4, c).contiguous()


def un_s2d(t):
    """inverse of s2d: (B, H/2, W/2, 4, C) -> (B, C, H, W)."""
    b, h2, w2, _, c = t.shape
    return t.reshape(b, h2, w2, 2, 2, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, 2 * h2, 2 * w2)


def rms_err(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


@pytest.mark.parametrize("batch,size", [(4, 64), (64, 64), (3, 128)])
def test_dconv0_fwd_bwd(batch, size):
    """Conv2d(3 -> 64, k5, s2, p2) + bias + LeakyReLU on mma.sync, and its backward (dx, dw, dbias), against torch
    fp32 on the bf16-rounded image / weights (reference core/models/hologan_discriminator.py:30,58)."""
    g = torch.Generator().manual_seed(batch + size)
    x = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).to(DEV)
    w = (torch.randn(64, 3, 5, 5, generator=g) * 0.05).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    xr = x.to(BF).float().requires_grad_(True)
    wr = w.to(BF).float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xr, wr, br, stride=2, padding=2), 0.2)
    xg, wg, bg = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y = ops.dconv0(xg, wg, bg, 0.2)
    assert tuple(y.shape) == (batch, size // 4, size // 4, 4, 64) and y.dtype == BF
    assert rel_err(un_s2d(y).float(), ref) < 2 ** -7
    dy = torch.randn(ref.shape, generator=g).to(BF).to(DEV)
    # the kernel's LeakyReLU mask comes from ITS bf16 output; use the same mask in the reference (pre-activations that
    # round across zero are measure-zero noise, not a kernel property)
    ref.backward(dy.float())
    (y.float() * s2d(dy).float()).sum().backward()
    assert rel_err(bg.grad, br.grad) < 2e-3
    assert rel_err(wg.grad, wr.grad) < 2e-3        # bf16 rounding of dpre = dy * 0.2 where the slope applies
    assert rel_err(xg.grad, xr.grad) < 2 ** -7     # weights and dpre enter the MMA as bf16
    # deterministic
    xg2, wg2, bg2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y2 = ops.dconv0(xg2, wg2, bg2, 0.2)
    (y2.float() * s2d(dy).float()).sum().backward()
    assert torch.equal(y, y2) and torch.equal(wg.grad, wg2.grad) and torch.equal(xg.grad, xg2.grad)


CONV5_CASES = [  # (batch, cin, cout, size_out)
    (8, 64, 128, 16), (8, 128, 256, 8), (8, 256, 512, 4),          # the three blocks at a small batch
    (64, 64, 128, 16), (64, 128, 256, 8), (64, 256, 512, 4),       # BASELINE cfg 2 (batch 64, 64x64): split-K plans of the bench
    (32, 64, 128, 32), (32, 256, 512, 8),                          # cfg 4 (batch 32, 128x128)
    (3, 128, 256, 8),                                              # ragged: last 128-row box partly out of range
]


@pytest.mark.parametrize("batch,cin,cout,size", CONV5_CASES)
def test_conv5s2_sn_fwd_dx_dw(batch, cin, cout, size):
    """Conv2d(k5, s2, p2) on the tap GEMMs with 1 / sigma folded into the weight pack: forward, dx and dw (w.r.t. the
    normalised weight, through the spectral-norm backward with an identity check) against torch fp32 on the bf16
    operands."""
    g = torch.Generator().manual_seed(cin + cout + batch)
    x = torch.randn(batch, cin, 2 * size, 2 * size, generator=g).to(BF).to(DEV)
    w = (torch.randn(cout, cin, 5, 5, generator=g) * 0.03).to(DEV)
    sigma = 1.7
    state = torch.zeros(int(_lib.load().hg_spectral_norm_state_floats(cout, cin, 25)), device=DEV)
    state[0], state[1] = sigma, 1.0 / sigma
    wn = torch.div(w.to(BF).float(), state[0])              # bf16 copy of the un-normalised weight, 1 / sigma in the epilogue
    xr = x.float().requires_grad_(True)
    wr = wn.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride=2, padding=2)
    dy = torch.randn(ref.shape, generator=g).to(BF).to(DEV)
    ref.backward(dy.float())
    # --- raw ABI calls: forward, dx, dw (gradient w.r.t. the normalised weight)
    P, st = ops._ptr, ops._stream()
    wk = torch.empty(25, cout, cin, dtype=BF, device=DEV)
    wt = torch.empty(25, cin, cout, dtype=BF, device=DEV)
    _lib.call("hg_conv5s2_pack_weight", P(w), P(None), P(wk), P(wt), cin, cout, st)
    assert torch.equal(wk, w.to(BF).reshape(cout, cin, 25).permute(2, 0, 1))
    assert torch.equal(wt, w.to(BF).reshape(cout, cin, 25).permute(2, 1, 0))
    wk2 = torch.empty_like(wk)
    _lib.call("hg_conv5s2_pack_weight", P(w), P(state), P(wk2), P(None), cin, cout, st)        # sigma folded into the pack
    assert torch.equal(wk2, torch.div(w, state[0]).to(BF).reshape(cout, cin, 25).permute(2, 0, 1))
    inv_sigma = ctypes.c_void_p(state.data_ptr() + 4)
    nb = _lib.load().hg_conv5s2_workspace_bytes(batch, cin, cout, size)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=DEV)
    xs = s2d(x)
    y = torch.empty(batch, size, size, cout, dtype=BF, device=DEV)
    _lib.call("hg_conv5s2_fwd", P(xs), P(wk), inv_sigma, P(y), P(ws), nb, batch, cin, cout, size, st)
    assert rel_err(y.permute(0, 3, 1, 2).float(), ref) < 2 ** -7
    dy_cl = dy.permute(0, 2, 3, 1).contiguous()
    dx = torch.empty_like(xs)
    _lib.call("hg_conv5s2_dx", P(dy_cl), P(wt), inv_sigma, P(dx), P(ws), nb, batch, cin, cout, size, st)
    assert rel_err(un_s2d(dx).float(), xr.grad) < 2 ** -7
    dw = torch.empty_like(w)
    _lib.call("hg_conv5s2_dw", P(dy_cl), P(xs), P(dw), P(ws), nb, batch, cin, cout, size, 0, st)
    dw2 = torch.empty_like(w)
    _lib.call("hg_conv5s2_dw", P(dy_cl), P(xs), P(dw2), P(ws), nb, batch, cin, cout, size, 0, st)
    assert torch.equal(dw, dw2)                    # fixed summation order
    assert rel_err(dw, wr.grad) < 1e-4             # fp32 accumulation of exact bf16 products
    y2 = torch.empty_like(y)
    _lib.call("hg_conv5s2_fwd", P(xs), P(wk), inv_sigma, P(y2), P(ws), nb, batch, cin, cout, size, st)
    assert torch.equal(y, y2)


def test_conv5s2_sn_autograd_matches_torch_spectral_norm():
    """ops.spectral_norm_sigma + ops.conv5s2_sn against torch.nn.utils.spectral_norm(Conv2d) in fp32: u / v buffers,
    output, and the gradient w.r.t. weight_orig (spectral-norm backward included)."""
    torch.manual_seed(3)
    conv = torch.nn.utils.spectral_norm(torch.nn.Conv2d(64, 128, 5, stride=2, padding=2, bias=False)).to(DEV)
    w = conv.weight_orig.detach().clone().requires_grad_(True)
    u, v = conv.weight_u.detach().clone(), conv.weight_v.detach().clone()
    x = torch.randn(8, 64, 32, 32, device=DEV).to(BF)
    xr = x.float().requires_grad_(True)
    conv.train()
    ref = conv(xr)
    dy = torch.randn_like(ref).to(BF)
    ref.backward(dy.float())
    (state,) = ops.spectral_norm_sigma([w], [u], [v], power_iteration=True)
    assert rel_err(u, conv.weight_u) < 1e-5 and rel_err(v, conv.weight_v) < 1e-5
    xs = s2d(x).requires_grad_(True)
    y = ops.conv5s2_sn(xs, w, state)
    assert rel_err(y.permute(0, 3, 1, 2).float(), ref) < 2 ** -6       # bf16 weights (the reference's are fp32)
    (y.float() * dy.permute(0, 2, 3, 1).float()).sum().backward()
    assert rel_err(w.grad, conv.weight_orig.grad) < 1e-2               # bf16 x and dy products, fp32 spectral-norm backward
    assert rel_err(un_s2d(xs.grad).float(), xr.grad) < 2 ** -6


@pytest.mark.parametrize("batch,size,c", [(8, 16, 128), (64, 8, 256), (5, 32, 128)])
def test_instance_norm_lrelu_s2d_store(batch, size, c):
    """InstanceNorm2d + LeakyReLU whose store (and the backward's dy read) is in 2x2 space-to-depth order."""
    g = torch.Generator().manual_seed(size + c)
    x = torch.randn(batch, size, size, c, generator=g).to(BF).to(DEV)
    dy = torch.randn(batch, size // 2, size // 2, 4, c, generator=g).to(BF).to(DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    plain = ops.instance_norm_act_channels_last(xa, 0.2, 1e-5)
    sd = ops.instance_norm_act_channels_last(xb, 0.2, 1e-5, s2d_out=True)
    assert tuple(sd.shape) == (batch, size // 2, size // 2, 4, c)
    assert torch.equal(un_s2d(sd), plain.permute(0, 3, 1, 2))
    (sd.float() * dy.float()).sum().backward()
    (plain.float() * un_s2d(dy).permute(0, 2, 3, 1).float()).sum().backward()
    assert torch.equal(xa.grad, xb.grad)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.leaky_relu(F.instance_norm(xr, eps=1e-5), 0.2)
    assert rel_err(plain.permute(0, 3, 1, 2).float(), ref) < 2 ** -7


@pytest.mark.parametrize("batch,hw,zdim", [(8, 16, 128), (64, 16, 128), (32, 64, 128), (5, 16, 96)])
def test_dheads_fwd_bwd(batch, hw, zdim):
    """linear1 / linear2 / linear3 heads on the channels-last activation (feature f = c * HW + hw of the reference's
    (c, h, w) flatten, core/models/hologan_discriminator.py:60-68) against torch fp32."""
    g = torch.Generator().manual_seed(batch + hw)
    side = int(hw ** 0.5)
    c = 512
    f = c * hw
    h = torch.randn(batch, side, side, c, generator=g).to(BF).to(DEV)
    ps = [(torch.randn(1, f, generator=g) * 0.02), torch.randn(1, generator=g) * 0.1,
          (torch.randn(128, f, generator=g) * 0.02), torch.randn(128, generator=g) * 0.1,
          (torch.randn(zdim, 128, generator=g) * 0.1), torch.randn(zdim, generator=g) * 0.1]
    ps = [p.to(DEV) for p in ps]
    pr = [p.clone().requires_grad_(True) for p in ps]
    hr = h.float().requires_grad_(True)
    flat = hr.permute(0, 3, 1, 2).reshape(batch, -1)                    # the reference's (c, h, w) flatten
    lr = F.linear(flat, pr[0], pr[1])
    zr = torch.tanh(F.linear(F.leaky_relu(F.linear(flat, pr[2], pr[3]), 0.2), pr[4], pr[5]))
    dl = torch.randn(batch, 1, generator=g).to(DEV)
    dz = torch.randn(batch, zdim, generator=g).to(DEV)
    ((lr * dl).sum() + (zr * dz).sum()).backward()
    pg = [p.clone().requires_grad_(True) for p in ps]
    hg = h.clone().requires_grad_(True)
    lg, zg = ops.dheads(hg, *pg, 0.2)
    assert lg.dtype == torch.float32 and tuple(lg.shape) == (batch, 1) and tuple(zg.shape) == (batch, zdim)
    assert rel_err(lg, lr) < 1e-5 and rel_err(zg, zr) < 1e-5
    ((lg * dl).sum() + (zg * dz).sum()).backward()
    for a, b in zip(pg, pr):
        assert rel_err(a.grad, b.grad) < 1e-5
    assert rel_err(hg.grad.float(), hr.grad) < 2 ** -7                  # dh is stored as bf16
    # generator step: only dh (parameter gradients not requested), logits gradient only
    hg2 = h.clone().requires_grad_(True)
    lg2, zg2 = ops.dheads(hg2, *ps, 0.2)
    (lg2 * dl).sum().backward()
    hr2 = h.float().requires_grad_(True)
    (F.linear(hr2.permute(0, 3, 1, 2).reshape(batch, -1), ps[0], ps[1]) * dl).sum().backward()
    assert rel_err(hg2.grad.float(), hr2.grad) < 2 ** -7


def _bf16_d_errors(net, x, z):
    xg = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=BF):
        logits, zp = net(xg)
    loss, _ = ops.hologan_g_loss(logits, zp, z)
    loss.backward()
    return logits, zp, xg.grad, dict(net.named_parameters())


@pytest.mark.parametrize("batch,img", [(8, 64), (64, 64), (8, 128)])
def test_discriminator_b200_path_vs_oracle(batch, img):
    """The whole discriminator on the hand-written kernels (Discriminator._forward_b200, no cuDNN / cuBLAS kernel) under
    bf16 autocast against the fp32 oracle: outputs within north_star's 2e-2; gradients (parameters and the input image)
    against the bf16 STORAGE floor -- the same fp32 CPU arithmetic with the stored activations / gradients / conv operands
    rounded to bf16 (oracle.discriminator_forward(bf16_storage=True)): rms-relative error <= 1.25x the floor's, and
    max-normalised <= max(2e-2, 2x the floor's) per tensor.  u / v buffers within 1e-5."""
    gen = torch.Generator().manual_seed(10 + batch + img)
    p = orc.init_discriminator_params(3, 64, 128, img, generator=gen)
    for k in list(p):                                       # trained-like biases (the reference initialises them to 0)
        if k.endswith("bias") and not k.startswith("blocks."):
            p[k] = torch.randn(p[k].shape, generator=gen) * 0.05
    x = torch.rand(batch, 3, img, img, generator=gen) * 2 - 1
    z = torch.rand(batch, 128, generator=gen) * 2 - 1

    def run_oracle(bf16_storage):
        pr = {k: (v.clone().requires_grad_(True) if not k.endswith(("_u", "_v")) else v.clone()) for k, v in p.items()}
        xr = x.clone().requires_grad_(True)
        lo, zo = orc.discriminator_forward(pr, xr, training=True, bf16_storage=bf16_storage)
        (F.binary_cross_entropy_with_logits(lo, torch.ones_like(lo)) + ((zo - z) ** 2).mean()).backward()
        return lo, zo, xr.grad, pr

    lo, zo, dxo, pro = run_oracle(False)
    lf, zf, dxf, prf = run_oracle(True)

    net = Discriminator(3, 64, 128, img_size=img).to(DEV)
    sd = {k: v for k, v in p.items()}
    for i in range(3):                                      # reference state_dict carries the alias keys too
        for s in ("bias", "weight_orig", "weight_u", "weight_v"):
            sd[f"blocks.{i}.conv2d_spec_norm.{s}"] = p[f"blocks.{i}.conv2d.{s}"]
    net.load_state_dict(sd)
    net.train()
    xd = x.to(DEV)
    with torch.autocast("cuda", dtype=BF):
        assert net._b200_ok(xd)
    lg, zg, dxg, named = _bf16_d_errors(net, xd, z.to(DEV))
    assert rel_err(lg.float(), lo) < 2e-2 and rel_err(zg.float(), zo) < 2e-2
    for i in range(3):
        k = f"blocks.{i}.conv2d."
        assert rel_err(net.blocks[i].conv2d.weight_u, pro[k + "weight_u"]) < 1e-5
        assert rel_err(net.blocks[i].conv2d.weight_v, pro[k + "weight_v"]) < 1e-5
    ours, floor = {"dx": (rel_err(dxg, dxo), rms_err(dxg, dxo))}, {"dx": (rel_err(dxf, dxo), rms_err(dxf, dxo))}
    for k, v in pro.items():
        if k.endswith(("_u", "_v")):
            continue
        if k.startswith("blocks.") and k.endswith("bias"):
            assert named[k].grad is None                    # not applied: InstanceNorm cancels it
            continue
        ours[k] = (rel_err(named[k].grad, v.grad), rms_err(named[k].grad, v.grad))
        floor[k] = (rel_err(prf[k].grad, v.grad), rms_err(prf[k].grad, v.grad))
    report = {k: tuple(round(e, 4) for e in ours[k] + floor[k]) for k in ours}
    bad = {k: report[k] for k in ours if ours[k][0] > max(2e-2, 2.0 * floor[k][0]) or ours[k][1] > max(1e-2, 1.25 * floor[k][1])}
    assert not bad, (bad, report)


def test_discriminator_b200_path_launches_no_library_kernel():
    """Every kernel of a forward + backward through the b200 path is ours (kernel names from the torch profiler)."""
    from torch.profiler import ProfilerActivity, profile
    torch.manual_seed(0)
    net = Discriminator(3, 64, 128).to(DEV)
    x = (torch.rand(8, 3, 64, 64, device=DEV) * 2 - 1).requires_grad_(True)
    z = torch.rand(8, 128, device=DEV) * 2 - 1

    def step():
        with torch.autocast("cuda", dtype=BF):
            logits, zp = net(x)
        loss, _ = ops.hologan_g_loss(logits, zp, z)
        loss.backward()

    step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    lib = [n for n in names if "hg::" not in n and
           any(t in n.lower() for t in ("cudnn", "cutlass", "nvjet", "gemm", "gemv", "convolve", "cublas", "implicit", "xmma"))]
    assert not lib, lib
    assert any("tap_gemm_kernel" in n for n in names) and any("c0_img2map_kernel" in n for n in names) \
        and any("dheads_fwd_partial_kernel" in n for n in names)
