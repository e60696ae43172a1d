"""tcgen05 implicit-GEMM kernels (convT fwd / dgrad / wgrad, plain GEMM) against torch's fp32 ops on the
same bf16-rounded operands.  The reference arithmetic of these ops IS torch's, so
F.conv_transpose{2,3}d on fp32 copies of the operands is the oracle; tolerance = bf16 output rounding."""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from lightning_gan_zoo_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
P = ops._ptr


def bf(t):
    return t.to(torch.bfloat16)


def to_channels_last(x):          # (B,C,*sp) -> (B,*sp,C) contiguous
    perm = [0] + list(range(2, x.dim())) + [1]
    return x.permute(*perm).contiguous()


def s2d_from_nc(y, ndim):         # (B,C,2S,..) -> (B,S,..,P,C)
    b, c = y.shape[:2]
    s = y.shape[2] // 2
    if ndim == 2:
        t = y.reshape(b, c, s, 2, s, 2).permute(0, 2, 4, 3, 5, 1)          # b,i,j,py,px,c
        return t.reshape(b, s, s, 4, c).contiguous()
    t = y.reshape(b, c, s, 2, s, 2, s, 2).permute(0, 2, 4, 6, 3, 5, 7, 1)  # b,iz,iy,ix,pz,py,px,c
    return t.reshape(b, s, s, s, 8, c).contiguous()


def nc_from_s2d(t, ndim):         # inverse of the above
    b = t.shape[0]
    s = t.shape[1]
    c = t.shape[-1]
    if ndim == 2:
        return t.reshape(b, s, s, 2, 2, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, 2 * s, 2 * s)
    return t.reshape(b, s, s, s, 2, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, c, 2 * s, 2 * s, 2 * s)


def pack(w, taps):
    cin, cout = w.shape[:2]
    wf = torch.empty(taps, cout, cin, dtype=torch.bfloat16, device=DEV)
    wd = torch.empty(taps, cin, cout, dtype=torch.bfloat16, device=DEV)
    _lib.call("hg_convt_pack_weight", P(w.contiguous()), P(wf), P(wd), cin, cout, taps, 0, 0, ops._stream())
    return wf, wd


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (256, 256, 128), (16384, 1024, 1024), (64, 2048, 128), (200, 48, 192)])
def test_plain_gemm(m, n, k):
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = bf(torch.randn(m, k, generator=g)).to(DEV)
    b = bf(torch.randn(n, k, generator=g) * 0.05).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    d = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
    _lib.call("hg_gemm_bf16_nt", P(a), P(b), P(bias), P(d), m, n, k, n, ctypes.c_float(0.0), ops._stream())
    ref = torch.relu(a.float() @ b.float().t() + bias)
    assert rel_err(d.float(), ref) < 2 ** -7


def test_pack_weight_layouts():
    w = torch.randn(64, 32, 4, 4, device=DEV)
    wf, wd = pack(w, 16)
    assert torch.equal(wf, bf(w.reshape(64, 32, 16).permute(2, 1, 0)))
    assert torch.equal(wd, bf(w.reshape(64, 32, 16).permute(2, 0, 1)))


CASES = [  # (ndim, kernel, batch, cin, cout, size)
    (2, 1, 4, 128, 64, 16),
    (2, 1, 8, 1024, 1024, 16),     # the 1x1 projection (a9)
    (2, 4, 2, 128, 64, 16),
    (2, 4, 4, 1024, 256, 16),      # block3 (a10)
    (2, 4, 4, 256, 64, 32),        # block4 (a10)
    (3, 3, 4, 512, 128, 4),        # block1 (a4)
    (3, 3, 3, 512, 128, 4),        # odd batch: last 128-row box is half out of range
    (3, 3, 2, 128, 64, 8),         # block2 (a4)
    # the bench configuration (BASELINE.json configs[1], batch 64): the split-K plans, tile counts and dual / single
    # launch shapes that bench.py actually runs
    (2, 4, 64, 1024, 256, 16),     # block3
    (2, 4, 64, 256, 64, 32),       # block4
    (3, 3, 64, 512, 128, 4),       # block1
    (3, 3, 64, 128, 64, 8),        # block2
    (2, 1, 64, 1024, 1024, 16),    # projection
]


def torch_convt(x, w, bias, ndim, kernel):
    if kernel == 1:
        return F.conv_transpose2d(x, w, bias)
    if ndim == 2 and kernel == 5:
        return F.conv_transpose2d(x, w, bias, stride=2, padding=2, output_padding=1)
    if ndim == 2:
        return F.conv_transpose2d(x, w, bias, stride=2, padding=1)
    return F.conv_transpose3d(x, w, bias, stride=2, padding=1, output_padding=1)


# k5 (s2, p2, op1): the transposed convolution whose dgrad / forward / wgrad are the discriminator's Conv2d(k5, s2, p2)
# forward / dgrad / wgrad.  Shapes = the duals of D's three spectral-norm blocks.
K5_CASES = [(2, 5, 4, 128, 64, 16), (2, 5, 4, 256, 128, 8), (2, 5, 8, 512, 256, 4), (2, 5, 3, 128, 64, 16)]


@pytest.mark.parametrize("ndim,kernel,batch,cin,cout,size", CASES + K5_CASES)
def test_convt_fwd_dgrad_wgrad(ndim, kernel, batch, cin, cout, size):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout)
    sp = (size,) * ndim
    taps = kernel ** ndim
    x = bf(torch.randn(batch, cin, *sp, generator=g)).to(DEV)
    w = bf(torch.randn(cin, cout, *((kernel,) * ndim), generator=g) * 0.05).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    wf, wd = pack(w.float(), taps)
    x_cl = to_channels_last(x)
    nclass = 1 if kernel == 1 else 2 ** ndim
    # ---- forward
    y_s2d = torch.empty((batch,) + sp + (nclass, cout), dtype=torch.bfloat16, device=DEV)
    _lib.call("hg_convt_fwd", P(x_cl), P(wf), P(bias), P(y_s2d), batch, cin, cout, ndim, size, kernel, ctypes.c_float(1.0),
              ops._stream())
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    yr = torch_convt(xr, wr, bias, ndim, kernel)
    got = nc_from_s2d(y_s2d, ndim) if kernel != 1 else y_s2d.reshape((batch,) + sp + (cout,)).permute(0, 3, 1, 2)
    assert rel_err(got.float(), yr) < 2 ** -7
    # ---- backward with a bf16 output gradient
    dy = bf(torch.randn(yr.shape, generator=g)).to(DEV)
    yr.backward(dy.float())
    dy_s2d = s2d_from_nc(dy, ndim) if kernel != 1 else to_channels_last(dy).reshape((batch,) + sp + (1, cout))
    dx = torch.empty_like(x_cl)
    _lib.call("hg_convt_dgrad", P(dy_s2d), P(wd), P(dx), batch, cin, cout, ndim, size, kernel, ops._stream())
    dx_nc = dx.permute(*([0, ndim + 1] + list(range(1, ndim + 1))))
    assert rel_err(dx_nc.float(), xr.grad) < 2 ** -7
    if cin % 128 == 0 and cout % 64 == 0:
        dw = ops.convt_wgrad(x_cl, dy_s2d, wr.shape, ndim, kernel)
        dw2 = ops.convt_wgrad(x_cl, dy_s2d, wr.shape, ndim, kernel)
        assert torch.equal(dw, dw2)               # split-K partials are reduced in a fixed order
        assert rel_err(dw, wr.grad) < 1e-4        # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("ndim,kernel,batch,cin,cout,size", [
    (2, 4, 64, 256, 64, 32), (3, 3, 64, 128, 64, 8), (3, 3, 64, 512, 128, 4), (3, 3, 3, 512, 128, 4), (2, 4, 4, 128, 32, 16),
    (3, 3, 2, 64, 32, 8), (2, 5, 4, 128, 64, 16), (2, 4, 2, 64, 96, 16)])
@pytest.mark.parametrize("boxes", [0, 2, 4])
def test_convt_fwd_shift_major_items(hg_option, boxes, ndim, kernel, batch, cin, cout, size):
    """Option TAPGEMM_SHARE_A: the forward of the narrow layers with every shifted activation box loaded once for all
    parity classes of a CTA (stacked weight boxes, one MMA per run of adjacent classes) against torch, for 0 (per-tap
    boxes), 2 and 4 weight boxes per item."""
    hg_option("TAPGEMM_SHARE_A", boxes)
    g = torch.Generator(device="cpu").manual_seed(cin + 3 * cout + boxes)
    sp = (size,) * ndim
    x = bf(torch.randn(batch, cin, *sp, generator=g)).to(DEV)
    w = bf(torch.randn(cin, cout, *((kernel,) * ndim), generator=g) * 0.05).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    wf, _ = pack(w.float(), kernel ** ndim)
    y_s2d = torch.empty((batch,) + sp + (2 ** ndim, cout), dtype=torch.bfloat16, device=DEV)
    _lib.call("hg_convt_fwd", P(to_channels_last(x)), P(wf), P(bias), P(y_s2d), batch, cin, cout, ndim, size, kernel,
              ctypes.c_float(0.2), ops._stream())
    ref = F.leaky_relu(torch_convt(x.float(), w.float(), bias, ndim, kernel), 0.2)
    assert rel_err(nc_from_s2d(y_s2d, ndim).float(), ref) < 2 ** -7


def test_fwd_relu_epilogue_and_errors():
    x = bf(torch.randn(2, 16, 16, 64)).to(DEV)
    w = torch.randn(64, 32, 1, 1, device=DEV) * 0.1
    wf, _ = pack(w, 1)
    y = torch.empty(2, 16, 16, 1, 32, dtype=torch.bfloat16, device=DEV)
    _lib.call("hg_convt_fwd", P(x), P(wf), P(None), P(y), 2, 64, 32, 2, 16, 1, ctypes.c_float(0.0), ops._stream())
    ref = torch.relu(x.float().reshape(-1, 64) @ bf(w).float().reshape(64, 32))
    assert rel_err(y.float().reshape(-1, 32), ref) < 2 ** -7
    with pytest.raises(_lib.HologanB200Error, match="Cin"):
        _lib.call("hg_convt_fwd", P(x), P(wf), P(None), P(y), 2, 48, 32, 2, 16, 1, ctypes.c_float(0.0), ops._stream())
    with pytest.raises(_lib.HologanB200Error, match="supported"):
        _lib.call("hg_convt_fwd", P(x), P(wf), P(None), P(y), 2, 64, 32, 2, 16, 3, ctypes.c_float(0.0), ops._stream())


def test_projection_channel_permutation():
    """pack / wgrad with perm=(C,S): GEMM K index y*C + c <-> torch channel c*S + (S-1-y) (reference :130-133)."""
    c, sz, cout, b = 8, 16, 64, 2
    cin = c * sz
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(cin, cout, 1, 1, generator=g) * 0.05).to(DEV)
    wf, wd = ops.pack_convt_weight(w, (c, sz))
    ref = w.reshape(c, sz, cout).flip(1).permute(1, 0, 2).reshape(cin, cout)          # [y*C + c][co]
    assert torch.equal(wd[0], bf(ref)) and torch.equal(wf[0], bf(ref.t()))
    x = bf(torch.randn(b, 16, 16, cin, generator=g)).to(DEV)
    dy = bf(torch.randn(b, 16, 16, 1, cout, generator=g)).to(DEV)
    dw = ops.convt_wgrad(x, dy, w.shape, 2, 1, (c, sz))
    dref = (x.float().reshape(-1, cin).t() @ dy.float().reshape(-1, cout))            # [k'][co]
    dref = dref.reshape(sz, c, cout).permute(1, 0, 2).flip(1).reshape(cin, cout, 1, 1)
    assert rel_err(dw, dref) < 1e-4
