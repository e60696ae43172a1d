"""ZMapping linear+ReLU and final conv+tanh kernels against torch fp32."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from lightning_gan_zoo_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("b,k,n", [(64, 128, 1024), (8, 128, 128), (3, 128, 512), (1, 64, 6)])
def test_linear_relu(b, k, n):
    g = torch.Generator().manual_seed(n)
    z = torch.rand(b, k, generator=g) * 2 - 1
    w = torch.randn(n, k, generator=g) * 0.05
    bias = torch.randn(n, generator=g) * 0.1
    dout = torch.randn(b, n, generator=g)
    zr, wr, br = (t.clone().requires_grad_(True) for t in (z, w, bias))
    ref = F.relu(F.linear(zr, wr, br))
    (ref * dout).sum().backward()
    zg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (z, w, bias))
    out = ops.linear_relu(zg, wg, bg)
    (out * dout.to(DEV)).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(zg.grad, zr.grad) < 1e-5 and rel_err(wg.grad, wr.grad) < 1e-5 and rel_err(bg.grad, br.grad) < 1e-5


@pytest.mark.parametrize("b,c,s,cout", [(4, 64, 64, 3), (2, 16, 32, 3), (3, 8, 16, 1), (2, 128, 8, 4)])
def test_final_conv_tanh(b, c, s, cout):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(c + s)
    x = torch.randn(b, c, s, s, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, c, 3, 3, generator=g) * 0.05
    bias = torch.randn(cout, generator=g) * 0.1
    dout = torch.randn(b, cout, s, s, generator=g)
    xr = x.float().clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = bias.clone().requires_grad_(True)
    ref = torch.tanh(F.conv2d(xr, wr, br, padding=1))
    (ref * dout).sum().backward()
    x_cl = x.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    wg = w.to(DEV).requires_grad_(True); bg = bias.to(DEV).requires_grad_(True)
    out = ops.final_conv_tanh(x_cl, wg, bg)
    (out * dout.to(DEV)).sum().backward()
    assert out.dtype == torch.float32 and tuple(out.shape) == (b, cout, s, s)
    assert rel_err(out, ref) < 1e-5
    assert rel_err(x_cl.grad.permute(0, 3, 1, 2).float(), xr.grad) < 2 ** -7      # dx stored in bf16
    assert rel_err(wg.grad, wr.grad) < 2e-5 and rel_err(bg.grad, br.grad) < 2e-5
    out2 = ops.final_conv_tanh(x_cl.detach(), wg, bg)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows,cols,slope", [(16384, 1024, 0.0), (300, 64, 0.2), (77, 24, 0.0), (5, 4096, 0.0)])
def test_act_bwd_bias(rows, cols, slope):
    g = torch.Generator().manual_seed(rows + cols)
    y = torch.randn(rows, cols, generator=g).to(torch.bfloat16).to(DEV)
    dy = torch.randn(rows, cols, generator=g).to(torch.bfloat16).to(DEV)
    dpre, db = ops.act_bwd_bias(y, dy, slope, True)
    ref = torch.where(y > 0, dy, (dy.float() * slope).to(torch.bfloat16))
    assert torch.equal(dpre, ref)
    assert rel_err(db, ref.float().sum(0)) < 1e-5
    dpre2, db2 = ops.act_bwd_bias(y, dy, slope, True)
    assert torch.equal(db, db2)                       # deterministic
    dpre3, none = ops.act_bwd_bias(y, dy, slope, False)
    assert none is None and torch.equal(dpre3, ref)
