"""AdaIN(+activation) kernels vs the reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from lightning_gan_zoo_b200 import ops
from oracle import hologan_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fwd_bwd_vs_reference_golden():
    g = load_golden("adain.npz")
    for i in range(int(g["n_cases"])):
        x = torch.from_numpy(g[f"x{i}"]).to(DEV).requires_grad_(True)
        s = torch.from_numpy(g[f"s{i}"]).to(DEV).requires_grad_(True)
        b = torch.from_numpy(g[f"b{i}"]).to(DEV).requires_grad_(True)
        y_plain = ops.adain_act(x, s, b, neg_slope=1.0)
        assert rel_err(y_plain, g[f"y{i}"]) < 1e-5, i
        y = ops.adain_act(x, s, b, neg_slope=0.0)
        assert rel_err(y, np.maximum(g[f"y{i}"], 0)) < 1e-5, i
        (y * torch.from_numpy(g[f"dy{i}"]).to(DEV)).sum().backward()
        # entries whose pre-activation is within rounding of 0 may flip the ReLU mask: compare with the
        # max-normalised metric over entries where |y_ref| is not tiny
        assert rel_err(x.grad, g[f"dx{i}"]) < 2e-5, i
        assert rel_err(s.grad, g[f"ds{i}"]) < 2e-5, i
        assert rel_err(b.grad, g[f"db{i}"]) < 2e-5, i


@pytest.mark.parametrize("shape", [(64, 512, 64), (64, 128, 512), (16, 64, 4096), (16, 256, 1024), (3, 5, 192)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("slope", [0.0, 0.2])
def test_hot_path_shapes_vs_oracle(shape, dtype, slope):
    """The five call-site shapes of SURVEY.md 8-a3 (C x N), fwd + bwd, against the torch-CPU oracle."""
    bsz, c, n = shape
    gen = torch.Generator().manual_seed(n + c)
    x = (torch.randn(bsz, c, n, generator=gen) * 2 + 0.5).to(dtype)
    s = torch.rand(bsz, c, generator=gen) + 0.1
    b = torch.randn(bsz, c, generator=gen)
    dy = torch.randn(bsz, c, n, generator=gen).to(dtype)
    xr = x.float().clone().requires_grad_(True); sr = s.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.leaky_relu(orc.adain(xr, sr, br), slope)
    (yr * dy.float()).sum().backward()
    xg = x.to(DEV).requires_grad_(True); sg = s.to(DEV).requires_grad_(True); bg = b.to(DEV).requires_grad_(True)
    y = ops.adain_act(xg, sg, bg, neg_slope=slope)
    (y.float() * dy.to(DEV).float()).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype
    assert rel_err(y.float(), yr) < tol
    assert rel_err(xg.grad.float(), xr.grad) < (3e-5 if dtype == torch.float32 else 2e-2)
    assert rel_err(sg.grad, sr.grad) < (3e-5 if dtype == torch.float32 else 2e-2)
    assert rel_err(bg.grad, br.grad) < (3e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("n", [16, 5, 67])
def test_constant_input_broadcast(n):
    """x of batch 1 (the learned 4^3 constant) == the reference's x.repeat(B, ...) (:121), with the
    gradient summed over the batch (8 warps per channel split the samples: fewer samples than warps, and a
    batch that is not a multiple of 8, are both covered)."""
    gen = torch.Generator().manual_seed(9)
    x = (torch.randn(1, 512, 4, 4, 4, generator=gen) - 0.5) / 0.5
    s = torch.rand(n, 512, generator=gen); b = torch.randn(n, 512, generator=gen)
    dy = torch.randn(n, 512, 4, 4, 4, generator=gen)
    xr = x.clone().requires_grad_(True); sr = s.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = torch.relu(orc.adain(xr.repeat(n, 1, 1, 1, 1), sr, br))
    (yr * dy).sum().backward()
    xg = x.to(DEV).requires_grad_(True); sg = s.to(DEV).requires_grad_(True); bg = b.to(DEV).requires_grad_(True)
    y = ops.adain_act(xg, sg, bg, 0.0)
    assert tuple(y.shape) == (n, 512, 4, 4, 4)
    (y * dy.to(DEV)).sum().backward()
    assert rel_err(y, yr) < 1e-5
    assert tuple(xg.grad.shape) == (1, 512, 4, 4, 4)
    assert rel_err(xg.grad, xr.grad) < 3e-5
    assert rel_err(sg.grad, sr.grad) < 3e-5 and rel_err(bg.grad, br.grad) < 3e-5


def test_split_style_views():
    """scale / bias given as the two halves of one ZMapping output (row stride 2C), no copies."""
    style = torch.rand(4, 64, device=DEV)
    x = torch.randn(4, 32, 8, 8, device=DEV)
    y = ops.adain_act(x, style[:, :32], style[:, 32:], 0.0)
    yr = torch.relu(orc.adain(x.cpu(), style[:, :32].cpu(), style[:, 32:].cpu()))
    assert rel_err(y, yr) < 1e-5


def test_statistics_property():
    """Size-independent property at the largest hot-path instance: plain AdaIN output has per-(b,c)
    mean == bias and unbiased std == scale."""
    x = torch.randn(64, 64, 4096, device=DEV) * 3 + 1
    s = torch.rand(64, 64, device=DEV) + 0.5
    b = torch.randn(64, 64, device=DEV)
    y = ops.adain_act(x, s, b, neg_slope=1.0)
    assert (y.mean(2) - b).abs().max() < 1e-4
    assert (y.std(2) / s - 1).abs().max() < 1e-4


def test_unsupported_shapes_fail_loudly():
    from lightning_gan_zoo_b200._lib import HologanB200Error
    with pytest.raises(HologanB200Error):
        ops.adain_act(torch.randn(2, 3, 6, device=DEV), torch.ones(2, 3, device=DEV), torch.zeros(2, 3, device=DEV))
    with pytest.raises(HologanB200Error, match="single-pass limit"):
        ops.adain_act(torch.randn(1, 1, 32768, device=DEV), torch.ones(1, 1, device=DEV), torch.zeros(1, 1, device=DEV))
