"""The drop-in Generator (B200 ops) against the reference's golden outputs, forward and backward."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import load_golden, params_sha, rel_err, sha16
from lightning_gan_zoo_b200.core.models.hologan_generator import Generator
from oracle import hologan_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("tag", ["p8", "p16"])
def test_generator_fp32_vs_reference_golden(tag):
    g = load_golden(f"generator_{tag}.npz")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    p = orc.init_generator_params(int(g["in_planes"]), 3, 128, 64, generator=gen, bias_std=0.05)
    if params_sha(p) != str(g["params_sha"]):
        pytest.skip("torch CPU RNG stream differs from the fixture's")
    bsz = g["z"].shape[0]
    _ = torch.rand(bsz, 128, generator=gen)
    dout = torch.randn(bsz, 3, 64, 64, generator=gen)
    assert sha16(dout) == str(g["dout_sha"])
    net = Generator(int(g["in_planes"]), 3, 128, SimpleNamespace(), 64).to(DEV)
    net.load_state_dict(p)
    z = torch.from_numpy(g["z"]).to(DEV).requires_grad_(True)
    out = net(z, view_in=g["view"])
    assert rel_err(out, g["out"]) < 1e-5
    (out * dout.to(DEV)).sum().backward()
    assert rel_err(z.grad, g["dz"]) < 1e-4
    named = dict(net.named_parameters())
    for k, (s, a) in zip([str(k) for k in g["grad_keys"]], g["grad_summary"]):
        if k.endswith("convTranspose.bias"):
            continue      # analytically zero (bias in front of an instance norm)
        got = named[k].grad.double().abs().sum().item()
        assert abs(got - a) <= 1e-4 * a, (k, got, a)
    for k in g.files:
        if k.startswith("grad::") and not k.endswith("convTranspose.bias"):
            assert rel_err(named[k[6:]].grad, g[k]) < 1e-4, k


def test_generator_accepts_tensor_views_and_samples_views():
    va = SimpleNamespace(azimuth_low=220, azimuth_high=320, elevation_low=70, elevation_high=110, scale_low=1,
                         scale_high=1, transX_low=0, transX_high=0, transY_low=0, transY_high=0, transZ_low=0,
                         transZ_high=0, batch_size=4)
    net = Generator(8, 3, 128, va, 64).to(DEV)
    z = torch.rand(4, 128, device=DEV) * 2 - 1
    view = orc.sample_view(4, np.random.RandomState(1))
    a = net(z, view_in=view)
    b = net(z, view_in=torch.from_numpy(view).float().to(DEV))
    assert rel_err(a, b) < 1e-6      # cuDNN's conv kernels are not bitwise run-to-run deterministic
    c = net(z)
    assert tuple(c.shape) == (4, 3, 64, 64) and c.abs().max() <= 1
    net128 = Generator(8, 3, 128, va, 128).to(DEV)
    assert tuple(net128(z).shape) == (4, 3, 128, 128)


@pytest.mark.parametrize("in_planes,bf16", [(8, False), (64, True)])
def test_render_views_equals_per_view_forward(in_planes, bf16):
    """View sweep (BASELINE cfg 5): the 3D trunk runs once per z, rotate + decoder per view; every column
    must equal the reference-style call `generator(z, view_in=view)` (core/figures/types.py:233-237)."""
    torch.manual_seed(5)
    net = Generator(in_planes, 3, 128, SimpleNamespace(), 64).to(DEV).eval()
    z = torch.rand(3, 128, device=DEV) * 2 - 1
    views = np.zeros((6, 6))
    views[:, 0] = np.deg2rad(np.linspace(220, 320, 6))
    views[:, 1] = np.deg2rad(90.0)
    views[:, 2] = 1.0
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16), torch.no_grad():
        assert net._use_tensor_core_path(z) == bf16
        sweep = net.render_views(z, views)
        assert tuple(sweep.shape) == (3, 6, 3, 64, 64)
        for v in range(6):
            one = net(z, view_in=np.repeat(views[v:v + 1], 3, axis=0))
            if bf16:
                assert torch.equal(sweep[:, v], one)       # our kernels only: deterministic
            else:
                assert rel_err(sweep[:, v], one) < 1e-6    # cuDNN fp32 convs are not bitwise run-to-run deterministic
        # per-latent views (B, V, 6)
        pv = np.stack([np.roll(views, i, axis=0) for i in range(3)])
        sweep2 = net.render_views(z, pv)
        one = net(z, view_in=pv[:, 2])
        assert rel_err(sweep2[:, 2], one) < 1e-6
    with pytest.raises(ValueError):
        net.render_views(z, np.zeros((4, 5)))


def _bf16_errors(net, p, z, view, dout):
    zg = z.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(zg, view_in=view)
    (out.float() * dout.to(DEV)).sum().backward()
    return out, zg.grad, dict(net.named_parameters())


def test_generator_bf16_tensor_core_path_vs_oracle():
    """Full-width generator (in_planes 64) under bf16 autocast: tcgen05 convs + AdaIN + rotate, forward
    and backward, against the fp32 oracle.

    Tolerances (max|a-b| / max|b|):
      * activations (output image): north_star's 2e-2;
      * gradients: bf16 rounding of the stored activations / gradients is amplified by every AdaIN backward
        (it projects out the mean and x_hat components of the incoming gradient, which carry most of its
        energy), so even stock PyTorch bf16 autocast (cuDNN convs) sits at 5-14 % for the deep layers on this
        network (tools/bf16_error_report.py).  The bar here: 2e-2 where the stock bf16 path reaches it,
        otherwise per tensor no worse than 2x, and on average over all tensors no worse than 1.15x, the
        stock bf16 path's error on the same inputs (single-sample errors scatter by ~1.5x).
    """
    gen = torch.Generator().manual_seed(77)
    p = orc.init_generator_params(64, 3, 128, 64, generator=gen, bias_std=0.05)
    bsz = 4
    z = torch.rand(bsz, 128, generator=gen) * 2 - 1
    view = orc.sample_view(bsz, np.random.RandomState(77))
    dout = torch.randn(bsz, 3, 64, 64, generator=gen)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    zr = z.clone().requires_grad_(True)
    ref = orc.generator_forward(pr, zr, view)
    (ref * dout).sum().backward()

    net = Generator(64, 3, 128, SimpleNamespace(), 64).to(DEV)
    net.load_state_dict(p)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert net._use_tensor_core_path(z.to(DEV))
    out, dz, named = _bf16_errors(net, p, z, view, dout)
    stock = Generator(64, 3, 128, SimpleNamespace(), 64).to(DEV)
    stock.load_state_dict(p)
    stock._use_tensor_core_path = lambda _z: False          # same module on torch/cuDNN bf16 convs
    out_s, dz_s, named_s = _bf16_errors(stock, p, z, view, dout)

    assert rel_err(out.float(), ref) < 2e-2
    ours, theirs = {"dz": rel_err(dz, zr.grad)}, {"dz": rel_err(dz_s, zr.grad)}
    for k, v in pr.items():
        if k.endswith("convTranspose.bias"):
            assert named[k].grad is None or named[k].grad.abs().max() == 0      # analytically zero
            continue
        ours[k], theirs[k] = rel_err(named[k].grad, v.grad), rel_err(named_s[k].grad, v.grad)
    bad = {k: (ours[k], theirs[k]) for k in ours if ours[k] > max(2e-2, 2.0 * theirs[k])}
    assert not bad, bad
    mean_ours, mean_theirs = sum(ours.values()) / len(ours), sum(theirs.values()) / len(theirs)
    assert mean_ours <= 1.15 * mean_theirs, (mean_ours, mean_theirs)
    assert rel_err(named["final_layer.weight"].grad, pr["final_layer.weight"].grad) < 2e-2
    assert rel_err(named["block4.zMapping.linear1.weight"].grad, pr["block4.zMapping.linear1.weight"].grad) < 3e-2
