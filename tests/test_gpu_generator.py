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
