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


@pytest.mark.parametrize("in_planes,bf16,tol", [(8, False, 1e-5), (64, True, 2e-2)])
def test_render_views_vs_oracle(in_planes, bf16, tol):
    """`Generator.render_views` (trunk once per latent, rotate + decoder per view) against the ORACLE's per-view forward
    (the reference's pattern: core/figures/types.py:300-322 calls generator(z, view_in=view) once per view)."""
    gen = torch.Generator().manual_seed(31)
    p = orc.init_generator_params(in_planes, 3, 128, 64, generator=gen, bias_std=0.05)
    z = torch.rand(3, 128, generator=gen) * 2 - 1
    views = np.zeros((5, 6))
    views[:, 0] = np.deg2rad(np.linspace(220, 320, 5))
    views[:, 1] = np.deg2rad(90.0)
    views[:, 2] = 1.0
    net = Generator(in_planes, 3, 128, SimpleNamespace(), 64).to(DEV).eval()
    net.load_state_dict(p)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16), torch.no_grad():
        sweep = net.render_views(z.to(DEV), views)
    for v in range(5):
        ref = orc.generator_forward(p, z, np.repeat(views[v:v + 1], 3, axis=0))
        assert rel_err(sweep[:, v].float(), ref) < tol, v


def _bf16_errors(net, p, z, view, dout):
    zg = z.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(zg, view_in=view)
    (out.float() * dout.to(DEV)).sum().backward()
    return out, zg.grad, dict(net.named_parameters())


def _rms_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


@pytest.mark.parametrize("bsz,seed", [(4, 77), (64, 64)])
def test_generator_bf16_tensor_core_path_vs_oracle(bsz, seed):
    """Full-width generator (in_planes 64) under bf16 autocast: tcgen05 convs + AdaIN + rotate, forward and backward,
    against the fp32 oracle -- at a small batch and at the bench configuration's batch 64 (BASELINE configs[1]: the
    split-K plans, tile counts and launch shapes bench.py runs).

    Tolerance (max|a-b| / max|b|): north_star's 2e-2 for the output image and for every gradient tensor
    whose bf16 STORAGE FLOOR allows it.  The floor is measured, not assumed: `orc.generator_forward(bf16_storage=True)`
    is the fp32 CPU oracle with nothing changed except that every activation / gradient a bf16 pipeline keeps in memory
    is rounded to bf16 where it is stored and the conv weights are bf16 copies -- no kernel of ours involved.  That alone
    puts the gradients of everything below block4 at 4-20 % (max-normalised; 4-10 % rms: profiles/r02b_bf16_error_report.txt):
    rounding flips the ReLU mask of ~0.05 % of the activations per layer and every AdaIN backward projects the mean and
    x_hat components (most of the energy) out of the incoming gradient, so the relative error of what is left grows
    layer by layer.  2e-2 is therefore unreachable for those tensors with ANY bf16-operand implementation, and the bar for
    them is the floor itself: rms-relative error <= 1.25 x the floor's, max-normalised <= 2 x the floor's (a single
    sample's max scatters by ~1.5x).  Measured: ours sits at 0.98-1.05 x the floor's rms on every tensor (and below the
    stock cuDNN bf16 path)."""
    gen = torch.Generator().manual_seed(seed)
    p = orc.init_generator_params(64, 3, 128, 64, generator=gen, bias_std=0.05)
    z = torch.rand(bsz, 128, generator=gen) * 2 - 1
    view = orc.sample_view(bsz, np.random.RandomState(seed))
    dout = torch.randn(bsz, 3, 64, 64, generator=gen)

    def oracle(bf16_storage):
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        zr = z.clone().requires_grad_(True)
        out = orc.generator_forward(pr, zr, view, bf16_storage=bf16_storage)
        (out * dout).sum().backward()
        return out.detach(), zr.grad, {k: v.grad for k, v in pr.items()}

    ref, dz_ref, g_ref = oracle(False)
    flo, dz_flo, g_flo = oracle(True)

    net = Generator(64, 3, 128, SimpleNamespace(), 64).to(DEV)
    net.load_state_dict(p)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert net._use_tensor_core_path(z.to(DEV))
    out, dz, named = _bf16_errors(net, p, z, view, dout)

    assert rel_err(out.float(), ref) < 2e-2
    ours = {"dz": (rel_err(dz, dz_ref), _rms_err(dz, dz_ref))}
    floor = {"dz": (rel_err(dz_flo, dz_ref), _rms_err(dz_flo, dz_ref))}
    for k, g in g_ref.items():
        if k.endswith("convTranspose.bias"):
            assert named[k].grad is None or named[k].grad.abs().max() == 0      # analytically zero
            continue
        ours[k] = (rel_err(named[k].grad, g), _rms_err(named[k].grad, g))
        floor[k] = (rel_err(g_flo[k], g), _rms_err(g_flo[k], g))
    report = {k: tuple(round(e, 4) for e in ours[k] + floor[k]) for k in ours}
    # (1) north_star's 2e-2 wherever bf16 storage permits it
    reachable = [k for k in ours if floor[k][0] <= 1e-2]
    assert "final_layer.weight" in reachable and "final_layer.bias" in reachable, report
    for k in reachable:
        assert ours[k][0] < 2e-2, (k, report[k])
    # (2) everywhere else: no worse than the storage format itself
    bad = {k: report[k] for k in ours if ours[k][0] > max(2e-2, 2.0 * floor[k][0]) or ours[k][1] > max(1e-2, 1.25 * floor[k][1])}
    assert not bad, (bad, report)
    mean_ours = sum(v[1] for v in ours.values()) / len(ours)
    mean_floor = sum(v[1] for v in floor.values()) / len(floor)
    assert mean_ours <= 1.1 * mean_floor, (mean_ours, mean_floor, report)
