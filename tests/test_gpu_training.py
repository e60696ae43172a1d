"""Discriminator mirror + training-step harness on the GPU against the reference's golden step."""
import numpy as np
import pytest
import torch

from conftest import load_golden, params_sha, rel_err, sha16
from lightning_gan_zoo_b200 import ops
from lightning_gan_zoo_b200.core.models.hologan_discriminator import Discriminator
from lightning_gan_zoo_b200.training import HologanConfig, HologanTrainer, optimizer_index
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


def _load_d(dp, out_planes):
    d = Discriminator(3, out_planes, 128).to(DEV)
    sd = {k: v.clone() for k, v in dp.items()}
    sd.update({k.replace(".conv2d.", ".conv2d_spec_norm."): v for k, v in sd.items() if k.startswith("blocks.")})
    d.load_state_dict(sd)
    return d


def test_training_step_fp32_vs_reference_golden():
    """HOLOGAN.training_step for both optimizer indices (tiny widths) against the reference's losses,
    gradients and spectral-norm buffers; D's InstanceNorm+LeakyReLU runs on the fused kernel."""
    g = load_golden("train_step_tiny.npz")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    dp = orc.init_discriminator_params(3, 8, 128, 64, generator=gen)
    gp = orc.init_generator_params(8, 3, 128, 64, generator=gen, bias_std=0.05)
    if params_sha(dp) != str(g["d_params_sha"]) or params_sha(gp) != str(g["g_params_sha"]):
        pytest.skip("torch CPU RNG stream differs from the fixture's")
    real = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    assert sha16(real) == str(g["real_sha"])
    cfg = HologanConfig(batch_size=4, gen_in_planes=8, disc_out_planes=8)
    tr = HologanTrainer(cfg, device=DEV, compute_dtype=torch.float32)
    tr.generator.load_state_dict(gp)
    tr.discriminator = _load_d(dp, 8)
    z = torch.from_numpy(g["z"]).to(DEV)
    for p in tr.discriminator.parameters():
        p.requires_grad_(True)
    loss_d = tr.training_step(real.to(DEV), z, g["view"], 0)
    assert abs(loss_d.item() - float(g["loss_d"])) < 2e-6
    loss_d.backward()
    named = dict(tr.discriminator.named_parameters())
    for k, (s, a) in zip([str(k) for k in g["d_grad_keys"]], g["d_grad_summary"]):
        if k.startswith("blocks.") and k.endswith("conv2d.bias"):
            continue
        got = named[k].grad.double().abs().sum().item()
        assert abs(got - a) <= 1e-4 * a, (k, got, a)
    for i in range(3):
        assert np.allclose(tr.discriminator.blocks[i].conv2d.weight_u.cpu().numpy(), g[f"u_after_dstep_{i}"], atol=1e-6)
    for p in tr.discriminator.parameters():
        p.requires_grad_(False)
    loss_g = tr.training_step(None, z, g["view"], 1)
    assert abs(loss_g.item() - float(g["loss_g"])) < 2e-6
    loss_g.backward()
    named = dict(tr.generator.named_parameters())
    for k, (s, a) in zip([str(k) for k in g["g_grad_keys"]], g["g_grad_summary"]):
        if k.endswith("convTranspose.bias"):
            continue
        got = named[k].grad.double().abs().sum().item()
        assert abs(got - a) <= 2e-4 * a, (k, got, a)


def test_schedule_and_lr():
    assert [optimizer_index(i) for i in range(7)] == [0, 1, 1, 0, 1, 1, 0]
    from lightning_gan_zoo_b200.training import hologan_lr_lambda
    f = hologan_lr_lambda(25)
    assert f(0) == 1.0 and f(12) == 1.0 and abs(f(13) - (1 - 0.5 / 12.5)) < 1e-12 and abs(f(25)) < 1e-12


def test_cuda_graph_step_matches_eager():
    """Graph replay of the whole step == eager launches (same kernels, same order), and enabling graphs
    leaves the training state untouched."""
    cfg = HologanConfig(batch_size=8)
    a = HologanTrainer(cfg, device=DEV, seed=3)
    b = HologanTrainer(cfg, device=DEV, seed=3)
    b.enable_cuda_graphs(8)
    for pa, pb in zip(a.generator.parameters(), b.generator.parameters()):
        assert torch.equal(pa, pb)
    gen = torch.Generator().manual_seed(0)
    for i in range(6):
        real = (torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1).to(DEV)
        z = (torch.rand(8, 128, generator=gen) * 2 - 1).to(DEV)
        view = orc.sample_view(8, np.random.RandomState(i))
        la = a.step(real, i, z=z, view=view)
        lb = b.step(real, i, z=z, view=view)
        assert abs(la.item() - lb.item()) < 5e-3 * max(1.0, abs(la.item())), (i, la.item(), lb.item())
    # weight matrices only: biases start at 0 and Adam moves them by +-lr per step whatever the (noisy) gradient
    worst = max(rel_err(pb, pa) for pa, pb in zip(a.generator.parameters(), b.generator.parameters()) if pa.dim() > 1)
    assert worst < 5e-2


@pytest.mark.parametrize("graphs", [False, True])
def test_bucketwise_adam_overlap_is_transparent(graphs, monkeypatch):
    """Gradient buckets are updated by hg_adam_apply on a side stream as soon as they are final (FlatAdam.tick +
    apply_range from the backward hooks); losses and parameters must equal the single hg_adam_step after the backward
    bit for bit."""
    cfg = HologanConfig(batch_size=8)
    monkeypatch.setenv("HG_ADAM_OVERLAP", "1")          # default on for world > 1 only
    a = HologanTrainer(cfg, device=DEV, seed=7)
    monkeypatch.setenv("HG_ADAM_OVERLAP", "0")
    b = HologanTrainer(cfg, device=DEV, seed=7)
    monkeypatch.delenv("HG_ADAM_OVERLAP")
    assert a._adam_overlap and not b._adam_overlap
    if graphs:
        a.enable_cuda_graphs(8)
        b.enable_cuda_graphs(8)
    gen = torch.Generator().manual_seed(2)
    for i in range(6):
        real = (torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1).to(DEV)
        z = (torch.rand(8, 128, generator=gen) * 2 - 1).to(DEV)
        view = orc.sample_view(8, np.random.RandomState(i))
        la, lb = a.step(real, i, z=z, view=view), b.step(real, i, z=z, view=view)
        assert la.item() == lb.item(), (i, la.item(), lb.item())
    torch.cuda.synchronize()
    for pa, pb in zip(list(a.generator.parameters()) + list(a.discriminator.parameters()),
                      list(b.generator.parameters()) + list(b.discriminator.parameters())):
        assert torch.equal(pa, pb)
    assert float(a.opt_g.kstate[0]) == float(b.opt_g.kstate[0]) == 4 and float(a.opt_d.kstate[0]) == 2


@pytest.mark.parametrize("graphs", [False, True])
def test_side_stream_wgrad_is_transparent(graphs, monkeypatch):
    """ops.WGRAD_SIDE_STREAM: the generator's weight-gradient GEMMs run on a side stream beside the backward's dgrad /
    AdaIN chain (only the optimizer reads their result).  Same kernels on the same data: bit-identical training."""
    from lightning_gan_zoo_b200 import ops as hops
    cfg = HologanConfig(batch_size=8)
    gen = torch.Generator().manual_seed(3)
    batches = [((torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1).to(DEV), (torch.rand(8, 128, generator=gen) * 2 - 1).to(DEV),
                orc.sample_view(8, np.random.RandomState(i))) for i in range(6)]

    def run(side):
        monkeypatch.setattr(hops, "WGRAD_SIDE_STREAM", side)
        t = HologanTrainer(cfg, device=DEV, seed=9)
        if graphs:
            t.enable_cuda_graphs(8)
        losses = [t.step(real, i, z=z, view=view).item() for i, (real, z, view) in enumerate(batches)]
        torch.cuda.synchronize()
        return losses, [p.detach().clone() for p in t.generator.parameters()]

    la, pa = run(True)
    lb, pb = run(False)
    assert la == lb
    for u, v in zip(pa, pb):
        assert torch.equal(u, v)


@pytest.mark.parametrize("graphs", [False, True])
def test_d_real_side_stream_is_transparent(graphs):
    """D step: D(real) forward / backward on a side stream beside the generator's forward and D(fake)'s backward.  Losses,
    parameters and the spectral-norm vectors must be those of the serial order, bit for bit."""
    cfg = HologanConfig(batch_size=8)
    a = HologanTrainer(cfg, device=DEV, seed=11)
    b = HologanTrainer(cfg, device=DEV, seed=11)
    assert a._d_real_side
    b._d_real_side = False
    if graphs:
        a.enable_cuda_graphs(8)
        b.enable_cuda_graphs(8)
    gen = torch.Generator().manual_seed(4)
    for i in range(7):
        real = (torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1).to(DEV)
        z = (torch.rand(8, 128, generator=gen) * 2 - 1).to(DEV)
        view = orc.sample_view(8, np.random.RandomState(i))
        la, lb = a.step(real, i, z=z, view=view), b.step(real, i, z=z, view=view)
        assert la.item() == lb.item(), (i, la.item(), lb.item())
    torch.cuda.synchronize()
    for pa, pb in zip(list(a.generator.parameters()) + list(a.discriminator.parameters()),
                      list(b.generator.parameters()) + list(b.discriminator.parameters())):
        assert torch.equal(pa, pb)
    for blk_a, blk_b in zip(a.discriminator.blocks, b.discriminator.blocks):
        assert torch.equal(blk_a.conv2d.weight_u, blk_b.conv2d.weight_u)


@pytest.mark.parametrize("graphs", [False, True])
def test_spectral_norm_prefetch_is_transparent(graphs):
    """The discriminator's power iterations run ahead on a side stream (Discriminator.prefetch_spectral_norm); u / v and
    the losses must be exactly what the in-line iteration gives -- same kernels on the same data, only earlier."""
    cfg = HologanConfig(batch_size=8)
    a = HologanTrainer(cfg, device=DEV, seed=5)
    b = HologanTrainer(cfg, device=DEV, seed=5)
    assert a._sn_prefetch
    b._sn_prefetch = False
    if graphs:
        a.enable_cuda_graphs(8)
        b.enable_cuda_graphs(8)
    gen = torch.Generator().manual_seed(1)
    for i in range(6):
        real = (torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1).to(DEV)
        z = (torch.rand(8, 128, generator=gen) * 2 - 1).to(DEV)
        view = orc.sample_view(8, np.random.RandomState(i))
        la, lb = a.step(real, i, z=z, view=view), b.step(real, i, z=z, view=view)
        assert la.item() == lb.item(), (i, la.item(), lb.item())
    torch.cuda.synchronize()
    for blk_a, blk_b in zip(a.discriminator.blocks, b.discriminator.blocks):
        assert torch.equal(blk_a.conv2d.weight_u, blk_b.conv2d.weight_u)
        assert torch.equal(blk_a.conv2d.weight_orig, blk_b.conv2d.weight_orig)
    assert not a.discriminator._sn_ready


@pytest.mark.parametrize("graphs", [False, True])
def test_step_host_matches_step(graphs):
    """The pipelined host-input entry (pinned staging, copy stream, asynchronous loss read-back) runs the same
    step as `step` on device-resident inputs: same losses, read one step late, and same weights afterwards."""
    cfg = HologanConfig(batch_size=8)
    a = HologanTrainer(cfg, device=DEV, seed=5)
    b = HologanTrainer(cfg, device=DEV, seed=5)
    if graphs:
        a.enable_cuda_graphs(8)
        b.enable_cuda_graphs(8)
    gen = torch.Generator().manual_seed(1)
    ref, got, pend = [], [], None
    for i in range(7):
        real = torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1
        if i % 2:
            real = real.pin_memory()
        z = torch.rand(8, 128, generator=gen) * 2 - 1
        view = orc.sample_view(8, np.random.RandomState(i))
        ref.append(a.step(real.to(DEV), i, z=z.to(DEV), view=view).item())
        nxt = b.step_host(real, i, z=z, view=view)
        if pend is not None:
            got.append(pend.item())
        pend = nxt
    got.append(float(pend))
    # same kernels on the same inputs (cuDNN's discriminator kernels may reorder their reductions between runs)
    for r, g in zip(ref, got):
        assert abs(r - g) < 5e-3 * max(1.0, abs(r)), (ref, got)
    worst = max(rel_err(pb, pa) for pa, pb in zip(a.generator.parameters(), b.generator.parameters()) if pa.dim() > 1)
    assert worst < 5e-2


@pytest.mark.parametrize("optimizer_idx", [0, 1])
def test_lightning_module_mirror_training_step_vs_oracle(optimizer_idx):
    """`core.lightning_module.HOLOGAN` built from the `+expt=hologan` configuration through the `_target_` strings
    (SURVEY 8-f4), full widths, bf16 autocast (every kernel ours): the step's loss against the oracle's restatement of
    core/lightning_module.py:209-237 on the same parameters, latents and views."""
    from lightning_gan_zoo_b200 import compat
    from lightning_gan_zoo_b200.config import instantiate, load_hologan_config
    compat.install()
    torch.manual_seed(3)
    cfg = load_hologan_config(overrides=["train.batch_size=8"])
    lm = instantiate(cfg.model.lm, cfg, "logs").to(DEV)
    bsz = 8
    view = orc.sample_view(bsz, np.random.RandomState(5))
    lm.generator.sample_view = lambda n: view
    real = torch.rand(bsz, 3, 64, 64, device=DEV) * 2 - 1
    gp = {k: v.detach().cpu().clone() for k, v in lm.generator.state_dict().items()}
    dp = {k: v.detach().cpu().clone() for k, v in lm.discriminator.state_dict().items() if "conv2d_spec_norm" not in k}
    torch.manual_seed(11)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = lm.training_step((real, None), 0, optimizer_idx)
    loss.backward()
    torch.manual_seed(11)
    z = lm.noise_distn.sample((bsz, 128))
    fake = orc.generator_forward(gp, z, view)
    ref, logs = orc.hologan_losses(optimizer_idx, dp, real.cpu(), fake, z)
    assert abs(loss.item() - ref.item()) <= 2e-2 * abs(ref.item()), (loss.item(), ref.item())
    key = "train/d_loss" if optimizer_idx == 0 else "train/g_loss"
    logged = lm.logged if hasattr(lm, "logged") else {}
    if logged:
        assert abs(float(logged[key]) - float(logs[key])) <= 2e-2 * abs(float(logs[key])) + 1e-3
        assert abs(float(logged["train/q_loss"]) - float(logs["train/q_loss"])) <= 2e-2 * float(logs["train/q_loss"]) + 1e-3
    stepped = lm.discriminator if optimizer_idx == 0 else lm.generator
    assert all(p.grad is not None for n, p in stepped.named_parameters() if not n.endswith("conv2d.bias") and "convTranspose.bias" not in n)


def test_flat_adam_matches_torch_adam():
    """hg_adam_step on flat buffers (FlatAdam) against torch.optim.Adam on the same parameters / gradients, three
    steps, with a gradient scale (the data-parallel 1 / world) and an lr change in between; state_dict round trip."""
    from lightning_gan_zoo_b200.training import FlatAdam, _FlatGrads
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(DEV)
    ref.load_state_dict(net.state_dict())
    fg = _FlatGrads(net.parameters())
    opt = FlatAdam(fg, 1e-3, betas=(0.9, 0.999))
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.5 ** e)
    rsched = torch.optim.lr_scheduler.LambdaLR(ropt, lambda e: 0.5 ** e)
    for it in range(3):
        x = torch.randn(8, 37, device=DEV)
        fg.begin()
        (net(x).square().sum() * 4.0).backward()
        fg.finish()
        ropt.zero_grad()
        ref(x).square().sum().backward()
        opt.step(grad_scale=0.25)
        ropt.step()
        sched.step(); rsched.step()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert rel_err(a, b) < 1e-6, it
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 3
    net2 = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).to(DEV)
    opt2 = FlatAdam(_FlatGrads(net2.parameters()), 1e-3)
    opt2.load_state_dict(ropt.state_dict())                    # a torch Adam state loads into the flat buffers
    assert rel_err(opt2.exp_avg_sq[:19 * 37], ropt.state[next(ref.parameters())]["exp_avg_sq"].flatten()) < 1e-6
    assert float(opt2.kstate[0]) == 3
