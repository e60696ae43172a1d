"""Host-side mirror of the reference interface (no GPU needed)."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from lightning_gan_zoo_b200 import compat, ops
from lightning_gan_zoo_b200.core.models import hologan_discriminator as D
from lightning_gan_zoo_b200.core.models import hologan_generator as G
from oracle import hologan_oracle as orc

VIEW_ARGS = SimpleNamespace(azimuth_low=220, azimuth_high=320, elevation_low=70, elevation_high=110, scale_low=1,
                            scale_high=1, transX_low=0, transX_high=0, transY_low=0, transY_high=0, transZ_low=0,
                            transZ_high=0, batch_size=32)


@pytest.mark.parametrize("tag,size", [("s16", 16), ("s8", 8)])
def test_view_to_affine_is_bit_exact(tag, size):
    g = load_golden(f"rotate_{tag}.npz")
    a = ops.view_to_affine(g["view"], size, size)
    assert a.dtype == torch.float32 and tuple(a.shape) == (g["view"].shape[0], 4, 4)
    assert np.array_equal(a.numpy(), g["a_inv"])
    # torch fp32 views (figure callbacks pass tensors: core/figures/types.py:235,312)
    a2 = ops.view_to_affine(torch.from_numpy(g["view"]).float(), size, size)
    assert np.array_equal(a2.numpy(), g["a_inv"])
    assert np.array_equal(a.numpy()[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (a.shape[0], 1)))


def test_view_to_affine_sweep():
    g = load_golden("rotate_sweep100.npz")
    assert np.array_equal(ops.view_to_affine(g["view"]).numpy(), g["a_inv"])
    with pytest.raises(ValueError):
        ops.view_to_affine(np.zeros((4, 5)))


def test_view_to_affine_batched_sweep_equals_per_view():
    """`Generator.render_views` converts all B*V views in one host pass: each matrix must carry the same bits as the
    per-view call the reference's figure callbacks make (core/figures/types.py:233-237)."""
    rs = np.random.RandomState(3)
    b, v = 5, 36
    views = np.zeros((b, v, 6))
    views[..., 0] = np.deg2rad(np.linspace(220, 320, v))[None]
    views[..., 1] = np.deg2rad(rs.randint(70, 110, (b, 1)))
    views[..., 2] = 1.0
    views[..., 3:] = rs.uniform(-1, 1, (b, 1, 3))
    a = ops.view_to_affine(views.reshape(b * v, 6)).reshape(b, v, 4, 4)
    for i in range(v):
        assert torch.equal(a[:, i], ops.view_to_affine(views[:, i]))


def test_state_dict_keys_match_reference():
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))
    g = G.Generator(64, 3, 128, VIEW_ARGS, 64, gpu=False)
    got = [[k, list(v.shape)] for k, v in g.state_dict().items()]
    assert got == spec["generator_64"]
    d = D.Discriminator(3, 64, 128)
    got = [[k, list(v.shape)] for k, v in d.state_dict().items()]
    assert got == spec["discriminator_64"]
    assert sum(p.numel() for p in g.parameters()) == 7795907
    assert sum(p.numel() for p in d.parameters()) == 5379969


def test_patched_128_heads():
    g = G.Generator(16, 3, 128, VIEW_ARGS, 128, gpu=False)
    assert tuple(g.final_layer.weight.shape) == (16, 3, 4, 4) and g.final_layer.stride == (2, 2)
    d = D.Discriminator(3, 8, 128, img_size=128)
    assert d.linear1.in_features == 8 * 8 * 8 * 8
    with pytest.raises(ValueError):
        G.Generator(16, 3, 128, VIEW_ARGS, 96, gpu=False)


def test_sample_view_rng_parity():
    g = G.Generator(8, 3, 128, VIEW_ARGS, 64, gpu=False)
    np.random.seed(7)
    mine = g.sample_view(16)
    theirs = orc.sample_view(16, np.random.RandomState(7))
    assert mine.dtype == np.float64 and np.array_equal(mine, theirs)
    deg = np.rad2deg(mine[:, 0])
    assert (deg > 219.999).all() and (deg < 319.001).all() and np.allclose(deg, np.round(deg))


def test_cpu_tensors_fail_loudly():
    g = G.Generator(8, 3, 128, VIEW_ARGS, 64, gpu=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(2, 128), view_in=np.zeros((2, 6)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.AdaIn(torch.zeros(1, 4, 8), torch.ones(1, 4), torch.zeros(1, 4))


def test_compat_aliases():
    compat.install()
    import importlib
    m = importlib.import_module("core.models.hologan_generator")
    assert m.Generator is G.Generator
    assert importlib.import_module("core.models.hologan_discriminator").Discriminator is D.Discriminator


def test_discriminator_matches_oracle_on_cpu():
    """The stock-torch discriminator mirror against the functional oracle (training-mode power iteration)."""
    gen = torch.Generator().manual_seed(5)
    dp = orc.init_discriminator_params(3, 8, 128, 64, generator=gen)
    d = D.Discriminator(3, 8, 128)
    sd = {k: v.clone() for k, v in dp.items()}
    sd.update({k.replace(".conv2d.", ".conv2d_spec_norm."): v for k, v in sd.items() if k.startswith("blocks.")})
    d.load_state_dict(sd)
    x = torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1
    d.train()
    l1, z1 = d(x)
    l2, z2 = orc.discriminator_forward({k: v.clone() for k, v in dp.items()}, x, training=True)
    assert torch.allclose(l1, l2, atol=1e-6) and torch.allclose(z1, z2, atol=1e-6)


def _cpu_trainer(seed):
    from lightning_gan_zoo_b200.training import HologanConfig, HologanTrainer
    cfg = HologanConfig(batch_size=4, gen_in_planes=8, disc_out_planes=8, num_epochs=4)
    return HologanTrainer(cfg, device="cpu", compute_dtype=torch.float32, seed=seed)


def _cpu_d_step(tr, seed):
    """One discriminator optimizer step on given fakes (the D branch of lightning_module.py:217-228) -- the generator has
    no CPU path, the discriminator mirror does."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(seed)
    real = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    fake = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    z = torch.rand(4, 128, generator=gen) * 2 - 1
    bce = F.binary_cross_entropy_with_logits
    tr.d_grads.zero()
    d_real, _ = tr.discriminator(real)
    d_fake, z_pred = tr.discriminator(fake)
    ((bce(d_real, torch.ones_like(d_real)) + bce(d_fake, torch.zeros_like(d_fake))) / 2 + torch.mean((z_pred - z) ** 2)).backward()
    tr.opt_d.step()


def test_checkpoint_round_trip_and_reference_layout(tmp_path):
    """Checkpoint / resume in the Lightning layout the reference writes (run_network.py:48-50,61-71): keys
    `generator.*` / `discriminator.*` exactly as in the reference's state_dict, optimizer order [D, G]; a resumed
    trainer continues bit-identically; a reference-style .ckpt (only `state_dict`) loads too."""
    a = _cpu_trainer(seed=1)
    for i in range(2):
        _cpu_d_step(a, 10 + i)
    a.end_epoch()
    a.sample_noise(3); a.sample_view(3)
    path = tmp_path / "model.ckpt"
    torch.save(a.checkpoint(epoch=1, global_step=2), path)
    ckpt = torch.load(path, weights_only=False)

    spec = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))
    want = {"generator." + k for k, _ in spec["generator_64"]} | {"discriminator." + k for k, _ in spec["discriminator_64"]}
    assert set(ckpt["state_dict"]) == want
    assert all(v.is_contiguous() and v.device.type == "cpu" for v in ckpt["state_dict"].values())
    assert ckpt["epoch"] == 1 and ckpt["global_step"] == 2 and len(ckpt["optimizer_states"]) == 2

    b = _cpu_trainer(seed=2)                                   # different init, different RNG streams
    b.load_checkpoint(ckpt)
    for (k, va), vb in zip(a.discriminator.state_dict().items(), b.discriminator.state_dict().values()):
        assert torch.equal(va, vb), k
    for va, vb in zip(a.generator.state_dict().values(), b.generator.state_dict().values()):
        assert torch.equal(va, vb)
    assert b.opt_d.param_groups[0]["lr"] == a.opt_d.param_groups[0]["lr"]
    assert torch.equal(a.sample_noise(5), b.sample_noise(5)) and np.array_equal(a.sample_view(5), b.sample_view(5))
    _cpu_d_step(a, 99)
    _cpu_d_step(b, 99)                                         # Adam moments and step count were restored
    for pa, pb in zip(a.discriminator.parameters(), b.discriminator.parameters()):
        assert torch.equal(pa, pb)

    # a reference checkpoint carries no optimizer layout we rely on: weights only
    c = _cpu_trainer(seed=3)
    c.load_checkpoint({"state_dict": ckpt["state_dict"]})
    for (k, v) in c.generator.state_dict().items():
        assert torch.equal(v, ckpt["state_dict"]["generator." + k]), k
    with pytest.raises(RuntimeError):
        c.load_checkpoint({"state_dict": {"generator.x": torch.zeros(1)}})             # strict: missing / mismatching keys


# per-dimension tap tables of lightning_gan_zoo_b200/csrc/conv_gemm.cu::dim_taps -- out[2 i + p] += in[i + s] * w[k]
_DIM_TAPS = {
    4: {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]},                     # k4 s2 p1
    3: {0: [(1, 0)], 1: [(0, 1), (2, 0)]},                              # k3 s2 p1 op1
    5: {0: [(0, 1), (2, 0), (4, -1)], 1: [(1, 1), (3, 0)]},             # k5 s2 p2 op1
}


def _convt2d_by_taps(x, w, kernel):
    """Transposed convolution as the tap GEMMs compute it: per output parity class a sum over (k, shift) taps of the
    zero-padded, shifted input times one weight slice (x: (B,Cin,S,S), w: (Cin,Cout,k,k)) -> (B,Cout,2S,2S)."""
    b, cin, s, _ = x.shape
    cout = w.shape[1]
    y = torch.zeros(b, cout, 2 * s, 2 * s, dtype=x.dtype)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    for py, ty in _DIM_TAPS[kernel].items():
        for px, tx in _DIM_TAPS[kernel].items():
            acc = torch.zeros(b, cout, s, s, dtype=x.dtype)
            for ky, sy in ty:
                for kx, sx in tx:
                    shifted = xp[:, :, 1 + sy:1 + sy + s, 1 + sx:1 + sx + s]              # in[i + s], zero outside
                    acc += torch.einsum("bchw,cd->bdhw", shifted, w[:, :, ky, kx])
            y[:, :, py::2, px::2] = acc
    return y


def _convt2d_dgrad_by_taps(dy, w, kernel):
    """The dgrad GEMM as conv_gemm.cu::hg_convt_dgrad sets it up: dX[i] = sum over (class, tap) of the class plane of
    dY shifted by -s, times W[k]^T (dy: (B,Cout,2S,2S), w: (Cin,Cout,k,k)) -> (B,Cin,S,S)."""
    b, cout, s2, _ = dy.shape
    s = s2 // 2
    dx = torch.zeros(b, w.shape[0], s, s, dtype=dy.dtype)
    for py, ty in _DIM_TAPS[kernel].items():
        for px, tx in _DIM_TAPS[kernel].items():
            plane = torch.nn.functional.pad(dy[:, :, py::2, px::2], (1, 1, 1, 1))        # dY_s2d[:, class (py, px)]
            for ky, sy in ty:
                for kx, sx in tx:
                    shifted = plane[:, :, 1 - sy:1 - sy + s, 1 - sx:1 - sx + s]          # dY_s2d[i - s]
                    dx += torch.einsum("bdhw,cd->bchw", shifted, w[:, :, ky, kx])
    return dx


@pytest.mark.parametrize("kernel,pad,opad", [(4, 1, 0), (5, 2, 1)])
def test_tap_tables_reproduce_transposed_conv_and_the_conv_duality(kernel, pad, opad):
    """The tap tables the tcgen05 implicit-GEMM kernels are driven by (conv_gemm.cu::dim_taps), restated here, against
    torch: k4 (the generator's ConvTranspose2d) and k5 (s2, p2, op1).  For k5 the adjoint of that transposed
    convolution is the discriminator's Conv2d(k5, s2, p2) (core/models/hologan_discriminator.py:12): the same tables
    run it as a 'dgrad' on a space-to-depth input -- checked through autograd."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(kernel)
    x = torch.randn(2, 6, 8, 8, generator=gen, dtype=torch.float64)
    w = torch.randn(6, 4, kernel, kernel, generator=gen, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, stride=2, padding=pad, output_padding=opad)
    got = _convt2d_by_taps(x, w, kernel)
    assert tuple(got.shape) == tuple(ref.shape) and torch.allclose(got, ref, atol=1e-12)
    if kernel == 5:
        # Conv2d(k5, s2, p2) with weight (Cout_c = 6, Cin_c = 4, 5, 5) == adjoint of the transposed conv above
        img = torch.randn(2, 4, 16, 16, generator=gen, dtype=torch.float64)
        conv = F.conv2d(img, w, stride=2, padding=2)                                      # (2, 6, 8, 8)
        xg = x.clone().requires_grad_(True)
        (_convt2d_by_taps(xg, w, 5) * img).sum().backward()                               # <convT(x), img> = <x, conv(img)>
        assert torch.allclose(xg.grad, conv, atol=1e-12)
        # ... and explicitly in the form ops.conv5x5_s2 calls it: hg_convt_dgrad on the space-to-depth image
        assert torch.allclose(_convt2d_dgrad_by_taps(img, w, 5), conv, atol=1e-12)
    else:
        dy = torch.randn(2, 4, 16, 16, generator=gen, dtype=torch.float64)
        xg = x.clone().requires_grad_(True)
        (F.conv_transpose2d(xg, w, stride=2, padding=pad) * dy).sum().backward()
        assert torch.allclose(_convt2d_dgrad_by_taps(dy, w, kernel), xg.grad, atol=1e-12)


def test_torch_library_registration():
    """`torch.ops.hologan.*`: every operator has a schema; the leaf ops have fake (meta) kernels and an
    autograd formula; there is no CPU kernel behind them."""
    import lightning_gan_zoo_b200.torch_ops as t
    for name in t.REGISTERED:
        assert hasattr(torch.ops.hologan, name), name
        assert getattr(torch.ops.hologan, name).default._schema is not None
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        vol = torch.empty(2, 16, 16, 16, 8, device="cuda", dtype=torch.bfloat16)
        a = torch.empty(2, 4, 4, device="cuda")
        out = torch.ops.hologan.rotate_resample(vol, a, 1, 1, 2)            # NDHWC -> PROJ
        assert tuple(out.shape) == (2, 16, 16, 16, 8) and out.dtype == torch.bfloat16
        g = torch.ops.hologan.rotate_resample_backward(out, a, 8, 16, 1, 1, 2)
        assert tuple(g.shape) == (2, 16, 16, 16, 8)
        y = torch.ops.hologan.convt_forward(torch.empty(2, 16, 16, 128, device="cuda", dtype=torch.bfloat16),
                                            torch.empty(16, 64, 128, device="cuda", dtype=torch.bfloat16), None, 2, 4, 1.0)
        assert tuple(y.shape) == (2, 16, 16, 4, 64)
    with pytest.raises(NotImplementedError):
        torch.ops.hologan.rotate_resample(torch.zeros(1, 2, 16, 16, 16), torch.eye(4).unsqueeze(0), 0, 0, 0)


def test_hologan_config_and_lightning_adapter():
    """the `+expt=hologan` configuration (shipped flat YAML, `${...}` interpolation, `_target_`
    instantiation) and the `core.lightning_module.HOLOGAN` mirror: the module tree is built from the reference's dotted
    `_target_` paths (resolved to this package by compat.install), `configure_optimizers` reproduces the [D, G, G]
    frequency schedule, Adam hyper-parameters and the LambdaLR of core/utils/hologan.py:3-9."""
    from lightning_gan_zoo_b200 import compat
    from lightning_gan_zoo_b200.config import instantiate, load_hologan_config
    compat.install()
    cfg = load_hologan_config(overrides=["generator.gpu=false", "generator.in_planes=8"])
    assert cfg.optimiser.betas == [0.9, 0.999] and cfg.optimiser.lr == 1e-4
    assert cfg.generator.view_args.batch_size == 32 and cfg.generator.img_size == 64
    assert cfg.model.noise_distn._target_ == "torch.distributions.uniform.Uniform"
    lm = instantiate(cfg.model.lm, cfg, "logs")
    import lightning_gan_zoo_b200.core.lightning_module as lmod
    import lightning_gan_zoo_b200.core.models.hologan_generator as gmod
    assert type(lm) is lmod.HOLOGAN and type(lm.generator) is gmod.Generator
    assert lm.generator.view_args.azimuth_high == 320 and tuple(lm.fixed_noise.shape) == (8, 128)
    assert float(lm.fixed_noise.min()) >= -1 and float(lm.fixed_noise.max()) <= 1
    (d, g) = lm.configure_optimizers()
    assert d["frequency"] == 1 and g["frequency"] == 2
    assert isinstance(d["optimizer"], torch.optim.Adam) and d["optimizer"].defaults["lr"] == 1e-4
    assert tuple(g["optimizer"].defaults["betas"]) == (0.9, 0.999)
    lam = g["lr_scheduler"].lr_lambdas[0]
    assert lam(0) == 1 and lam(12) == 1 and abs(lam(13) - (1 - 0.5 / 12.5)) < 1e-12 and abs(lam(25)) < 1e-12
    # the schedule Lightning derives from the frequencies
    from lightning_gan_zoo_b200.training import optimizer_index
    assert [optimizer_index(i, d["frequency"], g["frequency"]) for i in range(6)] == [0, 1, 1, 0, 1, 1]


@pytest.mark.skipif(not os.path.isdir("/root/reference/conf"), reason="reference tree not present (GPU box)")
def test_shipped_config_equals_reference_hydra_composition():
    """The flat YAML shipped with the package equals what Hydra composes from the reference's own conf tree for
    `+expt=hologan` on every key the hot path reads."""
    from lightning_gan_zoo_b200.config import load_hologan_config
    ours, ref = load_hologan_config(), load_hologan_config("/root/reference/conf")
    for key in ("name", "model", "optimisation", "optimiser", "disc_optimiser", "gen_optimiser", "generator", "noise_distn",
                "lr_scheduler"):
        a, b = ours[key], ref[key]
        assert (a.to_dict() if hasattr(a, "to_dict") else a) == (b.to_dict() if hasattr(b, "to_dict") else b), key
    t_ours, t_ref = ours.train.to_dict(), ref.train.to_dict()
    assert all(t_ref[k] == v for k, v in t_ours.items())
    d_ours, d_ref = ours.discriminator.to_dict(), ref.discriminator.to_dict()
    assert all(d_ref[k] == v for k, v in d_ours.items())
    # the reference's tree also hands img_size / final_sigmoid to the discriminator (which its own class rejects)
    assert d_ref["img_size"] == 64 and d_ref["final_sigmoid"] is False
    from lightning_gan_zoo_b200 import compat
    from lightning_gan_zoo_b200.config import instantiate
    compat.install()
    disc = instantiate(ref.discriminator)
    assert disc.linear1.in_features == 8192
