"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (imported from /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/gen_golden.py

For every case it (1) runs the unmodified reference modules
`core/models/hologan_generator.py` / `core/models/hologan_discriminator.py` on CPU fp32,
(2) runs the restatement in `oracle/hologan_oracle.py` on the same inputs and asserts
agreement (bit-exact where stated), and (3) stores inputs (or their seeds + a sha256 of
the regenerated tensors) and the REFERENCE's outputs as the fixture.  Test
infrastructure only -- nothing in the product imports this.
"""
from __future__ import annotations

import hashlib
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore", message="torch.meshgrid")

from core.models import hologan_generator as ref_g          # noqa: E402  (reference)
from core.models import hologan_discriminator as ref_d      # noqa: E402  (reference)
from oracle import hologan_oracle as orc                    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)


def sha(t) -> str:
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def params_sha(p) -> str:
    h = hashlib.sha256()
    for k in sorted(p):
        h.update(k.encode())
        h.update(p[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def views_mixed() -> np.ndarray:
    """config-range views + axis-aligned + identity + scaled / shifted (SURVEY.md R1 probes)."""
    deg = np.deg2rad
    rows = [
        (deg(220), deg(70), 1.0, 0, 0, 0),
        (deg(270), deg(90), 1.0, 0, 0, 0),      # axis aligned: thousands of integer coords
        (0.0, 0.0, 1.0, 0, 0, 0),               # identity: last plane of each axis -> 0
        (deg(301), deg(100), 0.7, 0, 0, 0),
        (deg(247), deg(85), 1.5, 0, 0, 0),
        (deg(319), deg(109), 1.0, 5.0, -5.0, 2.5),
        (deg(233), deg(75), 1.0, -3.25, 0.5, 4.0),
        (deg(285), deg(95), 0.85, 1.0, 1.0, -1.0),
    ]
    return np.asarray(rows, dtype=np.float64)


def ref_generator(in_planes, img_size=64):
    return ref_g.Generator(in_planes, 3, 128, SimpleNamespace(), img_size, gpu=False)


def case_rotate():
    g = ref_generator(8)
    gen = torch.Generator().manual_seed(101)
    view = views_mixed()
    b = view.shape[0]
    for size, ch, tag in ((16, 5, "s16"), (8, 3, "s8")):
        vol = torch.randn(b, ch, size, size, size, generator=gen)
        gout = torch.randn(b, ch, size, size, size, generator=gen)
        # --- reference: capture the arguments of `interpolation` (coords) -------------
        cap = {}
        orig = g.interpolation

        def spy(voxel, x, y, z, s, _o=orig, _c=cap):
            _c.update(x=x.detach().clone(), y=y.detach().clone(), z=z.detach().clone())
            return _o(voxel, x, y, z, s)
        g.interpolation = spy
        v = vol.clone().requires_grad_(True)
        out = g.transformation3d(v, view, size, size)
        (out * gout).sum().backward()
        g.interpolation = orig
        # --- oracle restatement ---------------------------------------------------------
        a = orc.view_to_affine(view, size, size)
        ox, oy, oz = orc.source_coords(a, size)
        assert torch.equal(ox, cap["x"]) and torch.equal(oy, cap["y"]) and torch.equal(oz, cap["z"]), \
            "oracle coords differ from reference"
        v2 = vol.clone().requires_grad_(True)
        out2 = orc.rotate_resample(v2, view)
        assert torch.equal(out2, out), "oracle rotate output differs from reference (bitwise)"
        (out2 * gout).sum().backward()
        err = (v2.grad - v.grad).abs().max().item() / v.grad.abs().max().item()
        assert err < 1e-6, err
        proj_ref = out.permute(0, 1, 3, 2, 4)
        proj_ref = proj_ref[:, :, torch.arange(size - 1, -1, -1).long(), :, :].reshape(b, -1, size, size)
        assert torch.equal(orc.project_depth_to_channels(out), proj_ref)
        coords = torch.stack([cap["x"], cap["y"], cap["z"]]).reshape(3, b, -1)
        np.savez_compressed(
            os.path.join(OUT, f"rotate_{tag}.npz"),
            view=view, a_inv=a.numpy(), vol=vol.numpy(), grad_out=gout.numpy(),
            coords=coords.numpy(), floor_idx=torch.floor(coords).to(torch.int32).numpy(),
            out=out.detach().numpy(), grad_vol=v.grad.numpy(),
            proj=proj_ref.detach().numpy() if tag == "s16" else np.zeros(0, np.float32),
        )
        print(f"rotate_{tag}: coords sha {sha(coords)} out sha {sha(out)} grad err vs oracle {err:.2e}")


def case_rotate_sweep():
    """100-view sweep of the config range: coords + floor hashes only."""
    g = ref_generator(8)
    az = np.arange(220, 320)
    el = 70 + (np.arange(100) % 40)
    view = np.zeros((100, 6))
    view[:, 0], view[:, 1], view[:, 2] = np.deg2rad(az), np.deg2rad(el), 1.0
    cap = {}
    orig = g.interpolation

    def spy(voxel, x, y, z, s):
        cap.update(x=x.clone(), y=y.clone(), z=z.clone())
        return orig(voxel, x, y, z, s)
    g.interpolation = spy
    vol = torch.zeros(100, 1, 16, 16, 16)
    g.transformation3d(vol, view, 16, 16)
    coords = torch.stack([cap["x"], cap["y"], cap["z"]]).reshape(3, 100, -1)
    a = orc.view_to_affine(view, 16, 16)
    oc = torch.stack(orc.source_coords(a, 16)).reshape(3, 100, -1)
    assert torch.equal(oc, coords)
    np.savez_compressed(os.path.join(OUT, "rotate_sweep100.npz"), view=view, a_inv=a.numpy(),
                        coords_sha=np.array(sha(coords)),
                        floor_sha=np.array(sha(torch.floor(coords).to(torch.int32))),
                        n_integer=np.array(int((coords == torch.floor(coords)).sum())))
    print("rotate_sweep100: coords sha", sha(coords), "exact-integer coords",
          int((coords == torch.floor(coords)).sum()))


def case_adain():
    gen = torch.Generator().manual_seed(202)
    store = {}
    for i, shape in enumerate([(2, 5, 4, 4, 4), (2, 3, 8, 8, 8), (1, 2, 16, 16, 16), (3, 4, 32, 32), (2, 2, 64, 64)]):
        x = (torch.randn(*shape, generator=gen) * 1.7 + 0.3).requires_grad_(True)
        s = torch.rand(shape[0], shape[1], generator=gen).requires_grad_(True)
        bb = torch.randn(shape[0], shape[1], generator=gen).requires_grad_(True)
        dy = torch.randn(*shape, generator=gen)
        y = ref_g.AdaIn(x, s, bb)
        yr = torch.relu(y)
        (yr * dy).sum().backward()
        x2, s2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, s, bb))
        y2 = orc.adain(x2, s2, b2)
        assert torch.equal(y2, y), "oracle adain differs from reference"
        (torch.relu(y2) * dy).sum().backward()
        assert torch.allclose(x2.grad, x.grad, rtol=0, atol=1e-6 * x.grad.abs().max().item())
        store.update({f"x{i}": x.detach().numpy(), f"s{i}": s.detach().numpy(), f"b{i}": bb.detach().numpy(),
                      f"dy{i}": dy.numpy(), f"y{i}": y.detach().numpy(), f"dx{i}": x.grad.numpy(),
                      f"ds{i}": s.grad.numpy(), f"db{i}": bb.grad.numpy()})
    # SURVEY.md appendix A known answer
    torch.manual_seed(1234)
    _ = ref_generator(64)
    _ = ref_d.Discriminator(3, 64, 128)
    _ = torch.rand(8, 128)
    vox = torch.randn(8, 64, 16, 16, 16)
    ka = ref_g.AdaIn(vox[:, :, :4, :4, :4].contiguous(), torch.linspace(0.5, 1.5, 512).view(8, 64),
                     torch.linspace(-1, 1, 512).view(8, 64))
    store["appendixA_absmean"] = np.array(ka.double().abs().mean().item())
    np.savez_compressed(os.path.join(OUT, "adain.npz"), n_cases=np.array(5), **store)
    print("adain: appendix-A abs-mean", ka.double().abs().mean().item())


def grads_summary(named_grads):
    keys = sorted(named_grads)
    return keys, np.array([[named_grads[k].double().sum().item(), named_grads[k].double().abs().sum().item()]
                           for k in keys])


def case_generator(in_planes, bsz, tag, seed, img_size=64, keep_grads=()):
    gen = torch.Generator().manual_seed(seed)
    p = orc.init_generator_params(in_planes, 3, 128, img_size, generator=gen, bias_std=0.05)
    z = torch.rand(bsz, 128, generator=gen) * 2 - 1
    rs = np.random.RandomState(seed)
    view = orc.sample_view(bsz, rs)
    if bsz >= 2:
        view[0, 0], view[0, 1] = np.deg2rad(270), np.deg2rad(90)
    dout = torch.randn(bsz, 3, img_size, img_size, generator=gen)
    g = ref_generator(in_planes, img_size)
    missing = g.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    zr = z.clone().requires_grad_(True)
    hooks = {}
    names = {"block1": "h1", "block2": "h2", "block3": "h4", "block4": "h5"}
    hs = [getattr(g, n).register_forward_hook(lambda m, i, o, _n=names[n]: hooks.__setitem__(_n, o.detach()))
          for n in names]
    out = g(zr, view_in=view)
    for h in hs:
        h.remove()
    (out * dout).sum().backward()
    ref_grads = {k: v.grad.clone() for k, v in g.named_parameters()}
    # oracle
    p2 = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    z2 = z.clone().requires_grad_(True)
    st = {}
    out2 = orc.generator_forward(p2, z2, view, img_size, stages=st)
    assert torch.equal(out2, out), f"oracle G output differs from reference ({tag})"
    for n in ("h1", "h2", "h4", "h5"):
        assert torch.equal(st[n], hooks[n]), n
    (out2 * dout).sum().backward()
    for k in ref_grads:
        den = ref_grads[k].abs().max().item() + 1e-30
        e = (p2[k].grad - ref_grads[k]).abs().max().item() / den
        assert e < 2e-5, (k, e)
    ez = (z2.grad - zr.grad).abs().max().item() / zr.grad.abs().max().item()
    assert ez < 2e-5, ez
    keys, summ = grads_summary(ref_grads)
    store = dict(seed=np.array(seed), in_planes=np.array(in_planes), img_size=np.array(img_size),
                 params_sha=np.array(params_sha(p)), z=z.numpy(), view=view, dout_sha=np.array(sha(dout)),
                 out=out.detach().numpy(), dz=zr.grad.numpy(), grad_keys=np.array(keys), grad_summary=summ,
                 h2_stats=np.array([hooks["h2"].double().sum().item(), hooks["h2"].double().abs().mean().item()]),
                 h5_stats=np.array([hooks["h5"].double().sum().item(), hooks["h5"].double().abs().mean().item()]))
    for k in keep_grads:
        store["grad::" + k] = ref_grads[k].numpy()
    np.savez_compressed(os.path.join(OUT, f"generator_{tag}.npz"), **store)
    print(f"generator_{tag}: out sum {out.double().sum().item():.10f} params sha {params_sha(p)}")


def case_discriminator_and_step():
    seed = 404
    gen = torch.Generator().manual_seed(seed)
    dp = orc.init_discriminator_params(3, 8, 128, 64, generator=gen)
    gp = orc.init_generator_params(8, 3, 128, 64, generator=gen, bias_std=0.05)
    bsz = 4
    real = torch.rand(bsz, 3, 64, 64, generator=gen) * 2 - 1
    z = torch.rand(bsz, 128, generator=gen) * 2 - 1
    view = orc.sample_view(bsz, np.random.RandomState(seed))
    d = ref_d.Discriminator(3, 8, 128)
    sd = {k: v.clone() for k, v in dp.items()}
    # the reference registers each spectrally-normalised conv under two names (`conv2d` and
    # `conv2d_spec_norm` are the same module, hologan_discriminator.py:12-15): alias the keys
    sd.update({k.replace(".conv2d.", ".conv2d_spec_norm."): v for k, v in sd.items() if k.startswith("blocks.")})
    d.load_state_dict(sd, strict=True)
    g = ref_generator(8)
    g.load_state_dict({k: v.clone() for k, v in gp.items()}, strict=True)
    bce = torch.nn.BCEWithLogitsLoss()
    store = dict(seed=np.array(seed), d_params_sha=np.array(params_sha(dp)), g_params_sha=np.array(params_sha(gp)),
                 real_sha=np.array(sha(real)), z=z.numpy(), view=view)
    d.train()
    # ---- optimizer_idx 0 (core/lightning_module.py:217-228), restated call-for-call on the reference modules
    fake = g(z, view_in=view)
    d_real, _ = d(real)
    l_real = bce(d_real, torch.ones_like(d_real))
    d_fake, zp = d(fake.detach())
    l_fake = bce(d_fake, torch.zeros_like(d_fake))
    loss_d = (l_real + l_fake) / 2 + torch.mean((zp - z) ** 2)
    d.zero_grad()
    loss_d.backward()
    dgr = {k: v.grad.clone() for k, v in d.named_parameters()}
    keys, summ = grads_summary(dgr)
    store.update(loss_d=np.array(loss_d.item()), d_real=d_real.detach().numpy(), d_fake=d_fake.detach().numpy(),
                 zp_d=zp.detach().numpy(), d_grad_keys=np.array(keys), d_grad_summary=summ)
    for i in range(3):
        store[f"u_after_dstep_{i}"] = d.blocks[i].conv2d.weight_u.detach().numpy().copy()
    # oracle on the same step
    dp2 = {k: (v.clone().requires_grad_(True) if "weight_u" not in k and "weight_v" not in k else v.clone())
           for k, v in dp.items()}
    fake2 = orc.generator_forward(gp, z, view)
    assert torch.equal(fake2, fake)
    loss2, _ = orc.hologan_losses(0, dp2, real, fake2, z)
    assert abs(loss2.item() - loss_d.item()) < 1e-6, (loss2.item(), loss_d.item())
    loss2.backward()
    for k in dgr:
        e = (dp2[k].grad - dgr[k]).abs().max().item() / (dgr[k].abs().max().item() + 1e-30)
        assert e < 2e-5, (k, e)
    for i in range(3):
        assert torch.allclose(dp2[f"blocks.{i}.conv2d.weight_u"], d.blocks[i].conv2d.weight_u, atol=1e-6)
    # ---- optimizer_idx 1 (:231-237)
    d.zero_grad(); g.zero_grad()
    fake = g(z, view_in=view)
    o, zp = d(fake)
    loss_g = bce(o, torch.ones_like(o)) + torch.mean((zp - z) ** 2)
    loss_g.backward()
    ggr = {k: v.grad.clone() for k, v in g.named_parameters()}
    keys, summ = grads_summary(ggr)
    store.update(loss_g=np.array(loss_g.item()), g_grad_keys=np.array(keys), g_grad_summary=summ,
                 fake=fake.detach().numpy())
    gp2 = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    dp3 = {k: v.detach().clone() for k, v in dp2.items()}
    fake3 = orc.generator_forward(gp2, z, view)
    loss3, _ = orc.hologan_losses(1, dp3, None, fake3, z)
    assert abs(loss3.item() - loss_g.item()) < 1e-6, (loss3.item(), loss_g.item())
    loss3.backward()
    for k in ggr:
        e = (gp2[k].grad - ggr[k]).abs().max().item() / (ggr[k].abs().max().item() + 1e-30)
        assert e < 5e-5, (k, e)
    np.savez_compressed(os.path.join(OUT, "train_step_tiny.npz"), **store)
    print(f"train_step_tiny: loss_d {loss_d.item():.8f} loss_g {loss_g.item():.8f}")


def case_appendix_a():
    """SURVEY.md Appendix A recipe on the full-size reference modules; stores scalars only."""
    torch.manual_seed(1234)
    g = ref_generator(64)
    d = ref_d.Discriminator(3, 64, 128)
    z = torch.rand(8, 128) * 2 - 1
    view = np.zeros((8, 6))
    view[:, 0] = np.deg2rad([220, 233, 247, 260, 270, 285, 301, 319])
    view[:, 1] = np.deg2rad([70, 75, 80, 85, 90, 95, 100, 109])
    view[:, 2] = 1
    vox = torch.randn(8, 64, 16, 16, 16)
    rot = g.transformation3d(vox, view, 16, 16)
    out = g(z, view_in=view)
    # oracle on the reference's own parameters
    p = {k: v.detach().clone() for k, v in g.state_dict().items()}
    assert torch.equal(orc.generator_forward(p, z, view), out)
    assert torch.equal(orc.rotate_resample(vox, view), rot)
    np.savez_compressed(os.path.join(OUT, "appendix_a.npz"), view=view,
                        rot_sum=np.array(rot.double().sum().item()),
                        rot_absmean=np.array(rot.double().abs().mean().item()),
                        out_sum=np.array(out.double().sum().item()),
                        out_absmean=np.array(out.double().abs().mean().item()),
                        vox_sha=np.array(sha(vox)))
    print("appendix_a: rot sum", rot.double().sum().item(), "out sum", out.double().sum().item())


def case_state_dict_spec():
    """Key order + shapes of the reference modules' state_dicts."""
    import json
    spec = {}
    for tag, mod in (("generator_64", ref_generator(64, 64)), ("discriminator_64", ref_d.Discriminator(3, 64, 128))):
        spec[tag] = [[k, list(v.shape)] for k, v in mod.state_dict().items()]
    with open(os.path.join(OUT, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f, indent=0)
    print("state_dict_spec:", {k: len(v) for k, v in spec.items()})


if __name__ == "__main__":
    case_state_dict_spec()
    case_rotate()
    case_rotate_sweep()
    case_adain()
    case_generator(8, 3, "p8", seed=303, keep_grads=("x", "block2.convTranspose.weight", "block4.convTranspose.bias",
                                                      "zMapping.linear1.weight"))
    case_generator(16, 2, "p16", seed=304)
    case_discriminator_and_step()
    case_appendix_a()
    print("golden fixtures written to", OUT)
