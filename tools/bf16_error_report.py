#!/usr/bin/env python
"""Per-tensor max-normalised error of the bf16 generator paths vs the fp32 oracle:
   tc    = tcgen05 implicit-GEMM path (this repo);
   lib   = same module with torch/cuDNN bf16 convs under autocast;
   floor = the fp32 CPU oracle itself with bf16 STORAGE emulated (oracle.generator_forward(bf16_storage=True): every
           activation / gradient a bf16 pipeline keeps in memory rounded to bf16, bf16 weight copies, fp32 arithmetic
           everywhere) -- the error any bf16-operand implementation carries, whatever its kernels;
   flips = fraction of ReLU masks that differ between the floor emulation and the fp32 oracle, per stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from types import SimpleNamespace
from lightning_gan_zoo_b200.core.models.hologan_generator import Generator
from oracle import hologan_oracle as orc

def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return (a - b).abs().max().item() / b.abs().max().item()

def relrms(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

for seed, bsz in ((77, 4), (78, 16)):
    gen = torch.Generator().manual_seed(seed)
    p = orc.init_generator_params(64, 3, 128, 64, generator=gen, bias_std=0.05)
    z = torch.rand(bsz, 128, generator=gen) * 2 - 1
    view = orc.sample_view(bsz, np.random.RandomState(seed))
    dout = torch.randn(bsz, 3, 64, 64, generator=gen)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    zr = z.clone().requires_grad_(True)
    st_ref = {}
    ref = orc.generator_forward(pr, zr, view, stages=st_ref); (ref * dout).sum().backward()
    res = {}
    pf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    zf = z.clone().requires_grad_(True)
    st_fl = {}
    fl = orc.generator_forward(pf, zf, view, stages=st_fl, bf16_storage=True); (fl * dout).sum().backward()
    floor = {"out": (rel(fl, ref), relrms(fl, ref)), "dz": (rel(zf.grad, zr.grad), relrms(zf.grad, zr.grad))}
    for k, v in pr.items():
        if not k.endswith("convTranspose.bias"):
            floor[k] = (rel(pf[k].grad, v.grad), relrms(pf[k].grad, v.grad))
    flips = {k: ((st_fl[k] > 0) != (st_ref[k] > 0)).float().mean().item() for k in ("h1", "h2", "h3", "h4", "h5")}
    for mode in ("tc", "lib"):
        net = Generator(64, 3, 128, SimpleNamespace(), 64).cuda(); net.load_state_dict(p)
        if mode == "lib":
            net._use_tensor_core_path = lambda z: False
        zg = z.cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(zg, view_in=view)
        (out.float() * dout.cuda()).sum().backward()
        named = dict(net.named_parameters())
        r = {"out": (rel(out.float(), ref), relrms(out.float(), ref)), "dz": (rel(zg.grad, zr.grad), relrms(zg.grad, zr.grad))}
        for k, v in pr.items():
            if named[k].grad is not None and not k.endswith("convTranspose.bias"):
                r[k] = (rel(named[k].grad, v.grad), relrms(named[k].grad, v.grad))
        res[mode] = r
    print(f"seed {seed} B {bsz}:   tensor           tc max / rms      lib max / rms      floor max / rms")
    for k in res["tc"]:
        a, b, c = res["tc"][k], res["lib"].get(k, (float('nan'),) * 2), floor.get(k, (float('nan'),) * 2)
        print(f"   {k:38s} {a[0]:.4f} {a[1]:.4f}     {b[0]:.4f} {b[1]:.4f}     {c[0]:.4f} {c[1]:.4f}")
    print("   ReLU mask flips (floor emulation vs fp32 oracle): " + ", ".join(f"{k} {v:.5f}" for k, v in flips.items()))
