#!/usr/bin/env python
"""Launch each hand-written HBM-bound kernel of the bf16 pipeline a few times at its hot-path shape (B = 64) so
that `ncu --set full -k regex:hg::` captures them in isolation:

    ncu --set full --clock-control none --import-source on -k regex:hg:: -o gpurun_out/prof python tools/ncu_targets.py [names...]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from lightning_gan_zoo_b200 import _lib, ops

DEV = torch.device("cuda")
B = 64
bf = torch.bfloat16
want = set(sys.argv[1:])


def on(name):
    return not want or name in want


def views(b, seed=0):
    rs = np.random.RandomState(seed)
    v = np.zeros((b, 6)); v[:, 0] = np.deg2rad(rs.randint(220, 320, b)); v[:, 1] = np.deg2rad(rs.randint(70, 110, b)); v[:, 2] = 1
    return v


REPS = int(os.environ.get("HG_NCU_REPS", "2"))
if on("adain_cl"):
    for ndim, size, classes, c in [(2, 16, 4, 256), (2, 32, 4, 64)]:
        n = size ** ndim * classes
        x = torch.randn(B, *([size] * ndim), classes, c, device=DEV).to(bf).requires_grad_(True)
        sc = torch.rand(B, c, device=DEV); bi = torch.randn(B, c, device=DEV)
        for _ in range(REPS):
            y = ops.adain_act_channels_last(x, sc, bi, ndim, classes)
            y.backward(torch.randn_like(y))
if on("rotate_cl"):
    s, c = 16, 64
    a = ops.view_to_affine(views(B), s, s).to(DEV)
    vol = torch.randn(B, s, s, s, c, device=DEV).to(bf)
    for _ in range(REPS):
        ops.rotate_fwd_raw(vol, a, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ)
        ops.rotate_bwd_raw(vol, a, c, s, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ)
if on("rotate"):
    s, c = 16, 64
    a = ops.view_to_affine(views(B), s, s).to(DEV)
    vol = torch.randn(B, c, s, s, s, device=DEV)
    for _ in range(REPS):
        ops.rotate_fwd_raw(vol, a, ops.HG_BORDER_REFERENCE)
        ops.rotate_bwd_raw(vol, a, c, s, ops.HG_BORDER_REFERENCE)
if on("final_conv"):
    x = torch.randn(B, 64, 64, 64, device=DEV).to(bf).requires_grad_(True)
    w = (torch.randn(3, 64, 3, 3, device=DEV) * 0.02).requires_grad_(True); bias = torch.zeros(3, device=DEV, requires_grad=True)
    for _ in range(REPS):
        o = ops.final_conv_tanh(x, w, bias)
        o.backward(torch.randn_like(o))
if on("pack"):
    for shape in [(1024, 256, 4, 4), (512, 128, 3, 3, 3), (1024, 1024, 1, 1)]:
        wt = torch.randn(*shape, device=DEV)
        for _ in range(REPS):
            ops.pack_convt_weight(wt)
if on("conv"):
    for name, ndim, k, cin, cout, size in [("block3", 2, 4, 1024, 256, 16), ("block4", 2, 4, 256, 64, 32), ("block2", 3, 3, 128, 64, 8),
                                           ("proj", 2, 1, 1024, 1024, 16), ("block1", 3, 3, 512, 128, 4)]:
        x = torch.randn(B, *([size] * ndim), cin, device=DEV).to(bf).requires_grad_(True)
        w = (torch.randn(cin, cout, *([k] * ndim), device=DEV) * 0.02).requires_grad_(True)
        for _ in range(REPS):
            y = ops.convt(x, w, None, ndim, k)
            y.backward(torch.randn_like(y))
if on("disc"):
    from lightning_gan_zoo_b200.core.models.hologan_discriminator import Discriminator
    d = Discriminator(3, 64, 128).to(DEV)
    x = (torch.rand(B, 3, 64, 64, device=DEV) * 2 - 1).requires_grad_(True)
    z = torch.rand(B, 128, device=DEV) * 2 - 1
    for _ in range(REPS):
        with torch.autocast("cuda", dtype=bf):
            lg, zp = d(x)
        loss, _ = ops.hologan_g_loss(lg, zp, z)
        loss.backward()
torch.cuda.synchronize()
print("done")
