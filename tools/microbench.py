#!/usr/bin/env python
"""Kernel microbenchmarks (BASELINE.json configs[2] and friends): CUDA-event timing, rotating buffers
larger than L2, achieved GB/s against the measured HBM peak.  Usage: python tools/microbench.py [rotate|adain|all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from lightning_gan_zoo_b200 import ops

DEV = torch.device("cuda")
PEAK = 6546.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def time_rot(fn, bufs, iters=30):
    for b in bufs[:3]:
        fn(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(bufs[i % len(bufs)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def views(b, seed=0):
    rs = np.random.RandomState(seed)
    v = np.zeros((b, 6)); v[:, 0] = np.deg2rad(rs.randint(220, 320, b)); v[:, 1] = np.deg2rad(rs.randint(70, 110, b)); v[:, 2] = 1
    return v


def rotate():
    print(f"# rotate-resample, HBM peak {PEAK} GB/s; algorithmic bytes = 2*B*C*S^3*sizeof")
    for (b, c, s) in [(64, 64, 16), (64, 128, 16), (64, 256, 16), (64, 64, 32), (64, 128, 32)]:
        for dt in (torch.float32, torch.bfloat16):
            a = ops.view_to_affine(views(b), s, s).to(DEV)
            nbytes = 2 * b * c * s ** 3 * (4 if dt == torch.float32 else 2)
            nbuf = max(2, int(600e6 // (nbytes // 2)) + 1)
            nbuf = min(nbuf, 12)
            bufs = [torch.randn(b, c, s, s, s, device=DEV, dtype=dt) for _ in range(nbuf)]
            for border, bn in ((ops.HG_BORDER_REFERENCE, "ref"), (ops.HG_BORDER_ZERO, "zero")):
                tf = time_rot(lambda v: ops.rotate_fwd_raw(v, a, border), bufs)
                tb = time_rot(lambda v: ops.rotate_bwd_raw(v, a, c, s, border), bufs)
                print(f"rotate ({b},{c},{s}^3) {str(dt)[6:]:8s} border={bn:4s} fwd {tf*1e6:8.1f} us {nbytes/tf/1e9:7.0f} GB/s "
                      f"({nbytes/tf/1e9/PEAK*100:4.1f}%)  bwd {tb*1e6:8.1f} us {nbytes/tb/1e9:7.0f} GB/s ({nbytes/tb/1e9/PEAK*100:4.1f}%)")
            del bufs


def adain():
    print(f"# AdaIN+ReLU (NC* layout); fwd bytes = 2*B*C*N*sizeof, bwd = 3*B*C*N*sizeof")
    for (b, c, n) in [(64, 512, 64), (64, 128, 512), (64, 64, 4096), (64, 256, 1024)]:
        for dt in (torch.float32, torch.bfloat16):
            es = 4 if dt == torch.float32 else 2
            nbuf = min(12, max(2, int(600e6 // (b * c * n * es)) + 1))
            xs = [torch.randn(b, c, n, device=DEV, dtype=dt) for _ in range(nbuf)]
            s = torch.rand(b, c, device=DEV); bb = torch.randn(b, c, device=DEV)
            mean = torch.empty(b, c, device=DEV); rstd = torch.empty(b, c, device=DEV)
            y = torch.empty_like(xs[0]); dx = torch.empty_like(xs[0]); ds = torch.empty(b, c, device=DEV); db = torch.empty(b, c, device=DEV)
            from lightning_gan_zoo_b200 import _lib
            P = ops._ptr
            code = ops._dtype_code(xs[0])

            def f(x):
                _lib.call("hg_adain_act_fwd", P(x), P(s), P(bb), P(y), P(mean), P(rstd), b, c, n, c * n, c, 1e-8, 0.0, 0, code, ops._stream())

            def g(x):
                _lib.call("hg_adain_act_bwd", P(x), P(y), P(s), P(bb), P(mean), P(rstd), P(dx), P(ds), P(db), b, c, n, c * n, c, c, 0.0, 0, code, ops._stream())
            tf = time_rot(f, xs); tb = time_rot(g, xs)
            fb, bbt = 2 * b * c * n * es, 3 * b * c * n * es
            print(f"adain ({b},{c},{n}) {str(dt)[6:]:8s} fwd {tf*1e6:7.1f} us {fb/tf/1e9:6.0f} GB/s ({fb/tf/1e9/PEAK*100:4.1f}%)"
                  f"  bwd {tb*1e6:7.1f} us {bbt/tb/1e9:6.0f} GB/s ({bbt/tb/1e9/PEAK*100:4.1f}%)")


def conv():
    """The 15 dense contractions of one generator fwd+bwd at B=64 (bf16, tcgen05), TFLOP/s against the measured
    bf16 matmul peak.  Algorithmic FLOPs = 2 * B * positions * Cin * Cout * taps_used."""
    import ctypes
    from lightning_gan_zoo_b200 import _lib
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    peak = peaks["bf16_tflops"]
    B = 64
    P = ops._ptr
    layers = [("block1", 3, 3, 512, 128, 4), ("block2", 3, 3, 128, 64, 8), ("proj1x1", 2, 1, 1024, 1024, 16),
              ("block3", 2, 4, 1024, 256, 16), ("block4", 2, 4, 256, 64, 32)]
    print(f"# conv kernels at B={B}; bf16 peak {peak} TFLOP/s (burst, measured)")
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for name, ndim, k, cin, cout, size in layers:
        sp = (size,) * ndim
        taps = k ** ndim
        ncls = 1 if k == 1 else 2 ** ndim
        pos = size ** ndim
        flops = 2.0 * B * pos * cin * cout * taps
        nb = 3
        xs = [torch.randn(B, *sp, cin, device=DEV).to(torch.bfloat16) for _ in range(nb)]
        dys = [torch.randn(B, *sp, ncls, cout, device=DEV).to(torch.bfloat16) for _ in range(nb)]
        w = torch.randn(cin, cout, *((k,) * ndim), device=DEV) * 0.02
        wf, wd = ops.pack_convt_weight(w)
        y = torch.empty_like(dys[0]); dx = torch.empty_like(xs[0]); dw = torch.empty_like(w)
        nws = _lib.load().hg_convt_wgrad_workspace_bytes(B, cin, cout, ndim, size, k)
        ws = torch.empty(nws, dtype=torch.uint8, device=DEV)
        st = ops._stream()
        f = lambda i: _lib.call("hg_convt_fwd", P(xs[i]), P(wf), P(None), P(y), B, cin, cout, ndim, size, k, ctypes.c_float(1.0), st)
        g = lambda i: _lib.call("hg_convt_dgrad", P(dys[i]), P(wd), P(dx), B, cin, cout, ndim, size, k, st)
        h = lambda i: _lib.call("hg_convt_wgrad", P(xs[i]), P(dys[i]), P(dw), P(ws), nws, B, cin, cout, ndim, size, k, 0, 0, 0, st)
        row = f"{name:8s} {flops/1e9:7.1f} GF"
        for tag, fn in (("fwd", f), ("dgrad", g), ("wgrad", h)):
            t = time_rot(fn, list(range(nb)), iters=12)
            tot[tag] += t
            row += f" | {tag} {t*1e6:7.1f} us {flops/t/1e12:6.0f} TF/s ({flops/t/1e12/peak*100:4.1f}%)"
        print(row)
    print("totals: " + ", ".join(f"{k} {v*1e6:.0f} us" for k, v in tot.items()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("rotate", "all"):
        rotate()
    if what in ("adain", "all"):
        adain()
    if what in ("conv", "all"):
        conv()
