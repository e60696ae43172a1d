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


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("rotate", "all"):
        rotate()
    if what in ("adain", "all"):
        adain()
