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


def time_rot(fn, bufs, iters=30, graph=True):
    """Seconds per call of fn over rotating buffers.  The calls are captured into ONE CUDA graph and the graph
    replay is timed with CUDA events, so host-side launch overhead (ctypes, allocator, autograd) does not hide
    kernels that are shorter than a Python call."""
    for b in bufs[:3]:
        fn(b)
    torch.cuda.synchronize()
    if not graph:                                   # autograd-driven calls cannot be stream-captured
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(bufs[i % len(bufs)])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(bufs[0])
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                fn(bufs[i % len(bufs)])
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def views(b, seed=0):
    rs = np.random.RandomState(seed)
    v = np.zeros((b, 6)); v[:, 0] = np.deg2rad(rs.randint(220, 320, b)); v[:, 1] = np.deg2rad(rs.randint(70, 110, b)); v[:, 2] = 1
    return v


def rotate(shapes=None, tune=0):
    print(f"# rotate-resample, HBM peak {PEAK} GB/s; algorithmic bytes = 2*B*C*S^3*sizeof")
    for (b, c, s) in shapes or [(64, 64, 16), (64, 128, 16), (64, 256, 16), (64, 64, 32), (64, 128, 32)]:
        for dt in (torch.float32, torch.bfloat16):
            a = ops.view_to_affine(views(b), s, s).to(DEV)
            nbytes = 2 * b * c * s ** 3 * (4 if dt == torch.float32 else 2)
            nbuf = max(2, int(600e6 // (nbytes // 2)) + 1)
            nbuf = min(nbuf, 12)
            bufs = [torch.randn(b, c, s, s, s, device=DEV, dtype=dt) for _ in range(nbuf)]
            for border, bn in ((ops.HG_BORDER_REFERENCE, "ref"), (ops.HG_BORDER_ZERO, "zero")):
                tf = time_rot(lambda v: ops.rotate_fwd_raw(v, a, border | tune), bufs)
                tb = time_rot(lambda v: ops.rotate_bwd_raw(v, a, c, s, border | tune), bufs)
                print(f"rotate ({b},{c},{s}^3) {str(dt)[6:]:8s} border={bn:4s} fwd {tf*1e6:8.1f} us {nbytes/tf/1e9:7.0f} GB/s "
                      f"({nbytes/tf/1e9/PEAK*100:4.1f}%)  bwd {tb*1e6:8.1f} us {nbytes/tb/1e9:7.0f} GB/s ({nbytes/tb/1e9/PEAK*100:4.1f}%)")
                if s == 32 and os.environ.get("HG_BENCH_SLAB32", "0") != "0":
                    # A/B against the kernels that were the default before r02a (per-channel slab fwd, smem scatter bwd)
                    from lightning_gan_zoo_b200 import _lib
                    _lib.set_option("ROTATE_SLAB32", 0)
                    ts = time_rot(lambda v: ops.rotate_fwd_raw(v, a, border | tune), bufs)
                    _lib.set_option("ROTATE_SLAB32", 1)
                    _lib.set_option("ROTATE_GATHER_BWD", 0)
                    tg = time_rot(lambda v: ops.rotate_bwd_raw(v, a, c, s, border | tune), bufs)
                    _lib.set_option("ROTATE_GATHER_BWD", 1)
                    print(f"rotate ({b},{c},{s}^3) {str(dt)[6:]:8s} border={bn:4s} fwd [old slab] {ts*1e6:8.1f} us "
                          f"{nbytes/ts/1e9:7.0f} GB/s ({nbytes/ts/1e9/PEAK*100:4.1f}%)  bwd [old scatter] {tg*1e6:8.1f} us "
                          f"{nbytes/tg/1e9:7.0f} GB/s ({nbytes/tg/1e9/PEAK*100:4.1f}%)")
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
        f = lambda i: _lib.call("hg_convt_fwd", P(xs[i]), P(wf), P(None), P(y), B, cin, cout, ndim, size, k, ctypes.c_float(1.0), ops._stream())
        g = lambda i: _lib.call("hg_convt_dgrad", P(dys[i]), P(wd), P(dx), B, cin, cout, ndim, size, k, ops._stream())
        h = lambda i: _lib.call("hg_convt_wgrad", P(xs[i]), P(dys[i]), P(dw), P(ws), nws, B, cin, cout, ndim, size, k, 0, 0, 0, ops._stream())
        row = f"{name:8s} {flops/1e9:7.1f} GF"
        for tag, fn in (("fwd", f), ("dgrad", g), ("wgrad", h)):
            t = time_rot(fn, list(range(nb)), iters=12)
            tot[tag] += t
            row += f" | {tag} {t*1e6:7.1f} us {flops/t/1e12:6.0f} TF/s ({flops/t/1e12/peak*100:4.1f}%)"
        if k != 1:          # forward with the AdaIN statistics epilogue
            nst = _lib.load().hg_convt_stats_floats(B, cout, ndim, size, k)
            st = torch.empty(nst, device=DEV)
            fs = lambda i: _lib.call("hg_convt_fwd_stats", P(xs[i]), P(wf), P(None), P(y), P(st), B, cin, cout, ndim, size, k, ctypes.c_float(1.0), ops._stream())
            t = time_rot(fs, list(range(nb)), iters=12)
            row += f" | fwd+stats {t*1e6:7.1f} us"
        print(row)
    print("totals: " + ", ".join(f"{k} {v*1e6:.0f} us" for k, v in tot.items()))


def dconv():
    """The discriminator's three Conv2d(k5, s2, p2) on the tcgen05 tap GEMMs (transposed-conv duality, ops.conv5x5_s2):
    forward = hg_convt_dgrad, dx = hg_convt_fwd, dw = hg_convt_wgrad, next to cuDNN bf16 channels-last on the same shapes."""
    import ctypes
    import torch.nn.functional as F
    from lightning_gan_zoo_b200 import _lib
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    peak = peaks["bf16_tflops"]
    P = ops._ptr
    for B, img in ((64, 64), (32, 128)):
        print(f"# D convs at B={B}, {img}x{img}; bf16 peak {peak} TFLOP/s")
        side = img // 4
        for name, cin, cout in (("D blk0", 64, 128), ("D blk1", 128, 256), ("D blk2", 256, 512)):
            size = side                                  # output extent
            flops = 2.0 * B * size * size * cin * cout * 25
            nb = 3
            xs = [torch.randn(B, size, size, 4, cin, device=DEV).to(torch.bfloat16) for _ in range(nb)]
            dys = [torch.randn(B, size, size, cout, device=DEV).to(torch.bfloat16) for _ in range(nb)]
            w = torch.randn(cout, cin, 5, 5, device=DEV) * 0.02
            wf, wd = ops.pack_convt_weight(w)
            y = torch.empty_like(dys[0]); dx = torch.empty_like(xs[0]); dw = torch.empty_like(w)
            nws = _lib.load().hg_convt_wgrad_workspace_bytes(B, cout, cin, 2, size, 5)
            ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=DEV)
            f = lambda i: _lib.call("hg_convt_dgrad", P(xs[i]), P(wd), P(y), B, cout, cin, 2, size, 5, ops._stream())
            g = lambda i: _lib.call("hg_convt_fwd", P(dys[i]), P(wf), P(None), P(dx), B, cout, cin, 2, size, 5, ctypes.c_float(1.0), ops._stream())
            h = lambda i: _lib.call("hg_convt_wgrad", P(dys[i]), P(xs[i]), P(dw), P(ws), nws, B, cout, cin, 2, size, 5, 0, 0, 0, ops._stream())
            # the split-K entry points the discriminator calls (hg_conv5s2_*), next to the unsplit transposed-conv duals
            nw5 = _lib.load().hg_conv5s2_workspace_bytes(B, cin, cout, size)
            ws5 = torch.empty(max(nw5, 16), dtype=torch.uint8, device=DEV)
            f5 = lambda i: _lib.call("hg_conv5s2_fwd", P(xs[i]), P(wd), P(None), P(y), P(ws5), nw5, B, cin, cout, size, ops._stream())
            g5 = lambda i: _lib.call("hg_conv5s2_dx", P(dys[i]), P(wf), P(None), P(dx), P(ws5), nw5, B, cin, cout, size, ops._stream())
            row = f"{name} {cin:3d}->{cout:3d} out {size:2d}^2 {flops/1e9:6.1f} GF"
            for tag, fn in (("fwd[split]", f5), ("dx[split]", g5)):
                t = time_rot(fn, list(range(nb)), iters=12)
                row += f" | {tag} {t*1e6:6.1f} us ({flops/t/1e12/peak*100:4.1f}%)"
            for tag, fn in (("fwd", f), ("dx", g), ("dw", h)):
                if tag == "dw" and nws < 0:
                    row += " | dw unsupported"
                    continue
                t = time_rot(fn, list(range(nb)), iters=12)
                row += f" | {tag} {t*1e6:6.1f} us {flops/t/1e12:5.0f} TF/s ({flops/t/1e12/peak*100:4.1f}%)"
            # cuDNN on the same shapes (channels-last bf16), forward and backward
            xc = [torch.randn(B, cin, 2 * size, 2 * size, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True) for _ in range(nb)]
            wc = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            tf = time_rot(lambda i: F.conv2d(xc[i], wc, None, stride=2, padding=2), list(range(nb)), iters=12)
            dyc = torch.randn(B, cout, size, size, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            def fb(i):
                xc[i].grad = None; wc.grad = None
                F.conv2d(xc[i], wc, None, stride=2, padding=2).backward(dyc)
            tfb = time_rot(fb, list(range(nb)), iters=12, graph=False)
            row += f" | cuDNN fwd {tf*1e6:6.1f} us, fwd+bwd {tfb*1e6:6.1f} us"
            print(row)
            side //= 2


def pipeline():
    """The HBM-bound kernels of the bf16 pipeline at their hot-path shapes (B = 64): channels-last AdaIN on
    s2d conv outputs, channels-last rotate (-> PROJ), final conv + tanh, weight packing, activation backward."""
    import ctypes
    from lightning_gan_zoo_b200 import _lib
    B = 64
    P = ops._ptr
    bf = torch.bfloat16
    print(f"# bf16 pipeline kernels at B={B}; HBM peak {PEAK} GB/s; bytes = algorithmic (DESIGN.md 4.1)")

    def report(name, t, nbytes):
        print(f"{name:44s} {t*1e6:8.1f} us {nbytes/t/1e9:7.0f} GB/s ({nbytes/t/1e9/PEAK*100:5.1f}%)")

    # AdaIN channels-last: (ndim, size, classes, C)  -- block1, block2, block3, block4 outputs; D blocks
    for name, ndim, size, classes, c, biased in [("adain_cl block1 (4^3 x8 -> 8^3, C128)", 3, 4, 8, 128, 0),
                                                 ("adain_cl block2 (8^3 x8 -> 16^3, C64)", 3, 8, 8, 64, 0),
                                                 ("adain_cl block3 (16^2 x4 -> 32^2, C256)", 2, 16, 4, 256, 0),
                                                 ("adain_cl block4 (32^2 x4 -> 64^2, C64)", 2, 32, 4, 64, 0),
                                                 ("inorm_cl D blk0 (16^2, C128)", 2, 16, 1, 128, 1),
                                                 ("inorm_cl D blk1 (8^2, C256)", 2, 8, 1, 256, 1),
                                                 ("inorm_cl D blk2 (4^2, C512)", 2, 4, 1, 512, 1)]:
        n = size ** ndim * classes
        nb = min(12, max(3, int(300e6 // (B * n * c * 2)) + 1))
        xs = [torch.randn(B, n, c, device=DEV).to(bf) for _ in range(nb)]
        dy = torch.randn(B, n, c, device=DEV).to(bf)
        sc = torch.rand(B, c, device=DEV); bi = torch.randn(B, c, device=DEV)
        mean = torch.empty(B, c, device=DEV); rstd = torch.empty(B, c, device=DEV)
        y = torch.empty_like(xs[0]); dx = torch.empty_like(xs[0]); ds = torch.empty(B, c, device=DEV); db = torch.empty(B, c, device=DEV)
        nws = _lib.load().hg_adain_cl_workspace_bytes(B, c, ndim, size, classes)
        wsp = torch.empty(max(nws, 16), dtype=torch.uint8, device=DEV)
        f = lambda x: _lib.call("hg_adain_cl_fwd", P(x), P(sc), P(bi), P(y), P(mean), P(rstd), P(wsp), nws, B, c, ndim, size,
                                classes, c, ctypes.c_float(1e-8), ctypes.c_float(0.0), biased, ops._stream())
        g = lambda x: _lib.call("hg_adain_cl_bwd", P(x), P(dy), P(sc), P(bi), P(mean), P(rstd), P(dx), P(ds), P(db), P(wsp), nws,
                                B, c, ndim, size, classes, c, c, ctypes.c_float(0.0), biased, ops._stream())
        # default dispatch / cluster single-pass kernels forced for the backward too / chunked two-kernel path only
        for tag, env in (("", {}), (" [cluster bwd]", {"ADAIN_CL_CLUSTER_BWD": 1}), (" [chunked]", {"ADAIN_CL_NO_CLUSTER": 1})):
            for k, v in env.items():
                _lib.set_option(k, v)
            tf = time_rot(f, xs); f(xs[0]); tb = time_rot(g, xs)
            for k in env:
                _lib.set_option(k, 0)
            report(name + " fwd" + tag, tf, 2 * B * n * c * 2)
            report(name + " bwd" + tag, tb, 3 * B * n * c * 2)
    # spectral norm of the discriminator's three 5x5 convolutions (one grouped call each way); bytes: fwd reads W three
    # times (W^T u, W v, scale) and writes the bf16 operand, bwd reads dW (bf16) + W twice and writes dW_orig
    ws_ = [torch.randn(co, ci, 5, 5, device=DEV).mul_(0.02).contiguous(memory_format=torch.channels_last).requires_grad_(True)
           for co, ci in ((128, 64), (256, 128), (512, 256))]
    us_ = [torch.randn(w.shape[0], device=DEV) for w in ws_]
    vs_ = [torch.randn(w[0].numel(), device=DEV) for w in ws_]
    nel = sum(w.numel() for w in ws_)
    with torch.no_grad():
        t = time_rot(lambda _: ops.spectral_norm_weights(ws_, us_, vs_, True, bf), [0, 1, 2])
    report("spectral_norm fwd (3 layers, 4.1 M weights)", t, nel * (3 * 4 + 2))
    def sn_fb(_):
        outs = ops.spectral_norm_weights(ws_, us_, vs_, True, bf)
        torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
    t2 = time_rot(sn_fb, [0, 1, 2], graph=False)
    report("spectral_norm fwd+bwd (eager, incl. autograd)", t2, nel * (3 * 4 + 2 + 2 * 2 + 2 * 4 + 4))
    # rotate channels-last -> PROJ
    s, c = 16, 64
    a = ops.view_to_affine(views(B), s, s).to(DEV)
    vols = [torch.randn(B, s, s, s, c, device=DEV).to(bf) for _ in range(10)]
    tf = time_rot(lambda v: ops.rotate_fwd_raw(v, a, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ), vols)
    tb = time_rot(lambda v: ops.rotate_bwd_raw(v, a, c, s, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ), vols)
    report("rotate_cl NDHWC->PROJ (64,16^3,64) fwd", tf, 2 * B * s ** 3 * c * 2)
    report("rotate_cl NDHWC->PROJ (64,16^3,64) bwd", tb, 2 * B * s ** 3 * c * 2)
    del vols
    # final conv + tanh
    xs = [torch.randn(B, 64, 64, 64, device=DEV).to(bf).requires_grad_(True) for _ in range(10)]
    w = (torch.randn(3, 64, 3, 3, device=DEV) * 0.02).requires_grad_(True); bias = torch.zeros(3, device=DEV, requires_grad=True)
    dout = torch.randn(B, 3, 64, 64, device=DEV)
    outs = {}
    def ff(x):
        outs["o"] = ops.final_conv_tanh(x, w, bias)
    tf = time_rot(ff, xs)
    def fb(x):
        o = ops.final_conv_tanh(x, w, bias)
        o.backward(dout)
    tfb = time_rot(fb, xs, graph=False)
    nbytes = B * 64 * 64 * 64 * 2 + B * 3 * 64 * 64 * 4
    report("final_conv_tanh (64,64^2,64->3) fwd", tf, nbytes)
    report("final_conv_tanh fwd+bwd (dx + dw)", tfb, 3 * nbytes)
    # discriminator ends + patched 128 head (disc_ends.cu): bytes = image fp32 + 64-channel map bf16
    for S, Bq in ((64, B), (128, 32)):
        x = [torch.rand(Bq, 3, S, S, device=DEV) * 2 - 1 for _ in range(6)]
        w0 = torch.randn(64, 3, 5, 5, device=DEV) * 0.05; b0 = torch.zeros(64, device=DEV)
        y = torch.empty(Bq, S // 4, S // 4, 4, 64, device=DEV, dtype=bf)
        dy = torch.randn(Bq, S // 4, S // 4, 4, 64, device=DEV).to(bf)
        dxi = torch.empty(Bq, 3, S, S, device=DEV); dw0 = torch.empty_like(w0); db0 = torch.empty(64, device=DEV)
        nb0 = _lib.load().hg_dconv0_bwd_workspace_bytes(Bq, S); ws0 = torch.empty(nb0, dtype=torch.uint8, device=DEV)
        nbytes = Bq * 3 * S * S * 4 + Bq * (S // 2) ** 2 * 64 * 2
        t = time_rot(lambda xi: _lib.call("hg_dconv0_fwd", P(xi), P(w0), P(b0), P(y), Bq, 3, 64, S, ctypes.c_float(0.2), ops._stream()), x)
        report(f"dconv0 fwd (B{Bq}, {S}^2)", t, nbytes)
        t = time_rot(lambda xi: _lib.call("hg_dconv0_bwd", P(xi), P(w0), P(y), P(dy), P(None), P(dw0), P(db0), P(ws0), nb0, Bq, 3, 64, S,
                                          ctypes.c_float(0.2), 0, ops._stream()), x)
        report(f"dconv0 bwd dw (B{Bq}, {S}^2)", t, nbytes + Bq * (S // 2) ** 2 * 64 * 2)
        t = time_rot(lambda xi: _lib.call("hg_dconv0_bwd", P(xi), P(w0), P(y), P(dy), P(dxi), P(None), P(None), P(ws0), nb0, Bq, 3, 64, S,
                                          ctypes.c_float(0.2), 0, ops._stream()), x)
        report(f"dconv0 bwd dx (B{Bq}, {S}^2)", t, nbytes + Bq * (S // 2) ** 2 * 64 * 2)
    hs = [torch.randn(B, 4, 4, 512, device=DEV).to(bf).requires_grad_(True) for _ in range(6)]
    hp = [torch.randn(1, 8192, device=DEV) * 0.02, torch.zeros(1, device=DEV), torch.randn(128, 8192, device=DEV) * 0.02,
          torch.zeros(128, device=DEV), torch.randn(128, 128, device=DEV) * 0.1, torch.zeros(128, device=DEV)]
    hp = [t_.requires_grad_(True) for t_ in hp]
    with torch.no_grad():
        t = time_rot(lambda h_: ops.dheads(h_, *hp, 0.2), hs)
    report("dheads fwd (B64, 8192 features)", t, B * 8192 * 2 + 129 * 8192 * 4)
    def hfb(h_):
        lg, zp = ops.dheads(h_, *hp, 0.2)
        (lg.sum() + zp.sum()).backward()
    t = time_rot(hfb, hs, graph=False)
    report("dheads fwd+bwd (eager, incl. autograd)", t, 3 * (B * 8192 * 2 + 129 * 8192 * 4))
    x128 = [torch.randn(32, 64, 64, 64, device=DEV).to(bf).requires_grad_(True) for _ in range(6)]
    w128 = (torch.randn(64, 3, 4, 4, device=DEV) * 0.05).requires_grad_(True); b128 = torch.zeros(3, device=DEV, requires_grad=True)
    with torch.no_grad():
        t = time_rot(lambda xi: ops.head128_tanh(xi, w128, b128), x128)
    report("head128 fwd (B32, 64^2 x64 -> 3 x 128^2)", t, 32 * 64 * 64 * 64 * 2 + 32 * 3 * 128 * 128 * 4)
    d128 = torch.randn(32, 3, 128, 128, device=DEV)
    def h128fb(xi):
        ops.head128_tanh(xi, w128, b128).backward(d128)
    t = time_rot(h128fb, x128, graph=False)
    report("head128 fwd+bwd (eager, incl. autograd)", t, 3 * (32 * 64 * 64 * 64 * 2 + 32 * 3 * 128 * 128 * 4))
    # weight packing (all five conv layers of G) and activation backward of the projection
    ws = [torch.randn(512, 128, 3, 3, 3, device=DEV), torch.randn(128, 64, 3, 3, 3, device=DEV), torch.randn(1024, 1024, 1, 1, device=DEV),
          torch.randn(1024, 256, 4, 4, device=DEV), torch.randn(256, 64, 4, 4, device=DEV)]
    def pk(_):
        for wt in ws:
            ops.pack_convt_weight(wt)
    t = time_rot(pk, [0, 1, 2])
    report("pack_weight x5 (7.5 M params: 4 B in, 2x2 B out)", t, sum(wt.numel() for wt in ws) * 8)
    ys = [torch.randn(B * 256, 1024, device=DEV).to(bf) for _ in range(6)]
    dyy = torch.randn(B * 256, 1024, device=DEV).to(bf)
    t = time_rot(lambda yv: ops.act_bwd_bias(yv, dyy, 0.0, True), ys)
    report("act_bwd_bias (16384 x 1024)", t, 3 * B * 256 * 1024 * 2)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("rotate", "all"):
        rotate()
    if what == "rotate16":                  # tuning: interleaved-tile kernels with 1 and 2 channel groups per CTA
        from lightning_gan_zoo_b200 import _lib
        for g in (512, 1024):
            print(f"## threads per CTA = {g}")
            rotate([(64, 64, 16), (64, 256, 16), (64, 64, 8)], tune=_lib.HG_TUNE_CTA1024 if g == 1024 else 0)
    if what == "rotate32":                  # BASELINE cfg 3 at 32^3 only (with HG_BENCH_SLAB32=1: the slab forward A/B)
        rotate([(64, 64, 32), (64, 128, 32)])
    if what in ("adain", "all"):
        adain()
    if what in ("conv", "all"):
        conv()
    if what in ("dconv", "all"):
        dconv()
    if what in ("pipeline", "all"):
        pipeline()
