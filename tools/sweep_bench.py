#!/usr/bin/env python
"""View-sweep inference throughput (BASELINE.json configs[4] / SURVEY 8d cfg 5) on ONE GPU:
B latents x V azimuths (linspace(220, 320, V) degrees at elevation 90, cf. core/figures/types.py:333-339), no grad,
bf16, images/s = B * V / time.  Two ways:
  per_view  -- the reference's pattern: `generator(z, view_in=view)` once per view (3D trunk recomputed V times)
  sweep     -- `Generator.render_views(z, views)`: trunk once per z, rotate + 2D decoder per view
Usage: python tools/sweep_bench.py [--batch 256] [--views 36] [--img-size 128] [--iters 5]
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from lightning_gan_zoo_b200.core.models.hologan_generator import Generator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--views", type=int, default=36)
    ap.add_argument("--img-size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = Generator(64, 3, 128, SimpleNamespace(), args.img_size).to(dev).eval()
    z = torch.rand(args.batch, 128, device=dev) * 2 - 1
    views = np.zeros((args.views, 6))
    views[:, 0] = np.deg2rad(np.linspace(220, 320, args.views))
    views[:, 1] = np.deg2rad(90.0)
    views[:, 2] = 1.0

    def per_view():
        out = None
        for v in range(args.views):
            out = net(z, view_in=np.repeat(views[v:v + 1], args.batch, axis=0))
        return out

    def sweep():
        return net.render_views(z, views)

    res = {}
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        for name, fn in (("per_view", per_view), ("sweep", sweep)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            res[name] = {"ms_per_sweep": ms, "images_per_s": args.batch * args.views / ms * 1e3}
    print(json.dumps({"metric": "hologan_sweep_images_per_s", "n_gpus": 1, "dtype": "bf16",
                      "config": {"workload": f"azimuth sweep, {args.batch} latents x {args.views} views, "
                                             f"{args.img_size}x{args.img_size}", "iters": args.iters}, **res}))


if __name__ == "__main__":
    main()
