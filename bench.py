#!/usr/bin/env python
"""bench.py -- HoloGAN training-step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 60 --warmup 9
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host cores

One "step" = one `training_step` + optimizer update on one batch, following the reference's
[D, G, G] optimizer schedule (conf/expt/hologan.yaml:16-17).  Workload (BASELINE.json configs[1]):
HoloGAN 64x64, batch 64 per GPU, bf16 compute, synthetic data, random-init weights.

Prints ONE JSON line (rank 0).  `value` times K steps with inputs resident in HBM; `e2e` times the
same K steps through the public API with pinned-host inputs copied H2D and the loss read back D2H
every step.  `roofline` is measured live for the dominant hand-written kernel; `cpu_baseline` times
the oracle port on the host (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "hologan_train_images_per_s"
UNIT = "images/s"
WORKLOAD = "HoloGAN 64x64 full training step ([D,G,G] schedule), batch 64 per GPU, bf16"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=9)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--img-size", type=int, default=64)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--mode", default="train", choices=["train", "sweep"],
                    help="train: the training step (BASELINE configs[1] / [3]); sweep: azimuth-sweep inference (configs[4])")
    ap.add_argument("--views", type=int, default=36, help="sweep mode: views per latent")
    ap.add_argument("--sweep-batch", type=int, default=256, help="sweep mode: latents per GPU")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# measured peaks / clocks
# --------------------------------------------------------------------------------------------------

def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index
        self.path = f"/tmp/hg_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------

def cpu_training_steps(batch: int, img_size: int, steps: int, warmup: int):
    """[D,G,G] training steps of the reference algorithm (oracle/hologan_oracle.py restates
    core/lightning_module.py:209-237 + the two networks) on the host; returns (images/s, ms/step, threads)."""
    from oracle import hologan_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(1)
    gp = {k: v.requires_grad_(True) for k, v in orc.init_generator_params(64, 3, 128, img_size, generator=gen).items()}
    dp = orc.init_discriminator_params(3, 64, 128, img_size, generator=gen)
    dp = {k: (v.requires_grad_(True) if not k.endswith(("_u", "_v")) else v) for k, v in dp.items()}
    opt_d = torch.optim.Adam([v for v in dp.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.999))
    opt_g = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    rs = np.random.RandomState(1)

    def one(i):
        idx = 0 if i % 3 == 0 else 1
        real = torch.rand(batch, 3, img_size, img_size, generator=gen) * 2 - 1
        z = torch.rand(batch, 128, generator=gen) * 2 - 1
        view = orc.sample_view(batch, rs)
        fake = orc.generator_forward(gp, z, view, img_size)
        if idx == 0:
            opt_d.zero_grad(set_to_none=True)
            loss, _ = orc.hologan_losses(0, dp, real, fake, z)
            loss.backward()
            opt_d.step()
        else:
            opt_g.zero_grad(set_to_none=True)
            # Lightning's toggle_optimizer freezes D during the G step: no D weight gradients
            frozen = {k: v.detach() for k, v in dp.items()}
            loss, _ = orc.hologan_losses(1, frozen, None, fake, z)
            loss.backward()
            opt_g.step()
            for k in frozen:
                if k.endswith(("_u", "_v")):
                    dp[k] = frozen[k]
        return float(loss.detach())

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    ips, ms, threads = cpu_training_steps(args.cpu_batch, args.img_size, steps, warmup)
    sample = f"batch {args.cpu_batch} per step (bounded sample of the batch-{args.batch} workload), fp32, torch CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"HoloGAN {args.img_size}x{args.img_size} full training step ([D,G,G] schedule), reference algorithm on the "
                               f"host CPU: batch {args.cpu_batch} per step, fp32 (bounded sample of the batch-{args.batch}-per-GPU bf16 workload)",
                   "img_size": args.img_size, "per_gpu_batch": args.batch, "cpu_batch": args.cpu_batch},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# roofline of the dominant hand-written kernel, measured live with CUDA events
# --------------------------------------------------------------------------------------------------

def _event_time_ms(fn, nbuf, iters):
    """Average milliseconds per call of fn(i) on torch's current stream (the stream the C ABI launches on)."""
    for i in range(min(nbuf, 3)):
        fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(iters):
        fn(i % nbuf)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def _ncu_traffic(kernel_key):
    """dram bytes per launch from the committed `ncu --set full` summary (profiles/ncu_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(kernel_key, {}).get("dram_bytes_per_launch")
    except (ValueError, OSError):
        return None


def conv_roofline(peaks, device):
    """Dominant kernel of the step: hg::tap_gemm_kernel (18 % of the step's kernel time, profiles/*_launches_summary.txt),
    timed at its largest instance -- block3 forward, ConvTranspose2d 1024->256 k4 s2 at 16x16, B = 64.
    Algorithmic FLOPs per launch = 2147.5 MFLOP/sample x 64 = 137.4 GFLOP.  Tensor-pipe roofline
    against the measured cuBLAS bf16 burst peak (the kernel is timed alone).  Three rotating operand sets."""
    import ctypes
    from lightning_gan_zoo_b200 import _lib, ops
    B, cin, cout, size, k, ndim = 64, 1024, 256, 16, 4, 2
    nb = 3
    xs = [torch.randn(B, size, size, cin, device=device).to(torch.bfloat16) for _ in range(nb)]
    w = torch.randn(cin, cout, k, k, device=device) * 0.02
    wf, _ = ops.pack_convt_weight(w)
    y = torch.empty(B, size, size, 4, cout, device=device, dtype=torch.bfloat16)
    P = ops._ptr
    fn = lambda i: _lib.call("hg_convt_fwd", P(xs[i]), P(wf), P(None), P(y), B, cin, cout, ndim, size, k,
                             ctypes.c_float(1.0), ops._stream())
    ms = _event_time_ms(fn, nb, 40)
    flops = 2.0 * B * size * size * cin * cout * k * k
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "hg::tap_gemm_kernel, block3 fwd (ConvT2d 1024->256 k4 s2 @16x16, B=64, bf16)",
            "achieved": achieved, "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst)",
            "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"], "traffic": _ncu_traffic("tap_gemm_block3_fwd"),
            "us_per_launch": ms * 1e3, "algorithmic_flops": flops}


def conv_roofline_flop_weighted(peaks, device):
    """All 15 dense contractions of one generator forward + backward at B = 64 (5 layers x forward / dgrad / wgrad on
    hg::tap_gemm_kernel / hg::wgrad_gemm_kernel), each timed alone with CUDA events: total algorithmic FLOPs / total time.
    The single best instance (`roofline`) must not be mistaken for the step: this is the FLOP-weighted figure."""
    import ctypes
    from lightning_gan_zoo_b200 import _lib, ops
    B, P = 64, ops._ptr
    layers = [("block1", 3, 3, 512, 128, 4), ("block2", 3, 3, 128, 64, 8), ("proj1x1", 2, 1, 1024, 1024, 16),
              ("block3", 2, 4, 1024, 256, 16), ("block4", 2, 4, 256, 64, 32)]
    tot_flops, tot_s, per = 0.0, 0.0, {}
    for name, ndim, k, cin, cout, size in layers:
        sp = (size,) * ndim
        ncls = 1 if k == 1 else 2 ** ndim
        flops = 2.0 * B * size ** ndim * cin * cout * k ** ndim
        xs = [torch.randn(B, *sp, cin, device=device).to(torch.bfloat16) for _ in range(3)]
        dys = [torch.randn(B, *sp, ncls, cout, device=device).to(torch.bfloat16) for _ in range(3)]
        w = torch.randn(cin, cout, *((k,) * ndim), device=device) * 0.02
        wf, wd = ops.pack_convt_weight(w)
        y, dx, dw = torch.empty_like(dys[0]), torch.empty_like(xs[0]), torch.empty_like(w)
        nws = _lib.load().hg_convt_wgrad_workspace_bytes(B, cin, cout, ndim, size, k)
        ws = torch.empty(nws, dtype=torch.uint8, device=device)
        st = ops._stream
        passes = {
            "fwd": lambda i: _lib.call("hg_convt_fwd", P(xs[i]), P(wf), P(None), P(y), B, cin, cout, ndim, size, k, ctypes.c_float(1.0), st()),
            "dgrad": lambda i: _lib.call("hg_convt_dgrad", P(dys[i]), P(wd), P(dx), B, cin, cout, ndim, size, k, st()),
            "wgrad": lambda i: _lib.call("hg_convt_wgrad", P(xs[i]), P(dys[i]), P(dw), P(ws), nws, B, cin, cout, ndim, size, k, 0, 0, 0, st()),
        }
        for tag, fn in passes.items():
            sec = _event_time_ms(fn, 3, 20) * 1e-3
            tot_flops += flops
            tot_s += sec
            per[f"{name}.{tag}"] = round(flops / sec / 1e12, 1)
    achieved = tot_flops / tot_s / 1e12
    return {"bound": "tensor", "kernel": "all 15 generator contractions (5 layers x fwd / dgrad / wgrad), FLOP-weighted, B=64, bf16",
            "achieved": achieved, "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst)", "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "traffic": None, "us_total": tot_s * 1e6, "algorithmic_flops": tot_flops,
            "tflops_per_kernel": per}


def rotate_roofline(peaks, device):
    """cfg 3 microbench ("voxel-rotate HBM GB/s" of the metric): (64, 64, 16^3) fp32 rotate-resample, forward and
    backward.  Algorithmic bytes per launch = read one volume + write one volume = 2*B*C*S^3*4 = 128 MiB
    (DESIGN.md 4.1).  Eight rotating 64 MiB inputs (512 MiB) so every launch reads from HBM, not the 126 MB L2."""
    from lightning_gan_zoo_b200 import ops
    b, c, s = 64, 64, 16
    nbuf = 8
    vols = [torch.randn(b, c, s, s, s, device=device) for _ in range(nbuf)]
    rs = np.random.RandomState(0)
    view = np.zeros((b, 6)); view[:, 0] = np.deg2rad(rs.randint(220, 320, b)); view[:, 1] = np.deg2rad(rs.randint(70, 110, b)); view[:, 2] = 1
    a = ops.view_to_affine(view).to(device)
    ms_f = _event_time_ms(lambda i: ops.rotate_fwd_raw(vols[i], a, ops.HG_BORDER_REFERENCE), nbuf, 40)
    ms_b = _event_time_ms(lambda i: ops.rotate_bwd_raw(vols[i], a, c, s, ops.HG_BORDER_REFERENCE), nbuf, 40)
    bytes_alg = 2 * b * c * s ** 3 * 4
    out = {}
    for tag, ms, key, kern in (("fwd", ms_f, "rotate_il_fwd", "hg::rotate_fwd_il_kernel<float,4,512,false>"),
                               ("bwd", ms_b, "rotate_il_bwd", "hg::rotate_adjoint_table_kernel<4> + hg::rotate_bwd_il_kernel<float,4,512>")):
        achieved = bytes_alg / (ms * 1e-3) / 1e9
        out[tag] = {"bound": "hbm", "kernel": kern + " (64,64,16^3) fp32", "achieved": achieved, "peak": peaks["hbm_gbs"],
                    "peak_source": peaks["source"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": _ncu_traffic(key), "us_per_launch": ms * 1e3, "algorithmic_bytes": bytes_alg}
    out["l2_policy"] = "8 rotating 64 MiB inputs (512 MiB) > 126 MB L2"
    del vols
    # cfg 3 at 32^3 (slab forward, table-free gather backward) and the bf16-pipeline kernel the training step runs
    extra = []
    s32 = 32
    a32 = ops.view_to_affine(view, s32, s32).to(device)
    v32 = [torch.randn(b, c, s32, s32, s32, device=device) for _ in range(3)]          # 3 x 512 MiB
    extra.append(("fwd_32", "hg::rotate_fwd_slab32_kernel<float> (64,64,32^3) fp32", 2 * b * c * s32 ** 3 * 4,
                  _event_time_ms(lambda i: ops.rotate_fwd_raw(v32[i], a32, ops.HG_BORDER_REFERENCE), 3, 10)))
    extra.append(("bwd_32", "hg::rotate_bwd_gather_kernel<float> (64,64,32^3) fp32", 2 * b * c * s32 ** 3 * 4,
                  _event_time_ms(lambda i: ops.rotate_bwd_raw(v32[i], a32, c, s32, ops.HG_BORDER_REFERENCE), 3, 10)))
    del v32
    vcl = [torch.randn(b, s, s, s, c, device=device).to(torch.bfloat16) for _ in range(nbuf)]
    extra.append(("fwd_cl", "hg::rotate_cl_fwd_kernel NDHWC->PROJ (64,16^3,64) bf16 (the training step's)", 2 * b * c * s ** 3 * 2,
                  _event_time_ms(lambda i: ops.rotate_fwd_raw(vcl[i], a, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ), nbuf, 40)))
    extra.append(("bwd_cl", "hg::rotate_adjoint_table_kernel + hg::rotate_cl_bwd_ell_kernel (64,16^3,64) bf16", 2 * b * c * s ** 3 * 2,
                  _event_time_ms(lambda i: ops.rotate_bwd_raw(vcl[i], a, c, s, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ), nbuf, 40)))
    for tag, kern, nbytes, ms in extra:
        achieved = nbytes / (ms * 1e-3) / 1e9
        out[tag] = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"],
                    "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None, "us_per_launch": ms * 1e3,
                    "algorithmic_bytes": nbytes}
    return out


# --------------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------------

def run_sweep(args):
    """BASELINE.json configs[4] (SURVEY 8d cfg 5): HoloGAN azimuth-sweep inference, `--sweep-batch` latents per GPU x
    `--views` azimuths (linspace(220, 320, V) degrees at elevation 90, the pattern of core/figures/types.py:300-322),
    no grad, bf16, patched 128 x 128 head by default (--img-size).  One "step" = one whole sweep per GPU.  Latents are
    sharded over the ranks; there is no collective on the data path.  `value`: latents resident in HBM; `e2e`: latents
    from pinned host memory, every image copied back D2H (on a copy stream, overlapped with the next views)."""
    import torch.distributed as dist
    from types import SimpleNamespace
    from lightning_gan_zoo_b200 import _lib
    from lightning_gan_zoo_b200.core.models.hologan_generator import Generator
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()
    B, V, S = args.sweep_batch, args.views, args.img_size if args.img_size != 64 or "--img-size" in sys.argv else 128
    K, W = args.steps, max(args.warmup, 3)
    torch.manual_seed(42)
    net = Generator(64, 3, 128, SimpleNamespace(), S).to(device).eval()
    views = np.zeros((V, 6))
    views[:, 0] = np.deg2rad(np.linspace(220, 320, V))
    views[:, 1] = np.deg2rad(90.0)
    views[:, 2] = 1.0
    gen = torch.Generator().manual_seed(7 + rank)
    z_host = (torch.rand(B, 128, generator=gen) * 2 - 1).pin_memory()
    z_dev = z_host.to(device)
    out_host = torch.empty((B, V, 3, S, S), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident(_i):
        with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
            net.render_views(z_dev, views)

    def e2e(_i):
        with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
            z = z_host.to(device, non_blocking=True)
            out = net.render_views(z, views)
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                out_host.copy_(out, non_blocking=True)
            out.record_stream(copy_stream)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        torch.cuda.current_stream().wait_stream(copy_stream)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(W):
        resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms_total = timed(resident, K)
    launches = _lib.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    e2e(0)
    torch.cuda.synchronize()
    ms_e2e = timed(e2e, K)
    imgs = world * B * V * K
    line = {
        "metric": "hologan_sweep_images_per_s", "value": imgs / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"HoloGAN {S}x{S} azimuth sweep inference, {V} views per z, batch {B} latents per GPU, bf16"
                               + (" (patched 128 head, SURVEY R4)" if S == 128 else ""),
                   "img_size": S, "views": V, "per_gpu_latents": B, "parallelism": f"dp{world} (latents sharded, no collective)",
                   "weights": "random init",
                   "l2": "no flush: one sweep writes %.1f GB of images per GPU" % (B * V * 3 * S * S * 4 / 1e9)},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 128 * 4,
                "d2h_bytes_per_step": B * V * 3 * S * S * 4, "ms_per_step": ms_e2e / K},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        _exit_without_nccl_teardown(rank)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.mode == "sweep":
        run_sweep(args)
        return
    import torch.distributed as dist
    from lightning_gan_zoo_b200 import _lib, ops
    from lightning_gan_zoo_b200.training import HologanConfig, HologanTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    cfg = HologanConfig(batch_size=args.batch, img_size=args.img_size)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    trainer = HologanTrainer(cfg, device=device, compute_dtype=dtype, rank=rank, world_size=world)
    B, S = args.batch, args.img_size
    K, W = args.steps, args.warmup

    # ---- synthetic inputs -------------------------------------------------------------------------
    pool = 6
    gen = torch.Generator().manual_seed(100 + rank)
    real_host = [torch.rand(B, 3, S, S, generator=gen).mul_(2).sub_(1).pin_memory() for _ in range(pool)]
    real_dev = [r.to(device) for r in real_host]
    z_dev = [trainer.sample_noise(B).to(device) for _ in range(pool)]
    a_dev = [ops.view_to_affine(trainer.sample_view(B)).to(device) for _ in range(pool)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(i):
        trainer.step(real_dev[i % pool], i, z=z_dev[i % pool], view=a_dev[i % pool])

    pending = []

    def e2e_step(i):
        # public API with HOST buffers (HologanTrainer.step_host): pinned real images, latents and views sampled on
        # the host (like the reference, lightning_module.py:212), copied H2D inside the timed region; the loss of
        # EVERY step is read back D2H -- one step late, after the next step has been launched, so the read does not
        # drain the GPU (e2e_finish reads the last one, still inside the timed region)
        pend = trainer.step_host(real_host[i % pool], i, z=trainer.sample_noise(B), view=trainer.sample_view(B))
        if pending:
            pending.pop().item()          # D2H read of the previous step's result
        pending.append(pend)

    def e2e_finish():
        while pending:
            pending.pop().item()

    def timed(fn, k, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    if not args.no_graphs:
        trainer.enable_cuda_graphs(B)
    for i in range(max(W, 3)):
        resident_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms_total = timed(resident_step, K)
    launches = _lib.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    for i in range(3):
        e2e_step(i)
    e2e_finish()
    ms_e2e = timed(e2e_step, K, e2e_finish)

    value = world * B * K / (ms_total * 1e-3)
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    h2d = B * 3 * S * S * 4 + B * cfg.noise_dim * 4 + B * 16 * 4

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD if (B, S, args.dtype) == (64, 64, "bf16") else
                   f"HoloGAN {S}x{S} training step, batch {B} per GPU, {args.dtype}",
                   "img_size": S, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "schedule": "[D,G,G]", "weights": "random init", "cuda_graphs": not args.no_graphs,
                   "l2": "no flush: one step touches >126 MB of distinct activations/gradients/weights+Adam state"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if rank == 0:
        peaks = measured_peaks()
        if not args.no_roofline:
            line["roofline"] = conv_roofline(peaks, device)
            line["roofline_flop_weighted"] = conv_roofline_flop_weighted(peaks, device)
            line["roofline_rotate"] = rotate_roofline(peaks, device)
        if world == 1 and not args.no_cpu_baseline:
            ips, ms, threads = cpu_training_steps(args.cpu_batch, S, steps=6, warmup=3)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"6 [D,G,G] steps at batch {args.cpu_batch} (fp32, torch CPU oracle port), "
                                              f"{ms:.0f} ms/step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        _exit_without_nccl_teardown(rank)


def _exit_without_nccl_teardown(rank: int):
    """Multi-rank runs end with a hard exit.  Measured on 2 x B200 (profiles/r01h_bench_n2.*): after the JSON line
    was printed, `dist.barrier(); dist.destroy_process_group()` never returned -- the step's CUDA graphs hold captured
    NCCL all-reduces, and tearing the communicator down under live graphs blocks.  Nothing after the timed region needs
    NCCL (the max-over-ranks all-reduce already happened): the other ranks wait on the rendezvous store (TCP, not NCCL)
    until rank 0 has printed its line, then every rank flushes and leaves with exit code 0."""
    import sys
    import time
    from datetime import timedelta
    import torch.distributed as dist
    try:
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("hg_bench_done", "1")
            time.sleep(0.5)                      # let the store answer the waiters before the server thread dies
        else:
            store.wait(["hg_bench_done"], timedelta(seconds=900))
    except Exception:                            # a missing store / closed connection must not turn into a hang or a failure
        pass
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
