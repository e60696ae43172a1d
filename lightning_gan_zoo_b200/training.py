"""Lightning-free HoloGAN training harness (data-parallel, one process per GPU).

Mirrors what the reference gets from `core.lightning_module.HOLOGAN` + `pl.Trainer(accelerator="ddp")`
(reference core/lightning_module.py:35-102, 209-237; run_network.py:66-72; conf/expt/hologan.yaml):

  * `training_step(batch, batch_idx, optimizer_idx)` -- the two losses of :217-237;
  * two Adam optimizers (lr 1e-4, betas (0.9, 0.999)) alternating with frequencies
    disc_freq = 1 / gen_freq = 2, i.e. the batch schedule [D, G, G, D, G, G, ...] (:75-87);
  * the LambdaLR schedule of core/utils/hologan.py:3-9;
  * DDP semantics: gradients are averaged over ranks once per optimizer step.  Here each network owns
    ONE flat gradient buffer (parameters' .grad are views into it), so the exchange is a single NCCL
    all-reduce per step (payload 21.5 MB for D, 31.2 MB for G in fp32) instead of DDP's buckets.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops
from .core.models.hologan_discriminator import Discriminator
from .core.models.hologan_generator import Generator


@dataclass
class HologanConfig:
    """Effective values of `+expt=hologan` (conf/expt/hologan.yaml:1-58, conf/config.yaml)."""
    noise_dim: int = 128
    batch_size: int = 32
    img_size: int = 64
    channels_img: int = 3
    num_epochs: int = 25
    lr: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    disc_freq: int = 1
    gen_freq: int = 2
    gen_in_planes: int = 64
    disc_out_planes: int = 64
    view_args: SimpleNamespace = field(default_factory=lambda: SimpleNamespace(
        elevation_low=70, elevation_high=110, azimuth_low=220, azimuth_high=320, scale_low=1, scale_high=1,
        transX_low=0, transX_high=0, transY_low=0, transY_high=0, transZ_low=0, transZ_high=0, batch_size=32))


def hologan_lr_lambda(total_epochs: int):
    """x1 until the middle epoch, then linear decay to 0 (core/utils/hologan.py:3-9)."""
    half = total_epochs / 2

    def factor(epoch: int) -> float:
        return 1.0 if epoch <= half else 1.0 - (epoch - half) / half
    return factor


def optimizer_index(batch_idx: int, disc_freq: int = 1, gen_freq: int = 2) -> int:
    """Lightning's `frequency` alternation: disc_freq batches on optimizer 0, then gen_freq on 1."""
    return 0 if (batch_idx % (disc_freq + gen_freq)) < disc_freq else 1


class _FlatGrads:
    """One contiguous gradient buffer per network; parameter .grad tensors are views into it.

    No per-step zero-fill and no per-parameter accumulation kernels: before a backward pass `begin()` detaches the
    views, so autograd hands every parameter its freshly computed gradient (AccumulateGrad keeps the incoming tensor
    instead of launching `grad += new`); `finish()` moves them into the flat buffer with ONE multi-tensor copy and
    re-attaches the views.  The tcgen05 wgrad kernels store straight into their view (`_hg_direct_grad`).  A view
    that no backward ever writes (a parameter without gradient) keeps its initial zeros.

    Data parallel: the buffer is laid out as `buckets` -- groups of parameters in the order their gradients become
    final during the backward pass -- followed by the remaining parameters.  `fire(i)` starts the (asynchronous) sum
    all-reduce of bucket i as soon as its gradients are in the buffer, so it overlaps the rest of the backward; the
    tail goes last and `wait()` joins everything before the optimizer reads the buffer (DDP's bucketed overlap,
    run_network.py:66-71, on a flat buffer)."""

    def __init__(self, params, direct: bool = False, accumulate_into=(), buckets=()):
        """direct: every parameter's view is offered to the kernels as a store target (one backward per step).
        accumulate_into: parameters whose kernels ADD into the view (several backward calls per step, e.g. the
        discriminator's spectral-norm weights: D runs on real and fake) -- those views are zeroed by `begin()`.
        buckets: sequence of parameter lists (see the class docstring)."""
        params = [p for p in params]
        pad = lambda n: (n + 3) // 4 * 4                  # every parameter starts 16-byte aligned (float4 / TMA-free kernels)
        ordered, seen, bounds = [], set(), []
        for bucket in buckets:
            start = sum(pad(p.numel()) for p in ordered)
            for p in bucket:
                if id(p) not in seen:
                    ordered.append(p)
                    seen.add(id(p))
            bounds.append((start, sum(pad(p.numel()) for p in ordered)))
        ordered += [p for p in params if id(p) not in seen]
        self.params = ordered
        acc_ids = {id(p) for p in accumulate_into}
        self.zero_views = []
        self.numel = sum(p.numel() for p in self.params)
        total = sum(pad(p.numel()) for p in self.params)
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(total, device=dev, dtype=dt)     # padding elements stay zero
        self.bucket_slices = bounds + [(bounds[-1][1] if bounds else 0, total)]
        self.views, self.offsets = [], []
        o = 0
        for p in self.params:
            chunk = self.flat[o:o + p.numel()]
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                n_, c, h, w = p.shape                     # same strides as the parameter (fused Adam needs that)
                view = chunk.view(n_, h, w, c).permute(0, 3, 1, 2)
            else:
                view = chunk.view_as(p)
            p.grad = view
            if direct or id(p) in acc_ids:
                p._hg_direct_grad = view                  # ops._direct_grad_target: wgrad kernels store / add here
            if id(p) in acc_ids:
                self.zero_views.append(view)
            self.views.append(view)
            self.offsets.append(o)
            o += pad(p.numel())
        self._handles, self._fired, self._counts = [], set(), {}
        self.bucket_update = None           # see fire()
        self._side, self._side_used = None, False

    def zero(self):
        self.flat.zero_()

    def begin(self):
        for p in self.params:
            p.grad = None
        if self.zero_views:
            with torch.no_grad():
                torch._foreach_zero_(self.zero_views)
        self._handles, self._fired, self._counts = [], set(), {}

    def finish(self):
        if self.flat.is_cuda:
            ops.join_side_stream(self.flat.device)      # weight-gradient GEMMs that ran beside the backward
        src, dst = [], []
        for p, view in zip(self.params, self.views):
            g = p.grad
            if g is not None and g.data_ptr() != view.data_ptr():
                src.append(g.detach())
                dst.append(view)
            p.grad = view
        if src:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)

    # ---- data-parallel exchange ---------------------------------------------------------------------
    def fire(self, i: int, world: int, after_calls: int = 1):
        """Bucket i's gradients are in the buffer (the hook the kernels' backward call; effective once it has been called
        `after_calls` times in the current step).  Without `bucket_update`: start the bucket's asynchronous sum all-reduce.
        With `bucket_update` (a callable (start, end), set by the trainer to FlatAdam.apply_range): all-reduce AND the
        optimizer update of the bucket run on a side stream, stream-ordered behind the gradients, overlapping the rest of
        the backward -- also on one GPU, where only the update is left."""
        self._counts[i] = self._counts.get(i, 0) + 1
        if i in self._fired or self._counts[i] < after_calls:
            return
        s, e = self.bucket_slices[i]
        if self.bucket_update is not None:
            self._fired.add(i)
            if e > s:
                cur = torch.cuda.current_stream(self.flat.device)
                if self._side is None:
                    self._side = ops.side_stream(self.flat.device, "comm")
                self._side.wait_stream(cur)
                ops.join_side_stream(self.flat.device, "wgrad", into=self._side)    # a side-stream wgrad may have produced this bucket
                with torch.cuda.stream(self._side):
                    if world > 1:
                        dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM)      # stream-ordered on the side stream
                    self.bucket_update(s, e)
                self._side_used = True
            return
        if world <= 1:
            return
        self._fired.add(i)
        if e > s:
            ops.join_side_stream(self.flat.device)      # a side-stream wgrad may have produced this bucket
            self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))

    def flush(self, world: int):
        """`bucket_update` mode: exchange + update every bucket `fire` has not handled yet (the tail), then join the side
        stream: afterwards all parameters of the network are updated on the current stream."""
        for i in range(len(self.bucket_slices)):
            self.fire(i, world, after_calls=0)
        if self._side_used:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
            self._side_used = False

    def all_reduce_sum(self, world: int):
        """All-reduce (sum) whatever `fire` has not started yet, then wait for every bucket."""
        if world <= 1:
            return
        for i in range(len(self.bucket_slices)):
            self.fire(i, world, after_calls=0)
        for h in self._handles:
            h.wait()
        self._handles = []

    def all_reduce_mean(self, world: int):
        if world > 1:
            self.all_reduce_sum(world)
            self.flat.div_(world)


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps) (core/lightning_module.py:75-87) on one flat buffer: the parameters are
    re-pointed to views of `flat_param` (in the order of `grads.params`), the moments are flat too, and `step()` is a
    single fused kernel (hg_adam_step) that also applies the data-parallel 1 / world gradient scale.  The per-parameter
    `state` exposes views of the flat moments plus a shared step counter, so `state_dict()` has torch Adam's layout
    (checkpoints stay interchangeable); `load_state_dict` copies a torch Adam state back into the flat buffers."""

    def __init__(self, grads: _FlatGrads, lr, betas=(0.9, 0.999), eps: float = 1e-8):
        params = grads.params
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        dev = grads.flat.device
        self.grads = grads
        self.flat_param = torch.zeros_like(grads.flat)
        self.exp_avg = torch.zeros_like(grads.flat)
        self.exp_avg_sq = torch.zeros_like(grads.flat)
        self.kstate = torch.zeros(4, device=dev)            # [0] steps taken, [1] lr / bc1, [2] 1 / sqrt(bc2)
        lr_t = self.param_groups[0]["lr"]
        if not isinstance(lr_t, torch.Tensor):
            self.param_groups[0]["lr"] = torch.tensor(float(lr_t), device=dev)
        with torch.no_grad():
            for p, o in zip(params, grads.offsets):       # same (16-byte aligned) layout as the gradient buffer
                n = p.numel()
                view = self.flat_param[o:o + n].view_as(p)
                view.copy_(p)
                p.data = view
                self.state[p] = {"step": self.kstate[0], "exp_avg": self.exp_avg[o:o + n].view_as(p),
                                 "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p)}

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        P = ops._ptr
        _lib.call("hg_adam_step", P(self.flat_param), P(self.grads.flat), P(self.exp_avg), P(self.exp_avg_sq),
                  self.flat_param.numel(), P(self.kstate), P(g["lr"]), float(b1), float(b2), float(g["eps"]), float(grad_scale),
                  ops._stream())

    # the same step bucket by bucket: tick() once (before the backward), apply_range() per gradient bucket
    @torch.no_grad()
    def tick(self):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        _lib.call("hg_adam_tick", ops._ptr(self.kstate), ops._ptr(g["lr"]), float(b1), float(b2), ops._stream())

    @torch.no_grad()
    def apply_range(self, start: int, end: int, grad_scale: float = 1.0):
        if end <= start:
            return
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        P = ops._ptr
        _lib.call("hg_adam_apply", P(self.flat_param[start:end]), P(self.grads.flat[start:end]), P(self.exp_avg[start:end]),
                  P(self.exp_avg_sq[start:end]), end - start, P(self.kstate), float(b1), float(b2), float(g["eps"]),
                  float(grad_scale), ops._stream())

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam (or FlatAdam) state dict over the same parameters in the same order."""
        packed = state_dict["state"]
        ids = state_dict["param_groups"][0]["params"]
        with torch.no_grad():
            step = None
            for idx, p in zip(ids, self.param_groups[0]["params"]):
                st = packed.get(idx)
                if st is None:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                step = st["step"]
            if step is not None:
                self.kstate[0] = float(step)
            lr = state_dict["param_groups"][0]["lr"]
            self.param_groups[0]["lr"].fill_(float(lr))


class PendingLoss:
    """Loss of a step launched by `HologanTrainer.step_host`: the value lands in pinned host memory through an
    asynchronous device-to-host copy; `item()` waits for that copy only."""

    def __init__(self, buf: torch.Tensor, event):
        self._buf, self._event = buf, event

    def item(self) -> float:
        self._event.synchronize()
        return float(self._buf[0])

    __float__ = item


class _FeedSlot:
    """Pinned host staging + device staging buffers of one in-flight step."""

    def __init__(self, batch, channels, size, noise_dim, device):
        self.real_pin = torch.empty(batch, channels, size, size).pin_memory()
        self.z_pin = torch.empty(batch, noise_dim).pin_memory()
        self.a_pin = torch.empty(batch, 4, 4).pin_memory()
        self.loss_pin = torch.zeros(1).pin_memory()
        self.real_dev = torch.empty(batch, channels, size, size, device=device)
        self.z_dev = torch.empty(batch, noise_dim, device=device)
        self.a_dev = torch.empty(batch, 4, 4, device=device)
        self.ready = torch.cuda.Event()      # H2D copies of this slot finished (recorded on the copy stream)
        self.done = torch.cuda.Event()       # the step consumed the device staging buffers and wrote the loss
        self.done.record()


class HologanTrainer:
    def __init__(self, cfg: Optional[HologanConfig] = None, device="cuda", compute_dtype=torch.bfloat16,
                 rank: int = 0, world_size: int = 1, seed: int = 42):
        self.cfg = cfg = cfg or HologanConfig()
        self.device = torch.device(device)
        self.rank, self.world = rank, world_size
        self.compute_dtype = compute_dtype
        torch.manual_seed(seed)          # same init on every rank (reference: seed_everything(42), run_network.py:27)
        self.generator = Generator(cfg.gen_in_planes, cfg.channels_img, cfg.noise_dim, cfg.view_args, cfg.img_size,
                                   gpu=self.device.type == "cuda").to(self.device)
        self.discriminator = Discriminator(cfg.channels_img, cfg.disc_out_planes, cfg.noise_dim,
                                           img_size=cfg.img_size).to(self.device)
        if self.world > 1:               # identical replicas even if a rank's RNG had diverged
            for t in list(self.generator.state_dict().values()) + list(self.discriminator.state_dict().values()):
                dist.broadcast(t, src=0)
        if self.device.type == "cuda" and compute_dtype == torch.bfloat16 and os.environ.get("HG_D_LIBRARY", "0") not in ("", "0"):
            # A/B switch HG_D_LIBRARY=1: the discriminator's convolutions on cuDNN run NHWC (its native tensor-core
            # layout), so the weights live channels-last too; the default (hand-written kernels) keeps torch's layout
            self.discriminator.to(memory_format=torch.channels_last)
        cuda = self.device.type == "cuda"
        sn_weights = [blk.conv2d.weight_orig for blk in self.discriminator.blocks] if cuda else ()
        # gradient buckets in the order the backward pass finishes them (data parallel: each is all-reduced as soon as
        # it is final, overlapping the rest of the backward; world 1: only the buffer layout)
        gen, disc = self.generator, self.discriminator
        g_buckets = [[gen.block3.convTranspose.weight, gen.block4.convTranspose.weight], [gen.convTranspose2d1.weight],
                     [gen.block2.convTranspose.weight], [gen.block1.convTranspose.weight]] if cuda else ()
        d_buckets = [[disc.blocks[2].conv2d.weight_orig], [disc.blocks[1].conv2d.weight_orig]] if cuda else ()
        self.d_grads = _FlatGrads(self.discriminator.parameters(), accumulate_into=sn_weights, buckets=d_buckets)
        # the generator's tcgen05 wgrad kernels store straight into the flat buffer (ops._direct_grad_target)
        self.g_grads = _FlatGrads(self.generator.parameters(), direct=cuda, buckets=g_buckets)
        # A/B switch: HG_NO_SN_PREFETCH=1 keeps the discriminator's spectral-norm iteration in line (in front of its first block)
        self._sn_prefetch = cuda and os.environ.get("HG_NO_SN_PREFETCH", "0") in ("", "0") and os.environ.get("HG_D_LIBRARY", "0") in ("", "0")
        self._flat_adam = cuda and os.environ.get("HG_D_LIBRARY", "0") in ("", "0") and os.environ.get("HG_TORCH_ADAM", "0") in ("", "0")
        if self._flat_adam:
            # one fused kernel per optimizer step over flat parameter / gradient / moment buffers (hg_adam_step)
            self.opt_d = FlatAdam(self.d_grads, cfg.lr, betas=(cfg.beta1, cfg.beta2))
            self.opt_g = FlatAdam(self.g_grads, cfg.lr, betas=(cfg.beta1, cfg.beta2))
        else:
            # torch's fused multi-tensor Adam; capturable (device-side step counter, tensor lr) so that a whole
            # optimizer step can live inside a CUDA graph
            lr = torch.tensor(cfg.lr, device=self.device) if cuda else cfg.lr
            kw = dict(betas=(cfg.beta1, cfg.beta2), fused=cuda, capturable=cuda)
            self.opt_d = torch.optim.Adam(self.d_grads.params, lr=lr, **kw)
            self.opt_g = torch.optim.Adam(self.g_grads.params, lr=lr.clone() if cuda else lr, **kw)
        # Bucket hooks: the all-reduce (world > 1) and, with the flat Adam, the optimizer update of a gradient bucket run on a
        # side stream as soon as the bucket is final, overlapping the rest of the backward.  A/B switches:
        # HG_NO_GRAD_OVERLAP=1 (one exchange + one update after the backward), HG_ADAM_OVERLAP=0 / 1 (update behind the
        # buckets off / on; default: on for world > 1 only -- on one GPU the HBM-bound update running beside the backward
        # kernels measured 1.1 % slower than one update at the end, profiles/r02H_adam_overlap.txt).
        overlap = cuda and os.environ.get("HG_NO_GRAD_OVERLAP", "0") in ("", "0")
        # D(real) on a side stream in the D step (needs the wgrad side stream: both passes' weight gradients are ordered there);
        # HG_D_REAL_SIDE=0 switches it off
        self._d_real_side = (cuda and ops.WGRAD_SIDE_STREAM and self._sn_prefetch and compute_dtype == torch.bfloat16
                             and os.environ.get("HG_D_REAL_SIDE", "1") not in ("", "0"))
        if self._d_real_side:
            # leaf gradients now arrive from two streams on purpose (autograd synchronises them)
            quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if quiet is not None:
                quiet(False)
        ao = os.environ.get("HG_ADAM_OVERLAP", "")
        self._adam_overlap = overlap and self._flat_adam and (ao == "1" or (ao == "" and self.world > 1))
        if self._adam_overlap:
            inv = 1.0 / self.world
            self.d_grads.bucket_update = lambda s, e: self.opt_d.apply_range(s, e, inv)
            self.g_grads.bucket_update = lambda s, e: self.opt_g.apply_range(s, e, inv)
        if overlap and (self.world > 1 or self._adam_overlap):
            w = self.world
            gen.block3.convTranspose.weight._hg_grad_ready = lambda: self.g_grads.fire(0, w)
            gen.convTranspose2d1.weight._hg_grad_ready = lambda: self.g_grads.fire(1, w)
            gen.block2.convTranspose.weight._hg_grad_ready = lambda: self.g_grads.fire(2, w)
            gen.block1.convTranspose.weight._hg_grad_ready = lambda: self.g_grads.fire(3, w)
            # the D step runs the discriminator on real and on fake: a bucket is final after the second backward
            disc.blocks[2].conv2d.weight_orig._hg_grad_ready = lambda: self.d_grads.fire(0, w, after_calls=2)
            disc.blocks[1].conv2d.weight_orig._hg_grad_ready = lambda: self.d_grads.fire(1, w, after_calls=2)
        self._graphs = None
        lam = hologan_lr_lambda(cfg.num_epochs)
        self.sched_d = torch.optim.lr_scheduler.LambdaLR(self.opt_d, lam)
        self.sched_g = torch.optim.lr_scheduler.LambdaLR(self.opt_g, lam)
        # decorrelated latent / view streams per rank (the reference's identical seed-42-per-rank is a quirk)
        self.noise_rng = torch.Generator().manual_seed(seed + 1000 * (rank + 1))
        self.view_rng = np.random.RandomState(seed + 1000 * (rank + 1))
        self.logs: Dict[str, torch.Tensor] = {}

    # ---- sampling (host side, like the reference) -----------------------------------------------
    def sample_noise(self, n: int) -> torch.Tensor:
        """U(-1,1) latent (conf/noise_distn/uniform.yaml), drawn on the host (lightning_module.py:212)."""
        return torch.rand(n, self.cfg.noise_dim, generator=self.noise_rng) * 2 - 1

    def sample_view(self, n: int) -> np.ndarray:
        a = self.cfg.view_args
        rs = self.view_rng
        view = np.zeros((n, 6))
        view[:, 0] = rs.randint(a.azimuth_low, a.azimuth_high, n).astype(np.float64) * math.pi / 180.0
        if a.elevation_low < a.elevation_high:
            view[:, 1] = rs.randint(a.elevation_low, a.elevation_high, n).astype(np.float64) * math.pi / 180.0
        view[:, 2] = float(rs.uniform(a.scale_low, a.scale_high))
        for col, (lo, hi) in enumerate(((a.transX_low, a.transX_high), (a.transY_low, a.transY_high),
                                        (a.transZ_low, a.transZ_high)), start=3):
            view[:, col] = lo + rs.random_sample(n) * (hi - lo)
        return view

    # ---- the step the metric counts -------------------------------------------------------------
    def _autocast(self):
        enabled = self.compute_dtype != torch.float32 and self.device.type == "cuda"
        return torch.autocast("cuda", dtype=self.compute_dtype, enabled=enabled)

    def training_step(self, real: torch.Tensor, z: torch.Tensor, view, optimizer_idx: int) -> torch.Tensor:
        """Losses of HOLOGAN.training_step (lightning_module.py:209-237).  `z` and `view` are explicit
        (the reference samples them inside); `real` is (B,3,H,W) in [-1,1] on the device."""
        bce = F.binary_cross_entropy_with_logits
        cuda = self.device.type == "cuda"
        if cuda and self._sn_prefetch and self.compute_dtype == torch.bfloat16:
            # the discriminator's spectral-norm power iterations (two forwards in the D step, one in a G step) overlap the
            # generator's forward on a side stream
            self.discriminator.prefetch_spectral_norm(2 if optimizer_idx == 0 else 1)
        if optimizer_idx == 0:
            if cuda and self._d_real_side:
                # D(real) does not depend on the generator: its forward runs on a side stream beside the generator's
                # (no-grad) forward, and autograd then runs its backward on that stream beside D(fake)'s.  The discriminator's
                # kernels at B = 64 fill a fraction of the GPU (<= 148 CTAs, 5-30 us each), so the two chains overlap
                # almost for free.  The spectral-norm states are consumed in the same order (real first).
                cur, side = torch.cuda.current_stream(self.device), ops.side_stream(self.device, "dreal")
                side.wait_stream(cur)
                with torch.cuda.stream(side), self._autocast():
                    d_real, _ = self.discriminator(real)
                real.record_stream(side)
                with torch.no_grad(), self._autocast():
                    fake = self.generator(z, view_in=view)
                with self._autocast():
                    d_fake, z_pred = self.discriminator(fake)
                cur.wait_stream(side)
                d_real.record_stream(cur)
            else:
                with torch.no_grad(), self._autocast():     # the D step detaches fake (:221): no G graph is needed
                    fake = self.generator(z, view_in=view)
                with self._autocast():
                    d_real, _ = self.discriminator(real)
                    d_fake, z_pred = self.discriminator(fake)
            if cuda:                                    # both losses + their gradients: one launch each way
                loss, parts = ops.hologan_d_loss(d_real, d_fake, z_pred, z)
                self.logs["train/d_loss"], self.logs["train/q_loss"] = parts[0], parts[1]
                return loss
            d_real, d_fake, z_pred = d_real.float(), d_fake.float(), z_pred.float()
            loss_d = (bce(d_real, torch.ones_like(d_real)) + bce(d_fake, torch.zeros_like(d_fake))) / 2
            q = torch.mean((z_pred - z) ** 2)
            self.logs["train/d_loss"], self.logs["train/q_loss"] = loss_d.detach(), q.detach()
            return loss_d + q
        with self._autocast():
            fake = self.generator(z, view_in=view)
            out, z_pred = self.discriminator(fake)
        if cuda:
            loss, parts = ops.hologan_g_loss(out, z_pred, z)
            self.logs["train/g_loss"], self.logs["train/q_loss"] = parts[0], parts[1]
            return loss
        out, z_pred = out.float(), z_pred.float()
        loss_g = bce(out, torch.ones_like(out))
        q = torch.mean((z_pred - z) ** 2)
        self.logs["train/g_loss"], self.logs["train/q_loss"] = loss_g.detach(), q.detach()
        return loss_g + q

    def step(self, real: torch.Tensor, batch_idx: int, z: Optional[torch.Tensor] = None, view=None) -> torch.Tensor:
        """One optimizer step of the [D, G, G] schedule on one batch; returns the (detached) loss.
        `view` may be a (B,6) host array or a (B,4,4) tensor of precomputed inverse transforms."""
        idx = optimizer_index(batch_idx, self.cfg.disc_freq, self.cfg.gen_freq)
        n = real.shape[0]
        if z is None:
            z = self.sample_noise(n).to(self.device, non_blocking=True)
        if view is None:
            view = self.sample_view(n)
        if self._graphs is not None and n == self._static["real"].shape[0]:
            return self._replay(real, z, view, idx)
        return self._eager_step(real, z, view, idx)

    def step_host(self, real: torch.Tensor, batch_idx: int, z: Optional[torch.Tensor] = None, view=None) -> PendingLoss:
        """`step` for HOST inputs, pipelined: the inputs go through pinned memory and a copy stream into device
        staging buffers (so the transfer of step i overlaps the kernels of step i-1), the step runs on the current
        stream, and the loss comes back through an asynchronous D2H copy -- read it with `.item()` whenever it is
        needed (reading the previous step's loss after launching the next one keeps the GPU busy).  Two slots
        alternate; a slot is reused only after the step that used it has finished."""
        if self.device.type != "cuda":
            raise RuntimeError("step_host needs a CUDA device (the B200 path has no CPU fallback)")
        n = real.shape[0]
        if z is None:
            z = self.sample_noise(n)
        if view is None:
            view = self.sample_view(n)
        a = view if isinstance(view, torch.Tensor) and view.dim() == 3 else ops.view_to_affine(view, 16, 16)
        feeds = getattr(self, "_feeds", None)
        if feeds is None or feeds[0].real_pin.shape[0] != n:
            feeds = self._feeds = [_FeedSlot(n, self.cfg.channels_img, self.cfg.img_size, self.cfg.noise_dim, self.device)
                                   for _ in range(2)]
            self._feed_next = 0
            self._copy_stream = torch.cuda.Stream(device=self.device)
        slot = feeds[self._feed_next]
        self._feed_next ^= 1
        slot.done.synchronize()                               # normally long finished: the caller lags one step at most
        src_real = real if real.is_pinned() else slot.real_pin.copy_(real)
        slot.z_pin.copy_(z)
        slot.a_pin.copy_(a)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            slot.real_dev.copy_(src_real, non_blocking=True)
            slot.z_dev.copy_(slot.z_pin, non_blocking=True)
            slot.a_dev.copy_(slot.a_pin, non_blocking=True)
            slot.ready.record(self._copy_stream)
        cur.wait_event(slot.ready)
        loss = self.step(slot.real_dev, batch_idx, z=slot.z_dev, view=slot.a_dev)
        slot.loss_pin.copy_(loss.detach().reshape(1), non_blocking=True)
        slot.done.record(cur)
        return PendingLoss(slot.loss_pin, slot.done)

    def _eager_step(self, real, z, view, idx):
        grads, opt = (self.d_grads, self.opt_d) if idx == 0 else (self.g_grads, self.opt_g)
        # Lightning's toggle_optimizer: only the stepped network's parameters require grad
        for p in self.discriminator.parameters():
            p.requires_grad_(idx == 0)
        grads.begin()
        if self._adam_overlap:
            opt.tick()                                  # step counter / bias corrections once, before any bucket update
            loss = self.training_step(real, z, view, idx)
            loss.backward()                             # hooks: buckets exchanged + updated on the side stream
            if idx == 0 and self._d_real_side:
                torch.cuda.current_stream(self.device).wait_stream(ops.side_stream(self.device, "dreal"))
            grads.finish()
            grads.flush(self.world)                     # the tail, then join
        else:
            loss = self.training_step(real, z, view, idx)
            loss.backward()
            if idx == 0 and self._d_real_side:
                torch.cuda.current_stream(self.device).wait_stream(ops.side_stream(self.device, "dreal"))
            grads.finish()
        if self._adam_overlap:
            pass
        elif self._flat_adam:
            grads.all_reduce_sum(self.world)            # buckets fired during the backward + the tail; 1 / world goes into Adam
            opt.step(grad_scale=1.0 / self.world)
        else:
            grads.all_reduce_mean(self.world)
            opt.step()
        if self.device.type == "cuda":          # bf16 operand copies of the stepped network: once per update, not per forward
            ops.refresh_packed_weights((self.discriminator if idx == 0 else self.generator).parameters())
        return loss.detach()

    # ---- CUDA graphs: the whole step (forward, backward, gradient all-reduce, Adam) replayed as one launch ----
    def enable_cuda_graphs(self, batch_size: Optional[int] = None) -> None:
        """Capture one graph per optimizer index on static input buffers.  The eager steps that stream
        capture needs as warm-up are rolled back, so enabling graphs does not change the training state."""
        if self.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        b = batch_size or self.cfg.batch_size
        s = self.cfg.img_size
        self._static = {"real": torch.zeros(b, self.cfg.channels_img, s, s, device=self.device),
                        "z": torch.zeros(b, self.cfg.noise_dim, device=self.device),
                        "a": torch.eye(4, device=self.device).repeat(b, 1, 1)}
        st = self._static
        opts = (self.opt_d, self.opt_g)
        model_t = list(self.generator.state_dict().values()) + list(self.discriminator.state_dict().values())
        model_saved = [t.clone() for t in model_t]
        fresh = [len(o.state) == 0 or (isinstance(o, FlatAdam) and float(o.kstate[0]) == 0) for o in opts]   # torch creates Adam state lazily
        opt_saved = [None if f else {id(p): {k: v.clone() for k, v in stt.items() if isinstance(v, torch.Tensor)}
                                     for p, stt in o.state.items()} for o, f in zip(opts, fresh)]
        # The capture stream -- the step's critical chain -- gets a HIGHER priority than the side streams (wgrad, sn, dreal,
        # comm: default priority): the CTAs of a critical-chain kernel are scheduled first, the side work fills the SMs that
        # its tail waves and launch gaps leave idle.  Stream priorities are recorded in the captured kernel nodes.
        # HG_CAPTURE_PRIORITY=0 keeps the default priority (A/B).
        prio = -1 if os.environ.get("HG_CAPTURE_PRIORITY", "1") not in ("", "0") else 0
        side = torch.cuda.Stream(device=self.device, priority=prio)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for idx in (0, 1, 1):
                self._eager_step(st["real"], st["z"], st["a"], idx)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graphs = {}
        for idx in (0, 1):
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count
            if self.world > 1:
                # the warm-up all-reduces must have left the NCCL watchdog's queue before capture starts, and the
                # watchdog thread's event queries must not invalidate the capture: thread-local error mode
                torch.cuda.synchronize(self.device)
                dist.barrier()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local" if self.world > 1 else "global"):
                loss = self._eager_step(st["real"], st["z"], st["a"], idx)
            graphs[idx] = (g, loss, _lib.launch_count - n0)
        with torch.no_grad():
            for t, v in zip(model_t, model_saved):
                t.copy_(v)
            for o, f, saved in zip(opts, fresh, opt_saved):
                for p, stt in o.state.items():
                    for k, v in stt.items():
                        if isinstance(v, torch.Tensor):
                            v.zero_() if f else v.copy_(saved[id(p)][k])
        self._refresh_packs()                  # the packed operand copies follow the rolled-back parameters
        self._graphs = graphs

    def _refresh_packs(self):
        if self.device.type == "cuda":
            ops.refresh_packed_weights(list(self.generator.parameters()) + list(self.discriminator.parameters()))

    def _replay(self, real, z, view, idx):
        st = self._static
        st["real"].copy_(real, non_blocking=True)
        st["z"].copy_(z, non_blocking=True)
        if isinstance(view, torch.Tensor) and view.dim() == 3:
            st["a"].copy_(view, non_blocking=True)
        else:
            size = 16
            st["a"].copy_(ops.view_to_affine(view, size, size), non_blocking=True)
        g, loss, launches = self._graphs[idx]
        g.replay()
        _lib.launch_count += launches          # kernels of libhologan_b200.so inside the replayed graph
        return loss

    # ---- checkpoint / resume: the Lightning checkpoint layout the reference writes and resumes from ----------
    # (run_network.py:48-50 ModelCheckpoint, :61-71 resume_from_checkpoint)
    def checkpoint(self, epoch: int = 0, global_step: int = 0) -> dict:
        """A dict in Lightning's checkpoint layout: `state_dict` with `generator.*` / `discriminator.*` keys in the
        reference's names and torch-native (contiguous, fp32) layouts, the two optimizer states in the reference's
        order ([D, G], lightning_module.py:75-87), the LambdaLR states and the host RNG streams.  `torch.save` it."""
        sd = {}
        for prefix, net in (("generator.", self.generator), ("discriminator.", self.discriminator)):
            for k, v in net.state_dict().items():
                sd[prefix + k] = v.detach().to("cpu").contiguous().clone()
        return {
            "epoch": int(epoch), "global_step": int(global_step), "state_dict": sd,
            "optimizer_states": [self.opt_d.state_dict(), self.opt_g.state_dict()],
            "lr_schedulers": [self.sched_d.state_dict(), self.sched_g.state_dict()],
            "hologan_b200": {"noise_rng": self.noise_rng.get_state(), "view_rng": self.view_rng.get_state()},
        }

    def load_checkpoint(self, ckpt: dict, strict: bool = True, load_optimizers: bool = True) -> None:
        """Load a checkpoint written by `checkpoint()` or by the reference (a Lightning `.ckpt`: only `state_dict` is
        required).  Parameters are copied in place (their memory format and the flat-gradient views stay valid);
        captured CUDA graphs are dropped because the optimizer state tensors are replaced -- call
        `enable_cuda_graphs()` again after loading."""
        sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        for prefix, net in (("generator.", self.generator), ("discriminator.", self.discriminator)):
            part = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
            if part or strict:
                net.load_state_dict(part, strict=strict)
        if load_optimizers and "optimizer_states" in ckpt:
            for opt, st in zip((self.opt_d, self.opt_g), ckpt["optimizer_states"]):
                opt.load_state_dict(st)
            for sch, st in zip((self.sched_d, self.sched_g), ckpt.get("lr_schedulers", ())):
                sch.load_state_dict(st)
        extra = ckpt.get("hologan_b200")
        if extra:
            self.noise_rng.set_state(extra["noise_rng"])
            self.view_rng.set_state(extra["view_rng"])
        self._refresh_packs()
        self._graphs = None

    def end_epoch(self):
        self.sched_d.step()
        self.sched_g.step()
