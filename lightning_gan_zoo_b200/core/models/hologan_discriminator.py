"""HoloGAN discriminator -- interface mirror of the reference's
`core.models.hologan_discriminator` (core/models/hologan_discriminator.py:7-78).

The discriminator is in the training-step metric.  Under bf16 autocast on CUDA with the
reference's widths it runs entirely on the kernels of libhologan_b200.so (`_forward_b200`); the fp32 parity path
and unsupported shapes use stock PyTorch modules.  The reference's `state_dict` keys are kept, including the
`conv2d_spec_norm` alias of each spectrally normalised convolution.  `img_size=128` uses the patched head sizes
(SURVEY.md R4).
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

from ... import ops


def truncated_normal_initializer(weight, mean=0, std=0.02):
    """Truncated N(mean, std) at +-2 sigma by picking the first in-range of 4 draws (:72-78)."""
    with torch.no_grad():
        draws = torch.randn(tuple(weight.shape) + (4,), dtype=weight.dtype, device=weight.device)
        in_range = (draws > -2) & (draws < 2)
        first = in_range.max(dim=-1, keepdim=True)[1]
        weight.copy_(draws.gather(-1, first).squeeze(-1) * std + mean)


class BasicBlock(nn.Module):
    """spectral-norm Conv5x5 s2 -> InstanceNorm2d -> LeakyReLU(0.2).  Reference :7-23."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.conv2d = nn.Conv2d(in_planes, out_planes, kernel_size=5, stride=2, padding=2)
        truncated_normal_initializer(self.conv2d.weight)
        nn.init.zeros_(self.conv2d.bias)
        # same module registered twice, like the reference (keys conv2d.* and conv2d_spec_norm.*)
        self.conv2d_spec_norm = nn.utils.spectral_norm(self.conv2d)
        self.instance_norm = nn.InstanceNorm2d(out_planes)

    def forward(self, x):
        h = self.conv2d_spec_norm(x)
        n = h.shape[2] * h.shape[3]
        if (h.is_cuda and h.dtype == torch.bfloat16 and h.shape[2] == h.shape[3] and h.shape[1] % 16 == 0 and 2 <= n <= 4096
                and h.is_contiguous(memory_format=torch.channels_last)):
            # channels-last bf16 pipeline: the norm + activation kernel works on the NHWC buffer in place of a
            # layout round trip (the tensor keeps its logical NCHW shape)
            y = ops.instance_norm_act_channels_last(h.permute(0, 2, 3, 1), 0.2, self.instance_norm.eps)
            return y.permute(0, 3, 1, 2)
        if h.is_cuda and h.dtype in (torch.float32, torch.bfloat16) and n % 8 == 0 and n <= 4096:
            # InstanceNorm2d + LeakyReLU(0.2) in one pass on the AdaIN kernel (biased variance, eps 1e-5)
            return ops.instance_norm_act(h.contiguous(), 0.2, self.instance_norm.eps)
        return F.leaky_relu(self.instance_norm(h), 0.2)


class Discriminator(nn.Module):
    def __init__(self, in_planes, out_planes, z_planes, img_size=64, final_sigmoid=False):
        """`img_size` / `final_sigmoid`: the reference's Hydra tree hands both to this constructor (conf/config.yaml:37-39
        merges a `discriminator: {img_size, final_sigmoid}` block into every experiment), although its own
        `Discriminator.__init__(in_planes, out_planes, z_planes)` accepts neither -- `instantiate(cfg.discriminator)` of
        `+expt=hologan` raises TypeError there.  Here `img_size` sizes the heads (patched 128 variant) and
        `final_sigmoid` must stay False (the loss is BCEWithLogits, conf/expt/hologan.yaml:12-13)."""
        super().__init__()
        if final_sigmoid:
            raise ValueError("the HoloGAN discriminator returns logits (BCEWithLogitsLoss); final_sigmoid must be False")
        self.conv2d = nn.Conv2d(in_planes, out_planes, kernel_size=5, stride=2, padding=2)
        truncated_normal_initializer(self.conv2d.weight)
        nn.init.zeros_(self.conv2d.bias)
        self.blocks = nn.Sequential(BasicBlock(out_planes, out_planes * 2),
                                    BasicBlock(out_planes * 2, out_planes * 4),
                                    BasicBlock(out_planes * 4, out_planes * 8))
        features = out_planes * 8 * (img_size // 16) ** 2       # 8192 at 64x64 (reference :41)
        self.linear1 = nn.Linear(features, 1)
        truncated_normal_initializer(self.linear1.weight)
        nn.init.zeros_(self.linear1.bias)
        # linear2 / linear3 keep torch's default weight init, as in the reference (:45-50 re-initialise
        # linear1 by mistake); only their biases are zeroed
        self.linear2 = nn.Linear(features, 128)
        nn.init.zeros_(self.linear2.bias)
        self.linear3 = nn.Linear(128, z_planes)
        nn.init.zeros_(self.linear3.bias)

    def _blocks_bf16(self, h):
        """The three spectral-norm blocks of the bf16 pipeline.  One grouped call normalises all three weights
        (power iteration on `weight_u` / `weight_v`, sigma, bf16 channels-last conv operand) instead of the hook's
        ~16 launches per layer; the convolution biases are not applied -- the InstanceNorm that follows removes any
        per-channel constant, so the output is unchanged and their gradient is identically zero (the reference only
        holds rounding noise there)."""
        convs = [blk.conv2d for blk in self.blocks]
        ws = ops.spectral_norm_weights([c.weight_orig for c in convs], [c.weight_u for c in convs],
                                       [c.weight_v for c in convs], power_iteration=self.training,
                                       out_dtype=torch.bfloat16)
        for blk, w in zip(self.blocks, ws):
            h = F.conv2d(h, w, None, stride=2, padding=2)
            h = ops.instance_norm_act_channels_last(h.permute(0, 2, 3, 1), 0.2, blk.instance_norm.eps).permute(0, 3, 1, 2)
        return h

    def _forward_b200(self, x):
        """The whole discriminator of the bf16 pipeline on hand-written kernels: first convolution +
        LeakyReLU on warp-level tensor cores writing the space-to-depth operand of block 0 directly, the three
        spectral-norm convolutions on the tcgen05 tap GEMMs (1 / sigma folded into the bf16 weight pack, K split over
        taps), InstanceNorm + LeakyReLU storing the next block's space-to-depth operand, and both heads in one pass.
        No cuDNN / cuBLAS kernel runs.  The convolution biases of the blocks are not applied: the InstanceNorm that
        follows removes any per-channel constant (their gradient is identically zero)."""
        self._b200_used = True
        h = ops.dconv0(x, self.conv2d.weight, self.conv2d.bias, 0.2)                  # (B, S/4, S/4, 4, 64)
        states = self._take_spectral_norm_states()
        for i, (blk, st) in enumerate(zip(self.blocks, states)):
            y = ops.conv5s2_sn(h, blk.conv2d.weight_orig, st)                         # (B, S', S', Cout)
            h = ops.instance_norm_act_channels_last(y, 0.2, blk.instance_norm.eps, s2d_out=i + 1 < len(self.blocks))
        return ops.dheads(h, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                          self.linear3.weight, self.linear3.bias, 0.2)

    def _spectral_norm_states(self):
        convs = [blk.conv2d for blk in self.blocks]
        return ops.spectral_norm_sigma([c.weight_orig for c in convs], [c.weight_u for c in convs],
                                       [c.weight_v for c in convs], power_iteration=self.training)

    def prefetch_spectral_norm(self, forwards: int = 1):
        """Run the power iteration + sigma of the next `forwards` forward passes NOW, on a side stream: the grouped
        spectral-norm kernels are tiny (3 .. 48 CTAs, ~40 us per forward) and depend only on the weights, so they overlap
        whatever the caller launches next on the current stream (the generator's forward in both the D and the G step)
        instead of sitting in front of the first block.  Each later `forward` consumes one prefetched state, in order --
        u / v advance exactly as if every forward had run its own iteration (reference
        core/models/hologan_discriminator.py:15: one power iteration per training-mode forward).  Works inside CUDA-graph
        capture (the side stream forks from and joins the capturing stream)."""
        w = self.blocks[0].conv2d.weight_orig
        if not w.is_cuda or not getattr(self, "_b200_used", False):     # only once a forward has taken the path that consumes them
            return
        cur = torch.cuda.current_stream(w.device)
        if getattr(self, "_sn_stream", None) is None or self._sn_stream.device != w.device:
            self._sn_stream = torch.cuda.Stream(device=w.device)
            self._sn_ready = []
        self._sn_stream.wait_stream(cur)                # the weights may have just been written by the optimizer
        with torch.cuda.stream(self._sn_stream):
            for _ in range(forwards):
                states = self._spectral_norm_states()
                ev = torch.cuda.Event()
                ev.record(self._sn_stream)
                for t in states:
                    t.record_stream(cur)
                self._sn_ready.append((states, ev, self.training))

    def _take_spectral_norm_states(self):
        ready = getattr(self, "_sn_ready", None)
        if ready:
            states, ev, training = ready.pop(0)
            if training == self.training:
                torch.cuda.current_stream(states[0].device).wait_event(ev)
                return states
            ready.clear()                               # mode changed since the prefetch: recompute in line
        return self._spectral_norm_states()

    def _b200_ok(self, x):
        import os
        if os.environ.get("HG_D_LIBRARY", "0") not in ("", "0"):                      # A/B switch: the cuDNN / cuBLAS pipeline
            return False
        if x.dim() != 4 or x.shape[-1] != x.shape[-2] or not ops.dconv0_supported(x.shape[1], self.conv2d.weight.shape[0], x.shape[-1]):
            return False
        side = x.shape[-1] // 2
        for blk in self.blocks:
            w = blk.conv2d.weight_orig
            side //= 2
            if not (w.is_contiguous() and ops.conv5s2_supported(w.shape[1], w.shape[0], side)):
                return False
        c = self.blocks[-1].conv2d.weight_orig.shape[0]
        return (ops.dheads_supported(x.shape[0], c, side * side) and self.linear2.weight.shape[0] == 128
                and self.linear3.weight.shape[1] == 128 and self.linear1.weight.shape[1] == c * side * side)

    def _bf16_pipeline_ok(self, x):
        if not (x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            return False
        w = self.blocks[0].conv2d.weight_orig
        side = x.shape[-1] // 4                                       # spatial extent after the first block
        return (x.shape[-1] == x.shape[-2] and w.shape[0] % 16 == 0 and side * side <= 4096 and (side // 4) ** 2 >= 2
                and all(b.conv2d.weight_orig.is_contiguous(memory_format=torch.channels_last) for b in self.blocks))

    def forward(self, x):
        bf16 = x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
        if bf16 and self._b200_ok(x):
            return self._forward_b200(x)
        if bf16:
            x = x.contiguous(memory_format=torch.channels_last)      # NHWC pipeline (3 MB at B = 64)
        h = F.leaky_relu(self.conv2d(x), 0.2)
        if bf16 and self._bf16_pipeline_ok(x):
            h = self._blocks_bf16(h).flatten(1)
        else:
            h = self.blocks(h).flatten(1)                             # logical (c, h, w) order, as the reference (:60)
        logits = self.linear1(h)
        z_prediction = torch.tanh(self.linear3(F.leaky_relu(self.linear2(h), 0.2)))
        return logits, z_prediction
