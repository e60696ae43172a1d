"""HoloGAN generator on the B200 path -- drop-in for the reference's
`core.models.hologan_generator` (ebartrum/lightning_gan_zoo @ 33c7f1b).

Same public names, constructor / forward signatures and `state_dict` keys as the reference
(core/models/hologan_generator.py:7-345), so Hydra's
`_target_: core.models.hologan_generator.Generator` (conf/expt/hologan.yaml:29-49) and reference
checkpoints keep working (see `lightning_gan_zoo_b200.compat.install`).  The arithmetic of the hot
path -- AdaIN(+ReLU), the learned-constant broadcast, the rigid-body rotate/resample and its
backward -- runs in `libhologan_b200.so` through `lightning_gan_zoo_b200.ops`.

Deviations from the reference, all deliberate:
  * `sample_view` uses np.float64 (the reference's `np.float` crashes on numpy >= 1.24, R7);
  * `img_size == 128` builds the patched, working head (`ConvTranspose2d(k4, s2, p1)`, R4);
  * no gradient is propagated into the view parameters (nobody consumes it);
  * `view_in` may also be a (B,4,4) fp32 tensor of precomputed inverse transforms
    (`ops.view_to_affine`), so a caller can keep them resident on the device.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from ... import ops


class ZMapping(nn.Module):
    """z -> (scale, bias) style vectors.  Reference :7-18."""

    def __init__(self, z_dimension, output_channel):
        super().__init__()
        self.output_channel = output_channel
        self.linear1 = nn.Linear(z_dimension, 2 * output_channel)
        nn.init.normal_(self.linear1.weight, std=0.02)
        nn.init.zeros_(self.linear1.bias)

    def style(self, x):
        """The packed (B, 2C) output [scale | bias]: the AdaIN kernels read both halves in place, so the hot path
        never slices it (two slices cost two zero-fills + two copies + an add in autograd's backward)."""
        # the style path stays fp32 even under autocast: it is 0.01 % of the FLOPs and every activation
        # of the block is scaled by it
        if x.is_cuda:
            return ops.linear_relu(x, self.linear1.weight, self.linear1.bias)        # fp32 SIMT kernel
        with torch.autocast(x.device.type, enabled=False):
            return F.relu(F.linear(x.float(), self.linear1.weight, self.linear1.bias))

    def forward(self, x):
        style = self.style(x)
        c = self.output_channel
        return style[:, :c], style[:, c:]


def AdaIn(features, scale, bias):
    """Plain adaptive instance norm (unbiased variance, eps 1e-8).  Reference :333-345."""
    return ops.adain_act(features, scale.float(), bias.float(), neg_slope=1.0)


class BasicBlock(nn.Module):
    """Transposed conv (x2 upsampling) -> AdaIN -> ReLU.  Reference :20-42."""

    def __init__(self, z_planes, in_planes, out_planes, transpose_dim):
        super().__init__()
        if transpose_dim == 2:
            self.convTranspose = nn.ConvTranspose2d(in_planes, out_planes, kernel_size=4, stride=2, padding=1)
        elif transpose_dim == 3:
            self.convTranspose = nn.ConvTranspose3d(in_planes, out_planes, kernel_size=3, stride=2, padding=1,
                                                    output_padding=1)
        else:
            raise ValueError("transpose_dim must be 2 or 3")
        nn.init.normal_(self.convTranspose.weight, std=0.02)
        nn.init.zeros_(self.convTranspose.bias)
        self.zMapping = ZMapping(z_planes, out_planes)

    def forward(self, h, z):
        h = self.convTranspose(h)
        return ops.adain_act(h, self.zMapping.style(z), None, neg_slope=0.0)      # AdaIN + ReLU in one pass


class Generator(nn.Module):
    def __init__(self, in_planes, out_planes, z_planes, view_args, img_size, view_planes=6, gpu=True):
        super().__init__()
        if img_size not in (64, 128):
            raise ValueError("img_size must be 64 or 128")
        self.device = torch.device("cuda" if gpu else "cpu")
        self.view_args = view_args
        self.img_size = img_size
        # border handling of the resampler: reference arithmetic (bit-exact) by default
        self.rotate_border = ops.HG_BORDER_REFERENCE

        self.x = nn.Parameter(((torch.randn(1, in_planes * 8, 4, 4, 4) - 0.5) / 0.5).to(self.device))
        self.zMapping = ZMapping(z_planes, in_planes * 8)
        self.block1 = BasicBlock(z_planes, in_planes * 8, in_planes * 2, transpose_dim=3)
        self.block2 = BasicBlock(z_planes, in_planes * 2, in_planes, transpose_dim=3)

        self.convTranspose2d1 = nn.ConvTranspose2d(in_planes * 16, in_planes * 16, kernel_size=1)
        nn.init.normal_(self.convTranspose2d1.weight, std=0.02)
        nn.init.zeros_(self.convTranspose2d1.bias)

        self.block3 = BasicBlock(z_planes, in_planes * 16, in_planes * 4, transpose_dim=2)
        self.block4 = BasicBlock(z_planes, in_planes * 4, in_planes, transpose_dim=2)

        if img_size == 64:
            self.final_layer = nn.Conv2d(in_planes, out_planes, kernel_size=3, padding=1)
        else:   # patched 128 head: the reference's lacks stride=2 and yields 65x65
            self.final_layer = nn.ConvTranspose2d(in_planes, out_planes, kernel_size=4, stride=2, padding=1)
        nn.init.normal_(self.final_layer.weight, std=0.02)
        nn.init.zeros_(self.final_layer.bias)

    # ---- views ------------------------------------------------------------------------------
    def sample_view(self, batch_size):
        """Random (azimuth, elevation, scale, tx, ty, tz) rows; same numpy RNG call order as the
        reference (:80-114): azimuth ints, elevation ints, one scale, then the three shifts."""
        a = self.view_args
        view = np.zeros((batch_size, 6), dtype=np.float64)
        view[:, 0] = np.random.randint(a.azimuth_low, a.azimuth_high, batch_size).astype(np.float64) * math.pi / 180.0
        if a.elevation_low < a.elevation_high:
            view[:, 1] = np.random.randint(a.elevation_low, a.elevation_high, batch_size).astype(np.float64) \
                * math.pi / 180.0
        view[:, 2] = float(np.random.uniform(a.scale_low, a.scale_high))
        for col, (lo, hi) in enumerate(((a.transX_low, a.transX_high), (a.transY_low, a.transY_high),
                                        (a.transZ_low, a.transZ_high)), start=3):
            view[:, col] = lo + np.random.random(batch_size) * (hi - lo)
        return view

    def _affine(self, view_params, size, new_size, device):
        if isinstance(view_params, torch.Tensor) and view_params.dim() == 3 and tuple(view_params.shape[1:]) == (4, 4):
            return view_params.to(device=device, dtype=torch.float32).contiguous()   # precomputed inverse transforms
        if size != new_size:
            raise NotImplementedError("resampling onto a grid of a different size is not on the hot path")
        return ops.view_to_affine(view_params, size, new_size).to(device, non_blocking=True)

    def transformation3d(self, voxel_array, view_params, size=16, new_size=16):
        """Rigid-body transform + trilinear resample of (B,C,S,S,S).  Reference :145-243."""
        a_inv = self._affine(view_params, size, new_size, voxel_array.device)
        return ops.rotate_resample(voxel_array.contiguous(), a_inv, self.rotate_border)

    # ---- forward ----------------------------------------------------------------------------
    def _use_tensor_core_path(self, z):
        """bf16 autocast + channel counts the tcgen05 implicit-GEMM kernels cover (in_planes % 64 == 0)."""
        if not (z.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            return False
        c0 = self.x.shape[1]
        return c0 % 512 == 0

    @staticmethod
    def _convt_adain(h, block, style, ndim, kernel):
        """Transposed convolution (tcgen05 tap GEMM, s2d output) -> AdaIN + ReLU with the depth-to-space shuffle in its
        store.  With the library option ADAIN_GEMM_STATS the GEMM epilogue also emits the AdaIN statistics partials and
        the normalise pass is a single streaming kernel (measured break-even with the default single-pass cluster
        kernel on B200, so off by default)."""
        from ... import _lib
        w = block.convTranspose.weight
        classes = 2 ** ndim
        if _lib.get_option("ADAIN_GEMM_STATS") and ops.convt_stats_supported(w.shape[1], ndim, h.shape[1]):
            y, st = ops.convt(h, w, None, ndim, kernel, stats=True)
            return ops.adain_act_channels_last(y, style, None, ndim=ndim, classes=classes, stats=st)
        y = ops.convt(h, w, None, ndim, kernel)                                  # (B,[S,]S,S,P,Cout) s2d
        return ops.adain_act_channels_last(y, style, None, ndim=ndim, classes=classes)

    def _trunk_tensor_core(self, z):
        """bf16 pipeline, view-independent half: styles, learned constant, the two ConvTranspose3d blocks.
        Returns (h2 NDHWC bf16, styles).  The ConvTranspose biases in front of an AdaIN are not applied: instance
        normalisation removes any per-channel constant, so the output is unchanged and their gradient is
        identically zero (the reference only holds rounding noise there)."""
        bf16 = torch.bfloat16
        # all five ZMappings read the same z: one launch (packed (B, 2C) styles, fp32)
        maps = [self.zMapping] + [b.zMapping for b in (self.block1, self.block2, self.block3, self.block4)]
        styles = ops.linear_relu_group(z, [m.linear1.weight for m in maps], [m.linear1.bias for m in maps])
        h0 = ops.adain_act(self.x, styles[0], None, neg_slope=0.0)               # (B,8P,4,4,4) fp32, NC*
        h = ops.nc_to_channels_last(h0.to(bf16))                                 # (B,4,4,4,8P)
        for block, style in ((self.block1, styles[1]), (self.block2, styles[2])):
            h = self._convt_adain(h, block, style, 3, 3)                          # (B,2S,2S,2S,Cout) NDHWC
        return h, styles

    def _decode_tensor_core(self, h, a_inv, styles):
        """bf16 pipeline, view-dependent half: rotate + depth fold, 1x1 projection, the two ConvTranspose2d
        blocks, final layer + tanh.  h: (B,S,S,S,C) NDHWC bf16, a_inv: (B,4,4) fp32 on the device."""
        n, size = h.shape[0], h.shape[1]
        # rotate + fold depth into channels in one kernel: out[b, z, x, (y, c)] is the projection's A operand
        rot = ops.rotate_resample(h, a_inv, ops.HG_BORDER_ZERO, ops.HG_NDHWC, ops.HG_PROJ)
        c = rot.shape[-1]
        a_proj = rot.reshape(n, size, size, size * c)
        # K index of the operand is y*C + c; the reference's folded channel c*S + j pairs with y = S-1-j:
        # the weight pack / wgrad kernels apply that permutation (perm = (C, S)), the parameter keeps its layout
        h = ops.convt(a_proj, self.convTranspose2d1.weight, self.convTranspose2d1.bias, 2, 1, neg_slope=0.0,
                      perm=(c, size))                                             # 1x1 conv + bias + ReLU
        h = h.reshape(n, size, size, -1)
        for block, style in ((self.block3, styles[3]), (self.block4, styles[4])):
            h = self._convt_adain(h, block, style, 2, 4)                          # (B,2S,2S,Cout) NHWC
        if self.img_size == 64 and ops.final_conv_supported(h.shape[-1], self.final_layer.weight.shape[0]):
            return ops.final_conv_tanh(h, self.final_layer.weight, self.final_layer.bias)   # direct conv + tanh, fp32 out
        if self.img_size == 128 and ops.head128_supported(h.shape[-1], self.final_layer.weight.shape[1], h.shape[1]):
            return ops.head128_tanh(h, self.final_layer.weight, self.final_layer.bias)      # patched-128 head on mma.sync
        # other widths: stock op on the NHWC buffer viewed as channels_last NCHW
        return torch.tanh(self.final_layer(h.permute(0, 3, 1, 2)))

    def _forward_tensor_core(self, z, view_in):
        """bf16 pipeline: channels-last activations, transposed convs / projection on the tcgen05 kernels
        (`ops.convt`), conv outputs in space-to-depth layout."""
        h, styles = self._trunk_tensor_core(z)
        size = h.shape[1]
        return self._decode_tensor_core(h, self._affine(view_in, size, size, z.device), styles)

    # ---- fp32 parity path, split the same way --------------------------------------------------
    def _trunk(self, z):
        h0 = ops.adain_act(self.x, self.zMapping.style(z), None, neg_slope=0.0)      # constant never repeated B times
        return self.block2(self.block1(h0, z), z)

    def _decode(self, h2, view_in, z):
        batch_size = z.shape[0]
        rot = self.transformation3d(h2, view_in, h2.shape[2], h2.shape[2])
        # fold depth into channels: out[b, c*S + j, r, col] = rot[b, c, r, S-1-j, col]  (:130-133)
        s = rot.shape[2]
        h2_2d = rot.permute(0, 1, 3, 2, 4).flip(2).reshape(batch_size, -1, s, s)
        h3 = F.relu(self.convTranspose2d1(h2_2d))
        h4 = self.block3(h3, z)
        h5 = self.block4(h4, z)
        return torch.tanh(self.final_layer(h5))

    # ---- view sweeps ---------------------------------------------------------
    @torch.no_grad()
    def render_views(self, z, views):
        """Every latent under every view: (B, V, out_planes, H, W).

        The reference's figure callbacks (core/figures/types.py:217-239 ElevationStep, :300-322 ElevationGif) call
        `generator(z, view_in=view)` once per view with the SAME z, recomputing the view-independent 3D trunk
        (constant -> AdaIN -> two ConvTranspose3d blocks) every time.  Here the trunk runs once per z and only
        rotate -> projection -> 2D decoder runs per view; `render_views(z, views)[:, v]` is bit-identical to
        `forward(z, view_in=views[v])` under the same autocast state.

        views: (V, 6) rows (azimuth, elevation, scale, tx, ty, tz) shared by all latents, or (B, V, 6);
        numpy float64 or torch, like `view_in`."""
        n = z.shape[0]
        views = views.detach().cpu().numpy() if isinstance(views, torch.Tensor) else np.asarray(views)
        if views.ndim == 2:
            views = np.broadcast_to(views[None], (n,) + views.shape)
        if views.ndim != 3 or views.shape[0] != n or views.shape[2] != 6:
            raise ValueError("views must be (V, 6) or (B, V, 6)")
        n_views = views.shape[1]
        tc = self._use_tensor_core_path(z)
        if tc:
            h, styles = self._trunk_tensor_core(z)
            size = h.shape[1]
        else:
            h = self._trunk(z)
            size = h.shape[2]
        # all B*V inverse transforms in one host pass + one H2D copy
        a_inv = ops.view_to_affine(np.ascontiguousarray(views.reshape(n * n_views, 6)), size, size)
        a_inv = a_inv.reshape(n, n_views, 4, 4).to(z.device, non_blocking=True)
        out = None
        for v in range(n_views):
            a_v = a_inv[:, v].contiguous()
            img = self._decode_tensor_core(h, a_v, styles) if tc else self._decode(h, a_v, z)
            if out is None:
                out = torch.empty((n, n_views) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
            out[:, v] = img
        return out

    def forward(self, z, view_in=None):
        batch_size = z.shape[0]
        if view_in is None:
            view_in = self.sample_view(batch_size)
        if self._use_tensor_core_path(z):
            return self._forward_tensor_core(z, view_in)

        return self._decode(self._trunk(z), view_in, z)
