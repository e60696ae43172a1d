"""`torch.ops.hologan.*`: the operators of the HoloGAN hot path registered with `torch.library`
("torch.ops.<ns>.* + autograd registration"), on top of the C ABI of include/hologan_b200.h.

Two kinds of registration:

* LEAF ops (`torch.library.custom_op`, CUDA only, fake/meta kernels for shape propagation): the rigid-body rotate +
  trilinear resample with its adjoint attached through `register_autograd` (`rotate_resample`,
  `rotate_resample_backward`) -- `lightning_gan_zoo_b200.ops.rotate_resample`, hence `Generator.transformation3d` and
  the bf16 pipeline, dispatch through it -- and the raw tcgen05 transposed-convolution passes `convt_forward`,
  `convt_dgrad`, `convt_wgrad` (forward-only leaves over packed weights).
* COMPOSITE ops (`CompositeImplicitAutograd`): `adain_act`, `adain_act_channels_last`, `final_conv_tanh`, `linear_relu`
  and the operators whose Python layer keeps per-parameter state (cached bf16 weight packs, gradient stores straight
  into a flat buffer) -- `convt`, `conv5s2_sn`, `dconv0`, `dheads` -- and the two losses.  They are visible to the
  dispatcher under `torch.ops.hologan.*`; autograd flows through the `torch.autograd.Function` inside.

There is no CPU kernel behind any of them: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops as _ops
from ._lib import HG_NCDHW

_NS = "hologan"
_CUDA = "cuda"


# ---- rotate + resample ------------------------------------------------------------------------------------

@torch.library.custom_op(f"{_NS}::rotate_resample", mutates_args=(), device_types=_CUDA)
def rotate_resample(vol: Tensor, a_inv: Tensor, border: int, in_layout: int, out_layout: int) -> Tensor:
    return _ops.rotate_fwd_raw(vol, a_inv, border, in_layout, out_layout)


@rotate_resample.register_fake
def _(vol, a_inv, border, in_layout, out_layout):
    if in_layout == HG_NCDHW:
        b, c, s = vol.shape[0], vol.shape[1], vol.shape[2]
    else:
        b, s, c = vol.shape[0], vol.shape[1], vol.shape[4]
    shape = (b, c, s, s, s) if out_layout == HG_NCDHW else (b, s, s, s, c)
    return vol.new_empty(shape)


@torch.library.custom_op(f"{_NS}::rotate_resample_backward", mutates_args=(), device_types=_CUDA)
def rotate_resample_backward(grad_out: Tensor, a_inv: Tensor, channels: int, size: int, border: int, in_layout: int,
                             out_layout: int) -> Tensor:
    return _ops.rotate_bwd_raw(grad_out, a_inv, channels, size, border, in_layout, out_layout)


@rotate_resample_backward.register_fake
def _(grad_out, a_inv, channels, size, border, in_layout, out_layout):
    b = grad_out.shape[0]
    shape = (b, channels, size, size, size) if in_layout == HG_NCDHW else (b, size, size, size, channels)
    return grad_out.new_empty(shape)


def _rotate_setup(ctx, inputs, output):
    vol, a_inv, border, in_layout, out_layout = inputs
    ctx.save_for_backward(a_inv)
    ctx.meta = (border, in_layout, out_layout, vol.shape[1] if in_layout == HG_NCDHW else vol.shape[4], vol.shape[2])


def _rotate_backward(ctx, grad_out):
    (a_inv,) = ctx.saved_tensors
    border, in_layout, out_layout, c, s = ctx.meta
    return rotate_resample_backward(grad_out.contiguous(), a_inv, c, s, border, in_layout, out_layout), None, None, None, None


rotate_resample.register_autograd(_rotate_backward, setup_context=_rotate_setup)


# ---- raw tcgen05 transposed-convolution passes (forward-only leaves) ------------------------------------------

@torch.library.custom_op(f"{_NS}::convt_forward", mutates_args=(), device_types=_CUDA)
def convt_forward(x_cl: Tensor, w_fwd: Tensor, bias: Optional[Tensor], ndim: int, kernel: int, neg_slope: float) -> Tensor:
    """y_s2d = act(convT(x) + bias) from the packed weight w_fwd [taps][Cout][Cin] (include/hologan_b200.h hg_convt_fwd)."""
    import ctypes
    from . import _lib
    b, size, cin = _ops._conv_dims(x_cl, ndim)
    cout = w_fwd.shape[1]
    nclass = 1 if kernel == 1 else 2 ** ndim
    y = torch.empty((b,) + (size,) * ndim + (nclass, cout), dtype=torch.bfloat16, device=x_cl.device)
    _lib.call("hg_convt_fwd", _ops._ptr(x_cl), _ops._ptr(w_fwd), _ops._ptr(bias), _ops._ptr(y), b, cin, cout, ndim, size, kernel,
              ctypes.c_float(neg_slope), _ops._stream())
    return y


@convt_forward.register_fake
def _(x_cl, w_fwd, bias, ndim, kernel, neg_slope):
    nclass = 1 if kernel == 1 else 2 ** ndim
    return x_cl.new_empty(tuple(x_cl.shape[:-1]) + (nclass, w_fwd.shape[1]))


@torch.library.custom_op(f"{_NS}::convt_dgrad", mutates_args=(), device_types=_CUDA)
def convt_dgrad(dy_s2d: Tensor, w_dgrad: Tensor, ndim: int, kernel: int) -> Tensor:
    from . import _lib
    b, size, cout = dy_s2d.shape[0], dy_s2d.shape[1], dy_s2d.shape[-1]
    cin = w_dgrad.shape[1]
    dx = torch.empty((b,) + (size,) * ndim + (cin,), dtype=torch.bfloat16, device=dy_s2d.device)
    _lib.call("hg_convt_dgrad", _ops._ptr(dy_s2d.contiguous()), _ops._ptr(w_dgrad), _ops._ptr(dx), b, cin, cout, ndim, size, kernel,
              _ops._stream())
    return dx


@convt_dgrad.register_fake
def _(dy_s2d, w_dgrad, ndim, kernel):
    return dy_s2d.new_empty(tuple(dy_s2d.shape[:ndim + 1]) + (w_dgrad.shape[1],))


@torch.library.custom_op(f"{_NS}::convt_wgrad", mutates_args=(), device_types=_CUDA)
def convt_wgrad(x_cl: Tensor, dy_s2d: Tensor, wshape: list[int], ndim: int, kernel: int) -> Tensor:
    return _ops.convt_wgrad(x_cl, dy_s2d.contiguous(), tuple(wshape), ndim, kernel)


@convt_wgrad.register_fake
def _(x_cl, dy_s2d, wshape, ndim, kernel):
    return x_cl.new_empty(tuple(wshape), dtype=torch.float32)


# ---- composite registrations ---------------------------------------------------------------------------------
_lib_def = torch.library.Library(_NS, "FRAGMENT")


def _composite(name: str, schema: str, fn):
    _lib_def.define(f"{name}{schema}")
    _lib_def.impl(name, fn, "CompositeImplicitAutograd")


_composite("adain_act", "(Tensor x, Tensor scale, Tensor? bias, float neg_slope=0.0, float eps=1e-8, bool biased_var=False) -> Tensor",
           lambda x, scale, bias, neg_slope=0.0, eps=1e-8, biased_var=False: _ops._AdaInAct.apply(x, scale, bias, neg_slope, eps, biased_var))
_composite("adain_act_channels_last",
           "(Tensor x, Tensor? scale, Tensor? bias, int ndim, int classes, float neg_slope=0.0, float eps=1e-8, bool biased_var=False, "
           "Tensor? stats=None) -> Tensor",
           lambda x, scale, bias, ndim, classes, neg_slope=0.0, eps=1e-8, biased_var=False, stats=None:
           _ops._AdaInChannelsLast.apply(x, scale, bias, ndim, classes, neg_slope, eps, biased_var, stats))
_composite("convt", "(Tensor x_cl, Tensor weight, Tensor? bias, int ndim, int kernel, float neg_slope=1.0, int perm_c=0, int perm_s=0) -> Tensor",
           lambda x_cl, weight, bias, ndim, kernel, neg_slope=1.0, perm_c=0, perm_s=0:
           _ops._ConvT.apply(x_cl, weight, bias, ndim, kernel, neg_slope, (perm_c, perm_s), False))
_composite("final_conv_tanh", "(Tensor x_cl, Tensor weight, Tensor bias) -> Tensor",
           lambda x_cl, weight, bias: _ops._FinalConvTanh.apply(x_cl, weight, bias))
_composite("linear_relu", "(Tensor z, Tensor weight, Tensor bias) -> Tensor",
           lambda z, weight, bias: _ops._LinearRelu.apply(z, weight, bias))
_composite("dconv0", "(Tensor x, Tensor weight, Tensor bias, float neg_slope=0.2) -> Tensor",
           lambda x, weight, bias, neg_slope=0.2: _ops._DConv0.apply(x, weight, bias, neg_slope))
_composite("conv5s2_sn", "(Tensor x_s2d, Tensor weight_orig, Tensor? state) -> Tensor",
           lambda x_s2d, weight_orig, state: _ops._Conv5s2SN.apply(x_s2d, weight_orig, state))
_composite("dheads", "(Tensor h, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor w3, Tensor b3, float neg_slope=0.2) -> (Tensor, Tensor)",
           lambda h, w1, b1, w2, b2, w3, b3, neg_slope=0.2: _ops._DHeads.apply(h, w1, b1, w2, b2, w3, b3, neg_slope))
_composite("hologan_d_loss", "(Tensor d_real, Tensor d_fake, Tensor z_pred, Tensor z) -> (Tensor, Tensor)",
           lambda d_real, d_fake, z_pred, z: _ops._GanLoss.apply(d_real, d_fake, z_pred, z, 1.0, 0.5, 0.0, 0.5))
_composite("hologan_g_loss", "(Tensor d_fake, Tensor z_pred, Tensor z) -> (Tensor, Tensor)",
           lambda d_fake, z_pred, z: _ops._GanLoss.apply(d_fake, None, z_pred, z, 1.0, 1.0, 0.0, 0.0))

REGISTERED = ("rotate_resample", "rotate_resample_backward", "convt_forward", "convt_dgrad", "convt_wgrad", "adain_act",
              "adain_act_channels_last", "convt", "final_conv_tanh", "linear_relu", "dconv0", "conv5s2_sn", "dheads",
              "hologan_d_loss", "hologan_g_loss")
