"""Torch-facing operators of the HoloGAN hot path.

Each operator is a thin `torch.autograd.Function` over the C ABI of `include/hologan_b200.h`
(raw device pointers + the current CUDA stream, through ctypes).  PyTorch only provides device
memory, streams and autograd bookkeeping here; all arithmetic on the path runs in
`libhologan_b200.so`.  There is no CPU implementation: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import HG_BF16, HG_BORDER_REFERENCE, HG_BORDER_ZERO, HG_F32, HG_NCDHW, HG_NDHWC, HG_PROJ  # noqa: F401

Tensor = torch.Tensor


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return HG_F32
    if t.dtype == torch.bfloat16:
        return HG_BF16
    raise TypeError(f"hologan_b200 ops support float32 and bfloat16 tensors, got {t.dtype}")


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hologan_b200 ops run on CUDA tensors only (no CPU fallback)")


def _ptr(t: Optional[Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------
# a5 + a6 (host part): view parameters -> inverse sampling transform
# ------------------------------------------------------------------------------------------------

def _homogeneous(batch: int) -> Tensor:
    return torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(batch, 1, 1)


def view_to_affine(view, size: int = 16, new_size: int = 16) -> Tensor:
    """(B,6) view = (azimuth, elevation, scale, tx, ty, tz) -> (B,4,4) fp32 CPU tensor
    A = inverse(Tn @ (T @ S @ (Rz @ Ry)) @ Tc).

    Mirrors `Generator.transformation3d` / `apply_transformation`
    (reference core/models/hologan_generator.py:148-221).  Deliberately computed on the HOST with
    torch's own fp32 CPU ops in the reference's association order: the sampling coordinates must be
    bit-identical to the reference CPU path, and its cos/sin (SLEEF) and inverse (LAPACK) are not
    reproducible by an independent implementation.  It is 64 bytes per sample.
    """
    if isinstance(view, torch.Tensor):
        v = view.detach().to("cpu")
    else:
        v = torch.as_tensor(np.asarray(view))
    if v.dim() != 2 or v.shape[1] != 6:
        raise ValueError(f"view must have shape (B, 6), got {tuple(v.shape)}")
    v = v.float()
    n = v.shape[0]
    az, el, sc = v[:, 0], v[:, 1], v[:, 2]

    r_az = _homogeneous(n)
    r_az[:, 0, 0], r_az[:, 0, 1] = az.cos(), az.sin()
    r_az[:, 1, 0], r_az[:, 1, 1] = -az.sin(), az.cos()
    r_el = _homogeneous(n)
    r_el[:, 0, 0], r_el[:, 0, 2] = el.cos(), el.sin()
    r_el[:, 2, 0], r_el[:, 2, 2] = -el.sin(), el.cos()
    rot = torch.matmul(r_az, r_el)

    s_mat = _homogeneous(n)
    s_mat[:, 0, 0] = s_mat[:, 1, 1] = s_mat[:, 2, 2] = sc
    t_mat = _homogeneous(n)
    t_mat[:, 0:3, 3] = v[:, 3:6]
    m = torch.matmul(torch.matmul(t_mat, s_mat), rot)

    to_origin = _homogeneous(n)
    to_origin[:, 0:3, 3] = -size * 0.5
    to_new = _homogeneous(n)
    to_new[:, 0:3, 3] = new_size * 0.5
    return torch.linalg.inv(torch.matmul(torch.matmul(to_new, m), to_origin)).contiguous()


# ------------------------------------------------------------------------------------------------
# a6 + a7: rotate + resample
# ------------------------------------------------------------------------------------------------

def rotate_fwd_raw(vol: Tensor, a_inv: Tensor, border: int = HG_BORDER_REFERENCE, in_layout: int = HG_NCDHW,
                   out_layout: int = HG_NCDHW, debug: bool = False):
    """Direct call of hg_rotate_fwd.  vol (B,C,S,S,S) [NCDHW] ; a_inv (B,4,4) fp32 on the same device."""
    _require_cuda(vol, a_inv)
    if not vol.is_contiguous():
        raise ValueError("vol must be contiguous")
    if a_inv.dtype != torch.float32 or tuple(a_inv.shape) != (vol.shape[0], 4, 4) or not a_inv.is_contiguous():
        raise ValueError("a_inv must be a contiguous fp32 (B,4,4) tensor")
    if in_layout == HG_NCDHW:
        b, c, s = vol.shape[0], vol.shape[1], vol.shape[2]
        cubic = vol.shape[2] == vol.shape[3] == vol.shape[4]
    else:
        b, s, c = vol.shape[0], vol.shape[1], vol.shape[4]
        cubic = vol.shape[1] == vol.shape[2] == vol.shape[3]
    if vol.dim() != 5 or not cubic:
        raise ValueError(f"vol must be a cubic 5-D volume, got {tuple(vol.shape)}")
    if out_layout == HG_NCDHW:
        out = torch.empty((b, c, s, s, s), dtype=vol.dtype, device=vol.device)
    elif out_layout == HG_NDHWC:
        out = torch.empty((b, s, s, s, c), dtype=vol.dtype, device=vol.device)
    else:
        out = torch.empty((b, s, s, s, c), dtype=vol.dtype, device=vol.device)       # [b, z, x, y, c]
    coords = idx = None
    if debug:
        coords = torch.empty((3, b, s ** 3), dtype=torch.float32, device=vol.device)
        idx = torch.empty((8, b, s ** 3), dtype=torch.int32, device=vol.device)
    _lib.call("hg_rotate_fwd", _ptr(vol), _ptr(a_inv), _ptr(out), _ptr(coords), _ptr(idx), b, c, s, in_layout,
              out_layout, _dtype_code(vol), border, _stream())
    return (out, coords, idx) if debug else out


def rotate_bwd_raw(grad_out: Tensor, a_inv: Tensor, channels: int, size: int, border: int = HG_BORDER_REFERENCE,
                   in_layout: int = HG_NCDHW, out_layout: int = HG_NCDHW) -> Tensor:
    _require_cuda(grad_out, a_inv)
    grad_out = grad_out.contiguous()
    b, c, s = grad_out.shape[0], channels, size
    shape = (b, c, s, s, s) if in_layout == HG_NCDHW else (b, s, s, s, c)
    grad_vol = torch.empty(shape, dtype=grad_out.dtype, device=grad_out.device)
    ws_bytes = _lib.load().hg_rotate_bwd_workspace_bytes(b, s, in_layout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grad_out.device) if ws_bytes else None
    _lib.call("hg_rotate_bwd", _ptr(grad_out), _ptr(a_inv), _ptr(grad_vol), _ptr(ws), ws_bytes, b, c, s, in_layout,
              out_layout, _dtype_code(grad_out), border, _stream())
    return grad_vol


class _RotateResample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, a_inv, border, in_layout, out_layout):
        ctx.save_for_backward(a_inv)
        ctx.meta = (border, in_layout, out_layout,
                    vol.shape[1] if in_layout == HG_NCDHW else vol.shape[4], vol.shape[2])
        ctx.set_materialize_grads(True)
        return rotate_fwd_raw(vol, a_inv, border, in_layout, out_layout)

    @staticmethod
    def backward(ctx, grad_out):
        (a_inv,) = ctx.saved_tensors
        border, in_layout, out_layout, c, s = ctx.meta
        return rotate_bwd_raw(grad_out, a_inv, c, s, border, in_layout, out_layout), None, None, None, None


def rotate_resample(vol: Tensor, a_inv: Tensor, border: int = HG_BORDER_REFERENCE, in_layout: int = HG_NCDHW,
                    out_layout: int = HG_NCDHW) -> Tensor:
    """Differentiable (w.r.t. `vol`) rigid-body rotate + trilinear resample: the registered custom op
    `torch.ops.hologan.rotate_resample` (lightning_gan_zoo_b200/torch_ops.py)."""
    from . import torch_ops  # noqa: F401  (registers the hologan:: operators on first use)
    return torch.ops.hologan.rotate_resample(vol, a_inv, int(border), int(in_layout), int(out_layout))


# ------------------------------------------------------------------------------------------------
# a1 + a3: AdaIN (+ activation)
# ------------------------------------------------------------------------------------------------

def _adain_dims(x: Tensor, scale: Tensor) -> Tuple[int, int, int, int]:
    batch = scale.shape[0]
    c = x.shape[1]
    n = 1
    for d in x.shape[2:]:
        n *= d
    if x.shape[0] == batch:
        xbs = c * n
    elif x.shape[0] == 1:
        xbs = 0                                   # batch-shared constant (reference :121 `repeat`)
    else:
        raise ValueError("x batch must equal the style batch or be 1")
    return batch, c, n, xbs


def _style_stride(scale: Tensor, bias: Tensor) -> int:
    """scale/bias may be the two halves of one (B, 2C) ZMapping output; both must share a row stride."""
    if scale.dtype != torch.float32 or bias.dtype != torch.float32:
        raise TypeError("scale / bias must be float32")
    if scale.dim() != 2 or scale.shape != bias.shape or scale.stride(1) != 1 or bias.stride(1) != 1 \
            or scale.stride(0) != bias.stride(0):
        raise ValueError("scale / bias must be (B, C) row-major views with equal row stride")
    return scale.stride(0)


def _style_ptrs(scale: Tensor, bias: Optional[Tensor], channels: int):
    """(scale pointer, bias pointer, row stride).  `bias is None` means `scale` is a packed ZMapping output
    (B, 2C) = [scale | bias] (reference hologan_generator.py:17-18 slices it): both halves are read in place."""
    if bias is not None:
        return _ptr(scale), _ptr(bias), _style_stride(scale, bias)
    if scale.dtype != torch.float32 or scale.dim() != 2 or scale.shape[1] != 2 * channels or not scale.is_contiguous():
        raise ValueError("packed style must be a contiguous fp32 (B, 2C) tensor")
    return _ptr(scale), ctypes.c_void_p(scale.data_ptr() + 4 * channels), 2 * channels


def _dstyle_ptrs(packed: bool, b: int, c: int, device):
    """Gradient buffers for the styles: one (B, 2C) tensor for a packed style, else a (2, B, C) pair."""
    if packed:
        d = torch.empty((b, 2 * c), dtype=torch.float32, device=device)
        return d, _ptr(d), ctypes.c_void_p(d.data_ptr() + 4 * c), 2 * c
    d = torch.empty((2, b, c), dtype=torch.float32, device=device)
    return d, _ptr(d[0]), _ptr(d[1]), c


class _AdaInAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, bias, neg_slope, eps, biased):
        _require_cuda(x, scale, bias)
        if not x.is_contiguous():
            x = x.contiguous()
        b, c, n, xbs = _adain_dims(x, scale)
        sp, bp, sbs = _style_ptrs(scale, bias, c)
        y = torch.empty((b,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        mean = torch.empty((b, c), dtype=torch.float32, device=x.device)
        rstd = torch.empty((b, c), dtype=torch.float32, device=x.device)
        _lib.call("hg_adain_act_fwd", _ptr(x), sp, bp, _ptr(y), _ptr(mean), _ptr(rstd), b, c, n, xbs,
                  sbs, float(eps), float(neg_slope), int(biased), _dtype_code(x), _stream())
        ctx.save_for_backward(x, scale, bias, mean, rstd)
        ctx.meta = (b, c, n, xbs, float(neg_slope), int(biased))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale, bias, mean, rstd = ctx.saved_tensors
        b, c, n, xbs, neg_slope, biased = ctx.meta
        sp, bp, sbs = _style_ptrs(scale, bias, c)
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dsb, dsp, dbp, dstride = _dstyle_ptrs(bias is None, b, c, x.device)
        _lib.call("hg_adain_act_bwd", _ptr(x), _ptr(dy), sp, bp, _ptr(mean), _ptr(rstd), _ptr(dx),
                  dsp, dbp, b, c, n, xbs, sbs, dstride, neg_slope, biased, _dtype_code(x), _stream())
        if bias is None:
            return dx, dsb, None, None, None, None
        return dx, dsb[0], dsb[1], None, None, None


def adain_act(x: Tensor, scale: Tensor, bias: Optional[Tensor], neg_slope: float = 0.0, eps: float = 1e-8,
              biased_var: bool = False) -> Tensor:
    """act(AdaIn(x, scale, bias)); `neg_slope=1.0` gives the plain AdaIn of the reference
    (core/models/hologan_generator.py:333-345), `0.0` fuses the ReLU of :41 / :124.
    x may have batch 1 (the learned constant): it is broadcast without being materialised.
    `bias=None`: `scale` is the packed (B, 2C) ZMapping output [scale | bias] -- no slices, one style gradient."""
    return _AdaInAct.apply(x, scale, bias, neg_slope, eps, biased_var)


def instance_norm_act(x: Tensor, neg_slope: float = 0.2, eps: float = 1e-5) -> Tensor:
    """InstanceNorm2d (no affine, biased variance) fused with LeakyReLU -- the discriminator's
    norm + activation (reference core/models/hologan_discriminator.py:16-17,21-22) on the AdaIN kernel."""
    b, c = x.shape[0], x.shape[1]
    ones = torch.ones((b, c), dtype=torch.float32, device=x.device)
    zeros = torch.zeros((b, c), dtype=torch.float32, device=x.device)
    return _AdaInAct.apply(x, ones, zeros, neg_slope, eps, True)


# ------------------------------------------------------------------------------------------------
# a4 / a9 / a10: transposed convolutions on the tcgen05 implicit-GEMM kernels
# ------------------------------------------------------------------------------------------------

def convt_supported(cin: int, cout: int) -> bool:
    """Shapes the tcgen05 path covers (fwd + dgrad + wgrad): Cin % 128 == 0 and Cout % 64 == 0."""
    return cin % 128 == 0 and cout % 64 == 0


# ---- packed bf16 operand copies of a parameter, refreshed only when the parameter changes -------------------------
# The GEMM kernels read bf16 copies of the fp32 master weights in their own layouts.  For an nn.Parameter the copies are
# cached on the parameter object and re-packed IN PLACE (same buffers: captured CUDA graphs keep pointing at them) when
# its version counter / storage changes, i.e. after an optimizer step -- not on every forward.  Code that updates
# parameters behind autograd's back (a replayed CUDA graph that contains the optimizer) calls refresh_packed_weights()
# right after the update, inside the same graph.

def _cached_pack(weight: Tensor, tag, alloc, fill):
    if not isinstance(weight, torch.nn.Parameter):
        bufs = alloc()
        fill(bufs)
        return bufs
    cache = weight.__dict__.setdefault("_hg_pack", {})
    ver = (weight._version, weight.data_ptr())
    ent = cache.get(tag)
    if ent is None or ent[1][0].device != weight.device:
        bufs = alloc()
        fill(bufs)
        cache[tag] = [ver, bufs, fill]
        return bufs
    if ent[0] != ver:
        fill(ent[1])
        ent[0] = ver
    return ent[1]


def refresh_packed_weights(params) -> None:
    """Re-pack (in place) every cached operand copy of the given parameters -- call right after an optimizer step."""
    for p in params:
        cache = getattr(p, "_hg_pack", None)
        if cache:
            for ent in cache.values():
                ent[2](ent[1])
                ent[0] = (p._version, p.data_ptr())


def pack_convt_weight(weight: Tensor, perm: Tuple[int, int] = (0, 0)) -> Tuple[Tensor, Tensor]:
    """torch ConvTranspose weight (Cin, Cout, k..) fp32 -> (w_fwd [t][Cout][Cin], w_dgrad [t][Cin][Cout]) bf16.
    perm = (C, S): input channels re-ordered for the HG_PROJ operand (packed y*C + c <-> torch c*S + S-1-y)."""
    _require_cuda(weight)
    cin, cout = weight.shape[0], weight.shape[1]
    taps = weight[0, 0].numel()

    def alloc():
        return (torch.empty((taps, cout, cin), dtype=torch.bfloat16, device=weight.device),
                torch.empty((taps, cin, cout), dtype=torch.bfloat16, device=weight.device))

    def fill(bufs):
        w = weight.detach().float().contiguous()
        _lib.call("hg_convt_pack_weight", _ptr(w), _ptr(bufs[0]), _ptr(bufs[1]), cin, cout, taps, perm[0], perm[1], _stream())

    return _cached_pack(weight, ("convt", perm), alloc, fill)


def convt_wgrad(x_cl: Tensor, dy_s2d: Tensor, wshape, ndim: int, kernel: int, perm: Tuple[int, int] = (0, 0),
                accumulate_into: Optional[Tensor] = None, overwrite: bool = False) -> Optional[Tensor]:
    """Weight gradient in the torch parameter layout (fp32).  With `accumulate_into` (a contiguous fp32 tensor of
    the parameter's shape, e.g. its live .grad) the result is ADDED there (or, with `overwrite`, stored there) and
    None is returned."""
    b, size, cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[-1]
    cout = dy_s2d.shape[-1]
    nbytes = _lib.load().hg_convt_wgrad_workspace_bytes(b, cin, cout, ndim, size, kernel)
    if nbytes < 0:
        raise _lib.HologanB200Error(f"hg_convt_wgrad: unsupported shape Cin={cin} Cout={cout} size={size}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_cl.device)
    if accumulate_into is not None:
        dw, acc = accumulate_into, (0 if overwrite else 1)
    else:
        dw, acc = torch.empty(tuple(wshape), dtype=torch.float32, device=x_cl.device), 0
    _lib.call("hg_convt_wgrad", _ptr(x_cl), _ptr(dy_s2d), _ptr(dw), _ptr(ws), nbytes, b, cin, cout, ndim, size, kernel,
              perm[0], perm[1], acc, _stream())
    return None if accumulate_into is not None else dw


def act_bwd_bias(y: Tensor, dy: Tensor, neg_slope: float, want_bias: bool) -> Tuple[Tensor, Optional[Tensor]]:
    """(dpre, dbias): gradient through `y = act(pre + bias)` in one pass over the bf16 tensors (last dim = columns)."""
    _require_cuda(y, dy)
    if y.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16 or not y.is_contiguous() or not dy.is_contiguous():
        raise ValueError("y / dy must be contiguous bf16 tensors")
    cols = y.shape[-1]
    rows = y.numel() // cols
    dpre = torch.empty_like(dy)
    db = ws = None
    nbytes = 0
    if want_bias:
        nbytes = _lib.load().hg_act_bwd_bias_workspace_bytes(rows, cols)
        if nbytes < 0:
            raise _lib.HologanB200Error(f"hg_act_bwd_bias: unsupported shape rows={rows} cols={cols}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=y.device)
        db = torch.empty(cols, dtype=torch.float32, device=y.device)
    _lib.call("hg_act_bwd_bias", _ptr(y), _ptr(dy), _ptr(dpre), _ptr(db), _ptr(ws), nbytes, rows, cols,
              ctypes.c_float(neg_slope), _stream())
    return dpre, db


def _direct_grad_target(param, shape) -> Tuple[Optional[Tensor], bool]:
    """(buffer, overwrite): where the wgrad kernel may put the parameter's gradient itself, if the owner opted in.
    `param._hg_direct_grad = True`: accumulate into the live .grad (autograd's extra read-modify-write pass over the
    gradient is skipped).  `param._hg_direct_grad = <tensor>` (HologanTrainer: a view of its flat gradient buffer,
    one backward per step): STORE the gradient there -- the buffer then needs no zero-fill either."""
    mode = None if param is None else getattr(param, "_hg_direct_grad", None)
    if mode is None or mode is False:
        return None, False
    g, overwrite = (param.grad, False) if mode is True else (mode, True)
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or tuple(g.shape) != tuple(shape):
        return None, False
    return g, overwrite


# Side streams for work that runs BESIDE the backward's critical path.  "wgrad": the weight-gradient GEMMs (only the
# optimizer needs their result).  "comm": the data-parallel bucket exchange and the bucket-wise optimizer update
# (training._FlatGrads), which waits for the wgrad stream before it touches a bucket -- two streams, so that a wgrad never
# queues behind a collective.
_SIDE_STREAMS = {}
_SIDE_PENDING = set()
WGRAD_SIDE_STREAM = os.environ.get("HG_WGRAD_SIDE", "1") not in ("", "0")     # A/B switch (profiles/r02I_wgrad_side.txt)


def _dev_key(device) -> int:
    dev = torch.device(device)
    return dev.index if dev.index is not None else torch.cuda.current_device()


def side_stream(device, which: str = "wgrad") -> "torch.cuda.Stream":
    key = (_dev_key(device), which)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key[0])
    return st


def side_stream_mark(device, which: str = "wgrad") -> None:
    _SIDE_PENDING.add((_dev_key(device), which))


def join_side_stream(device, which: str = "wgrad", into=None) -> None:
    """Make `into` (default: the current stream) wait for everything enqueued on the side stream so far (no-op if the side
    stream has not been used since the last join into the current stream)."""
    if torch.device(device).type != "cuda" or not _SIDE_PENDING:
        return
    key = (_dev_key(device), which)
    if key in _SIDE_PENDING:
        (into if into is not None else torch.cuda.current_stream(key[0])).wait_stream(_SIDE_STREAMS[key])
        if into is None:
            _SIDE_PENDING.discard(key)


def _grad_ready(param) -> None:
    """A parameter's gradient has just been written into its flat-buffer view: run the owner's hook, if any
    (HologanTrainer starts the bucket's data-parallel all-reduce from it, overlapping the rest of the backward)."""
    cb = None if param is None else getattr(param, "_hg_grad_ready", None)
    if cb is not None:
        cb()


def _conv_dims(x_cl: Tensor, ndim: int):
    if x_cl.dtype != torch.bfloat16 or not x_cl.is_contiguous() or x_cl.dim() != ndim + 2:
        raise ValueError("activations must be contiguous channels-last bf16 tensors (B, [S,] S, S, C)")
    return x_cl.shape[0], x_cl.shape[1], x_cl.shape[-1]


class _ConvT(torch.autograd.Function):
    """y_s2d = act(convT(x) + bias) with x (B,[S,]S,S,Cin) bf16 and y_s2d (B,[S,]S,S,P,Cout) bf16."""

    @staticmethod
    def forward(ctx, x_cl, weight, bias, ndim, kernel, neg_slope, perm, want_stats):
        _require_cuda(x_cl, weight)
        b, size, cin = _conv_dims(x_cl, ndim)
        cout = weight.shape[1]
        if weight.shape[0] != cin:
            raise ValueError("weight / activation channel mismatch")
        wf, wd = pack_convt_weight(weight, perm)
        nclass = 1 if kernel == 1 else 2 ** ndim
        y = torch.empty((b,) + (size,) * ndim + (nclass, cout), dtype=torch.bfloat16, device=x_cl.device)
        bias_f = None if bias is None else bias.detach().float().contiguous()
        stats = None
        if want_stats:          # AdaIN statistics partials from the GEMM epilogue (consumed by adain_act_channels_last)
            n = _lib.load().hg_convt_stats_floats(b, cout, ndim, size, kernel)
            if n < 0:
                raise _lib.HologanB200Error("hg_convt_stats_floats: unsupported shape")
            stats = torch.empty(n, dtype=torch.float32, device=x_cl.device)
            _lib.call("hg_convt_fwd_stats", _ptr(x_cl), _ptr(wf), _ptr(bias_f), _ptr(y), _ptr(stats), b, cin, cout, ndim, size,
                      kernel, ctypes.c_float(neg_slope), _stream())
        else:
            _lib.call("hg_convt_fwd", _ptr(x_cl), _ptr(wf), _ptr(bias_f), _ptr(y), b, cin, cout, ndim, size, kernel,
                      ctypes.c_float(neg_slope), _stream())
        ctx.save_for_backward(x_cl, wd, y if neg_slope != 1.0 else None)
        ctx.meta = (b, cin, cout, ndim, size, kernel, neg_slope, tuple(weight.shape), bias is not None, perm)
        ctx.weight_param = weight if isinstance(weight, torch.nn.Parameter) else None
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y

    @staticmethod
    def backward(ctx, dy, *_unused):
        x_cl, wd, y = ctx.saved_tensors
        b, cin, cout, ndim, size, kernel, neg_slope, wshape, has_bias, perm = ctx.meta
        dy = dy.contiguous()
        dx = dw = db = None
        want_db = has_bias and ctx.needs_input_grad[2]
        nclass = dy.shape[-2]
        if y is not None:                      # activation (+ bias) fused in the forward epilogue: one pass
            if nclass == 1 or not want_db:
                dy, db = act_bwd_bias(y, dy, neg_slope, want_db)
            else:
                dy, _ = act_bwd_bias(y, dy, neg_slope, False)
        if want_db and db is None:
            db = dy.reshape(-1, nclass, cout).float().sum(dim=(0, 1))
        target, overwrite = _direct_grad_target(ctx.weight_param, wshape) if ctx.needs_input_grad[1] else (None, False)
        on_side = WGRAD_SIDE_STREAM and target is not None and ctx.needs_input_grad[0]
        if on_side:
            # The weight gradient goes straight into the owner's flat buffer and only the optimizer reads it: fork it onto
            # the side stream BEFORE the dgrad is enqueued, so the two GEMMs (and whatever the backward launches next) run
            # concurrently.  The owner joins the side stream before its optimizer step (training._FlatGrads.finish).
            cur, side = torch.cuda.current_stream(dy.device), side_stream(dy.device)
            side.wait_stream(cur)                       # dy is final on the current stream
            with torch.cuda.stream(side):
                convt_wgrad(x_cl, dy, wshape, ndim, kernel, perm, accumulate_into=target, overwrite=overwrite)
            x_cl.record_stream(side)
            dy.record_stream(side)
            side_stream_mark(dy.device)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x_cl)
            _lib.call("hg_convt_dgrad", _ptr(dy), _ptr(wd), _ptr(dx), b, cin, cout, ndim, size, kernel, _stream())
        if ctx.needs_input_grad[1]:
            if not on_side:
                dw = convt_wgrad(x_cl, dy, wshape, ndim, kernel, perm, accumulate_into=target, overwrite=overwrite)
            _grad_ready(ctx.weight_param)
        return dx, dw, db, None, None, None, None, None


def convt(x_cl: Tensor, weight: Tensor, bias: Optional[Tensor], ndim: int, kernel: int, neg_slope: float = 1.0,
          perm: Tuple[int, int] = (0, 0), stats: bool = False):
    """Transposed convolution / 1x1 projection on the tcgen05 tap GEMM.  `stats=True` returns (y_s2d, stats): the AdaIN
    statistics partials of y computed in the GEMM epilogue -- pass them to `adain_act_channels_last(..., stats=stats)`."""
    return _ConvT.apply(x_cl, weight, bias, ndim, kernel, neg_slope, perm, stats)


def convt_stats_supported(cout: int, ndim: int, size: int) -> bool:
    return cout % 32 == 0 and (size ** ndim) % 32 == 0


# ---- layout glue (pure data movement) -------------------------------------------------------------

def s2d_to_nc(y_s2d: Tensor, ndim: int) -> Tensor:
    """(B,[S,]S,S,P,C) space-to-depth conv output -> torch layout (B,C,2S,[2S,]2S)."""
    b, s, c = y_s2d.shape[0], y_s2d.shape[1], y_s2d.shape[-1]
    if ndim == 2:
        return y_s2d.reshape(b, s, s, 2, 2, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, 2 * s, 2 * s)
    return y_s2d.reshape(b, s, s, s, 2, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, c, 2 * s, 2 * s, 2 * s)


def nc_to_channels_last(x: Tensor) -> Tensor:
    """(B,C,*sp) -> contiguous (B,*sp,C)."""
    return x.permute(0, *range(2, x.dim()), 1).contiguous()


# ------------------------------------------------------------------------------------------------
# channels-last AdaIN (+activation) on conv outputs in space-to-depth layout
# ------------------------------------------------------------------------------------------------

class _AdaInChannelsLast(torch.autograd.Function):
    """x (B,[S,]S,S,P,C) bf16 (s2d conv output, P = 2^ndim) or (B,[S,]S,S,C) (P = 1)  ->
    y (B,[2S,]2S,2S,C) bf16 channels-last.  scale / bias None = 1 / 0 (no style gradients)."""

    @staticmethod
    def forward(ctx, x, scale, bias, ndim, classes, neg_slope, eps, biased, stats=None):
        _require_cuda(x, scale, bias)
        if x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise ValueError("x must be a contiguous bf16 tensor")
        b, size, c = x.shape[0], x.shape[1], x.shape[-1]
        sp, bp, sbs = _style_ptrs(scale, bias, c) if scale is not None else (_ptr(None), _ptr(None), 0)
        up = 2 if classes > 1 else 1
        if classes == -4:       # plain (B, S, S, C) in, 2x2 space-to-depth order out: the next stride-2 conv's operand
            y = torch.empty((b, size // 2, size // 2, 4, c), dtype=torch.bfloat16, device=x.device)
        else:
            y = torch.empty((b,) + (up * size,) * ndim + (c,), dtype=torch.bfloat16, device=x.device)
        mean = torch.empty((b, c), dtype=torch.float32, device=x.device)
        rstd = torch.empty((b, c), dtype=torch.float32, device=x.device)
        nbytes = _lib.load().hg_adain_cl_workspace_bytes(b, c, ndim, size, classes)
        if nbytes < 0:
            raise _lib.HologanB200Error(f"hg_adain_cl_fwd: unsupported shape C={c} size={size} classes={classes}")
        if stats is not None:       # statistics partials from the producing GEMM's epilogue: merge + one streaming pass
            _lib.call("hg_adain_cl_fwd_stats", _ptr(x), _ptr(stats), sp, bp, _ptr(y), _ptr(mean), _ptr(rstd), b, c, ndim, size,
                      classes, sbs, ctypes.c_float(eps), ctypes.c_float(neg_slope), int(biased), _stream())
        else:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
            _lib.call("hg_adain_cl_fwd", _ptr(x), sp, bp, _ptr(y), _ptr(mean), _ptr(rstd), _ptr(ws), nbytes, b, c,
                      ndim, size, classes, sbs, ctypes.c_float(eps), ctypes.c_float(neg_slope), int(biased), _stream())
        ctx.save_for_backward(x, scale, bias, mean, rstd)
        ctx.meta = (b, c, ndim, size, classes, float(neg_slope), int(biased))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale, bias, mean, rstd = ctx.saved_tensors
        b, c, ndim, size, classes, neg_slope, biased = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        if scale is not None:
            sp, bp, sbs = _style_ptrs(scale, bias, c)
            dsb, dsp, dbp, dstride = _dstyle_ptrs(bias is None, b, c, x.device)
        else:
            sp = bp = dsp = dbp = _ptr(None)
            dsb, sbs, dstride = None, 0, c
        nbytes = _lib.load().hg_adain_cl_workspace_bytes(b, c, ndim, size, classes)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
        _lib.call("hg_adain_cl_bwd", _ptr(x), _ptr(dy), sp, bp, _ptr(mean), _ptr(rstd), _ptr(dx),
                  dsp, dbp, _ptr(ws), nbytes, b, c, ndim,
                  size, classes, sbs, dstride, ctypes.c_float(neg_slope), biased, _stream())
        if dsb is None:
            return dx, None, None, None, None, None, None, None, None
        if bias is None:
            return dx, dsb, None, None, None, None, None, None, None
        return dx, dsb[0], dsb[1], None, None, None, None, None, None


def adain_act_channels_last(x: Tensor, scale: Tensor, bias: Optional[Tensor], ndim: int, classes: int,
                            neg_slope: float = 0.0, eps: float = 1e-8, stats: Optional[Tensor] = None) -> Tensor:
    """`bias=None`: `scale` is the packed (B, 2C) style [scale | bias] (see adain_act).
    `stats`: the statistics partials `convt(..., stats=True)` produced for x in its GEMM epilogue."""
    return _AdaInChannelsLast.apply(x, scale, bias, ndim, classes, neg_slope, eps, False, stats)


def instance_norm_act_channels_last(x_nhwc: Tensor, neg_slope: float = 0.2, eps: float = 1e-5, s2d_out: bool = False) -> Tensor:
    """InstanceNorm2d (no affine, biased variance) + LeakyReLU on a channels-last (B,H,W,C) bf16 tensor, H == W
    -- the discriminator's norm + activation (reference core/models/hologan_discriminator.py:16-17,21-22).
    `s2d_out`: the result is stored as (B, H/2, W/2, 4, C) (2x2 space-to-depth order), the input layout of `conv5s2_sn`."""
    return _AdaInChannelsLast.apply(x_nhwc, None, None, 2, -4 if s2d_out else 1, neg_slope, eps, True)


# ------------------------------------------------------------------------------------------------
# a2: ZMapping linear + ReLU (fp32), a11: final conv + tanh
# ------------------------------------------------------------------------------------------------

class _LinearRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, weight, bias):
        _require_cuda(z, weight, bias)
        z = z.float().contiguous()
        w = weight.float().contiguous()
        bb = bias.float().contiguous()
        b, k = z.shape
        n = w.shape[0]
        out = torch.empty((b, n), dtype=torch.float32, device=z.device)
        _lib.call("hg_linear_relu_fwd", _ptr(z), _ptr(w), _ptr(bb), _ptr(out), b, k, n, _stream())
        ctx.save_for_backward(z, w, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, w, out = ctx.saved_tensors
        b, k = z.shape
        n = w.shape[0]
        dout = dout.float().contiguous()
        dw = torch.empty_like(w)
        db = torch.empty(n, dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        _lib.call("hg_linear_relu_bwd", _ptr(z), _ptr(w), _ptr(out), _ptr(dout), _ptr(dw), _ptr(db), _ptr(dz), b, k, n, 0,
                  _stream())
        return dz, dw, db


def linear_relu(z: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """relu(z @ weight.T + bias) in fp32 -- the ZMapping of reference hologan_generator.py:16-17."""
    return _LinearRelu.apply(z, weight, bias)


class _LinearReluGroup(torch.autograd.Function):
    """styles_l = relu(z @ W_l^T + b_l) for several ZMappings of the same z: one launch forward, one backward."""

    @staticmethod
    def forward(ctx, z, n_layers, *wb):
        weights, biases = wb[:n_layers], wb[n_layers:]
        _require_cuda(z, *weights, *biases)
        z = z.float().contiguous()
        ws = [w.float().contiguous() for w in weights]
        bs = [b.float().contiguous() for b in biases]
        b, k = z.shape
        ns = [w.shape[0] for w in ws]
        if any(w.shape[1] != k for w in ws):
            raise ValueError("linear_relu_group: every weight must be (N_l, K) with K = z.shape[1]")
        outs = [torch.empty((b, n), dtype=torch.float32, device=z.device) for n in ns]
        vp = ctypes.c_void_p
        _lib.call("hg_linear_relu_group_fwd", n_layers, _ptr(z), _c_array(vp, [w.data_ptr() for w in ws]),
                  _c_array(vp, [t.data_ptr() for t in bs]), _c_array(vp, [o.data_ptr() for o in outs]),
                  _c_array(ctypes.c_int, ns), b, k, _stream())
        ctx.save_for_backward(z, *outs)
        ctx.meta = (n_layers, ns, k)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        z, outs = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        n_layers, ns, k = ctx.meta
        douts = [d.float().contiguous() for d in douts]
        dws = [torch.empty((n, k), dtype=torch.float32, device=z.device) for n in ns]
        dbs = [torch.empty(n, dtype=torch.float32, device=z.device) for n in ns]
        vp = ctypes.c_void_p
        _lib.call("hg_linear_relu_group_bwd", n_layers, _ptr(z), _c_array(vp, [o.data_ptr() for o in outs]),
                  _c_array(vp, [d.data_ptr() for d in douts]), _c_array(vp, [d.data_ptr() for d in dws]),
                  _c_array(vp, [d.data_ptr() for d in dbs]), _c_array(ctypes.c_int, ns), z.shape[0], k, _stream())
        return (None, None) + tuple(dws) + tuple(dbs)


def linear_relu_group(z: Tensor, weights, biases):
    """[relu(z @ W^T + b) for W, b in zip(weights, biases)] -- the generator's five ZMappings of one latent batch
    (reference hologan_generator.py:34,54) in a single launch.  `z` gets no gradient (use `linear_relu` for that)."""
    if z.requires_grad:
        return tuple(linear_relu(z, w, b) for w, b in zip(weights, biases))
    return _LinearReluGroup.apply(z, len(weights), *weights, *biases)


def final_conv_supported(cin: int, cout: int) -> bool:
    lanes = cin // 8
    return cin % 8 == 0 and 8 <= cin <= 256 and (lanes & (lanes - 1)) == 0 and 1 <= cout <= 4


class _FinalConvTanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, weight, bias):
        _require_cuda(x_cl, weight, bias)
        if x_cl.dtype != torch.bfloat16 or not x_cl.is_contiguous() or x_cl.dim() != 4:
            raise ValueError("x must be a contiguous (B,S,S,C) bf16 tensor")
        b, s, cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[3]
        w = weight.detach().float().contiguous()
        bb = bias.detach().float().contiguous()
        cout = w.shape[0]
        out = torch.empty((b, cout, s, s), dtype=torch.float32, device=x_cl.device)
        _lib.call("hg_final_conv_tanh_fwd", _ptr(x_cl), _ptr(w), _ptr(bb), _ptr(out), b, cin, cout, s, _stream())
        ctx.save_for_backward(x_cl, w, out)
        ctx.params = (weight if isinstance(weight, torch.nn.Parameter) else None, bias if isinstance(bias, torch.nn.Parameter) else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_cl, w, out = ctx.saved_tensors
        b, s, cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[3]
        cout = w.shape[0]
        dout = dout.float().contiguous()
        dev = x_cl.device
        nbytes = _lib.load().hg_final_conv_tanh_bwd_workspace_bytes(b, cin, cout, s)
        dx = torch.empty_like(x_cl) if ctx.needs_input_grad[0] else None
        # weight / bias gradient straight into the owner's flat buffer on the wgrad side stream (see _ConvT.backward)
        tw, ow = _direct_grad_target(ctx.params[0], w.shape)
        tb, ob = _direct_grad_target(ctx.params[1], (cout,))
        if WGRAD_SIDE_STREAM and dx is not None and tw is not None and tb is not None and ow and ob:
            cur, side = torch.cuda.current_stream(dev), side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ws2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.call("hg_final_conv_tanh_bwd", _ptr(x_cl), _ptr(w), _ptr(out), _ptr(dout), _ptr(None), _ptr(tw), _ptr(tb),
                          _ptr(ws2), nbytes, b, cin, cout, s, _stream())
            for t in (x_cl, out, dout):
                t.record_stream(side)
            side_stream_mark(dev)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("hg_final_conv_tanh_bwd", _ptr(x_cl), _ptr(w), _ptr(out), _ptr(dout), _ptr(dx), _ptr(None), _ptr(None),
                      _ptr(ws), nbytes, b, cin, cout, s, _stream())
            return dx, None, None
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dw = torch.empty_like(w)
        db = torch.empty(cout, dtype=torch.float32, device=dev)
        _lib.call("hg_final_conv_tanh_bwd", _ptr(x_cl), _ptr(w), _ptr(out), _ptr(dout), _ptr(dx), _ptr(dw), _ptr(db),
                  _ptr(ws), nbytes, b, cin, cout, s, _stream())
        return dx, dw, db


def final_conv_tanh(x_cl: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """tanh(conv2d(x, weight, bias, kernel 3, padding 1)) with x (B,S,S,C) bf16 NHWC -> (B,Cout,S,S) fp32 NCHW."""
    return _FinalConvTanh.apply(x_cl, weight, bias)


class _Head128Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, weight, bias):
        _require_cuda(x_cl, weight, bias)
        if x_cl.dtype != torch.bfloat16 or not x_cl.is_contiguous() or x_cl.dim() != 4 or x_cl.shape[1] != x_cl.shape[2]:
            raise ValueError("x must be a contiguous (B,S,S,C) bf16 tensor")
        b, s2, cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[3]
        w = weight.detach().float().contiguous()
        bb = bias.detach().float().contiguous()
        cout = w.shape[1]
        if tuple(w.shape) != (cin, cout, 4, 4):
            raise ValueError("weight must be the ConvTranspose2d parameter (Cin, Cout, 4, 4)")
        out = torch.empty((b, cout, 2 * s2, 2 * s2), dtype=torch.float32, device=x_cl.device)
        _lib.call("hg_head128_fwd", _ptr(x_cl), _ptr(w), _ptr(bb), _ptr(out), b, cin, cout, 2 * s2, _stream())
        ctx.save_for_backward(x_cl, w, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_cl, w, out = ctx.saved_tensors
        b, s2, cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[3]
        cout = w.shape[1]
        dout = dout.float().contiguous()
        nbytes = _lib.load().hg_head128_bwd_workspace_bytes(b, 2 * s2)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x_cl.device)
        dx = torch.empty_like(x_cl) if ctx.needs_input_grad[0] else None
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dw = torch.empty_like(w) if want_w else None
        db = torch.empty(cout, dtype=torch.float32, device=x_cl.device) if want_w else None
        _lib.call("hg_head128_bwd", _ptr(x_cl), _ptr(w), _ptr(out), _ptr(dout), _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws), nbytes, b,
                  cin, cout, 2 * s2, _stream())
        return dx, dw, db


def head128_tanh(x_cl: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """tanh(conv_transpose2d(x, weight, bias, kernel 4, stride 2, padding 1)): the patched 128 x 128 head with
    x (B,S,S,64) bf16 NHWC -> (B,3,2S,2S) fp32 NCHW, weight the torch ConvTranspose2d parameter (64, 3, 4, 4)."""
    return _Head128Tanh.apply(x_cl, weight, bias)


def head128_supported(cin: int, cout: int, size_in: int) -> bool:
    return cin == 64 and cout == 3 and size_in % 16 == 0


# ------------------------------------------------------------------------------------------------
# a13: the losses of HOLOGAN.training_step
# ------------------------------------------------------------------------------------------------

class _GanLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, z_pred, z, ta, wa, tb, wb):
        _require_cuda(a, b, z_pred, z)
        a, z_pred = a.contiguous(), z_pred.contiguous()
        b = None if b is None else b.contiguous()
        z = z.float().contiguous()
        if z_pred.dtype != a.dtype or (b is not None and b.dtype != a.dtype) or z_pred.numel() != z.numel():
            raise ValueError("logits / z_pred must share a dtype; z_pred and z must have the same size")
        total = torch.empty((), dtype=torch.float32, device=a.device)
        parts = torch.empty(2, dtype=torch.float32, device=a.device)
        nb = 0 if b is None else b.numel()
        _lib.call("hg_gan_loss_fwd", _ptr(a), a.numel(), float(ta), float(wa), _ptr(b), nb, float(tb), float(wb),
                  _ptr(z_pred), _ptr(z), z.numel(), _dtype_code(a), _ptr(total), _ptr(parts), _stream())
        ctx.save_for_backward(a, b, z_pred, z)
        ctx.meta = (float(ta), float(wa), float(tb), float(wb))
        ctx.mark_non_differentiable(parts)
        ctx.set_materialize_grads(False)
        return total, parts

    @staticmethod
    def backward(ctx, gtotal, _gparts):
        a, b, z_pred, z = ctx.saved_tensors
        ta, wa, tb, wb = ctx.meta
        if gtotal is None:
            return (None,) * 8
        gtotal = gtotal.float().contiguous()
        da, dzp = torch.empty_like(a), torch.empty_like(z_pred)
        db = None if b is None else torch.empty_like(b)
        nb = 0 if b is None else b.numel()
        _lib.call("hg_gan_loss_bwd", _ptr(gtotal), _ptr(a), a.numel(), ta, wa, _ptr(b), nb, tb, wb, _ptr(z_pred), _ptr(z),
                  z.numel(), _dtype_code(a), _ptr(da), _ptr(db), _ptr(dzp), _stream())
        return da, db, dzp, None, None, None, None, None


def hologan_d_loss(d_real: Tensor, d_fake: Tensor, z_pred: Tensor, z: Tensor) -> Tuple[Tensor, Tensor]:
    """Discriminator-step loss of the reference (core/lightning_module.py:219-229):
    (BCE(D(real), 1) + BCE(D(fake), 0)) / 2 + mean((z_pred - z)^2).  Returns (total, [d_loss, q_loss])."""
    return _GanLoss.apply(d_real, d_fake, z_pred, z, 1.0, 0.5, 0.0, 0.5)


def hologan_g_loss(d_fake: Tensor, z_pred: Tensor, z: Tensor) -> Tuple[Tensor, Tensor]:
    """Generator-step loss (core/lightning_module.py:231-237): BCE(D(fake), 1) + mean((z_pred - z)^2).
    Returns (total, [g_loss, q_loss])."""
    return _GanLoss.apply(d_fake, None, z_pred, z, 1.0, 1.0, 0.0, 0.0)


# ------------------------------------------------------------------------------------------------
# a14: spectral normalisation of the discriminator's convolutions (all layers in one call)
# ------------------------------------------------------------------------------------------------

def _c_array(ctype, values):
    return (ctype * len(values))(*values)


def _sn_geometry(weights):
    w0 = weights[0]
    if all(w.is_contiguous() for w in weights):
        channels_last = 0
    elif all(w.dim() == 4 and w.is_contiguous(memory_format=torch.channels_last) for w in weights):
        channels_last = 1
    else:
        raise ValueError("spectral_norm_weights: weights must all be contiguous or all channels_last")
    for w in weights:
        if w.dtype != torch.float32 or w.dim() < 2 or w.device != w0.device:
            raise ValueError("spectral_norm_weights: fp32 weights of >= 2 dims on one device")
    cout = [w.shape[0] for w in weights]
    cin = [w.shape[1] for w in weights]
    taps = [w[0, 0].numel() for w in weights]
    return channels_last, cout, cin, taps


class _SpectralNormWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, us, vs, iterate, out_dtype, eps, *weights):
        _require_cuda(*weights, *us, *vs)
        n = len(weights)
        channels_last, cout, cin, taps = _sn_geometry(weights)
        for w, u, v, co, ci, t in zip(weights, us, vs, cout, cin, taps):
            if u.dtype != torch.float32 or v.dtype != torch.float32 or u.numel() != co or v.numel() != ci * t \
                    or not u.is_contiguous() or not v.is_contiguous():
                raise ValueError("spectral_norm_weights: u (Cout) / v (Cin*taps) must be contiguous fp32 buffers")
        lib = _lib.load()
        ci_arr, co_arr, t_arr = _c_array(ctypes.c_int, cin), _c_array(ctypes.c_int, cout), _c_array(ctypes.c_int, taps)
        nbytes = lib.hg_spectral_norm_workspace_bytes(n, co_arr, ci_arr, t_arr)
        if nbytes < 0:
            raise _lib.HologanB200Error(f"hg_spectral_norm_fwd: unsupported layer count {n}")
        dev = weights[0].device
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        outs = [torch.empty_like(w, dtype=out_dtype) for w in weights]           # same physical layout
        states = [torch.empty(lib.hg_spectral_norm_state_floats(co, ci, t), dtype=torch.float32, device=dev)
                  for co, ci, t in zip(cout, cin, taps)]
        vp = ctypes.c_void_p
        _lib.call("hg_spectral_norm_fwd", n, _c_array(vp, [w.data_ptr() for w in weights]),
                  _c_array(vp, [u.data_ptr() for u in us]), _c_array(vp, [v.data_ptr() for v in vs]),
                  _c_array(vp, [o.data_ptr() for o in outs]), _c_array(vp, [s.data_ptr() for s in states]),
                  co_arr, ci_arr, t_arr, channels_last, int(bool(iterate)), float(eps),
                  HG_BF16 if out_dtype == torch.bfloat16 else HG_F32, _ptr(ws), nbytes, _stream())
        ctx.save_for_backward(*weights, *states)
        ctx.meta = (n, channels_last, cout, cin, taps, out_dtype)
        ctx.params = [w if isinstance(w, torch.nn.Parameter) else None for w in weights]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        n, channels_last, cout, cin, taps, out_dtype = ctx.meta
        weights, states = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        grads = [g.to(out_dtype).contiguous(memory_format=fmt) for g in grads]
        # HologanTrainer's flat gradient buffer: accumulate straight into the parameter's view (the D step runs the
        # discriminator twice, so two backward calls add into it; the trainer zeroes these views before the step)
        targets = [_direct_grad_target(p, w.shape)[0] if p is not None else None for p, w in zip(ctx.params, weights)]
        direct = all(t is not None and t.stride() == w.stride() for t, w in zip(targets, weights))
        dws = targets if direct else [torch.empty_like(w) for w in weights]
        lib = _lib.load()
        ci_arr, co_arr, t_arr = _c_array(ctypes.c_int, cin), _c_array(ctypes.c_int, cout), _c_array(ctypes.c_int, taps)
        nbytes = lib.hg_spectral_norm_workspace_bytes(n, co_arr, ci_arr, t_arr)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=weights[0].device)
        vp = ctypes.c_void_p
        _lib.call("hg_spectral_norm_bwd", n, _c_array(vp, [g.data_ptr() for g in grads]),
                  _c_array(vp, [w.data_ptr() for w in weights]), _c_array(vp, [s.data_ptr() for s in states]),
                  _c_array(vp, [d.data_ptr() for d in dws]), co_arr, ci_arr, t_arr, int(direct),
                  HG_BF16 if out_dtype == torch.bfloat16 else HG_F32, _ptr(ws), nbytes, _stream())
        return (None, None, None, None, None) + ((None,) * n if direct else tuple(dws))


def spectral_norm_weights(weights, us, vs, power_iteration: bool = True, out_dtype=torch.float32, eps: float = 1e-12):
    """W_i / sigma_i for a group of (<= 4) weights, with one power iteration that updates `us[i]` / `vs[i]` in place
    when `power_iteration` -- torch.nn.utils.spectral_norm's training-mode forward as the reference uses it
    (core/models/hologan_discriminator.py:15).  Outputs have `out_dtype` and the weights' memory format."""
    return _SpectralNormWeights.apply(list(us), list(vs), power_iteration, out_dtype, eps, *weights)


# ------------------------------------------------------------------------------------------------
# a14 / f1: the discriminator on hand-written kernels (first conv, spectral-norm conv blocks, heads)
# ------------------------------------------------------------------------------------------------

class _DConv0(torch.autograd.Function):
    """y_s2d = leaky_relu(Conv2d(3 -> 64, k5, s2, p2)(x) + bias): x (B,3,S,S) fp32 NCHW -> (B,S/4,S/4,4,64) bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, neg_slope):
        _require_cuda(x, weight, bias)
        x = x.float().contiguous()
        w = weight.detach().float().contiguous()
        bb = bias.detach().float().contiguous()
        b, cin, s = x.shape[0], x.shape[1], x.shape[2]
        cout = w.shape[0]
        if x.dim() != 4 or x.shape[3] != s or tuple(w.shape[1:]) != (cin, 5, 5):
            raise ValueError("dconv0: x must be (B, Cin, S, S) and weight (Cout, Cin, 5, 5)")
        y = torch.empty((b, s // 4, s // 4, 4, cout), dtype=torch.bfloat16, device=x.device)
        _lib.call("hg_dconv0_fwd", _ptr(x), _ptr(w), _ptr(bb), _ptr(y), b, cin, cout, s, ctypes.c_float(neg_slope), _stream())
        ctx.save_for_backward(x, w, y)
        ctx.neg_slope = float(neg_slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        b, cin, s = x.shape[0], x.shape[1], x.shape[2]
        cout = w.shape[0]
        dy = dy.contiguous()
        want_dx, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if want_dx else None
        dw = torch.empty_like(w) if want_dw else None
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if want_dw else None
        nbytes = _lib.load().hg_dconv0_bwd_workspace_bytes(b, s)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.call("hg_dconv0_bwd", _ptr(x), _ptr(w), _ptr(y), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws), nbytes, b, cin,
                  cout, s, ctypes.c_float(ctx.neg_slope), 0, _stream())
        return dx, dw, db, None


def dconv0(x: Tensor, weight: Tensor, bias: Tensor, neg_slope: float = 0.2) -> Tensor:
    """First discriminator convolution + LeakyReLU (reference core/models/hologan_discriminator.py:30,58) on the
    warp-level tensor cores; the result is in the space-to-depth order `conv5s2_sn` reads."""
    return _DConv0.apply(x, weight, bias, neg_slope)


def dconv0_supported(cin: int, cout: int, size: int) -> bool:
    return cin == 3 and cout == 64 and size % 32 == 0


def spectral_norm_sigma(weights, us, vs, power_iteration: bool = True, eps: float = 1e-12):
    """Power iteration + sigma for a group of weights WITHOUT materialising W / sigma: returns one state tensor per
    layer ([0] sigma, [1] 1 / sigma, then the u and v used), consumed by `conv5s2_sn` (which folds 1 / sigma into its
    bf16 weight pack) and by its backward.  `us` / `vs` are updated in place when `power_iteration` (training-mode
    forward of torch.nn.utils.spectral_norm, reference core/models/hologan_discriminator.py:15).  No autograd."""
    with torch.no_grad():
        weights = [w.detach() for w in weights]
        _require_cuda(*weights, *us, *vs)
        n = len(weights)
        channels_last, cout, cin, taps = _sn_geometry(weights)
        if channels_last:
            raise ValueError("spectral_norm_sigma: contiguous weights only")
        lib = _lib.load()
        ci_arr, co_arr, t_arr = _c_array(ctypes.c_int, cin), _c_array(ctypes.c_int, cout), _c_array(ctypes.c_int, taps)
        nbytes = lib.hg_spectral_norm_workspace_bytes(n, co_arr, ci_arr, t_arr)
        dev = weights[0].device
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        states = [torch.empty(lib.hg_spectral_norm_state_floats(co, ci, t), dtype=torch.float32, device=dev)
                  for co, ci, t in zip(cout, cin, taps)]
        vp = ctypes.c_void_p
        _lib.call("hg_spectral_norm_fwd", n, _c_array(vp, [w.data_ptr() for w in weights]),
                  _c_array(vp, [u.data_ptr() for u in us]), _c_array(vp, [v.data_ptr() for v in vs]),
                  _c_array(vp, [0] * n), _c_array(vp, [s.data_ptr() for s in states]),
                  co_arr, ci_arr, t_arr, 0, int(bool(power_iteration)), float(eps), HG_F32, _ptr(ws), nbytes, _stream())
    return states


class _Conv5s2SN(torch.autograd.Function):
    """y = Conv2d(k5, s2, p2)(x, weight_orig / sigma) (no bias) on the tcgen05 tap GEMMs; x_s2d (B,S,S,4,Cin) bf16 ->
    y (B,S,S,Cout) bf16.  `state` comes from `spectral_norm_sigma` (None: plain weight).  The backward produces the
    gradient w.r.t. weight_orig (spectral-norm backward included)."""

    @staticmethod
    def forward(ctx, x_s2d, weight, state):
        _require_cuda(x_s2d, weight)
        if x_s2d.dtype != torch.bfloat16 or not x_s2d.is_contiguous() or x_s2d.dim() != 5 or x_s2d.shape[3] != 4:
            raise ValueError("x_s2d must be a contiguous bf16 tensor (B, S, S, 4, Cin)")
        b, size, cin = x_s2d.shape[0], x_s2d.shape[1], x_s2d.shape[4]
        w = weight.detach()
        cout = w.shape[0]
        if w.dtype != torch.float32 or not w.is_contiguous() or tuple(w.shape[1:]) != (cin, 5, 5):
            raise ValueError("weight must be a contiguous fp32 (Cout, Cin, 5, 5) tensor")
        dev = x_s2d.device
        # bf16 copies of the UN-normalised weight (cached: packed once per optimizer step); 1 / sigma, which moves with
        # every forward's power iteration, is applied in the GEMM epilogue: conv(x, W / sigma) = conv(x, W) / sigma
        wk, wt = _cached_pack(weight, "conv5s2",
                              lambda: (torch.empty((25, cout, cin), dtype=torch.bfloat16, device=dev),
                                       torch.empty((25, cin, cout), dtype=torch.bfloat16, device=dev)),
                              lambda bufs: _lib.call("hg_conv5s2_pack_weight", _ptr(weight.detach()), _ptr(None), _ptr(bufs[0]),
                                                     _ptr(bufs[1]), cin, cout, _stream()))
        inv_sigma = ctypes.c_void_p(0 if state is None else state.data_ptr() + 4)
        nbytes = _lib.load().hg_conv5s2_workspace_bytes(b, cin, cout, size)
        if nbytes < 0:
            raise _lib.HologanB200Error(f"hg_conv5s2: unsupported shape Cin={cin} Cout={cout} size={size}")
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        y = torch.empty((b, size, size, cout), dtype=torch.bfloat16, device=dev)
        _lib.call("hg_conv5s2_fwd", _ptr(x_s2d), _ptr(wk), inv_sigma, _ptr(y), _ptr(ws), nbytes, b, cin, cout, size, _stream())
        ctx.save_for_backward(x_s2d, wt, w, state)
        ctx.meta = (b, size, cin, cout, nbytes)
        ctx.weight_param = weight if isinstance(weight, torch.nn.Parameter) else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x_s2d, wt, w, state = ctx.saved_tensors
        b, size, cin, cout, nbytes = ctx.meta
        dy = dy.contiguous()
        dev = dy.device
        dx = dw = None
        target = None
        if ctx.needs_input_grad[1] and state is not None and ctx.weight_param is not None:
            target = _direct_grad_target(ctx.weight_param, w.shape)[0]
        direct = target is not None and target.stride() == w.stride()

        def weight_grad():
            ws2 = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            dwn = torch.empty_like(w)                                   # gradient w.r.t. the normalised weight
            _lib.call("hg_conv5s2_dw", _ptr(dy), _ptr(x_s2d), _ptr(dwn), _ptr(ws2), nbytes, b, cin, cout, size, 0, _stream())
            if state is None:
                return dwn
            dw_orig = target if direct else torch.empty_like(w)
            lib = _lib.load()
            co_arr, ci_arr, t_arr = _c_array(ctypes.c_int, [cout]), _c_array(ctypes.c_int, [cin]), _c_array(ctypes.c_int, [25])
            sn_bytes = lib.hg_spectral_norm_workspace_bytes(1, co_arr, ci_arr, t_arr)
            sn_ws = torch.empty(sn_bytes, dtype=torch.uint8, device=dev)
            vp = ctypes.c_void_p
            _lib.call("hg_spectral_norm_bwd", 1, _c_array(vp, [dwn.data_ptr()]), _c_array(vp, [w.data_ptr()]),
                      _c_array(vp, [state.data_ptr()]), _c_array(vp, [dw_orig.data_ptr()]), co_arr, ci_arr, t_arr,
                      int(direct), HG_F32, _ptr(sn_ws), sn_bytes, _stream())
            return None if direct else dw_orig

        # the weight gradient (conv dw + spectral-norm backward) adds straight into the owner's flat buffer: it runs on the
        # wgrad side stream beside the dx chain (see _ConvT.backward); the two passes of a D step stay ordered on that stream
        on_side = WGRAD_SIDE_STREAM and direct and ctx.needs_input_grad[0]
        if on_side:
            cur, side = torch.cuda.current_stream(dev), side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                weight_grad()
            for t in (dy, x_s2d, state):
                t.record_stream(side)
            side_stream_mark(dev)
        if ctx.needs_input_grad[0]:
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            dx = torch.empty_like(x_s2d)
            inv_sigma = ctypes.c_void_p(0 if state is None else state.data_ptr() + 4)
            _lib.call("hg_conv5s2_dx", _ptr(dy), _ptr(wt), inv_sigma, _ptr(dx), _ptr(ws), nbytes, b, cin, cout, size, _stream())
        if ctx.needs_input_grad[1]:
            if not on_side:
                dw = weight_grad()
            if direct:
                _grad_ready(ctx.weight_param)
        return dx, dw, None


def conv5s2_sn(x_s2d: Tensor, weight_orig: Tensor, state: Optional[Tensor]) -> Tensor:
    """Spectrally normalised Conv2d(kernel 5, stride 2, padding 2, no bias) of a space-to-depth input (reference
    core/models/hologan_discriminator.py:12-15,20) on the tcgen05 kernels."""
    return _Conv5s2SN.apply(x_s2d, weight_orig, state)


def conv5s2_supported(cin: int, cout: int, size_out: int) -> bool:
    return cin % 64 == 0 and cout % 128 == 0 and size_out >= 2 and (size_out & (size_out - 1)) == 0


class _DHeads(torch.autograd.Function):
    """(logits (B,1), z_pred (B,zdim)) fp32 from the last block's channels-last activation h (B,H,W,C) bf16."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2, b2, w3, b3, neg_slope):
        _require_cuda(h, w1, w2, w3)
        if h.dtype != torch.bfloat16 or not h.is_contiguous() or h.dim() != 4:
            raise ValueError("h must be a contiguous bf16 (B, H, W, C) tensor")
        b, hw, c = h.shape[0], h.shape[1] * h.shape[2], h.shape[3]
        zdim = w3.shape[0]
        ps = [t.detach().float().contiguous() for t in (w1, b1, w2, b2, w3, b3)]
        if tuple(ps[0].shape) != (1, c * hw) or tuple(ps[2].shape) != (128, c * hw) or ps[4].shape[1] != 128:
            raise ValueError("dheads: linear1 (1, F), linear2 (128, F), linear3 (zdim, 128) with F = C * H * W")
        dev = h.device
        nbytes = _lib.load().hg_dheads_workspace_bytes(b, c, hw, zdim)
        if nbytes < 0:
            raise _lib.HologanB200Error(f"hg_dheads: unsupported shape B={b} C={c} HW={hw}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        logits = torch.empty((b, 1), dtype=torch.float32, device=dev)
        t2 = torch.empty((b, 128), dtype=torch.float32, device=dev)
        zp = torch.empty((b, zdim), dtype=torch.float32, device=dev)
        _lib.call("hg_dheads_fwd", _ptr(h), *[_ptr(t) for t in ps], _ptr(logits), _ptr(t2), _ptr(zp), _ptr(ws), nbytes, b, c, hw,
                  zdim, ctypes.c_float(neg_slope), _stream())
        ctx.save_for_backward(h, ps[0], ps[2], ps[4], t2, zp)
        ctx.meta = (b, c, hw, zdim, float(neg_slope), nbytes)
        ctx.set_materialize_grads(False)
        return logits, zp

    @staticmethod
    def backward(ctx, dlogits, dzp):
        h, w1, w2, w3, t2, zp = ctx.saved_tensors
        b, c, hw, zdim, neg_slope, nbytes = ctx.meta
        dev = h.device
        dlogits = None if dlogits is None else dlogits.float().contiguous()
        dzp = None if dzp is None else dzp.float().contiguous()
        want_params = any(ctx.needs_input_grad[1:7])
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        grads = [None] * 6
        if want_params:
            grads = [torch.empty_like(w1), torch.empty(1, dtype=torch.float32, device=dev), torch.empty_like(w2),
                     torch.empty(128, dtype=torch.float32, device=dev), torch.empty_like(w3),
                     torch.empty(zdim, dtype=torch.float32, device=dev)]
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("hg_dheads_bwd", _ptr(h), _ptr(w1), _ptr(w2), _ptr(w3), _ptr(t2), _ptr(zp), _ptr(dlogits), _ptr(dzp), _ptr(dh),
                  *[_ptr(g) for g in grads], _ptr(ws), nbytes, b, c, hw, zdim, ctypes.c_float(neg_slope), _stream())
        return (dh,) + tuple(grads) + (None,)


def dheads(h: Tensor, w1, b1, w2, b2, w3, b3, neg_slope: float = 0.2) -> Tuple[Tensor, Tensor]:
    """logits = linear1(flat(h)), z_pred = tanh(linear3(leaky_relu(linear2(flat(h))))) with flat = the reference's
    (c, h, w) flatten (core/models/hologan_discriminator.py:60-68); h is channels-last."""
    return _DHeads.apply(h, w1, b1, w2, b2, w3, b3, neg_slope)


def dheads_supported(batch: int, channels: int, hw: int) -> bool:
    return 1 <= batch <= 64 and hw <= 64 and 64 % hw == 0 and channels % (64 // hw) == 0
