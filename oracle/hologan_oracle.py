"""CPU oracle for the HoloGAN generator hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU, fp32) of the algorithm in the
reference `ebartrum/lightning_gan_zoo` (commit 33c7f1b).  It is NOT product code:
only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product path (`lightning_gan_zoo_b200`)
never imports anything under `oracle/` and has no CPU fallback.

Parity pinning: the reference ships no tests / golden vectors, so
this restatement is pinned against *outputs of the reference itself*, executed in
the build container by `oracle/gen_golden.py` (which imports /root/reference) and
committed under `tests/golden/`.  `tests/test_oracle_golden.py` replays them.

Every function cites the reference lines it follows (paths relative to the
reference root).  The dense contractions (conv_transpose / conv / linear) are
third-party arithmetic in the reference too (PyTorch, unpinned by the reference;
torch 2.11.0 CPU in this image) and are called through `torch.nn.functional`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# a5 / a6: view -> inverse homogeneous transform
# --------------------------------------------------------------------------------------

def _as_f32_column(view, col: int) -> Tensor:
    # core/models/hologan_generator.py:148-151,172,181-185 -- each view column is
    # converted with torch.as_tensor(...).float(); numpy f64 input is rounded to f32 here.
    v = view[:, col]
    return torch.as_tensor(v).float().reshape(-1).cpu()


def _eye4(batch: int) -> Tensor:
    return torch.eye(4, dtype=torch.float32).repeat(batch, 1, 1)


def view_to_transform(view) -> Tensor:
    """M = T @ S @ (Rz @ Ry), (B,4,4) fp32.  hologan_generator.py:145-196.

    The matrices are assembled entry-wise (same values as the reference's
    torch.cat construction) and multiplied in the reference's association order.
    """
    theta = _as_f32_column(view, 0)
    gamma = _as_f32_column(view, 1)
    scale = _as_f32_column(view, 2)
    shift = [_as_f32_column(view, 3 + i) for i in range(3)]
    b = theta.shape[0]

    rz = _eye4(b)                                   # :156-160
    rz[:, 0, 0] = theta.cos();  rz[:, 0, 1] = theta.sin()
    rz[:, 1, 0] = -theta.sin(); rz[:, 1, 1] = theta.cos()

    ry = _eye4(b)                                   # :163-167
    ry[:, 0, 0] = gamma.cos();  ry[:, 0, 2] = gamma.sin()
    ry[:, 2, 0] = -gamma.sin(); ry[:, 2, 2] = gamma.cos()

    rot = torch.matmul(rz, ry)                      # :169

    sc = _eye4(b)                                   # :174-178
    for i in range(3):
        sc[:, i, i] = scale

    tr = _eye4(b)                                   # :187-191
    for i in range(3):
        tr[:, i, 3] = shift[i]

    m = torch.matmul(tr, sc)                        # :193
    return torch.matmul(m, rot)                     # :194


def transform_to_inverse(m: Tensor, size: int, new_size: int) -> Tensor:
    """A = inverse(Tn @ M @ Tc), (B,4,4).  hologan_generator.py:198-221."""
    b = m.shape[0]
    tc = _eye4(b)
    tc[:, :3, 3] = -size * 0.5                      # :203-208
    tn = _eye4(b)
    tn[:, :3, 3] = new_size * 0.5                   # :212-217
    a = torch.matmul(tn, m)                         # :219
    a = torch.matmul(a, tc)                         # :220
    return a.inverse()                              # :221


def view_to_affine(view, size: int = 16, new_size: int = 16) -> Tensor:
    """(B,6) view -> (B,4,4) inverse transform used for sampling."""
    return transform_to_inverse(view_to_transform(view), size, new_size)


def lattice(new_size: int) -> Tensor:
    """(4, new_size^3) homogeneous integer lattice, x fastest then y then z.
    hologan_generator.py:323-331 (meshgrid called with (depth,height,width), 'ij')."""
    r = torch.arange(new_size)
    zz, yy, xx = torch.meshgrid(r, r, r, indexing="ij")
    flat = [t.reshape(-1).float() for t in (xx, yy, zz)]
    return torch.stack(flat + [torch.ones_like(flat[0])], dim=0)


def source_coords(a_inv: Tensor, new_size: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Per-output-voxel source coordinates (x,y,z), each (B*new_size^3,).
    hologan_generator.py:225-232.  Uses the same batched matmul as the reference so the
    fp32 bits are the reference's (see oracle/rotate_oracle.c for the fmaf-chain form)."""
    b = a_inv.shape[0]
    g = lattice(new_size).unsqueeze(0).repeat(b, 1, 1)
    t = torch.matmul(a_inv, g)
    return t[:, 0, :].reshape(-1), t[:, 1, :].reshape(-1), t[:, 2, :].reshape(-1)


# --------------------------------------------------------------------------------------
# a7: clamped-corner trilinear gather
# --------------------------------------------------------------------------------------

def trilinear_clamped(vol: Tensor, x: Tensor, y: Tensor, z: Tensor) -> Tensor:
    """hologan_generator.py:245-321.  vol (B,C,S,S,S); x,y,z (B*N,) -> (B*N, C).

    Corner indices are floor / floor+1 clamped to [0,S-1]; weights use the CLAMPED corner
    as a float against the UNCLAMPED coordinate, so out-of-range samples cancel to ~0
    instead of being border-extended.  x indexes tensor dim 4, y dim 3,
    z dim 2.  Corner order a..h and the (wx*wy)*wz product order follow :309-320.
    """
    bsz, ch, s2, s3, s4 = vol.shape
    n = x.numel() // bsz
    lo = {}
    hi = {}
    for name, c, lim in (("x", x, s4), ("y", y, s3), ("z", z, s2)):
        f = torch.floor(c).long()
        lo[name] = f.clamp(0, lim - 1)              # :249-261
        hi[name] = (f + 1).clamp(0, lim - 1)
    base = (torch.arange(bsz) * (s2 * s3 * s4)).repeat_interleave(n)     # :263-265
    flat = vol.permute(0, 2, 3, 4, 1).reshape(-1, ch)                    # :290

    def fetch(zc, yc, xc):
        return flat[base + zc * (s3 * s4) + yc * s4 + xc]                # :268-299

    fx0, fx1 = lo["x"].float(), hi["x"].float()
    fy0, fy1 = lo["y"].float(), hi["y"].float()
    fz0, fz1 = lo["z"].float(), hi["z"].float()
    ux, lx = fx1 - x, x - fx0
    uy, ly = fy1 - y, y - fy0
    uz, lz = fz1 - z, z - fz0
    # (corner z, corner y, corner x, weight) in the reference's a,b,c,d,e,f,g,h order
    terms = (
        (lo["z"], lo["y"], lo["x"], ux * uy * uz),   # a
        (lo["z"], hi["y"], lo["x"], ux * ly * uz),   # b
        (lo["z"], lo["y"], hi["x"], lx * uy * uz),   # c
        (lo["z"], hi["y"], hi["x"], lx * ly * uz),   # d
        (hi["z"], lo["y"], lo["x"], ux * uy * lz),   # e
        (hi["z"], hi["y"], lo["x"], ux * ly * lz),   # f
        (hi["z"], lo["y"], hi["x"], lx * uy * lz),   # g
        (hi["z"], hi["y"], hi["x"], lx * ly * lz),   # h
    )
    acc = None
    for zc, yc, xc, w in terms:                      # :320 left-to-right sum
        t = w.unsqueeze(1) * fetch(zc, yc, xc)
        acc = t if acc is None else acc + t
    return acc


def rotate_resample(vol: Tensor, view=None, *, a_inv: Optional[Tensor] = None,
                    new_size: Optional[int] = None) -> Tensor:
    """transformation3d: (B,C,S,S,S) -> (B,C,S',S',S').  hologan_generator.py:145-243."""
    size = vol.shape[2]
    new_size = size if new_size is None else new_size
    if a_inv is None:
        a_inv = view_to_affine(view, size, new_size)
    x, y, z = source_coords(a_inv, new_size)
    out = trilinear_clamped(vol, x, y, z)
    bsz, ch = vol.shape[:2]
    return out.reshape(bsz, new_size, new_size, new_size, ch).permute(0, 4, 1, 2, 3)   # :241-242


def project_depth_to_channels(rot: Tensor) -> Tensor:
    """hologan_generator.py:130-133: out[b, c*S + j, r, col] = rot[b, c, r, S-1-j, col]."""
    b, c, s = rot.shape[:3]
    t = rot.permute(0, 1, 3, 2, 4)
    t = torch.flip(t, dims=[2])
    return t.reshape(b, c * s, s, s)


# --------------------------------------------------------------------------------------
# a2 / a3: style mapping and AdaIN
# --------------------------------------------------------------------------------------

def zmapping(z: Tensor, weight: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    """hologan_generator.py:7-18: relu(linear) split into (scale, bias)."""
    out = F.relu(F.linear(z, weight, bias))
    c = weight.shape[0] // 2
    return out[:, :c], out[:, c:]


def adain(x: Tensor, scale: Tensor, bias: Tensor) -> Tensor:
    """hologan_generator.py:333-345: unbiased variance, eps=1e-8 inside rsqrt."""
    b, c = x.shape[:2]
    flat = x.reshape(b, c, -1)
    bshape = (b, c) + (1,) * (x.dim() - 2)
    mu = flat.mean(2).reshape(bshape)
    var = flat.var(2).reshape(bshape)                # torch default: N-1
    y = (x - mu) * torch.rsqrt(var + 1e-8)
    return scale.reshape(bshape) * y + bias.reshape(bshape)


# --------------------------------------------------------------------------------------
# generator forward (functional over a reference-keyed state dict)
# --------------------------------------------------------------------------------------

class _StoreBf16(torch.autograd.Function):
    """Value stored as bf16 (round to nearest even), its gradient stored as bf16 too."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _OperandBf16(torch.autograd.Function):
    """bf16 copy of an fp32 master weight: the forward value is rounded, the gradient stays fp32."""

    @staticmethod
    def forward(ctx, w):
        return w.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def _ident(t: Tensor) -> Tensor:
    return t


def _style_block(h: Tensor, z: Tensor, p: Dict[str, Tensor], prefix: str, dims: int, store=_ident,
                 operand=_ident) -> Tensor:
    """BasicBlock.forward, hologan_generator.py:37-42."""
    w, bia = operand(p[prefix + ".convTranspose.weight"]), p[prefix + ".convTranspose.bias"]
    if dims == 3:
        h = F.conv_transpose3d(h, w, bia, stride=2, padding=1, output_padding=1)      # :29-30
    else:
        h = F.conv_transpose2d(h, w, bia, stride=2, padding=1)                        # :26-27
    s, b = zmapping(z, p[prefix + ".zMapping.linear1.weight"], p[prefix + ".zMapping.linear1.bias"])
    return store(F.relu(adain(store(h), s, b)))


def generator_forward(p: Dict[str, Tensor], z: Tensor, view, img_size: int = 64,
                      patched128: bool = True, stages: Optional[dict] = None, bf16_storage: bool = False) -> Tensor:
    """Generator.forward, hologan_generator.py:116-143.  `p` uses the reference's
    state_dict keys.  img_size==128 needs `patched128` (the reference's 128
    branch lacks stride=2; the patched variant uses ConvTranspose2d(k4,s2,p1)).

    `bf16_storage=True` is NOT the reference: it is the same fp32 CPU arithmetic with every activation (and its
    gradient) that a bf16 pipeline keeps in memory rounded to bf16 where it is stored, and bf16 copies of the
    convolution weights -- the error floor of ANY implementation with bf16 operands and fp32 accumulation, used by
    the tests to tell kernel defects from the price of the storage format (tests/test_gpu_generator.py)."""
    store, operand = (_StoreBf16.apply, _OperandBf16.apply) if bf16_storage else (_ident, _ident)
    bsz = z.shape[0]
    x = p["x"].repeat(bsz, 1, 1, 1, 1)                                                # :121
    s0, b0 = zmapping(z, p["zMapping.linear1.weight"], p["zMapping.linear1.bias"])    # :122
    h0 = store(F.relu(adain(x, s0, b0)))                                              # :123-124
    h1 = _style_block(h0, z, p, "block1", 3, store, operand)                          # :126
    h2 = _style_block(h1, z, p, "block2", 3, store, operand)                          # :127
    rot = store(rotate_resample(h2, view))                                            # :129
    h2d = project_depth_to_channels(rot)                                              # :130-133
    h3 = store(F.relu(F.conv_transpose2d(h2d, operand(p["convTranspose2d1.weight"]), p["convTranspose2d1.bias"])))  # :135-136
    h4 = _style_block(h3, z, p, "block3", 2, store, operand)                          # :138
    h5 = _style_block(h4, z, p, "block4", 2, store, operand)                          # :139
    if img_size == 64:
        h6 = F.conv2d(h5, p["final_layer.weight"], p["final_layer.bias"], padding=1)  # :70
    elif img_size == 128:
        if not patched128:
            raise ValueError("reference 128x128 branch is broken (SURVEY.md R4)")
        h6 = F.conv_transpose2d(h5, p["final_layer.weight"], p["final_layer.bias"], stride=2, padding=1)
    else:
        raise ValueError("img_size must be 64 or 128")
    out = torch.tanh(h6)                                                              # :142
    if stages is not None:
        stages.update(h0=h0, h1=h1, h2=h2, rot=rot, h2d=h2d, h3=h3, h4=h4, h5=h5, out=out)
    return out


def init_generator_params(in_planes: int = 64, out_planes: int = 3, z_planes: int = 128,
                          img_size: int = 64, generator: Optional[torch.Generator] = None,
                          bias_std: float = 0.0) -> Dict[str, Tensor]:
    """Random parameters with the reference's shapes and init distributions
    (hologan_generator.py:11-13,32-33,49,60-62,70-75).  RNG consumption order is the
    oracle's own (fixtures pin a sha256 of the result).

    `bias_std > 0` replaces the reference's all-zero bias init by N(0, bias_std), i.e. a
    "trained-like" state.  Parity fixtures use it because with exactly-zero biases the ReLU
    after the 1x1 projection sits on pre-activations that are pure rounding noise wherever the
    rotated volume is out of range (~1e-9), which makes the REFERENCE's own gradient there a
    coin flip (observed: 0.9 % change of sum|d bias| between two fp32 evaluation orders)."""
    g = generator

    def nrm(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    p: Dict[str, Tensor] = {}
    c0 = in_planes * 8
    p["x"] = (torch.randn(1, c0, 4, 4, 4, generator=g) - 0.5) / 0.5

    def zmap(prefix, c):
        p[prefix + ".linear1.weight"] = nrm(2 * c, z_planes)
        p[prefix + ".linear1.bias"] = torch.zeros(2 * c)

    zmap("zMapping", c0)
    chans3 = [(c0, in_planes * 2), (in_planes * 2, in_planes)]
    for i, (ci, co) in enumerate(chans3, start=1):
        p[f"block{i}.convTranspose.weight"] = nrm(ci, co, 3, 3, 3)
        p[f"block{i}.convTranspose.bias"] = torch.zeros(co)
        zmap(f"block{i}.zMapping", co)
    cp = in_planes * 16
    p["convTranspose2d1.weight"] = nrm(cp, cp, 1, 1)
    p["convTranspose2d1.bias"] = torch.zeros(cp)
    chans2 = [(cp, in_planes * 4), (in_planes * 4, in_planes)]
    for i, (ci, co) in enumerate(chans2, start=3):
        p[f"block{i}.convTranspose.weight"] = nrm(ci, co, 4, 4)
        p[f"block{i}.convTranspose.bias"] = torch.zeros(co)
        zmap(f"block{i}.zMapping", co)
    if img_size == 64:
        p["final_layer.weight"] = nrm(out_planes, in_planes, 3, 3)
    else:
        p["final_layer.weight"] = nrm(in_planes, out_planes, 4, 4)
    p["final_layer.bias"] = torch.zeros(out_planes)
    if bias_std > 0:
        for k in sorted(p):
            if k.endswith(".bias"):
                p[k] = torch.randn(p[k].shape, generator=g) * bias_std
    return p


# --------------------------------------------------------------------------------------
# a14: discriminator (functional) and a13: training-step losses
# --------------------------------------------------------------------------------------

def _l2_normalize(v: Tensor, eps: float = 1e-12) -> Tensor:
    return v / v.norm().clamp_min(eps)


def spectral_weight(w_orig: Tensor, u: Tensor, v: Tensor, training: bool
                    ) -> Tuple[Tensor, Tensor, Tensor]:
    """torch.nn.utils.spectral_norm (legacy hook, n_power_iterations=1, dim=0, eps=1e-12) as
    applied at hologan_discriminator.py:15.  Returns (weight, new_u, new_v); u/v only move in
    training mode and carry no gradient."""
    mat = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v = _l2_normalize(torch.mv(mat.t(), u))
            u = _l2_normalize(torch.mv(mat, v))
    sigma = torch.dot(u, torch.mv(mat, v))
    return w_orig / sigma, u, v


def discriminator_forward(p: Dict[str, Tensor], x: Tensor, training: bool = True,
                          update_buffers: bool = True, bf16_storage: bool = False) -> Tuple[Tensor, Tensor]:
    """Discriminator.forward, hologan_discriminator.py:56-70 (+ BasicBlock :19-23).
    `p` holds the reference state_dict keys (weight_orig / weight_u / weight_v for the
    spectrally-normalised convs).  In training mode the u/v entries of `p` are replaced
    by the power-iteration result, like the reference's buffers."""
    # bf16_storage: NOT the reference -- the same fp32 arithmetic with the tensors a bf16 pipeline keeps in memory
    # (activations, their gradients, convolution operands) rounded to bf16 where they are stored: the error floor of
    # any bf16-operand implementation (see generator_forward)
    store, operand = (_StoreBf16.apply, _OperandBf16.apply) if bf16_storage else (_ident, _ident)
    h = store(F.leaky_relu(F.conv2d(operand(x), operand(p["conv2d.weight"]), p["conv2d.bias"], stride=2, padding=2), 0.2))
    for i in range(3):
        k = f"blocks.{i}.conv2d."
        w, u, v = spectral_weight(p[k + "weight_orig"], p[k + "weight_u"], p[k + "weight_v"], training)
        if training and update_buffers:
            p[k + "weight_u"], p[k + "weight_v"] = u, v
        h = store(F.conv2d(h, operand(w), p[k + "bias"], stride=2, padding=2))
        h = F.instance_norm(h, eps=1e-5)                       # InstanceNorm2d defaults (:16)
        h = store(F.leaky_relu(h, 0.2))
    flat = h.reshape(x.shape[0], -1)
    logits = F.linear(flat, p["linear1.weight"], p["linear1.bias"])
    enc = F.leaky_relu(F.linear(flat, p["linear2.weight"], p["linear2.bias"]), 0.2)
    z_pred = torch.tanh(F.linear(enc, p["linear3.weight"], p["linear3.bias"]))
    return logits, z_pred


def init_discriminator_params(in_planes: int = 3, out_planes: int = 64, z_planes: int = 128,
                              img_size: int = 64, generator: Optional[torch.Generator] = None
                              ) -> Dict[str, Tensor]:
    """Shapes of hologan_discriminator.py:25-54 (img_size 128 uses the R4 patch:
    linear in_features = out_planes*8*(img_size//16)**2)."""
    g = generator

    def tn(*shape):
        # truncated normal by resampling 4 candidates (:72-78)
        t = torch.randn(*shape, 4, generator=g)
        ok = (t < 2) & (t > -2)
        idx = ok.max(-1, keepdim=True)[1]
        return t.gather(-1, idx).squeeze(-1) * 0.02

    p: Dict[str, Tensor] = {}
    p["conv2d.weight"] = tn(out_planes, in_planes, 5, 5)
    p["conv2d.bias"] = torch.zeros(out_planes)
    c = out_planes
    for i in range(3):
        k = f"blocks.{i}.conv2d."
        w = tn(2 * c, c, 5, 5)
        p[k + "weight_orig"] = w
        p[k + "bias"] = torch.zeros(2 * c)
        p[k + "weight_u"] = _l2_normalize(torch.randn(2 * c, generator=g))
        p[k + "weight_v"] = _l2_normalize(torch.randn(c * 25, generator=g))
        c *= 2
    feat = c * (img_size // 16) ** 2
    p["linear1.weight"] = tn(1, feat)
    p["linear1.bias"] = torch.zeros(1)
    bound = 1.0 / math.sqrt(feat)
    p["linear2.weight"] = (torch.rand(128, feat, generator=g) * 2 - 1) * bound
    p["linear2.bias"] = torch.zeros(128)
    bound = 1.0 / math.sqrt(128)
    p["linear3.weight"] = (torch.rand(z_planes, 128, generator=g) * 2 - 1) * bound
    p["linear3.bias"] = torch.zeros(z_planes)
    return p


def hologan_losses(optimizer_idx: int, d_params: Dict[str, Tensor], real: Optional[Tensor],
                   fake: Tensor, z: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
    """HOLOGAN.training_step losses, core/lightning_module.py:209-237."""
    bce = F.binary_cross_entropy_with_logits
    logs: Dict[str, Tensor] = {}
    if optimizer_idx == 0:                                                      # :217-228
        d_real, _ = discriminator_forward(d_params, real)
        l_real = bce(d_real, torch.ones_like(d_real))
        d_fake, z_pred = discriminator_forward(d_params, fake.detach())
        l_fake = bce(d_fake, torch.zeros_like(d_fake))
        d_loss = (l_real + l_fake) / 2
        q = torch.mean((z_pred - z) ** 2)
        logs["train/d_loss"], logs["train/q_loss"] = d_loss.detach(), q.detach()
        return d_loss + q, logs
    out, z_pred = discriminator_forward(d_params, fake)                          # :231-237
    g_loss = bce(out, torch.ones_like(out))
    q = torch.mean((z_pred - z) ** 2)
    logs["train/g_loss"], logs["train/q_loss"] = g_loss.detach(), q.detach()
    return g_loss + q, logs


def hologan_lr_lambda(total_epochs: int):
    """core/utils/hologan.py:3-9."""
    half = total_epochs / 2

    def f(epoch):
        return 1 if epoch <= half else 1 - ((epoch - half) / half)
    return f


def sample_view(batch: int, rng: np.random.RandomState, azimuth=(220, 320), elevation=(70, 110),
                scale=(1.0, 1.0), trans=((0, 0), (0, 0), (0, 0))) -> np.ndarray:
    """hologan_generator.py:80-114 with np.float -> np.float64; same RNG
    call order: azimuth ints, elevation ints, scalar scale, then x/y/z shifts."""
    theta = rng.randint(azimuth[0], azimuth[1], batch).astype(np.float64) * math.pi / 180.0
    if elevation[0] < elevation[1]:
        gamma = rng.randint(elevation[0], elevation[1], batch).astype(np.float64) * math.pi / 180.0
    else:
        gamma = np.zeros(batch)
    sc = float(rng.uniform(scale[0], scale[1]))
    view = np.zeros((batch, 6))
    view[:, 0], view[:, 1], view[:, 2] = theta, gamma, sc
    for i, (lo, hi) in enumerate(trans):
        view[:, 3 + i] = lo + rng.random_sample(batch) * (hi - lo)
    return view
