"""Rotate-resample kernels vs the reference's golden outputs and the oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, np_ptr, rel_err
from lightning_gan_zoo_b200 import ops
from oracle import hologan_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("tag", ["s16", "s8"])
def test_fwd_fp32_bit_exact_vs_reference(tag):
    g = load_golden(f"rotate_{tag}.npz")
    vol = torch.from_numpy(g["vol"]).to(DEV)
    a = torch.from_numpy(g["a_inv"]).to(DEV)
    out, coords, idx = ops.rotate_fwd_raw(vol, a, ops.HG_BORDER_REFERENCE, debug=True)
    b, c, s = vol.shape[:3]
    # "grid coordinates" and "sampling indices": bit-exact
    assert np.array_equal(coords.cpu().numpy(), g["coords"])
    cx, cy, cz = (torch.from_numpy(g["coords"][i]).reshape(-1) for i in range(3))
    f = lambda t, d: (torch.floor(t).long() + d).clamp(0, s - 1)
    base = (torch.arange(b) * s ** 3).repeat_interleave(s ** 3)
    ref_idx = []
    for dz, dy, dx in ((0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)):
        ref_idx.append(base + f(cz, dz) * s * s + f(cy, dy) * s + f(cx, dx))      # reference :268-287
    ref_idx = torch.stack(ref_idx).reshape(8, b, -1).to(torch.int32)
    assert torch.equal(idx.cpu(), ref_idx)
    # fp32 forward with the reference border arithmetic: bit-exact
    assert np.array_equal(out.cpu().numpy(), g["out"])


@pytest.mark.parametrize("tag", ["s16", "s8"])
def test_fwd_zero_border_and_bf16(tag):
    g = load_golden(f"rotate_{tag}.npz")
    vol = torch.from_numpy(g["vol"]).to(DEV)
    a = torch.from_numpy(g["a_inv"]).to(DEV)
    out = ops.rotate_fwd_raw(vol, a, ops.HG_BORDER_ZERO)
    assert rel_err(out, g["out"]) < 1e-5          # fp32 tolerance of north_star
    inside = np.all((g["coords"] >= 0) & (g["coords"] < vol.shape[2] - 1), axis=0)       # (B, N)
    mask = torch.from_numpy(inside).reshape(vol.shape[0], 1, *vol.shape[2:]).expand_as(out)
    assert torch.equal(out.cpu()[~mask], torch.zeros_like(out.cpu()[~mask]))
    assert rel_err(out.cpu()[mask], torch.from_numpy(g["out"])[mask]) < 1e-6
    ob = ops.rotate_fwd_raw(vol.bfloat16(), a, ops.HG_BORDER_REFERENCE)
    assert ob.dtype == torch.bfloat16
    ref_b = orc.rotate_resample(vol.bfloat16().float().cpu(), a_inv=a.cpu())
    assert rel_err(ob.float(), ref_b) < 2 ** -8   # one bf16 rounding of the output
    assert rel_err(ob.float(), g["out"]) < 2e-2   # bf16 tolerance of north_star vs the fp32 reference


@pytest.mark.parametrize("tag", ["s16", "s8"])
@pytest.mark.parametrize("border", [ops.HG_BORDER_REFERENCE, ops.HG_BORDER_ZERO])
def test_bwd_vs_reference(tag, border):
    g = load_golden(f"rotate_{tag}.npz")
    a = torch.from_numpy(g["a_inv"]).to(DEV)
    vol = torch.from_numpy(g["vol"]).to(DEV).requires_grad_(True)
    out = ops.rotate_resample(vol, a, border)
    (out * torch.from_numpy(g["grad_out"]).to(DEV)).sum().backward()
    # views 0-4 (config range, axis aligned, identity, scale 0.7 / 1.5): north_star's 1e-5.  Views 5-7 shift
    # the grid by up to 5 voxels: their far-out-of-range samples carry weights ~1e2 that cancel pairwise, so
    # the reference's own index_put_ sums hold rounding residue of that order -> 5e-5 there.
    assert rel_err(vol.grad[:5], g["grad_vol"][:5]) < 1e-5
    assert rel_err(vol.grad[5:], g["grad_vol"][5:]) < 5e-5
    gb = ops.rotate_bwd_raw(torch.from_numpy(g["grad_out"]).to(DEV).bfloat16(), a, vol.shape[1], vol.shape[2], border)
    assert rel_err(gb.float(), g["grad_vol"]) < 2e-2


def test_sweep_coords_hash():
    g = load_golden("rotate_sweep100.npz")
    a = torch.from_numpy(g["a_inv"]).to(DEV)
    vol = torch.zeros(100, 1, 16, 16, 16, device=DEV)
    _, coords, _ = ops.rotate_fwd_raw(vol, a, debug=True)
    from conftest import sha16
    assert sha16(coords.cpu()) == str(g["coords_sha"])
    assert sha16(torch.floor(coords.cpu()).to(torch.int32)) == str(g["floor_sha"])


def test_identity_view_zeroes_last_planes():
    """identity view returns vol on [:15]^3 and exact 0 on index 15 of every axis."""
    vol = torch.randn(2, 3, 16, 16, 16, device=DEV)
    view = np.zeros((2, 6)); view[:, 2] = 1.0
    a = ops.view_to_affine(view).to(DEV)
    out = ops.rotate_fwd_raw(vol, a)
    assert torch.equal(out[:, :, :15, :15, :15], vol[:, :, :15, :15, :15])
    assert out[:, :, 15].abs().max() == 0 and out[:, :, :, 15].abs().max() == 0 and out[..., 15].abs().max() == 0


@pytest.mark.parametrize("shape", [(64, 64, 16), (8, 64, 32), (3, 7, 16), (2, 1, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_full_size_vs_c_oracle_and_adjoint(shape, dtype, c_oracle):
    """BASELINE cfg 3 sizes (and ragged channel counts): forward against the C oracle, and the
    size-independent adjoint identity <R v, g> == <v, R^T g> linking forward and backward."""
    b, c, s = shape
    gen = torch.Generator().manual_seed(b * 1000 + c)
    vol = torch.randn(b, c, s, s, s, generator=gen).to(dtype)
    gout = torch.randn(b, c, s, s, s, generator=gen).to(dtype)
    view = orc.sample_view(b, np.random.RandomState(c))
    a_cpu = ops.view_to_affine(view, s, s)
    a = a_cpu.to(DEV)
    out = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_REFERENCE)
    ref = np.empty((b, c, s, s, s), np.float32)
    vf = np.ascontiguousarray(vol.float().numpy())
    c_oracle.orc_rotate_fwd(np_ptr(vf), np_ptr(a_cpu.numpy()), np_ptr(ref), b, c, s)
    if dtype == torch.float32:
        assert np.array_equal(out.cpu().numpy(), ref)                  # bit-exact at full size
    else:
        assert rel_err(out.float(), ref) < 2 ** -8
    gv = ops.rotate_bwd_raw(gout.to(DEV), a, c, s, ops.HG_BORDER_REFERENCE)
    lhs = (out.double() * gout.to(DEV).double()).sum().item()
    rhs = (vol.to(DEV).double() * gv.double()).sum().item()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    scale = (out.double().abs() * gout.to(DEV).double().abs()).sum().item()
    assert abs(lhs - rhs) <= tol * scale / 10
    if dtype == torch.float32:
        refg = np.empty_like(ref)
        gf = np.ascontiguousarray(gout.float().numpy())
        c_oracle.orc_rotate_bwd(np_ptr(gf), np_ptr(a_cpu.numpy()), np_ptr(refg), b, c, s)
        assert rel_err(gv, refg) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("scale", [1.0, 0.7, 1.8])
def test_slab32_forward_vs_c_oracle(dtype, scale, c_oracle, hg_option):
    """32^3 forward on source-slab tiles (rotate_slab.cu, the default since r02a: 650-805 us vs 877-1052 us at
    (64,64,32^3) fp32) against the C oracle (bit-exact in fp32) and against the per-channel slab kernel it replaced
    (option ROTATE_SLAB32 = 0), both border modes."""
    b, c, s = 5, 16, 32
    gen = torch.Generator().manual_seed(int(scale * 10))
    vol = torch.randn(b, c, s, s, s, generator=gen).to(dtype)
    view = orc.sample_view(b, np.random.RandomState(11))
    view[:, 2] = scale
    view[1:, 3:6] = np.random.RandomState(12).uniform(-3, 3, (b - 1, 3))
    view[0, 0], view[0, 1] = np.deg2rad(270), np.deg2rad(90)
    a_cpu = ops.view_to_affine(view, s, s)
    a = a_cpu.to(DEV)
    hg_option("ROTATE_SLAB32", 0)
    base = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_REFERENCE)
    base_z = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_ZERO)
    hg_option("ROTATE_SLAB32", 1)
    out = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_REFERENCE)
    out_z = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_ZERO)
    assert torch.equal(out, base)          # same arithmetic, same order -> same bits
    assert rel_err(out_z.float(), base_z.float()) < (1e-6 if dtype == torch.float32 else 2 ** -8)
    if dtype == torch.float32:
        ref = np.empty((b, c, s, s, s), np.float32)
        c_oracle.orc_rotate_fwd(np_ptr(np.ascontiguousarray(vol.numpy())), np_ptr(a_cpu.numpy()), np_ptr(ref), b, c, s)
        assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("scale", [1.0, 0.7, 1.8])
def test_gather32_backward_vs_c_oracle(scale, c_oracle, hg_option):
    """32^3 backward as a table-free per-voxel gather (rotate_slab.cu, the default since r02a: 2.0 ms vs 5.6 ms for
    the shared-memory scatter at (64,64,32^3) fp32) against the C oracle's scatter adjoint; ragged channel count;
    deterministic; and against the scatter kernel it replaced (option ROTATE_GATHER_BWD = 0)."""
    b, c, s = 4, 7, 32
    gen = torch.Generator().manual_seed(int(scale * 10) + 1)
    gout = torch.randn(b, c, s, s, s, generator=gen)
    view = orc.sample_view(b, np.random.RandomState(21))
    view[:, 2] = scale
    view[1:, 3:6] = np.random.RandomState(22).uniform(-3, 3, (b - 1, 3))
    a_cpu = ops.view_to_affine(view, s, s)
    a = a_cpu.to(DEV)
    refg = np.empty((b, c, s, s, s), np.float32)
    c_oracle.orc_rotate_bwd(np_ptr(np.ascontiguousarray(gout.numpy())), np_ptr(a_cpu.numpy()), np_ptr(refg), b, c, s)
    assert _lib_option("ROTATE_GATHER_BWD") == 1
    gv = ops.rotate_bwd_raw(gout.to(DEV), a, c, s, ops.HG_BORDER_REFERENCE)
    gv2 = ops.rotate_bwd_raw(gout.to(DEV), a, c, s, ops.HG_BORDER_REFERENCE)
    gvb = ops.rotate_bwd_raw(gout.to(DEV).bfloat16(), a, c, s, ops.HG_BORDER_ZERO)
    assert torch.equal(gv, gv2)                                      # deterministic
    assert rel_err(gv, refg) < 1e-5
    assert rel_err(gvb.float(), refg) < 2e-2
    hg_option("ROTATE_GATHER_BWD", 0)
    old = ops.rotate_bwd_raw(gout.to(DEV), a, c, s, ops.HG_BORDER_REFERENCE)
    # the scatter kernel adds the reference's ~1e-7 out-of-range residues in shared-memory order: 1.2e-4 at scale 0.7
    assert rel_err(old, refg) < 3e-4


def _lib_option(name):
    from lightning_gan_zoo_b200 import _lib
    return _lib.get_option(name)


@pytest.mark.parametrize("scale", [0.6, 1.0, 1.5, 2.5])
@pytest.mark.parametrize("threads", [512, 1024])
@pytest.mark.parametrize("size", [16, 8])
def test_scaled_shifted_views_fwd_bwd(scale, threads, size, c_oracle):
    """Views with scale != 1 and shifts: the lattice shrinks / grows, so the adjoint table of the NCDHW gather
    backward gets long rows (scale 2.5 overflows its workspace -> the cell-table fallback).  Forward stays
    bit-exact, backward within 1e-5 of the C oracle's scatter adjoint; both CTA sizes of the tile pipeline."""
    from lightning_gan_zoo_b200 import _lib
    b, c, s = 6, 8, size
    gen = torch.Generator().manual_seed(int(scale * 10) + size)
    vol = torch.randn(b, c, s, s, s, generator=gen)
    gout = torch.randn(b, c, s, s, s, generator=gen)
    view = orc.sample_view(b, np.random.RandomState(7))
    view[:, 2] = scale
    view[:, 3:6] = np.random.RandomState(8).uniform(-2, 2, (b, 3))
    a_cpu = ops.view_to_affine(view, s, s)
    a = a_cpu.to(DEV)
    tune = _lib.HG_TUNE_CTA1024 if threads == 1024 else 0
    out = ops.rotate_fwd_raw(vol.to(DEV), a, ops.HG_BORDER_REFERENCE | tune)
    gv = ops.rotate_bwd_raw(gout.to(DEV), a, c, s, ops.HG_BORDER_REFERENCE | tune)
    gvb = ops.rotate_bwd_raw(gout.to(DEV).bfloat16(), a, c, s, ops.HG_BORDER_ZERO | tune)
    ref = np.empty((b, c, s, s, s), np.float32)
    refg = np.empty_like(ref)
    c_oracle.orc_rotate_fwd(np_ptr(np.ascontiguousarray(vol.numpy())), np_ptr(a_cpu.numpy()), np_ptr(ref), b, c, s)
    c_oracle.orc_rotate_bwd(np_ptr(np.ascontiguousarray(gout.numpy())), np_ptr(a_cpu.numpy()), np_ptr(refg), b, c, s)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert rel_err(gv, refg) < 1e-5
    assert rel_err(gvb.float(), refg) < 2e-2


def test_linearity():
    vol1 = torch.randn(4, 8, 16, 16, 16, device=DEV)
    vol2 = torch.randn(4, 8, 16, 16, 16, device=DEV)
    a = ops.view_to_affine(orc.sample_view(4, np.random.RandomState(3))).to(DEV)
    lhs = ops.rotate_fwd_raw(vol1 + 2 * vol2, a, ops.HG_BORDER_ZERO)
    rhs = ops.rotate_fwd_raw(vol1, a, ops.HG_BORDER_ZERO) + 2 * ops.rotate_fwd_raw(vol2, a, ops.HG_BORDER_ZERO)
    assert rel_err(lhs, rhs) < 1e-5


def test_errors():
    vol = torch.randn(1, 2, 16, 16, 16, device=DEV)
    a = torch.eye(4, device=DEV).unsqueeze(0)
    with pytest.raises(Exception):
        ops.rotate_fwd_raw(vol.double(), a)
    with pytest.raises(ValueError):
        ops.rotate_fwd_raw(vol, a[:, :3])
    with pytest.raises(RuntimeError):
        ops.rotate_fwd_raw(vol.cpu(), a.cpu())
    from lightning_gan_zoo_b200._lib import HologanB200Error
    with pytest.raises(HologanB200Error, match="size must be"):
        ops.rotate_fwd_raw(torch.randn(1, 1, 12, 12, 12, device=DEV), a)
