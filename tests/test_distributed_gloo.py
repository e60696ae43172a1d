"""Data-parallel host logic on CPU: world_size 2, gloo backend.

The generator has no CPU path, so these tests drive the pieces of `HologanTrainer` that do not touch the
CUDA library: replica synchronisation at construction, per-rank latent/view streams, and the flat-buffer
gradient all-reduce (mean) + Adam step on the discriminator mirror, which must equal one process stepping
on the concatenated batch (every op of the path is per-sample: no cross-sample statistics).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import rel_err
from lightning_gan_zoo_b200.training import HologanConfig, HologanTrainer, _FlatGrads

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _d_loss(disc, real, fake, z):
    """The D branch of HOLOGAN.training_step (core/lightning_module.py:217-228) on given fakes."""
    bce = F.binary_cross_entropy_with_logits
    d_real, _ = disc(real)
    d_fake, z_pred = disc(fake)
    return (bce(d_real, torch.ones_like(d_real)) + bce(d_fake, torch.zeros_like(d_fake))) / 2 + torch.mean((z_pred - z) ** 2)


def _worker(rank: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(2)
        torch.manual_seed(1000 + rank)                  # diverged RNG before construction: the ctor must re-sync
        torch.rand(rank + 1)
        cfg = HologanConfig(batch_size=4, gen_in_planes=8, disc_out_planes=8)
        tr = HologanTrainer(cfg, device="cpu", compute_dtype=torch.float32, rank=rank, world_size=WORLD, seed=42 + 7 * rank)
        # 1. identical replicas after the constructor's broadcast
        flat = torch.cat([t.detach().flatten().float() for t in list(tr.generator.state_dict().values()) +
                          list(tr.discriminator.state_dict().values())])
        gathered = [torch.empty_like(flat) for _ in range(WORLD)]
        dist.all_gather(gathered, flat)
        assert torch.equal(gathered[0], gathered[1]), "replicas differ after construction"
        # 2. decorrelated latent / view streams per rank
        z_local = tr.sample_noise(4)
        zs = [torch.empty_like(z_local) for _ in range(WORLD)]
        dist.all_gather(zs, z_local)
        assert not torch.equal(zs[0], zs[1])
        v = torch.from_numpy(tr.sample_view(4))
        vs = [torch.empty_like(v) for _ in range(WORLD)]
        dist.all_gather(vs, v)
        assert not torch.equal(vs[0], vs[1])
        # 3. one data-parallel D step == one single-process step on the concatenated batch
        gen = torch.Generator().manual_seed(5)
        real = torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1
        fake = torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1
        z = torch.rand(8, 128, generator=gen) * 2 - 1
        sl = slice(4 * rank, 4 * rank + 4)
        tr.d_grads.zero()
        _d_loss(tr.discriminator, real[sl], fake[sl], z[sl]).backward()
        tr.d_grads.all_reduce_mean(WORLD)
        tr.opt_d.step()
        after = torch.cat([p.detach().flatten() for p in tr.discriminator.parameters()])
        u_after = torch.cat([b.conv2d.weight_u.flatten() for b in tr.discriminator.blocks])
        torch.save({"params": after, "u": u_after, "grad": tr.d_grads.flat.clone()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_d_step_matches_single_process(tmp_path):
    mp.spawn(_worker, args=(_free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["params"], r1["params"]) and torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["u"], r1["u"])

    # single-process reference: same seed-42 init (rank 0's constructor seed), full batch of 8
    cfg = HologanConfig(batch_size=8, gen_in_planes=8, disc_out_planes=8)
    tr = HologanTrainer(cfg, device="cpu", compute_dtype=torch.float32, seed=42)
    gen = torch.Generator().manual_seed(5)
    real = torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1
    fake = torch.rand(8, 3, 64, 64, generator=gen) * 2 - 1
    z = torch.rand(8, 128, generator=gen) * 2 - 1
    tr.d_grads.zero()
    # the two ranks each run D on real-shard then fake-shard: two power iterations per rank, like this
    loss = 0.5 * (_d_loss(tr.discriminator, real[:4], fake[:4], z[:4]))
    loss.backward()
    # second shard through the SAME u/v state a rank would have had: rebuild a fresh replica for it
    tr2 = HologanTrainer(cfg, device="cpu", compute_dtype=torch.float32, seed=42)
    tr2.d_grads.zero()
    (0.5 * _d_loss(tr2.discriminator, real[4:], fake[4:], z[4:])).backward()
    full_grad = tr.d_grads.flat + tr2.d_grads.flat
    assert rel_err(r0["grad"], full_grad) < 1e-5          # max|a-b| / max|b| (SURVEY 8c metric)
    tr.d_grads.flat.copy_(full_grad)
    tr.opt_d.step()
    want = torch.cat([p.detach().flatten() for p in tr.discriminator.parameters()])
    # Adam normalises every coordinate to +-lr, so noise-level gradients may flip: compare the update
    assert (r0["params"] - want).abs().max().item() <= 2.0 * cfg.lr + 1e-9
    assert ((r0["params"] - want).abs() > 1e-6).float().mean().item() < 0.02
    # spectral-norm u depends on the weights only -> identical on every rank without a broadcast
    u = torch.cat([b.conv2d.weight_u.flatten() for b in tr.discriminator.blocks])
    assert torch.allclose(r0["u"], u, atol=1e-6)


def test_flat_grads_views_and_zero():
    lin = torch.nn.Linear(4, 3)
    fg = _FlatGrads(lin.parameters())
    lin(torch.ones(2, 4)).sum().backward()
    assert fg.numel == 15 and fg.flat.numel() == 16 and fg.flat.abs().sum() > 0      # padded to a float4 multiple
    assert lin.weight.grad.data_ptr() == fg.flat.data_ptr()
    fg.zero()
    assert lin.weight.grad.abs().sum() == 0 and lin.bias.grad.abs().sum() == 0
    fg.all_reduce_mean(1)           # world 1: no process group needed


_EXIT_SCRIPT = """
import sys, time, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import bench
torch.cuda.synchronize = lambda *a, **k: None      # no GPU here: the handshake itself is what is tested
dist.init_process_group("gloo")
rank = dist.get_rank()
if rank == 0:
    time.sleep(1.0)                                 # rank 0 still measures its rooflines while the others wait
    print("LINE", flush=True)
bench._exit_without_nccl_teardown(rank)
raise SystemExit(3)                                 # never reached
"""


def test_bench_multi_rank_exit_handshake(tmp_path):
    """bench.py leaves multi-rank runs through a store handshake + hard exit (no NCCL teardown under live CUDA
    graphs, which hung on 2 x B200): under torchrun every rank must end with code 0, after rank 0 printed."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "exit_check.py"
    script.write_text(_EXIT_SCRIPT.format(root=root))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(WORLD),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LINE" in r.stdout


def _bucket_worker(rank: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
        early = [net[2].weight, net[2].bias]                           # final first in the backward pass
        fg = _FlatGrads(net.parameters(), buckets=[early, [net[1].weight]])
        assert fg.params[0] is net[2].weight and fg.params[2] is net[1].weight and len(fg.bucket_slices) == 3
        x = torch.full((2, 6), float(rank + 1))
        fg.begin()
        net(x).sum().backward()
        fg.finish()
        before = fg.flat.clone()
        fg.fire(0, WORLD)                                              # bucket 0 starts while "the backward goes on"
        fg.fire(1, WORLD, after_calls=2)                               # not yet: needs two calls
        assert 1 not in fg._fired
        fg.fire(1, WORLD, after_calls=2)
        assert 1 in fg._fired
        fg.all_reduce_sum(WORLD)                                       # tail bucket + wait
        torch.save({"before": before, "after": fg.flat.clone()}, os.path.join(out_dir, f"b{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_bucketed_async_all_reduce_equals_flat_sum(tmp_path):
    """Gradient buckets fired one by one (the overlap schedule of HologanTrainer) sum to the same buffer as one flat
    all-reduce: every element exactly once, on every rank."""
    mp.spawn(_bucket_worker, args=(_free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    r0, r1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert torch.equal(r0["after"], r1["after"])
    assert torch.allclose(r0["after"], r0["before"] + r1["before"])
