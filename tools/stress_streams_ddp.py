"""Stress under torchrun (world 2): 120 steps at B=64 per GPU with the bucketed exchange + bucket-wise Adam + side streams
against one all-reduce and one Adam after a serial backward.  With two ranks a sum has two addends, so the results must agree
bit for bit whatever algorithm NCCL picks per message size.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/stress_streams_ddp.py"""
import os, sys, torch, numpy as np
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from lightning_gan_zoo_b200 import ops
from lightning_gan_zoo_b200.training import HologanTrainer, HologanConfig
from oracle import hologan_oracle as orc
B = 64
a = HologanTrainer(HologanConfig(batch_size=B), device="cuda", seed=21, rank=rank, world_size=world)
os.environ["HG_NO_GRAD_OVERLAP"] = "1"
ops.WGRAD_SIDE_STREAM = False
b = HologanTrainer(HologanConfig(batch_size=B), device="cuda", seed=21, rank=rank, world_size=world)
b._sn_prefetch = False; b._d_real_side = False
del os.environ["HG_NO_GRAD_OVERLAP"]
assert a._adam_overlap and not b._adam_overlap
ops.WGRAD_SIDE_STREAM = True
a.enable_cuda_graphs(B)
ops.WGRAD_SIDE_STREAM = False
b.enable_cuda_graphs(B)
ops.WGRAD_SIDE_STREAM = True
gen = torch.Generator().manual_seed(5 + rank)
bad = 0
for i in range(120):
    real = (torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1).cuda()
    z = (torch.rand(B, 128, generator=gen) * 2 - 1).cuda()
    view = ops.view_to_affine(orc.sample_view(B, np.random.RandomState(1000 * rank + i)), 16, 16).cuda()
    la, lb = a.step(real, i, z=z, view=view), b.step(real, i, z=z, view=view)
    bad += int(la.item() != lb.item())
torch.cuda.synchronize()
neq = sum(int(not torch.equal(pa, pb)) for pa, pb in zip(list(a.generator.parameters()) + list(a.discriminator.parameters()),
                                                        list(b.generator.parameters()) + list(b.discriminator.parameters())))
print(f"rank {rank}: steps with different loss: {bad}, parameters that differ: {neq}", flush=True)
dist.barrier()
os._exit(0)
