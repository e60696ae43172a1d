"""Stress: 300 steps at the bench configuration (B=64, CUDA graphs) with every side stream on vs everything serial --
losses must agree bit for bit at every step, parameters at the end."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lightning_gan_zoo_b200 import ops
from lightning_gan_zoo_b200.training import HologanTrainer, HologanConfig
from oracle import hologan_oracle as orc
B = 64
a = HologanTrainer(HologanConfig(batch_size=B), device="cuda", seed=21)
ops.WGRAD_SIDE_STREAM = False
b = HologanTrainer(HologanConfig(batch_size=B), device="cuda", seed=21)
b._sn_prefetch = False; b._d_real_side = False
ops.WGRAD_SIDE_STREAM = True
a.enable_cuda_graphs(B)
ops.WGRAD_SIDE_STREAM = False
b.enable_cuda_graphs(B)
ops.WGRAD_SIDE_STREAM = True
gen = torch.Generator().manual_seed(5)
bad = 0
for i in range(300):
    real = (torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1).cuda()
    z = (torch.rand(B, 128, generator=gen) * 2 - 1).cuda()
    view = ops.view_to_affine(orc.sample_view(B, np.random.RandomState(i)), 16, 16).cuda()
    la, lb = a.step(real, i, z=z, view=view), b.step(real, i, z=z, view=view)
    if la.item() != lb.item():
        bad += 1
        if bad < 5: print("step", i, la.item(), lb.item())
torch.cuda.synchronize()
neq = sum(int(not torch.equal(pa, pb)) for pa, pb in zip(list(a.generator.parameters()) + list(a.discriminator.parameters()),
                                                        list(b.generator.parameters()) + list(b.discriminator.parameters())))
print("steps with different loss:", bad, " parameters that differ:", neq, " last loss", la.item())
