#!/usr/bin/env python
"""torch.profiler view of the training step: GPU kernel time per step, top kernels.  python tools/profile_step.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from lightning_gan_zoo_b200 import ops
from lightning_gan_zoo_b200.training import HologanConfig, HologanTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(os.environ.get("HG_BATCH", "64"))
dev = torch.device("cuda")
tr = HologanTrainer(HologanConfig(batch_size=B), device=dev)
real = torch.rand(B, 3, 64, 64, device=dev) * 2 - 1
zs = [tr.sample_noise(B).to(dev) for _ in range(3)]
avs = [ops.view_to_affine(tr.sample_view(B)).to(dev) for _ in range(3)]
for i in range(6):
    tr.step(real, i, z=zs[i % 3], view=avs[i % 3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        tr.step(real, i, z=zs[i % 3], view=avs[i % 3])
    torch.cuda.synchronize()
from torch.autograd import DeviceType
kern = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
tot = sum(e.self_device_time_total for e in kern)
print(f"GPU kernel time per step: {tot / steps / 1e3:.3f} ms over {steps} steps ({len(kern)} distinct kernels)")
def cat(k):
    if "hg::" in k: return "libhologan_b200"
    if any(t in k for t in ("cudnn", "cutlass", "nvjet", "implicit_convolve", "convolve_common", "gemv", "sgemm", "nhwcAddPadding")): return "cuDNN/cuBLAS (D convs, final conv, linears)"
    if "Memset" in k or "Memcpy" in k: return "memset/memcpy"
    if "multi_tensor_apply" in k or "Adam" in k: return "fused Adam"
    return "torch elementwise/reduce/copy"
cats = {}
for e in kern:
    cats[cat(e.key)] = cats.get(cat(e.key), 0) + e.self_device_time_total
for k, v in sorted(cats.items(), key=lambda kv: -kv[1]):
    print(f"  {v / steps / 1e3:7.3f} ms  {k}")
print(f"kernel launches per step: {sum(e.count for e in kern) / steps:.0f}")
for e in sorted(kern, key=lambda e: -e.self_device_time_total)[:70]:
    print(f"{e.self_device_time_total / steps:9.1f} us/step  x{e.count / steps:6.1f}  {e.key[:120]}")
