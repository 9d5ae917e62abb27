"""Kernel timeline of the data-parallel step (rank 0): every kernel of 2 CUDA-graph replays with start / duration, NCCL
kernels included -- shows where the all-reduces sit relative to the backward kernels and what they cost them.
    torchrun --nproc-per-node 2 tools/dp_timeline.py out.json      (or plain python for the 1-GPU reference timeline)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.autograd import DeviceType  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import xshinnosuke_b200 as xs  # noqa: E402
from xshinnosuke_b200 import nn  # noqa: E402
from xshinnosuke_b200.core.device import rt  # noqa: E402
from xshinnosuke_b200.recipes import ResNet18  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
rt.ensure()
xs.set_math_mode("tf32")
rs = np.random.RandomState(rank)
x, y = xs.tensor(rs.rand(128, 3, 224, 224).astype(np.float32)), xs.tensor(rs.randint(0, 100, (128,)))
np.random.seed(0)
net = ResNet18(num_classes=100)
opt = xs.optim.Adam(net.parameters(), lr=1e-3)
crit = nn.CrossEntropyLoss()
dp = None
if world > 1:
    from xshinnosuke_b200.parallel import DataParallel
    dp = DataParallel(net, opt, bucket_mb=float(os.environ.get("BUCKET_MB", "8")))


def step():
    opt.zero_grad()
    loss = crit(net(x), y)
    loss.backward()
    opt.step()
    return loss


g = xs.GraphedStep(step, warmup=3)
for _ in range(5):
    g()
torch.cuda.synchronize()
if dp is not None:
    dp.dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        g()
    torch.cuda.synchronize()
if rank == 0:
    ks = [(ev.name, ev.time_range.start, ev.time_range.elapsed_us()) for ev in prof.events() if ev.device_type == DeviceType.CUDA]
    ks.sort(key=lambda k: k[1])
    t0 = ks[0][1]
    json.dump([(n, s - t0, d) for n, s, d in ks], open(sys.argv[1], "w"))
    print("kernels", len(ks), "span us", ks[-1][1] + ks[-1][2] - t0)
if dp is not None:
    dp.dist.barrier()
    torch.distributed.destroy_process_group()
