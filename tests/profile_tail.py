"""Runs the frame tail a few times on a smooth and on a noise soft mask (for `ncu --metrics gpu__time_duration.sum`).
usage: python tests/profile_tail.py [H W] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfloodnet_b200 import tail  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda')
g = torch.Generator().manual_seed(5)
coarse = torch.randn(1, 2, 10, 16, generator=g).to(dev) * 4
smooth = torch.softmax(torch.nn.functional.interpolate(coarse, size=(480, 864), mode='bicubic', align_corners=False), dim=1)[0].contiguous()
noise = torch.rand((2, 480, 864), generator=g).to(dev)
ft = tail.FrameTail((H, W), [(W // 4, H // 4), (W // 2, H // 5)], dev)
for name, src in (('smooth', smooth), ('noise', noise)):
    for _ in range(reps):
        ft(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ft(src)
    e1.record()
    torch.cuda.synchronize()
    print(name, 'stats', ft.stats.tolist(), 'levels', ft.levels.tolist(), 'ms/frame', e0.elapsed_time(e1) / 20)
