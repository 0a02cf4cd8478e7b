"""GPU time of every frame of the bench clip (one CUDA event per frame): python tests/debug_frame_times.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vfloodnet_b200 as vfn  # noqa: E402

dev = torch.device('cuda', 0)
from vfloodnet_b200 import _lib  # noqa: E402
if os.environ.get('VFN_PDL') is not None:
    _lib.load().vfn_debug_set_pdl(int(os.environ['VFN_PDL']))
clip = bench.to_device(bench.make_clip(seed=100, frames=100, frac_merge=0.1, pin=False), dev)
for rep in range(3):
    fb = vfn.FeatureBank(2, bench.BUDGET, dev)
    m = vfn.Matcher(update_bank=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(101)]
    sub = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(100)]
    fb.init_bank(list(clip['keys0']), list(clip['vals0']))
    torch.cuda.synchronize()
    ev[0].record()
    for t in range(100):
        q_in, q_out, pk, pv = clip['frames'][t]
        p, r1, q_local = clip['urr']
        out = m(fb, q_in, q_out)
        sub[t][0].record()
        p_up, unc, conf, lm = vfn.urr_pre(p, r1.expand(2, -1, -1, -1), (1, 2, bench.R1_H, bench.R1_W))
        prob = vfn.urr_post(p_up, unc, conf, q_local)
        sub[t][1].record()
        fb.update(pk, pv, t + 1)
        ev[t + 1].record()
    torch.cuda.synchronize()
    if rep == 2:
        n = [fb.bank_n(c) for c in range(2)]
        print('final bank', n, 'total ms', ev[0].elapsed_time(ev[100]))
        for t in (0, 1, 2, 5, 10, 20, 30, 40, 50, 60, 70, 80, 90, 99):
            print(f'frame {t + 1:3d}: total {ev[t].elapsed_time(ev[t + 1]):.3f} ms  read {ev[t].elapsed_time(sub[t][0]):.3f}  '
                  f'urr {sub[t][0].elapsed_time(sub[t][1]):.3f}  update {sub[t][1].elapsed_time(ev[t + 1]):.3f}')
