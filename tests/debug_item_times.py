"""Per-item timing of the CTA-pair phase-B kernel (clock64 stamps from the issuing warp): python tests/debug_item_times.py [N] [HW]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as vfn  # noqa: E402
from vfloodnet_b200 import _lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 1620
lib = _lib.load()
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(0)
keys, vals = zip(*[synth.gen_bank(g, n) for _ in range(2)])
info = [synth.gen_info(g, n, 50) for _ in range(2)]
q_in, q_out = [t.to(dev) for t in synth.gen_query(g, hw)]
fb = vfn.FeatureBank(2, 10 ** 7, dev)
fb.load_state(list(keys), list(vals), info)
m = vfn.Matcher(update_bank=True)
for _ in range(3):
    m(fb, q_in, q_out)
ts = torch.zeros((74, 64, 8), dtype=torch.int64, device=dev)
lib.vfn_debug_set_tstamp(ts.data_ptr())
m(fb, q_in, q_out)
torch.cuda.synchronize()
lib.vfn_debug_set_tstamp(None)
t = ts.cpu().numpy()
ok = t[:, :, 4] > 1
start, first, last, end, tiles = [t[:, :, i][ok].astype(np.float64) for i in range(5)]
fill = first - start
steady = (last - first) / (tiles - 1)
drain = end - last
total = end - start
print(f'N={n} HW={hw}: items {int(ok.sum())} (per cluster {ok.sum(1).min()}..{ok.sum(1).max()}), tiles/item {tiles.mean():.1f}')
print(f'  fill  (item start -> P of tile 0 ready)   {fill.mean():9.0f} clk  (min {fill.min():.0f}, max {fill.max():.0f})')
print(f'  steady tile period                        {steady.mean():9.1f} clk  (min {steady.min():.1f}, max {steady.max():.1f})')
print(f'  drain (P of last tile ready -> item end)  {drain.mean():9.0f} clk  (min {drain.min():.0f}, max {drain.max():.0f})')
print(f'  item total {total.mean():.0f} clk = {total.mean() / tiles.mean():.1f} per tile; cluster busy {np.array([(t[c, :, 3][ok[c]] - t[c, :, 0][ok[c]]).sum() for c in range(74)]).mean():.0f} clk')
span = t[:, :, 3][ok].max() - t[:, :, 0][ok].min()
print(f'  kernel span (first start -> last end) {span:.0f} clk')
