"""Where does the tcgen05 readout error come from on REAL (random-init, BN-calibrated) model features?
python tests/debug_readout_precision.py [frames]   (GPU; needs baseline/_ref)
Per sampled frame: error of the product's read against an fp64 read of the same bank, next to (a) the fp32 SIMT read,
(b) a torch emulation of the operand-format scheme (fp16 hi + e4m3/e5m2 corrections) with fp64 accumulation."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as vfn
from baseline import refshim, model_clip as MC

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device('cuda', 0)
ns = refshim.load()
model = MC.build_reference_model(ns, dev)
clip = MC.make_clip(frames)
f16 = lambda x: x.to(torch.float16).to(x.dtype)
e5 = lambda x: x.to(torch.float8_e5m2).to(x.dtype)
e4 = lambda x: x.clamp(-448, 448).to(torch.float8_e4m3fn).to(x.dtype)
box = {}
model.global_matcher.register_forward_hook(lambda m, i, o: box.__setitem__('q', (i[1].detach().clone(), i[2].detach().clone())))


def on_frame(t, frame, score, pm, k4, v4, fb):
    if t not in (1, 3, 6, 9, frames):
        return
    q_in, q_out = box['q']
    keys = [k.clone() for k in fb.keys]
    vals = [v.clone() for v in fb.values]
    info = [i.clone() for i in fb.info]
    outs = {}
    for impl in (1, 2):
        g = vfn.FeatureBank(2, 10 ** 7, dev, impl=impl)
        g.load_state(keys, vals, info)
        outs[impl] = vfn.Matcher(update_bank=False)(g, q_in, q_out)[0, :, :512]
    for c in range(2):
        K, V, Q = keys[c].double(), vals[c].double(), q_in[0].double()
        S = (K.t() @ Q) / math.sqrt(128)
        P = torch.softmax(S, dim=0)
        O = V @ P
        Pf, Vt = P.float().t(), V.float().t()
        Pp = Pf * 256.0
        Ph = f16(Pp); Pl = Pp - Ph
        Vh = f16(Vt); Vl = Vt - Vh
        emu = (Ph.double() @ Vh.double() + e4(Pl).double() @ e4(Vt).double() + e4(Pp).double() @ e5(Vl).double()) / 256.0
        p32 = torch.softmax((keys[c].t() @ q_in[0]) / math.sqrt(128), dim=0)
        o32 = vals[c] @ p32
        err = lambda x: (x.double() - O).abs().max().item()
        print(f'frame {t} obj {c} N {K.shape[1]} |V|max {V.abs().max().item():.2f} |O|max {O.abs().max().item():.2f} '
              f'Pmax-median {P.max(dim=0).values.median().item():.3f}:  tcgen05 {err(outs[2][c]):.2e}  simt {err(outs[1][c]):.2e}  '
              f'scheme emulation {err(emu.t()):.2e}  torch fp32 {err(o32):.2e}', flush=True)


MC.run_clip(model, ns.FeatureBank, clip, dev, on_frame=on_frame, keep_masks=False)

# ---- teacher-forced mask sensitivity on the calibrated model: reference vs patched model from the same bank state
ours = MC.patched_copy(model, vfn)
fb_ref = ns.FeatureBank(2, MC.BUDGET, dev)
with torch.no_grad():
    k4, v4 = model.memorize(clip[0].to(dev), MC.first_mask().to(dev))
    fb_ref.init_bank(k4, v4)
    for t in range(1, frames + 1):
        frame = clip[t].to(dev)
        st = ([k.clone() for k in fb_ref.keys], [v.clone() for v in fb_ref.values], [i.clone() for i in fb_ref.info])
        score_r, _ = model.segment(frame, fb_ref)
        g = vfn.FeatureBank(2, MC.BUDGET, dev)
        g.load_state(*st)
        score_o, _ = ours.segment(frame, g)
        pm_r = torch.softmax(score_r, 1)
        a, b = pm_r[0].argmax(0), torch.softmax(score_o, 1)[0].argmax(0)
        print(f'frame {t}: teacher-forced mask IoU {MC.iou(a, b):.5f}  pixels differ {int((a != b).sum())}  water '
              f'{float((a == 1).float().mean()):.3f}  max |score diff| {(score_r - score_o).abs().max().item():.2e}  '
              f'mean {(score_r - score_o).abs().mean().item():.2e}', flush=True)
        k4, v4 = model.memorize(frame, pm_r)
        fb_ref.update(k4, v4, t)
