"""Kernel-level timing at a fixed bank size (regime C): python tests/profile_kernels.py [N] [HW] [reps]
Prints average CUDA-event time per kernel kind (library hooks) and the implied algorithmic rates."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as vfn
from vfloodnet_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 1620
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _lib.load()
if os.environ.get('VFN_URR_MODE') is not None:
    lib.vfn_debug_set_urr_stream(int(os.environ['VFN_URR_MODE']))
if os.environ.get('VFN_PAIR') is not None:
    lib.vfn_debug_set_pair(int(os.environ['VFN_PAIR']))
g = torch.Generator().manual_seed(0)
keys, vals = zip(*[synth.gen_bank(g, n) for _ in range(2)])
info = [synth.gen_info(g, n, 50) for _ in range(2)]
q_in, q_out = synth.gen_query(g, hw)
pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw, 0.5) for c in range(2)])
dev = torch.device('cuda', 0)
q_in, q_out = q_in.to(dev), q_out.to(dev)
pk, pv = [k.to(dev) for k in pk], [v.to(dev) for v in pv]
m = vfn.Matcher(update_bank=True)
urr_p, urr_r1, _ = [t.to(dev) for t in synth.gen_urr_inputs(g, 2, 240, 432)]
urr_r1 = urr_r1.expand(2, -1, -1, -1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fresh():
    fb = vfn.FeatureBank(2, 10 ** 7, dev)
    fb.load_state(list(keys), list(vals), info)
    return fb


fb = fresh()
m(fb, q_in, q_out)
vfn.urr_pre(urr_p, urr_r1, (1, 2, 240, 432))
fb.update(pk, pv, 51)
torch.cuda.synchronize()
lib.vfn_profile_enable(1)
for _ in range(reps):
    fb = fresh()
    flush.zero_()
    m(fb, q_in, q_out)
    flush.zero_()
    vfn.urr_pre(urr_p, urr_r1, (1, 2, 240, 432))
    flush.zero_()
    fb.update(pk, pv, 51)
torch.cuda.synchronize()
prof = (ctypes.c_double * 24)()
lib.vfn_profile_collect(prof, 8)
names = ['read_phase_a', 'read_phase_b', 'match', 'compact_move', 'merge', 'append', 'urr_local']
print(f'N={n}/object, HW={hw}, 2 objects, reps={reps}')
for i, nm in enumerate(names):
    cnt, ms, work = prof[3 * i], prof[3 * i + 1], prof[3 * i + 2]
    if cnt:
        rate = work / (ms * 1e-3) if ms > 0 else 0
        unit = 'TFLOP/s' if i < 3 else 'GB/s'
        print(f'{nm:14s} launches {int(cnt):4d}  avg {ms / cnt:8.4f} ms   {rate / (1e12 if i < 3 else 1e9):9.1f} {unit} (algorithmic)')
