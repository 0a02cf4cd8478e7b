"""Reference arm of the hot-path benchmark: the reference's own read + update on plain torch ops (CPU, or the same
torch ops on a CUDA device = BASELINE.json configs[1]'s "reference torch ops").

kind == 'reference': the UNMODIFIED `FeatureBank` and `Matcher` classes of the reference (baseline/_ref through
baseline/refshim.py).  The URR block is not separately callable in the reference (it is inlined in `Decoder.forward`
between convolutions, AFB_URR.py:214-237), so that one stage runs through the oracle's restatement of those lines.
kind == 'port': the reference is not staged; everything runs through oracle/afb_oracle.py.

Bench / test infrastructure only; none of the product's kernels or modules are on this path.
"""
from __future__ import annotations

import torch

from oracle import afb_oracle as O
from . import refshim


class RefArm:
    def __init__(self, budget: int, device='cpu', obj_n: int = 2, prefer_reference: bool = True):
        self.device = torch.device(device)
        self.budget, self.obj_n = budget, obj_n
        if prefer_reference and refshim.available():
            ns = refshim.load()
            self.kind = 'reference'
            self.fb = ns.FeatureBank(obj_n, budget, self.device)
            self.matcher = ns.Matcher(update_bank=True)
        else:
            self.kind = 'port'
            self.fb = O.OracleFeatureBank(obj_n, budget, self.device)
            self.matcher = None

    def D(self, t):
        return t.to(self.device)

    def init(self, keys, vals):
        """test_video_seg.py:100-101 (the bank adopts and later mutates the tensors: clones)"""
        self.fb.init_bank([self.D(k).clone() for k in keys], [self.D(v).clone() for v in vals])

    def load(self, keys, vals, info):
        """adopt a bank state ((d,N) keys / values, (N,2) info per object), e.g. a snapshot of a run in flight"""
        self.fb.init_bank([self.D(k).contiguous().clone() for k in keys], [self.D(v).contiguous().clone() for v in vals])
        self.fb.info = [self.D(i).contiguous().clone() for i in info]

    def sizes(self):
        return [int(k.shape[1]) for k in self.fb.keys]

    def frame(self, q_in, q_out, pk, pv, urr, frame_idx):
        """one hot-path step in the order of test_video_seg.py:108-112: read -> URR -> update"""
        fb = self.fb
        if self.matcher is not None:
            out = self.matcher(fb, q_in, q_out)                                   # AFB_URR.py:136-178, unmodified
        else:
            out = O.matcher_forward(fb.keys, fb.values, fb.info, q_in, q_out, 1e-3, update_bank=True).out
        prob = None
        if urr is not None:
            p, r1, q_local, feature_shape = urr
            p_up, unc, conf, _lm = O.urr_pre(p, r1, feature_shape)                # AFB_URR.py:214-231
            prob = O.urr_post(p_up, unc, conf, q_local)                           # AFB_URR.py:233-237
        fb.update(pk, pv, frame_idx)                                              # FeatureBank.py:53-143, unmodified
        return out, prob
