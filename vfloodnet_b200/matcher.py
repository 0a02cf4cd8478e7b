"""Drop-in Matcher (reference: video_module/model/AFB_URR.py:130-178): the memory read.

forward(feature_bank, q_in, q_out) -> (bs=1, obj_n, 2*d_val, HW), and, when update_bank, the LFU usage
count side effect info[i][:,1] += log(cnt+1).  One C-ABI call (vfn_memread); no CPU / OOM fallback
(the reference's `except RuntimeError` CPU path, AFB_URR.py:147-157, is deliberately not carried over).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from ._lib import VFN_Q_IN_EM, check, em_backed, on_device, ptr, stream_ptr
from .feature_bank import FeatureBank


class Matcher(nn.Module):
    def __init__(self, thres_valid=1e-3, update_bank=False, impl=None):
        super().__init__()
        self.thres_valid = thres_valid
        self.update_bank = update_bank
        self.impl = impl          # None -> use the bank's setting
        self._ws = None
        self.last_lse = None
        self.want_lse = False
        self.launches = 0

    def _workspace(self, lib, fb: FeatureBank, hw: int, d_key: int, d_val: int):
        if any(s is None for s in fb._slabs):
            raise RuntimeError('Matcher.forward on an empty feature bank: call init_bank() first '
                               '(the reference fails on fb.keys[i].size() of None, AFB_URR.py:141)')
        n_max = max(max(s.cap for s in fb._slabs), 1)
        need = lib.vfn_memread_workspace_bytes(fb.obj_n, n_max, hw, d_key, d_val)
        if self._ws is None or self._ws.numel() < need or self._ws.device != fb.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=fb.device)
        return self._ws

    def forward(self, feature_bank, q_in, q_out):
        if not isinstance(feature_bank, FeatureBank):
            raise TypeError('vfloodnet_b200.Matcher needs a vfloodnet_b200.FeatureBank (device slabs); '
                            'there is no fallback for other bank types')
        fb = feature_bank
        with on_device(fb.device):
            return self._forward(fb, q_in, q_out)

    def _forward(self, fb, q_in, q_out):
        lib = _lib.load()
        if q_in.dim() != 3 or q_in.shape[0] != 1:
            raise ValueError('inference read expects q_in of shape (1, d_key, HW)')   # bs>1 is training only
        q_in = q_in.to(fb.device, torch.float32)
        q_em = em_backed(q_in[0])          # entry-major storage behind the (1, d_key, HW) shape (KeyValueHead): no copy
        if not q_em:
            q_in = q_in.contiguous()
        q_out = q_out.to(fb.device, torch.float32).contiguous()
        _, d_key, hw = q_in.shape
        d_val = q_out.shape[1]
        out = torch.empty((1, fb.obj_n, 2 * d_val, hw), dtype=torch.float32, device=fb.device)
        lse = torch.empty((fb.obj_n, hw), dtype=torch.float32, device=fb.device) if self.want_lse else None
        ws = self._workspace(lib, fb, hw, d_key, d_val)
        impl = fb.impl if self.impl is None else self.impl
        # the tcgen05 read takes the live bank sizes from device memory: updates still in flight need not be finished
        banks = fb.bank_array(bounds_ok=True, impl=int(impl))
        l0 = lib.vfn_launch_count()
        check(lib.vfn_memread(banks, fb.obj_n, ptr(q_in), ptr(q_out), hw, float(self.thres_valid),
                              int(bool(self.update_bank)), ptr(out), ptr(lse), ptr(ws), ws.numel(),
                              int(impl) | (VFN_Q_IN_EM if q_em else 0),
                              stream_ptr()), 'vfn_memread')
        self.last_lse = lse
        self.launches += lib.vfn_launch_count() - l0      # 4 on the tcgen05 path: phase A, LSE combine, phase B, combine
        return out
