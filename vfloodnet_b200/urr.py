"""Uncertain-region refinement (reference: Decoder.forward, AFB_URR.py:214-237; myutils/data.py:42-48).

`urr_pre` / `urr_post` are the two fused CUDA stages either side of the three small local convolutions
(local_convFM / local_ResMM / local_pred2 stay cuDNN).  `decoder_forward` is a drop-in body for
`Decoder.forward` that routes the URR block through them; `patch_model` installs it (and the CUDA Matcher)
on a reference AFB_URR instance.
"""
from __future__ import annotations

import types

import torch
from torch.nn import functional as NF

from . import _lib
from ._lib import check, on_device, ptr, stream_ptr


def urr_pre(p: torch.Tensor, r1: torch.Tensor, feature_shape, out_local_match: torch.Tensor = None):
    """p: (obj_n, 2, h/2, w/2) logits of pred2; r1: (obj_n, C, h, w) (may be an expand()ed view of (1,C,h,w)).
    Returns p_up (obj_n,2,h,w), uncertainty (obj_n,1,h,w) [expanded view], r1_conf (obj_n,1,h,w),
    local_match (obj_n,2C,h,w).  out_local_match: write local_match into this (contiguous fp32) tensor instead of a
    fresh one (a CUDA-graph input buffer, vfloodnet_b200.graphed)."""
    lib = _lib.load()
    bs, obj_n, h, w = feature_shape
    if bs != 1:
        raise ValueError('URR kernels implement the inference path (bs == 1)')
    c = r1.shape[1]
    dev = p.device
    p = p.to(torch.float32).contiguous()
    if r1.stride(0) == 0 or r1.shape[0] == 1:
        r1c = r1[0].to(torch.float32).contiguous()
        obj_stride = 0
    else:
        r1c = r1.to(torch.float32).contiguous()
        obj_stride = c * h * w
    f32 = dict(dtype=torch.float32, device=dev)
    p_up = torch.empty((obj_n, 2, h, w), **f32)
    seg = torch.empty((obj_n, h, w), **f32)
    unc = torch.empty((h, w), **f32)
    conf = torch.empty((obj_n, 1, h, w), **f32)
    avg = torch.empty((obj_n, h, w), **f32)
    if out_local_match is not None:
        if tuple(out_local_match.shape) != (obj_n, 2 * c, h, w) or out_local_match.dtype != torch.float32 or \
                not out_local_match.is_contiguous() or out_local_match.device != dev:
            raise ValueError('out_local_match must be a contiguous fp32 (obj_n, 2C, h, w) tensor on the device of p')
        local_match = out_local_match
    else:
        local_match = torch.empty((obj_n, 2 * c, h, w), **f32)
    with on_device(dev):
        check(lib.vfn_urr_pre(ptr(p), ptr(r1c), obj_stride, obj_n, c, h, w, ptr(p_up), ptr(seg), ptr(unc), ptr(conf),
                              ptr(avg), ptr(local_match), stream_ptr()), 'vfn_urr_pre')
    return p_up, unc.view(1, 1, h, w).expand(obj_n, -1, -1, -1), conf, local_match


def urr_post(p_up: torch.Tensor, uncertainty: torch.Tensor, r1_conf: torch.Tensor, q_local: torch.Tensor):
    """Returns prob (obj_n, 2h, 2w) = softmax(bilinear_x2(p_up + unc * conf * q_local), dim=1)[:, 1]."""
    lib = _lib.load()
    obj_n, _, h, w = p_up.shape
    unc_plane = uncertainty[0, 0].contiguous()
    q_local = q_local.to(torch.float32).contiguous()
    prob = torch.empty((obj_n, 2 * h, 2 * w), dtype=torch.float32, device=p_up.device)
    with on_device(p_up.device):
        check(lib.vfn_urr_post(ptr(p_up), ptr(unc_plane), ptr(r1_conf), ptr(q_local), obj_n, h, w, ptr(prob),
                               stream_ptr()), 'vfn_urr_post')
    return prob


def decoder_forward(self, patch_match, r3, r2, r1=None, feature_shape=None):
    """Body for reference Decoder.forward (AFB_URR.py:208-239) with the URR block on the CUDA kernels."""
    p = self.ResMM(self.convFM(patch_match))
    p = self.RF3(r3, p)
    p = self.RF2(r2, p)
    p = self.pred2(NF.relu(p))
    # [r1 ; r1_local] (106 MB at 480p) goes into one buffer kept on the decoder: it is consumed by local_convFM right
    # below, and a fresh 100 MB allocation per frame is what the caching allocator handles worst
    bs, obj_n, h, w = feature_shape
    lm = self.__dict__.get('_vfn_local_match')
    if lm is None or tuple(lm.shape) != (obj_n, 2 * r1.shape[1], h, w) or lm.device != p.device:
        lm = torch.empty((obj_n, 2 * r1.shape[1], h, w), dtype=torch.float32, device=p.device)
        self.__dict__['_vfn_local_match'] = lm
    p_up, uncertainty, r1_conf, local_match = urr_pre(p, r1, feature_shape, out_local_match=lm)
    q = self.local_ResMM(self.local_convFM(local_match))
    q = self.local_pred2(NF.relu(q))
    return urr_post(p_up, uncertainty, r1_conf, q)


def patch_model(model, update_bank=None):
    """Make a reference AFB_URR instance use the B200 hot path: swaps `global_matcher` for the CUDA Matcher and
    binds the fused-URR Decoder.forward.  The rest of the model (encoders, KeyValue, decoder convs) is untouched."""
    from .matcher import Matcher
    ub = model.global_matcher.update_bank if update_bank is None else update_bank
    model.global_matcher = Matcher(thres_valid=model.global_matcher.thres_valid, update_bank=ub)
    model.decoder.forward = types.MethodType(decoder_forward, model.decoder)
    return model
