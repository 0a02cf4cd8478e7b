"""The `segment` / `memorize` glue either side of the hot path, fused with the boundary (reference:
AFB_URR.segment / AFB_URR.memorize / Refine.forward, video_module/model/AFB_URR.py:113-127,255-318).

What the reference does around the read and the decoder, and what changes here (`fuse_model`):

  * AFB_URR.py:289-296 `expand(...).reshape(...)` materialises one copy of r3, r2 and r1 PER OBJECT (r1 alone is
    26.5 MB per object at 480p) although every object sees the same query-frame features.  Here nothing is expanded:
    the URR kernels read r1 through a zero object stride, and the skip branches of the two `Refine` blocks
    (`ResFS(convFS(f))`, AFB_URR.py:121 - a function of the query frame only) are evaluated ONCE at batch 1 and
    broadcast-added to the per-object up-sampled stream (`s + interpolate(pm)`, :122).  At 2 objects that removes
    3 of the 5 convolutions per object of each Refine block for the second object (122 of the decoder's ~420 GFLOP at
    480p) - same modules, same weights, same arithmetic per element; only the batch the shared convolutions run at
    differs (cuDNN may pick another algorithm: results agree to convolution rounding, tested).
  * KeyValue (AFB_URR.py:94-111): `KeyValueHead` (keyvalue.py) computes both 3x3 convolutions as one tcgen05 implicit
    GEMM with fp32-grade operand splits and writes keys / values ENTRY-MAJOR - the layout the read's query operand and
    the bank update's candidate rows have - so the (d, HW) -> (HW, d) transposes of the preparation kernels disappear.
    The tensors handed back keep the reference's shapes ((B, 128, HW), (B, 512, HW)): they are transposed VIEWS of the
    entry-major storage, which `Matcher` and `FeatureBank.update` recognise by their strides.

`fuse_model(model)` is opt-in on top of `patch_model` (which swaps only the matcher and the URR block): the unfused
patched model stays the reference's own module graph, and tests compare the two.
"""
from __future__ import annotations

import types

import torch
from torch.nn import functional as NF

from .urr import urr_post, urr_pre


def pad_divide_by16(x: torch.Tensor):
    """myutils.pad_divide_by([x], 16, x.shape[-2:]) (myutils/data.py:134-151) for one tensor: the padding is split
    `lw = int(delta / 2)`, `uw = delta - lw`; returns (padded, (lw, uw, lh, uh))"""
    h, w = x.shape[-2:]
    nh, nw = (h + 15) // 16 * 16, (w + 15) // 16 * 16
    lh, lw = int((nh - h) / 2), int((nw - w) / 2)
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return NF.pad(x, pad), pad


def refine_shared(rf, s, pm):
    """Refine.forward (AFB_URR.py:120-126) with the skip branch `s = ResFS(convFS(f))` given (shape (1, C, h, w),
    shared by all objects) and the per-object stream pm (obj_n, C, h/2, w/2)"""
    m = s + NF.interpolate(pm, scale_factor=rf.scale_factor, mode='bilinear', align_corners=False)
    return rf.ResMM(m)


def refine_skip(rf, f):
    """the object-independent half of Refine.forward (AFB_URR.py:121)"""
    return rf.ResFS(rf.convFS(f))


def decoder_trunk_shared(dec, patch_match, s3, s2):
    """Decoder.forward up to the coarse logits (AFB_URR.py:209-212) with the Refine skip branches pre-computed"""
    p = dec.ResMM(dec.convFM(patch_match))
    p = refine_shared(dec.RF3, s3, p)
    p = refine_shared(dec.RF2, s2, p)
    return dec.pred2(NF.relu(p))


def decoder_local(dec, p, r1, obj_n):
    """the URR block (AFB_URR.py:214-237) on the CUDA kernels; r1 is the UN-expanded (1, 64, h, w) encoder output"""
    _, c, h, w = r1.shape
    lm = dec.__dict__.get('_vfn_local_match')
    if lm is None or tuple(lm.shape) != (obj_n, 2 * c, h, w) or lm.device != p.device:
        lm = torch.empty((obj_n, 2 * c, h, w), dtype=torch.float32, device=p.device)
        dec.__dict__['_vfn_local_match'] = lm
    p_up, unc, conf, local_match = urr_pre(p, r1.expand(obj_n, -1, -1, -1), (1, obj_n, h, w), out_local_match=lm)
    q = dec.local_ResMM(dec.local_convFM(local_match))
    q = dec.local_pred2(NF.relu(q))
    return urr_post(p_up, unc, conf, q)


def finish_score(prob, obj_n, pad):
    """AFB_URR.py:302-316 (inference branch): probabilities -> clamped logits, padding removed"""
    score = prob.view(1, obj_n, *prob.shape[-2:])
    score = torch.clamp(score, 1e-7, 1 - 1e-7)
    score = torch.log(score / (1 - score))
    if pad[2] + pad[3] > 0:
        score = score[:, :, pad[2]:score.shape[2] - pad[3], :]
    if pad[0] + pad[1] > 0:
        score = score[:, :, :, pad[0]:score.shape[3] - pad[1]]
    return score


def segment_fused(self, frame, fb_global):
    """Body for AFB_URR.segment (AFB_URR.py:274-318), inference branch, without per-object copies; the training branch
    (uncertainty loss, bs > 1) is the reference's own method."""
    if self.training or frame.shape[0] != 1:
        return self._vfn_ref_segment(frame, fb_global)
    obj_n = fb_global.obj_n
    frame, pad = pad_divide_by16(frame)
    r4, r3, r2, r1 = self.encoder_q(frame)
    k4, v4 = self.keyval_r4(r4)
    res_global = self.global_matcher(fb_global, k4, v4)                       # (1, obj_n, 1024, HW)
    res_global = res_global.reshape(obj_n, res_global.shape[2], r4.shape[2], r4.shape[3])
    dec = self.decoder
    p = decoder_trunk_shared(dec, res_global, refine_skip(dec.RF3, r3), refine_skip(dec.RF2, r2))
    prob = decoder_local(dec, p, r1, obj_n)
    return finish_score(prob, obj_n, pad), None


def fuse_model(model, keyvalue: bool = True, keyvalue_passes: int = 3, fold_bn: bool = False):
    """On top of `patch_model`: bind the copy-free `segment` glue, (keyvalue=True) replace `model.keyval_r4` by the
    tcgen05 `KeyValueHead` built from its weights, and (fold_bn=True) bind the encoders' forward passes
    with BatchNorm folded into the convolutions and conv + bias + ReLU (+ add) as single cuDNN calls
    (vfloodnet_b200.folded).  Returns the model."""
    from .urr import patch_model
    from .matcher import Matcher
    if not isinstance(model.global_matcher, Matcher):
        patch_model(model)
    if '_vfn_ref_segment' not in model.__dict__:
        model.__dict__['_vfn_ref_segment'] = model.segment
    model.segment = types.MethodType(segment_fused, model)
    if keyvalue:
        from .keyvalue import KeyValueHead
        if not isinstance(model.keyval_r4, KeyValueHead):
            model.keyval_r4 = KeyValueHead.from_reference(model.keyval_r4, passes=keyvalue_passes).train(model.training)
    if fold_bn:
        from .folded import fold_encoders
        fold_encoders(model)
    return model
