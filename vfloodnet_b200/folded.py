"""Encoder throughput (reference: EncoderM / EncoderQ, video_module/model/AFB_URR.py:33-93).

In inference (`model.eval()`, test_video_seg.py:88) every BatchNorm of the two ResNet-50 trunks is a fixed per-channel
affine map.  The reference runs it as its own kernel after every convolution, followed by a ReLU kernel and, at the
end of a bottleneck, an add kernel: ~170 launches per encoder pass, each re-reading and re-writing the activation.
`fold_encoders` keeps the reference's modules and parameters untouched and binds forward passes that evaluate the SAME
function with the affine map folded into the preceding convolution (`w' = w * g / sqrt(var + eps)`,
`b' = beta - mean * g / sqrt(var + eps)`), and convolution + bias + ReLU (+ residual add) issued as ONE cuDNN call
(`torch.cudnn_convolution_relu` / `cudnn_convolution_add_relu`): 53 launches per encoder pass.  EncoderM's three stem
convolutions (frame, mask, background: AFB_URR.py:56) are one convolution over the 5 stacked input channels.

The folded weights are cached and rebuilt when a parameter or a BatchNorm statistic changes (tensor versions).  The
results differ from the unfolded modules by convolution rounding only (the scale is applied to the weights before
the products instead of after the sum); tests/test_gpu_keyvalue.py compares them and the masks.
cuDNN stays the convolution engine: this file is launch structure, not a kernel.
"""
from __future__ import annotations

import types

import torch
from torch.nn import functional as NF


def _fold(conv, bn):
    """(w', b') of bn(conv(x)) in eval mode"""
    g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * g.view(-1, 1, 1, 1)
    b = bn.bias - bn.running_mean * g
    if conv.bias is not None:
        b = b + conv.bias * g
    return w.detach().contiguous(), b.detach().contiguous()


def _conv_relu(x, w, b, stride, padding):
    return torch.cudnn_convolution_relu(x, w, b, stride, padding, (1, 1), 1)


def _conv_add_relu(x, w, z, b, stride, padding):
    return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, stride, padding, (1, 1), 1)


class _FoldedBottleneck:
    """torchvision Bottleneck.forward (conv1-bn1-relu, conv2-bn2-relu, conv3-bn3, + identity / downsample, relu)"""

    def __init__(self, blk):
        self.c1 = _fold(blk.conv1, blk.bn1)
        self.c2 = _fold(blk.conv2, blk.bn2)
        self.c3 = _fold(blk.conv3, blk.bn3)
        self.s2, self.p2 = tuple(blk.conv2.stride), tuple(blk.conv2.padding)
        self.down = None
        if blk.downsample is not None:
            self.down = _fold(blk.downsample[0], blk.downsample[1])
            self.sd = tuple(blk.downsample[0].stride)

    def __call__(self, x):
        o = _conv_relu(x, *self.c1, (1, 1), (0, 0))
        o = _conv_relu(o, *self.c2, self.s2, self.p2)
        idn = x if self.down is None else NF.conv2d(x, self.down[0], self.down[1], self.sd)
        return _conv_add_relu(o, self.c3[0], idn, self.c3[1], (1, 1), (0, 0))


class _FoldedTrunk:
    def __init__(self, enc):
        self.enc = enc
        self.key = None
        self._tensors = None

    def _version(self):
        # in-place updates (load_state_dict, optimiser steps, BatchNorm statistics) bump the tensors' versions; a move of
        # the module (`.to(device)`) re-points every parameter, which the first one's pointer shows.  The tensor list is
        # walked once: parameters() / buffers() cost 0.7 ms per call on this trunk.
        if self._tensors is None:
            self._tensors = list(self.enc.parameters()) + list(self.enc.buffers())
        ts = self._tensors
        return (ts[0].data_ptr(), len(ts)) + tuple(t._version for t in ts)

    def refresh(self):
        key = self._version()
        if key == self.key:
            return
        e = self.enc
        with torch.no_grad():
            if hasattr(e, 'conv1_m'):      # EncoderM: conv1(f) + conv1_m(m) + conv1_o(o) = one conv over 5 channels
                class _Cat:                 # the three stems share kernel size / stride / padding (AFB_URR.py:36-40)
                    pass
                cat = _Cat()
                cat.weight = torch.cat([e.conv1.weight, e.conv1_m.weight, e.conv1_o.weight], dim=1)
                cat.bias = None
                self.stem = _fold(cat, e.bn1)
            else:
                self.stem = _fold(e.conv1, e.bn1)
            self.stem_sp = (tuple(e.conv1.stride), tuple(e.conv1.padding))
            self.layers = [[_FoldedBottleneck(b) for b in layer] for layer in (e.res2, e.res3, e.res4)]
        self.key = key

    def trunk(self, x):
        """stem input (already normalised / stacked) -> r4, r3, r2, r1"""
        r1 = _conv_relu(x, *self.stem, *self.stem_sp)
        x = self.enc.maxpool(r1)
        outs = []
        for layer in self.layers:
            for blk in layer:
                x = blk(x)
            outs.append(x)
        r2, r3, r4 = outs
        return r4, r3, r2, r1


def _encoder_q_forward(self, in_f):
    """EncoderQ.forward (AFB_URR.py:81-92)"""
    if self.training or in_f.device.type != 'cuda':      # the fused cuDNN calls exist on CUDA only
        return self._vfn_ref_forward(in_f)
    t = self._vfn_folded
    t.refresh()
    f = (in_f - self.mean) / self.std
    return t.trunk(f)


def _encoder_m_forward(self, in_f, in_m, in_o):
    """EncoderM.forward (AFB_URR.py:53-64)"""
    if self.training or in_f.device.type != 'cuda':
        return self._vfn_ref_forward(in_f, in_m, in_o)
    t = self._vfn_folded
    t.refresh()
    f = (in_f - self.mean) / self.std
    r4, _r3, _r2, r1 = t.trunk(torch.cat([f, in_m, in_o], dim=1))
    return r4, r1


def fold_encoders(model):
    """bind the folded forward passes on model.encoder_q / model.encoder_m (inference only; train mode falls through to
    the reference's own forward).  Parameters, buffers and state_dict are untouched.  Returns the model."""
    for enc, fwd in ((model.encoder_q, _encoder_q_forward), (model.encoder_m, _encoder_m_forward)):
        if '_vfn_folded' in enc.__dict__:
            continue
        enc.__dict__['_vfn_ref_forward'] = enc.forward
        enc.__dict__['_vfn_folded'] = _FoldedTrunk(enc)
        enc.forward = types.MethodType(fwd, enc)
    return model
