"""Drop-in KeyValue head (reference: KeyValue, video_module/model/AFB_URR.py:94-111) on the tcgen05 implicit-GEMM
kernel `vfn_keyvalue` (csrc/vfn_kv.cu), SURVEY.md 8(f) n3.

Same constructor data as the reference module - it WRAPS the reference's two `nn.Conv2d` (attribute names `Key` /
`Value`, so `state_dict()` keys do not change) - and the same `forward(x) -> (key, val)` shapes
((B, keydim, H*W), (B, valdim, H*W)).  What differs:

  * both 3x3 convolutions run as one fp32-grade tensor-core GEMM (fp16 hi/lo operand splits, three passes, fp32
    accumulate: the result of a true-fp32 convolution up to summation order, whatever `cudnn.allow_tf32` says;
    passes=1 gives the TF32 class of the reference's default math);
  * the results are stored ENTRY-MAJOR ((B, H*W, d): one row per pixel, the layout of the memory read's query operand
    and of the bank's candidate rows).  The tensors handed back are transposed views with the reference's shapes;
    `vfloodnet_b200.Matcher` and `FeatureBank.update` recognise them by their strides and skip the transposes of their
    preparation step.  For B == 1 (the query frame of `segment`) the value tensor is written in the reference's
    (1, valdim, H*W) layout instead, because the read copies it unchanged into `[mem ; q_out]` (AFB_URR.py:159).

No CPU path: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from ._lib import check, on_device, ptr, stream_ptr


class KeyValueHead(nn.Module):
    def __init__(self, key_conv: nn.Conv2d, value_conv: nn.Conv2d, passes: int = 3):
        super().__init__()
        for conv in (key_conv, value_conv):
            if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.stride != (1, 1) or \
                    conv.dilation != (1, 1) or conv.groups != 1:
                raise ValueError('KeyValueHead implements 3x3 / padding 1 / stride 1 convolutions (AFB_URR.py:100-101)')
        if key_conv.in_channels != value_conv.in_channels:
            raise ValueError('Key and Value must read the same feature map')
        if passes not in (1, 3):
            raise ValueError('passes: 3 (fp32 grade) or 1 (TF32 class)')
        self.Key, self.Value = key_conv, value_conv
        self.train(key_conv.training)        # a head built around the convolutions of an eval() model is in eval mode
        self.keydim, self.valdim = key_conv.out_channels, value_conv.out_channels
        self.passes = passes
        self._packed = None
        self._packed_key = None
        self._ws = {}         # workspace per input shape (a CUDA graph captured at one shape keeps its buffer)
        self.launches = 0

    @classmethod
    def from_reference(cls, kv: nn.Module, passes: int = 3) -> 'KeyValueHead':
        """kv: a reference `KeyValue` instance (its convolutions are adopted, not copied)"""
        return cls(kv.Key, kv.Value, passes=passes)

    # packed tensor-core operands of the weights; rebuilt when a weight tensor is replaced or modified in place
    def _weights(self, lib, dev):
        ps = [self.Key.weight, self.Key.bias, self.Value.weight, self.Value.bias]
        key = tuple((p.data_ptr(), p._version) if p is not None else None for p in ps) + (str(dev),)
        if self._packed is None or self._packed_key != key:
            c_in = self.Key.in_channels
            n = lib.vfn_kv_packed_weights_bytes(c_in, self.keydim, self.valdim)
            if n == 0:
                raise ValueError('unsupported KeyValue dimensions')
            packed = torch.empty(n, dtype=torch.uint8, device=dev)
            wk, wv = self.Key.weight.detach().float().contiguous(), self.Value.weight.detach().float().contiguous()
            bk = self.Key.bias.detach().float().contiguous() if self.Key.bias is not None else None
            bv = self.Value.bias.detach().float().contiguous() if self.Value.bias is not None else None
            check(lib.vfn_kv_pack_weights(ptr(wk), ptr(bk), ptr(wv), ptr(bv), c_in, self.keydim, self.valdim, ptr(packed),
                                          stream_ptr()), 'vfn_kv_pack_weights')
            self._packed, self._packed_key = packed, key
        return self._packed

    def forward(self, x: torch.Tensor, layout: str = 'auto'):
        """x: (B, C, h, w).  layout: 'auto' (values dimension-major when B == 1, else entry-major), 'em' or 'dm' (both
        outputs in that layout).  Returns key (B, keydim, h*w), val (B, valdim, h*w)."""
        if self.training:
            # training (train_video_seg.py) needs autograd through the two convolutions: the reference's own forward
            # (AFB_URR.py:103-111).  The library path is the inference path.
            key, val = self.Key(x), self.Value(x)
            return key.view(*key.shape[:2], -1), val.view(*val.shape[:2], -1)
        if x.device.type != 'cuda':
            raise RuntimeError('KeyValueHead runs on the CUDA library only (no CPU fallback)')
        lib = _lib.load()
        b, c, h, w = x.shape
        if c != self.Key.in_channels:
            raise ValueError('input channels do not match the convolution weights')
        dev, hw = x.device, h * w
        x = x.detach().to(torch.float32).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        key_layout = 'em' if layout == 'auto' else layout
        val_layout = ('dm' if b == 1 else 'em') if layout == 'auto' else layout
        key_em = torch.empty((b, hw, self.keydim), **f32) if key_layout == 'em' else None
        key_dm = torch.empty((b, self.keydim, hw), **f32) if key_layout == 'dm' else None
        val_em = torch.empty((b, hw, self.valdim), **f32) if val_layout == 'em' else None
        val_dm = torch.empty((b, self.valdim, hw), **f32) if val_layout == 'dm' else None
        with on_device(dev):
            packed = self._weights(lib, dev)
            need = lib.vfn_keyvalue_workspace_bytes(b, c, h, w, self.keydim, self.valdim)
            ws = self._ws.get((b, h, w, str(dev)))
            if ws is None:
                ws = self._ws[(b, h, w, str(dev))] = torch.empty(need, dtype=torch.uint8, device=dev)
            check(lib.vfn_keyvalue(ptr(x), b, c, h, w, ptr(packed), self.keydim, self.valdim, self.passes, ptr(key_em),
                                   ptr(val_em), ptr(key_dm), ptr(val_dm), ptr(ws), ws.numel(), stream_ptr()),
                  'vfn_keyvalue')
        self.launches += 5
        key = key_em.transpose(1, 2) if key_em is not None else key_dm
        val = val_em.transpose(1, 2) if val_em is not None else val_dm
        return key, val
