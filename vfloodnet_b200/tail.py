"""Device-resident tail of the reference's frame loop.

After ``fb.update`` the reference resizes the soft mask to the original frame size, takes the arg-max, copies the
full-resolution prediction to the host, keeps its largest connected component with OpenCV and writes a PNG
(test_video_seg.py:114-118, myutils/data.py:19-39); the estimation layer later re-reads that PNG and scans one column
per reference object for the water line (estimation/reference_tracking.py:190-204).  Here the same four steps run as
CUDA kernels on the frame's stream and only the water levels (a few floats) need to leave the GPU; the mask stays
available as a device tensor for callers that still want the PNG (``vfloodnet_b200.wire``).

Same names as the reference where a function is replaced: ``postprocessing_pred``.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib
from ._lib import check, on_device, ptr, stream_ptr

water_label_id = 1   # estimation/reference_tracking.py:21


def key_points_from_bbox(ref_bbox) -> list:
    """(x, y, w, h) boxes of the reference objects -> key points (bottom centre), reference_tracking.py:192-195."""
    pts = []
    for box in ref_bbox:
        x, y, w, h = [int(v) for v in box]
        pts.append((int(x + w / 2), int(y + h)))
    return pts


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'{what}: expected a CUDA tensor (there is no CPU path)')


def resize_argmax(pred_mask: torch.Tensor, ori_size, antialias: bool = True) -> torch.Tensor:
    """``argmax(TF.resize(pred_mask, ori_size, BICUBIC)[0], dim=0)`` as uint8 (H, W), on the device."""
    _need_cuda(pred_mask, 'resize_argmax')
    if pred_mask.dim() == 4:
        if pred_mask.shape[0] != 1:
            raise ValueError('resize_argmax implements the inference path (bs == 1)')
        pred_mask = pred_mask[0]
    obj_n, h, w = pred_mask.shape
    H, W = int(ori_size[0]), int(ori_size[1])
    src = pred_mask.to(torch.float32).contiguous()
    pred = torch.empty((H, W), dtype=torch.uint8, device=src.device)
    with on_device(src.device):
        check(_lib.load().vfn_tail_resize_argmax(ptr(src), obj_n, h, w, H, W, int(bool(antialias)), ptr(pred),
                                                 stream_ptr()), 'vfn_tail_resize_argmax')
    return pred


def postprocessing_pred(pred: torch.Tensor, return_stats: bool = False):
    """myutils.postprocessing_pred for a binary uint8 (H, W) CUDA tensor: largest 8-connected component, ties resolved in
    OpenCV's label order, empty prediction -> all ones.  ``return_stats``: also an int32[4] device tensor
    {foreground pixels, components, kept size, kept root id}."""
    _need_cuda(pred, 'postprocessing_pred')
    if pred.dtype != torch.uint8 or pred.dim() != 2:
        raise ValueError('postprocessing_pred expects a uint8 (H, W) prediction')
    lib = _lib.load()
    pred = pred.contiguous()
    H, W = pred.shape
    ws_bytes = lib.vfn_tail_workspace_bytes(H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pred.device)
    mask = torch.empty_like(pred)
    stats = torch.empty(4, dtype=torch.int32, device=pred.device)
    with on_device(pred.device):
        check(lib.vfn_tail_largest_component(ptr(pred), H, W, ptr(mask), ptr(stats), ptr(ws), ws_bytes, stream_ptr()),
              'vfn_tail_largest_component')
    return (mask, stats) if return_stats else mask


class FrameTail:
    """One video stream's tail state: output size, key points, the running water-level estimates and the scratch memory.

    ``tail(pred_mask)`` -> ``(mask, levels)``: ``mask`` uint8 (H, W) and ``levels`` float32 (n_pts) device tensors, valid
    until the next call.  ``levels`` persists between frames exactly like ``waterlevel_list[-1]`` in the reference
    (a frame whose column shows no water keeps the previous estimate; the initial estimate is 0;
    a level of 1 px is NaN).  Nothing synchronises with the host; read ``levels`` when (and if) it is needed.
    """

    def __init__(self, ori_size: Tuple[int, int], key_pts: Sequence[Tuple[int, int]] = (), device='cuda',
                 antialias: bool = True, label_id: int = water_label_id):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('FrameTail needs a CUDA device (there is no CPU path)')
        self.lib = _lib.load()
        self.H, self.W = int(ori_size[0]), int(ori_size[1])
        self.antialias = bool(antialias)
        self.label_id = int(label_id)
        self.n_pts = len(key_pts)
        kp = torch.tensor([[int(x), int(y)] for x, y in key_pts], dtype=torch.int32).reshape(-1, 2)
        self.key_pts = kp.to(self.device)
        self.levels = torch.zeros(max(self.n_pts, 1), dtype=torch.float32, device=self.device)[:self.n_pts]
        self.ws_bytes = self.lib.vfn_tail_workspace_bytes(self.H, self.W)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.pred = torch.empty((self.H, self.W), dtype=torch.uint8, device=self.device)
        self.mask = torch.empty_like(self.pred)
        self.stats = torch.empty(4, dtype=torch.int32, device=self.device)

    def __call__(self, pred_mask: torch.Tensor):
        _need_cuda(pred_mask, 'FrameTail')
        if pred_mask.dim() == 4:
            if pred_mask.shape[0] != 1:
                raise ValueError('FrameTail implements the inference path (bs == 1)')
            pred_mask = pred_mask[0]
        obj_n, h, w = pred_mask.shape
        src = pred_mask.to(torch.float32).contiguous()
        with on_device(self.device):
            check(self.lib.vfn_frame_tail(ptr(src), obj_n, h, w, self.H, self.W, int(self.antialias),
                                          ptr(self.key_pts) if self.n_pts else None, self.n_pts, self.label_id,
                                          ptr(self.pred), ptr(self.mask), ptr(self.stats),
                                          ptr(self.levels) if self.n_pts else None, ptr(self.ws), self.ws_bytes,
                                          stream_ptr()), 'vfn_frame_tail')
        return self.mask, self.levels
