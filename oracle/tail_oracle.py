"""CPU oracle for the frame-loop tail.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
legs may import it; ``vfloodnet_b200`` never does.

Restates, on the CPU, what the reference does to every frame after ``fb.update`` (all under /root/reference):
  test_video_seg.py:114-115          pred = argmax(TF.resize(pred_mask, ori_size, BICUBIC)[0], dim=0) as uint8
  myutils/data.py:19-39              postprocessing_pred: 8-connected components (cv2, CCL_GRANA), keep the largest
  estimation/reference_tracking.py:190-204   per key point: first water pixel below it in its column -> level in px

Parity pinning: the reference has no tests for these functions.  ``tests/golden/make_golden_tail.py`` runs the
reference's own ``myutils.data.postprocessing_pred`` (cv2) and torchvision's ``TF.resize`` in the build container and
commits inputs/outputs as ``tests/golden/tail_*.npz``; ``tests/test_oracle_tail.py`` holds this file to them bit-exactly
(labels) / at 1e-6 (resize).  The water-level scan is embedded in a 150-line CLI function that reads PNGs and a tracker,
so it cannot be imported on its own: its loop is restated below line by line ("parity unpinned" upstream; pinned here by
hand-computed cases).

Third-party pieces on this path and how they are restated:
  * ``TF.resize`` on a tensor (torchvision >= 0.17, antialias default True) = ``F.interpolate(mode='bicubic',
    align_corners=False, antialias=True)``: Pillow-style separable cubic, a = -0.5, window truncated at the borders and
    re-normalised.  torchvision 0.9.1 (the reference's README pin) has no antialias flag: a = -0.75, clamped indices.
    Both are offered (``antialias=``); the default follows the torchvision installed in this image.
  * ``cv2.connectedComponentsWithAlgorithm(.., 8, CV_32S, CCL_GRANA)`` (OpenCV 4.13 in this image): labels are numbered
    in raster order of the first 2x2 block a component touches (block-based decision tree, provisional labels are
    created block by block and flattened in increasing order).  Only this ORDER matters to the caller (ties between
    equally large components go to the lowest label); it is reproduced with ``scipy.ndimage.label`` + that key.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as NF
from scipy import ndimage

_S8 = np.ones((3, 3), dtype=np.int32)


def resize_bicubic(pred_mask: torch.Tensor, ori_size, antialias: bool = True) -> torch.Tensor:
    """``TF.resize(pred_mask, ori_size, InterpolationMode.BICUBIC)`` for a float (1, obj_n, h, w) tensor
    (test_video_seg.py:114)."""
    return NF.interpolate(pred_mask.float(), size=tuple(int(v) for v in ori_size), mode='bicubic', align_corners=False,
                          antialias=antialias)


def resize_argmax(pred_mask: torch.Tensor, ori_size, antialias: bool = True):
    """test_video_seg.py:114-115.  Returns (pred uint8 (H, W), margin (H, W)): margin = top1 - top2 of the resized
    scores, so a test can exclude pixels whose arg-max is inside fp32 rounding noise."""
    up = resize_bicubic(pred_mask, ori_size, antialias)[0]
    pred = torch.argmax(up, dim=0).numpy().astype(np.uint8)
    top = up.topk(k=min(2, up.shape[0]), dim=0).values
    margin = (top[0] - top[1]) if up.shape[0] > 1 else torch.full_like(top[0], np.inf)
    return pred, margin.numpy()


def grana_order_labels(pred: np.ndarray):
    """8-connected components of ``pred != 0`` numbered like cv2's CCL_GRANA: by the block-raster position of the first
    2x2 block of each component.  Returns (label_cnt including background, labels int32)."""
    sl, n = ndimage.label(pred != 0, _S8)
    if n == 0:
        return 1, np.zeros(pred.shape, np.int32)
    ys, xs = np.nonzero(sl)
    wb = (pred.shape[1] + 1) // 2
    key = np.full(n + 1, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(key, sl[ys, xs], (ys // 2).astype(np.int64) * wb + xs // 2)
    order = np.argsort(key[1:], kind='stable') + 1
    rank = np.zeros(n + 1, np.int32)
    rank[order] = np.arange(1, n + 1, dtype=np.int32)
    return n + 1, rank[sl]


def postprocessing_pred(pred: np.ndarray) -> np.ndarray:
    """myutils/data.py:19-39 for a binary prediction (obj_n == 2 always, Water_DS.py:93-94).

    label_cnt == 2 (one component): the prediction itself.  Otherwise the largest foreground component, ties to the
    lowest label.  With NO foreground (label_cnt == 1) the reference's loop skips label 0, leaves ``max_label = 0`` and
    returns ``labels == 0``: an all-ones mask.  Kept as is.
    """
    label_cnt, labels = grana_order_labels(pred)
    if label_cnt == 2:
        out = labels if labels[0, 0] == pred[0, 0] else 1 - labels
        return out.astype(np.uint8)
    max_cnt, max_label = 0, 0
    sizes = np.bincount(labels.ravel(), minlength=label_cnt)
    for i in range(1, label_cnt):          # label 0 is the background: `pred[mask][0] == 0` -> continue
        if sizes[i] > max_cnt:
            max_cnt, max_label = int(sizes[i]), i
    return (labels == max_label).astype(np.uint8)


def waterlevel_scan(water_mask: np.ndarray, key_pts, water_label_id: int = 1, prev=None):
    """estimation/reference_tracking.py:190-204: for key point (x, y) walk down the column from y + 1; the first pixel
    carrying the water label gives ``level = y' - y``; a level of exactly 1 is recorded as NaN; if the column holds no
    water the previous frame's estimate is kept (``copy.deepcopy(waterlevel_list[-1])``, initially 0)."""
    est = [0.0] * len(key_pts) if prev is None else list(prev)
    for t, (kx, ky) in enumerate(key_pts):
        for y in range(int(ky) + 1, water_mask.shape[0]):
            if water_mask[y][int(kx)] == water_label_id:
                est[t] = float(y - int(ky))
                if est[t] == 1:
                    est[t] = float('nan')
                break
    return est


def frame_tail(pred_mask: torch.Tensor, ori_size, key_pts, prev=None, antialias: bool = True):
    """The whole tail of one frame: resize -> argmax -> largest component -> water levels."""
    pred, margin = resize_argmax(pred_mask, ori_size, antialias)
    mask = postprocessing_pred(pred)
    return mask, waterlevel_scan(mask, key_pts, 1, prev), pred, margin
