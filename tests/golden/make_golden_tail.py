"""Golden vectors for the frame-loop tail, produced by the UNMODIFIED reference functions.

Run in the build container only:   python tests/golden/make_golden_tail.py

  * ``myutils.data.postprocessing_pred`` (myutils/data.py:19-39, cv2 CCL_GRANA) on hand-made and random predictions,
    including equal-size ties, an empty prediction, a full one and a single component.
  * ``TF.resize(pred_mask, ori_size, BICUBIC)`` + argmax exactly as test_video_seg.py:114-115 (torchvision of this image).
  * ``myutils.data.save_seg_mask`` + ``load_image_in_PIL(path, 'P')`` (myutils/data.py:51-55, est_waterlevel.py reads
    these files): the PNG bytes' decoded index plane and palette.

Output (committed): tests/golden/tail_cc.npz, tail_resize.npz, tail_png.npz
"""
import io
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_shims  # noqa: E402


def blobs(rng, h, w, n_blob, rmax):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(n_blob):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(1, rmax)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    return m.astype(np.uint8)


def main():
    install_shims()
    import myutils
    from myutils.data import postprocessing_pred, save_seg_mask, color_palette
    from torchvision.transforms import functional as TF, InterpolationMode

    rng = np.random.default_rng(7)
    cases = {}
    # hand-made: ties between equally large components in several block alignments
    t = np.zeros((6, 9), np.uint8); t[0, 5:7] = 1; t[1, 0:2] = 1                 # row-0 component starts in a later block
    cases['tie_block_order'] = t
    t = np.zeros((7, 7), np.uint8); t[0, 6] = 1; t[2, 0] = 1; t[4, 3] = 1
    cases['tie_singletons'] = t
    t = np.zeros((5, 8), np.uint8); t[1, 1] = 1; t[0, 2] = 1; t[3, 4:6] = 1       # diagonal (8-connectivity) vs a run
    cases['tie_diagonal'] = t
    cases['empty'] = np.zeros((9, 11), np.uint8)
    cases['full'] = np.ones((8, 5), np.uint8)
    cases['one_component'] = blobs(rng, 24, 31, 1, 9)
    t = np.zeros((12, 12), np.uint8); t[::2, ::2] = 1
    cases['isolated_grid'] = t
    t = np.ones((16, 16), np.uint8); t[5, :] = 0; t[:, 9] = 0; t[5, 9] = 1; t[4, 9] = 0
    cases['cross_cut'] = t
    t = np.zeros((20, 30), np.uint8)
    for y in range(20):                                                          # spiral-ish snake: long union chains
        if y % 4 == 0: t[y, :] = 1
        if y % 4 == 1: t[y, -1] = 1
        if y % 4 == 2: t[y, :] = 1
        if y % 4 == 3: t[y, 0] = 1
    t[8, :] = 0
    cases['snake'] = t
    for i in range(12):
        h, w = int(rng.integers(3, 70)), int(rng.integers(3, 90))
        cases[f'noise{i}'] = (rng.random((h, w)) < [0.3, 0.45, 0.55, 0.7][i % 4]).astype(np.uint8)
    for i in range(4):
        cases[f'blobs{i}'] = blobs(rng, 135, 240, 14, 30)
    cc = {}
    for k, p in cases.items():
        cc[k + '.pred'] = p
        cc[k + '.out'] = postprocessing_pred(p.copy())
    np.savez_compressed(os.path.join(HERE, 'tail_cc.npz'), **cc)

    # resize + argmax (test_video_seg.py:114-115)
    g = torch.Generator().manual_seed(3)
    rz = {}
    for name, (h, w, oh, ow) in {'up_2x25': (48, 86, 108, 192), 'up_odd': (30, 53, 67, 121), 'down': (60, 80, 25, 33),
                                 'same': (20, 24, 20, 24)}.items():
        logit = torch.nn.functional.interpolate(torch.randn(1, 2, h // 4 + 1, w // 4 + 1, generator=g) * 3, size=(h, w),
                                                mode='bilinear', align_corners=False)
        pm = torch.softmax(logit, dim=1)
        up = TF.resize(pm, (oh, ow), InterpolationMode.BICUBIC)
        rz[name + '.pred_mask'] = pm.numpy()
        rz[name + '.up'] = up.numpy()
        rz[name + '.pred'] = torch.argmax(up[0], dim=0).numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, 'tail_resize.npz'), **rz)

    # PNG wire format (myutils/data.py:51-55 writer, myutils load_image_in_PIL reader)
    pred = cases['blobs0']
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'a.png')
        save_seg_mask(pred, path, color_palette)
        img = myutils.load_image_in_PIL(path, 'P')
        raw = open(path, 'rb').read()
        png = {'pred': pred, 'decoded': np.asarray(img), 'palette': np.asarray(img.getpalette(), np.uint8),
               'color_palette': np.asarray(color_palette, np.int32), 'mode': np.array(img.mode),
               'png_bytes': np.frombuffer(raw, np.uint8)}
    np.savez_compressed(os.path.join(HERE, 'tail_png.npz'), **png)
    print('wrote tail_cc.npz tail_resize.npz tail_png.npz')


if __name__ == '__main__':
    main()
