"""The frame-tail oracle (oracle/tail_oracle.py) against vectors produced by the reference's own functions
(tests/golden/make_golden_tail.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import tail_oracle as T

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _cases(npz, suffix):
    return sorted(k[:-len(suffix)] for k in npz.files if k.endswith(suffix))


def test_postprocessing_matches_reference_cv2_on_every_case():
    z = np.load(os.path.join(GOLD, 'tail_cc.npz'))
    names = _cases(z, '.pred')
    assert len(names) >= 25
    for k in names:
        out = T.postprocessing_pred(z[k + '.pred'])
        assert out.dtype == np.uint8
        assert np.array_equal(out, z[k + '.out']), k


def test_empty_prediction_becomes_all_water_like_the_reference():
    z = np.load(os.path.join(GOLD, 'tail_cc.npz'))
    assert z['empty.out'].min() == 1                       # reference quirk, myutils/data.py:27-37
    assert T.postprocessing_pred(z['empty.pred']).min() == 1


def test_tie_goes_to_first_block_not_first_pixel():
    z = np.load(os.path.join(GOLD, 'tail_cc.npz'))
    out = z['tie_block_order.out']
    # the row-1 component sits in block column 0 and wins over the row-0 component that starts in a later block
    assert out[1, 0] == 1 and out[0, 5] == 0


@pytest.mark.parametrize('name', ['up_2x25', 'up_odd', 'down', 'same'])
def test_resize_argmax_matches_torchvision(name):
    z = np.load(os.path.join(GOLD, 'tail_resize.npz'))
    pm = torch.from_numpy(z[name + '.pred_mask'])
    size = z[name + '.up'].shape[-2:]
    up = T.resize_bicubic(pm, size)
    assert (up.numpy() - z[name + '.up']).__abs__().max() <= 1e-6
    pred, margin = T.resize_argmax(pm, size)
    clear = margin > 1e-6
    assert np.array_equal(pred[clear], z[name + '.pred'][clear])


def test_waterlevel_scan_hand_cases():
    m = np.zeros((10, 6), np.uint8)
    m[7:, 2] = 1          # water from row 7 in column 2
    m[4, 3] = 1           # directly below key point (3, 3): level 1 -> NaN
    est = T.waterlevel_scan(m, [(2, 3), (3, 3), (5, 0)], 1, prev=[9.0, 9.0, 5.0])
    assert est[0] == 4.0
    assert np.isnan(est[1])
    assert est[2] == 5.0   # no water in the column: previous estimate kept
    assert T.waterlevel_scan(m, [(5, 0)]) == [0.0]
    assert T.waterlevel_scan(m, [(2, 9)], prev=[3.0]) == [3.0]      # key point on the last row: empty range
