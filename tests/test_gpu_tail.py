"""GPU parity of the frame-loop tail: resize+argmax, largest connected component and the water-level
column scan through the C ABI, against (a) vectors produced by the reference's own functions (cv2 / torchvision,
tests/golden/make_golden_tail.py) and (b) the CPU oracle (oracle/tail_oracle.py) on seeded inputs up to 4K.

Bars: component labels / kept component / water levels bit-exact; arg-max bit-exact wherever the two resized class
scores differ by more than 1e-5 (inside that band the reference's own CPU and CUDA interpolation kernels disagree)."""
import os

import numpy as np
import pytest
import torch

from oracle import tail_oracle as TO

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
MARGIN = 1e-5


@pytest.fixture(scope='module')
def tail():
    from vfloodnet_b200 import tail as t
    assert torch.cuda.is_available()
    return t


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def blobs(rng, h, w, n_blob, rmax):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(n_blob):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(2, rmax)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    return m.astype(np.uint8)


def soft_mask(seed, h, w, coarse=8):
    g = torch.Generator().manual_seed(seed)
    logit = torch.nn.functional.interpolate(torch.randn(1, 2, h // coarse + 2, w // coarse + 2, generator=g) * 3,
                                            size=(h, w), mode='bicubic', align_corners=False)
    return torch.softmax(logit, dim=1)


def test_largest_component_golden_cv2(tail):
    z = np.load(os.path.join(GOLD, 'tail_cc.npz'))
    names = sorted(k[:-5] for k in z.files if k.endswith('.pred'))
    assert len(names) >= 25
    for k in names:
        mask, stats = tail.postprocessing_pred(dev(z[k + '.pred']), return_stats=True)
        assert np.array_equal(mask.cpu().numpy(), z[k + '.out']), k
        fg, comps, kept, root = stats.tolist()
        assert fg == int((z[k + '.pred'] != 0).sum()), k
        if fg:
            assert kept == int(z[k + '.out'].sum()) and root >= 0, k
        else:
            assert comps == 0 and root == -1 and z[k + '.out'].min() == 1, k


@pytest.mark.parametrize('name', ['up_2x25', 'up_odd', 'down', 'same'])
def test_resize_argmax_golden_torchvision(tail, name):
    z = np.load(os.path.join(GOLD, 'tail_resize.npz'))
    up = z[name + '.up'][0]
    pred = tail.resize_argmax(dev(z[name + '.pred_mask']), up.shape[-2:]).cpu().numpy()
    clear = np.abs(up[1] - up[0]) > MARGIN
    assert clear.mean() > 0.99
    assert np.array_equal(pred[clear], z[name + '.pred'][clear])


@pytest.mark.parametrize('antialias', [True, False])
@pytest.mark.parametrize('h,w,H,W', [(480, 854, 1080, 1920), (480, 854, 480, 854), (480, 854, 2160, 3840),
                                     (270, 480, 97, 131)])
def test_resize_argmax_vs_oracle(tail, antialias, h, w, H, W):
    pm = soft_mask(h + W, h, w)
    ref, margin = TO.resize_argmax(pm, (H, W), antialias)
    pred = tail.resize_argmax(pm.cuda(), (H, W), antialias).cpu().numpy()
    clear = margin > MARGIN
    assert clear.mean() > 0.995
    assert np.array_equal(pred[clear], ref[clear])


def test_resize_three_classes_first_max_wins(tail):
    pm = torch.zeros(1, 3, 12, 16)
    pm[0, 1] = 0.5
    pm[0, 2] = 0.5                        # exact tie between classes 1 and 2 everywhere -> argmax takes the first
    pred = tail.resize_argmax(pm.cuda(), (12, 16)).cpu().numpy()
    assert (pred == 1).all()


@pytest.mark.parametrize('h,w,kind', [(1080, 1920, 'blobs'), (1080, 1920, 'noise'), (2160, 3840, 'blobs'),
                                      (481, 853, 'noise'), (1, 37, 'noise'), (53, 1, 'noise'), (1080, 1920, 'stripes')])
def test_largest_component_vs_oracle(tail, h, w, kind):
    rng = np.random.default_rng(h * 7 + w)
    if kind == 'blobs':
        pred = blobs(rng, h, w, 40, max(3, min(h, w) // 6))
    elif kind == 'noise':
        pred = (rng.random((h, w)) < 0.55).astype(np.uint8)    # near the 8-connectivity percolation threshold
    else:
        pred = np.zeros((h, w), np.uint8)
        pred[::4, :] = 1                                       # equal-size rows: a 270-way tie
        pred[2::4, ::2] = 1
    ref = TO.postprocessing_pred(pred)
    mask, stats = tail.postprocessing_pred(dev(pred), return_stats=True)
    assert np.array_equal(mask.cpu().numpy(), ref)
    assert stats[0].item() == int(pred.sum()) and stats[2].item() == int(ref.sum())


def test_component_count_matches_cv2_order_oracle(tail):
    rng = np.random.default_rng(11)
    pred = (rng.random((300, 400)) < 0.4).astype(np.uint8)
    cnt, labels = TO.grana_order_labels(pred)
    _, stats = tail.postprocessing_pred(dev(pred), return_stats=True)
    assert stats[1].item() == cnt - 1


def test_idempotent_and_subset(tail):
    rng = np.random.default_rng(3)
    pred = dev(blobs(rng, 1080, 1920, 30, 200))
    m1 = tail.postprocessing_pred(pred)
    m2 = tail.postprocessing_pred(m1)
    assert torch.equal(m1, m2)                       # one component left: applying it again changes nothing
    assert bool((m1 <= pred).all())


@pytest.mark.parametrize('h,w,H,W,antialias', [(480, 854, 1080, 1920, True), (480, 854, 1080, 1920, False),
                                               (480, 864, 483, 857, True), (270, 480, 2160, 3840, True),
                                               (480, 854, 300, 500, True)])
def test_frame_tail_vs_oracle(tail, h, w, H, W, antialias):
    """vfn_frame_tail (fused kernels) over several frames of one stream: arg-max against the oracle outside the
    rounding band, then - from the arg-max the GPU produced - largest component and carried-over water levels exactly."""
    key_pts = [(W // 5, H // 10), (W // 2, 5), (W - 1, H - 1), (0, 0), (3 * W // 4, 2 * H // 3), (5 * W, 3)]
    ft = tail.FrameTail((H, W), key_pts, antialias=antialias)
    prev = None
    for f in range(4):
        pm = soft_mask(100 + f, h, w, coarse=60)
        ref_pred, margin = TO.resize_argmax(pm, (H, W), antialias)
        mask, levels = ft(pm.cuda())
        pred = ft.pred.cpu().numpy()
        clear = margin > MARGIN
        assert clear.mean() > 0.995
        assert np.array_equal(pred[clear], ref_pred[clear])
        want_mask = TO.postprocessing_pred(pred)
        assert np.array_equal(mask.cpu().numpy(), want_mask)
        want = np.array(TO.waterlevel_scan(want_mask, key_pts[:5], 1, prev))
        prev = list(want)
        got = levels.cpu().numpy()
        assert np.array_equal(np.isnan(got[:5]), np.isnan(want))
        assert np.array_equal(got[:5][~np.isnan(want)], want[~np.isnan(want)])
        assert got[5] == 0.0                           # column outside the image: estimate untouched
        st = ft.stats.tolist()
        assert st[0] == int((pred != 0).sum()) and st[2] == int(want_mask.sum()) or st[0] == 0


def test_waterlevel_hand_cases(tail):
    m = np.zeros((10, 6), np.uint8)
    m[7:, 2] = 1
    m[4, 3] = 1
    lib_pts = [(2, 3), (3, 3), (5, 0), (2, 9)]
    from vfloodnet_b200 import _lib
    lib = _lib.load()
    mask = dev(m)
    kp = torch.tensor(lib_pts, dtype=torch.int32).cuda()
    lv = torch.tensor([9.0, 9.0, 5.0, 3.0]).cuda()
    _lib.check(lib.vfn_tail_waterlevel(mask.data_ptr(), 10, 6, kp.data_ptr(), 4, 1, lv.data_ptr(), None))
    torch.cuda.synchronize()
    got = lv.cpu().numpy()
    want = TO.waterlevel_scan(m, lib_pts, 1, prev=[9.0, 9.0, 5.0, 3.0])
    assert got[0] == want[0] == 4.0 and np.isnan(got[1]) and np.isnan(want[1])
    assert got[2] == want[2] == 5.0 and got[3] == want[3] == 3.0


def test_tail_rejects_cpu_tensors_and_small_workspace(tail):
    with pytest.raises(RuntimeError):
        tail.postprocessing_pred(torch.zeros(4, 4, dtype=torch.uint8))
    from vfloodnet_b200 import _lib
    lib = _lib.load()
    p = torch.zeros(8, 8, dtype=torch.uint8, device='cuda')
    st = torch.zeros(4, dtype=torch.int32, device='cuda')
    ws = torch.zeros(16, dtype=torch.uint8, device='cuda')
    rc = lib.vfn_tail_largest_component(p.data_ptr(), 8, 8, p.data_ptr(), st.data_ptr(), ws.data_ptr(), 16, None)
    assert rc == -2 and b'workspace' in lib.vfn_last_error()
