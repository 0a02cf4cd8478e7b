"""Mask PNG wire format: our writer/reader against a file written by the reference's
``myutils.save_seg_mask`` and read by its ``load_image_in_PIL(path, 'P')`` (tests/golden/make_golden_tail.py)."""
import io
import os

import numpy as np
import pytest

from vfloodnet_b200 import wire

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_palette_is_the_reference_palette():
    z = np.load(os.path.join(GOLD, 'tail_png.npz'))
    assert list(z['color_palette']) == wire.color_palette
    assert len(wire.color_palette) == 768


def test_our_reader_decodes_the_reference_file():
    z = np.load(os.path.join(GOLD, 'tail_png.npz'))
    plane, pal = wire.decode_seg_mask(z['png_bytes'].tobytes())
    assert np.array_equal(plane, z['decoded']) and np.array_equal(plane, z['pred'])
    assert np.array_equal(pal.ravel(), z['palette'][:pal.size])


def test_pillow_reads_our_file_like_the_reference_file(tmp_path):
    Image = pytest.importorskip('PIL.Image')
    z = np.load(os.path.join(GOLD, 'tail_png.npz'))
    path = wire.mask_path(str(tmp_path), 'frame_0001')
    assert path.endswith('frame_0001.png')
    wire.save_seg_mask(z['pred'], path)
    img = Image.open(path)
    img.load()
    assert img.mode == str(z['mode']) == 'P'
    assert np.array_equal(np.asarray(img.convert('P')), z['decoded'])
    assert np.array_equal(np.asarray(img.getpalette(), np.uint8), z['palette'])
    assert np.array_equal(wire.load_seg_mask(path), z['pred'])


@pytest.mark.parametrize('shape', [(1, 1), (3, 7), (480, 854)])
def test_round_trip_and_errors(shape):
    rng = np.random.default_rng(sum(shape))
    pred = rng.integers(0, 4, shape).astype(np.uint8)
    plane, pal = wire.decode_seg_mask(wire.encode_seg_mask(pred))
    assert np.array_equal(plane, pred) and pal.shape == (256, 3)
    with pytest.raises(ValueError):
        wire.encode_seg_mask(pred.astype(np.int32))
    with pytest.raises(ValueError):
        wire.decode_seg_mask(b'not a png')


def test_reader_handles_filtered_scanlines():
    Image = pytest.importorskip('PIL.Image')
    rng = np.random.default_rng(5)
    pred = (rng.random((40, 61)) < 0.5).astype(np.uint8) * 3
    img = Image.fromarray(pred)
    img.putpalette(wire.color_palette)
    for opt in (False, True):
        buf = io.BytesIO()
        img.save(buf, format='PNG', optimize=opt)
        assert np.array_equal(wire.decode_seg_mask(buf.getvalue())[0], pred)
