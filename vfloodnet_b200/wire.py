"""Mask wire format between the video path and the estimation layer.

The reference writes every frame's prediction as an 8-bit palette PNG named after the frame
(``myutils.save_seg_mask``, myutils/data.py:51-55; test_video_seg.py:64-66,117-118) and ``est_waterlevel.py`` reads it
back with ``load_image_in_PIL(path, 'P')`` (est_waterlevel.py:26-28, reference_tracking.py:167).  What a reader sees is
the index plane (class ids) and the 256-entry palette; this module writes exactly that: colour type 3, bit depth 8,
one PLTE chunk carrying ``color_palette``, filter 0 scanlines, zlib stream.  The encoder is self-contained (zlib only)
so a stream worker does not need Pillow; any PNG reader - Pillow included - decodes the same indices and palette.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

# myutils/data.py:16 - background black, class 1 (water) dark blue, 2 green, 3 dark red, the rest grey
color_palette = [0, 0, 0, 0, 0, 128, 0, 128, 0, 128, 0, 0] + [100, 100, 100] * 252

_SIG = b'\x89PNG\r\n\x1a\n'


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)


def encode_seg_mask(pred, palette=color_palette, level: int = 6) -> bytes:
    """uint8 (H, W) class-id plane -> bytes of a palette PNG."""
    if hasattr(pred, 'detach'):                       # torch tensor (device or host): one D2H copy of H*W bytes
        pred = pred.detach().cpu().numpy()
    pred = np.ascontiguousarray(pred)
    if pred.dtype != np.uint8 or pred.ndim != 2:
        raise ValueError('save_seg_mask expects a uint8 (H, W) array')
    pal = bytes(int(v) & 0xff for v in palette)
    if len(pal) % 3 or not 3 <= len(pal) <= 768:
        raise ValueError('palette must hold 1..256 RGB triples')
    h, w = pred.shape
    rows = np.zeros((h, w + 1), np.uint8)             # filter type 0 in front of every scanline
    rows[:, 1:] = pred
    ihdr = struct.pack('>IIBBBBB', w, h, 8, 3, 0, 0, 0)
    return _SIG + _chunk(b'IHDR', ihdr) + _chunk(b'PLTE', pal) + _chunk(b'IDAT', zlib.compress(rows.tobytes(), level)) \
        + _chunk(b'IEND', b'')


def save_seg_mask(pred, seg_path, palette=color_palette):
    """Drop-in for ``myutils.save_seg_mask(pred, seg_path, palette)``."""
    data = encode_seg_mask(pred, palette)
    with open(seg_path, 'wb') as f:
        f.write(data)


def decode_seg_mask(data: bytes):
    """Inverse of :func:`encode_seg_mask` for files of this writer's shape (8-bit palette, no interlace; any filter).
    Returns (index plane uint8 (H, W), palette uint8 (n, 3)).  Used to read masks back without Pillow."""
    if data[:8] != _SIG:
        raise ValueError('not a PNG file')
    pos, idat, pal, hdr = 8, [], None, None
    while pos < len(data):
        n, tag = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if struct.unpack('>I', data[pos + 8 + n:pos + 12 + n])[0] != (zlib.crc32(tag + body) & 0xffffffff):
            raise ValueError(f'bad CRC in chunk {tag!r}')
        pos += 12 + n
        if tag == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', body)
        elif tag == b'PLTE':
            pal = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif tag == b'IDAT':
            idat.append(body)
        elif tag == b'IEND':
            break
    if hdr is None or pal is None:
        raise ValueError('missing IHDR/PLTE')
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or ctype != 3 or interlace != 0:
        raise ValueError('only 8-bit, non-interlaced palette PNGs are supported')
    raw = np.frombuffer(zlib.decompress(b''.join(idat)), np.uint8).reshape(h, w + 1)
    out = np.zeros((h, w), np.uint8)
    prev = np.zeros(w, np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 0xff
        elif f == 1:
            cur = np.cumsum(line) & 0xff                       # bpp = 1: left neighbour
        else:                                                   # Average / Paeth: sequential in x
            cur = np.zeros(w, np.int32)
            for x in range(w):
                a = int(cur[x - 1]) if x else 0
                b = int(prev[x])
                c = int(prev[x - 1]) if x else 0
                if f == 3:
                    pr = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pr = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
                cur[x] = (int(line[x]) + pr) & 0xff
        out[y] = cur
        prev = cur
    return out, pal


def load_seg_mask(path):
    """Index plane of a mask PNG, as ``np.asarray(load_image_in_PIL(path, 'P'))`` gives for the reference's files."""
    with open(path, 'rb') as f:
        return decode_seg_mask(f.read())[0]


def mask_path(seg_dir: str, frame_name: str) -> str:
    """test_video_seg.py:117: ``{seg_dir}/{frame stem}.png``"""
    return os.path.join(seg_dir, f'{frame_name}.png')
