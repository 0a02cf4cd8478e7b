"""Seeded synthetic inputs at the drop-in boundary (regime C): bank tensors, query features and
update candidates with a controlled merge/append mix.  Pure torch; used by tests/ and bench.py on both arms."""
from __future__ import annotations

import math

import torch

S_K = 1.58   # key/query element scale so that logits/sqrt(128) have sigma ~2.5


def _dev(g: torch.Generator):
    """tensors are produced on the generator's device (a CUDA generator builds a clip in HBM without the host)"""
    return g.device


def gen_bank(g: torch.Generator, n: int, d_key=128, d_val=512, s_k=S_K):
    d = _dev(g)
    return torch.randn(d_key, n, generator=g, device=d) * s_k, torch.randn(d_val, n, generator=g, device=d)


def gen_query(g: torch.Generator, hw: int, d_key=128, d_val=512, s_k=S_K):
    d = _dev(g)
    return torch.randn(1, d_key, hw, generator=g, device=d) * s_k, torch.randn(1, d_val, hw, generator=g, device=d)


def gen_candidates(g: torch.Generator, key: torch.Tensor, value: torch.Tensor, hw: int, frac_merge=0.5, dup=True,
                   s_k=S_K, noise=0.1):
    """`frac_merge` of the candidates are noisy copies of random bank columns (cos ~0.995 -> merged, every 4th one
    duplicated so several candidates hit the same slot); the rest are fresh draws (cos ~0 -> appended)."""
    d_k, n = key.shape
    d_v = value.shape[0]
    d = _dev(g)
    n_m = int(hw * frac_merge)
    src = torch.randint(0, n, (n_m,), generator=g, device=d)
    if dup and n_m >= 4:
        src[1::4] = src[0::4][: len(src[1::4])]
    k_m = key[:, src] + noise * s_k * torch.randn(d_k, n_m, generator=g, device=d)
    v_m = value[:, src] + noise * torch.randn(d_v, n_m, generator=g, device=d)
    k_f = torch.randn(d_k, hw - n_m, generator=g, device=d) * s_k
    v_f = torch.randn(d_v, hw - n_m, generator=g, device=d)
    perm = torch.randperm(hw, generator=g, device=d)
    return torch.cat([k_m, k_f], 1)[:, perm].contiguous(), torch.cat([v_m, v_f], 1)[:, perm].contiguous()


def gen_info(g: torch.Generator, n: int, frame_idx: int):
    d = _dev(g)
    info = torch.zeros(n, 2, device=d)
    info[:, 0] = torch.randint(0, max(frame_idx, 1), (n,), generator=g, device=d).float()
    info[:, 1] = torch.rand(n, generator=g, device=d) * 50
    return info


def gen_urr_inputs(g: torch.Generator, obj_n: int, h: int, w: int, c=64):
    """p: coarse logits (obj_n,2,h/2,w/2); r1 (1,c,h,w) shared by objects; q_local: stand-in for the output of the
    three local convolutions (obj_n,2,h,w)."""
    d = _dev(g)
    p = torch.randn(obj_n, 2, h // 2, w // 2, generator=g, device=d) * 2
    r1 = torch.randn(1, c, h, w, generator=g, device=d).relu()
    q_local = torch.randn(obj_n, 2, h, w, generator=g, device=d)
    return p, r1, q_local


class ClipGenerator:
    """A synthetic clip at the boundary: per frame (q_in, q_out) for the read and (prev_key, prev_value) lists for the
    update, produced from a seeded generator and a small pool of 'scene prototypes' so that a fraction of the
    candidates re-occur (merge) and the rest are new (append) - the 480p 2-object configuration of BASELINE.json."""

    def __init__(self, seed=0, obj_n=2, hw=1620, d_key=128, d_val=512, frac_merge=0.1, n_init=None, device='cpu'):
        self.g = torch.Generator(device=device).manual_seed(seed)
        self.obj_n, self.hw, self.d_key, self.d_val, self.frac_merge = obj_n, hw, d_key, d_val, frac_merge
        self.n_init = n_init or hw

    def init(self):
        keys, vals = zip(*[gen_bank(self.g, self.n_init, self.d_key, self.d_val) for _ in range(self.obj_n)])
        self._last = [(k, v) for k, v in zip(keys, vals)]
        return [k.clone() for k in keys], [v.clone() for v in vals]

    def frame(self):
        q_in, q_out = gen_query(self.g, self.hw, self.d_key, self.d_val)
        pk, pv = [], []
        for c in range(self.obj_n):
            k, v = gen_candidates(self.g, self._last[c][0], self._last[c][1], self.hw, self.frac_merge)
            pk.append(k); pv.append(v)
            self._last[c] = (k, v)     # next frame's merge candidates resemble this frame's features
        return q_in, q_out, pk, pv
