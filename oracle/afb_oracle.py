"""CPU oracle for the AFB-URR memory-propagation hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``vfloodnet_b200``) never routes through it.

It is a plain-PyTorch (CPU, fp32) restatement of the reference algorithm, written
from the reference's behaviour, with every decision (match index, merge set, append
set, evicted set, threshold sequence, usage counts) returned explicitly so the CUDA
path can be compared decision by decision.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this
path.  The oracle is pinned instead against OUTPUTS OF THE
REFERENCE ITSELF, imported in the build container from /root/reference by
``tests/golden/make_golden.py`` (committed), which writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against those vectors.

The one third-party arithmetic dependency on the path is
``torch_scatter.scatter_mean`` (rusty1s/pytorch_scatter, pinned ``torch-scatter==2.0.8``
in the reference README.md:58; not vendored, not installable here).  Its published
algorithm is restated in :func:`scatter_mean_2_0_8`; no reference test pins that
boundary, so that single function is "parity unpinned" upstream and is pinned here by
a hand-computed example in ``tests/test_oracle_golden.py``.

Reference files followed (all under /root/reference):
  video_module/model/FeatureBank.py:10-149   FeatureBank (init/append/update/remove)
  video_module/model/AFB_URR.py:130-178      Matcher.forward (read + usage count)
  video_module/model/AFB_URR.py:208-239      Decoder.forward URR block
  myutils/data.py:42-48                      calc_uncertainty
  myutils/data.py:134-151                    pad_divide_by
  test_video_seg.py:99-112                   per-frame loop order
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as NF


# --------------------------------------------------------------------------------------
# torch_scatter.scatter_mean, version 2.0.8 (FeatureBank.py:5,78,92 call sites)
# --------------------------------------------------------------------------------------
def scatter_mean_2_0_8(src: torch.Tensor, index: torch.Tensor, dim: int, out: torch.Tensor) -> torch.Tensor:
    """``scatter_mean(src, index, dim, out=out)`` as published in torch-scatter 2.0.8.

    out.scatter_add_(dim, index, src); count = scatter_add(ones_like(index)) along the same
    dim with dim_size = out.size(dim); count clamped to >= 1; out.true_divide_(count).
    On CPU ``scatter_add_`` accumulates sequentially in index order, i.e. ascending source
    position, which is the summation order the CUDA merge kernel reproduces.
    """
    out.scatter_add_(dim, index, src)
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = torch.zeros(out.size(), dtype=src.dtype, device=src.device)
    count.scatter_add_(dim, index, ones)
    count.clamp_(min=1)
    out.true_divide_(count)
    return out


# --------------------------------------------------------------------------------------
# Memory read  (AFB_URR.py:136-178)
# --------------------------------------------------------------------------------------
@dataclass
class ReadResult:
    out: torch.Tensor                      # (bs, obj_n, d_v + d_v, HW)
    cnt: List[torch.Tensor]                # per object (N,) usage counts  (AFB_URR.py:165)
    lse: List[torch.Tensor]                # per object (HW,) natural-log log-sum-exp of scaled logits
    p: Optional[List[torch.Tensor]] = None  # per object (bs, N, HW) softmax, only if keep_p


def matcher_forward(keys, values, info, q_in, q_out, thres_valid=1e-3, update_bank=False,
                    keep_p=False) -> ReadResult:
    """Matcher.forward restated (AFB_URR.py:136-178).

    keys[i]: (d_k, N_i), values[i]: (d_v, N_i), info[i]: (N_i, 2) mutated in place when
    update_bank (AFB_URR.py:174); q_in: (bs, d_k, HW); q_out: (bs, d_v, HW).
    """
    obj_n = len(keys)
    outs, cnts, lses, ps = [], [], [], []
    for i in range(obj_n):
        d_key, bank_n = keys[i].size()
        s = torch.matmul(keys[i].transpose(0, 1), q_in) / math.sqrt(d_key)   # :144
        p = NF.softmax(s, dim=1)                                             # :145 over memory
        mem = torch.matmul(values[i], p)                                     # :146
        outs.append(torch.cat([mem, q_out], dim=1))                          # :159
        lses.append(torch.logsumexp(s[0], dim=0))
        ones = torch.ones_like(p)
        zeros = torch.zeros_like(p)
        bank_cnt = torch.where(p > thres_valid, ones, zeros).sum(dim=2)[0]   # :165 batch elem 0
        cnts.append(bank_cnt)
        if update_bank:
            info[i][:, 1] += torch.log(bank_cnt + 1)                         # :174
        if keep_p:
            ps.append(p)
    out = torch.stack(outs, dim=0).transpose(0, 1)                           # :176
    return ReadResult(out=out, cnt=cnts, lse=lses, p=ps if keep_p else None)


# --------------------------------------------------------------------------------------
# Feature bank (FeatureBank.py)
# --------------------------------------------------------------------------------------
@dataclass
class RemoveDecision:
    thresholds: List[int]          # T sequence tried (FeatureBank.py:123,136)
    keep_mask: torch.Tensor        # (N_before,) bool: survivors of the final threshold
    balance: float


@dataclass
class UpdateDecision:
    match_idx: torch.Tensor        # (HW,) int64  j*        (FeatureBank.py:67)
    match_corr: torch.Tensor       # (HW,) fp32   c*        (FeatureBank.py:68)
    margin: torch.Tensor           # (HW,) fp32   top1 - top2 cosine (diagnostic, not in reference)
    merge_q: torch.Tensor          # ascending query positions merged      (:71)
    merge_slot: torch.Tensor       # bank slot of each merged query        (:72)
    touched: torch.Tensor          # sorted unique touched slots           (:73)
    append_q: torch.Tensor         # ascending query positions appended    (:100)
    remove: Optional[RemoveDecision] = None
    n_before: int = 0
    n_after: int = 0


class OracleFeatureBank:
    """Same constructor, attributes and methods as the reference FeatureBank (FeatureBank.py:8-149)."""

    def __init__(self, obj_n, memory_budget, device='cpu', update_rate=0.1, thres_close=0.95):
        self.obj_n = obj_n
        self.update_rate = update_rate
        self.thres_close = thres_close
        self.device = device
        self.info = [None for _ in range(obj_n)]
        self.peak_n = np.zeros(obj_n)
        self.replace_n = np.zeros(obj_n)
        self.class_budget = memory_budget // obj_n                # :20
        if obj_n == 2:
            self.class_budget = 0.8 * self.class_budget           # :21-22 (a Python float)
        self.keys = None
        self.values = None
        self.last_decisions: List[UpdateDecision] = []

    def init_bank(self, keys, values, frame_idx=0):               # :27-36
        self.keys = keys
        self.values = values
        for c in range(self.obj_n):
            _, bank_n = keys[c].shape
            self.info[c] = torch.zeros((bank_n, 2), device=self.device)
            self.info[c][:, 0] = frame_idx
            self.peak_n[c] = max(self.peak_n[c], self.info[c].shape[0])

    def append(self, keys, values, frame_idx=0):                  # :38-51
        if self.keys:
            for c in range(self.obj_n):
                self.keys[c] = torch.cat([self.keys[c], keys[c]], dim=1)
                self.values[c] = torch.cat([self.values[c], values[c]], dim=1)
                _, bank_n = keys[c].shape
                new_info = torch.ones((bank_n, 2), device=self.device) * 20   # :46
                new_info[:, 0] = frame_idx
                self.info[c] = torch.cat([self.info[c], new_info], dim=0)
                self.peak_n[c] = max(self.peak_n[c], self.info[c].shape[0])
        else:
            self.init_bank(keys, values, frame_idx)

    def update(self, prev_key, prev_value, frame_idx, update_rate=-1):   # :53-115
        if update_rate == -1:
            update_rate = self.update_rate
        self.last_decisions = []
        for c in range(self.obj_n):
            d_key, bank_n = self.keys[c].shape
            d_val, _ = self.values[c].shape

            normed_keys = NF.normalize(self.keys[c], dim=0)                       # :63
            normed_prev_key = NF.normalize(prev_key[c], dim=0)                    # :64
            mag_keys = self.keys[c].norm(p=2, dim=0)                              # :65
            corr = torch.mm(normed_keys.transpose(0, 1), normed_prev_key)         # :66
            related_bank_idx = corr.argmax(dim=0, keepdim=True)                   # :67
            related_bank_corr = torch.gather(corr, 0, related_bank_idx)           # :68
            if bank_n >= 2:
                top2 = corr.topk(2, dim=0).values
                margin = top2[0] - top2[1]
            else:
                margin = torch.full_like(related_bank_corr[0], float('inf'))

            sel = (related_bank_corr[0] > self.thres_close).nonzero(as_tuple=False)      # :71
            slot_of_sel = related_bank_idx[0, sel[:, 0]]                                 # :72
            touched, _cnt = slot_of_sel.unique(dim=0, return_counts=True)                # :73

            key_bank_update = torch.zeros((d_key, bank_n), dtype=torch.float, device=self.device)   # :76
            key_bank_idx = slot_of_sel.unsqueeze(0).expand(d_key, -1)                    # :77
            scatter_mean_2_0_8(normed_prev_key[:, sel[:, 0]], key_bank_idx, 1, key_bank_update)   # :78
            self.keys[c][:, touched] = mag_keys[touched] * (                             # :81-84
                (1 - update_rate) * normed_keys[:, touched] + update_rate * key_bank_update[:, touched])

            normed_values = NF.normalize(self.values[c], dim=0)                          # :87
            normed_prev_value = NF.normalize(prev_value[c], dim=0)                       # :88
            mag_values = self.values[c].norm(p=2, dim=0)                                 # :89
            val_bank_update = torch.zeros((d_val, bank_n), dtype=torch.float, device=self.device)   # :90
            val_bank_idx = slot_of_sel.unsqueeze(0).expand(d_val, -1)                    # :91
            scatter_mean_2_0_8(normed_prev_value[:, sel[:, 0]], val_bank_idx, 1, val_bank_update)  # :92
            self.values[c][:, touched] = mag_values[touched] * (                         # :94-97
                (1 - update_rate) * normed_values[:, touched] + update_rate * val_bank_update[:, touched])

            app = (related_bank_corr[0] <= self.thres_close).nonzero(as_tuple=False)     # :100

            dec = UpdateDecision(match_idx=related_bank_idx[0].clone(), match_corr=related_bank_corr[0].clone(),
                                 margin=margin, merge_q=sel[:, 0].clone(), merge_slot=slot_of_sel.clone(),
                                 touched=touched.clone(), append_q=app[:, 0].clone(), n_before=bank_n)

            if self.class_budget < bank_n + app.shape[0]:                                # :102
                dec.remove = self._remove(c, app.shape[0], frame_idx)                    # :103

            self.keys[c] = torch.cat([self.keys[c], prev_key[c][:, app[:, 0]]], dim=1)          # :105
            self.values[c] = torch.cat([self.values[c], prev_value[c][:, app[:, 0]]], dim=1)    # :106-107
            new_info = torch.zeros((app.shape[0], 2), device=self.device)                       # :109
            new_info[:, 0] = frame_idx
            self.info[c] = torch.cat([self.info[c], new_info], dim=0)                           # :111
            self.peak_n[c] = max(self.peak_n[c], self.info[c].shape[0])                         # :113
            self.info[c][:, 1] = torch.clamp(self.info[c][:, 1], 0, 1e5)                        # :115
            dec.n_after = self.info[c].shape[0]
            self.last_decisions.append(dec)

    def _remove(self, class_idx, request_n, frame_idx) -> RemoveDecision:    # :117-143
        old_size = self.keys[class_idx].shape[1]
        LFU = frame_idx - self.info[class_idx][:, 0]                          # :121
        LFU = self.info[class_idx][:, 1] / LFU                                # :122
        thres_dynamic = int(LFU.min()) + 1                                    # :123
        thresholds = [thres_dynamic]
        keep_total = torch.ones(old_size, dtype=torch.bool, device=LFU.device)
        alive = torch.arange(old_size, device=LFU.device)
        while True:
            selected = LFU > thres_dynamic                                    # :127
            self.keys[class_idx] = self.keys[class_idx][:, selected]
            self.values[class_idx] = self.values[class_idx][:, selected]
            self.info[class_idx] = self.info[class_idx][selected]
            LFU = LFU[selected]
            alive = alive[selected]
            balance = (self.class_budget - self.keys[class_idx].shape[1]) - request_n   # :134
            if balance < 0:
                thres_dynamic = int(LFU.min()) + 1                            # :136 raises on empty, like the reference
                thresholds.append(thres_dynamic)
            else:
                break
        keep_total[:] = False
        keep_total[alive] = True
        new_size = self.keys[class_idx].shape[1]
        self.replace_n[class_idx] += old_size - new_size                      # :140-141
        return RemoveDecision(thresholds=thresholds, keep_mask=keep_total, balance=balance)

    def remove(self, class_idx, request_n, frame_idx):
        return self._remove(class_idx, request_n, frame_idx).balance

    def print_peak_mem(self):                                                 # :145-149
        ur = self.peak_n / self.class_budget
        rr = self.replace_n / self.class_budget
        print(f'Obj num: {self.obj_n}.', f'Budget / obj: {self.class_budget}.', f'UR: {ur}.', f'Replace: {rr}.')


# --------------------------------------------------------------------------------------
# URR block of Decoder.forward (AFB_URR.py:214-237) without the three small convolutions
# --------------------------------------------------------------------------------------
def calc_uncertainty(score: torch.Tensor) -> torch.Tensor:                    # myutils/data.py:42-48
    score_top, _ = score.topk(k=2, dim=1)
    uncertainty = score_top[:, 0] / (score_top[:, 1] + 1e-8)
    return torch.exp(1 - uncertainty).unsqueeze(1)


def urr_pre(p: torch.Tensor, r1: torch.Tensor, feature_shape, local_size: int = 7):
    """AFB_URR.py:214-231.  p: (bs*obj_n, 2, h/2, w/2) coarse logits from pred2; r1: (bs*obj_n, 64, h, w).

    Returns (p_up, uncertainty, r1_conf, local_match) exactly as the reference computes them:
    p_up (bs*obj_n,2,h,w), uncertainty (bs*obj_n,1,h,w), r1_conf (bs*obj_n,1,h,w),
    local_match (bs*obj_n,128,h,w) = cat([r1, r1_local]).
    """
    pad = local_size // 2
    p = NF.interpolate(p, scale_factor=2, mode='bilinear', align_corners=False)       # :214
    bs, obj_n, h, w = feature_shape
    rough_seg = NF.softmax(p, dim=1)[:, 1]                                            # :217
    rough_seg = rough_seg.view(bs, obj_n, h, w)
    rough_seg = NF.softmax(rough_seg, dim=1)                                          # :219
    uncertainty = calc_uncertainty(rough_seg)                                         # :222
    uncertainty = uncertainty.expand(-1, obj_n, -1, -1).reshape(bs * obj_n, 1, h, w)  # :223
    rough_seg = rough_seg.view(bs * obj_n, 1, h, w)                                   # :225
    r1_weighted = r1 * rough_seg                                                      # :226
    r1_local = NF.avg_pool2d(r1_weighted, local_size, stride=1, padding=pad)          # :227
    r1_local = r1_local / (NF.avg_pool2d(rough_seg, local_size, stride=1, padding=pad) + 1e-8)   # :228
    r1_conf = NF.max_pool2d(rough_seg, local_size, stride=1, padding=pad)             # :229
    local_match = torch.cat([r1, r1_local], dim=1)                                    # :231
    return p, uncertainty, r1_conf, local_match


def urr_post(p_up: torch.Tensor, uncertainty: torch.Tensor, r1_conf: torch.Tensor, q_local: torch.Tensor):
    """AFB_URR.py:233-237.  q_local = local_pred2(relu(local_ResMM(local_convFM(local_match)))) (cuDNN, out of scope).

    Returns the final per-object foreground probability (bs*obj_n, 2h, 2w).
    """
    q = r1_conf * q_local                                                             # :233
    p = p_up + uncertainty * q                                                        # :235
    p = NF.interpolate(p, scale_factor=2, mode='bilinear', align_corners=False)       # :236
    return NF.softmax(p, dim=1)[:, 1]                                                 # :237


def keyvalue_forward(x, wk, bk, wv, bv):
    """KeyValue.forward (AFB_URR.py:103-111): two 3x3 / padding 1 convolutions of the same feature map, flattened to
    (B, d, H*W).  Evaluated in the dtype of x (tests pass float64 for the exact value, float32 for the reference's)."""
    key = NF.conv2d(x, wk.to(x.dtype), None if bk is None else bk.to(x.dtype), padding=1)     # :104 (Key, :100)
    key = key.view(*key.shape[:2], -1)                                                        # :105
    val = NF.conv2d(x, wv.to(x.dtype), None if bv is None else bv.to(x.dtype), padding=1)     # :107 (Value, :101)
    val = val.view(*val.shape[:2], -1)                                                        # :108
    return key, val


def pad_divide_by(in_list, d, in_size):                                               # myutils/data.py:134-151
    h, w = in_size
    new_h = h + d - h % d if h % d > 0 else h
    new_w = w + d - w % d if w % d > 0 else w
    lh, uh = int((new_h - h) / 2), int(new_h - h) - int((new_h - h) / 2)
    lw, uw = int((new_w - w) / 2), int(new_w - w) - int((new_w - w) / 2)
    pad_array = (int(lw), int(uw), int(lh), int(uh))
    return [NF.pad(x, pad_array) for x in in_list], pad_array


# --------------------------------------------------------------------------------------
# One hot-path step at the drop-in boundary, in the order of test_video_seg.py:108-112
# --------------------------------------------------------------------------------------
def hot_path_step(fb: OracleFeatureBank, q_in, q_out, prev_key, prev_value, frame_idx,
                  urr_in=None, thres_valid=1e-3):
    """read (segment's Matcher call) -> [URR] -> update.  Returns (ReadResult, urr_prob or None)."""
    rr = matcher_forward(fb.keys, fb.values, fb.info, q_in, q_out, thres_valid, update_bank=True)
    prob = None
    if urr_in is not None:
        p, r1, q_local, feature_shape = urr_in
        p_up, unc, conf, _local_match = urr_pre(p, r1, feature_shape)
        prob = urr_post(p_up, unc, conf, q_local)
    fb.update(prev_key, prev_value, frame_idx)
    return rr, prob
