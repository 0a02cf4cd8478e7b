"""Property tests (hypothesis): invariants of the reference's bank arithmetic that hold for ANY
input, checked on the CPU for the oracle (which is pinned to the reference by the golden vectors) and on the GPU for
the product against the oracle on random small shapes - ragged sizes, single slots, duplicate keys, all-merge /
all-append thresholds, budgets that force one or several LFU thresholds."""
import math

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import afb_oracle as O

# derandomize: the examples are a fixed function of the test, so a run here and a run on the GPU box see the same cases
CPU = settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
GPU = settings(max_examples=16, deadline=None, derandomize=True, database=None,
               suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _bank(seed, n, d_k=16, d_v=24, scale=1.58):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(d_k, n, generator=g) * scale, torch.randn(d_v, n, generator=g), g


# ---------------------------------------------------------------------------------------------------
# oracle invariants (CPU)
# ---------------------------------------------------------------------------------------------------
@CPU
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 70), hw=st.integers(1, 40))
def test_read_is_a_convex_combination_and_counts_are_bounded(seed, n, hw):
    k, v, g = _bank(seed, n)
    q_in, q_out = torch.randn(1, 16, hw, generator=g) * 1.58, torch.randn(1, 24, hw, generator=g)
    info = [torch.zeros(n, 2)]
    rr = O.matcher_forward([k], [v], info, q_in, q_out, 1e-3, update_bank=True, keep_p=True)
    p = rr.p[0][0]
    assert torch.allclose(p.sum(dim=0), torch.ones(hw), atol=1e-5)                    # softmax over MEMORY (AFB_URR.py:145)
    mem = rr.out[0, 0, :24]
    assert (mem <= v.max(dim=1, keepdim=True).values + 1e-4).all() and (mem >= v.min(dim=1, keepdim=True).values - 1e-4).all()
    assert torch.equal(rr.out[0, 0, 24:], q_out[0])                                   # [mem ; q_out]  (:159)
    cnt = rr.cnt[0]
    assert (cnt >= 0).all() and (cnt <= hw).all() and cnt.sum() <= hw * 1000          # at most 1/thres slots per query
    assert torch.allclose(info[0][:, 1], torch.log(cnt + 1))                          # :174


@CPU
@given(seed=st.integers(0, 10 ** 6), n=st.integers(2, 60), hw=st.integers(1, 30),
       thres=st.sampled_from([-1.0, 0.3, 0.95, 2.0]), budget=st.integers(4, 200))
def test_update_invariants(seed, n, hw, thres, budget):
    """every candidate merges or appends (never both, NaN-free input); appended columns are raw copies in ascending
    candidate order; merged slots keep direction-blended keys of no larger norm; the budget holds whenever remove() ran;
    surviving rows keep their order; peak_n / replace_n are consistent"""
    k, v, g = _bank(seed, n)
    pk, pv = torch.randn(16, hw, generator=g) * 1.58, torch.randn(24, hw, generator=g)
    if hw > 2:
        pk[:, 0] = k[:, n // 2] * 1.7                     # an exact direction match: cosine 1
    fb = O.OracleFeatureBank(1, budget, 'cpu', thres_close=thres)
    fb.init_bank([k.clone()], [v.clone()])
    fb.info[0][:, 0] = torch.sort(torch.randint(0, 5, (n,), generator=g).float()).values
    fb.info[0][:, 1] = torch.rand(n, generator=g) * 30
    info0 = fb.info[0].clone()
    norms0 = k.norm(dim=0)
    try:
        fb.update([pk.clone()], [pv.clone()], 6)
    except (RuntimeError, ValueError, IndexError):
        return      # the reference raises when remove() empties the bank while still over budget (App. A item 10)
    d = fb.last_decisions[0]
    assert len(d.merge_q) + len(d.append_q) == hw and not set(d.merge_q.tolist()) & set(d.append_q.tolist())
    assert torch.equal(d.append_q, torch.sort(d.append_q).values)
    n_app = len(d.append_q)
    if n_app:
        assert torch.equal(fb.keys[0][:, -n_app:], pk[:, d.append_q]) and torch.equal(fb.values[0][:, -n_app:], pv[:, d.append_q])
        assert (fb.info[0][-n_app:, 0] == 6).all() and (fb.info[0][-n_app:, 1] == 0).all()
    if d.remove is not None:
        assert fb.keys[0].shape[1] <= fb.class_budget                      # FeatureBank.py:134
        kept = d.remove.keep_mask
        assert torch.equal(fb.info[0][:int(kept.sum()), 0], info0[kept, 0])                       # order preserved
        T = d.remove.thresholds
        assert all(b > a for a, b in zip(T, T[1:]))                        # thresholds strictly increase (:136)
        assert fb.replace_n[0] == n - int(kept.sum())
    else:
        assert fb.keys[0].shape[1] == n + n_app
        if len(d.touched):
            new_norm = fb.keys[0][:, d.touched].norm(dim=0)
            assert (new_norm <= norms0[d.touched] * (1 + 1e-5)).all()      # blend of unit vectors: norm never grows (:81-84)
    assert fb.peak_n[0] >= fb.keys[0].shape[1] and (fb.info[0][:, 1] <= 1e5).all()


@CPU
@given(vals=st.lists(st.floats(0, 50, allow_nan=False, width=32), min_size=2, max_size=60), request=st.integers(0, 20),
       budget=st.integers(1, 60))
def test_lfu_search_matches_a_direct_restatement(vals, request, budget):
    """sharded.lfu_threshold_search (one rank) == the reference's loop (FeatureBank.py:121-138) on any LFU vector"""
    from vfloodnet_b200 import sharded
    lfu = torch.tensor(vals, dtype=torch.float32)
    keep = torch.ones(len(vals), dtype=torch.bool)
    T = int(lfu.min()) + 1
    seq, ok = [T], True
    while True:
        keep = keep & (lfu > T)
        if (budget - int(keep.sum())) - request >= 0:
            break
        if not keep.any():
            ok = False
            break
        T = int(lfu[keep].min()) + 1
        seq.append(T)
    solo = sharded.ThreadComm.make(1)[0]
    if not ok:
        with pytest.raises(RuntimeError):
            sharded.lfu_threshold_search(lfu, float(budget), request, solo)
        return
    kl, kg, Tf, thr = sharded.lfu_threshold_search(lfu, float(budget), request, solo)
    assert thr == seq and Tf == seq[-1] and kl == kg == int(keep.sum())


# ---------------------------------------------------------------------------------------------------
# product vs oracle on random shapes (GPU)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 3000), hw=st.integers(1, 700), impl=st.sampled_from([1, 2]))
def test_gpu_read_matches_oracle_on_random_shapes(seed, n, hw, impl):
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import synth
    g = torch.Generator().manual_seed(seed)
    ns = [n, max(1, n // 3)]
    keys, vals = zip(*[synth.gen_bank(g, m) for m in ns])
    info = [synth.gen_info(g, m, 9) for m in ns]
    q_in, q_out = synth.gen_query(g, hw)
    info_o = [i.clone() for i in info]
    rr = O.matcher_forward(list(keys), list(vals), info_o, q_in, q_out, 1e-3, update_bank=True, keep_p=True)
    fb = vfn.FeatureBank(2, 10 ** 7, 'cuda', impl=impl)
    fb.load_state(list(keys), list(vals), info)
    out = vfn.Matcher(update_bank=True)(fb, q_in.cuda(), q_out.cuda())
    assert (out.cpu() - rr.out).abs().max().item() <= (1e-4 if impl == 1 else 1e-3)
    eps = 1e-5 if impl == 1 else 2e-4
    for c in range(2):
        cnt = torch.round(torch.exp((fb.info[c][:, 1].cpu() - info[c][:, 1]).double()) - 1).long()
        p = rr.p[c][0]
        assert torch.all(cnt >= (p > 1e-3 * (1 + eps)).sum(dim=1)) and torch.all(cnt <= (p > 1e-3 * (1 - eps)).sum(dim=1))


@pytest.mark.gpu
@GPU
@given(seed=st.integers(0, 10 ** 6), n=st.integers(2, 2500), hw=st.integers(1, 600), impl=st.sampled_from([1, 2]),
       frac=st.sampled_from([0.0, 0.4, 1.0]), slack=st.integers(-200, 400))
def test_gpu_update_matches_oracle_on_random_shapes(seed, n, hw, impl, frac, slack):
    """match / merge / append / LFU eviction against the oracle for random bank sizes, candidate counts, merge
    fractions and budgets around the bank size (slack < hw forces remove(); very negative slack may empty the bank)"""
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import synth
    g = torch.Generator().manual_seed(seed)
    keys, vals = zip(*[synth.gen_bank(g, n) for _ in range(2)])
    info = [synth.gen_info(g, n, 12) for _ in range(2)]
    pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw, frac) for c in range(2)])
    budget = int(2.5 * max(n + slack, 2))                    # class_budget = 0.8 * (budget // 2) ~ n + slack
    ofb = O.OracleFeatureBank(2, budget, 'cpu')
    ofb.init_bank([k.clone() for k in keys], [v.clone() for v in vals])
    ofb.info = [i.clone() for i in info]
    fb = vfn.FeatureBank(2, budget, 'cuda', impl=impl)
    fb.load_state(list(keys), list(vals), info)
    try:
        ofb.update([k.clone() for k in pk], [v.clone() for v in pv], 12)
    except (RuntimeError, ValueError, IndexError):
        with pytest.raises((RuntimeError, ValueError)):
            fb.update([k.cuda() for k in pk], [v.cuda() for v in pv], 12)
        return
    fb.update([k.cuda() for k in pk], [v.cuda() for v in pv], 12)
    for c in range(2):
        d, dg = ofb.last_decisions[c], fb.last_decisions[c]
        clear = d.margin > 4e-6
        assert torch.equal(dg['match_idx'].cpu().long()[clear], d.match_idx[clear])
        if not bool(clear.all()):
            continue                                          # a near tie may legitimately flip a merge target
        assert dg['evicted'] == (d.remove is not None)
        assert fb.bank_n(c) == ofb.keys[c].shape[1]
        assert torch.equal(fb.info[c][:, 0].cpu(), ofb.info[c][:, 0])
        np.testing.assert_allclose(fb.keys[c].cpu().numpy(), ofb.keys[c].numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(fb.values[c].cpu().numpy(), ofb.values[c].numpy(), rtol=1e-5, atol=1e-5)
        if d.remove is not None:
            assert fb.last_thresholds_obj[c] == d.remove.thresholds
    assert np.array_equal(fb.replace_n, ofb.replace_n)
