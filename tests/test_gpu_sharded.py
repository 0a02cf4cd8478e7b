"""GPU parity of the SHARDED feature bank (BASELINE config 5) on one device: `world` ranks run as threads
over a ThreadComm, each with its own shard; after every update the gathered bank must equal the single (unsharded)
bank bit for bit - match decisions, merge arithmetic, LFU thresholds, kept set, append order - and the split-memory
read must reproduce the single-bank read.  The NCCL form of the same code is tests/multi_gpu_check.py (torchrun)."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vfn():
    import vfloodnet_b200 as v
    assert torch.cuda.is_available()
    return v


def run_ranks(world, fn):
    """fn(comm) on `world` threads; returns the per-rank results, re-raises the first exception"""
    from vfloodnet_b200.sharded import ThreadComm
    comms = ThreadComm.make(world)
    res, err = [None] * world, [None] * world

    def body(r):
        try:
            torch.cuda.set_device(0)
            res[r] = fn(comms[r])
        except BaseException as e:      # noqa: BLE001 - reported below
            err[r] = e
            comms[r]._s.barrier.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    first = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)] or \
            [e for e in err if e is not None]
    if first:
        raise first[0]
    return res


@pytest.mark.parametrize('world,impl', [(2, 2), (3, 2), (2, 1)])
def test_sharded_update_equals_single_bank(vfn, world, impl):
    from vfloodnet_b200 import synth
    from vfloodnet_b200.sharded import ShardedFeatureBank
    g = torch.Generator().manual_seed(5 + world)
    n0, hw, frames, budget = 2500, 800, 7, 9000          # class_budget 3600 -> LFU eviction from frame 3 on
    keys, vals = zip(*[synth.gen_bank(g, n0) for _ in range(2)])
    clip = []
    for t in range(frames):
        pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw, 0.4) for c in range(2)])
        usage = [torch.rand(20000, generator=g) * (12.0 * (t + 1)) for _ in range(2)]   # info[:,1] teacher forcing
        clip.append((list(pk), list(pv), usage))
    q_in, q_out = synth.gen_query(g, hw)

    # single bank
    full = vfn.FeatureBank(2, budget, 'cuda', impl=impl)
    full.init_bank([k.clone() for k in keys], [v.clone() for v in vals])
    trace = []
    for t, (pk, pv, usage) in enumerate(clip):
        for c in range(2):
            n = full.bank_n(c)
            full.info[c][:, 1] = usage[c][:n].cuda()
        full.update([k.cuda() for k in pk], [v.cuda() for v in pv], t + 1)
        trace.append(dict(keys=[full.keys[c].clone() for c in range(2)], values=[full.values[c].clone() for c in range(2)],
                          info=[full.info[c].clone() for c in range(2)], n=[full.bank_n(c) for c in range(2)],
                          thr=[list(full.last_thresholds_obj[c] or []) if full.last_decisions[c]['evicted'] else []
                               for c in range(2)],
                          evicted=[full.last_decisions[c]['evicted'] for c in range(2)],
                          corr=[full.last_decisions[c]['match_corr'].clone() for c in range(2)],
                          replace_n=full.replace_n.copy(), peak_n=full.peak_n.copy()))
    m = vfn.Matcher(update_bank=True)
    out_full = m(full, q_in.cuda(), q_out.cuda())
    info_full = [full.info[c].clone() for c in range(2)]
    assert any(any(tr['evicted']) for tr in trace), 'the clip must exercise LFU eviction'
    assert any(len(th) > 1 for tr in trace for th in tr['thr']), 'the clip must exercise a multi-step threshold search'

    def rank_body(comm):
        sfb = ShardedFeatureBank(2, budget, 'cuda', comm=comm, impl=impl)
        sfb.init_bank([k.clone() for k in keys], [v.clone() for v in vals])
        for t, (pk, pv, usage) in enumerate(clip):
            for c in range(2):
                k_g, v_g, i_g, _ = sfb.gather_state(c)
                i_g = i_g.clone()
                i_g[:, 1] = usage[c][:i_g.shape[0]].cuda()
                sfb.scatter_info(c, i_g)
            sfb.update([k.cuda() for k in pk], [v.cuda() for v in pv], t + 1)
            tr = trace[t]
            for c in range(2):
                k_g, v_g, i_g, seq = sfb.gather_state(c)
                assert sfb.n_global[c] == tr['n'][c] == k_g.shape[1], (t, c, sfb.n_global[c], tr['n'][c])
                assert bool((seq[1:] > seq[:-1]).all())
                assert sfb.last_decisions[c]['evicted'] == tr['evicted'][c]
                if tr['evicted'][c]:
                    assert sfb.last_thresholds_obj[c] == tr['thr'][c], (t, c, sfb.last_thresholds_obj[c], tr['thr'][c])
                assert torch.equal(sfb.last_decisions[c]['match_corr'], tr['corr'][c]), (t, c)
                assert torch.equal(k_g, tr['keys'][c]), (t, c, (k_g - tr['keys'][c]).abs().max().item())
                assert torch.equal(v_g, tr['values'][c]), (t, c)
                assert torch.equal(i_g, tr['info'][c]), (t, c)
            assert (sfb.replace_n == tr['replace_n']).all() and (sfb.peak_n == tr['peak_n']).all()
        out = sfb.read(q_in.cuda(), q_out.cuda())
        info = [sfb.gather_state(c)[2] for c in range(2)]
        return out, info, [sfb.n_local(c) for c in range(2)]

    res = run_ranks(world, rank_body)
    tol = 2e-5 if impl == 1 else 2e-4
    for out, info, n_loc in res:
        assert (out - out_full).abs().max().item() <= tol
        for c in range(2):
            d = (info[c][:, 1] - info_full[c][:, 1]).abs()
            assert int((d > 1e-6).sum()) <= 2          # usage counts are local and exact up to threshold-band flips
    # the append chunks keep the shards balanced
    sizes = [r[2] for r in res]
    for c in range(2):
        assert sum(s[c] for s in sizes) == trace[-1]['n'][c]
        assert max(s[c] for s in sizes) - min(s[c] for s in sizes) < 0.35 * trace[-1]['n'][c]


def test_sharded_match_tie_goes_to_earliest_slot(vfn):
    """exact duplicate slots living on different shards: the earliest-inserted one (lowest reference index) wins"""
    from vfloodnet_b200 import synth
    from vfloodnet_b200.sharded import ShardedFeatureBank
    g = torch.Generator().manual_seed(3)
    n0, hw = 600, 256
    k, v = synth.gen_bank(g, n0)
    k[:, 450] = k[:, 7]                  # slot 450 (rank 1) duplicates slot 7 (rank 0)
    k[:, 599] = k[:, 301]                # both on rank 1
    cand_k = k[:, torch.arange(hw) % n0].clone()
    cand_k[:, 0], cand_k[:, 1] = k[:, 450], k[:, 599]
    cand_v = torch.randn(512, hw, generator=g)

    def rank_body(comm):
        sfb = ShardedFeatureBank(1, 10 ** 6, 'cuda', comm=comm)
        sfb.init_bank([k.clone()], [v.clone()])
        sfb.update([cand_k.cuda()], [cand_v.cuda()], 1)
        return sfb.last_decisions[0]['match_seq'].cpu()

    for seq in run_ranks(2, rank_body):
        assert seq[0].item() == 7 and seq[1].item() == 301
        assert seq[2:].tolist() == [(i % n0) if (i % n0) not in (450, 599) else {450: 7, 599: 301}[i % n0]
                                    for i in range(2, hw)]
