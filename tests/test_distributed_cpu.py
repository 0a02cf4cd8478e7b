"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the sharded-bank exchange steps (LSE combine, partial
readout reduction, arg-max combine with global tie-break) and the stream-parallel partition used by bench.py."""
import math
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from vfloodnet_b200 import sharded
        g = torch.Generator().manual_seed(7)          # same data on every rank
        n, hw, dk, dv = 301, 37, 16, 24
        K = torch.randn(dk, n, generator=g) * 1.58
        V = torch.randn(dv, n, generator=g)
        Q = torch.randn(dk, hw, generator=g) * 1.58
        s_full = (K.t() @ Q) / math.sqrt(dk)
        p_full = torch.softmax(s_full, dim=0)
        mem_full = V @ p_full
        lo, hi = sharded.shard_range(n, rank, world)
        s = s_full[lo:hi]
        m = s.max(dim=0).values
        l = torch.exp(s - m).sum(dim=0)
        lse = sharded.combine_lse(torch.stack([m, l], dim=-1))
        ok_lse = torch.allclose(lse, torch.logsumexp(s_full, dim=0), atol=1e-5)
        p_loc = torch.exp(s - lse)                      # normalised with the GLOBAL lse
        cnt_loc = (p_loc > 1e-3).sum(dim=1)
        ok_cnt = torch.equal(cnt_loc, (p_full[lo:hi] > 1e-3).sum(dim=1))
        mem = sharded.reduce_readout(V[:, lo:hi] @ p_loc)
        ok_mem = torch.allclose(mem, mem_full, atol=1e-5)
        # arg-max combine with exact duplicates across shards: lowest global slot must win
        Kn = torch.nn.functional.normalize(K, dim=0)
        Kn[:, n - 5] = Kn[:, 3]                        # duplicate of slot 3 lives in the last shard
        cand = Kn[:, [3, 100, 200]]
        corr = Kn.t() @ cand
        c_loc, i_loc = corr[lo:hi].max(dim=0)
        best, idx = sharded.combine_match(c_loc, i_loc + lo)
        ok_match = idx.tolist() == corr.argmax(dim=0).tolist() and idx[0].item() == 3
        # empty shard (-inf, 0) must not poison the combine
        ml_e = torch.stack([m, l], dim=-1) if rank == 0 else torch.stack([torch.full_like(m, -math.inf), torch.zeros_like(l)], -1)
        lse_e = sharded.combine_lse(ml_e)
        ok_empty = torch.allclose(lse_e, torch.logsumexp(s_full[:sharded.shard_range(n, 0, world)[1]], dim=0), atol=1e-5)
        q.put((rank, bool(ok_lse), bool(ok_cnt), bool(ok_mem), bool(ok_match), bool(ok_empty)))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), f'rank {r[0]}: lse/cnt/mem/match/empty = {r[1:]}'


def test_shard_ranges_cover_and_preserve_order():
    from vfloodnet_b200.sharded import shard_range
    for n in (0, 1, 7, 100000):
        for w in (1, 2, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_bench_reference_arm_rank_gating(monkeypatch, capsys):
    """under torchrun only rank 0 runs and prints the reference arm; other ranks exit without work"""
    import bench
    import argparse
    args = argparse.Namespace(gpus=2, steps=1, warmup=0, frames=100, frac_merge=0.1)
    called = []
    stub = dict(kind='port', secs=[2.0], cnt=[4], sizes=[[10, 10]], desc='stub', frames_done=4, arm=None)
    monkeypatch.setattr(bench, 'reference_free_run', lambda *a, **k: (called.append(1) or stub))
    bench.main_reference(args, rank=1, world=2)
    assert not called and capsys.readouterr().out == ''
    bench.main_reference(args, rank=0, world=2)
    out = capsys.readouterr().out
    assert called and '"impl": "reference"' in out
    import json
    line = json.loads(out)
    assert line['value'] == 2.0 and line['e2e']['value'] == 2.0 and line['cpu_baseline']['kind'] == 'port'


def test_bench_reference_arm_free_runs_the_clip():
    """the reference arm consumes the SAME clip as our arm (ClipGenerator(seed=100)), free-running: after t frames the
    bank holds init + the appended candidates of those frames; every frame lands in exactly one step"""
    import bench
    from vfloodnet_b200 import synth
    bench.select_workload('480p-2obj-100frame-clip-hotpath')
    old = (bench.HW_H, bench.HW_W, bench.R1_H, bench.R1_W)
    try:
        bench.HW_H, bench.HW_W, bench.R1_H, bench.R1_W = 4, 6, 16, 24          # a tiny grid: seconds on the CPU
        r = bench.reference_free_run(0.25, 7, seed=100, steps=3)
        assert r['cnt'] == [3, 2, 2] and r['frames_done'] == 7 and all(s > 0 for s in r['secs'])
        gen = synth.ClipGenerator(seed=100, obj_n=2, hw=24, frac_merge=0.25)
        gen.init()
        assert r['sizes'][0][0] > 24 and r['sizes'][-1][0] == r['arm'].sizes()[0]
        assert r['kind'] in ('reference', 'port')
    finally:
        bench.HW_H, bench.HW_W, bench.R1_H, bench.R1_W = old


# ---------------------------------------------------------------------------------------------------
# sharded LFU eviction: the threshold search over two shards must take the oracle's (= reference's) T sequence and
# keep exactly the oracle's survivors; the thread communicator must agree with the gloo one
# ---------------------------------------------------------------------------------------------------
def _lfu_case(seed):
    from oracle import afb_oracle as O
    g = torch.Generator().manual_seed(seed)
    n, frame, budget_obj, request = 4001, 40, 2000, 700
    info = torch.zeros(n, 2)
    info[:, 0] = torch.randint(0, frame, (n,), generator=g).float()
    info[:, 1] = torch.rand(n, generator=g) * 400
    info[::7, 1] = 0.0                                   # fresh entries: LFU 0
    info[5, 1] = 5.0 * (frame - info[5, 0])              # LFU exactly 5.0 (strict > must evict it at T=5)
    ofb = O.OracleFeatureBank(1, budget_obj, 'cpu')
    ofb.init_bank([torch.zeros(8, n)], [torch.zeros(8, n)])
    ofb.info[0] = info.clone()
    dec = ofb._remove(0, request, frame)
    lfu = info[:, 1] / (frame - info[:, 0])
    return lfu, float(ofb.class_budget), request, dec


def _lfu_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from vfloodnet_b200 import sharded
        ok = True
        for seed in (0, 1, 2):
            lfu, budget, request, dec = _lfu_case(seed)
            # interleaved ownership (what the chunked appends produce), local order preserved
            own = torch.arange(lfu.numel()) % world == rank
            kl, kg, T, thr = sharded.lfu_threshold_search(lfu[own], budget, request, sharded.DistComm())
            keep = lfu[own] > float(T)
            ok = ok and thr == dec.thresholds and kg == int(dec.keep_mask.sum()) and kl == int(keep.sum())
            ok = ok and torch.equal(keep, dec.keep_mask[own]) and len(thr) > 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_lfu_threshold_search_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lfu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_thread_comm_matches_reference_search_and_raises_like_it():
    import threading
    import pytest
    from vfloodnet_b200 import sharded
    lfu, budget, request, dec = _lfu_case(3)
    world = 3
    comms = sharded.ThreadComm.make(world)
    out = [None] * world

    def body(r):
        own = torch.arange(lfu.numel()) % world == r
        out[r] = sharded.lfu_threshold_search(lfu[own], budget, request, comms[r])

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=60) for t in th]
    assert all(o is not None and o[3] == dec.thresholds and o[1] == int(dec.keep_mask.sum()) for o in out)
    assert sum(o[0] for o in out) == out[0][1]
    # error behaviour on one rank == the reference's exceptions (FeatureBank.py:123,136)
    solo = sharded.ThreadComm.make(1)[0]
    with pytest.raises(ValueError):
        sharded.lfu_threshold_search(torch.tensor([1.0, float('nan')]), 10.0, 1, solo)
    with pytest.raises(RuntimeError):
        sharded.lfu_threshold_search(torch.tensor([1.0, 2.0]), 1.0, 5, solo)


def test_bench_workload_selection(monkeypatch):
    """--workload switches the query grid, the r1 grid, the pre-filled bank and the frame numbering together"""
    import importlib
    bench = importlib.import_module('bench')
    try:
        frames = bench.select_workload('1080p-2obj-bank-at-capacity')
        assert (bench.HW_H * bench.HW_W, bench.R1_H, bench.R1_W) == (8160, 544, 960)
        assert bench.N_INIT == 100000 and bench.START_FRAME == 50 and frames == 30
        assert bench.select_workload('480p-2obj-100frame-clip-hotpath', 7) == 7
    finally:
        assert bench.select_workload('480p-2obj-100frame-clip-hotpath') == 100
    assert (bench.HW_H * bench.HW_W, bench.R1_H, bench.R1_W, bench.N_INIT, bench.START_FRAME) == (1620, 240, 432, None, 0)
    cfg = bench.config_dict(type('A', (), dict(frames=100, frac_merge=0.1))())
    assert set(cfg) == {'workload', 'hw', 'budget', 'frames', 'frac_merge', 'l2_policy'}     # same keys in both arms
