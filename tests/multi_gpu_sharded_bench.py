"""Sharded feature bank over the GPUs of one node: parity against the single-GPU bank and the 4K measurement of
BASELINE.json configs[4].  Run under torchrun (also reachable as
`bench.py --workload 4k-2obj-sharded-bank --gpus N`):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      tests/multi_gpu_sharded_bench.py [--frames 24] [--budget 2000000] [--no-4k]

Part 1 (480p sizes, both exchange back ends: kernels over peer memory / NCCL collectives): a clip with merges, appends and
LFU evictions; after every frame the bank gathered from the shards must equal the single-GPU bank BIT FOR BIT (keys,
values, info) and the split-memory read must match the single-GPU read.
Part 2 (4K: HW = 32400, --budget slots over two objects, bank at capacity): read and update latency of the sharded bank
with both back ends, the exchange cost alone, and the single-GPU read / update of the SAME bank on rank 0.
One JSON line on rank 0 (also written to gpurun_out/sharded_bench_n{N}.json)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vfloodnet_b200 as vfn            # noqa: E402
from vfloodnet_b200 import sharded, synth   # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def parity(dev, rank, world, peer, frames):
    """returns dict(frames, evictions, max_read_err, ms_read, ms_update) and asserts bit-exactness on rank 0"""
    g = torch.Generator().manual_seed(17)
    n0, hw, budget = 20000, 1620, 60000           # class_budget 24000 -> eviction after a few frames
    keys, vals = zip(*[synth.gen_bank(g, n0) for _ in range(2)])
    clip = []
    for t in range(frames):
        pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw, 0.3) for c in range(2)])
        usage = [(torch.rand(60000, generator=g) * (12.0 * (t + 1))).to(dev) for _ in range(2)]   # info[:,1] teacher forcing
        clip.append(([k.to(dev) for k in pk], [v.to(dev) for v in pv], usage))
    q_in, q_out = [t.to(dev) for t in synth.gen_query(g, hw)]
    sfb = sharded.ShardedFeatureBank(2, budget, dev, peer=peer)
    sfb.init_bank(list(keys), list(vals))
    full = m = None
    if rank == 0:
        full = vfn.FeatureBank(2, budget, dev)
        full.init_bank(list(keys), list(vals))
        m = vfn.Matcher(update_bank=False)
    # communicators of every collective the update can issue are created before anything is timed
    sharded.lfu_threshold_search(torch.ones(4, device=dev) * (rank + 2), 1e9, 0, sfb.comm)
    torch.cuda.synchronize()
    dist.barrier()
    t_upd = t_read = 0.0
    n_evict, max_err = 0, 0.0
    for t, (pk, pv, usage) in enumerate(clip):
        # identical usage columns on both sides, so that threshold-band count flips of a read cannot change LFU decisions
        for c in range(2):
            i_g = sfb.gather_state(c)[2].clone()
            i_g[:, 1] = usage[c][:i_g.shape[0]]
            sfb.scatter_info(c, i_g)
            if rank == 0:
                full.info[c][:, 1] = usage[c][:full.bank_n(c)]
        torch.cuda.synchronize()
        dist.barrier()
        e0 = ev()
        out = sfb.read(q_in, q_out, update_bank=False)
        e1 = ev()
        sfb.update(pk, pv, t + 1)
        e2 = ev()
        torch.cuda.synchronize()
        if t >= 2:
            t_read += e0.elapsed_time(e1)
            t_upd += e1.elapsed_time(e2)
        n_evict += int(sfb.last_decisions[0]['evicted'])
        states = [sfb.gather_state(c) for c in range(2)]
        if rank == 0:
            ref = m(full, q_in, q_out)
            full.update(pk, pv, t + 1)
            err = (out - ref).abs().max().item()
            max_err = max(max_err, err)
            assert err < 2e-4, (t, err)
            for c in range(2):
                k_g, v_g, i_g, _ = states[c]
                assert k_g.shape[1] == full.bank_n(c), (t, c, k_g.shape[1], full.bank_n(c))
                assert torch.equal(k_g, full.keys[c]) and torch.equal(v_g, full.values[c]), (t, c)
                assert torch.equal(i_g, full.info[c]), (t, c)
    k = max(frames - 2, 1)
    return dict(frames=frames, evictions=n_evict, max_read_err_vs_single=max_err, ms_read=t_read / k, ms_update=t_upd / k,
                bank=[sfb.n_global[c] for c in range(2)], gathered_bank_bit_exact=True)


def fill_shard(dev, rank, world, n_global, frame, seed):
    """this rank's contiguous range of a synthetic bank at capacity, generated on the device (seeded per object and rank)"""
    keys, vals, info, seq = [], [], [], []
    for c in range(2):
        lo, hi = sharded.shard_range(n_global, rank, world)
        g = torch.Generator(device=dev).manual_seed(seed * 1000 + c * 64 + rank)
        keys.append(torch.randn(128, hi - lo, generator=g, device=dev) * synth.S_K)
        vals.append(torch.randn(512, hi - lo, generator=g, device=dev))
        i = torch.zeros(hi - lo, 2, device=dev)
        i[:, 0] = torch.sort(torch.randint(0, frame, (hi - lo,), generator=g, device=dev).float()).values
        i[:, 1] = torch.rand(hi - lo, generator=g, device=dev) * 50
        info.append(i)
        seq.append(torch.arange(lo, hi, dtype=torch.int64, device=dev))
    return keys, vals, info, seq


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a = ev()
    for _ in range(reps):
        fn()
    b = ev()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_4k(dev, rank, world, budget, reps):
    hw, frame = 135 * 240, 60
    class_budget = int(0.8 * (budget // 2))
    res = dict(hw=hw, budget=budget, slots_per_object=class_budget, slots_per_object_per_gpu=class_budget // world)
    g = torch.Generator(device=dev).manual_seed(7)
    q_in = torch.randn(1, 128, hw, generator=g, device=dev) * synth.S_K
    q_out = torch.randn(1, 512, hw, generator=g, device=dev)
    keys, vals, info, seq = fill_shard(dev, rank, world, class_budget, frame, seed=3)
    outs = {}
    for peer in (True, False):
        name = 'peer' if peer else 'nccl'
        sfb = sharded.ShardedFeatureBank(2, budget, dev, peer=peer)
        sfb.local.load_state(keys, vals, info)
        for c in range(2):
            sfb.seq[c], sfb.next_seq[c], sfb.n_global[c] = seq[c], class_budget, class_budget
        res[f'ms_read_{name}'] = timed(lambda: sfb.read(q_in, q_out, update_bank=False), reps)
        outs[name] = sfb.read(q_in, q_out, update_bank=False)
        # the exchange alone: the same reader against one-tile shards (128 slots per object): phase A/B are then a few
        # microseconds and what remains is barriers + combine kernels (or the collectives) + launch overhead
        tiny = vfn.FeatureBank(2, 10 ** 6, dev)
        tiny.load_state([k[:, :128] for k in keys], [v[:, :128] for v in vals], [i[:128] for i in info])
        rd = sharded.ShardedReader(update_bank=False, peer=peer)
        rd.px = sfb.reader.px
        res[f'ms_exchange_{name}'] = timed(lambda: rd(tiny, q_in, q_out), reps * 4)
        if peer:
            res['peer_barriers_per_read'] = 2 if 2 * 512 * hw * 4 <= sharded.PeerExchange.TWO_SHOT_BYTES else 3
        # update at capacity with every candidate new (all append -> LFU eviction on every rank)
        pk = [torch.randn(128, hw, generator=g, device=dev) * synth.S_K for _ in range(2)]
        pv = [torch.randn(512, hw, generator=g, device=dev) for _ in range(2)]
        sfb.update(pk, pv, frame)                 # untimed: allocates the ping-pong slabs and grows the shard's capacity
        sfb.local.load_state(keys, vals, info)    # back to the bank at capacity (the allocations stay cached)
        for c in range(2):
            sfb.seq[c], sfb.next_seq[c], sfb.n_global[c] = seq[c], class_budget, class_budget
        torch.cuda.synchronize()
        dist.barrier()
        a = ev()
        sfb.update(pk, pv, frame)
        b = ev()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f'ms_update_{name}'] = float(t.item())
        res[f'update_{name}'] = dict(evicted=bool(sfb.last_decisions[0]['evicted']), bank_after=[sfb.n_global[c] for c in range(2)],
                                     thresholds=sfb.last_thresholds_obj[0])
        del sfb, tiny, rd
        torch.cuda.empty_cache()
    res['max_abs_peer_vs_nccl_read'] = (outs['peer'] - outs['nccl']).abs().max().item()
    assert res['max_abs_peer_vs_nccl_read'] <= 1e-4, res['max_abs_peer_vs_nccl_read']
    # the same bank on ONE GPU (rank 0): gather the shards' raw tensors over NCCL
    full_k = [torch.cat(_gather(k.t().contiguous(), dev, world)).t().contiguous() for k in keys]
    full_v = [torch.cat(_gather(v.t().contiguous(), dev, world)).t().contiguous() for v in vals]
    full_i = [torch.cat(_gather(i, dev, world)) for i in info]
    if rank == 0:
        full = vfn.FeatureBank(2, budget, dev)
        full.load_state(full_k, full_v, full_i)
        del full_k, full_v, full_i
        m = vfn.Matcher(update_bank=False)
        m(full, q_in, q_out)
        torch.cuda.synchronize()
        a = ev()
        for _ in range(max(reps // 2, 1)):
            ref = m(full, q_in, q_out)
        b = ev()
        torch.cuda.synchronize()
        res['ms_read_single_gpu'] = a.elapsed_time(b) / max(reps // 2, 1)
        res['max_abs_sharded_vs_single_read'] = (outs['peer'] - ref).abs().max().item()
        assert res['max_abs_sharded_vs_single_read'] <= 2e-4, res['max_abs_sharded_vs_single_read']
        res['read_speedup_vs_single_gpu'] = res['ms_read_single_gpu'] / res['ms_read_peer']
        # the same update (every candidate new, bank at capacity -> LFU eviction) on the single-GPU bank
        gq = torch.Generator(device=dev).manual_seed(7)
        torch.randn(1, 128, hw, generator=gq, device=dev); torch.randn(1, 512, hw, generator=gq, device=dev)   # q_in, q_out
        pk1 = [torch.randn(128, hw, generator=gq, device=dev) * synth.S_K for _ in range(2)]
        pv1 = [torch.randn(512, hw, generator=gq, device=dev) for _ in range(2)]
        st0 = ([full.keys[c].clone() for c in range(2)], [full.values[c].clone() for c in range(2)],
               [full.info[c].clone() for c in range(2)])
        full.update(pk1, pv1, frame)                    # untimed: allocations
        full.load_state(*st0)
        torch.cuda.synchronize()
        a = ev()
        full.update(pk1, pv1, frame)
        b = ev()
        torch.cuda.synchronize()
        res['ms_update_single_gpu'] = a.elapsed_time(b)
        res['update_single_gpu_bank_after'] = [full.bank_n(c) for c in range(2)]
        res['update_speedup_vs_single_gpu'] = res['ms_update_single_gpu'] / res['ms_update_peer']
        res['algorithmic_tflops_read_sharded'] = 1280.0 * class_budget * hw * 2 / (res['ms_read_peer'] * 1e-3) / 1e12
    dist.barrier()
    return res


def _gather(t, dev, world):
    n = torch.tensor([t.shape[0]], device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    mx = int(max(x.item() for x in ns))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:int(k.item())] for o, k in zip(out, ns)]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=24)
    ap.add_argument('--budget', type=int, default=2000000)
    ap.add_argument('--reps', type=int, default=6)
    ap.add_argument('--no-4k', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    args = ap.parse_args(argv)
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    line = {'workload': '4k-2obj-sharded-bank', 'n_gpus': world}
    if not args.no_parity:
        for peer in (True, False):
            line['parity_480p_' + ('peer' if peer else 'nccl')] = parity(dev, rank, world, peer, args.frames)
    if not args.no_4k:
        line['bench_4k'] = bench_4k(dev, rank, world, args.budget, args.reps)
    if rank == 0:
        print(json.dumps(line))
        out = os.path.join(ROOT, 'gpurun_out')
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f'sharded_bench_n{world}.json'), 'w') as f:
            json.dump(line, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
