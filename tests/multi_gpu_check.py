"""2+ GPU check of the sharded-bank read over NCCL (run under torchrun on a multi-GPU box; not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Each rank owns a contiguous slice of the bank; the combined read must equal the single-GPU read."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as vfn
from vfloodnet_b200 import sharded, synth

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
n, hw = int(os.environ.get('VFN_N', 100000)), int(os.environ.get('VFN_HW', 1620))
g = torch.Generator().manual_seed(3)
keys, vals = zip(*[synth.gen_bank(g, n) for _ in range(2)])
info = [synth.gen_info(g, n, 9) for _ in range(2)]
q_in, q_out = synth.gen_query(g, hw)
lo, hi = sharded.shard_range(n, rank, world)
fb = vfn.FeatureBank(2, 10 ** 7, dev)
fb.load_state([k[:, lo:hi] for k in keys], [v[:, lo:hi] for v in vals], [i[lo:hi] for i in info])
reader = sharded.ShardedReader(update_bank=True)
out = reader(fb, q_in, q_out)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(10):
    out = reader(fb, q_in, q_out)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
if rank == 0:
    full = vfn.FeatureBank(2, 10 ** 7, dev)
    full.load_state(list(keys), list(vals), info)
    ref = vfn.Matcher(update_bank=False)(full, q_in, q_out)
    err = (out - ref).abs().max().item()
    print(f'sharded read over {world} GPUs: N={n} HW={hw}  max|sharded - single| = {err:.3e}  {dt*1e3:.3f} ms/read')
    assert err < 2e-4

# ---- sharded UPDATE over NCCL: gathered bank must equal the single-GPU bank bit for bit -------------------------
g = torch.Generator().manual_seed(17)
n0, hw2, frames, budget = 20000, 1620, 6, 60000           # class_budget 24000 -> eviction after a few frames
keys, vals = zip(*[synth.gen_bank(g, n0) for _ in range(2)])
clip = []
for t in range(frames):
    pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw2, 0.3) for c in range(2)])
    usage = [(torch.rand(60000, generator=g) * (12.0 * (t + 1))).to(dev) for _ in range(2)]   # info[:,1] teacher forcing
    clip.append(([k.to(dev) for k in pk], [v.to(dev) for v in pv], usage))
q_in, q_out = synth.gen_query(g, hw2)
sfb = sharded.ShardedFeatureBank(2, budget, dev)
sfb.init_bank(list(keys), list(vals))
full = vfn.FeatureBank(2, budget, dev) if rank == 0 else None
if rank == 0:
    full.init_bank(list(keys), list(vals))
    m = vfn.Matcher(update_bank=False)
torch.cuda.synchronize()
dist.barrier()
t_upd = t_read = 0.0
n_evict = 0
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
    t0 = time.perf_counter()
    out = sfb.read(q_in, q_out, update_bank=False)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sfb.update(pk, pv, t + 1)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t_read += t1 - t0
    t_upd += t2 - t1
    n_evict += int(sfb.last_decisions[0]['evicted'])
    states = [sfb.gather_state(c) for c in range(2)]
    if rank == 0:
        ref = m(full, q_in, q_out)
        full.update(pk, pv, t + 1)
        assert (out - ref).abs().max().item() < 2e-4
        for c in range(2):
            k_g, v_g, i_g, _ = states[c]
            assert k_g.shape[1] == full.bank_n(c), (t, c, k_g.shape[1], full.bank_n(c))
            assert torch.equal(k_g, full.keys[c]) and torch.equal(v_g, full.values[c]), (t, c)
            assert torch.equal(i_g, full.info[c]), (t, c)
if rank == 0:
    print(f'sharded read+update over {world} GPUs: {frames} frames, bank {[sfb.n_global[c] for c in range(2)]} slots, '
          f'{n_evict} evictions, gathered bank == single-GPU bank (bit-exact keys/values); '
          f'{t_read / frames * 1e3:.2f} ms/read, {t_upd / frames * 1e3:.2f} ms/update')
dist.destroy_process_group()
