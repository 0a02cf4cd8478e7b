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
dist.destroy_process_group()
