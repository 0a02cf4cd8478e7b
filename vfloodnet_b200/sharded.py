"""Feature bank sharded over GPUs: split-memory read with a log-sum-exp combine (BASELINE config 5).

Each rank owns a contiguous range of every object's bank slots in its own FeatureBank; the query features are
replicated.  Softmax over memory is associative through (max, sum-exp), so the read needs exactly two exchange steps:

    phase A (local kernels)  ->  all_gather of (m, l) per (object, query)      [obj_n * HW * 2 floats per rank]
    global LSE               ->  phase B (local kernels) against the GLOBAL LSE [usage counts are exact and local]
    partial readouts         ->  all_reduce(sum) of (obj_n, d_val, HW)          [3.3 MB/object at 480p, 66 MB at 4K]

`combine_lse` and `reduce_readout` are backend-agnostic torch.distributed code (NCCL on the GPU box, gloo in the CPU
tests); the kernels behind phase A / phase B are vfn_memread_phase_a / vfn_memread_phase_b (include/vfn.h).
Stream-parallel operation (independent videos, one group per GPU) needs no collective at all: see bench.py --gpus N.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """contiguous slot range [lo, hi) of rank `rank` (insertion order preserved across ranks)"""
    return n * rank // world, n * (rank + 1) // world


def combine_lse(ml_local: torch.Tensor, group=None) -> torch.Tensor:
    """ml_local (..., 2) = local (max, sum exp(s - max)), natural-log domain  ->  global LSE (...)"""
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(ml_local) for _ in range(world)]
    dist.all_gather(gathered, ml_local.contiguous(), group=group)
    ml = torch.stack(gathered, dim=0)                     # (world, ..., 2)
    m, l = ml[..., 0], ml[..., 1]
    M = m.max(dim=0).values
    safe = torch.where(torch.isinf(M), torch.zeros_like(M), M)   # ranks with empty shards contribute (-inf, 0)
    L = (l * torch.exp(m - safe)).sum(dim=0)
    return safe + torch.log(L)


def reduce_readout(partial: torch.Tensor, group=None) -> torch.Tensor:
    """sum of the per-rank partial readouts sum_{i in shard} p_ij v_i (p normalised with the global LSE)"""
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def combine_match(corr_local: torch.Tensor, idx_global: torch.Tensor, group=None):
    """Global arg-max of the cosine match across shards, ties -> lowest GLOBAL slot (== the reference's lowest index,
    because shards keep insertion order).  corr_local (HW,) fp32, idx_global (HW,) int64 global slot ids."""
    world = dist.get_world_size(group)
    cs = [torch.empty_like(corr_local) for _ in range(world)]
    ix = [torch.empty_like(idx_global) for _ in range(world)]
    dist.all_gather(cs, corr_local.contiguous(), group=group)
    dist.all_gather(ix, idx_global.contiguous(), group=group)
    c, i = torch.stack(cs), torch.stack(ix)
    best = c.max(dim=0).values
    cand = torch.where(c == best.unsqueeze(0), i, torch.full_like(i, torch.iinfo(torch.int64).max))
    return best, cand.min(dim=0).values


class ShardedReader:
    """Split-memory Matcher.forward over a process group; `fb` is this rank's vfloodnet_b200.FeatureBank shard."""

    def __init__(self, thres_valid=1e-3, update_bank=True, group=None):
        self.thres_valid, self.update_bank, self.group = thres_valid, update_bank, group
        self._ws = None

    def __call__(self, fb, q_in: torch.Tensor, q_out: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        from . import _lib
        from ._lib import check, ptr, stream_ptr
        lib = _lib.load()
        dev = fb.device
        q_in = q_in.to(dev, torch.float32).contiguous()
        q_out = q_out.to(dev, torch.float32).contiguous()
        _, d_key, hw = q_in.shape
        d_val = q_out.shape[1]
        n_max = max(s.cap for s in fb._slabs)
        need = lib.vfn_memread_workspace_bytes(fb.obj_n, n_max, hw, d_key, d_val)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        banks = fb.bank_array()
        ml = torch.empty((fb.obj_n, hw, 2), dtype=torch.float32, device=dev)
        check(lib.vfn_memread_phase_a(banks, fb.obj_n, ptr(q_in), hw, ptr(ml), ptr(self._ws), self._ws.numel(), fb.impl,
                                      stream_ptr()), 'memread_phase_a')
        lse = combine_lse(ml, self.group).contiguous()                               # exchange step 1
        partial = torch.empty((fb.obj_n, d_val, hw), dtype=torch.float32, device=dev)
        check(lib.vfn_memread_phase_b(banks, fb.obj_n, ptr(q_in), hw, ptr(lse), float(self.thres_valid),
                                      int(self.update_bank), ptr(partial), ptr(self._ws), self._ws.numel(), fb.impl,
                                      stream_ptr()), 'memread_phase_b')
        mem = reduce_readout(partial, self.group)                                    # exchange step 2
        out = torch.cat([mem, q_out.expand(fb.obj_n, -1, -1)], dim=1).unsqueeze(0)   # (1, obj_n, 2*d_val, HW)
        return out
