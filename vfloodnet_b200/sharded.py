"""Feature bank sharded over GPUs: split-memory read with a log-sum-exp combine (BASELINE config 5).

Each rank owns a contiguous range of every object's bank slots in its own FeatureBank; the query features are
replicated.  Softmax over memory is associative through (max, sum-exp), so the read needs exactly two exchange steps:

    phase A (local kernels)  ->  all_gather of (m, l) per (object, query)      [obj_n * HW * 2 floats per rank]
    global LSE               ->  phase B (local kernels) against the GLOBAL LSE [usage counts are exact and local]
    partial readouts         ->  all_reduce(sum) of (obj_n, d_val, HW)          [3.3 MB/object at 480p, 66 MB at 4K]

`combine_lse` and `reduce_readout` are backend-agnostic torch.distributed code (NCCL on the GPU box, gloo in the CPU
tests); the kernels behind phase A / phase B are vfn_memread_phase_a / vfn_memread_phase_b (include/vfn.h).
Stream-parallel operation (independent videos, one group per GPU) needs no collective at all: see bench.py --gpus N.

`ShardedFeatureBank` is the bank itself sharded over ranks (FeatureBank.update / remove, FeatureBank.py:53-143):

    local match              ->  all_gather of (c*, sequence id of j*) per query   [arg-max combine, ties -> earliest slot]
    merge                    ->  purely local on the rank that owns the matched slot (candidates are replicated)
    LFU eviction             ->  per threshold iteration: all_reduce(min) of the survivors' LFU minimum and
                                 all_reduce(sum) of the kept counts (two scalars), then a local order-preserving compaction
    append                   ->  the append set (ascending query order) is cut into `world` contiguous chunks, rank r
                                 appends chunk r; every slot carries a global sequence id (insertion order), so "lowest
                                 index" of the reference == lowest sequence id across shards
All exchange steps go through a small communicator interface (`DistComm` = torch.distributed over NCCL / gloo,
`ThreadComm` = ranks as threads of one process, used to test several shards on one GPU).
"""
from __future__ import annotations

import threading
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


class DistComm:
    """collectives over a torch.distributed process group (NCCL on NVLink/NVSwitch; gloo in the CPU tests)"""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, t: torch.Tensor) -> List[torch.Tensor]:
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    def all_reduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
        dist.all_reduce(t, op={'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN, 'max': dist.ReduceOp.MAX}[op],
                        group=self.group)
        return t


class ThreadComm:
    """`world` ranks as threads of ONE process sharing one device (tests: a sharded bank on a single GPU).
    make(world) returns one communicator per rank; collectives meet on a barrier."""

    class _Shared:
        def __init__(self, world):
            self.slots = [None] * world
            self.barrier = threading.Barrier(world)

    def __init__(self, shared, rank, world):
        self._s, self.rank, self.world = shared, rank, world

    @classmethod
    def make(cls, world: int):
        sh = cls._Shared(world)
        return [cls(sh, r, world) for r in range(world)]

    def all_gather(self, t: torch.Tensor) -> List[torch.Tensor]:
        self._s.slots[self.rank] = t.contiguous()
        self._s.barrier.wait()
        out = [x.clone() for x in self._s.slots]
        self._s.barrier.wait()
        return out

    def all_reduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
        st = torch.stack(self.all_gather(t))
        r = {'sum': st.sum(dim=0), 'min': st.min(dim=0).values, 'max': st.max(dim=0).values}[op]
        t.copy_(r.to(t.dtype))
        return t


def _comm(group=None, comm=None):
    return comm if comm is not None else DistComm(group)


def shard_range(n: int, rank: int, world: int):
    """contiguous slot range [lo, hi) of rank `rank` (insertion order preserved across ranks)"""
    return n * rank // world, n * (rank + 1) // world


def combine_lse(ml_local: torch.Tensor, group=None, comm=None) -> torch.Tensor:
    """ml_local (..., 2) = local (max, sum exp(s - max)), natural-log domain  ->  global LSE (...)"""
    ml = torch.stack(_comm(group, comm).all_gather(ml_local), dim=0)      # (world, ..., 2)
    m, l = ml[..., 0], ml[..., 1]
    M = m.max(dim=0).values
    safe = torch.where(torch.isinf(M), torch.zeros_like(M), M)   # ranks with empty shards contribute (-inf, 0)
    L = (l * torch.exp(m - safe)).sum(dim=0)
    return safe + torch.log(L)


def reduce_readout(partial: torch.Tensor, group=None, comm=None) -> torch.Tensor:
    """sum of the per-rank partial readouts sum_{i in shard} p_ij v_i (p normalised with the global LSE)"""
    return _comm(group, comm).all_reduce(partial, 'sum')


def combine_match(corr_local: torch.Tensor, idx_global: torch.Tensor, group=None, comm=None):
    """Global arg-max of the cosine match across shards, ties -> lowest GLOBAL slot (== the reference's lowest index,
    because shards keep insertion order).  corr_local (HW,) fp32, idx_global (HW,) int64 global slot / sequence ids.
    A NaN score (NaN candidate) yields best = NaN and the int64 maximum as slot: such a query goes nowhere."""
    cm = _comm(group, comm)
    c, i = torch.stack(cm.all_gather(corr_local)), torch.stack(cm.all_gather(idx_global))
    best = c.max(dim=0).values
    cand = torch.where(c == best.unsqueeze(0), i, torch.full_like(i, torch.iinfo(torch.int64).max))
    return best, cand.min(dim=0).values


def lfu_threshold_search(lfu: torch.Tensor, class_budget, request_n: int, comm):
    """FeatureBank.remove's threshold search (FeatureBank.py:117-138) over sharded LFU values: `lfu` holds this rank's
    slots; per iteration one exchange of the survivors' LFU minimum and one of the kept counts (scalars).
    T = int(min LFU) + 1; keep LFU > T (strict); while class_budget - kept - request_n < 0: T = int(min of survivors) + 1.
    Returns (kept_local, kept_global, T_final, [T sequence]); raises like the reference on a non-finite minimum
    (int() of nan/inf, :123) and when every slot is gone while the budget is still exceeded (min of empty, :136)."""
    dev = lfu.device

    def exchange(keep):
        """one collective: (survivors of this rank, their LFU minimum) of every rank -> (global count, global minimum)"""
        sel = lfu if keep is None else lfu[keep]
        mn = sel.min().double() if sel.numel() else torch.tensor(float('inf'), dtype=torch.float64, device=dev)
        pair = torch.stack([torch.tensor(float(sel.numel()), dtype=torch.float64, device=dev), mn])
        allp = torch.stack(comm.all_gather(pair))                       # torch.min propagates NaN, like the reference
        return int(allp[:, 0].sum().item()), float(allp[:, 1].min().item())

    _, mn = exchange(None)
    if not np.isfinite(mn):
        raise ValueError('FeatureBank.remove: LFU minimum is not finite (the reference raises in int(), '
                         'FeatureBank.py:123)')
    T = int(mn) + 1                                                                       # :123
    thresholds = []
    while True:
        keep = lfu > float(T)                                                             # strict >  (:127)
        kept_local = int(keep.sum().item())
        kept_global, mn = exchange(keep)
        thresholds.append(T)
        balance = (class_budget - kept_global) - request_n                                # :134
        if balance >= 0:
            break
        if kept_global == 0:
            raise RuntimeError('FeatureBank.remove: every entry was evicted and the budget is still exceeded '
                               '(the reference raises on LFU.min() of an empty tensor, FeatureBank.py:136)')
        T = int(mn) + 1                                                                   # :136
    return kept_local, kept_global, T, thresholds


class PeerExchange:
    """The exchange steps over PEER memory (NVLink / NVSwitch) instead of library collectives: every rank keeps its
    partial results in a symmetric-memory buffer that all GPUs of the node map, and the combine kernels of
    libvfn_sm100a.so (vfn_lse_combine_peers, vfn_reduce_peers / vfn_gather_peers, vfn_match_combine_peers) load their
    peers' partials directly.  A cross-GPU barrier on the stream (signal pads of the same symmetric allocation) stands
    before each combine; none is needed after it, because a rank overwrites a region only after the NEXT barrier of
    the sequence, which every peer reaches after it has finished reading (read: ml | barrier | lse, phase B -> po |
    barrier | reduce; update: pair | barrier | combine).  One process per GPU, torch.distributed initialised (NCCL)."""

    TWO_SHOT_BYTES = 16 << 20     # partial readouts beyond this are reduced slice-wise (reduce-scatter + all-gather)

    def __init__(self, group, device, obj_n: int, hw: int, d_val: int):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device(device)
        self.obj_n, self.hw, self.d_val = obj_n, hw, d_val
        al = lambda n: (n + 255) // 256 * 256
        self.off = {}
        total = 0
        for name, nbytes in (('ml', obj_n * hw * 8), ('po', obj_n * d_val * hw * 4), ('pair', obj_n * hw * 16)):
            self.off[name] = (total, nbytes)
            total += al(nbytes)
        self.buf = symm.empty(total, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self._ptrs = {}
        for name, (o, _n) in self.off.items():
            self._ptrs[name] = (C.c_void_p * self.world)(*[b + o for b in base])
        self.barriers = 0

    def fits(self, obj_n, hw, d_val):
        return (obj_n, hw, d_val) == (self.obj_n, self.hw, self.d_val)

    def local(self, name, shape, dtype):
        o, n = self.off[name]
        return self.buf[o:o + n].view(dtype).view(*shape)

    def peers(self, name):
        return self._ptrs[name]

    def barrier(self):
        self.hdl.barrier(channel=0)
        self.barriers += 1


class ShardedReader:
    """Split-memory Matcher.forward over a process group; `fb` is this rank's vfloodnet_b200.FeatureBank shard.
    peer = True: the two exchange steps run as kernels over peer memory (PeerExchange); otherwise through the
    communicator's collectives (NCCL / gloo / thread ranks)."""

    def __init__(self, thres_valid=1e-3, update_bank=True, group=None, comm=None, peer=False):
        self.thres_valid, self.update_bank, self.group, self.comm = thres_valid, update_bank, group, comm
        self.peer = peer
        self.px: Optional[PeerExchange] = None
        self._ws = None
        self._bufs = {}

    def _buf(self, name, shape, dev):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != dev:
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            self._bufs[name] = t
        return t

    def exchange(self, fb, hw, d_val) -> Optional['PeerExchange']:
        if not self.peer:
            return None
        if self.px is None or not self.px.fits(fb.obj_n, hw, d_val):
            self.px = PeerExchange(self.group, fb.device, fb.obj_n, hw, d_val)
        return self.px

    def __call__(self, fb, q_in: torch.Tensor, q_out: torch.Tensor) -> torch.Tensor:
        from . import _lib
        from ._lib import check, on_device, ptr, stream_ptr
        lib = _lib.load()
        dev = fb.device
        q_in = q_in.to(dev, torch.float32).contiguous()
        q_out = q_out.to(dev, torch.float32).contiguous()
        _, d_key, hw = q_in.shape
        d_val = q_out.shape[1]
        obj_n = fb.obj_n
        if any(s is None for s in fb._slabs):
            raise RuntimeError('sharded read on an empty feature bank: call init_bank() first')
        n_max = max(s.cap for s in fb._slabs)
        need = lib.vfn_memread_workspace_bytes(obj_n, n_max, hw, d_key, d_val)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        banks = fb.bank_array()
        px = self.exchange(fb, hw, d_val)
        with on_device(dev):
            st = stream_ptr()
            ml = px.local('ml', (obj_n, hw, 2), torch.float32) if px else self._buf('ml', (obj_n, hw, 2), dev)
            check(lib.vfn_memread_phase_a(banks, obj_n, ptr(q_in), hw, ptr(ml), ptr(self._ws), self._ws.numel(), fb.impl,
                                          st), 'memread_phase_a')
            if px:                                                                              # exchange step 1
                lse = self._buf('lse', (obj_n, hw), dev)
                px.barrier()
                check(lib.vfn_lse_combine_peers(px.peers('ml'), px.world, obj_n * hw, ptr(lse), st), 'lse_combine_peers')
            else:
                lse = combine_lse(ml, self.group, self.comm).contiguous()
            partial = px.local('po', (obj_n, d_val, hw), torch.float32) if px else self._buf('po', (obj_n, d_val, hw), dev)
            check(lib.vfn_memread_phase_b(banks, obj_n, ptr(q_in), hw, ptr(lse), float(self.thres_valid),
                                          int(self.update_bank), ptr(partial), ptr(self._ws), self._ws.numel(), fb.impl,
                                          st), 'memread_phase_b')
            out = torch.empty((1, obj_n, 2 * d_val, hw), dtype=torch.float32, device=dev)
            if px:                                                                              # exchange step 2
                total = obj_n * d_val * hw
                mem = self._buf('mem', (obj_n, d_val, hw), dev)
                px.barrier()
                if total * 4 <= px.TWO_SHOT_BYTES or px.world == 1:
                    check(lib.vfn_reduce_peers(px.peers('po'), px.world, 0, total, ptr(mem), st), 'reduce_peers')
                else:   # reduce-scatter into my slice of my own buffer, then all-gather the slices
                    sl = (total // px.world) // 4 * 4
                    first = px.rank * sl
                    count = sl if px.rank < px.world - 1 else total - first
                    check(lib.vfn_reduce_peers(px.peers('po'), px.world, first, count, ptr(partial), st), 'reduce_peers')
                    px.barrier()
                    check(lib.vfn_gather_peers(px.peers('po'), px.world, px.rank, sl, total, ptr(mem), st), 'gather_peers')
            else:
                mem = reduce_readout(partial, self.group, self.comm)
            out[0, :, :d_val] = mem
            out[0, :, d_val:] = q_out[0]
        return out


class ShardedFeatureBank:
    """FeatureBank (FeatureBank.py:8-149) with every object's slots sharded over the ranks of a communicator.

    Each rank keeps its slots in a private `vfloodnet_b200.FeatureBank` (`self.local`: the same HBM slabs and kernels
    as the single-GPU bank) plus one int64 *sequence id* per slot = the slot's position in the global insertion order.
    Local slot order is always ascending in sequence id (appends are ascending, compaction preserves order), so the
    reference's "lowest index wins" is "lowest sequence id wins" across shards.  Candidates (`prev_key`, `prev_value`)
    are replicated on every rank, as SURVEY 8e prescribes (they are tiny next to the bank).

    The global decisions (match, merge pairs, append set, eviction thresholds, kept set) are identical to those of one
    unsharded bank holding the same slots: `tests/test_gpu_parity.py::test_sharded_update_equals_single_bank` checks the
    gathered bank bit for bit.  `class_budget`, `peak_n`, `replace_n` keep their reference meaning (global counts).
    """

    def __init__(self, obj_n, memory_budget, device, update_rate=0.1, thres_close=0.95, *, comm=None, group=None,
                 impl: int = 0, peer: bool = False):
        from .feature_bank import FeatureBank
        self.comm = _comm(group, comm)
        self.rank, self.world = self.comm.rank, self.comm.world
        self.local = FeatureBank(obj_n, memory_budget, device, update_rate, thres_close, impl=impl)
        self.local.defer = False                       # every step below needs exact local sizes
        self.obj_n, self.device = obj_n, self.local.device
        self.update_rate, self.thres_close = update_rate, thres_close
        self.class_budget = self.local.class_budget    # GLOBAL budget per object (FeatureBank.py:20-22)
        self.peak_n = np.zeros(obj_n)
        self.replace_n = np.zeros(obj_n)
        self.seq: List[Optional[torch.Tensor]] = [None] * obj_n    # (n_local,) int64, ascending
        self.next_seq = [0] * obj_n
        self.n_global = [0] * obj_n
        self.last_decisions = [None] * obj_n
        self.last_thresholds_obj = [None] * obj_n
        self.group, self.peer = group, peer
        self.reader = ShardedReader(comm=self.comm, group=group, peer=peer)

    # ---- sizes -----------------------------------------------------------------------------------
    def n_local(self, c: int) -> int:
        return self.local._n[c]

    def _sum_int(self, v: int) -> int:
        t = torch.tensor([v], dtype=torch.int64, device=self.device)
        return int(self.comm.all_reduce(t, 'sum').item())

    # ---- reference methods ------------------------------------------------------------------------
    def init_bank(self, keys, values, frame_idx=0):
        """FeatureBank.py:27-36 with the (replicated) inputs cut into contiguous per-rank ranges"""
        ks, vs = [], []
        for c in range(self.obj_n):
            n = keys[c].shape[1]
            lo, hi = shard_range(n, self.rank, self.world)
            if hi - lo < 1:
                raise ValueError('every shard needs at least one slot at init_bank (n >= world size)')
            ks.append(keys[c][:, lo:hi])
            vs.append(values[c][:, lo:hi])
            self.seq[c] = torch.arange(lo, hi, dtype=torch.int64, device=self.device)
            self.next_seq[c] = n
            self.n_global[c] = n
            self.peak_n[c] = max(self.peak_n[c], n)
        self.local.init_bank(ks, vs, frame_idx)

    def read(self, q_in, q_out, update_bank=True):
        """Matcher.forward over the shards (AFB_URR.py:136-178): two exchange steps, see ShardedReader"""
        self.reader.update_bank = update_bank
        return self.reader(self.local, q_in, q_out)

    def update(self, prev_key, prev_value, frame_idx, update_rate=-1):
        """FeatureBank.py:53-115 for all objects; collective (every rank calls it with the same candidates)"""
        import ctypes as C
        from . import _lib
        from ._lib import check, ptr, stream_ptr
        from .feature_bank import _Slab
        if update_rate == -1:
            update_rate = self.update_rate
        lib, fb, dev = _lib.load(), self.local, self.device
        nan = float('nan')
        obj_n = self.obj_n
        hw = prev_key[0].shape[1]
        px = self.reader.exchange(fb, hw, prev_value[0].shape[0])
        st = stream_ptr()
        # ---- phase I, every object: candidates, local match, and this rank's (c*, sequence id) for the exchange
        pre = []
        for c in range(obj_n):
            pk = prev_key[c].to(dev, torch.float32).contiguous()
            pv = prev_value[c].to(dev, torch.float32).contiguous()
            d_key, hw_c = pk.shape
            d_val = pv.shape[0]
            n_loc = fb._n[c]
            s = fb._slabs[c]
            if (d_key, d_val) != (s.d_key, s.d_val) or pv.shape[1] != hw or hw_c != hw:
                raise ValueError('candidate dims do not match the bank')
            # (1) candidates -> entry-major raw + normalised (FeatureBank.py:64,88)
            ck, nck = fb._buf(f'sh_ck{c}', (hw, d_key), torch.float32), fb._buf(f'sh_nck{c}', (hw, d_key), torch.float32)
            cv, ncv = fb._buf(f'sh_cv{c}', (hw, d_val), torch.float32), fb._buf(f'sh_ncv{c}', (hw, d_val), torch.float32)
            check(lib.vfn_prep_rows(ptr(pk), d_key, hw, ptr(ck), ptr(nck), None, None, 1.0, st), 'prep_rows')
            check(lib.vfn_prep_rows(ptr(pv), d_val, hw, ptr(cv), ptr(ncv), None, None, 1.0, st), 'prep_rows')
            # (2) local match (FeatureBank.py:66-68)
            idx = fb._buf(f'sh_idx{c}', (hw,), torch.int32)
            corr = fb._buf(f'sh_corr{c}', (hw,), torch.float32)
            if n_loc > 0:
                mws = fb._buf('sh_mws', (lib.vfn_bank_match_workspace_bytes(n_loc, hw),), torch.uint8)
                bank = fb.bank_struct(c)
                check(lib.vfn_bank_match(C.byref(bank), ptr(nck), hw, ptr(idx), ptr(corr), ptr(mws), mws.numel(),
                                         int(fb.impl), st), 'bank_match')
            else:                                  # an empty shard never wins: (-inf, last possible sequence id)
                idx.zero_()
                corr.fill_(-float('inf'))
            if px:
                pair = px.local('pair', (obj_n, hw, 2), torch.int64)[c]
                check(lib.vfn_match_pack(ptr(corr), ptr(idx), ptr(self.seq[c]) if n_loc > 0 else None, n_loc, hw,
                                         ptr(pair), st), 'match_pack')
                seq_loc = None
            elif n_loc > 0:
                seq_loc = self.seq[c][idx.long()]
            else:
                seq_loc = torch.full((hw,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
            pre.append(dict(ck=ck, nck=nck, cv=cv, ncv=ncv, idx=idx, corr=corr, seq_loc=seq_loc, d_key=d_key, d_val=d_val))
        # ---- exchange step 1, all objects at once: arg-max combine over the shards (ties -> earliest slot)
        if px:
            best_all = fb._buf('sh_best', (obj_n, hw), torch.float32)
            seq_all = fb._buf('sh_bseq', (obj_n, hw), torch.int64)
            px.barrier()
            check(lib.vfn_match_combine_peers(px.peers('pair'), px.world, obj_n * hw, ptr(best_all), ptr(seq_all), st),
                  'match_combine_peers')
            pair_all = px.local('pair', (obj_n, hw, 2), torch.int64)
            my_seq = pair_all[..., 1]
            my_corr = torch.stack([p['corr'] for p in pre])
        else:
            my_corr = torch.stack([p['corr'] for p in pre])
            my_seq = torch.stack([p['seq_loc'] for p in pre])
            best_all, seq_all = combine_match(my_corr, my_seq, comm=self.comm)
        mine_all = (my_corr == best_all) & (my_seq == seq_all)
        is_append_all = best_all <= self.thres_close
        n_append_all = [int(v) for v in is_append_all.sum(dim=1).tolist()]          # the reference's nonzero() sync, once
        # ---- phase II, per object: plan, merge, LFU eviction against the global budget, append, clamp
        for c in range(obj_n):
            p = pre[c]
            ck, nck, cv, ncv, idx, corr = p['ck'], p['nck'], p['cv'], p['ncv'], p['idx'], p['corr']
            d_key, d_val = p['d_key'], p['d_val']
            n_loc = fb._n[c]
            s = fb._slabs[c]
            best, best_seq, mine = best_all[c], seq_all[c], mine_all[c]
            # (3) global classification (FeatureBank.py:71,100): strict > merges, <= appends, NaN goes nowhere
            is_merge = best > self.thres_close
            is_append = is_append_all[c]
            pos = torch.cumsum(is_append.to(torch.int64), 0) - 1
            n_append = n_append_all[c]
            a_lo, a_hi = shard_range(n_append, self.rank, self.world)
            my_append = is_append & (pos >= a_lo) & (pos < a_hi)
            # local plan: owned merges keep their score, my chunk of the append set keeps its score, the rest -> NaN
            corr_plan = torch.where((mine & is_merge) | my_append, best, torch.full_like(best, nan)).contiguous()
            merge_q = fb._buf(f'sh_mq{c}', (hw,), torch.int32)
            merge_slot = fb._buf(f'sh_ms{c}', (hw,), torch.int32)
            run_off = fb._buf(f'sh_ro{c}', (hw + 1,), torch.int32)
            append_q = fb._buf(f'sh_aq{c}', (hw,), torch.int32)
            counts = fb._buf(f'sh_cnt{c}', (4,), torch.int32)
            pws = fb._buf('sh_pws', (lib.vfn_bank_plan_workspace_bytes(hw),), torch.uint8)
            check(lib.vfn_bank_plan(ptr(idx), ptr(corr_plan), hw, float(self.thres_close), ptr(merge_q), ptr(merge_slot),
                                    ptr(run_off), ptr(append_q), ptr(counts), None, ptr(pws), pws.numel(), st), 'bank_plan')
            if n_loc > 0:
                bank = s.struct(n_loc)
                check(lib.vfn_bank_merge(C.byref(bank), ptr(nck), ptr(ncv), ptr(merge_q), ptr(merge_slot), ptr(run_off),
                                         ptr(counts), hw, float(update_rate), st), 'bank_merge')
            # (4) LFU eviction against the GLOBAL budget (FeatureBank.py:102-103 -> remove, :117-143)
            n_before = self.n_global[c]
            evicted, thresholds, kept_global = False, [], n_before
            if self.class_budget < n_before + n_append:
                evicted = True
                kept_local, kept_global, T, thresholds, lfu = self._evict_search(c, frame_idx, n_append)
                if n_loc > 0:
                    alt = fb._alt[c]
                    if alt is None or alt.cap < s.cap:
                        alt = _Slab(s.d_key, s.d_val, s.cap, dev, s.kh is not None)
                    plan = torch.tensor([0, kept_local, len(thresholds), T], dtype=torch.int32, device=dev)
                    cws = fb._buf(f'sh_cws{c}', (lib.vfn_bank_compact_workspace_bytes(n_loc),), torch.uint8)
                    src, dst = s.struct(n_loc), alt.struct(0)
                    check(lib.vfn_bank_compact(C.byref(src), C.byref(dst), ptr(lfu), ptr(plan), ptr(cws), cws.numel(), st),
                          'bank_compact')
                    self.seq[c] = self.seq[c][lfu > float(T)]
                    fb._slabs[c], fb._alt[c] = alt, s
                    fb._n[c] = n_loc = kept_local
                    s = alt
                self.replace_n[c] += n_before - kept_global                               # FeatureBank.py:140-141
            # (5) append my chunk of the append set (ascending query order), then clamp (FeatureBank.py:105-115)
            n_mine = a_hi - a_lo
            if n_mine > 0:
                fb._ensure_capacity(c, n_loc + n_mine, d_key, d_val, slack=hw)
                s = fb._slabs[c]
                bank = s.struct(n_loc)
                check(lib.vfn_bank_append_rows(C.byref(bank), ptr(ck), ptr(cv), ptr(nck), ptr(append_q), n_mine, None,
                                               float(frame_idx), 0.0, st), 'append_rows')
                self.seq[c] = torch.cat([self.seq[c], self.next_seq[c] + torch.arange(a_lo, a_hi, dtype=torch.int64,
                                                                                      device=dev)])
                fb._n[c] = n_loc = n_loc + n_mine
            self.next_seq[c] += n_append
            self.n_global[c] = kept_global + n_append
            if n_loc > 0:
                bank = fb._slabs[c].struct(n_loc)
                check(lib.vfn_bank_clamp_info(C.byref(bank), n_loc, st), 'clamp_info')
            fb._set_live(c)
            self.peak_n[c] = max(self.peak_n[c], self.n_global[c])                        # FeatureBank.py:113
            self.last_thresholds_obj[c] = thresholds
            self.last_decisions[c] = dict(match_corr=best, match_seq=best_seq, is_merge=is_merge, is_append=is_append,
                                          n_append=n_append, evicted=evicted, kept=kept_global)

    def _evict_search(self, c: int, frame_idx, request_n: int):
        fb = self.local
        info = fb._slabs[c].info[:fb._n[c]]
        lfu = (info[:, 1] / (float(frame_idx) - info[:, 0])).contiguous()                # :121-122 (fp32, IEEE divide)
        kept_local, kept_global, T, thresholds = lfu_threshold_search(lfu, self.class_budget, request_n, self.comm)
        return kept_local, kept_global, T, thresholds, lfu

    def print_peak_mem(self):
        ur = self.peak_n / self.class_budget
        rr = self.replace_n / self.class_budget
        print(f'Obj num: {self.obj_n}.', f'Budget / obj: {self.class_budget}.', f'UR: {ur}.', f'Replace: {rr}.')

    # ---- test / parity helpers --------------------------------------------------------------------
    def _gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """concatenate per-rank (n_r, ...) tensors in rank order (ragged sizes: padded to the maximum)"""
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=self.device)
        ns = [int(x.item()) for x in self.comm.all_gather(n)]
        pad = torch.zeros((max(ns),) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        pad[:t.shape[0]] = t
        parts = self.comm.all_gather(pad)
        return torch.cat([p[:k] for p, k in zip(parts, ns)])

    def gather_state(self, c: int):
        """(keys (d_key,N), values (d_val,N), info (N,2), seq (N)) of object c in GLOBAL order (= reference order)"""
        n, s = self.local._n[c], self.local._slabs[c]
        seq = self._gather_rows(self.seq[c])
        order = torch.argsort(seq)
        g = lambda a: self._gather_rows(a[:n])[order]
        return g(s.keys).t(), g(s.values).t(), g(s.info), seq[order]

    def scatter_info(self, c: int, info_full: torch.Tensor):
        """teacher forcing: overwrite info with the rows of a (N,2) tensor given in GLOBAL order"""
        seq_all = torch.sort(self._gather_rows(self.seq[c])).values
        rows = torch.searchsorted(seq_all, self.seq[c])
        n = self.local._n[c]
        self.local._slabs[c].info[:n].copy_(info_full.to(self.device, torch.float32)[rows])
