"""Drop-in FeatureBank (reference: video_module/model/FeatureBank.py:8-149) on B200 HBM.

Same constructor, attributes (obj_n, keys, values, info, peak_n, replace_n, class_budget, update_rate,
thres_close, device) and methods (init_bank, append, update, remove, print_peak_mem) as the reference class.
The bank lives in capacity-sized entry-major device slabs; ``keys[i]`` / ``values[i]`` /
``info[i]`` are (d, N) / (d, N) / (N, 2) VIEWS of those slabs, valid until the next update()/remove()/append().
All arithmetic runs in libvfn_sm100a.so; there is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import collections
import ctypes as C
import math
from typing import List, Optional

import numpy as np
import torch

from . import _lib
import functools

from ._lib import VfnBank, VfnUpdateIO, check, em_backed, on_device, ptr, stream_ptr


def _on_bank_device(fn):
    """run a FeatureBank method with the bank's GPU as the current CUDA device: the library launches on the calling
    thread's current device and stream_ptr() is that device's current stream (a bank may live on any GPU, like the
    reference's `device` argument)"""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with on_device(self.device):
            return fn(self, *a, **k)
    return wrapped


class _Slab:
    """One object's device arrays (struct vfn_bank)."""

    def __init__(self, d_key: int, d_val: int, cap: int, device, operands: bool):
        self.d_key, self.d_val, self.cap = d_key, d_val, cap
        f32 = dict(dtype=torch.float32, device=device)
        self.keys = torch.empty((cap, d_key), **f32)
        self.values = torch.empty((cap, d_val), **f32)
        self.info = torch.zeros((cap, 2), **f32)
        self.nk = torch.empty((cap, d_key), **f32)
        self.cnt = torch.zeros((cap,), dtype=torch.int32, device=device)
        if operands:
            # tensor-core operands: zero-filled so that rows beyond n never hold NaN/Inf bit patterns
            f16 = dict(dtype=torch.float16, device=device)
            u8 = dict(dtype=torch.uint8, device=device)
            self.nkh = torch.zeros((cap, d_key), **f16)
            self.nkl = torch.zeros((cap, d_key), **f16)
            self.kh = torch.zeros((cap, d_key), **f16)
            self.kl = torch.zeros((cap, d_key), **f16)
            self.vh = torch.zeros((cap, d_val), **f16)
            self.v8 = torch.zeros((cap, d_val), **u8)
            self.vl = torch.zeros((cap, d_val), **u8)
        else:
            self.nkh = self.nkl = self.kh = self.kl = self.vh = self.v8 = self.vl = None

    _ARRAYS = ('keys', 'values', 'info', 'nk', 'nkh', 'nkl', 'kh', 'kl', 'vh', 'v8', 'vl', 'cnt')

    def struct(self, n: int, n_live=None, n_min: Optional[int] = None) -> VfnBank:
        """n_live: device int32[2] live-count cell of the bank (vfn.h); then n / n_min are upper / lower bounds"""
        return VfnBank(self.d_key, self.d_val, self.cap, n, ptr(self.keys), ptr(self.values), ptr(self.info),
                       ptr(self.nk), ptr(self.nkh), ptr(self.nkl), ptr(self.kh), ptr(self.kl), ptr(self.vh),
                       ptr(self.v8), ptr(self.vl), ptr(self.cnt), ptr(n_live) if self.kh is not None else None,
                       n if n_min is None else n_min)

    def copy_rows_from(self, other: '_Slab', n: int):
        for name in self._ARRAYS:
            a, b = getattr(self, name), getattr(other, name)
            if a is not None:
                a[:n].copy_(b[:n])


class _ViewList:
    """List-like exposing per-object views, so reference-style code (`fb.keys[i].size()`, `fb.info[i][:,1] += ..`) works."""

    def __init__(self, fb: 'FeatureBank', kind: str):
        self._fb, self._kind = fb, kind

    def __len__(self):
        return self._fb.obj_n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        self._fb._resolve()
        s, n = self._fb._slabs[i], self._fb._n[i]
        if s is None:
            return None
        if self._kind == 'keys':
            return s.keys[:n].t()
        if self._kind == 'values':
            return s.values[:n].t()
        return s.info[:n]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __bool__(self):
        return len(self) > 0


class FeatureBank:
    """See module docstring.  Extra keyword-only knobs (not in the reference): `impl` (0 auto, 1 SIMT fp32,
    2 tcgen05) selects the read kernels used by vfloodnet_b200.Matcher on this bank."""

    def __init__(self, obj_n, memory_budget, device, update_rate=0.1, thres_close=0.95, *, impl: int = 0):
        self.obj_n = obj_n
        self.update_rate = update_rate
        self.thres_close = thres_close
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ValueError('vfloodnet_b200.FeatureBank lives in GPU memory; device must be a CUDA device '
                             '(there is no CPU fallback)')
        self._peak_n = np.zeros(obj_n)
        self.replace_n = np.zeros(obj_n)
        self.class_budget = memory_budget // obj_n              # FeatureBank.py:20
        if obj_n == 2:
            self.class_budget = 0.8 * self.class_budget         # FeatureBank.py:21-22 (float)
        self.impl = impl
        self._lib = _lib.load()
        self._slabs: List[Optional[_Slab]] = [None] * obj_n
        self._alt: List[Optional[_Slab]] = [None] * obj_n      # ping-pong target of eviction compaction
        self._n = [0] * obj_n
        self._scratch = {}
        self._h_plan = torch.zeros((obj_n, 72), dtype=torch.int32).pin_memory()
        self._last_decisions = [None] * obj_n   # device tensors of the last update (tests / debugging)
        self.launches = 0                       # kernels launched by this bank (bench accounting)
        self.last_thresholds = []               # T sequence of the last remove() (tests / debugging)
        self.last_thresholds_obj = [None] * obj_n
        # Deferred completion of update() (vfn.h: vfn_bank_update / vfn_bank_update_finish) with a device-resident live
        # count per object (vfn_bank::n_live): an update that cannot evict does not read |A| back.  Its counts travel to a
        # pinned slot, the kernels of the following read / update take the live count from device memory, and the host
        # only keeps bounds (n_min <= live <= n).  Up to `run_ahead` updates stay unfinished; update t waits for the
        # counts of update t - run_ahead (normally long there), which refreshes the exact size and bounds the host's
        # lead over the GPU.  Anything that needs an exact size (keys/values/info views, bank_n, an update that may evict,
        # capacity growth) finishes everything first (_resolve).
        self.defer = True
        self._run_ahead = 3
        self._ring = 0
        self._size_ring(self._run_ahead + 1)    # pinned count slots / events: one more than the updates in flight
        self._seq = 0
        self._pending = collections.deque()     # unfinished updates, oldest first
        self._n_hi = [0] * obj_n                # upper bound of the live count (== _n when nothing is pending)
        self._n_live = torch.zeros((obj_n, 2), dtype=torch.int32, device=self.device)

    def _size_ring(self, ring: int):
        if ring > self._ring:
            self._ring = ring
            self._h_pinned = torch.zeros((ring, self.obj_n * 80), dtype=torch.int32).pin_memory()
            self._events = [None] * ring

    @property
    def run_ahead(self):
        """updates that may stay unfinished behind the host (0: every update is finished before the next call)"""
        return self._run_ahead

    @run_ahead.setter
    def run_ahead(self, depth: int):
        depth = int(depth)
        if depth < 0:
            raise ValueError('run_ahead must be >= 0')
        self._resolve()                         # nothing in flight while the ring of count slots is replaced
        self._run_ahead = depth
        self._size_ring(depth + 1)

    # ---- reference attribute surface -------------------------------------------------------------
    @property
    def keys(self):
        return _ViewList(self, 'keys') if any(s is not None for s in self._slabs) else None

    @property
    def values(self):
        return _ViewList(self, 'values') if any(s is not None for s in self._slabs) else None

    @property
    def info(self):
        return _ViewList(self, 'info')

    @property
    def peak_n(self):
        self._resolve()
        return self._peak_n

    @property
    def last_decisions(self):
        """per object: counts of the last finished update + its decision tensors.  The tensors are scratch buffers shared
        by all updates: they hold the data of the most recently ISSUED update (read them before the next update())."""
        self._resolve()
        return self._last_decisions

    def bank_n(self, class_idx: int) -> int:
        self._resolve()
        return self._n[class_idx]

    def bank_struct(self, class_idx: int) -> VfnBank:
        """exact view of one object's bank (finishes deferred updates first)"""
        self._resolve()
        return self._slabs[class_idx].struct(self._n[class_idx], self._n_live[class_idx])

    def _struct_bounds(self, c: int) -> VfnBank:
        """view with bounds: n = upper bound, n_min = last exact size; the kernels read the live count on the device"""
        return self._slabs[c].struct(self._n_hi[c], self._n_live[c], self._n[c])

    def _finish_oldest(self):
        """complete the oldest deferred update(): wait for its counts, advance the exact sizes (FeatureBank.py:105-113)"""
        pend = self._pending.popleft()
        slot, banks, io, dec, hw = pend
        self._events[slot].synchronize()
        check(self._lib.vfn_bank_update_finish(banks, self.obj_n, io, self._h_pinned[slot].data_ptr()),
              'bank_update_finish')
        later = sum(p[4] for p in self._pending)
        for c in range(self.obj_n):
            r = io[c]
            self._n[c] = int(banks[c].n)
            self._n_hi[c] = self._n[c] + later
            self._peak_n[c] = max(self._peak_n[c], self._n[c])                       # FeatureBank.py:113
            self._last_decisions[c] = dict(n_merge=int(r.n_merge), n_runs=int(r.n_runs), n_append=int(r.n_append),
                                           evicted=False, **dec[c])

    def _drain(self, keep: int):
        while len(self._pending) > keep:
            self._finish_oldest()

    def _resolve(self):
        """finish every deferred update: bank sizes are exact afterwards"""
        self._drain(0)

    @_on_bank_device
    def _set_live(self, c: int):
        """device-resident live count := exact host count (after ingest / remove, which bypass vfn_bank_update)"""
        s = self._slabs[c]
        if s is not None and s.kh is not None:
            bank = s.struct(self._n[c], self._n_live[c])
            check(self._lib.vfn_bank_set_live(C.byref(bank), self._n[c], stream_ptr()), 'bank_set_live')
            self.launches += 1
        self._n_hi[c] = self._n[c]

    def can_use_bounds(self, impl: int) -> bool:
        """True when a read may run against the live count without finishing the pending updates"""
        return bool(self._pending) and impl != 1 and all(s is not None and s.kh is not None for s in self._slabs)

    def bank_array(self, bounds_ok: bool = False, impl: int = 0):
        """struct vfn_bank[obj_n].  bounds_ok: the caller's kernels accept a device-resident live count (tcgen05 read),
        so pending updates need not be finished"""
        arr = (VfnBank * self.obj_n)()
        use_bounds = bounds_ok and self.can_use_bounds(impl)
        for c in range(self.obj_n):
            arr[c] = self._struct_bounds(c) if use_bounds else self.bank_struct(c)
        return arr

    # ---- internals -------------------------------------------------------------------------------
    def _use_operands(self, d_key, d_val):
        return d_key == 128 and d_val == 512

    def _budget_cap(self):
        return int(math.ceil(self.class_budget))

    @_on_bank_device
    def _ensure_capacity(self, c: int, needed: int, d_key: int, d_val: int, slack: int = 0):
        """grow geometrically up to the budget (+ one frame of candidates: update() needs cap >= n + hw)"""
        s = self._slabs[c]
        if s is not None and s.cap >= needed:
            return
        cap = max(needed, 4096)
        limit = max(self._budget_cap() + slack, needed)
        if s is not None:
            cap = max(cap, min(2 * s.cap, limit))
        else:
            cap = max(cap, min(4 * needed, limit))
        new = _Slab(d_key, d_val, cap, self.device, self._use_operands(d_key, d_val))
        if s is not None:
            new.copy_rows_from(s, self._n[c])
        self._slabs[c] = new
        self._alt[c] = None

    def _buf(self, name, shape, dtype):
        t = self._scratch.get(name)
        numel = int(np.prod(shape))
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self._scratch[name] = t
        return t[:numel].view(*shape)

    @_on_bank_device
    def _ingest(self, c: int, key_dm: torch.Tensor, val_dm: torch.Tensor, info0: float, info1: float):
        """append all columns of (d, n) tensors as new slots (init_bank / append)."""
        self._resolve()
        lib, st = self._lib, stream_ptr()
        key_dm = key_dm.to(self.device, torch.float32).contiguous()
        val_dm = val_dm.to(self.device, torch.float32).contiguous()
        d_key, n_new = key_dm.shape
        d_val = val_dm.shape[0]
        self._ensure_capacity(c, self._n[c] + n_new, d_key, d_val)
        ck = self._buf(f'ing_ck{c}', (n_new, d_key), torch.float32)
        cv = self._buf(f'ing_cv{c}', (n_new, d_val), torch.float32)
        check(lib.vfn_prep_rows(ptr(key_dm), d_key, n_new, ptr(ck), None, None, None, 1.0, st), 'prep_rows')
        check(lib.vfn_prep_rows(ptr(val_dm), d_val, n_new, ptr(cv), None, None, None, 1.0, st), 'prep_rows')
        bank = self._slabs[c].struct(self._n[c])      # exact host count; the device-resident count is set below
        check(lib.vfn_bank_append_rows(C.byref(bank), ptr(ck), ptr(cv), None, None, n_new, None, float(info0),
                                       float(info1), st), 'append_rows')
        self.launches += 3
        self._n[c] += n_new
        self._set_live(c)
        self._peak_n[c] = max(self._peak_n[c], self._n[c])

    # ---- reference methods -----------------------------------------------------------------------
    def init_bank(self, keys, values, frame_idx=0):
        """FeatureBank.py:27-36.  keys[i]: (d_key, n), values[i]: (d_val, n); copied into the slabs."""
        self._resolve()
        for c in range(self.obj_n):
            self._slabs[c], self._alt[c], self._n[c] = None, None, 0
            self._ingest(c, keys[c], values[c], frame_idx, 0.0)

    def append(self, keys, values, frame_idx=0):
        """FeatureBank.py:38-51 (info column 1 starts at 20 for appended entries, :46)."""
        if any(s is not None for s in self._slabs):
            for c in range(self.obj_n):
                self._ingest(c, keys[c], values[c], frame_idx, 20.0)
        else:
            self.init_bank(keys, values, frame_idx)

    @_on_bank_device
    def update(self, prev_key, prev_value, frame_idx, update_rate=-1):
        """FeatureBank.py:53-115: cosine match -> merge -> (LFU evict) -> append -> clamp, for all objects, as ONE call
        into the library (vfn_bank_update orders the kernel launches in C++)."""
        if update_rate == -1:
            update_rate = self.update_rate
        lib, st = self._lib, stream_ptr()
        obj_n = self.obj_n
        pk = [prev_key[c].to(self.device, torch.float32) for c in range(obj_n)]
        pv = [prev_value[c].to(self.device, torch.float32) for c in range(obj_n)]
        # candidates handed over entry-major (transposed views of (HW, d) storage: KeyValueHead) are read as they lie
        em = all(em_backed(x) for x in pk) and all(em_backed(x) for x in pv)
        if not em:
            pk, pv = [x.contiguous() for x in pk], [x.contiguous() for x in pv]
        d_key, hw = pk[0].shape
        d_val = pv[0].shape[0]
        for c in range(obj_n):
            s = self._slabs[c]
            if pk[c].shape != (s.d_key, hw) or pv[c].shape != (s.d_val, hw):
                raise ValueError('candidate dims do not match the bank')
        # run ahead of the GPU only while no object can reach its budget or its slab's capacity even at the upper bound
        # of its size (then FeatureBank.py:102 cannot fire and nothing on the host depends on |A|)
        self._drain(max(self.run_ahead - 1, 0) if self.defer else 0)
        bounds = (self.defer and self.run_ahead > 0 and self.impl != 1 and
                  all(s.kh is not None and self._n_hi[c] + hw <= self.class_budget and self._n_hi[c] + hw <= s.cap
                      for c, s in enumerate(self._slabs)))
        if not bounds:
            self._resolve()
        banks, alts = (VfnBank * obj_n)(), (VfnBank * obj_n)()
        io = (VfnUpdateIO * obj_n)()
        dec = []
        for c in range(obj_n):
            s = self._slabs[c]
            if not bounds:
                self._ensure_capacity(c, self._n[c] + hw, s.d_key, s.d_val, slack=hw)
                s = self._slabs[c]
                if self.class_budget < self._n[c] + hw:           # remove() may run: keep the ping-pong slab ready
                    alt = self._alt[c]
                    if alt is None or alt.cap < s.cap:
                        self._alt[c] = _Slab(s.d_key, s.d_val, s.cap, self.device, s.kh is not None)
                banks[c] = s.struct(self._n[c], self._n_live[c])
            else:
                banks[c] = self._struct_bounds(c)
            alts[c] = self._alt[c].struct(0) if self._alt[c] is not None else VfnBank()
            d = dict(match_idx=self._buf(f'midx{c}', (hw,), torch.int32),
                     match_corr=self._buf(f'mcorr{c}', (hw,), torch.float32),
                     merge_q=self._buf(f'merge_q{c}', (hw,), torch.int32),
                     merge_slot=self._buf(f'merge_slot{c}', (hw,), torch.int32),
                     run_off=self._buf(f'run_off{c}', (hw + 1,), torch.int32),
                     append_q=self._buf(f'append_q{c}', (hw,), torch.int32))
            dec.append(d)
            io[c].d_prev_key_dm, io[c].d_prev_value_dm = ptr(pk[c]), ptr(pv[c])
            io[c].prev_layout = 1 if em else 0
            io[c].d_match_idx, io[c].d_match_corr = ptr(d['match_idx']), ptr(d['match_corr'])
            io[c].d_merge_q, io[c].d_merge_slot = ptr(d['merge_q']), ptr(d['merge_slot'])
            io[c].d_run_off, io[c].d_append_q = ptr(d['run_off']), ptr(d['append_q'])
        n_max = max(int(banks[c].n) for c in range(obj_n))
        ws_bytes = lib.vfn_bank_update_workspace_bytes(obj_n, n_max, hw, d_key, d_val)
        ws = self._buf('upd_ws', (ws_bytes,), torch.uint8)
        l0 = lib.vfn_launch_count()
        ev, slot = None, self._seq % self._ring
        if self.defer:
            if self._events[slot] is None:
                self._events[slot] = torch.cuda.Event()
                self._events[slot].record()           # creates the underlying cudaEvent_t
            ev = self._events[slot].cuda_event
        check(lib.vfn_bank_update(banks, alts, obj_n, io, hw, float(frame_idx), float(update_rate),
                                  float(self.thres_close), float(self.class_budget), ptr(ws), ws.numel(),
                                  self._h_pinned[slot].data_ptr(), self.impl, ev, st), 'bank_update')
        self.launches += lib.vfn_launch_count() - l0
        if io[0].deferred:
            self._seq += 1
            self._pending.append((slot, banks, io, dec, hw))
            for c in range(obj_n):
                self._n_hi[c] = int(banks[c].n) + hw
            if not any(s.kh is not None for s in self._slabs):
                self._resolve()       # no device-resident count on this bank: the next call needs the exact size
            return
        err = None
        for c in range(obj_n):
            r = io[c]
            if r.evicted:
                self.last_thresholds = [int(r.thresholds[k]) for k in range(min(r.n_iter, 64))]   # first 64 kept
                self.last_thresholds_obj[c] = list(self.last_thresholds)
                if r.evict_status == 1:
                    err = err or RuntimeError('FeatureBank.remove: every entry was evicted and the budget is still '
                                              'exceeded (the reference raises on LFU.min() of an empty tensor, '
                                              'FeatureBank.py:136)')
                elif r.evict_status == 2:
                    err = err or ValueError('FeatureBank.remove: LFU minimum is not finite (the reference raises in '
                                            'int(), FeatureBank.py:123)')
            if r.swapped:
                self._slabs[c], self._alt[c] = self._alt[c], self._slabs[c]
                self.replace_n[c] += r.n_before - r.kept                         # FeatureBank.py:140-141
            self._n[c] = self._n_hi[c] = int(banks[c].n)
            self._peak_n[c] = max(self._peak_n[c], self._n[c])                       # FeatureBank.py:113
            self._last_decisions[c] = dict(n_merge=int(r.n_merge), n_runs=int(r.n_runs), n_append=int(r.n_append),
                                           evicted=bool(r.evicted), **dec[c])
        if err is not None:
            raise err

    @_on_bank_device
    def _launch_evict_plan(self, c: int, request_n: int, frame_idx):
        self._resolve()
        lib, st = self._lib, stream_ptr()
        s, n = self._slabs[c], self._n[c]
        lfu = self._buf(f'lfu{c}', (max(n, 1),), torch.float32)
        plan = self._buf(f'plan{c}', (72,), torch.int32)
        bank = s.struct(n, self._n_live[c])
        check(lib.vfn_bank_evict_plan(C.byref(bank), float(frame_idx), float(self.class_budget), int(request_n),
                                      ptr(plan), self._h_plan[c].data_ptr(), ptr(lfu), st), 'evict_plan')
        self.launches += 1

    @_on_bank_device
    def _finish_evict(self, c: int, request_n: int):
        lib, st = self._lib, stream_ptr()
        hp = self._h_plan.numpy()[c]
        status, kept, n_iter = int(hp[0]), int(hp[1]), int(hp[2])
        self.last_thresholds = [int(v) for v in hp[4:4 + min(n_iter, 64)]]
        if status == 1:
            raise RuntimeError('FeatureBank.remove: every entry was evicted and the budget is still exceeded '
                               '(the reference raises on LFU.min() of an empty tensor, FeatureBank.py:136)')
        if status == 2:
            raise ValueError('FeatureBank.remove: LFU minimum is not finite (the reference raises in int(), '
                             'FeatureBank.py:123)')
        s, n = self._slabs[c], self._n[c]
        alt = self._alt[c]
        if alt is None or alt.cap < s.cap:
            alt = _Slab(s.d_key, s.d_val, s.cap, self.device, s.kh is not None)
        src, dst = s.struct(n), alt.struct(0)
        lfu = self._scratch[f'lfu{c}']
        plan = self._scratch[f'plan{c}']
        cws = self._buf(f'cws{c}', (lib.vfn_bank_compact_workspace_bytes(n),), torch.uint8)
        check(lib.vfn_bank_compact(C.byref(src), C.byref(dst), ptr(lfu), ptr(plan), ptr(cws), cws.numel(), st),
              'bank_compact')
        # algorithmic bytes of the compaction: evicted rows read once, kept rows read + written
        lib.vfn_profile_add_work(3, 4.0 * (s.d_key + s.d_val + 2) * ((n - kept) + 2.0 * kept))
        self.launches += 3
        self._slabs[c], self._alt[c] = alt, s
        self._n[c] = kept
        self._set_live(c)
        self.replace_n[c] += n - kept                                             # FeatureBank.py:140-141
        return (self.class_budget - kept) - request_n

    @_on_bank_device
    def remove(self, class_idx, request_n, frame_idx):
        """FeatureBank.py:117-143; returns `balance`."""
        self._launch_evict_plan(class_idx, request_n, frame_idx)
        torch.cuda.current_stream().synchronize()
        return self._finish_evict(class_idx, request_n)

    def print_peak_mem(self):
        ur = self.peak_n / self.class_budget
        rr = self.replace_n / self.class_budget
        print(f'Obj num: {self.obj_n}.', f'Budget / obj: {self.class_budget}.', f'UR: {ur}.', f'Replace: {rr}.')

    # ---- test / parity helpers (not in the reference) ---------------------------------------------
    @_on_bank_device
    def load_state(self, keys, values, info):
        """Teacher forcing: overwrite the bank with (d,N) keys/values and (N,2) info tensors."""
        self._resolve()
        for c in range(self.obj_n):
            self._slabs[c], self._alt[c], self._n[c] = None, None, 0
            self._ingest(c, keys[c], values[c], 0.0, 0.0)
            self._slabs[c].info[:self._n[c]].copy_(info[c].to(self.device, torch.float32))
