#!/usr/bin/env python
"""bench.py - hot-path throughput of the AFB-URR memory propagation on B200 (contract: task prompt section 4).

A "step" is one pass of the hot path over one synthetic 480p / 2-object / 100-frame clip at the drop-in boundary:
per frame  read (Matcher.forward) -> URR pre/post -> bank update (match, merge, LFU evict, append), starting from
init_bank.  `value` = frames/s with the clip resident in HBM; `e2e` = the same through the public Python API with
HOST (pinned) inputs copied in and the refined mask copied out every frame.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA kernels)
    python bench.py --impl reference ...                           # reference algorithm on the host cores (oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HW_H, HW_W = 30, 54          # r4 grid of a 480x864 padded frame  (HW = 1620)
R1_H, R1_W = 240, 432
D_KEY, D_VAL = 128, 512
BUDGET = 250000               # test_video_seg.py:24  -> class_budget 100000.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE tc_phase_b_pair_kernel launch, from the committed `ncu --set full`
# capture (profiles/r1m_ncu_summary.md): N = 100000 slots/object, 2 objects, HW = 1620.  The operand arrays that launch
# has to stream once are 2560 B/slot (kh, kl, vh, v8, vl) = 512 MB; its algorithmic work is 3.32e11 flop.
NCU_PHASE_B_TRAFFIC = {'bytes': 611.70e6 + 91.63e6,
                       'note': 'per launch at N=100000 slots/object (ncu capture profiles/r1m_ncu_summary.md); operand '
                               'bytes streamed once = 512 MB; the bench launches average fewer slots'}
TAIL_SIZE = (1080, 1920)      # original frame size the mask is resized back to (test_video_seg.py:103,114)
TAIL_KEY_PTS = [(480, 300), (960, 200), (1440, 400), (1800, 100)]
# --workload: the default is BASELINE.json configs[1] (the metric's configuration); '1080p-capacity' is configs[2]'s regime
# (long-video stress: native 1080p query grid, bank pre-filled to its 100000-slot budget so that LFU eviction fires on
# every frame), offered for measurement only - the driver's line is always the default workload.
WORKLOADS = {
    '480p-2obj-100frame-clip-hotpath': dict(hw=(30, 54), r1=(240, 432), frames=100, n_init=None, start_frame=0),
    '1080p-2obj-bank-at-capacity': dict(hw=(68, 120), r1=(544, 960), frames=30, n_init=100000, start_frame=50),
}
WORKLOAD = '480p-2obj-100frame-clip-hotpath'
START_FRAME, N_INIT = 0, None


def select_workload(name, frames=None):
    global HW_H, HW_W, R1_H, R1_W, WORKLOAD, START_FRAME, N_INIT
    w = WORKLOADS[name]
    (HW_H, HW_W), (R1_H, R1_W) = w['hw'], w['r1']
    WORKLOAD, START_FRAME, N_INIT = name, w['start_frame'], w['n_init']
    return frames or w['frames']


METRIC = '480p frames/sec (1/2/4/8 B200); mem-read tensor util; bank-update HBM GB/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=None)
    ap.add_argument('--workload', default='480p-2obj-100frame-clip-hotpath', choices=list(WORKLOADS))
    ap.add_argument('--frac-merge', type=float, default=0.1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--read-impl', type=int, default=0, help='0 auto (tcgen05), 1 fp32 SIMT, 2 tcgen05')
    args = ap.parse_args()
    args.frames = select_workload(args.workload, args.frames)
    return args


# ---------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------
def make_clip(seed, frames, frac_merge, pin):
    from vfloodnet_b200 import synth
    gen = synth.ClipGenerator(seed=seed, obj_n=2, hw=HW_H * HW_W, d_key=D_KEY, d_val=D_VAL, frac_merge=frac_merge,
                              n_init=N_INIT)
    keys0, vals0 = gen.init()
    fr = []
    for _ in range(frames):
        q_in, q_out, pk, pv = gen.frame()
        fr.append((q_in, q_out, pk, pv))
    g = torch.Generator().manual_seed(seed + 1000)
    urr = synth.gen_urr_inputs(g, 2, R1_H, R1_W)
    if pin:
        P = lambda t: t.pin_memory()
        keys0, vals0 = [P(k) for k in keys0], [P(v) for v in vals0]
        fr = [(P(a), P(b), [P(k) for k in pk], [P(v) for v in pv]) for a, b, pk, pv in fr]
        urr = tuple(P(t) for t in urr)
    clip = dict(keys0=keys0, vals0=vals0, frames=fr, urr=urr)
    if N_INIT:   # a bank that has lived for START_FRAME frames: insertion frames in order, usage counts spread (regime C)
        info0 = []
        for _ in range(2):
            i = synth.gen_info(g, N_INIT, START_FRAME)
            i[:, 0] = torch.sort(i[:, 0]).values
            info0.append(P(i) if pin else i)
        clip['info0'] = info0
    return clip


def to_device(clip, dev):
    D = lambda t: t.to(dev, non_blocking=True)
    out = dict(keys0=[D(k) for k in clip['keys0']], vals0=[D(v) for v in clip['vals0']],
               frames=[(D(a), D(b), [D(k) for k in pk], [D(v) for v in pv]) for a, b, pk, pv in clip['frames']],
               urr=tuple(D(t) for t in clip['urr']))
    if 'info0' in clip:
        out['info0'] = [D(i) for i in clip['info0']]
    return out


def run_clip_gpu(vfn, clip, dev, read_impl, host_inputs=False, out_host=None, exact_sizes=False, tail=None,
                 levels_host=None):
    """one step: the whole clip through the drop-in API.  Returns (bank, last readout, last refined mask).
    host_inputs: every frame's tensors start in pinned HOST memory; their H2D copies are issued on a side stream one
    frame ahead (double buffering) and the refined mask is copied back to the host every frame."""
    fb = vfn.FeatureBank(2, BUDGET, dev, impl=read_impl)
    if exact_sizes:
        fb.defer = False
    m = vfn.Matcher(update_bank=True)
    cur = torch.cuda.current_stream(dev)
    if host_inputs:
        copy_stream = _copy_stream(dev)
        H = lambda t: t.to(dev, non_blocking=True)
        # two resident staging sets on the device (frame t uses set t & 1): no allocation inside the loop - allocating
        # each frame's inputs on the side stream made one clip in ~10 take 30 ms longer (gpurun_out/t12, r1m2: the
        # caching allocator has to cudaMalloc when record_stream'ed blocks are not reusable yet)
        stg = _staging(dev, clip)
        done = [None, None]

        def stage(t):
            b = t & 1
            q_in, q_out, pk, pv = clip['frames'][t]
            src = (q_in, q_out, *pk, *pv, *clip['urr'])
            with torch.cuda.stream(copy_stream):
                if done[b] is not None:
                    copy_stream.wait_event(done[b])          # frame t-2 has consumed this set
                for d_, s_ in zip(stg[b], src):
                    d_.copy_(s_, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            q = stg[b]
            return (q[0], q[1], [q[2], q[3]], [q[4], q[5]], (q[6], q[7], q[8])), ev

        _start_bank(fb, clip, H)
        copy_stream.wait_stream(cur)
        nxt = stage(0)
    out = prob = None
    for t in range(len(clip['frames'])):
        if host_inputs:
            (q_in, q_out, pk, pv, urr), ev = nxt
            if t + 1 < len(clip['frames']):
                nxt = stage(t + 1)
            cur.wait_event(ev)
        else:
            if t == 0:
                _start_bank(fb, clip, lambda x: x)
            q_in, q_out, pk, pv = clip['frames'][t]
            urr = clip['urr']
        p, r1, q_local = urr
        out = m(fb, q_in, q_out)
        p_up, unc, conf, local_match = vfn.urr_pre(p, r1.expand(2, -1, -1, -1), (1, 2, R1_H, R1_W))
        prob = vfn.urr_post(p_up, unc, conf, q_local)
        fb.update(pk, pv, START_FRAME + t + 1)
        if tail is not None:
            # device-resident loop tail: resize to the original frame size, arg-max, largest component,
            # water-level column scan; only the levels go back to the host
            _, levels = tail(prob)
            levels_host.copy_(levels, non_blocking=True)
        elif out_host is not None:
            # the refined mask goes back on its own stream: a 3.3 MB read-back on the compute stream would hold up the
            # bank update of the same frame for ~70 us (the result of frame t is consumed by the host, not by frame t+1)
            d2h = _d2h_stream(dev)
            d2h.wait_stream(cur)
            with torch.cuda.stream(d2h):
                out_host.copy_(prob, non_blocking=True)
            prob.record_stream(d2h)
        if host_inputs:
            done[t & 1] = torch.cuda.Event()
            done[t & 1].record(cur)
    if out_host is not None and tail is None:
        cur.wait_stream(_d2h_stream(dev))          # the step's result is on the host when the step's work is done
    return fb, out, prob


def _start_bank(fb, clip, H):
    if 'info0' in clip:          # pre-filled bank (capacity workload)
        fb.load_state([H(k) for k in clip['keys0']], [H(v) for v in clip['vals0']], [H(i) for i in clip['info0']])
    else:                        # test_video_seg.py:100-101
        fb.init_bank([H(k) for k in clip['keys0']], [H(v) for v in clip['vals0']])


_COPY_STREAMS = {}
_D2H_STREAMS = {}
_STAGING = {}


def _staging(dev, clip):
    q_in, q_out, pk, pv = clip['frames'][0]
    src = (q_in, q_out, *pk, *pv, *clip['urr'])
    key = (str(dev), tuple(tuple(t.shape) for t in src))
    if key not in _STAGING:
        assert len(src) == 9, 'two objects: q_in, q_out, 2 keys, 2 values, 3 URR tensors'
        _STAGING[key] = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in src] for _ in range(2)]
    return _STAGING[key]


def _d2h_stream(dev):
    if dev not in _D2H_STREAMS:
        _D2H_STREAMS[dev] = torch.cuda.Stream(dev)
    return _D2H_STREAMS[dev]


def _copy_stream(dev):
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(dev)
    return _COPY_STREAMS[dev]


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region.  Samples through NVML in-process (the library behind
    nvidia-smi; no fork of a CUDA process inside the timed region), falling back to the nvidia-smi CLI."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    BITS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def mark(self):
        """samples from here on belong to the timed region"""
        self.first = len(self.samples)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    try:
                        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    try:
                        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                    except Exception:
                        pw = None
                    self.samples.append([str(sm), str(self.sm_max)] +
                                        [('Active' if r & b else 'Not Active') for b in self.BITS.values()] + [pw])
                else:
                    o = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                        str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.samples.append([x.strip() for x in o.split(',')])
            except Exception:
                pass
            # sparse on purpose: NVML queries contend with kernel launches for a driver lock (profiles/r1h: a 20 Hz
            # sampler on rank 0 tripled the step time of a 2-GPU run; nvidia-smi forks cost 25 % at 10 Hz on one GPU)
            time.sleep(0.3 if self.nvml is not None else 0.6)

    def summary(self):
        samples = self.samples[getattr(self, 'first', 0):] or self.samples[-1:]
        if not samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock query unavailable']}
        sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
        names = list(self.BITS)
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in samples)]
        pw = sorted(s[6] for s in samples if len(s) > 6 and s[6] is not None)
        out = {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': int(samples[0][1]), 'reasons': reasons,
               'samples': len(samples), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}
        if pw:
            out['power_w'] = round(pw[len(pw) // 2], 1)      # the tensor kernels run the board at its power limit
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port, torch CPU, all host threads) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def expected_bank_size(frame, frac_merge):
    hw = HW_H * HW_W
    if N_INIT:
        return int(min(N_INIT, 0.8 * (BUDGET // 2)))
    return int(min(hw + (1 - frac_merge) * hw * frame, 0.8 * (BUDGET // 2)))


def default_samples():
    return (START_FRAME + 10,) if N_INIT else (25, 50, 75, 100)


def cpu_sample(frac_merge, sample_frames=None, seed=0, device='cpu'):
    """Times oracle read + URR + update at the bank sizes the clip has at `sample_frames` (synthetic bank contents of
    the clip's analytic size trajectory).  Returns (frames_per_sec, description, seconds).
    device='cuda:N' runs the same plain-torch restatement with ATen/cuBLAS kernels on the GPU ("reference torch ops" of
    BASELINE.json configs[1]); timed with a device synchronisation on both sides of every step."""
    from oracle import afb_oracle as O
    from vfloodnet_b200 import synth
    sample_frames = sample_frames or default_samples()
    on_gpu = str(device) != 'cpu'
    if not on_gpu:
        torch.set_num_threads(os.cpu_count())
    hw = HW_H * HW_W
    total = 0.0
    D = (lambda t: t.to(device)) if on_gpu else (lambda t: t)
    for f in sample_frames:
        g = torch.Generator().manual_seed(seed + f)
        n = expected_bank_size(f, frac_merge)
        fb = O.OracleFeatureBank(2, BUDGET, device)
        keys, vals = zip(*[synth.gen_bank(g, n) for _ in range(2)])
        fb.init_bank([D(k) for k in keys], [D(v) for v in vals])
        for c in range(2):
            fb.info[c] = D(synth.gen_info(g, n, f))
        q_in, q_out = synth.gen_query(g, hw)
        pk, pv = zip(*[synth.gen_candidates(g, keys[c], vals[c], hw, frac_merge) for c in range(2)])
        p, r1, q_local = synth.gen_urr_inputs(g, 2, R1_H, R1_W)
        args = (D(q_in), D(q_out), [D(k) for k in pk], [D(v) for v in pv])
        urr = (D(p), D(r1).expand(2, -1, -1, -1), D(q_local), (1, 2, R1_H, R1_W))
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.hot_path_step(fb, *args, f, urr_in=urr)
        if on_gpu:
            torch.cuda.synchronize()
        total += time.perf_counter() - t0
    where = 'torch CUDA ops (ATen/cuBLAS fp32) on the same GPU' if on_gpu else 'torch CPU fp32'
    desc = (f'oracle port ({where}) of read+URR+update on frames {list(sample_frames)} of the clip, bank sizes '
            f'{[expected_bank_size(f, frac_merge) for f in sample_frames]} slots/object (analytic trajectory)')
    return len(sample_frames) / total, desc, total


def main_reference(args, rank, world):
    if rank != 0:
        return
    for _ in range(args.warmup and 1):
        cpu_sample(args.frac_merge, sample_frames=(5,))
    t0 = time.perf_counter()
    n_frames = 0
    for _ in range(args.steps):
        fps, desc, secs = cpu_sample(args.frac_merge)
        n_frames += len(default_samples())
    el = time.perf_counter() - t0
    v = n_frames / el
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'hw': HW_H * HW_W, 'budget': BUDGET,
                       'frames': args.frames, 'frac_merge': args.frac_merge},
            'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': desc},
            'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main_ours(args, rank, world, local_rank):
    import ctypes
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # one independent stream (clip) per GPU: weak scaling, no data-path collective (stream-parallel)
    host_clip = make_clip(seed=100 + rank, frames=args.frames, frac_merge=args.frac_merge, pin=True)
    dev_clip = to_device(host_clip, dev)
    out_host = torch.empty((2, 2 * R1_H, 2 * R1_W), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up: the first NVML queries of a process are slow and hold a driver lock
    # that kernel launches also take (gpurun_out/q1: a sampler started at the timed region doubled its step time)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get('VFN_BENCH_NO_SAMPLER'):
        sampler.start()
    for _ in range(args.warmup):
        run_clip_gpu(vfn, dev_clip, dev, args.read_impl)
    barrier()
    sampler.mark()
    # timed region: K clips, no per-kernel events (bracketing every kernel with timing events serialises the stream:
    # profiles/r1h measured 1.48 ms/frame with them against 1.16 ms without)
    l0 = lib.vfn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    step_ev = []
    fb = out = prob = None
    for _ in range(args.steps):
        # one bank per stream: the previous clip's bank is released before the next clip builds its own, exactly as in
        # the warm-up (gpurun_out/host1: with the old bank still referenced the second timed clip had to cudaMalloc a
        # second set of slabs inside the timed region: one step of 780 ms among steps of 105 ms)
        fb = out = prob = None
        fb, out, prob = run_clip_gpu(vfn, dev_clip, dev, args.read_impl)
        step_ev.append(torch.cuda.Event(enable_timing=True))
        step_ev[-1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ms_steps = [a.elapsed_time(b) for a, b in zip([e0] + step_ev[:-1], step_ev)]
    launches = lib.vfn_launch_count() - l0
    final_n = [fb.bank_n(c) for c in range(2)]
    # roofline pass: the same K clips again with the library's CUDA events around each dominant kernel (recorded on the
    # launching stream), bank sizes read back every frame so that the algorithmic work per launch is exact
    sampler.stop_flag = True                      # the clocks line covers the timed region only
    lib.vfn_profile_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        run_clip_gpu(vfn, dev_clip, dev, args.read_impl, exact_sizes=True)
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    prof = (ctypes.c_double * 24)()
    _lib.check(lib.vfn_profile_collect(prof, 8), 'profile_collect')
    lib.vfn_profile_enable(0)

    # e2e: host inputs, copies inside the timed region
    for _ in range(2):     # two untimed clips: the side-stream allocations of the host-input path settle (gpurun_out/t12)
        run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, out_host=out_host)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, out_host=out_host)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    # e2e with the device-resident tail: per frame the refined probabilities are resized to 1080x1920, reduced to the
    # largest water component and scanned for the water level at 4 key points; 16 bytes per frame return to the host
    # instead of the 3.3 MB probability map (the reference copies a full-resolution mask and runs OpenCV on the host)
    from vfloodnet_b200 import tail as vtail
    ft = vtail.FrameTail(TAIL_SIZE, TAIL_KEY_PTS, dev)
    levels_host = torch.empty(len(TAIL_KEY_PTS), dtype=torch.float32).pin_memory()
    for _ in range(2):
        run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, tail=ft, levels_host=levels_host)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(args.steps):
        run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, tail=ft, levels_host=levels_host)
    u1.record()
    barrier()
    ms_e2e_tail = u0.elapsed_time(u1)
    # the tail alone, back to back on one stream (working set ~25 MB: L2 resident, stated in the line), on a smooth
    # soft mask (a handful of water bodies, what a trained model emits) and on per-pixel noise (10^5 components: the
    # labelling's worst case, and what the URR output of this benchmark's random inputs looks like)
    gt = torch.Generator().manual_seed(5)
    coarse = torch.randn(1, 2, 10, 16, generator=gt).to(dev) * 4
    smooth = torch.softmax(torch.nn.functional.interpolate(coarse, size=(2 * R1_H, 2 * R1_W), mode='bicubic',
                                                           align_corners=False), dim=1)[0].contiguous()
    noise = torch.rand((2, 2 * R1_H, 2 * R1_W), generator=gt).to(dev)
    tail_ms = {}
    tail_launches = 0
    for nm, src in (('smooth', smooth), ('noise', noise)):
        for _ in range(5):
            ft(src)
        tl0 = lib.vfn_launch_count()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(50):
            ft(src)
        v1.record()
        torch.cuda.synchronize()
        tail_ms[nm] = v0.elapsed_time(v1) / 50
        tail_launches = (lib.vfn_launch_count() - tl0) // 50
    tail_stats = ft.stats.tolist()

    t_ms = torch.tensor([ms, ms_e2e, ms_e2e_tail], dtype=torch.float64, device=dev)
    per_rank = [[ms, ms_e2e, ms_e2e_tail]]
    if dist is not None:
        allt = [torch.empty_like(t_ms) for _ in range(world)]
        dist.all_gather(allt, t_ms)
        per_rank = [t.tolist() for t in allt]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_tail = t_ms.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    frames_total = args.frames * args.steps * world
    value = frames_total / (ms / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)
    numel = lambda ts: sum(t.numel() for t in ts)
    fr = host_clip['frames'][0]
    h2d_frame = 4 * (fr[0].numel() + fr[1].numel() + numel(fr[2]) + numel(fr[3]) + numel(host_clip['urr']))
    h2d = args.frames * h2d_frame + 4 * (numel(host_clip['keys0']) + numel(host_clip['vals0']))
    d2h = args.frames * out_host.numel() * 4

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'bf16_tflops_sustained of MEASURED_PEAKS.json' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    k = lambda i: (prof[3 * i], prof[3 * i + 1], prof[3 * i + 2])
    nb, msb, wb = k(1)
    na, msa, wa = k(0)
    ach = (wb / (msb * 1e-3) / 1e12) if msb > 0 else 0.0
    roofline = {'bound': 'tensor', 'kernel': 'tc_phase_b_kernel (read phase B: P=softmax, O+=P.V, usage counts)',
                'achieved': ach, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach / tf_peak if tf_peak else None,
                'traffic': NCU_PHASE_B_TRAFFIC['bytes'], 'traffic_note': NCU_PHASE_B_TRAFFIC['note'],
                'peak_source': peak_src, 'launches': int(nb), 'avg_ms': msb / nb if nb else None,
                'timing': 'CUDA events around each launch, in a second pass over the same K clips (ms_per_step_profiled); '
                          'the timed region itself carries no per-kernel events',
                'ms_per_step_profiled': ms_prof / args.steps,
                'algorithmic_flop_per_launch': wb / nb if nb else None,
                'read_total': {'achieved': ((wa + wb) / ((msa + msb) * 1e-3) / 1e12) if msa + msb > 0 else 0.0,
                               'phase_a_avg_ms': msa / na if na else None, 'unit': 'TFLOP/s'}}
    extra = {}
    names = {2: 'match', 3: 'compact_move', 4: 'merge', 5: 'append', 6: 'urr_local'}
    for i, nm in names.items():
        n_i, ms_i, w_i = k(i)
        if n_i:
            d = {'launches': int(n_i), 'avg_ms': ms_i / n_i}
            if w_i > 0:
                rate = w_i / (ms_i * 1e-3)
                if i == 2:
                    d.update(achieved=rate / 1e12, unit='TFLOP/s')
                else:
                    d.update(achieved=rate / 1e9, unit='GB/s', frac_of_hbm_peak=rate / 1e9 / hbm_peak)
            extra[nm] = d
    tail_bytes = 4.0 * smooth.numel() + TAIL_SIZE[0] * TAIL_SIZE[1]        # read the soft mask once, write the u8 mask
    extra['frame_tail'] = {'launches_per_frame': int(tail_launches), 'avg_ms': tail_ms['smooth'],
                           'avg_ms_noise_input': tail_ms['noise'], 'out_size': list(TAIL_SIZE),
                           'achieved': tail_bytes / (tail_ms['smooth'] * 1e-3) / 1e9, 'unit': 'GB/s',
                           'frac_of_hbm_peak': tail_bytes / (tail_ms['smooth'] * 1e-3) / 1e9 / hbm_peak,
                           'noise_input_stats': dict(zip(['fg_pixels', 'components', 'kept', 'root'], tail_stats)),
                           'note': 'resize+argmax, largest 8-connected component, water-level scan; working set is L2 '
                                   'resident (25 MB), latency bound: 7 dependent launches'}
    line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16 hi/lo + f8 split operands, f32 accumulate (read, match); f32 (update, URR)',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'hw': HW_H * HW_W, 'budget': BUDGET,
                       'frames': args.frames, 'frac_merge': args.frac_merge, 'streams_per_gpu': 1,
                       'final_bank_slots': final_n, 'l2_policy': 'inputs_exceed_l2 (bank operands 0.5-1.1 GB >> 126 MB)',
                       'read_impl': args.read_impl},
            'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'e2e_tail': {'value': frames_total / (ms_e2e_tail / 1e3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': args.frames * 4 * len(TAIL_KEY_PTS),
                         'note': 'e2e + device-resident loop tail (1080x1920 mask, largest component, 4 water levels)'},
            'gpu_launches': int(launches), 'roofline': roofline, 'kernels': extra, 'clocks': sampler.summary(),
            'ms_per_rank': [[round(x / args.steps, 2) for x in r] for r in per_rank],
            'ms_steps': [round(x, 2) for x in ms_steps]}
    if world == 1:
        # BASELINE.json configs[1]: "... vs reference torch ops" - the plain-torch restatement on this GPU (baseline only)
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            cpu_sample(args.frac_merge, sample_frames=(5,), device=dev)                   # cuBLAS / allocator warm-up
            fps_t, desc_t, secs_t = cpu_sample(args.frac_merge, device=dev)
            line['torch_gpu_baseline'] = {'value': fps_t, 'unit': 'frames/s', 'kind': 'port', 'sample': desc_t,
                                          'seconds': secs_t}
        except Exception as e:                                                            # reported, never hidden
            line['torch_gpu_baseline'] = {'value': None, 'error': f'{type(e).__name__}: {e}'[:300]}
        torch.cuda.empty_cache()
    if not args.no_cpu_baseline and world == 1:
        fps, desc, secs = cpu_sample(args.frac_merge)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': desc, 'seconds': secs}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        main_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun when asked for N>1 GPUs directly
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
                   '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        main_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
