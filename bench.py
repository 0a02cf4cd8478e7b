#!/usr/bin/env python
"""bench.py - hot-path throughput of the AFB-URR memory propagation on B200 (contract: task prompt section 4).

A "step" is one pass of the hot path over one synthetic 480p / 2-object / 100-frame clip at the drop-in boundary:
per frame  read (Matcher.forward) -> URR pre/post -> bank update (match, merge, LFU evict, append), starting from
init_bank.  `value` = frames/s with the clip resident in HBM; `e2e` = the same through the public Python API with
HOST (pinned) inputs copied in and the refined mask copied out every frame.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA kernels)
    python bench.py --impl reference ...                           # the reference's own read+update on the host cores:
                                                                   # the SAME clip, free-running (true bank trajectory)
    python bench.py --workload 480p-model-clip                     # whole reference model, unpatched vs patch_model
    python bench.py --workload 1080p-2obj-bank-at-capacity --frames 2000      # BASELINE configs[2]
    python bench.py --workload 480p-64-streams --gpus G                         # BASELINE configs[3]
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
# capture (profiles/r3d_ncu_summary.md): N = 100000 slots/object, 2 objects, HW = 1620.  The operand arrays that launch
# has to stream once are 2560 B/slot (kh, kl, vh, v8, vl) = 512 MB; its algorithmic work is 3.32e11 flop.
NCU_PHASE_B_TRAFFIC = {'bytes': 612.65e6 + 92.52e6,
                       'note': 'per launch at N=100000 slots/object (ncu capture profiles/r3d_ncu_summary.md); operand '
                               'bytes streamed once = 512 MB; the bench launches average fewer slots'}
TAIL_SIZE = (1080, 1920)      # original frame size the mask is resized back to (test_video_seg.py:103,114)
TAIL_KEY_PTS = [(480, 300), (960, 200), (1440, 400), (1800, 100)]
# --workload: the default is BASELINE.json configs[1] (the metric's configuration); '1080p-capacity' is configs[2]'s regime
# (long-video stress: native 1080p query grid, bank pre-filled to its 100000-slot budget so that LFU eviction fires on
# every frame), offered for measurement only - the driver's line is always the default workload.
WORKLOADS = {
    '480p-2obj-100frame-clip-hotpath': dict(hw=(30, 54), r1=(240, 432), frames=100, n_init=None, start_frame=0),
    '1080p-2obj-bank-at-capacity': dict(hw=(68, 120), r1=(544, 960), frames=30, n_init=100000, start_frame=50),
    # whole reference model (encoders, KeyValue, decoder convolutions stay cuDNN) around the hot path: the unmodified
    # reference on the GPU against the same weights with vfloodnet_b200.patch_model (BASELINE configs[1])
    '480p-model-clip': dict(hw=(30, 54), r1=(240, 432), frames=100, n_init=None, start_frame=0),
    # BASELINE configs[3]: 64 independent 480p water-level streams, 64/G per GPU, several of them in flight per GPU on
    # their own CUDA streams; the step is one pass over all of a rank's streams (read -> URR -> update -> frame tail)
    '480p-64-streams': dict(hw=(30, 54), r1=(240, 432), frames=100, n_init=None, start_frame=0),
    # BASELINE configs[4]: one 4K stream (HW = 32400), its bank sharded over the GPUs of the node, split-memory read with
    # the log-sum-exp combine and the readout reduction as kernels over peer memory (tests/multi_gpu_sharded_bench.py)
    '4k-2obj-sharded-bank': dict(hw=(135, 240), r1=(1080, 1920), frames=24, n_init=None, start_frame=0),
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
    ap.add_argument('--no-torch-baseline', action='store_true', help='skip the reference-torch-ops-on-this-GPU leg')
    ap.add_argument('--no-affinity', action='store_true', help='do not bind the process to the GPU-local CPUs')
    ap.add_argument('--total-streams', type=int, default=64, help='480p-64-streams: video streams over all GPUs')
    ap.add_argument('--concurrency', type=int, default=8, help='480p-64-streams: streams in flight per GPU')
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
                 levels_host=None, before_frame=None, copy_only=False):
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
        if copy_only:          # the host->device leg alone (same staging, same events), no kernels
            if host_inputs:
                done[t & 1] = torch.cuda.Event()
                done[t & 1].record(cur)
            continue
        if before_frame is not None:
            before_frame(t, fb, (q_in, q_out, pk, pv))
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
# Reference arm: the reference's own read + update (unmodified classes from baseline/_ref when staged, else the oracle
# port) on plain torch ops, over THE SAME CLIP as our arm (same ClipGenerator seed), free-running from init_bank so that
# every frame sees the true bank state (sizes, usage counts, evictions) of the clip.
# ---------------------------------------------------------------------------------------------------
def config_dict(args):
    """identical in both arms (the driver compares it)"""
    return {'workload': WORKLOAD, 'hw': HW_H * HW_W, 'budget': BUDGET, 'frames': args.frames,
            'frac_merge': args.frac_merge,
            'l2_policy': 'inputs_exceed_l2 (bank operands 0.5-1.1 GB per object >> 126 MB L2)'}


def _ref_clip_source(seed, frames, frac_merge):
    """frame generator of the clip (CPU tensors, produced one frame ahead of use: a whole clip is 4 GB)"""
    from vfloodnet_b200 import synth
    gen = synth.ClipGenerator(seed=seed, obj_n=2, hw=HW_H * HW_W, d_key=D_KEY, d_val=D_VAL, frac_merge=frac_merge,
                              n_init=N_INIT)
    keys0, vals0 = gen.init()
    g = torch.Generator().manual_seed(seed + 1000)
    urr = synth.gen_urr_inputs(g, 2, R1_H, R1_W)
    return gen, keys0, vals0, urr


def reference_free_run(frac_merge, frames, seed, device='cpu', steps=1, timed_from=0, max_seconds=None):
    """Runs the clip free from init_bank through the reference arm.  Every frame is timed on its own; frame t belongs to
    step (t mod steps) - each step is a bounded sample (every steps-th frame) of the one clip, and together the steps
    are the whole clip.  Returns dict(kind, seconds_per_step[steps], frames_per_step[steps], sizes, desc)."""
    from baseline.ref_arm import RefArm
    on_gpu = str(device) != 'cpu'
    if not on_gpu:
        torch.set_num_threads(os.cpu_count())
    else:
        torch.backends.cuda.matmul.allow_tf32 = False      # the reference's default: true fp32 GEMMs
    gen, keys0, vals0, urr = _ref_clip_source(seed, frames, frac_merge)
    arm = RefArm(BUDGET, device)
    arm.init(keys0, vals0)
    if N_INIT:
        from vfloodnet_b200 import synth
        g = torch.Generator().manual_seed(seed + 1000)
        synth.gen_urr_inputs(g, 2, R1_H, R1_W)
        info0 = []
        for _ in range(2):
            i = synth.gen_info(g, N_INIT, START_FRAME)
            i[:, 0] = torch.sort(i[:, 0]).values
            info0.append(i)
        arm.load(keys0, vals0, info0)
    p, r1, q_local = [arm.D(t) for t in urr]
    urr_in = (p, r1.expand(2, -1, -1, -1), q_local, (1, 2, R1_H, R1_W))
    secs, cnt, sizes = [0.0] * steps, [0] * steps, []
    t_all = time.perf_counter()
    done = 0
    for t in range(frames):
        q_in, q_out, pk, pv = gen.frame()
        a = (arm.D(q_in), arm.D(q_out), [arm.D(k) for k in pk], [arm.D(v) for v in pv])
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        arm.frame(*a, urr_in, START_FRAME + t + 1)
        if on_gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sizes.append(arm.sizes())
        done += 1
        if t >= timed_from:
            secs[t % steps] += dt
            cnt[t % steps] += 1
        if max_seconds is not None and time.perf_counter() - t_all > max_seconds:
            break
    where = 'torch CUDA ops (ATen/cuBLAS fp32, allow_tf32=False) on the same GPU' if on_gpu else 'torch CPU fp32, all host threads'
    what = ('unmodified reference FeatureBank + Matcher (baseline/_ref), URR block via the oracle restatement'
            if arm.kind == 'reference' else 'oracle port of the reference (baseline/_ref not staged)')
    desc = (f'{what}; {where}; read+URR+update of frames 1..{done} of the same clip (seed {seed}), free-running from '
            f'init_bank: true bank trajectory, mean {sum(sum(x) for x in sizes) / (2 * len(sizes)):.0f} slots/object, '
            f'final {sizes[-1]}')
    return dict(kind=arm.kind, secs=secs, cnt=cnt, sizes=sizes, desc=desc, frames_done=done, arm=arm)


def main_reference(args, rank, world):
    if rank != 0:
        return
    if WORKLOAD == '480p-model-clip':
        return main_reference_model(args)
    if args.warmup:                                   # one short untimed pass (allocator, thread pools, MKL plans)
        reference_free_run(args.frac_merge, min(3, args.frames), seed=99, steps=1)
    r = reference_free_run(args.frac_merge, args.frames, seed=100, steps=args.steps)
    el = sum(r['secs'])
    n_frames = sum(r['cnt'])
    v = n_frames / el
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_dict(args),
            'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': r['kind'],
                             'sample': r['desc'] + f'; step k = frames t with t mod {args.steps} == k '
                                                   f'({n_frames} frames in all)'},
            'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'final_bank_slots': r['sizes'][-1]}
    print(json.dumps(line))


def cpu_sample_true_state(vfn, dev_clip, dev, read_impl, frac_merge, sample_frames):
    """cpu_baseline of OUR arm: a bounded sample of the same clip.  The GPU run is paused before the read of each sample
    frame, its bank (the true state of the clip at that frame) is copied to the host, and the reference arm (CPU) is
    timed on that one frame.  sample_frames are 0-based clip positions."""
    from baseline.ref_arm import RefArm
    torch.set_num_threads(os.cpu_count())
    secs, sizes, kind = [], [], [None]
    p, r1, q_local = [t.cpu() for t in dev_clip['urr']]
    urr_in = (p, r1.expand(2, -1, -1, -1), q_local, (1, 2, R1_H, R1_W))

    def before_frame(t, fb, frame):
        if t not in sample_frames:
            return
        torch.cuda.synchronize()
        arm = RefArm(BUDGET, 'cpu')
        arm.load([fb.keys[c].cpu() for c in range(2)], [fb.values[c].cpu() for c in range(2)],
                 [fb.info[c].cpu() for c in range(2)])
        kind[0] = arm.kind
        q_in, q_out, pk, pv = frame
        a = (q_in.cpu(), q_out.cpu(), [k.cpu() for k in pk], [v.cpu() for v in pv])
        sizes.append(arm.sizes())
        t0 = time.perf_counter()
        arm.frame(*a, urr_in, START_FRAME + t + 1)
        secs.append(time.perf_counter() - t0)

    run_clip_gpu(vfn, dev_clip, dev, read_impl, before_frame=before_frame)
    total = sum(secs)
    desc = (f'reference arm ({kind[0]}; torch CPU fp32, all host threads) on frames {[t + 1 for t in sample_frames]} of the '
            f'same clip, each started from the TRUE bank state of the clip at that frame (copied from the GPU run), bank '
            f'sizes {sizes} slots/object')
    return len(secs) / total, desc, total, kind[0]


# ---------------------------------------------------------------------------------------------------
# host placement: bind the process (and so its pinned allocations: first touch) to the CPUs / NUMA node of its GPU
# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_node(index):
    """at 8 GPUs the per-frame H2D copies of ranks 0-3 ran at 2/3 of the speed of ranks 4-7 - pinned buffers
    had been allocated wherever the launcher happened to start the process.  Must run BEFORE any pin_memory().
    Returns a description for the bench line."""
    info = {'bound': False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        try:
            info['pci'] = pynvml.nvmlDeviceGetPciInfo(h).busId
            if isinstance(info['pci'], bytes):
                info['pci'] = info['pci'].decode()
        except Exception:
            pass
        node = None
        if info.get('pci'):
            pth = f"/sys/bus/pci/devices/{info['pci'].lower()[-12:]}/numa_node"
            if os.path.exists(pth):
                node = int(open(pth).read().strip())
        info['numa_node'] = node
        cpus = None
        if node is not None and node >= 0 and os.path.exists(f'/sys/devices/system/node/node{node}/cpulist'):
            cpus = set()
            for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        else:
            try:                                                    # NVML's ideal CPU set of the device
                n_words = (os.cpu_count() + 63) // 64
                mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
                cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
            except Exception:
                cpus = None
        allowed = os.sched_getaffinity(0)
        if cpus:
            cpus = cpus & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
        else:
            info['cpus'] = len(allowed)
    except Exception as e:
        info['error'] = f'{type(e).__name__}: {e}'[:120]
    return info


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
    placement = {'bound': False, 'note': 'disabled'} if args.no_affinity else bind_to_gpu_node(local_rank)
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
    # the host->device leg of that loop alone (every rank at once): what the host memory / PCIe path sustains
    run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, copy_only=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        run_clip_gpu(vfn, host_clip, dev, args.read_impl, host_inputs=True, copy_only=True)
    c1.record()
    barrier()
    ms_copy = c0.elapsed_time(c1)

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

    t_ms = torch.tensor([ms, ms_e2e, ms_e2e_tail, ms_copy], dtype=torch.float64, device=dev)
    per_rank = [[ms, ms_e2e, ms_e2e_tail, ms_copy]]
    if dist is not None:
        allt = [torch.empty_like(t_ms) for _ in range(world)]
        dist.all_gather(allt, t_ms)
        per_rank = [t.tolist() for t in allt]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_tail, ms_copy = t_ms.tolist()
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
                # executed MMA work per algorithmic unit of this kernel: per 128 queries x 64 slots x 256 channels it issues
                # 3 S passes (24 x 32 clk) + f16 and two f8 readout products (8 x 128 clk) = 1792 MMA-clk for 512 algorithmic
                # (O only; S is credited to phase A).  Reported next to `frac`, never instead of it.
                'executed_mma': {'ratio_to_algorithmic': 1792.0 / 512.0, 'achieved': ach * 1792.0 / 512.0, 'unit': 'TFLOP/s bf16-equivalent',
                                 'frac': (ach * 1792.0 / 512.0) / tf_peak if tf_peak else None},
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
    copy_gbs = h2d * args.steps / (ms_copy * 1e-3) / 1e9
    line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16 hi/lo + f8 split operands, f32 accumulate (read, match); f32 (update, URR)',
            'data': 'synthetic', 'config': config_dict(args),
            'run': {'streams_per_gpu': 1, 'final_bank_slots': final_n, 'read_impl': args.read_impl, 'seed': 100,
                    'host_placement': placement},
            'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'e2e_tail': {'value': frames_total / (ms_e2e_tail / 1e3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': args.frames * 4 * len(TAIL_KEY_PTS),
                         'note': 'e2e + device-resident loop tail (1080x1920 mask, largest component, 4 water levels)'},
            'h2d_copy_only': {'gb_per_s_per_gpu': copy_gbs, 'ms_per_step': ms_copy / args.steps,
                              'frames_per_s_bound': frames_total / (ms_copy / 1e3),
                              'note': 'the e2e loop with its kernels removed: pinned-host -> device copies of every '
                                      "frame's boundary tensors on all ranks at once (max over ranks); e2e cannot exceed it"},
            'gpu_launches': int(launches), 'roofline': roofline, 'kernels': extra, 'clocks': sampler.summary(),
            'ms_per_rank': [[round(x / args.steps, 2) for x in r] for r in per_rank],
            'ms_steps': [round(x, 2) for x in ms_steps]}
    if world == 1 and not args.no_torch_baseline:
        # BASELINE.json configs[1]: "... vs reference torch ops": the reference's own classes on this GPU over the same clip
        try:
            reference_free_run(args.frac_merge, 3, seed=99, device=dev)                   # cuBLAS / allocator warm-up
            r = reference_free_run(args.frac_merge, args.frames, seed=100, device=dev)
            line['torch_gpu_baseline'] = {'value': sum(r['cnt']) / sum(r['secs']), 'unit': 'frames/s', 'kind': r['kind'],
                                          'sample': r['desc'], 'seconds': sum(r['secs']),
                                          'final_bank_slots': r['sizes'][-1]}
            del r
        except Exception as e:                                                            # reported, never hidden
            line['torch_gpu_baseline'] = {'value': None, 'error': f'{type(e).__name__}: {e}'[:300]}
        torch.cuda.empty_cache()
    if not args.no_cpu_baseline and world == 1:
        n_f = args.frames
        sample = sorted({min(n_f - 1, int((i + 0.5) * n_f / 8)) for i in range(8)})      # midpoints of 8 equal parts
        fps, desc, secs, kind = cpu_sample_true_state(vfn, dev_clip, dev, args.read_impl, args.frac_merge, sample)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': kind,
                                'sample': desc, 'seconds': secs}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# --workload 480p-model-clip: the whole reference model around the hot path
# ---------------------------------------------------------------------------------------------------
def _stage_table(tot, frames):
    """per-frame ms: CNN = segment + memorize minus the stages timed inside them"""
    seg, mem, upd = tot.get('segment', 0.0), tot.get('memorize', 0.0), tot.get('update', 0.0)
    read, urr = tot.get('read', 0.0), tot.get('urr', 0.0)
    d = {'cnn': (seg - read - urr + mem) / frames, 'read': read / frames, 'update': upd / frames,
         'total': (seg + mem + upd) / frames}
    if 'urr' in tot:
        d['urr'] = urr / frames
    else:
        d['urr'] = None          # inlined in Decoder.forward between convolutions in the unpatched reference: inside cnn
    return {k: (round(v, 4) if v is not None else None) for k, v in d.items()}


def main_model_clip(args, rank, world, local_rank):
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import _lib
    from baseline import refshim, model_clip as MC
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    placement = {'bound': False, 'note': 'disabled'} if args.no_affinity else bind_to_gpu_node(local_rank)
    if not refshim.available():
        print(json.dumps({'metric': METRIC, 'config': config_dict(args),
                          'unavailable': 'reference model definition not staged (python baseline/make_ref.py)'}))
        return
    ns = refshim.load()
    model_ref = MC.build_reference_model(ns, dev)
    model_ours = MC.patched_copy(model_ref, vfn)
    host_clip = MC.make_clip(args.frames, seed=rank, pin=True)
    dev_clip = [f.to(dev) for f in host_clip]
    torch.cuda.synchronize()

    def clip_ours(frames, **kw):
        return MC.run_clip(model_ours, vfn.FeatureBank, frames, dev, budget=BUDGET, **kw)

    def clip_ref(frames, **kw):
        return MC.run_clip(model_ref, ns.FeatureBank, frames, dev, budget=BUDGET, **kw)

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get('VFN_BENCH_NO_SAMPLER'):
        sampler.start()
    for _ in range(max(args.warmup, 1)):
        clip_ours(dev_clip, keep_masks=False)
    torch.cuda.synchronize()
    sampler.mark()
    l0 = lib.vfn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = None
    for _ in range(args.steps):
        o = None          # one bank at a time: a second live bank makes the caching allocator cudaMalloc inside the region
        o = clip_ours(dev_clip, keep_masks=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.vfn_launch_count() - l0
    final_n = [o['fb'].bank_n(c) for c in range(2)]
    o = None
    sampler.stop_flag = True
    # e2e: every frame starts in pinned host memory, the arg-max mask returns to the host
    for _ in range(2):
        clip_ours(host_clip, keep_masks=False, frames_on_host=True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        clip_ours(host_clip, keep_masks=False, frames_on_host=True)
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1)
    # per-stage times, our arm: events around segment / memorize / update, hooks around the read; the URR kernels are timed
    # on their own with tensors of the model's shapes (they sit between convolutions inside Decoder.forward)
    torch.cuda.empty_cache()
    stage_wall = []
    for _pass in range(2):          # the second pass is reported (the first one re-warms the allocator after the e2e legs)
        tm = MC.StageTimer(dev)
        hooks = MC.instrument(model_ours, tm)
        ours_run = None
        try:
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            ours_run = clip_ours(dev_clip, timer=tm)
            torch.cuda.synchronize()
            stage_wall.append(round((time.perf_counter() - w0) * 1e3 / args.frames, 3))
        finally:
            for h in hooks:
                h.remove()
    tot_ours = tm.totals()
    from vfloodnet_b200 import synth
    g = torch.Generator(device=dev).manual_seed(9)
    up, ur1, uq = synth.gen_urr_inputs(g, 2, R1_H, R1_W)
    for _ in range(3):
        a_ = vfn.urr_pre(up, ur1.expand(2, -1, -1, -1), (1, 2, R1_H, R1_W))
        vfn.urr_post(a_[0], a_[1], a_[2], uq)
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(20):
        a_ = vfn.urr_pre(up, ur1.expand(2, -1, -1, -1), (1, 2, R1_H, R1_W))
        vfn.urr_post(a_[0], a_[1], a_[2], uq)
    u1.record()
    torch.cuda.synchronize()
    tot_ours['urr'] = u0.elapsed_time(u1) / 20 * args.frames
    del a_, up, ur1, uq
    st_ours = _stage_table(tot_ours, args.frames)
    # the same patched model with its convolution stages replayed as CUDA graphs (vfloodnet_b200.GraphedAFBURR, n4)
    graphed = None
    try:
        gm = vfn.GraphedAFBURR(model_ours, tuple(dev_clip[0].shape))
        run_g = lambda frames, **kw: MC.run_clip(gm, vfn.FeatureBank, frames, dev, budget=BUDGET, **kw)
        g0, g1, g2, g3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        gr = None
        for _ in range(2):          # every leg is preceded by untimed passes of its own kind (allocator, pinned buffers)
            run_g(dev_clip, keep_masks=False)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(args.steps):
            gr = None
            gr = run_g(dev_clip, keep_masks=False)
        g1.record()
        gr_n = [gr['fb'].bank_n(c) for c in range(2)]
        gr = None
        for _ in range(2):
            run_g(host_clip, keep_masks=False, frames_on_host=True)
        torch.cuda.synchronize()
        g2.record()
        for _ in range(args.steps):
            run_g(host_clip, keep_masks=False, frames_on_host=True)
        g3.record()
        torch.cuda.synchronize()
        graphed = {'value': args.frames * args.steps / (g0.elapsed_time(g1) / 1e3), 'unit': 'frames/s',
                   'e2e': args.frames * args.steps / (g2.elapsed_time(g3) / 1e3),
                   'final_bank_slots': gr_n,
                   'note': 'encoder_q+KeyValue, decoder trunk, local head and memorize replayed as four CUDA graphs; read, '
                           'URR and update as in the eager patched model'}
        del gm
    except Exception as e:                                                                # reported, never hidden
        graphed = {'value': None, 'error': f'{type(e).__name__}: {e}'[:300]}
    # the same weights with vfloodnet_b200.fuse_model on top (segment glue without per-object copies, the
    # Refine skip branches once per frame, KeyValue head on the tcgen05 implicit GEMM writing bank / query layout),
    # eager and with the convolution stages as CUDA graphs
    fused = None
    try:
        model_fused = vfn.fuse_model(MC.patched_copy(model_ref, vfn), fold_bn=True)

        def leg(model, frames, host=False, pipeline=False):
            run = lambda: MC.run_clip(model, vfn.FeatureBank, frames, dev, budget=BUDGET, keep_masks=False,
                                      frames_on_host=host, pipeline=pipeline)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = None
            for _ in range(args.steps):
                r = None
                r = run()
            b.record()
            torch.cuda.synchronize()
            return args.frames * args.steps / (a.elapsed_time(b) / 1e3), [r['fb'].bank_n(c) for c in range(2)]

        f_eager, f_n = leg(model_fused, dev_clip)
        tmf = MC.StageTimer(dev)
        hooks = MC.instrument(model_fused, tmf)
        try:
            MC.run_clip(model_fused, vfn.FeatureBank, dev_clip, dev, budget=BUDGET, keep_masks=False, timer=tmf)
        finally:
            for h in hooks:
                h.remove()
        tot_f = tmf.totals()
        tot_f['urr'] = tot_ours['urr']
        gmf = vfn.GraphedAFBURR(model_fused, tuple(dev_clip[0].shape))
        f_graph, fg_n = leg(gmf, dev_clip)
        f_graph_e2e, _ = leg(gmf, host_clip, host=True)
        # frame-level pipelining: the next frame's encoder graph on a side stream, overlapping memorize + update
        f_pipe, fp_n = leg(gmf, dev_clip, pipeline=True)
        f_pipe_e2e, _ = leg(gmf, host_clip, host=True, pipeline=True)
        kv = model_fused.keyval_r4
        fused = {'value': f_eager, 'unit': 'frames/s', 'graphed': f_graph, 'graphed_e2e': f_graph_e2e,
                 'graphed_pipelined': f_pipe, 'graphed_pipelined_e2e': f_pipe_e2e, 'final_bank_slots_pipelined': fp_n,
                 'final_bank_slots': f_n, 'final_bank_slots_graphed': fg_n,
                 'stages_ms_per_frame': _stage_table(tot_f, args.frames), 'keyvalue_passes': kv.passes,
                 'note': 'fuse_model: Refine skip branches evaluated once per frame (not per object), r1 / r2 / r3 never '
                         'expanded, KeyValue as one fp32-grade tcgen05 implicit GEMM (3 passes of fp16 hi/lo operands) '
                         'handing keys / values over entry-major; encoders with BatchNorm folded and conv + bias + ReLU '
                         '(+ add) as single cuDNN calls (n4)'}
        del gmf, model_fused
    except Exception as e:                                                                # reported, never hidden
        fused = {'value': None, 'error': f'{type(e).__name__}: {e}'[:300]}
    # the unmodified reference on the same GPU: same weights, same clip
    torch.backends.cuda.matmul.allow_tf32 = False
    clip_ref(dev_clip[:4], keep_masks=False)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    clip_ref(dev_clip, keep_masks=False)
    r1.record()
    torch.cuda.synchronize()
    ms_ref = r0.elapsed_time(r1)
    tm_r = MC.StageTimer(dev)
    hooks = MC.instrument(model_ref, tm_r)
    ref_run = clip_ref(dev_clip, timer=tm_r)
    for h in hooks:
        h.remove()
    st_ref = _stage_table(tm_r.totals(), args.frames)
    ious = [MC.iou(a, b) for a, b in zip(ref_run['masks'], ours_run['masks'])]
    n_ref = [int(ref_run['fb'].keys[c].shape[1]) for c in range(2)]
    if rank != 0:
        return
    frames_total = args.frames * args.steps
    frame_bytes = host_clip[1].numel() * 4
    line = {'metric': METRIC, 'value': frames_total / (ms / 1e3), 'unit': 'frames/s', 'n_gpus': 1, 'steps': args.steps,
            'warmup': max(args.warmup, 1), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 convolutions (cuDNN, TF32 allowed: the reference default); hot path as in '
                                          'the default workload', 'data': 'synthetic',
            'config': config_dict(args),
            'run': {'final_bank_slots': final_n, 'host_placement': placement,
                    'model': f'reference AFB_URR (random init seed {MC.MODEL_SEED}, BN statistics and output scale '
                             'calibrated), patch_model + vfloodnet_b200.FeatureBank'},
            'e2e': {'value': frames_total / (ms_e2e / 1e3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': (args.frames + 1) * frame_bytes,
                    'd2h_bytes_per_step': args.frames * host_clip[1].shape[-1] * host_clip[1].shape[-2]},
            'gpu_launches': int(launches), 'stages_ms_per_frame': st_ours,
            'stages_pass_wall_ms_per_frame': stage_wall, 'graphed_convolutions': graphed, 'fused_glue': fused,
            'reference_gpu': {'value': args.frames / (ms_ref / 1e3), 'unit': 'frames/s', 'kind': 'reference',
                              'sample': 'unmodified reference AFB_URR + FeatureBank (baseline/_ref), torch CUDA ops on '
                                        'the same GPU, same weights, same clip, free-running',
                              'stages_ms_per_frame': st_ref, 'final_bank_slots': n_ref},
            'parity': {'mask_iou_frames_1_to_5': [round(x, 5) for x in ious[:5]],
                       'frames_with_iou_ge_0_999': next((t for t, x in enumerate(ious) if x < 0.999), len(ious)),
                       'note': 'two FREE-RUNNING clips with TF32 convolutions: an untrained model is an unstable recurrence and '
                               'any two arms part ways within ~8 frames; the gated comparison (teacher-forced, exact arm, '
                               'fp32 convolutions) is tests/test_gpu_dropin.py'},
            'clocks': sampler.summary()}
    print(json.dumps(line))


def main_reference_model(args):
    """the unmodified reference model on the host cores: the first frames of the same clip, free-running, bounded"""
    from baseline import refshim, model_clip as MC
    if not refshim.available():
        print(json.dumps({'impl': 'reference', 'unavailable': 'reference not staged (python baseline/make_ref.py)'}))
        return
    torch.set_num_threads(os.cpu_count())
    ns = refshim.load()
    model = MC.build_reference_model(ns, 'cpu')
    n = min(args.frames, int(os.environ.get('VFN_REF_MODEL_FRAMES', '12')))
    clip = MC.make_clip(n)
    tm = MC.StageTimer('cpu')
    hooks = MC.instrument(model, tm)
    MC.run_clip(model, ns.FeatureBank, clip, 'cpu', budget=BUDGET, timer=tm, warm_frames=min(args.warmup, 1),
                keep_masks=False)
    for h in hooks:
        h.remove()
    tot = tm.totals()
    timed = n - min(args.warmup, 1)
    secs = (tot['segment'] + tot['memorize'] + tot['update']) / 1e3
    v = timed / secs
    print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * secs / args.steps,
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                      'data': 'synthetic', 'config': config_dict(args),
                      'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': 'reference',
                                       'sample': f'unmodified reference model, frames 1..{n} of the clip (bank still small)'},
                      'stages_ms_per_frame': _stage_table(tot, timed),
                      'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                      'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------------
# --workload 480p-64-streams (BASELINE configs[3])
# ---------------------------------------------------------------------------------------------------
def main_streams(args, rank, world, local_rank):
    """64 seeded video streams, stream s on GPU s mod G; per GPU `concurrency` of them run at a time, each on its own
    CUDA stream with its own bank, matcher workspace and frame tail (no shared state, no inter-GPU traffic).  The host
    issues frame t of every stream of a wave round-robin, so kernels of different streams overlap on the device (the
    small-bank frames are launch-latency bound on their own: 0.32 ms for 0.1 ms of work).  Clips are generated in HBM by
    a seeded CUDA generator (stream seed 1000 + s) before the timed region; a stream's result (bank sizes, replace_n,
    the water levels of every frame) does not depend on G or on its GPU - `stream_checksums` lets two runs be compared."""
    import hashlib
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import _lib, synth, tail as vtail
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    placement = {'bound': False, 'note': 'disabled'} if args.no_affinity else bind_to_gpu_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    mine = [s for s in range(args.total_streams) if s % world == rank]
    conc = max(1, min(args.concurrency, len(mine)))
    hw = HW_H * HW_W
    clips = {}
    for s_id in mine:
        gen = synth.ClipGenerator(seed=1000 + s_id, obj_n=2, hw=hw, frac_merge=args.frac_merge, device=dev)
        keys0, vals0 = gen.init()
        frames = [gen.frame() for _ in range(args.frames)]
        g = torch.Generator(device=dev).manual_seed(5000 + s_id)
        clips[s_id] = dict(keys0=keys0, vals0=vals0, frames=frames, urr=synth.gen_urr_inputs(g, 2, R1_H, R1_W))
    cuda_streams = [torch.cuda.Stream(dev) for _ in range(conc)]
    levels_all = {s_id: torch.zeros((args.frames, len(TAIL_KEY_PTS)), device=dev) for s_id in mine}
    results = {}

    def run_pass(record):
        for w0 in range(0, len(mine), conc):
            wave = mine[w0:w0 + conc]
            st = []
            for i, s_id in enumerate(wave):
                with torch.cuda.stream(cuda_streams[i]):
                    fb = vfn.FeatureBank(2, BUDGET, dev, impl=args.read_impl)
                    c = clips[s_id]
                    fb.init_bank(c['keys0'], c['vals0'])
                    st.append((fb, vfn.Matcher(update_bank=True), vtail.FrameTail(TAIL_SIZE, TAIL_KEY_PTS, dev)))
            for t in range(args.frames):
                for i, s_id in enumerate(wave):
                    fb, m, ft = st[i]
                    c = clips[s_id]
                    q_in, q_out, pk, pv = c['frames'][t]
                    p, r1, q_local = c['urr']
                    with torch.cuda.stream(cuda_streams[i]):
                        m(fb, q_in, q_out)
                        p_up, unc, conf, _lm = vfn.urr_pre(p, r1.expand(2, -1, -1, -1), (1, 2, R1_H, R1_W))
                        prob = vfn.urr_post(p_up, unc, conf, q_local)
                        fb.update(pk, pv, t + 1)
                        _, levels = ft(prob)
                        if record:
                            levels_all[s_id][t].copy_(levels)
            for i, s_id in enumerate(wave):
                cuda_streams[i].synchronize()
                if record:
                    fb = st[i][0]
                    results[s_id] = dict(bank=[fb.bank_n(c) for c in range(2)], replace_n=fb.replace_n.tolist())
            del st

    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get('VFN_BENCH_NO_SAMPLER'):
        sampler.start()
    for _ in range(args.warmup):
        run_pass(False)
    barrier()
    sampler.mark()
    l0 = lib.vfn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        run_pass(k == args.steps - 1)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.vfn_launch_count() - l0
    sampler.stop_flag = True
    # the same pass with ONE stream in flight per GPU: what the concurrency buys
    conc_saved, conc = conc, 1
    run_pass(False)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    run_pass(False)
    s1.record()
    barrier()
    ms_serial = s0.elapsed_time(s1)
    conc = conc_saved
    sums = {}
    for s_id in mine:
        h = hashlib.sha256(levels_all[s_id].cpu().numpy().tobytes())
        h.update(json.dumps(results[s_id], sort_keys=True).encode())
        sums[s_id] = h.hexdigest()[:16]
    t_ms = torch.tensor([ms, ms_serial], dtype=torch.float64, device=dev)
    all_sums = [sums]
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        all_sums = [None] * world
        dist.all_gather_object(all_sums, sums)
    ms, ms_serial = t_ms.tolist()
    if rank == 0:
        merged = {}
        for d in all_sums:
            merged.update(d)
        frames_total = args.frames * args.total_streams * args.steps
        fleet = hashlib.sha256(json.dumps(sorted(merged.items())).encode()).hexdigest()[:16]
        line = {'metric': METRIC, 'value': frames_total / (ms / 1e3), 'unit': 'frames/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f16 hi/lo + f8 split operands, f32 accumulate (read, match); f32 (update, URR, tail)',
                'data': 'synthetic (seeded CUDA generator, stream seed 1000 + s)', 'config': config_dict(args),
                'run': {'total_streams': args.total_streams, 'streams_per_gpu': len(mine), 'in_flight_per_gpu': conc,
                        'host_placement': placement, 'read_impl': args.read_impl},
                'one_stream_in_flight': {'value': args.frames * args.total_streams / (ms_serial / 1e3), 'unit': 'frames/s',
                                         'note': 'the same streams, one at a time per GPU'},
                'gpu_launches': int(launches), 'clocks': sampler.summary(),
                'stream_checksums': {'all_streams': fleet, 'first': [merged[s] for s in sorted(merged)[:4]],
                                     'note': 'sha256 over every frame\'s water levels + final bank sizes + replace_n, per '
                                             'stream; all_streams must not depend on --gpus'},
                'stream_results_sample': {str(s): results[s] for s in mine[:2]}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# --workload 1080p-2obj-bank-at-capacity --frames 2000   (BASELINE configs[2] as written: long-video stress)
# ---------------------------------------------------------------------------------------------------
def main_long(args, rank, world, local_rank):
    """2000 frames at the native 1080p grid (HW = 8160) against a bank at its 100000-slot budget: LFU eviction on every
    frame.  Frames are synthesised ON THE DEVICE in chunks of 50 (a resident clip would be 146 GB) by a seeded CUDA
    generator, outside the timed regions; every chunk is timed with CUDA events.  Reports frames/s over all chunks, the
    first / last 100 frames, the allocator high-water mark at the start and at the end (growth = leak), peak_n, replace_n
    and the info clamp (FeatureBank.py:113-115,140-141)."""
    import vfloodnet_b200 as vfn
    from vfloodnet_b200 import _lib, synth
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    hw, chunk = HW_H * HW_W, 50
    gen = synth.ClipGenerator(seed=100 + rank, obj_n=2, hw=hw, frac_merge=args.frac_merge, n_init=N_INIT, device=dev)
    keys0, vals0 = gen.init()
    g = torch.Generator(device=dev).manual_seed(1100 + rank)
    p, r1, q_local = synth.gen_urr_inputs(g, 2, R1_H, R1_W)
    info0 = []
    for _ in range(2):
        i = synth.gen_info(g, N_INIT, START_FRAME)
        i[:, 0] = torch.sort(i[:, 0]).values
        info0.append(i)
    fb = vfn.FeatureBank(2, BUDGET, dev, impl=args.read_impl)
    fb.load_state(keys0, vals0, info0)
    del keys0, vals0
    m = vfn.Matcher(update_bank=True)
    chunks_ms, sizes, mem = [], [], []
    l0 = lib.vfn_launch_count()
    done = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    while done < args.frames:
        n = min(chunk, args.frames - done)
        frames = [gen.frame() for _ in range(n)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t, (q_in, q_out, pk, pv) in enumerate(frames):
            m(fb, q_in, q_out)
            p_up, unc, conf, _lm = vfn.urr_pre(p, r1.expand(2, -1, -1, -1), (1, 2, R1_H, R1_W))
            vfn.urr_post(p_up, unc, conf, q_local)
            fb.update(pk, pv, START_FRAME + done + t + 1)
        e1.record()
        torch.cuda.synchronize()
        chunks_ms.append(e0.elapsed_time(e1))
        done += n
        sizes.append([fb.bank_n(c) for c in range(2)])
        mem.append(torch.cuda.max_memory_allocated(dev) / 2 ** 30)
        del frames
    sampler.stop_flag = True
    total_ms = sum(chunks_ms)
    per = chunk
    info_max = max(float(fb.info[c][:, 1].max()) for c in range(2))
    line = {'metric': METRIC, 'value': args.frames / (total_ms / 1e3), 'unit': 'frames/s', 'n_gpus': 1, 'steps': 1,
            'warmup': 0, 'ms_per_step': total_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 hi/lo + f8 split operands, f32 accumulate (read, match); f32 (update, URR)',
            'data': 'synthetic (seeded CUDA generator, 50-frame chunks synthesised on the device between timed regions)',
            'config': config_dict(args),
            'long_run': {'frames': args.frames, 'ms_per_frame_first_100': sum(chunks_ms[:100 // per]) / 100,
                         'ms_per_frame_last_100': sum(chunks_ms[-(100 // per):]) / 100,
                         'ms_per_frame_min_chunk': min(chunks_ms) / per, 'ms_per_frame_max_chunk': max(chunks_ms) / per,
                         'bank_slots_first_chunk': sizes[0], 'bank_slots_last_chunk': sizes[-1],
                         'peak_n': fb.peak_n.tolist(), 'replace_n': fb.replace_n.tolist(),
                         'replace_over_budget': (fb.replace_n / fb.class_budget).tolist(),
                         'allocator_high_water_gib_after_first_chunk': mem[0], 'allocator_high_water_gib_at_end': mem[-1],
                         'info_col1_max': info_max, 'info_clamp': 1e5},
            'gpu_launches': int(lib.vfn_launch_count() - l0), 'clocks': sampler.summary()}
    print(json.dumps(line))


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
        if WORKLOAD == '4k-2obj-sharded-bank':
            sys.path.insert(0, os.path.join(ROOT, 'tests'))
            import multi_gpu_sharded_bench
            os.environ.setdefault('RANK', '0'); os.environ.setdefault('WORLD_SIZE', '1'); os.environ.setdefault('LOCAL_RANK', '0')
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29519')
            multi_gpu_sharded_bench.main(['--frames', str(args.frames)])
        elif WORKLOAD == '480p-model-clip':
            main_model_clip(args, rank, world, local_rank)
        elif WORKLOAD == '480p-64-streams':
            main_streams(args, rank, world, local_rank)
        elif WORKLOAD == '1080p-2obj-bank-at-capacity' and args.frames > 200:
            main_long(args, rank, world, local_rank)
        else:
            main_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
