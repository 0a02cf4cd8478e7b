"""Where does a frame's wall time go?  python tests/debug_host_overhead.py  (GPU box; not a pytest)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vfloodnet_b200 as vfn

dev = torch.device('cuda', 0)
clip = bench.to_device(bench.make_clip(100, 100, 0.1, pin=False), dev)
for rep in range(3):
    fb = vfn.FeatureBank(2, bench.BUDGET, dev)
    m = vfn.Matcher(update_bank=True)
    fb.init_bank(clip['keys0'], clip['vals0'])
    torch.cuda.synchronize()
    acc = {k: [0.0, 0.0] for k in ('read', 'urr', 'update')}   # [issue (cpu) s, total-with-sync s]
    t_all = time.perf_counter()
    for t, (q_in, q_out, pk, pv) in enumerate(clip['frames']):
        p, r1, q_local = clip['urr']
        a = time.perf_counter(); out = m(fb, q_in, q_out); b = time.perf_counter(); torch.cuda.synchronize(); c = time.perf_counter()
        acc['read'][0] += b - a; acc['read'][1] += c - a
        a = time.perf_counter()
        p_up, unc, conf, lm = vfn.urr_pre(p, r1.expand(2, -1, -1, -1), (1, 2, bench.R1_H, bench.R1_W))
        prob = vfn.urr_post(p_up, unc, conf, q_local)
        b = time.perf_counter(); torch.cuda.synchronize(); c = time.perf_counter()
        acc['urr'][0] += b - a; acc['urr'][1] += c - a
        a = time.perf_counter(); fb.update(pk, pv, t + 1); b = time.perf_counter(); torch.cuda.synchronize(); c = time.perf_counter()
        acc['update'][0] += b - a; acc['update'][1] += c - a
    tot = time.perf_counter() - t_all
    print(f'rep {rep}: total {tot*1e3:.1f} ms for 100 frames (stage-synchronised); per frame ms: ' +
          ', '.join(f'{k}: issue {v[0]*10:.3f} / done {v[1]*10:.3f}' for k, v in acc.items()))
# free-running (no per-stage sync)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bench.run_clip_gpu(vfn, clip, dev, 0)
    torch.cuda.synchronize(); print(f'free-running clip: {(time.perf_counter()-t0)*1e3:.1f} ms')
# python-only cost of an update: time the calls with a tiny bank (GPU work negligible)
