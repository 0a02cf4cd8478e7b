"""Bring-up diagnostics for the tcgen05 read (not a pytest): python tests/debug_tc.py N HW [seed]
Compares impl=2 (tcgen05) against impl=1 (fp32 SIMT) and the CPU oracle, and checks the dumped S^T tile."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as v
from vfloodnet_b200 import _lib, synth
from oracle import afb_oracle as O


def main():
    n, hw = int(sys.argv[1]), int(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    ns = [n, max(1, n - 37)]
    keys, vals = zip(*[synth.gen_bank(g, k) for k in ns])
    info = [synth.gen_info(g, k, 10) for k in ns]
    q_in, q_out = synth.gen_query(g, hw)
    res = {}
    for impl in (1, 2):
        fb = v.FeatureBank(2, 10 ** 6, 'cuda', impl=impl)
        fb.load_state(list(keys), list(vals), info)
        m = v.Matcher(update_bank=True)
        m.want_lse = True
        dump = None
        if impl == 2:
            dump = torch.full((128 * 128,), float('nan'), device='cuda')
            lib.vfn_debug_set_dump(dump.data_ptr())
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out = m(fb, q_in.cuda(), q_out.cuda())
        t1.record()
        torch.cuda.synchronize()
        lib.vfn_debug_set_dump(None)
        res[impl] = dict(out=out.cpu(), lse=m.last_lse.cpu(), info=[fb.info[c].cpu() for c in range(2)],
                         dump=None if dump is None else dump.cpu(), ms=t0.elapsed_time(t1))
        print(f'impl {impl}: {res[impl]["ms"]:.3f} ms (first call, includes workspace alloc)')
    a, b = res[1], res[2]
    print('lse   max|tc-simt| =', (a['lse'] - b['lse']).abs().max().item())
    print('mem   max|tc-simt| =', (a['out'][0, :, :512] - b['out'][0, :, :512]).abs().max().item())
    print('q_out equal        =', torch.equal(a['out'][0, :, 512:], b['out'][0, :, 512:]))
    for c in range(2):
        d = (a['info'][c][:, 1] - b['info'][c][:, 1]).abs()
        print(f'info[{c}] max diff = {d.max().item():.3e}  rows differing = {(d > 1e-6).sum().item()} / {ns[c]}')
    # dumped S^T tile of phase B (CTA 0: object 0, query tile 0, slots 0..63), log2 domain
    s_ref = (q_in[0].t()[:128] @ keys[0][:, :64]) * (math.log2(math.e) / math.sqrt(128))
    d = b['dump'][:128 * 64].view(128, 64)[: min(128, hw), : min(64, ns[0])]
    r = s_ref[: min(128, hw), : min(64, ns[0])]
    print('S tile max err     =', (d - r).abs().max().item(), ' ref max |s| =', r.abs().max().item())
    if n <= 30000:
        rr = O.matcher_forward(list(keys), list(vals), [i.clone() for i in info], q_in, q_out, 1e-3, True)
        print('mem   max|tc-oracle|   =', (b['out'] - rr.out).abs().max().item())
        print('mem   max|simt-oracle| =', (a['out'] - rr.out).abs().max().item())
        print('lse   max|tc-oracle|   =', max((b['lse'][c] - rr.lse[c]).abs().max().item() for c in range(2)))


if __name__ == '__main__':
    main()
