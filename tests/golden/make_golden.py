"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Two import shims, neither touching arithmetic of the reference's own code:
  * ``matplotlib`` stub  - myutils/__init__.py:2 imports plot_depth -> matplotlib (import-only).
  * ``torch_scatter``    - rusty1s/pytorch_scatter is not installed/vendored; ``scatter_mean`` is
    provided with the torch-scatter 2.0.8 published semantics (scatter_add_, count, clamp(1), true_divide_).
    This shim is written independently of oracle/ so the oracle's restatement is checked against it.

Outputs (committed): tests/golden/read_*.npz, update_*.npz, urr_*.npz, keyvalue.npz, misc.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('VFN_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)

    ts = types.ModuleType('torch_scatter')

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        assert out is not None
        out.scatter_add_(dim, index, src)
        ones = torch.ones_like(src)
        cnt = torch.zeros_like(out).scatter_add_(dim, index, ones)
        cnt[cnt < 1] = 1
        out.true_divide_(cnt)
        return out

    ts.scatter_mean = scatter_mean
    sys.modules['torch_scatter'] = ts
    sys.path.insert(0, REF)


def t2n(x):
    return x.detach().cpu().numpy().copy()   # copy: the reference mutates bank tensors in place later


def gen_bank(g, d_k, d_v, n, s_k=1.58):
    return torch.randn(d_k, n, generator=g) * s_k, torch.randn(d_v, n, generator=g)


def gen_candidates(g, key, value, hw, frac_merge=0.5, dup=True, s_k=1.58):
    """Candidates: a fraction are noisy copies of bank columns (cos ~0.995 -> merge, with duplicates so
    several candidates hit one slot), the rest fresh (cos ~0 -> append).  SURVEY 8(d) regime C."""
    d_k, n = key.shape
    d_v = value.shape[0]
    n_m = int(hw * frac_merge)
    src = torch.randint(0, n, (n_m,), generator=g)
    if dup and n_m >= 4:
        src[1::4] = src[0::4][: len(src[1::4])]
    k_m = key[:, src] + 0.1 * s_k * torch.randn(d_k, n_m, generator=g)
    v_m = value[:, src] + 0.1 * torch.randn(d_v, n_m, generator=g)
    k_f = torch.randn(d_k, hw - n_m, generator=g) * s_k
    v_f = torch.randn(d_v, hw - n_m, generator=g)
    k = torch.cat([k_m, k_f], dim=1)
    v = torch.cat([v_m, v_f], dim=1)
    perm = torch.randperm(hw, generator=g)
    return k[:, perm].contiguous(), v[:, perm].contiguous()


def golden_read(name, seed, d_k, d_v, ns, hw):
    from video_module.model.FeatureBank import FeatureBank
    from video_module.model.AFB_URR import Matcher
    g = torch.Generator().manual_seed(seed)
    obj_n = len(ns)
    fb = FeatureBank(obj_n, 10 ** 6, 'cpu')
    keys, vals = zip(*[gen_bank(g, d_k, d_v, n) for n in ns])
    fb.init_bank([k.clone() for k in keys], [v.clone() for v in vals], frame_idx=0)
    for c in range(obj_n):
        fb.info[c][:, 1] = torch.rand(ns[c], generator=g) * 5
    info0 = [fb.info[c].clone() for c in range(obj_n)]
    q_in = torch.randn(1, d_k, hw, generator=g) * 1.58
    q_out = torch.randn(1, d_v, hw, generator=g)
    m = Matcher(thres_valid=1e-3, update_bank=True)
    out = m(fb, q_in, q_out)
    d = {'q_in': t2n(q_in), 'q_out': t2n(q_out), 'out': t2n(out), 'obj_n': obj_n}
    for c in range(obj_n):
        d[f'key{c}'] = t2n(keys[c]); d[f'val{c}'] = t2n(vals[c])
        d[f'info_before{c}'] = t2n(info0[c]); d[f'info_after{c}'] = t2n(fb.info[c])
    np.savez(os.path.join(HERE, f'read_{name}.npz'), **d)
    print('read', name, out.shape)


def golden_update(name, seed, d_k, d_v, n0, hw, frames, budget, thres_close=0.95, frac_merge=0.5, s_k=1.58):
    """Frame loop of Matcher (usage counts) + FeatureBank.update, state saved after every frame."""
    from video_module.model.FeatureBank import FeatureBank
    from video_module.model.AFB_URR import Matcher
    g = torch.Generator().manual_seed(seed)
    obj_n = 2
    fb = FeatureBank(obj_n, budget, 'cpu', update_rate=0.1, thres_close=thres_close)
    keys, vals = zip(*[gen_bank(g, d_k, d_v, n0, s_k) for _ in range(obj_n)])
    fb.init_bank([k.clone() for k in keys], [v.clone() for v in vals])
    m = Matcher(thres_valid=1e-3, update_bank=True)
    d = {'obj_n': obj_n, 'budget': budget, 'frames': frames, 'thres_close': thres_close,
         'class_budget': float(fb.class_budget)}
    for c in range(obj_n):
        d[f'key_init{c}'] = t2n(keys[c]); d[f'val_init{c}'] = t2n(vals[c])
    for t in range(1, frames + 1):
        q_in = torch.randn(1, d_k, hw, generator=g) * s_k
        q_out = torch.randn(1, d_v, hw, generator=g)
        out = m(fb, q_in, q_out)
        pk, pv = zip(*[gen_candidates(g, fb.keys[c], fb.values[c], hw, frac_merge, s_k=s_k) for c in range(obj_n)])
        d[f'f{t}_q_in'] = t2n(q_in); d[f'f{t}_q_out'] = t2n(q_out); d[f'f{t}_out'] = t2n(out)
        for c in range(obj_n):
            d[f'f{t}_pk{c}'] = t2n(pk[c]); d[f'f{t}_pv{c}'] = t2n(pv[c])
            d[f'f{t}_info_read{c}'] = t2n(fb.info[c])
        fb.update([k.clone() for k in pk], [v.clone() for v in pv], t)
        for c in range(obj_n):
            d[f'f{t}_key{c}'] = t2n(fb.keys[c]); d[f'f{t}_val{c}'] = t2n(fb.values[c])
            d[f'f{t}_info{c}'] = t2n(fb.info[c])
        d[f'f{t}_peak_n'] = fb.peak_n.copy(); d[f'f{t}_replace_n'] = fb.replace_n.copy()
        print('update', name, 'frame', t, [fb.keys[c].shape[1] for c in range(obj_n)], fb.replace_n)
    np.savez(os.path.join(HERE, f'update_{name}.npz'), **d)


def golden_append_api(seed=5):
    """FeatureBank.append (unused by the CLIs but part of the API surface, FeatureBank.py:38-51)."""
    from video_module.model.FeatureBank import FeatureBank
    g = torch.Generator().manual_seed(seed)
    fb = FeatureBank(2, 1000, 'cpu')
    k0, v0 = zip(*[gen_bank(g, 8, 16, 5) for _ in range(2)])
    k1, v1 = zip(*[gen_bank(g, 8, 16, 3) for _ in range(2)])
    fb.append([k.clone() for k in k0], [v.clone() for v in v0], frame_idx=2)   # empty bank -> init_bank
    fb.append([k.clone() for k in k1], [v.clone() for v in v1], frame_idx=7)
    d = {}
    for c in range(2):
        d[f'k0_{c}'] = t2n(k0[c]); d[f'v0_{c}'] = t2n(v0[c]); d[f'k1_{c}'] = t2n(k1[c]); d[f'v1_{c}'] = t2n(v1[c])
        d[f'key{c}'] = t2n(fb.keys[c]); d[f'val{c}'] = t2n(fb.values[c]); d[f'info{c}'] = t2n(fb.info[c])
    d['peak_n'] = fb.peak_n.copy()
    fb3 = FeatureBank(3, 1000, 'cpu')
    d['class_budget_obj2'] = float(fb.class_budget); d['class_budget_obj3'] = float(fb3.class_budget)
    return d


def golden_urr(name, seed, H, W):
    from video_module.model.AFB_URR import Decoder
    torch.manual_seed(seed)
    dec = Decoder('cpu').eval()
    obj_n = 2
    patch = torch.randn(obj_n, 1024, H // 16, W // 16) * 0.5
    r3 = torch.randn(obj_n, 512, H // 8, W // 8).relu()
    r2 = torch.randn(obj_n, 256, H // 4, W // 4).relu()
    r1 = torch.randn(1, 64, H // 2, W // 2).relu().expand(obj_n, -1, -1, -1).contiguous()
    cap = {}
    dec.pred2.register_forward_hook(lambda m, i, o: cap.__setitem__('pred2', o.detach().clone()))
    dec.local_convFM.register_forward_hook(lambda m, i, o: cap.__setitem__('local_match', i[0].detach().clone()))
    dec.local_pred2.register_forward_hook(lambda m, i, o: cap.__setitem__('q_local', o.detach().clone()))
    with torch.no_grad():
        out = dec(patch, r3, r2, r1, (1, obj_n, H // 2, W // 2))
    np.savez(os.path.join(HERE, f'urr_{name}.npz'), p=t2n(cap['pred2']), r1=t2n(r1),
             local_match=t2n(cap['local_match']), q_local=t2n(cap['q_local']), out=t2n(out),
             feature_shape=np.array([1, obj_n, H // 2, W // 2]))
    print('urr', name, out.shape)


def golden_keyvalue(seed=21):
    """KeyValue.forward (AFB_URR.py:94-111) of the unmodified reference module on seeded inputs: indim 64 (the kernel
    serves any multiple of 64), keydim 128, valdim 512, two images of 6 x 7; and the decoder trunk (convFM .. pred2,
    AFB_URR.py:209-212) with per-object copies of r3 / r2, for the copy-free glue (round 2)."""
    from video_module.model.AFB_URR import KeyValue, Decoder
    torch.manual_seed(seed)
    kv = KeyValue(64, keydim=128, valdim=512).eval()
    x = torch.randn(2, 64, 6, 7).relu() * 3
    with torch.no_grad():
        k, v = kv(x)
    d = dict(x=t2n(x), wk=t2n(kv.Key.weight), bk=t2n(kv.Key.bias), wv=t2n(kv.Value.weight), bv=t2n(kv.Value.bias),
             key=t2n(k), val=t2n(v))
    dec = Decoder('cpu').eval()
    H, W, obj_n = 32, 48, 2
    patch = torch.randn(obj_n, 1024, H // 16, W // 16) * 0.5
    r3 = torch.randn(1, 512, H // 8, W // 8).relu()
    r2 = torch.randn(1, 256, H // 4, W // 4).relu()
    with torch.no_grad():
        p = dec.ResMM(dec.convFM(patch))
        p = dec.RF3(r3.expand(obj_n, -1, -1, -1).reshape(obj_n, *r3.shape[1:]), p)       # AFB_URR.py:291-292
        p = dec.RF2(r2.expand(obj_n, -1, -1, -1).reshape(obj_n, *r2.shape[1:]), p)
        p = dec.pred2(torch.relu(p))
    d.update(trunk_patch=t2n(patch), trunk_r3=t2n(r3), trunk_r2=t2n(r2), trunk_out=t2n(p), trunk_seed=np.array(seed))
    np.savez(os.path.join(HERE, 'keyvalue.npz'), **d)
    print('keyvalue', k.shape, v.shape, p.shape)


def golden_misc():
    import myutils
    g = torch.Generator().manual_seed(11)
    d = golden_append_api()
    score = torch.softmax(torch.randn(1, 2, 9, 13, generator=g) * 3, dim=1)
    d['unc_in'] = t2n(score); d['unc_out'] = t2n(myutils.calc_uncertainty(score))
    x = torch.randn(1, 3, 30, 53, generator=g)
    (y,), pad = myutils.pad_divide_by([x], 16, (30, 53))
    d['pad_in'] = t2n(x); d['pad_out'] = t2n(y); d['pad_array'] = np.array(pad)
    np.savez(os.path.join(HERE, 'misc.npz'), **d)
    print('misc ok', pad)


if __name__ == '__main__':
    install_shims()
    torch.set_num_threads(1)   # deterministic reduction order for the committed vectors
    if 'keyvalue' in sys.argv[1:]:             # added in round 2 (n3): generate this file alone
        golden_keyvalue()
        sys.exit(0)
    if 'real_dims_evict' in sys.argv[1:]:      # added in round 2: regenerate this file alone
        golden_update('real_dims_evict', 13, 128, 512, 100, 64, 3, budget=350, s_k=3.0)
        sys.exit(0)
    golden_read('real_dims', 1, 128, 512, (96, 77), 40)
    golden_read('one_slot', 2, 16, 24, (1, 3), 7)
    golden_update('small_evict', 3, 16, 24, 20, 24, 8, budget=150)          # class_budget 60.0 -> eviction fires
    golden_update('small_evict2', 9, 16, 24, 40, 32, 8, budget=200, s_k=3.0)               # peaked softmax -> LFU spread, partial evictions
    golden_update('small_allmerge', 4, 16, 24, 30, 16, 3, budget=10 ** 5, thres_close=-1.0, frac_merge=1.0)
    golden_update('small_allappend', 6, 16, 24, 30, 16, 3, budget=10 ** 5, thres_close=2.0, frac_merge=0.0)
    golden_update('real_dims', 7, 128, 512, 64, 36, 2, budget=260)          # class_budget 104.0
    # d_k=128 / d_v=512 WITH LFU eviction (class_budget 140.0, peaked softmax): the only dims the tcgen05 match/read serve
    golden_update('real_dims_evict', 13, 128, 512, 100, 64, 3, budget=350, s_k=3.0)
    golden_urr('h32w48', 8, 32, 48)
    golden_keyvalue()
    golden_misc()
