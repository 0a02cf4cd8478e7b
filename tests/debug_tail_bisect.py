import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from vfloodnet_b200 import tail, _lib
from oracle import tail_oracle as TO
lib = _lib.load()
rng = np.random.default_rng(1080 * 7 + 1920)
pred = (rng.random((1080, 1920)) < 0.55).astype(np.uint8)
cnt, labels = TO.grana_order_labels(pred)
ref = TO.postprocessing_pred(pred)
d = torch.from_numpy(pred).cuda()
for flags in (0, 2, 0):
    lib.vfn_debug_set_tail(flags)
    res = []
    for rep in range(5):
        m, st = tail.postprocessing_pred(d, return_stats=True)
        res.append((st.tolist(), bool(np.array_equal(m.cpu().numpy(), ref))))
    print('flags', flags, 'want comps', cnt - 1, 'kept', int(ref.sum()), res)
