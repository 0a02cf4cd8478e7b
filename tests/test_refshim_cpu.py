"""CPU checks of the reference staging (baseline/) and of the oracle against the WHOLE unmodified reference model loop
(SURVEY T4): `AFB_URR.memorize/segment` + `FeatureBank.update` of the reference vs the same model driving the oracle's
FeatureBank / Matcher restatement, on a small frame so that it runs in seconds."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import afb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def ref():
    from baseline import refshim
    if not refshim.available():
        pytest.skip('reference neither staged (baseline/_ref) nor present (/root/reference)')
    return refshim.load()


def test_staged_reference_is_unmodified():
    """every file under baseline/_ref/ is byte-identical to its source in /root/reference (when both are present)"""
    staged = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(staged, 'video_module')):
        pytest.skip('baseline/_ref not staged')
    man = json.load(open(os.path.join(staged, 'MANIFEST.json')))
    assert any(k.endswith('FeatureBank.py') for k in man['files'])
    for rel, sha in man['files'].items():
        assert hashlib.sha256(open(os.path.join(staged, rel), 'rb').read()).hexdigest() == sha, rel
        src = os.path.join('/root/reference', rel)
        if os.path.exists(src):
            assert hashlib.sha256(open(src, 'rb').read()).hexdigest() == sha, f'{rel} differs from the reference'


def test_product_package_never_imports_the_reference():
    pkg = os.path.join(ROOT, 'vfloodnet_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'baseline' not in src and 'refshim' not in src and 'video_module' not in src.replace(
                'video_module/model', ''), fn


class _OracleMatcher(torch.nn.Module):
    """the oracle's read behind the reference Matcher's call signature"""

    def __init__(self):
        super().__init__()
        self.update_bank, self.thres_valid = True, 1e-3

    def forward(self, fb, q_in, q_out):
        return O.matcher_forward(fb.keys, fb.values, fb.info, q_in, q_out, self.thres_valid, update_bank=True).out


@pytest.mark.parametrize('budget,thres', [(1200, 0.95), (1200, 0.70)])
def test_oracle_equals_reference_in_the_model_loop(ref, budget, thres):
    """the reference loop (test_video_seg.py:99-112) with the reference's FeatureBank/Matcher and with the oracle's:
    identical masks, bank sizes, insertion frames, replace_n; keys within fp32 noise.  64x96 frames -> HW = 24; the
    budget forces LFU evictions, thres 0.70 gives a merge/append mix on real (random-init, BN-calibrated) features."""
    import copy
    from baseline import model_clip as MC
    torch.set_num_threads(4)
    h, w, frames = 64, 96, 30
    calib = [MC.make_frame(1000 + i, 0, h, w) for i in range(3)]
    model = MC.build_reference_model(ref, 'cpu', calib_frames=calib)
    clip = MC.make_clip(frames, h=h, w=w)
    r = MC.run_clip(model, ref.FeatureBank, clip, 'cpu', budget=budget, thres_close=thres)
    model_o = copy.deepcopy(model)
    model_o.global_matcher = _OracleMatcher()
    o = MC.run_clip(model_o, O.OracleFeatureBank, clip, 'cpu', budget=budget, thres_close=thres)
    assert r['fb'].replace_n.sum() > 0, 'the small budget must force evictions'
    for a, b in zip(r['masks'], o['masks']):
        assert torch.equal(a, b)
    for c in range(2):
        assert r['fb'].keys[c].shape == o['fb'].keys[c].shape
        assert torch.equal(r['fb'].info[c][:, 0], o['fb'].info[c][:, 0])
        np.testing.assert_allclose(r['fb'].info[c][:, 1].numpy(), o['fb'].info[c][:, 1].numpy(), atol=1e-5)
        np.testing.assert_allclose(r['fb'].keys[c].numpy(), o['fb'].keys[c].numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(r['fb'].values[c].numpy(), o['fb'].values[c].numpy(), rtol=1e-5, atol=1e-5)
    assert np.array_equal(r['fb'].replace_n, o['fb'].replace_n) and np.array_equal(r['fb'].peak_n, o['fb'].peak_n)


def test_reference_arm_kinds(ref):
    """bench.py's reference arm drives the unmodified classes when they are staged and gives the oracle's results"""
    from baseline.ref_arm import RefArm
    from vfloodnet_b200 import synth
    gen = synth.ClipGenerator(seed=3, obj_n=2, hw=60, frac_merge=0.3)
    keys0, vals0 = gen.init()
    a, b = RefArm(600, 'cpu'), RefArm(600, 'cpu', prefer_reference=False)
    assert a.kind == 'reference' and b.kind == 'port'
    a.init(keys0, vals0); b.init(keys0, vals0)
    for t in range(8):
        q_in, q_out, pk, pv = gen.frame()
        oa, _ = a.frame(q_in, q_out, [k.clone() for k in pk], [v.clone() for v in pv], None, t + 1)
        ob, _ = b.frame(q_in, q_out, [k.clone() for k in pk], [v.clone() for v in pv], None, t + 1)
        assert torch.allclose(oa, ob, atol=1e-6)
        assert a.sizes() == b.sizes()
    assert a.fb.replace_n.sum() > 0 and np.array_equal(a.fb.replace_n, b.fb.replace_n)
