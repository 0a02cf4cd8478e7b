"""CPU checks of the round-2 glue around the hot path: the oracle's KeyValue restatement against
the golden vectors of the unmodified reference module, the copy-free decoder trunk and the folded encoders against the
reference's own modules (torch CPU), and the stride test that recognises entry-major hand-overs.  No CUDA code runs
here: the kernels themselves are tested in tests/test_gpu_keyvalue.py."""
import os

import numpy as np
import pytest
import torch
from torch.nn import functional as NF

from oracle import afb_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_oracle_keyvalue_matches_the_reference_module():
    """tests/golden/keyvalue.npz: KeyValue.forward of the unmodified reference (make_golden.py golden_keyvalue)"""
    g = np.load(os.path.join(GOLD, 'keyvalue.npz'))
    k, v = O.keyvalue_forward(_t(g['x']), _t(g['wk']), _t(g['bk']), _t(g['wv']), _t(g['bv']))
    assert k.shape == g['key'].shape and v.shape == g['val'].shape
    np.testing.assert_allclose(k.numpy(), g['key'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(v.numpy(), g['val'], rtol=0, atol=1e-5)
    # the float64 evaluation the GPU tests use as the exact value agrees with it to fp32 rounding
    k64, v64 = O.keyvalue_forward(_t(g['x']).double(), _t(g['wk']), _t(g['bk']), _t(g['wv']), _t(g['bv']))
    assert float((k64 - k.double()).abs().max()) < 1e-4 and float((v64 - v.double()).abs().max()) < 1e-4


def test_pad_divide_by16_is_the_reference_padding():
    from vfloodnet_b200 import glue
    g = np.load(os.path.join(GOLD, 'misc.npz'))
    y, pad = glue.pad_divide_by16(_t(g['pad_in']))
    assert tuple(pad) == tuple(int(p) for p in g['pad_array'])
    assert torch.equal(y, _t(g['pad_out']))
    for h, w in ((480, 854), (480, 853), (1080, 1920), (16, 16), (17, 31)):
        x = torch.zeros(1, 1, h, w)
        (yo,), po = O.pad_divide_by([x], 16, (h, w))
        yg, pg = glue.pad_divide_by16(x)
        assert tuple(po) == tuple(pg) and yo.shape == yg.shape


def test_em_backed_recognises_transposed_views_only():
    from vfloodnet_b200._lib import em_backed
    em = torch.zeros(2, 40, 16)
    view = em.transpose(1, 2)                   # (2, 16, 40): what KeyValueHead returns
    assert em_backed(view[0]) and em_backed(view[1])
    assert not em_backed(view[0].contiguous())
    assert not em_backed(torch.zeros(16, 40))
    assert not em_backed(torch.zeros(16, 40).double().t().t())
    assert not em_backed(view[0][:, ::2])       # a strided slice is not a whole row set
    assert not em_backed(torch.zeros(40, 16, dtype=torch.float16).t())


@pytest.fixture(scope='module')
def ref():
    from baseline import refshim
    if not refshim.available():
        pytest.skip('reference neither staged (baseline/_ref) nor present (/root/reference)')
    return refshim.load()


def test_shared_decoder_trunk_matches_the_reference_decoder(ref):
    """glue.decoder_trunk_shared (Refine skip branches once, broadcast over the objects) against the golden output of the
    reference Decoder modules fed with per-object copies of r3 / r2 (AFB_URR.py:209-212,289-292).  The Decoder weights
    are re-created by replaying make_golden.py's seeded construction order."""
    import sys
    from vfloodnet_b200 import glue
    g = np.load(os.path.join(GOLD, 'keyvalue.npz'))
    kv_mod = sys.modules['video_module.model.AFB_URR']
    torch.manual_seed(int(g['trunk_seed']))
    kv = kv_mod.KeyValue(64, keydim=128, valdim=512)
    x = torch.randn(2, 64, 6, 7)
    assert torch.equal(x.relu() * 3, _t(g['x'])) and torch.equal(kv.Key.weight.detach(), _t(g['wk']))
    dec = ref.Decoder('cpu').eval()
    patch = torch.randn(2, 1024, 2, 3) * 0.5
    assert torch.equal(patch, _t(g['trunk_patch']))
    with torch.no_grad():
        r3, r2 = _t(g['trunk_r3']), _t(g['trunk_r2'])
        p = glue.decoder_trunk_shared(dec, patch, glue.refine_skip(dec.RF3, r3), glue.refine_skip(dec.RF2, r2))
    np.testing.assert_allclose(p.numpy(), g['trunk_out'], rtol=0, atol=2e-4 * float(np.abs(g['trunk_out']).max()))


def test_folded_encoders_compute_the_reference_encoders(ref, monkeypatch):
    """vfloodnet_b200.folded: BatchNorm folded into the convolutions.  On the CPU the two fused cuDNN calls are replaced
    by their definitions (relu(conv + b), relu(conv + b + z)); the folding, the 5-channel stem of EncoderM and the
    bottleneck wiring are what is checked against the reference modules."""
    from baseline import model_clip
    from vfloodnet_b200 import folded
    monkeypatch.setattr(folded, '_conv_relu', lambda x, w, b, s, p: torch.relu(NF.conv2d(x, w, b, s, p)))
    monkeypatch.setattr(folded, '_conv_add_relu', lambda x, w, z, b, s, p: torch.relu(NF.conv2d(x, w, b, s, p) + z))
    model = model_clip.build_reference_model(ref, 'cpu', calibrate_outputs=False,
                                             calib_frames=[model_clip.make_frame(1000 + i, h=64, w=96) for i in range(2)])
    f = model_clip.make_frame(3, h=64, w=96)
    m = model_clip.first_mask(64, 96)
    mk = m[0].unsqueeze(1).float()
    args_m = (f.expand(2, -1, -1, -1), mk, (1 - mk).clamp(0, 1))
    with torch.no_grad():
        want_q = model.encoder_q(f)
        want_m = model.encoder_m(*args_m)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        folded.fold_encoders(model)
        got_q = model.encoder_q(f)
        got_m = model.encoder_m(*args_m)
        for a, b in list(zip(want_q, got_q)) + list(zip(want_m, got_m)):
            assert a.shape == b.shape
            assert float((a - b).abs().max()) <= 2e-4 * float(a.abs().max())
        assert all(torch.equal(sd[k], v) for k, v in model.state_dict().items())       # parameters untouched
        # a changed BatchNorm statistic must invalidate the folded weights
        model.encoder_q.bn1.running_mean.add_(0.5)
        again = model.encoder_q(f)[3]
        assert float((again - got_q[3]).abs().max()) > 1e-3
        # train mode falls through to the reference's own forward
        model.train()
        r_train = model.encoder_q(f)
        model.eval()
        assert len(r_train) == 4


def test_keyvalue_head_training_mode_is_the_reference_forward_and_eval_has_no_cpu_path():
    """train mode keeps autograd (the reference's two convolutions, AFB_URR.py:103-111); eval mode on a CPU tensor
    raises instead of falling back"""
    import vfloodnet_b200 as vfn
    torch.manual_seed(0)
    key, val = torch.nn.Conv2d(64, 128, 3, padding=1), torch.nn.Conv2d(64, 512, 3, padding=1)
    assert not vfn.KeyValueHead(key.eval(), val.eval()).training      # the head follows the mode of what it wraps
    head = vfn.KeyValueHead(key, val)
    x = torch.randn(2, 64, 5, 6)
    head.train()
    k, v = head(x)
    assert k.shape == (2, 128, 30) and v.shape == (2, 512, 30) and k.requires_grad
    k0, v0 = O.keyvalue_forward(x, key.weight, key.bias, val.weight, val.bias)
    assert torch.equal(k, k0) and torch.equal(v, v0)
    head.eval()
    with pytest.raises(RuntimeError):
        head(x)
    assert set(head.state_dict().keys()) == {'Key.weight', 'Key.bias', 'Value.weight', 'Value.bias'}
