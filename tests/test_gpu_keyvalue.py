"""KeyValue head on tcgen05 (vfn_keyvalue) and the copy-free segment glue (vfloodnet_b200.glue).

Reference: KeyValue.forward (video_module/model/AFB_URR.py:94-111) = two cuDNN 3x3 convolutions; AFB_URR.segment /
Refine.forward (AFB_URR.py:113-127,274-318).  The checker for the convolution is torch's own conv2d evaluated in
float64 (exact to ~1e-15: what the reference computes, without its rounding), with the reference's fp32 / TF32 cuDNN
results beside it: the bar for passes=3 is "as close to exact as the true-fp32 cuDNN convolution, within a factor",
for passes=1 "as close as the TF32 cuDNN convolution".  Layout variants must agree bit for bit, and the entry-major
hand-over to Matcher / FeatureBank.update must give the same results as the reference layout.
"""
import json
import os

import pytest
import torch
from torch.nn import functional as NF

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, d):
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f'keyvalue_report_{name}.json'), 'w') as f:
        json.dump(d, f, indent=1)


def _head(c_in=1024, dk=128, dv=512, seed=0, passes=3, bias=True):
    import vfloodnet_b200 as vfn
    g = torch.Generator().manual_seed(seed)
    key = torch.nn.Conv2d(c_in, dk, 3, padding=1, bias=bias)
    val = torch.nn.Conv2d(c_in, dv, 3, padding=1, bias=bias)
    with torch.no_grad():
        for conv in (key, val):
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (9 * c_in)) ** 0.5)
            if bias:
                conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.1)
    key, val = key.cuda().eval(), val.cuda().eval()
    return vfn.KeyValueHead(key, val, passes=passes)


def _features(b, c, h, w, seed=1):
    """post-ReLU-like feature map: half zeros, heavy tail (|x| up to ~30 like a BN-calibrated r4)"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, c, h, w, generator=g)
    x = torch.relu(x) * (1 + 4 * torch.rand(b, c, 1, 1, generator=g))
    return x.cuda()


def _exact(head, x):
    k = NF.conv2d(x.double(), head.Key.weight.double(), head.Key.bias.double() if head.Key.bias is not None else None,
                  padding=1)
    v = NF.conv2d(x.double(), head.Value.weight.double(),
                  head.Value.bias.double() if head.Value.bias is not None else None, padding=1)
    return k.flatten(2), v.flatten(2)


def _cudnn(head, x, tf32):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        with torch.no_grad():
            return head.Key(x).flatten(2), head.Value(x).flatten(2)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize('b', [1, 2])
def test_keyvalue_480p_fp32_grade(b):
    """r4 of a 480p frame (30 x 54), B = 1 (segment) and B = 2 (memorize with two objects)"""
    head = _head()
    x = _features(b, 1024, 30, 54)
    with torch.no_grad():
        k, v = head(x)
    assert k.shape == (b, 128, 1620) and v.shape == (b, 512, 1620)
    ke, ve = _exact(head, x)
    k32, v32 = _cudnn(head, x, tf32=False)
    ktf, vtf = _cudnn(head, x, tf32=True)
    scale = float(ve.detach().abs().max())
    err = max(float((k.double() - ke).abs().max()), float((v.double() - ve).abs().max()))
    err32 = max(float((k32.double() - ke).abs().max()), float((v32.double() - ve).abs().max()))
    errtf = max(float((ktf.double() - ke).abs().max()), float((vtf.double() - ve).abs().max()))
    _report(f'480p_b{b}', dict(out_absmax=scale, err_vs_exact=err, cudnn_fp32_err_vs_exact=err32,
                              cudnn_tf32_err_vs_exact=errtf, rel=err / scale))
    # fp32 grade: within 4x of the true-fp32 cuDNN convolution's own distance from exact arithmetic, and two orders of
    # magnitude inside the TF32 result
    assert err <= max(4 * err32, 1e-5 * scale), (err, err32, errtf, scale)
    assert err <= 0.05 * errtf, (err, errtf)


def test_keyvalue_golden_of_the_reference_module():
    """tests/golden/keyvalue.npz: KeyValue.forward of the unmodified reference (CPU fp32) on seeded inputs; the kernel
    with the same weights must reproduce it to fp32 convolution rounding, in both layouts"""
    import numpy as np
    import vfloodnet_b200 as vfn
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'keyvalue.npz'))
    key = torch.nn.Conv2d(64, 128, 3, padding=1)
    val = torch.nn.Conv2d(64, 512, 3, padding=1)
    with torch.no_grad():
        key.weight.copy_(torch.from_numpy(g['wk'])); key.bias.copy_(torch.from_numpy(g['bk']))
        val.weight.copy_(torch.from_numpy(g['wv'])); val.bias.copy_(torch.from_numpy(g['bv']))
    head = vfn.KeyValueHead(key.cuda(), val.cuda()).eval()
    x = torch.from_numpy(g['x']).cuda()
    want_k, want_v = torch.from_numpy(g['key']).cuda(), torch.from_numpy(g['val']).cuda()
    assert not head.training
    for layout in ('auto', 'em', 'dm'):
        with torch.no_grad():
            n0 = head.launches
            k, v = head(x, layout=layout)
            assert head.launches > n0            # the library path ran (train mode would run the two cuDNN convolutions)
        assert k.shape == want_k.shape and v.shape == want_v.shape
        assert float((k - want_k).abs().max()) <= 5e-6 * float(want_k.abs().max())
        assert float((v - want_v).abs().max()) <= 5e-6 * float(want_v.abs().max())


def test_keyvalue_single_pass_is_tf32_class():
    head = _head(passes=1)
    x = _features(1, 1024, 30, 54)
    with torch.no_grad():
        k, v = head(x)
    ke, ve = _exact(head, x)
    ktf, vtf = _cudnn(head, x, tf32=True)
    err = max(float((k.double() - ke).abs().max()), float((v.double() - ve).abs().max()))
    errtf = max(float((ktf.double() - ke).abs().max()), float((vtf.double() - ve).abs().max()))
    _report('480p_pass1', dict(err_vs_exact=err, cudnn_tf32_err_vs_exact=errtf))
    assert err <= 2 * errtf, (err, errtf)


@pytest.mark.parametrize('shape', [(1, 64, 7, 5), (3, 128, 9, 33), (1, 1024, 68, 120), (2, 256, 1, 1), (1, 64, 3, 130)])
def test_keyvalue_shapes_and_borders(shape):
    """ragged sizes: maps narrower / wider than a 32-pixel packing chunk, one pixel, 1080p (68 x 120), three images
    (zero padding at every border and between stacked images is what the padded-raster GEMM must get right)"""
    b, c, h, w = shape
    head = _head(c_in=c, seed=3)
    x = _features(b, c, h, w, seed=4)
    with torch.no_grad():
        k, v = head(x)
    ke, ve = _exact(head, x)
    scale = float(ve.abs().max()) + 1e-30
    err = max(float((k.double() - ke).abs().max()), float((v.double() - ve).abs().max()))
    # the tensor core's fp32 accumulation truncates: ~2^-25 relative per accumulation step, chains capped at 576 steps
    # (KV_CHAIN_MAX); cuDNN's own true-fp32 result sits at 2e-6 .. 2e-5 of the output scale from exact arithmetic
    assert err <= 1.5e-5 * scale, (shape, err, scale)


def test_keyvalue_layouts_agree_bitwise():
    head = _head(seed=5)
    x = _features(2, 1024, 30, 54, seed=6)
    with torch.no_grad():
        k_em, v_em = head(x, layout='em')
        k_dm, v_dm = head(x, layout='dm')
    assert k_dm.is_contiguous() and v_dm.is_contiguous() and not k_em.is_contiguous()
    assert torch.equal(k_em, k_dm) and torch.equal(v_em, v_dm)
    x1 = x[:1].contiguous()
    with torch.no_grad():
        k1, v1 = head(x1)                                 # auto: keys entry-major, values dimension-major
    assert v1.is_contiguous() and not k1.is_contiguous()
    # the K split depends on the batch (work distribution), the values on the summation order only: fp32 rounding apart
    assert float((k1[0] - k_dm[0]).abs().max()) <= 1e-5 * float(k_dm.abs().max())


def test_keyvalue_no_bias_zero_input_and_weight_update():
    head = _head(c_in=64, seed=7, bias=False)
    x = torch.zeros(1, 64, 6, 6, device='cuda')
    with torch.no_grad():
        k, v = head(x)
    assert float(k.abs().max()) == 0 and float(v.abs().max()) == 0
    x = _features(1, 64, 6, 6, seed=8)
    with torch.no_grad():
        k0, _ = head(x)
        head.Key.weight.mul_(2.0)                       # in-place change: the packed operands must follow
        k1, _ = head(x)
    assert torch.allclose(k1, 2 * k0, rtol=1e-6, atol=0)


def test_keyvalue_rejects_what_it_cannot_do():
    import vfloodnet_b200 as vfn
    head = _head(c_in=64)
    with pytest.raises(RuntimeError):
        head(torch.zeros(1, 64, 4, 4))                                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        vfn.KeyValueHead(torch.nn.Conv2d(64, 128, 1), torch.nn.Conv2d(64, 512, 1))
    bad = vfn.KeyValueHead(torch.nn.Conv2d(48, 128, 3, padding=1).cuda(), torch.nn.Conv2d(48, 512, 3, padding=1).cuda()).eval()
    with pytest.raises((ValueError, RuntimeError)):
        bad(torch.zeros(1, 48, 4, 4, device='cuda'))                     # c_in not a multiple of 64


# ---------------------------------------------------------------------------------------------------
# entry-major hand-over to the read and the update
# ---------------------------------------------------------------------------------------------------
def _bank(vfn, n=5000, hw=1620, seed=0):
    g = torch.Generator().manual_seed(seed)
    keys = [torch.randn(128, n, generator=g).cuda() * 0.6 for _ in range(2)]
    vals = [torch.randn(512, n, generator=g).cuda() for _ in range(2)]
    fb = vfn.FeatureBank(2, 250000, torch.device('cuda', 0))
    fb.init_bank(keys, vals)
    return fb


def test_read_and_update_take_entry_major_views():
    import vfloodnet_b200 as vfn
    hw = 1620
    g = torch.Generator().manual_seed(11)
    q_em = (torch.randn(1, hw, 128, generator=g) * 0.6).cuda()
    v_dm = torch.randn(1, 512, hw, generator=g).cuda()
    q_view = q_em.transpose(1, 2)
    m = vfn.Matcher(update_bank=True)
    fb_a, fb_b = _bank(vfn), _bank(vfn)
    out_a = m(fb_a, q_view, v_dm)
    out_b = m(fb_b, q_view.contiguous(), v_dm)
    assert torch.equal(out_a, out_b)
    assert all(torch.equal(fb_a.info[c], fb_b.info[c]) for c in range(2))
    # update: candidates as transposed views of (hw, d) storage vs contiguous (d, hw); some re-occur (merges)
    ck = [(torch.randn(hw, 128, generator=g) * 0.6).cuda() for _ in range(2)]
    cv = [torch.randn(hw, 512, generator=g).cuda() for _ in range(2)]
    for c in range(2):
        ck[c][:200] = fb_a.keys[c][:, :200].t() * 1.01
    fb_a.update([x.t() for x in ck], [x.t() for x in cv], 1)
    fb_b.update([x.t().contiguous() for x in ck], [x.t().contiguous() for x in cv], 1)
    for c in range(2):
        assert fb_a.bank_n(c) == fb_b.bank_n(c)
        assert fb_a.last_decisions[c]['n_merge'] == fb_b.last_decisions[c]['n_merge'] > 0
        assert torch.equal(fb_a.keys[c], fb_b.keys[c]) and torch.equal(fb_a.values[c], fb_b.values[c])
        assert torch.equal(fb_a.info[c], fb_b.info[c])


# ---------------------------------------------------------------------------------------------------
# fuse_model on the reference AFB_URR
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def env():
    from baseline import refshim, model_clip
    if not refshim.available():
        pytest.skip('reference not staged: run `python baseline/make_ref.py` in the build container')
    import vfloodnet_b200 as vfn
    ns = refshim.load()
    dev = torch.device('cuda', 0)
    ref = model_clip.build_reference_model(ns, dev)
    ours = model_clip.patched_copy(ref, vfn)
    fused = vfn.fuse_model(model_clip.patched_copy(ref, vfn))
    folded = vfn.fuse_model(model_clip.patched_copy(ref, vfn), fold_bn=True)
    return dict(ns=ns, vfn=vfn, MC=model_clip, dev=dev, ref=ref, ours=ours, fused=fused, folded=folded)


def test_fuse_model_surface(env):
    """fuse_model keeps every parameter and state_dict key of the reference; only segment and keyval_r4 change"""
    ref, fused, vfn = env['ref'], env['fused'], env['vfn']
    assert isinstance(fused.keyval_r4, vfn.KeyValueHead)
    sr, sf = ref.state_dict(), fused.state_dict()
    assert sr.keys() == sf.keys()
    assert all(torch.equal(sr[k], sf[k]) for k in sr)


def test_fused_segment_and_memorize_vs_patched_model(env):
    """teacher-forced over a clip: from the same bank state, the fused model's memorize output, readout, score and mask
    against the unfused patched model's (true-fp32 convolutions in both: the bar is convolution rounding)"""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    ours, fused = env['ours'], env['fused']
    frames = int(os.environ.get('VFN_FUSED_FRAMES', '10'))
    clip = MC.make_clip(frames)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    stats = dict(k_err=[], v_err=[], prob_err=[], iou=[], pixels_differ=[])
    try:
        with torch.no_grad():
            f0 = clip[0].to(dev)
            m0 = MC.first_mask().to(dev)
            k4, v4 = ours.memorize(f0, m0)
            k4f, v4f = fused.memorize(f0, m0)
            fb = vfn.FeatureBank(2, MC.BUDGET, dev)
            fb.init_bank(k4, v4)
            fbf = vfn.FeatureBank(2, MC.BUDGET, dev)
            for t in range(1, frames + 1):
                frame = clip[t].to(dev)
                fbf.load_state([k.clone() for k in fb.keys], [v.clone() for v in fb.values], [i.clone() for i in fb.info])
                score, _ = ours.segment(frame, fb)
                score_f, _ = fused.segment(frame, fbf)
                assert score.shape == score_f.shape
                pm, pmf = torch.softmax(score, 1), torch.softmax(score_f, 1)
                am, amf = pm[0].argmax(0), pmf[0].argmax(0)
                stats['prob_err'].append(float((pm - pmf).abs().max()))
                stats['iou'].append(MC.iou(am, amf))
                stats['pixels_differ'].append(int((am != amf).sum()))
                k4, v4 = ours.memorize(frame, pm)
                k4f, v4f = fused.memorize(frame, pm)
                ks, vs = float(torch.stack(k4).abs().max()), float(torch.stack(v4).abs().max())
                stats['k_err'].append(max(float((a - b).abs().max()) for a, b in zip(k4, k4f)) / ks)
                stats['v_err'].append(max(float((a - b).abs().max()) for a, b in zip(v4, v4f)) / vs)
                fb.update(k4, v4, t)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    _report('fused_vs_patched', stats)
    # cuDNN's own true-fp32 result at B = 2 sits 1.7e-5 of the output scale from exact arithmetic (keyvalue_report_480p_b2)
    assert max(stats['k_err']) <= 5e-5 and max(stats['v_err']) <= 5e-5, stats
    assert min(stats['iou']) >= 0.9995, stats


def test_fused_model_free_running_bank_trajectory(env):
    """a short free run of the fused model (entry-major candidates into update, entry-major query into the read):
    same bank sizes as the unfused patched model while the masks agree"""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    clip = [f.to(dev) for f in MC.make_clip(6)]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        a = MC.run_clip(env['ours'], vfn.FeatureBank, clip, dev)
        b = MC.run_clip(env['fused'], vfn.FeatureBank, clip, dev)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ious = [MC.iou(x, y) for x, y in zip(a['masks'], b['masks'])]
    _report('fused_free_run', dict(iou=ious, bank_a=[a['fb'].bank_n(c) for c in range(2)],
                                   bank_b=[b['fb'].bank_n(c) for c in range(2)]))
    assert ious[0] >= 0.9995, ious
    assert [a['fb'].bank_n(c) for c in range(2)] == [b['fb'].bank_n(c) for c in range(2)]


def test_graphed_fused_model_equals_eager_fused(env):
    """GraphedAFBURR over the fused model (Refine skip branches inside the encoder graph, KeyValueHead inside the encoder
    and memorize graphs, entry-major views of graph-owned buffers handed to the read and the update) against the eager
    fused model over a free-running clip"""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    clip = [f.to(dev) for f in MC.make_clip(8)]
    scores = {}

    def keep(tag):
        def cb(t, frame, score, pm, k4, v4, fb):
            scores.setdefault(tag, []).append(score.clone())
        return cb

    gm = vfn.GraphedAFBURR(env['fused'], tuple(clip[0].shape))
    assert gm.fused
    e = MC.run_clip(env['fused'], vfn.FeatureBank, clip, dev, on_frame=keep('eager'))
    g = MC.run_clip(gm, vfn.FeatureBank, clip, dev, on_frame=keep('graph'))
    worst = max((a - b).abs().max().item() for a, b in zip(scores['eager'], scores['graph']))
    ious = [MC.iou(a, b) for a, b in zip(e['masks'], g['masks'])]
    _report('graphed_fused_vs_eager', dict(max_score_diff=worst, min_iou=min(ious)))
    assert min(ious) >= 0.9999 and worst <= 1e-3, (min(ious), worst)
    assert [e['fb'].bank_n(c) for c in range(2)] == [g['fb'].bank_n(c) for c in range(2)]


def test_folded_encoders_vs_reference_modules(env):
    """BatchNorm folded into the convolutions, conv + bias + ReLU (+ add) as one cuDNN call.  Same
    function, other rounding: encoder outputs against the reference modules' (true-fp32 convolutions), state_dict
    untouched, and teacher-forced masks of the fully fused model against the unfused patched model."""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    ref, ours, folded = env['ref'], env['ours'], env['folded']
    sr, sf = ref.state_dict(), folded.state_dict()
    assert sr.keys() == sf.keys() and all(torch.equal(sr[k], sf[k]) for k in sr)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    stats = dict(enc=[], iou=[], pixels_differ=[])
    try:
        with torch.no_grad():
            clip = MC.make_clip(6)
            f0, m0 = clip[0].to(dev), MC.first_mask().to(dev)
            fp = torch.nn.functional.pad(f0, (5, 5, 0, 0))
            for a, b in zip(ref.encoder_q(fp), folded.encoder_q(fp)):
                stats['enc'].append(float((a - b).abs().max() / a.abs().max()))
            assert max(stats['enc']) <= 5e-4, stats
            k4, v4 = ours.memorize(f0, m0)
            fb = vfn.FeatureBank(2, MC.BUDGET, dev)
            fb.init_bank(k4, v4)
            fbf = vfn.FeatureBank(2, MC.BUDGET, dev)
            for t in range(1, 7):
                frame = clip[t].to(dev)
                fbf.load_state([k.clone() for k in fb.keys], [v.clone() for v in fb.values], [i.clone() for i in fb.info])
                score, _ = ours.segment(frame, fb)
                score_f, _ = folded.segment(frame, fbf)
                pm, pmf = torch.softmax(score, 1), torch.softmax(score_f, 1)
                am, amf = pm[0].argmax(0), pmf[0].argmax(0)
                stats['iou'].append(MC.iou(am, amf))
                stats['pixels_differ'].append(int((am != amf).sum()))
                k4, v4 = ours.memorize(frame, pm)
                fb.update(k4, v4, t)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    _report('folded_vs_patched', stats)
    assert min(stats['iou']) >= 0.999, stats


def test_pipelined_graphed_model_equals_unpipelined(env):
    """GraphedAFBURR.prefetch: the next frame's encoder stage on a side stream while memorize + update of the current
    frame run.  Same kernels on the same inputs: scores, masks and banks must be IDENTICAL to the un-pipelined graphed
    loop, with device frames and with pinned host frames."""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    host = MC.make_clip(8, pin=True)
    clip = [f.to(dev) for f in host]
    gm = vfn.GraphedAFBURR(env['folded'], tuple(clip[0].shape))
    scores = {}

    def keep(tag):
        def cb(t, frame, score, pm, k4, v4, fb):
            scores.setdefault(tag, []).append(score.clone())
        return cb

    a = MC.run_clip(gm, vfn.FeatureBank, clip, dev, on_frame=keep('plain'))
    b = MC.run_clip(gm, vfn.FeatureBank, clip, dev, on_frame=keep('pipe'), pipeline=True)
    c = MC.run_clip(gm, vfn.FeatureBank, host, dev, on_frame=keep('pipe_host'), pipeline=True, frames_on_host=True)
    torch.cuda.synchronize()
    for tag in ('pipe', 'pipe_host'):
        assert all(torch.equal(x, y) for x, y in zip(scores['plain'], scores[tag])), tag
    for r in (b, c):
        assert all(torch.equal(x, y) for x, y in zip(a['masks'], r['masks']))
        assert [a['fb'].bank_n(k) for k in range(2)] == [r['fb'].bank_n(k) for k in range(2)]
        assert all(torch.equal(a['fb'].keys[k], r['fb'].keys[k]) for k in range(2))


def test_fused_model_other_frame_size_with_padding_on_both_axes(env):
    """360 x 636 frames: pad_divide_by(16) pads both axes (368 x 640); the fused glue must un-pad like the reference
    (AFB_URR.py:312-316) and agree with the unfused patched model"""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    ours, fused = env['ours'], env['folded']
    h, w = 360, 636
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            f0, f1 = MC.make_frame(0, h=h, w=w).to(dev), MC.make_frame(1, h=h, w=w).to(dev)
            m0 = MC.first_mask(h, w).to(dev)
            k4, v4 = ours.memorize(f0, m0)
            assert k4[0].shape == (128, 23 * 40)
            fb, fbf = vfn.FeatureBank(2, MC.BUDGET, dev), vfn.FeatureBank(2, MC.BUDGET, dev)
            fb.init_bank(k4, v4)
            fbf.init_bank(*fused.memorize(f0, m0))
            s, _ = ours.segment(f1, fb)
            sf, _ = fused.segment(f1, fbf)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert s.shape == sf.shape == (1, 2, h, w)
    assert MC.iou(s[0].argmax(0), sf[0].argmax(0)) >= 0.999
    gm = vfn.GraphedAFBURR(fused, (1, 3, h, w))
    with torch.no_grad():
        sg, _ = gm.segment(f1, fbf)
    assert sg.shape == (1, 2, h, w)


def test_prefetch_of_another_frame_is_discarded(env):
    """GraphedAFBURR.prefetch(A) followed by segment(B): the side stream's work is waited for and the encoder stage is
    run again for B - the result is that of a plain segment(B)"""
    vfn, MC, dev = env['vfn'], env['MC'], env['dev']
    clip = [f.to(dev) for f in MC.make_clip(3)]
    gm = vfn.GraphedAFBURR(env['fused'], tuple(clip[0].shape))
    with torch.no_grad():
        fb = vfn.FeatureBank(2, MC.BUDGET, dev)
        fb.init_bank(*gm.memorize(clip[0], MC.first_mask().to(dev)))
        fb.update_bank = False
        gm.global_matcher.update_bank = False            # two reads of the same bank state must not differ by usage counts
        try:
            want, _ = gm.segment(clip[2], fb)
            want = want.clone()
            gm.prefetch(clip[1])
            got, _ = gm.segment(clip[2], fb)
            assert torch.equal(want, got)
            gm.prefetch(clip[2])
            got2, _ = gm.segment(clip[2], fb)
            assert torch.equal(want, got2)
        finally:
            gm.global_matcher.update_bank = True
