"""Where the convolution time of a frame goes: per-module CUDA-event times of the reference
AFB_URR's stages at 480p, with TF32 convolutions (the reference default) and with true fp32, plus the KeyValue head
alone (3x3, 1024 -> 128 / 512 on the (B, 1024, 30, 54) r4 map) at B = 1 (segment) and B = 2 (memorize).
Diagnostic script (run on the GPU box), not a test:  python tests/debug_cnn_times.py > gpurun_out/<tag>/cnn_times.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import model_clip, refshim  # noqa: E402


def time_fn(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device('cuda:0')
    ns = refshim.load()
    model = model_clip.build_reference_model(ns, dev)
    out = {}
    f = model_clip.make_frame(1).to(dev)
    m = model_clip.first_mask().to(dev)
    [fp], _ = ns.myutils.pad_divide_by([f], 16, f.shape[-2:])
    (fm, mm), _ = ns.myutils.pad_divide_by([f, m], 16, f.shape[-2:])
    fm2 = fm.expand(2, -1, -1, -1)
    mk = mm[0].unsqueeze(1).float()
    mk_inv = (1 - mk).clamp(0, 1)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        tag = 'tf32' if tf32 else 'fp32'
        with torch.no_grad():
            r4, r3, r2, r1 = model.encoder_q(fp)
            r4m, r1m = model.encoder_m(fm2, mk, mk_inv)
            res = {}
            res['encoder_q_b1'] = time_fn(lambda: model.encoder_q(fp))
            res['encoder_m_b2'] = time_fn(lambda: model.encoder_m(fm2, mk, mk_inv))
            res['keyval_b1'] = time_fn(lambda: model.keyval_r4(r4))
            res['keyval_b2'] = time_fn(lambda: model.keyval_r4(r4m))
            res['keyval_key_only_b1'] = time_fn(lambda: model.keyval_r4.Key(r4))
            res['keyval_value_only_b1'] = time_fn(lambda: model.keyval_r4.Value(r4))
            fb = ns.FeatureBank(2, model_clip.BUDGET, dev)
            k4, v4 = model.memorize(f, m)
            fb.init_bank(k4, v4)
            res['segment_total_small_bank'] = time_fn(lambda: model.segment(f, fb), iters=10, warm=3)
            res['memorize_total'] = time_fn(lambda: model.memorize(f, m), iters=10, warm=3)
            # decoder alone on the tensors segment() would hand it
            kq, vq = model.keyval_r4(r4)
            rg = model.global_matcher(fb, kq, vq).reshape(2, 1024, *r4.shape[-2:])
            r3e = r3.expand(2, -1, -1, -1)
            r2e = r2.expand(2, -1, -1, -1)
            r1e = r1.expand(2, -1, -1, -1)
            fs = (1, 2, r1.shape[2], r1.shape[3])
            res['decoder_b2'] = time_fn(lambda: model.decoder(rg, r3e, r2e, r1e, fs), iters=10, warm=3)
            import vfloodnet_b200 as vfn
            from vfloodnet_b200 import glue
            for passes in (3, 1):
                head = vfn.KeyValueHead.from_reference(model.keyval_r4, passes=passes)
                res[f'vfn_keyvalue_b1_p{passes}'] = time_fn(lambda: head(r4))
                res[f'vfn_keyvalue_b2_p{passes}'] = time_fn(lambda: head(r4m))
            dec = model.decoder
            res['refine_skips_b1'] = time_fn(lambda: (glue.refine_skip(dec.RF3, r3), glue.refine_skip(dec.RF2, r2)),
                                             iters=10, warm=3)
            s3, s2 = glue.refine_skip(dec.RF3, r3), glue.refine_skip(dec.RF2, r2)
            res['decoder_trunk_shared_b2'] = time_fn(lambda: glue.decoder_trunk_shared(dec, rg, s3, s2), iters=10, warm=3)
            res['decoder_trunk_reference_b2'] = time_fn(
                lambda: dec.pred2(torch.relu(dec.RF2(r2e, dec.RF3(r3e, dec.ResMM(dec.convFM(rg)))))), iters=10, warm=3)
            import copy
            from vfloodnet_b200 import folded
            mf = folded.fold_encoders(copy.deepcopy(model))
            res['folded_encoder_q_b1'] = time_fn(lambda: mf.encoder_q(fp))
            res['folded_encoder_m_b2'] = time_fn(lambda: mf.encoder_m(fm2, mk, mk_inv))
            a4 = mf.encoder_q(fp)[0]
            res['folded_r4_rel_diff'] = float((a4 - r4).abs().max() / r4.abs().max())
            del mf
        out[tag] = {k: round(v, 6) for k, v in res.items()}
    out['r4_absmax'] = float(r4.abs().max())
    out['r4m_absmax'] = float(r4m.abs().max())
    out['key_weight_absmax'] = float(model.keyval_r4.Key.weight.abs().max())
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
