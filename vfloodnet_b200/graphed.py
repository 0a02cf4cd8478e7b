"""The convolutional stages around the hot path as CUDA graphs ("CUDA graphs for the whole frame
step"; reference: AFB_URR.memorize / AFB_URR.segment, video_module/model/AFB_URR.py:255-318).

Once read + URR + update take ~1 ms per frame, a frame of the reference loop is dominated by its ~350 eager cuDNN / ATen
launches (encoder_q, KeyValue, decoder, encoder_m for every object): 6.8 of 7.6 ms on a B200, most of it launch-bound.
`GraphedAFBURR` wraps a reference `AFB_URR` instance (after `patch_model`) and captures its four static-shape stages once:

    enc   frame -> pad -> encoder_q -> KeyValue             (AFB_URR.py:279-285)
    dec   readout, r3, r2 -> convFM .. pred2 -> coarse logits (AFB_URR.py:209-212; the expand of r3 / r2 included)
    loc   [r1 ; r1_local] -> local_convFM, local_ResMM, local_pred2   (AFB_URR.py:232-233)
    mem   frame, mask -> pad -> encoder_m -> KeyValue        (AFB_URR.py:257-272), captured by calling model.memorize

Between the graphs the library's kernels run as before (vfn_memread, vfn_urr_pre / _post, vfn_bank_update): their shapes
follow the bank.  The modules, weights and the arithmetic are the reference's own; only the launch mechanism changes, so
`segment` / `memorize` return what the patched model returns (tests/test_gpu_dropin.py::test_graphed_model_equals_eager).
Same method names and signatures as the reference model: the frame loop of test_video_seg.py:99-112 runs unchanged.
"""
from __future__ import annotations

import torch
from torch.nn import functional as NF

from .glue import decoder_trunk_shared, refine_skip
from .urr import urr_post, urr_pre


def _pad16(x):
    """myutils.pad_divide_by([x], 16, x.shape[-2:]) (myutils/data.py:134-151) for one tensor"""
    h, w = x.shape[-2:]
    nh, nw = (h + 15) // 16 * 16, (w + 15) // 16 * 16
    lh, lw = int((nh - h) / 2), int((nw - w) / 2)
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return NF.pad(x, pad), pad


class GraphedAFBURR:
    def __init__(self, model, frame_shape, obj_n: int = 2, warmup: int = 3, fused=None):
        """model: a reference AFB_URR (eval mode, on a CUDA device) with vfloodnet_b200.patch_model applied.
        frame_shape: (1, 3, H, W) of the frames the loop will feed (test_video_seg.py:107 after the resize).
        fused: capture the copy-free glue of vfloodnet_b200.glue (Refine skip branches once per frame, inside the
        encoder graph); default: whatever the model runs eagerly (fuse_model applied or not)."""
        self.model, self.obj_n = model, obj_n
        self.fused = ('_vfn_ref_segment' in model.__dict__) if fused is None else bool(fused)
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('GraphedAFBURR needs the model on a CUDA device')
        self.device = dev
        b, c, h, w = frame_shape
        if b != 1:
            raise ValueError('inference path: one frame at a time (bs == 1)')
        self.frame = torch.zeros(frame_shape, device=dev)          # input of the memorize graph
        self.frame_q = torch.zeros(frame_shape, device=dev)        # input of the encoder graph (the query frame)
        self._side = torch.cuda.Stream(dev)                        # prefetch(): the next frame's encoder stage
        self._enc_done = None
        self._prefetched = None                                    # the object handed to prefetch(), until it is consumed
        self._seg_src = None                                       # the object handed to the last segment()
        self.mask = torch.zeros((1, obj_n, h, w), device=dev)
        self.mask[:, 0] = 1
        with torch.no_grad():
            self._capture(warmup)

    # ---- the four stages as plain functions of the static buffers ---------------------------------
    def _enc(self):
        f, pad = _pad16(self.frame_q)
        r4, r3, r2, r1 = self.model.encoder_q(f)
        k4, v4 = self.model.keyval_r4(r4)
        if self.fused:      # the object-independent halves of the two Refine blocks (AFB_URR.py:121) belong to the frame
            d = self.model.decoder
            r3, r2 = refine_skip(d.RF3, r3), refine_skip(d.RF2, r2)
        return k4, v4, r3, r2, r1, pad, tuple(r4.shape[-2:])

    def _dec(self):
        d, n = self.model.decoder, self.obj_n
        if self.fused:
            return decoder_trunk_shared(d, self.res_global, self.r3, self.r2)
        r3 = self.r3.unsqueeze(1).expand(-1, n, -1, -1, -1).reshape(n, *self.r3.shape[1:])      # AFB_URR.py:291-292
        r2 = self.r2.unsqueeze(1).expand(-1, n, -1, -1, -1).reshape(n, *self.r2.shape[1:])
        p = d.ResMM(d.convFM(self.res_global))
        p = d.RF3(r3, p)
        p = d.RF2(r2, p)
        return d.pred2(NF.relu(p))

    def _loc(self):
        d = self.model.decoder
        q = d.local_ResMM(d.local_convFM(self.local_match))
        return d.local_pred2(NF.relu(q))

    def _mem(self):
        # lists of per-object (d, HW) tensors, as the model returns them: views of the graph's own output buffers (with
        # KeyValueHead: transposed views of entry-major storage, which FeatureBank.update reads as it lies)
        return self.model.memorize(self.frame, self.mask)

    def _graph(self, fn, warmup):
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        return g, out

    def _capture(self, warmup):
        with torch.cuda.device(self.device):
            self.g_enc, (self.k4, self.v4, self.r3, self.r2, self.r1, self.pad, self.grid4) = \
                self._graph(self._enc, warmup)       # fused: r3 / r2 hold the Refine skip branches; grid4 = r4 grid
            n = self.obj_n
            self.res_global = torch.zeros((n, 2 * self.v4.shape[1]) + self.grid4, device=self.device)
            self.g_dec, self.p = self._graph(self._dec, warmup)
            c1, h1, w1 = self.r1.shape[1:]
            self.local_match = torch.zeros((n, 2 * c1, h1, w1), device=self.device)
            self.g_loc, self.q = self._graph(self._loc, warmup)
            self.g_mem, (self.mk4, self.mv4) = self._graph(self._mem, warmup)

    # ---- reference method surface ------------------------------------------------------------------
    @property
    def global_matcher(self):
        return self.model.global_matcher

    @property
    def decoder(self):
        return self.model.decoder

    def eval(self):
        return self

    @torch.no_grad()
    def memorize(self, frame, mask):
        """AFB_URR.memorize (AFB_URR.py:255-272): lists of (128, HW) / (512, HW) per object (views of static buffers,
        valid until the next memorize)"""
        if frame is not self._seg_src:           # segment() left a copy of the frame it was given in self.frame
            self.frame.copy_(frame, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self.g_mem.replay()
        return list(self.mk4), list(self.mv4)

    @torch.no_grad()
    def prefetch(self, frame):
        """Frame-level software pipelining (no reference counterpart; optional).  Call it right after `segment(t)` returned
        with frame t+1 (a device tensor or a pinned host tensor): the part of `segment` that depends on the frame only -
        encoder_q, KeyValue, the Refine skip branches: the whole encoder graph - is issued on a side stream, where it
        overlaps `memorize(t)` and `fb.update(t)` of the current frame (convolutions at batch 1-2 leave SMs idle).  The next
        `segment` must be given the SAME object; it then only waits for the side stream.  Results are those of the
        un-pipelined loop (same kernels, same inputs)."""
        cur = torch.cuda.current_stream(self.device)
        if self._prefetched is not None:
            cur.wait_event(self._enc_done)
        self._side.wait_stream(cur)          # every reader of the encoder buffers issued so far (read, decoder, URR of frame t)
        with torch.cuda.stream(self._side):
            self.frame_q.copy_(frame, non_blocking=True)
            self.g_enc.replay()
            self._enc_done = torch.cuda.Event()
            self._enc_done.record(self._side)
        self._prefetched = frame

    @torch.no_grad()
    def segment(self, frame, fb_global):
        """AFB_URR.segment, inference branch (AFB_URR.py:274-318): returns (score, None)"""
        n = fb_global.obj_n
        if n != self.obj_n:
            raise ValueError('object count differs from the captured graphs')
        cur = torch.cuda.current_stream(self.device)
        if self._prefetched is not None:
            cur.wait_event(self._enc_done)       # the side stream is done with the encoder buffers (whatever it ran)
        if self._prefetched is not frame:
            self.frame_q.copy_(frame, non_blocking=True)
            self.g_enc.replay()
        self._prefetched = None
        self.frame.copy_(self.frame_q, non_blocking=True)    # memorize(frame, .) of this frame finds it there
        self._seg_src = frame
        res = self.model.global_matcher(fb_global, self.k4, self.v4)                   # (1, n, 1024, HW): the read
        self.res_global.copy_(res.reshape(n, -1, *self.grid4))
        self.g_dec.replay()
        r1 = self.r1.expand(n, -1, -1, -1)                                             # stride-0 view: never materialised
        fs = (1, n, self.r1.shape[2], self.r1.shape[3])
        p_up, unc, conf, _lm = urr_pre(self.p, r1, fs, out_local_match=self.local_match)
        self.g_loc.replay()
        prob = urr_post(p_up, unc, conf, self.q)                                       # (n, H, W)
        score = prob.view(1, n, *prob.shape[-2:])
        score = torch.clamp(score, 1e-7, 1 - 1e-7)                                     # AFB_URR.py:308-309
        score = torch.log(score / (1 - score))
        pad = self.pad
        if pad[2] + pad[3] > 0:
            score = score[:, :, pad[2]:score.shape[2] - pad[3], :]
        if pad[0] + pad[1] > 0:
            score = score[:, :, :, pad[0]:score.shape[3] - pad[1]]
        return score, None
