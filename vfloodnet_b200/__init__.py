"""vfloodnet_b200: B200-native (sm_100a) AFB-URR memory-propagation hot path of V-FloodNet.

Public surface mirrors the reference operator API for this path:
    FeatureBank  (video_module/model/FeatureBank.py)
    Matcher      (video_module/model/AFB_URR.py:130-178)
    urr_pre / urr_post / decoder_forward / patch_model   (AFB_URR.py:208-239)
    GraphedAFBURR (memorize / segment of AFB_URR.py:255-318 with the convolution stages as CUDA graphs)
    KeyValueHead  (KeyValue, AFB_URR.py:94-111, as a tcgen05 implicit GEMM writing bank / query layout)
    fuse_model    (segment glue without per-object copies, AFB_URR.py:287-297, + KeyValueHead)
    fold_encoders (EncoderM / EncoderQ, AFB_URR.py:33-93, BatchNorm folded, conv+bias+ReLU as one cuDNN call)
Everything computes in libvfn_sm100a.so (include/vfn.h); importing this package without the built library, or
calling it without a CUDA device, fails loudly - there is no CPU fallback.
"""
from .feature_bank import FeatureBank
from .matcher import Matcher
from .urr import urr_pre, urr_post, decoder_forward, patch_model
from .graphed import GraphedAFBURR
from .keyvalue import KeyValueHead
from .glue import fuse_model, segment_fused
from .folded import fold_encoders

__all__ = ['FeatureBank', 'Matcher', 'urr_pre', 'urr_post', 'decoder_forward', 'patch_model', 'GraphedAFBURR',
           'KeyValueHead', 'fuse_model', 'segment_fused', 'fold_encoders']
