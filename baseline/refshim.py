"""Import the UNMODIFIED reference (`video_module`, `myutils`) from baseline/_ref/ (staged by baseline/make_ref.py)
or, in the build container, straight from /root/reference.

Test / bench infrastructure only: tests/, __graft_entry__ and bench.py's reference legs use it; the product package
`vfloodnet_b200` never imports it.  Two import shims, neither touching arithmetic of the reference's own code
(SURVEY.md 8c):

  * `matplotlib` stub - myutils/__init__.py:2 imports plot_depth -> matplotlib.pyplot; import-only, absent here.
  * `torch_scatter`  - rusty1s/pytorch_scatter (pin torch-scatter==2.0.8, reference README.md:58) is neither vendored
    nor installable offline.  `scatter_mean(src, index, dim, out=)` is provided with that version's published
    semantics: out.scatter_add_; count = scatter_add of ones, clamped to >= 1; out.true_divide_(count).
    FeatureBank.py:78,92 are the only call sites.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, '_ref')
_state = {}


def ref_root():
    """directory holding `video_module/` and `myutils/`, or None"""
    for d in (STAGED, os.environ.get('VFN_REFERENCE', '/root/reference')):
        if d and os.path.isdir(os.path.join(d, 'video_module', 'model')) and os.path.isdir(os.path.join(d, 'myutils')):
            return d
    return None


def available():
    return ref_root() is not None


def install_shims():
    import torch
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType('matplotlib')
            plt = types.ModuleType('matplotlib.pyplot')
            mpl.pyplot = plt
            sys.modules['matplotlib'] = mpl
            sys.modules['matplotlib.pyplot'] = plt
    if 'torch_scatter' not in sys.modules:
        try:
            import torch_scatter  # noqa: F401
        except Exception:
            ts = types.ModuleType('torch_scatter')

            def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
                assert out is not None, 'the reference always passes out= (FeatureBank.py:78,92)'
                out.scatter_add_(dim, index, src)
                cnt = torch.zeros_like(out).scatter_add_(dim, index, torch.ones_like(src))
                cnt.clamp_(min=1)
                out.true_divide_(cnt)
                return out

            ts.scatter_mean = scatter_mean
            sys.modules['torch_scatter'] = ts


def load():
    """Returns a namespace with the reference's own classes: AFB_URR, FeatureBank, Matcher, Decoder, myutils."""
    if 'ns' in _state:
        return _state['ns']
    root = ref_root()
    if root is None:
        raise ImportError('the reference is not staged: run `python baseline/make_ref.py` in the build container '
                          '(copies video_module/ and myutils/ into git-ignored baseline/_ref/)')
    install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import myutils
        from video_module.model.AFB_URR import AFB_URR, Matcher, Decoder
        from video_module.model.FeatureBank import FeatureBank
    ns = types.SimpleNamespace(root=root, myutils=myutils, AFB_URR=AFB_URR, Matcher=Matcher, Decoder=Decoder,
                               FeatureBank=FeatureBank)
    _state['ns'] = ns
    return ns
