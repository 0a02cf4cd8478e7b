"""ctypes binding of libvfn_sm100a.so (include/vfn.h).  No CPU fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libvfn_sm100a.so')

c_i32, c_i64, c_f32, c_f64, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_size_t


class VfnBank(C.Structure):
    """struct vfn_bank (include/vfn.h)"""
    _fields_ = [('d_key', c_i32), ('d_val', c_i32), ('cap', c_i64), ('n', c_i64),
                ('keys', c_vp), ('values', c_vp), ('info', c_vp), ('nk', c_vp), ('nkh', c_vp), ('nkl', c_vp),
                ('kh', c_vp), ('kl', c_vp), ('vh', c_vp), ('v8', c_vp), ('vl', c_vp), ('cnt', c_vp),
                ('n_live', c_vp), ('n_min', c_i64)]


class VfnUpdateIO(C.Structure):
    """struct vfn_update_io (include/vfn.h)"""
    _fields_ = [('d_prev_key_dm', c_vp), ('d_prev_value_dm', c_vp), ('d_match_idx', c_vp), ('d_match_corr', c_vp),
                ('d_merge_q', c_vp), ('d_merge_slot', c_vp), ('d_run_off', c_vp), ('d_append_q', c_vp),
                ('n_merge', c_i32), ('n_runs', c_i32), ('n_append', c_i32), ('evicted', c_i32), ('swapped', c_i32),
                ('evict_status', c_i32), ('kept', c_i32), ('n_iter', c_i32), ('thresholds', c_i32 * 64),
                ('n_before', c_i64), ('deferred', c_i32), ('prev_layout', c_i32)]


BANK_P = C.POINTER(VfnBank)
IO_P = C.POINTER(VfnUpdateIO)

# name -> (restype, argtypes); mirrors include/vfn.h one to one
SIGNATURES = {
    'vfn_version': (c_i32, []),
    'vfn_last_error': (C.c_char_p, []),
    'vfn_device_is_sm100': (c_i32, []),
    'vfn_abi_sizeof_bank': (c_i32, []),
    'vfn_abi_sizeof_update_io': (c_i32, []),
    'vfn_prep_rows': (c_i32, [c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp]),
    'vfn_bank_append_rows': (c_i32, [BANK_P, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_f32, c_f32, c_vp]),
    'vfn_bank_set_live': (c_i32, [BANK_P, c_i64, c_vp]),
    'vfn_bank_refresh': (c_i32, [BANK_P, c_i64, c_i64, c_vp]),
    'vfn_memread_workspace_bytes': (c_sz, [c_i32, c_i64, c_i64, c_i32, c_i32]),
    'vfn_memread': (c_i32, [BANK_P, c_i32, c_vp, c_vp, c_i64, c_f32, c_i32, c_vp, c_vp, c_vp, c_sz, c_i32, c_vp]),
    'vfn_memread_phase_a': (c_i32, [BANK_P, c_i32, c_vp, c_i64, c_vp, c_vp, c_sz, c_i32, c_vp]),
    'vfn_memread_phase_b': (c_i32, [BANK_P, c_i32, c_vp, c_i64, c_vp, c_f32, c_i32, c_vp, c_vp, c_sz, c_i32, c_vp]),
    'vfn_lse_combine': (c_i32, [c_vp, c_i32, c_i64, c_vp, c_vp]),
    'vfn_lse_combine_peers': (c_i32, [C.POINTER(c_vp), c_i32, c_i64, c_vp, c_vp]),
    'vfn_reduce_peers': (c_i32, [C.POINTER(c_vp), c_i32, c_i64, c_i64, c_vp, c_vp]),
    'vfn_gather_peers': (c_i32, [C.POINTER(c_vp), c_i32, c_i32, c_i64, c_i64, c_vp, c_vp]),
    'vfn_match_pack': (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    'vfn_match_combine_peers': (c_i32, [C.POINTER(c_vp), c_i32, c_i64, c_vp, c_vp, c_vp]),
    'vfn_bank_match_workspace_bytes': (c_sz, [c_i64, c_i64]),
    'vfn_bank_match': (c_i32, [BANK_P, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_i32, c_vp]),
    'vfn_bank_plan_workspace_bytes': (c_sz, [c_i64]),
    'vfn_bank_plan': (c_i32, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'vfn_bank_merge': (c_i32, [BANK_P, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_vp]),
    'vfn_bank_evict_plan': (c_i32, [BANK_P, c_f32, c_f64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'vfn_bank_compact_workspace_bytes': (c_sz, [c_i64]),
    'vfn_bank_compact': (c_i32, [BANK_P, BANK_P, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'vfn_bank_clamp_info': (c_i32, [BANK_P, c_i64, c_vp]),
    'vfn_bank_update_workspace_bytes': (c_sz, [c_i32, c_i64, c_i64, c_i32, c_i32]),
    'vfn_bank_update': (c_i32, [BANK_P, BANK_P, c_i32, IO_P, c_i64, c_f32, c_f32, c_f32, c_f64, c_vp, c_sz, c_vp,
                                c_i32, c_vp, c_vp]),
    'vfn_bank_update_finish': (c_i32, [BANK_P, c_i32, IO_P, c_vp]),
    'vfn_kv_packed_weights_bytes': (c_sz, [c_i32, c_i32, c_i32]),
    'vfn_kv_pack_weights': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'vfn_keyvalue_workspace_bytes': (c_sz, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32]),
    'vfn_keyvalue': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                             c_vp, c_sz, c_vp]),
    'vfn_urr_pre': (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vfn_urr_post': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'vfn_tail_workspace_bytes': (c_sz, [c_i32, c_i32]),
    'vfn_tail_resize_argmax': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'vfn_tail_largest_component': (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'vfn_tail_waterlevel': (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    'vfn_frame_tail': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp,
                               c_vp, c_vp, c_sz, c_vp]),
    'vfn_profile_enable': (c_i32, [c_i32]),
    'vfn_profile_collect': (c_i32, [c_vp, c_i32]),
    'vfn_profile_add_work': (c_i32, [c_i32, c_f64]),
    'vfn_launch_count': (c_i64, []),
    'vfn_debug_set_dump': (c_i32, [c_vp]),
    'vfn_debug_set_tstamp': (c_i32, [c_vp]),
    'vfn_debug_set_pair': (c_i32, [c_i32]),
    'vfn_debug_set_urr_stream': (c_i32, [c_i32]),
    'vfn_debug_set_tail': (c_i32, [c_i32]),
    'vfn_debug_set_pdl': (c_i32, [c_i32]),
}

_lib = None
VFN_VERSION = 102      # include/vfn.h
VFN_Q_IN_EM = 0x100    # vfn_memread impl bit 8: q_in is entry-major


class VfnError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """dlopen the in-tree library.  Raises ImportError loudly if it is absent and cannot be built.
    With nvcc at hand a library older than its sources (csrc/*, include/vfn.h) is rebuilt first (build.build() is a no-op
    when it is current), so a stale .so with other struct layouts is never loaded silently; without nvcc (the GPU box
    receives the prebuilt library) the ABI version export is checked instead."""
    global _lib
    if _lib is not None:
        return _lib
    have_nvcc = bool(shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc'))
    if build_if_missing and have_nvcc:
        from . import build as _build
        if _build.needs_build():
            _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing and nvcc is not available: the CUDA extension must be built '
                          f'(python -m vfloodnet_b200.build); there is no CPU fallback')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.vfn_version() != VFN_VERSION or lib.vfn_abi_sizeof_bank() != C.sizeof(VfnBank) or \
            lib.vfn_abi_sizeof_update_io() != C.sizeof(VfnUpdateIO):
        raise ImportError(f'{LIB_PATH} does not match this binding (version {lib.vfn_version()} vs {VFN_VERSION}, '
                          f'struct sizes {lib.vfn_abi_sizeof_bank()}/{lib.vfn_abi_sizeof_update_io()} vs '
                          f'{C.sizeof(VfnBank)}/{C.sizeof(VfnUpdateIO)}): rebuild with python -m vfloodnet_b200.build')
    if os.environ.get('VFN_PAIR'):          # debug override of the tensor-kernel selection mask (vfn_debug_set_pair)
        lib.vfn_debug_set_pair(int(os.environ['VFN_PAIR']))
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().vfn_last_error().decode('utf-8', 'replace')
        raise VfnError(f'{what or "libvfn"} failed with code {rc}: {msg}')


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def em_backed(t) -> bool:
    """True for a (d, n) fp32 tensor that is the transposed VIEW of contiguous (n, d) entry-major storage - what
    vfloodnet_b200.KeyValueHead hands out in place of the reference's (d, n) KeyValue outputs.  The library then reads
    the rows as they lie (VFN_Q_IN_EM / vfn_update_io.prev_layout) instead of transposing a contiguous copy."""
    import torch
    return (t.dim() == 2 and t.dtype == torch.float32 and t.shape[0] > 1 and t.shape[1] > 1 and
            t.stride(0) == 1 and t.stride(1) == t.shape[0])


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (None: the current device)"""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def on_device(device):
    """context: make `device` the current CUDA device for the library calls inside (the library launches on the calling
    thread's current device; banks, matchers and tails may live on any GPU of the process, like the reference's)"""
    import torch
    return torch.cuda.device(device)
