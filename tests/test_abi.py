"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/vfn.h declares; the ctypes signature table covers the header one to one; host-side constructor logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'vfn.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vfn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    fns = header_functions()
    for must in ['vfn_memread', 'vfn_memread_phase_a', 'vfn_memread_phase_b', 'vfn_bank_match', 'vfn_bank_plan',
                 'vfn_bank_merge', 'vfn_bank_evict_plan', 'vfn_bank_compact', 'vfn_bank_append_rows', 'vfn_urr_pre',
                 'vfn_urr_post', 'vfn_last_error', 'vfn_version']:
        assert must in fns


def test_library_loads_and_exports_every_declared_symbol():
    from vfloodnet_b200 import _lib
    lib = _lib.load()
    assert lib.vfn_version() == 102
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), f'{name} declared in include/vfn.h but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_struct_layout_matches_header():
    from vfloodnet_b200._lib import VfnBank, VfnUpdateIO
    # 2 x int32, 2 x int64, 12 pointers, n_live pointer, n_min int64
    assert ctypes.sizeof(VfnBank) == 8 + 16 + 12 * 8 + 16
    assert VfnBank.keys.offset == 24 and VfnBank.cnt.offset == 24 + 11 * 8
    assert VfnBank.n_live.offset == 24 + 12 * 8 and VfnBank.n_min.offset == 24 + 13 * 8
    # 8 pointers, 8 + 64 int32, 1 int64, 2 int32
    assert ctypes.sizeof(VfnUpdateIO) == 8 * 8 + 72 * 4 + 8 + 8
    assert VfnUpdateIO.thresholds.offset == 8 * 8 + 8 * 4 and VfnUpdateIO.n_before.offset == 8 * 8 + 72 * 4
    assert VfnUpdateIO.deferred.offset == 8 * 8 + 72 * 4 + 8


def test_struct_layout_against_the_c_compiler(tmp_path):
    """sizeof / offsetof as gcc sees include/vfn.h == the ctypes mirror"""
    import os
    import shutil
    import subprocess
    from vfloodnet_b200._lib import VfnBank, VfnUpdateIO
    if shutil.which('gcc') is None:
        import pytest
        pytest.skip('gcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields_b = [f[0] for f in VfnBank._fields_]
    fields_u = [f[0] for f in VfnUpdateIO._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "vfn.h"\nint main(void){\n'
    src += 'printf("%zu\\n", sizeof(vfn_bank));\n'
    src += ''.join(f'printf("%zu\\n", offsetof(vfn_bank, {f}));\n' for f in fields_b)
    src += 'printf("%zu\\n", sizeof(vfn_update_io));\n'
    src += ''.join(f'printf("%zu\\n", offsetof(vfn_update_io, {f}));\n' for f in fields_u)
    src += 'return 0;}\n'
    c = tmp_path / 'layout.c'
    c.write_text(src)
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(root, 'include'), str(c), '-o', str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [ctypes.sizeof(VfnBank)] + [getattr(VfnBank, f).offset for f in fields_b]
    want += [ctypes.sizeof(VfnUpdateIO)] + [getattr(VfnUpdateIO, f).offset for f in fields_u]
    assert vals == want


def test_argument_errors_are_reported_not_fatal():
    from vfloodnet_b200 import _lib
    lib = _lib.load()
    rc = lib.vfn_prep_rows(None, 0, 0, None, None, None, None, 1.0, None)
    assert rc == -1 and b'prep_rows' in lib.vfn_last_error()
    assert lib.vfn_bank_plan_workspace_bytes(1620) == 2048 * 8
    assert lib.vfn_memread_workspace_bytes(2, 100000, 1620, 128, 512) > 0


def test_no_cpu_fallback():
    import torch
    import vfloodnet_b200 as v
    with pytest.raises(ValueError):
        v.FeatureBank(2, 250000, 'cpu')
    from oracle import afb_oracle as O
    m = v.Matcher(update_bank=True)
    with pytest.raises(TypeError):
        m(O.OracleFeatureBank(2, 100), torch.zeros(1, 128, 4), torch.zeros(1, 512, 4))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'vfloodnet_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('no CPU', ''), f'{f} mentions the oracle'
