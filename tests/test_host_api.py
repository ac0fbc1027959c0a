"""CPU-side checks of the host API: the C-ABI library loads and exports every symbol of
include/dsb200.h, constructor validation mirrors the reference, optimizers / schedules."""
import ctypes
import os

import numpy as np
import pytest

from helpers import healpix_L


def test_library_loads_and_exports_every_declared_symbol():
    from deepsphere_tf1_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    dll = ctypes.CDLL(_lib.LIB_PATH)
    assert len(_lib.PROTOS) >= 30
    for name in _lib.PROTOS:
        assert hasattr(dll, name), name
    assert _lib.LIB.load().dsb200_abi_version() == 2


def test_product_has_no_cpu_fallback():
    import torch
    from deepsphere_tf1_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.fill(torch.zeros(4), 1.0)


def test_exponential_decay_end_points_of_the_cosmo_demo():
    """experiments/cosmo/demo.ipynb:847 prints 2.0e-4 -> 4.7e-5 over 1440 steps for
    exponential_decay(2e-4, step, decay_steps=1, decay_rate=0.999)."""
    from deepsphere_tf1_b200 import optimizers
    assert optimizers.exponential_decay(2e-4, 0, 1, 0.999) == pytest.approx(2.0e-4)
    assert optimizers.exponential_decay(2e-4, 1440, 1, 0.999) == pytest.approx(4.7e-5, rel=2e-2)


def test_constructor_errors_match_the_reference():
    from deepsphere_tf1_b200 import models
    L = [healpix_L(2), healpix_L(1)]
    common = dict(num_epochs=1, scheduler=lambda s: 1e-3, optimizer=None, verbose=False)
    with pytest.raises(ValueError, match='same length'):
        models.cgcnn(L, F=[4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], **common)
    with pytest.raises(ValueError, match='greater or equal'):
        models.cgcnn(L, F=[4, 4], K=[3, 3], p=[0, 1], batch_norm=[True, True], M=[], **common)
    with pytest.raises(ValueError, match='powers of two'):
        models.cgcnn(L, F=[4, 4], K=[3, 3], p=[3, 1], batch_norm=[True, True], M=[], **common)
    with pytest.raises(ValueError, match='last'):
        models.cgcnn(L, F=[4, 4], K=[3, 3], p=[4, 4], batch_norm=[True, True], M=[], **common)
    with pytest.raises(ValueError, match='Unknown statistical'):
        models.cgcnn(L, F=[4, 4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='median', **common)


def test_integration_stub_matches_the_header():
    """The ctypes stub printed in INTEGRATION.md declares as many arguments as include/dsb200.h does (an argument short
    shifts the stream handle into the workspace slot)."""
    import re
    from deepsphere_tf1_b200 import _lib
    text = open(os.path.join(os.path.dirname(_lib.HEADER), '..', 'INTEGRATION.md')).read()
    found = re.findall(r'lib\.(dsb200_\w+)\.argtypes = \[([^\]]*)\]', text)
    assert len(found) >= 2
    for name, args in found:
        n = len([a for a in args.split(',') if a.strip()])
        assert n == len(_lib.PROTOS[name][1]), (name, n, len(_lib.PROTOS[name][1]))
    for name, args in re.findall(r'lib\.(dsb200_\w+)\(([^)]*)\)', text):
        if name in _lib.PROTOS and name != 'dsb200_last_error':
            n = len([a for a in args.split(',') if a.strip()])
            assert n == len(_lib.PROTOS[name][1]), ('call of ' + name, n, len(_lib.PROTOS[name][1]))


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """include/dsb200.h is the contract for bindings from other host languages: it must compile as strict C99 (no C++, no
    torch types) and every prototype must resolve against libdsb200.so at link time.  The program takes the address of every
    declared entry point and calls the two that need no device."""
    import shutil
    import subprocess
    from deepsphere_tf1_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('no C compiler')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = sorted(_lib.PROTOS)
    src = ['#include <stdio.h>', '#include "dsb200.h"', 'int main(void) {', '    const void* entry[] = {']
    src += ['        (const void*)&{},'.format(n) for n in names]
    src += ['    };', '    unsigned n = 0, i;', '    for (i = 0; i < sizeof(entry) / sizeof(entry[0]); ++i) n += entry[i] != 0;',
            '    printf("%d %u %s\\n", dsb200_abi_version(), n, dsb200_last_error()[0] ? "error" : "clean");', '    return 0;', '}']
    c = tmp_path / 'abi.c'
    c.write_text('\n'.join(src) + '\n')
    exe = str(tmp_path / 'abi')
    libdir = os.path.dirname(_lib.LIB_PATH)
    # -pedantic would object to the function -> object pointer casts of the table above, not to the header
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(root, 'include'), str(c), '-o', exe,
                        '-L', libdir, '-ldsb200', '-Wl,-rpath,' + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ['2', str(len(names)), 'clean']
    # and the header alone, strictly
    h = tmp_path / 'only_header.c'
    h.write_text('#include "dsb200.h"\nint dsb200_header_compiles;\n')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(root, 'include'), '-c', str(h),
                        '-o', str(tmp_path / 'only_header.o')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
