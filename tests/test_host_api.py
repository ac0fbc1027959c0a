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
